// SURVEY §8f rank 1 ("next" row): the 6-layer multi-scale deformable-attention encoder of mmdet's
// MSDeformAttnPixelDecoder (cfg configs/mask2former/pairnet.py:38-66, call site pairnet_head.py:262), the
// upstream that produces the memories the hot path consumes.
//
//   per layer:  q = x + pos
//               value = value_proj(x) ; [offsets | logits] = [sampling_offsets ; attention_weights](q)
//               out   = sum_{level,point} softmax(logits) * bilinear(value_level, ref + offsets/(W,H))   (gather-bound)
//               x     = LN(x + output_proj(out)) ; x = LN(x + FFN(x))
//
// Linears (M = B * 21 950 tokens) run on the tcgen05 3xTF32 GEMM (umma_gemm.cu); producers emit their
// outputs already split hi/lo.  The sampling kernel maps one warp to one (token, head): lane = channel, so every
// bilinear tap is one coalesced 128-byte line served from L2 (the 45 MB value tensor is L2 resident).
#include "common.cuh"

namespace pn {

constexpr int MSDA_MAX_LEVELS = 4;
struct MsdaGeom {
  int L, P;                       // levels, points
  int h[MSDA_MAX_LEVELS], w[MSDA_MAX_LEVELS], start[MSDA_MAX_LEVELS];
  int nq;                         // tokens per image
};

__device__ __forceinline__ float rna_tf32m(float v) {  // == cvt.rna.tf32.f32 for finite inputs, 2 integer ops (umma_ptx.cuh)
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}

// value [B,nq,256] ; ol [B*nq, ldo]: cols [0, 8*L*P*2) offsets ((h,l,p),xy), then 8*L*P attention logits ((h),(l,p))
// out (hi, lo) [B*nq,256]
// The kernel is instruction-issue bound, not bandwidth bound (the value tensor lives in L2), so the mapping
// minimises instructions per gathered byte: 8 lanes own one (token, head) and each lane gathers a float4
// (4 of the 32 head channels); a warp covers 4 consecutive tokens.  Lane j of a group prepares sampling points j
// and j+8 (softmax weight x bilinear weights, clamped tap addresses) once; the tap loop only broadcasts them
// inside the 8-lane group and issues 4 x LDG.128 + 16 FFMA per point.
constexpr int MSDA_TOK = 32;  // tokens per CTA (8 warps x 4)
struct MsdaPoint {
  float w00, w01, w10, w11;  // attention weight x bilinear weight (0 for taps outside the map)
  int r0, r1, x0, x1;        // clamped row offsets (start + y*W) and columns
};
__device__ __forceinline__ MsdaPoint msda_point(float aw, float ref_x, float ref_y, float ox, float oy, int W, int H,
                                                int start) {
  // pixel coordinates in the sampled level: (ref + off/(W,H)) * (W,H) - 0.5   (grid_sample, align_corners=False)
  const float x = (ref_x + ox / (float)W) * (float)W - 0.5f;
  const float y = (ref_y + oy / (float)H) * (float)H - 0.5f;
  const float xf = floorf(x), yf = floorf(y);
  const int x0 = (int)xf, y0 = (int)yf;
  const float lx = x - xf, ly = y - yf;
  const bool xin0 = x0 >= 0 && x0 < W, xin1 = x0 + 1 >= 0 && x0 + 1 < W;
  const bool yin0 = y0 >= 0 && y0 < H, yin1 = y0 + 1 >= 0 && y0 + 1 < H;
  MsdaPoint p;
  p.w00 = (xin0 && yin0) ? aw * (1.f - lx) * (1.f - ly) : 0.f;
  p.w01 = (xin1 && yin0) ? aw * lx * (1.f - ly) : 0.f;
  p.w10 = (xin0 && yin1) ? aw * (1.f - lx) * ly : 0.f;
  p.w11 = (xin1 && yin1) ? aw * lx * ly : 0.f;
  p.x0 = min(max(x0, 0), W - 1);
  p.x1 = min(max(x0 + 1, 0), W - 1);
  p.r0 = start + min(max(y0, 0), H - 1) * W;
  p.r1 = start + min(max(y0 + 1, 0), H - 1) * W;
  return p;
}
__device__ __forceinline__ float seg8_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
}
__device__ __forceinline__ float seg8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v + __shfl_xor_sync(0xffffffffu, v, 1);
}

__global__ void __launch_bounds__(256) msda_sample_kernel(const float* __restrict__ value, const float* __restrict__ ol,
                                                           int ldo, float* __restrict__ out_hi, float* __restrict__ out_lo,
                                                           const MsdaGeom g, int B) {
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const int head = blockIdx.y, b = blockIdx.z;
  int q = blockIdx.x * MSDA_TOK + (threadIdx.x >> 3);
  const bool qvalid = q < g.nq;
  q = qvalid ? q : g.nq - 1;  // keep the whole warp converged for the shuffles; stores are predicated
  const long long tok = (long long)b * g.nq + q;
  const int LP = g.L * g.P;   // <= 16
  int ql = 0;
  while (ql + 1 < g.L && q >= g.start[ql + 1]) ++ql;
  const int qi = q - g.start[ql];
  const float ref_x = ((float)(qi % g.w[ql]) + 0.5f) / (float)g.w[ql];
  const float ref_y = ((float)(qi / g.w[ql]) + 0.5f) / (float)g.h[ql];

  const float* row = ol + (size_t)tok * ldo;
  const int pa = sub, pb = sub + 8;  // the two sampling points this lane prepares
  float la = -INFINITY, lb = -INFINITY;
  float2 oa = make_float2(0.f, 0.f), ob = oa;
  if (pa < LP) {
    la = __ldg(row + NH * LP * 2 + head * LP + pa);
    oa = __ldg(reinterpret_cast<const float2*>(row + (head * LP + pa) * 2));
  }
  if (pb < LP) {
    lb = __ldg(row + NH * LP * 2 + head * LP + pb);
    ob = __ldg(reinterpret_cast<const float2*>(row + (head * LP + pb) * 2));
  }
  const float mx = seg8_max(fmaxf(la, lb));
  const float ea = (pa < LP) ? __expf(la - mx) : 0.f, eb = (pb < LP) ? __expf(lb - mx) : 0.f;
  const float inv = 1.f / seg8_sum(ea + eb);
  MsdaPoint A, Bp;
  {
    const int lvl = min(pa / g.P, g.L - 1);
    A = msda_point(ea * inv, ref_x, ref_y, oa.x, oa.y, g.w[lvl], g.h[lvl], g.start[lvl]);
    const int lvb = min(pb / g.P, g.L - 1);
    Bp = msda_point(eb * inv, ref_x, ref_y, ob.x, ob.y, g.w[lvb], g.h[lvb], g.start[lvb]);
  }
  const float4* vbase = reinterpret_cast<const float4*>(value + (size_t)b * g.nq * D + head * HD) + sub;
  // The prepared points go through shared memory: broadcasting 8 values per point with shuffles + selects cost 14 of the
  // 42 instructions of a tap-loop iteration; two 16-byte LDS (same address for the 8 lanes of a group) replace them and the
  // four tap offsets are pre-multiplied (measured 162 -> 153 us stand-alone).  Tried and dropped: 8 x 4 token patches per
  // CTA instead of 32 consecutive tokens (158 us: the gathers are bound by L1 wavefronts -- four 128-byte segments per
  // LDG.128 -- not by L2 locality).
  __shared__ float4 s_w[MSDA_TOK][16];
  __shared__ int4 s_o[MSDA_TOK][16];
  const int ti = threadIdx.x >> 3;
  if (pa < LP) {
    s_w[ti][pa] = make_float4(A.w00, A.w01, A.w10, A.w11);
    s_o[ti][pa] = make_int4((A.r0 + A.x0) * (D / 4), (A.r0 + A.x1) * (D / 4), (A.r1 + A.x0) * (D / 4), (A.r1 + A.x1) * (D / 4));
  }
  if (pb < LP) {
    s_w[ti][pb] = make_float4(Bp.w00, Bp.w01, Bp.w10, Bp.w11);
    s_o[ti][pb] = make_int4((Bp.r0 + Bp.x0) * (D / 4), (Bp.r0 + Bp.x1) * (D / 4), (Bp.r1 + Bp.x0) * (D / 4),
                            (Bp.r1 + Bp.x1) * (D / 4));
  }
  __syncwarp();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int i = 0; i < LP; ++i) {
    const float4 wq = s_w[ti][i];
    const int4 oq = s_o[ti][i];
    const float w00 = wq.x, w01 = wq.y, w10 = wq.z, w11 = wq.w;
    const float4 v00 = __ldg(vbase + oq.x);
    const float4 v01 = __ldg(vbase + oq.y);
    const float4 v10 = __ldg(vbase + oq.z);
    const float4 v11 = __ldg(vbase + oq.w);
    acc.x = fmaf(w00, v00.x, fmaf(w01, v01.x, fmaf(w10, v10.x, fmaf(w11, v11.x, acc.x))));
    acc.y = fmaf(w00, v00.y, fmaf(w01, v01.y, fmaf(w10, v10.y, fmaf(w11, v11.y, acc.y))));
    acc.z = fmaf(w00, v00.z, fmaf(w01, v01.z, fmaf(w10, v10.z, fmaf(w11, v11.z, acc.z))));
    acc.w = fmaf(w00, v00.w, fmaf(w01, v01.w, fmaf(w10, v10.w, fmaf(w11, v11.w, acc.w))));
  }
  if (!qvalid) return;
  const size_t o4 = ((size_t)tok * D + head * HD) / 4 + sub;
  if (out_lo) {
    float4 hi;
    hi.x = rna_tf32m(acc.x); hi.y = rna_tf32m(acc.y); hi.z = rna_tf32m(acc.z); hi.w = rna_tf32m(acc.w);
    reinterpret_cast<float4*>(out_hi)[o4] = hi;
    reinterpret_cast<float4*>(out_lo)[o4] = make_float4(rna_tf32m(acc.x - hi.x), rna_tf32m(acc.y - hi.y),
                                                        rna_tf32m(acc.z - hi.z), rna_tf32m(acc.w - hi.w));
  } else {
    reinterpret_cast<float4*>(out_hi)[o4] = acc;
  }
}

// x [M,256], pos [pos_mod,256] -> x (hi,lo), q = x + pos (hi,lo)
__global__ void __launch_bounds__(256) split_add_kernel(const float* __restrict__ x, const float* __restrict__ pos,
                                                         int pos_mod, float* __restrict__ x_hi, float* __restrict__ x_lo,
                                                         float* __restrict__ q_hi, float* __restrict__ q_lo, size_t M) {
  const size_t i4 = (size_t)blockIdx.x * 256 + threadIdx.x;  // float4 index
  if (i4 >= M * (D / 4)) return;
  const size_t m = i4 / (D / 4);
  const int c4 = (int)(i4 % (D / 4));
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i4);
  const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + (m % pos_mod) * (D / 4) + c4);
  const float a[4] = {v.x, v.y, v.z, v.w};
  const float qv[4] = {v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w};
  float h[4], l[4], qh[4], ql[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    h[t] = rna_tf32m(a[t]); l[t] = rna_tf32m(a[t] - h[t]);
    qh[t] = rna_tf32m(qv[t]); ql[t] = rna_tf32m(qv[t] - qh[t]);
  }
  reinterpret_cast<float4*>(x_hi)[i4] = make_float4(h[0], h[1], h[2], h[3]);
  reinterpret_cast<float4*>(x_lo)[i4] = make_float4(l[0], l[1], l[2], l[3]);
  reinterpret_cast<float4*>(q_hi)[i4] = make_float4(qh[0], qh[1], qh[2], qh[3]);
  reinterpret_cast<float4*>(q_lo)[i4] = make_float4(ql[0], ql[1], ql[2], ql[3]);
}

struct EncBuffers {
  float *x, *x_hi, *x_lo, *q_hi, *q_lo;     // [M,256]
  float *value, *ol;                        // [M,256], [M,ldo]
  float *att_hi, *att_lo, *proj;            // [M,256]
  float *x1, *x1_hi, *x1_lo;                // [M,256]
  float *h_hi, *h_lo;                       // [M,ffn]
  float *y;                                 // [M,256]
  float *w_hi, *w_lo;                       // split weights of the current layer
  float *b_ol;                              // [ldo] concatenated biases
};

static void enc_take(Workspace& ws, EncBuffers& b, size_t M, int ffn, int ldo) {
  b.x = ws.take<float>(M * D); b.x_hi = ws.take<float>(M * D); b.x_lo = ws.take<float>(M * D);
  b.q_hi = ws.take<float>(M * D); b.q_lo = ws.take<float>(M * D);
  b.value = ws.take<float>(M * D); b.ol = ws.take<float>(M * ldo);
  b.att_hi = ws.take<float>(M * D); b.att_lo = ws.take<float>(M * D); b.proj = ws.take<float>(M * D);
  b.x1 = ws.take<float>(M * D); b.x1_hi = ws.take<float>(M * D); b.x1_lo = ws.take<float>(M * D);
  b.h_hi = ws.take<float>(M * ffn); b.h_lo = ws.take<float>(M * ffn);
  b.y = ws.take<float>(M * D);
  const size_t wmax = (size_t)D * D * 2 + (size_t)ldo * D + (size_t)2 * ffn * D;
  b.w_hi = ws.take<float>(wmax); b.w_lo = ws.take<float>(wmax);
  b.b_ol = ws.take<float>(ldo);
}


static inline size_t enc_wmax(int ffn, int ldo) { return (size_t)D * D * 2 + (size_t)ldo * D + (size_t)2 * ffn * D; }
// prepared blob of one layer: [TF32 hi plane | TF32 lo plane | bf16 hi plane | bf16 lo plane | offsets;attention bias]
// (the bf16 planes of the 3xBF16 GEMM variant take wmax / 2 floats each); both representations are kept so that
// PN_OPT_ENC_BF16X3 can be flipped without re-preparing
static inline size_t enc_layer_floats(int ffn, int ldo) { return 3 * enc_wmax(ffn, ldo) + (size_t)round_up(ldo, 64); }
// TF32 hi/lo splits of one layer's six weight matrices + the concatenated [offsets ; attention] bias
static int enc_prepare_layer(const PnMsdaEncoderLayer& Lw, float* w_hi, float* w_lo, float* b_ol, int ffn, int ldo,
                             int n_off, int n_att, cudaStream_t st, float* w16_hi = nullptr, float* w16_lo = nullptr) {
  float* wo_hi = w_hi + (size_t)D * D;            float* wo_lo = w_lo + (size_t)D * D;
  float* wp_hi = wo_hi + (size_t)ldo * D;         float* wp_lo = wo_lo + (size_t)ldo * D;
  float* w1_hi = wp_hi + (size_t)D * D;           float* w1_lo = wp_lo + (size_t)D * D;
  float* w2_hi = w1_hi + (size_t)ffn * D;         float* w2_lo = w1_lo + (size_t)ffn * D;
  PN_TRY(launch_split_tf32(Lw.value_proj.w, w_hi, w_lo, (size_t)D * D, st));
  PN_TRY(launch_split_tf32(Lw.sampling_offsets.w, wo_hi, wo_lo, (size_t)n_off * D, st));
  PN_TRY(launch_split_tf32(Lw.attention_weights.w, wo_hi + (size_t)n_off * D, wo_lo + (size_t)n_off * D, (size_t)n_att * D,
                           st));
  PN_TRY(launch_split_tf32(Lw.output_proj.w, wp_hi, wp_lo, (size_t)D * D, st));
  PN_TRY(launch_split_tf32(Lw.ffn1.w, w1_hi, w1_lo, (size_t)ffn * D, st));
  PN_TRY(launch_split_tf32(Lw.ffn2.w, w2_hi, w2_lo, (size_t)ffn * D, st));
  if (w16_hi) {  // the same six matrices as bf16 hi / lo planes, identical element offsets
    uint16_t* h = reinterpret_cast<uint16_t*>(w16_hi);
    uint16_t* l = reinterpret_cast<uint16_t*>(w16_lo);
    const size_t o_wo = (size_t)D * D, o_wp = o_wo + (size_t)ldo * D, o_w1 = o_wp + (size_t)D * D, o_w2 = o_w1 + (size_t)ffn * D;
    PN_TRY(launch_split_bf16(Lw.value_proj.w, h, l, (size_t)D * D, st));
    PN_TRY(launch_split_bf16(Lw.sampling_offsets.w, h + o_wo, l + o_wo, (size_t)n_off * D, st));
    PN_TRY(launch_split_bf16(Lw.attention_weights.w, h + o_wo + (size_t)n_off * D, l + o_wo + (size_t)n_off * D,
                             (size_t)n_att * D, st));
    PN_TRY(launch_split_bf16(Lw.output_proj.w, h + o_wp, l + o_wp, (size_t)D * D, st));
    PN_TRY(launch_split_bf16(Lw.ffn1.w, h + o_w1, l + o_w1, (size_t)ffn * D, st));
    PN_TRY(launch_split_bf16(Lw.ffn2.w, h + o_w2, l + o_w2, (size_t)ffn * D, st));
  }
  cudaError_t e = cudaMemcpyAsync(b_ol, Lw.sampling_offsets.b, sizeof(float) * n_off, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(b_ol + n_off, Lw.attention_weights.b, sizeof(float) * n_att, cudaMemcpyDeviceToDevice, st);
  PN_REQUIRE(e == cudaSuccess, (int)e, "msda_encoder: memcpy: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace pn

using namespace pn;

extern "C" {

size_t pn_msda_encoder_prepared_bytes(const PnMsdaEncoderWeights* w) {
  if (!w || w->num_layers < 1 || w->num_layers > PN_MAX_LAYERS) return 0;
  const int ldo = (int)round_up(NH * w->num_levels * w->num_points * 3, 4);
  return (size_t)w->num_layers * enc_layer_floats(w->ffn_dims, ldo) * sizeof(float) + 256;
}

/* builds the static operands of the encoder's tcgen05 GEMMs once per weight version; set w->prepared = blob afterwards */
int pn_msda_encoder_prepare(const PnMsdaEncoderWeights* w, void* blob, size_t blob_bytes, pn_stream_t stream) {
  PN_REQUIRE(w && blob && blob_bytes >= pn_msda_encoder_prepared_bytes(w) && ((uintptr_t)blob & 255) == 0, PN_ERR_BAD_ARG,
             "msda_encoder_prepare: bad blob");
  const int ffn = w->ffn_dims;
  const int LP = w->num_levels * w->num_points;
  const int n_off = NH * LP * 2, n_att = NH * LP;
  const int ldo = (int)round_up(n_off + n_att, 4);
  for (int i = 0; i < w->num_layers; ++i) {
    float* base = reinterpret_cast<float*>(blob) + (size_t)i * enc_layer_floats(ffn, ldo);
    const size_t wm = enc_wmax(ffn, ldo);
    PN_TRY(enc_prepare_layer(w->layers[i], base, base + wm, base + 3 * wm, ffn, ldo, n_off, n_att, as_stream(stream),
                             base + 2 * wm, base + 2 * wm + wm / 2));
  }
  return 0;
}

size_t pn_msda_encoder_workspace_bytes(int B, int nq, int ffn_dims, int num_levels, int num_points) {
  Workspace ws(nullptr, 0);
  EncBuffers b;
  const int ldo = (int)round_up(NH * num_levels * num_points * 3, 4);
  enc_take(ws, b, (size_t)B * nq, ffn_dims, ldo);
  return ws.off + 1024;
}

int pn_msda_encoder_forward(const PnMsdaEncoderWeights* w, const float* x_in, const float* pos, const int* h,
                            const int* wd, float* x_out, int B, void* wsp, size_t ws_bytes, pn_stream_t stream) {
  PN_REQUIRE(w && x_in && pos && h && wd && x_out && wsp, PN_ERR_BAD_ARG, "msda_encoder: null argument");
  PN_REQUIRE(w->num_levels >= 1 && w->num_levels <= MSDA_MAX_LEVELS && w->num_points >= 1 &&
                 w->num_levels * w->num_points <= 16,
             PN_ERR_UNSUPPORTED, "msda_encoder: levels*points must be <= 16");
  PN_REQUIRE(w->num_layers >= 1 && w->num_layers <= PN_MAX_LAYERS && w->ffn_dims % 32 == 0, PN_ERR_BAD_ARG,
             "msda_encoder: bad layer count / ffn dims");
  cudaStream_t st = as_stream(stream);
  MsdaGeom g{};
  g.L = w->num_levels; g.P = w->num_points;
  int nq = 0;
  for (int l = 0; l < g.L; ++l) { g.h[l] = h[l]; g.w[l] = wd[l]; g.start[l] = nq; nq += h[l] * wd[l]; }
  g.nq = nq;
  const size_t M = (size_t)B * nq;
  const int ffn = w->ffn_dims;
  const int LP = g.L * g.P;
  const int n_off = NH * LP * 2, n_att = NH * LP, n_ol = n_off + n_att;
  const int ldo = (int)round_up(n_ol, 4);
  Workspace ws(wsp, ws_bytes);
  EncBuffers b;
  enc_take(ws, b, M, ffn, ldo);
  PN_REQUIRE(ws.ok(), PN_ERR_WORKSPACE, "msda_encoder: workspace too small (%zu needed, %zu given)", ws.off, ws.cap);

  const int Mi = (int)M;
  // raw mode: activations enter the tcgen05 GEMM as plain fp32 and are split inside the SM (TMEM), so no
  // producer materialises hi/lo copies; otherwise every producer emits its output pre-split.
  const bool raw = get_option(OPT_UMMA_RAW_A) != 0;
  // 3xBF16 (umma_gemm.cu, W16 variant): the encoder's six GEMMs per layer on the kind::f16 pipe at twice the TF32 rate,
  // ~1e-5 of the scale instead of ~1e-6 -- its inputs come from single-pass-TF32 cuDNN convolutions (1e-3 class).
  // Needs the prepared blob (bf16 weight planes), raw-A mode, tensor cores on and the fp32-parity pass count.
  const int w16 = (raw && w->prepared && get_option(OPT_ENC_BF16X3) && get_option(OPT_TENSOR_CORES) &&
                   !get_option(OPT_SINGLE_PASS) && get_option(OPT_UMMA_TMA_STORE) && ffn % 64 == 0) ? 1 : 0;
  float* q_raw = b.q_hi;  // raw mode: q = x + pos lives here
  if (raw) {
    PN_TRY(launch_add_rows(x_in, pos, q_raw, B, nq, st));
  } else {
    split_add_kernel<<<cdiv((long long)M * (D / 4), 256), 256, 0, st>>>(x_in, pos, nq, b.x_hi, b.x_lo, b.q_hi, b.q_lo, M);
    PN_TRY(check_launch("split_add_kernel"));
  }
  const float* x_cur = x_in;
  for (int i = 0; i < w->num_layers; ++i) {
    const PnMsdaEncoderLayer& Lw = w->layers[i];
    // split this layer's weights: [value_proj | sampling_offsets ; attention_weights | output_proj | ffn1 | ffn2]
    // this layer's split weights [value_proj | sampling_offsets ; attention_weights | output_proj | ffn1 | ffn2] and
    // concatenated offset / attention biases: from the prepared blob (static: built once by pn_msda_encoder_prepare)
    // or, without one, split into the workspace on every call (6 launches + 2 copies per layer)
    float *w_hi_l = b.w_hi, *w_lo_l = b.w_lo, *b_ol_l = b.b_ol;
    if (w->prepared) {
      float* base = reinterpret_cast<float*>(const_cast<void*>(w->prepared)) + (size_t)i * enc_layer_floats(ffn, ldo);
      const size_t wm = enc_wmax(ffn, ldo);
      w_hi_l = base; w_lo_l = base + wm; b_ol_l = base + 3 * wm;
      if (w16) { w_hi_l = base + 2 * wm; w_lo_l = base + 2 * wm + wm / 2; }
    } else {
      PN_TRY(enc_prepare_layer(Lw, w_hi_l, w_lo_l, b_ol_l, ffn, ldo, n_off, n_att, st));
    }
    // sub-matrix offsets are in ELEMENTS: fp32 containers (TF32 planes) or 2-byte bf16 (3xBF16 planes)
    auto at = [&](float* base, size_t elems) -> float* {
      return w16 ? reinterpret_cast<float*>(reinterpret_cast<uint16_t*>(base) + elems) : base + elems;
    };
    const size_t o_wo = (size_t)D * D, o_wp = o_wo + (size_t)ldo * D, o_w1 = o_wp + (size_t)D * D, o_w2 = o_w1 + (size_t)ffn * D;
    float* wv_hi = w_hi_l;                          float* wv_lo = w_lo_l;
    float* wo_hi = at(w_hi_l, o_wo);                float* wo_lo = at(w_lo_l, o_wo);
    float* wp_hi = at(w_hi_l, o_wp);                float* wp_lo = at(w_lo_l, o_wp);
    float* w1_hi = at(w_hi_l, o_w1);                float* w1_lo = at(w_lo_l, o_w1);
    float* w2_hi = at(w_hi_l, o_w2);                float* w2_lo = at(w_lo_l, o_w2);
    {  // value = x Wv^T + bv ; ol = q [Wo;Wa]^T + [bo;ba]
      UmmaOperand o[2] = {{b.x_hi, b.x_lo, D, wv_hi, wv_lo, D, Lw.value_proj.b, b.value, D, Mi, D, D},
                          {b.q_hi, b.q_lo, D, wo_hi, wo_lo, D, b_ol_l, b.ol, ldo, Mi, n_ol, D}};
      if (raw) {
        o[0].a_hi = x_cur; o[0].a_lo = nullptr; o[0].a_is_raw = 1;
        o[1].a_hi = q_raw; o[1].a_lo = nullptr; o[1].a_is_raw = 1;
      }
      o[0].w_bf16 = o[1].w_bf16 = w16;
      PN_TRY(launch_umma_gemm(o, 2, 3, st));
    }
    {
      dim3 grid(cdiv(nq, MSDA_TOK), NH, B);
      msda_sample_kernel<<<grid, 256, 0, st>>>(b.value, b.ol, ldo, b.att_hi, raw ? nullptr : b.att_lo, g, B);
      PN_TRY(check_launch("msda_sample_kernel"));
    }
    {
      UmmaOperand o{b.att_hi, raw ? nullptr : b.att_lo, D, wp_hi, wp_lo, D, Lw.output_proj.b, b.proj, D, Mi, D, D};
      o.a_is_raw = raw;
      o.w_bf16 = w16;
      PN_TRY(launch_umma_gemm(&o, 1, 3, st));
      LnArgs n{};
      n.x = b.proj; n.nparts = 1; n.resid = x_cur; n.gamma = Lw.norm[0].gamma; n.beta = Lw.norm[0].beta;
      n.y = b.x1; n.M = Mi;
      if (!raw) { n.y_hi = b.x1_hi; n.y_lo = b.x1_lo; }
      PN_TRY(launch_layernorm(n, st));
    }
    {
      UmmaOperand o1{b.x1_hi, b.x1_lo, D, w1_hi, w1_lo, D, Lw.ffn1.b, b.h_hi, ffn, Mi, ffn, D, b.h_lo, 1};
      if (raw) { o1.a_hi = b.x1; o1.a_lo = nullptr; o1.a_is_raw = 1; o1.C_lo = nullptr; }  // h stays raw fp32
      o1.w_bf16 = w16;
      PN_TRY(launch_umma_gemm(&o1, 1, 3, st));
      UmmaOperand o2{b.h_hi, raw ? nullptr : b.h_lo, ffn, w2_hi, w2_lo, ffn, Lw.ffn2.b, b.y, D, Mi, D, ffn};
      o2.a_is_raw = raw;
      o2.w_bf16 = w16;
      PN_TRY(launch_umma_gemm(&o2, 1, 3, st));
      const bool last = (i + 1 == w->num_layers);
      LnArgs n{};
      n.x = b.y; n.nparts = 1; n.resid = b.x1; n.gamma = Lw.norm[1].gamma; n.beta = Lw.norm[1].beta;
      n.y = last ? x_out : b.x; n.M = Mi;
      if (!last) {
        n.pos = pos; n.pos_mod = nq;
        if (raw) {
          n.ypos = q_raw;
        } else {
          n.y_hi = b.x_hi; n.y_lo = b.x_lo; n.ypos_hi = b.q_hi; n.ypos_lo = b.q_lo;
        }
      }
      PN_TRY(launch_layernorm(n, st));
      x_cur = b.x;
    }
  }
  return 0;
}

/* stand-alone sampling core (stage test): value [B,nq,256], ol [B*nq, 8*L*P*3] -> out [B*nq,256] */
int pn_msda_sample(const float* value, const float* ol, float* out, const int* h, const int* wd, int num_levels,
                   int num_points, int B, pn_stream_t stream) {
  PN_REQUIRE(value && ol && out && h && wd, PN_ERR_BAD_ARG, "msda_sample: null argument");
  PN_REQUIRE(num_levels >= 1 && num_levels <= MSDA_MAX_LEVELS && num_levels * num_points <= 16, PN_ERR_UNSUPPORTED,
             "msda_sample: levels*points must be <= 16");
  MsdaGeom g{};
  g.L = num_levels; g.P = num_points;
  int nq = 0;
  for (int l = 0; l < g.L; ++l) { g.h[l] = h[l]; g.w[l] = wd[l]; g.start[l] = nq; nq += h[l] * wd[l]; }
  g.nq = nq;
  dim3 grid(cdiv(nq, MSDA_TOK), NH, B);
  msda_sample_kernel<<<grid, 256, 0, as_stream(stream)>>>(value, ol, NH * num_levels * num_points * 3, out,
                                                                     nullptr, g, B);
  return check_launch("msda_sample_kernel");
}

}  // extern "C"
