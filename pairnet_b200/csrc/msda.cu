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

__device__ __forceinline__ float rna_tf32m(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// value [B,nq,256] ; ol [B*nq, ldo]: cols [0, 8*L*P*2) offsets ((h,l,p),xy), then 8*L*P attention logits ((h),(l,p))
// out (hi, lo) [B*nq,256]
// CTA = MSDA_TOK consecutive tokens (neighbours along x) of ONE head: neighbouring queries sample overlapping
// taps of the same 128-byte head slice, so most gathers hit L1 instead of L2.
constexpr int MSDA_TOK = 16;
__global__ void __launch_bounds__(MSDA_TOK * 32) msda_sample_kernel(const float* __restrict__ value,
                                                                     const float* __restrict__ ol, int ldo,
                                                                     float* __restrict__ out_hi,
                                                                     float* __restrict__ out_lo, const MsdaGeom g,
                                                                     int B) {
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int q = blockIdx.x * MSDA_TOK + (threadIdx.x >> 5);
  if (q >= g.nq) return;
  const long long tok = (long long)b * g.nq + q;
  const int LP = g.L * g.P;
  // reference point of this token (cell centre of its own level, normalised)
  int ql = 0;
  while (ql + 1 < g.L && q >= g.start[ql + 1]) ++ql;
  const int qi = q - g.start[ql];
  const float ref_x = ((float)(qi % g.w[ql]) + 0.5f) / (float)g.w[ql];
  const float ref_y = ((float)(qi / g.w[ql]) + 0.5f) / (float)g.h[ql];

  const float* row = ol + (size_t)tok * ldo;
  // lanes [0, LP): one sampling point each
  float logit = -INFINITY, ox = 0.f, oy = 0.f;
  if (lane < LP) {
    logit = __ldg(row + NH * LP * 2 + head * LP + lane);
    const float2 o = __ldg(reinterpret_cast<const float2*>(row + (head * LP + lane) * 2));
    ox = o.x; oy = o.y;
  }
  float mx = logit;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  float e = (lane < LP) ? __expf(logit - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  const float aw = e / sum;
  // pixel coordinates in the sampled level: (ref + off / (W,H)) * (W,H) - 0.5   (grid_sample, align_corners=False)
  int lvl = 0, W = 1, H = 1, st = 0;
  float px = 0.f, py = 0.f;
  if (lane < LP) {
    lvl = lane / g.P;
    W = g.w[lvl]; H = g.h[lvl]; st = g.start[lvl];
    px = (ref_x + ox / (float)W) * (float)W - 0.5f;
    py = (ref_y + oy / (float)H) * (float)H - 0.5f;
  }
  const float* vbase = value + (size_t)b * g.nq * D + head * HD + lane;
  float acc = 0.f;
  for (int i = 0; i < LP; ++i) {
    const float x = __shfl_sync(0xffffffffu, px, i), y = __shfl_sync(0xffffffffu, py, i);
    const float a = __shfl_sync(0xffffffffu, aw, i);
    const int w_ = __shfl_sync(0xffffffffu, W, i), h_ = __shfl_sync(0xffffffffu, H, i);
    const int s_ = __shfl_sync(0xffffffffu, st, i);
    const float xf = floorf(x), yf = floorf(y);
    const int x0 = (int)xf, y0 = (int)yf;
    const float lx = x - xf, ly = y - yf;
    const float w00 = (1.f - lx) * (1.f - ly), w01 = lx * (1.f - ly), w10 = (1.f - lx) * ly, w11 = lx * ly;
    const bool xin0 = x0 >= 0 && x0 < w_, xin1 = x0 + 1 >= 0 && x0 + 1 < w_;
    const bool yin0 = y0 >= 0 && y0 < h_, yin1 = y0 + 1 >= 0 && y0 + 1 < h_;
    float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
    if (yin0 && xin0) v00 = __ldg(vbase + (size_t)(s_ + y0 * w_ + x0) * D);
    if (yin0 && xin1) v01 = __ldg(vbase + (size_t)(s_ + y0 * w_ + x0 + 1) * D);
    if (yin1 && xin0) v10 = __ldg(vbase + (size_t)(s_ + (y0 + 1) * w_ + x0) * D);
    if (yin1 && xin1) v11 = __ldg(vbase + (size_t)(s_ + (y0 + 1) * w_ + x0 + 1) * D);
    acc = fmaf(a, w00 * v00 + w01 * v01 + w10 * v10 + w11 * v11, acc);
  }
  const size_t o = (size_t)tok * D + head * HD + lane;
  if (out_lo) {
    const float hi = rna_tf32m(acc);
    out_hi[o] = hi;
    out_lo[o] = rna_tf32m(acc - hi);
  } else {
    out_hi[o] = acc;
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

}  // namespace pn

using namespace pn;

extern "C" {

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
                 w->num_levels * w->num_points <= 32,
             PN_ERR_UNSUPPORTED, "msda_encoder: levels*points must be <= 32");
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
  auto memcpy_d2d = [&](void* d, const void* s, size_t bytes) -> int {
    cudaError_t e = cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, st);
    PN_REQUIRE(e == cudaSuccess, (int)e, "msda_encoder: memcpy: %s", cudaGetErrorString(e));
    return 0;
  };
  split_add_kernel<<<cdiv((long long)M * (D / 4), 256), 256, 0, st>>>(x_in, pos, nq, b.x_hi, b.x_lo, b.q_hi, b.q_lo, M);
  PN_TRY(check_launch("split_add_kernel"));
  const float* x_cur = x_in;
  for (int i = 0; i < w->num_layers; ++i) {
    const PnMsdaEncoderLayer& Lw = w->layers[i];
    // split this layer's weights: [value_proj | sampling_offsets ; attention_weights | output_proj | ffn1 | ffn2]
    float* wv_hi = b.w_hi;                          float* wv_lo = b.w_lo;
    float* wo_hi = wv_hi + (size_t)D * D;           float* wo_lo = wv_lo + (size_t)D * D;
    float* wp_hi = wo_hi + (size_t)ldo * D;         float* wp_lo = wo_lo + (size_t)ldo * D;
    float* w1_hi = wp_hi + (size_t)D * D;           float* w1_lo = wp_lo + (size_t)D * D;
    float* w2_hi = w1_hi + (size_t)ffn * D;         float* w2_lo = w1_lo + (size_t)ffn * D;
    PN_TRY(launch_split_tf32(Lw.value_proj.w, wv_hi, wv_lo, (size_t)D * D, st));
    PN_TRY(launch_split_tf32(Lw.sampling_offsets.w, wo_hi, wo_lo, (size_t)n_off * D, st));
    PN_TRY(launch_split_tf32(Lw.attention_weights.w, wo_hi + (size_t)n_off * D, wo_lo + (size_t)n_off * D,
                             (size_t)n_att * D, st));
    PN_TRY(launch_split_tf32(Lw.output_proj.w, wp_hi, wp_lo, (size_t)D * D, st));
    PN_TRY(launch_split_tf32(Lw.ffn1.w, w1_hi, w1_lo, (size_t)ffn * D, st));
    PN_TRY(launch_split_tf32(Lw.ffn2.w, w2_hi, w2_lo, (size_t)ffn * D, st));
    PN_TRY(memcpy_d2d(b.b_ol, Lw.sampling_offsets.b, sizeof(float) * n_off));
    PN_TRY(memcpy_d2d(b.b_ol + n_off, Lw.attention_weights.b, sizeof(float) * n_att));
    {  // value = x Wv^T + bv ; ol = q [Wo;Wa]^T + [bo;ba]
      UmmaOperand o[2] = {{b.x_hi, b.x_lo, D, wv_hi, wv_lo, D, Lw.value_proj.b, b.value, D, Mi, D, D},
                          {b.q_hi, b.q_lo, D, wo_hi, wo_lo, D, b.b_ol, b.ol, ldo, Mi, n_ol, D}};
      PN_TRY(launch_umma_gemm(o, 2, 3, st));
    }
    {
      dim3 grid(cdiv(nq, MSDA_TOK), NH, B);
      msda_sample_kernel<<<grid, MSDA_TOK * 32, 0, st>>>(b.value, b.ol, ldo, b.att_hi, b.att_lo, g, B);
      PN_TRY(check_launch("msda_sample_kernel"));
    }
    {
      UmmaOperand o{b.att_hi, b.att_lo, D, wp_hi, wp_lo, D, Lw.output_proj.b, b.proj, D, Mi, D, D};
      PN_TRY(launch_umma_gemm(&o, 1, 3, st));
      LnArgs n{};
      n.x = b.proj; n.nparts = 1; n.resid = x_cur; n.gamma = Lw.norm[0].gamma; n.beta = Lw.norm[0].beta;
      n.y = b.x1; n.y_hi = b.x1_hi; n.y_lo = b.x1_lo; n.M = Mi;
      PN_TRY(launch_layernorm(n, st));
    }
    {
      UmmaOperand o1{b.x1_hi, b.x1_lo, D, w1_hi, w1_lo, D, Lw.ffn1.b, b.h_hi, ffn, Mi, ffn, D, b.h_lo, 1};
      PN_TRY(launch_umma_gemm(&o1, 1, 3, st));
      UmmaOperand o2{b.h_hi, b.h_lo, ffn, w2_hi, w2_lo, ffn, Lw.ffn2.b, b.y, D, Mi, D, ffn};
      PN_TRY(launch_umma_gemm(&o2, 1, 3, st));
      const bool last = (i + 1 == w->num_layers);
      LnArgs n{};
      n.x = b.y; n.nparts = 1; n.resid = b.x1; n.gamma = Lw.norm[1].gamma; n.beta = Lw.norm[1].beta;
      n.y = last ? x_out : b.x; n.M = Mi;
      if (!last) {
        n.y_hi = b.x_hi; n.y_lo = b.x_lo; n.pos = pos; n.pos_mod = nq; n.ypos_hi = b.q_hi; n.ypos_lo = b.q_lo;
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
  PN_REQUIRE(num_levels >= 1 && num_levels <= MSDA_MAX_LEVELS && num_levels * num_points <= 32, PN_ERR_UNSUPPORTED,
             "msda_sample: levels*points must be <= 32");
  MsdaGeom g{};
  g.L = num_levels; g.P = num_points;
  int nq = 0;
  for (int l = 0; l < g.L; ++l) { g.h[l] = h[l]; g.w[l] = wd[l]; g.start[l] = nq; nq += h[l] * wd[l]; }
  g.nq = nq;
  dim3 grid(cdiv(nq, MSDA_TOK), NH, B);
  msda_sample_kernel<<<grid, MSDA_TOK * 32, 0, as_stream(stream)>>>(value, ol, NH * num_levels * num_points * 3, out,
                                                                     nullptr, g, B);
  return check_launch("msda_sample_kernel");
}

}  // extern "C"
