// C-ABI entry points (include/pairnet_b200.h) and the host-side orchestration of the kernel graph:
// Mask2Former masked-attention decoder -> Pair Proposal Network -> Relation Fusion -> output gathers.
// Everything is enqueued on the caller's stream with no synchronisation or allocation, so a whole
// CrossHead2.forward is one CUDA-graph-capturable launch sequence.
#include <stdarg.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: the calls are no-ops unless a profiler injected its library

#include "common.cuh"

namespace pn {

// NVTX range per stage of the hot path (tracing): visible in `ncu --nvtx` / Nsight Systems timelines of an
// eager (non-graph) forward.  PN_OPT_NVTX = 0 removes even the function-pointer check.
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name);
  ~NvtxRange() { if (on) nvtxRangePop(); }
};

static thread_local char g_err[512] = "ok";
static thread_local int g_launches = 0;
// Process-wide configuration knobs (pn_set_option): read at launch time by every entry point, meant to be set once
// before the first forward (tests flip them between calls on one thread); they are not per-call state.
static int g_options[OPT_COUNT] = {1, 0, 0, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1};
int get_option(int key) { return (key >= 0 && key < OPT_COUNT) ? g_options[key] : 0; }

NvtxRange::NvtxRange(const char* name) : on(get_option(OPT_NVTX) != 0) { if (on) nvtxRangePushA(name); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { ++g_launches; }
int sm_count() {
  static int cached[PN_MAX_DEVICES] = {0};
  const int dev = current_device();
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}
int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// ---- side stream: independent chains of one forward (memory-side K/V projections, the output head) run on a
// second stream, forked from / joined to the caller's stream with events, so that under CUDA-graph capture they
// become parallel branches.  OPT_OVERLAP = 0 keeps everything on the caller's stream.
struct Side {
  cudaStream_t s2 = nullptr;
  cudaEvent_t ev[256];  // ring: an event is re-recorded only long after the wait on its previous record was enqueued (~30 per forward)
  int next = 0;
  bool ok = false;
};
static Side* get_side() {
  static thread_local Side sides[PN_MAX_DEVICES];  // streams / events belong to the device that was current at creation
  Side& side = sides[current_device()];
  if (!side.ok) {
    if (cudaStreamCreateWithFlags(&side.s2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (auto& e : side.ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    side.ok = true;
  }
  return &side;
}
static cudaEvent_t side_record(Side* sd, cudaStream_t on) {
  cudaEvent_t e = sd->ev[sd->next];
  sd->next = (sd->next + 1) % 256;
  cudaEventRecord(e, on);
  return e;
}
static int side_wait(cudaStream_t who, cudaEvent_t e) {
  cudaError_t r = cudaStreamWaitEvent(who, e, 0);
  PN_REQUIRE(r == cudaSuccess, (int)r, "cudaStreamWaitEvent: %s", cudaGetErrorString(r));
  return 0;
}

static int copy_async(void* d, const void* s, size_t bytes, cudaStream_t st) {
  ++g_launches;
  cudaError_t e = cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) { set_error("cudaMemcpyAsync: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

static int linear1(const float* A, int lda, const PnLinear& L, float* C, int ldc, int M, int N, int K, int relu,
                   cudaStream_t st) {
  GemmBatch b{};
  b.p[0] = make_linear(A, lda, L.w, L.b, C, ldc, M, N, K, relu);
  b.p[0].w_static = 1;  // a model parameter: the skinny kernel fetches it ahead of the PDL wait
  b.count = 1;
  return launch_gemm(b, st);
}

// Linear-ReLU-Linear-ReLU-Linear on [M,256]
static int mlp3(const float* x, const PnMlp3& m, float* t1, float* t2, float* y, int M, cudaStream_t st) {
  PN_TRY(linear1(x, D, m.l[0], t1, D, M, D, D, 1, st));
  PN_TRY(linear1(t1, D, m.l[1], t2, D, M, D, D, 1, st));
  return linear1(t2, D, m.l[2], y, D, M, D, D, 0, st);
}

// K / V^T of one layer in the split form the tensor-core attention consumes
struct FaKV {
  const float *k_hi, *k_lo, *vt_hi, *vt_lo;
  int ldv;
  int ldk = 0, k_col0 = 0, vt_img_rows = 0, vt_row0 = 0;  // strided views (0 = dense), see FaArgs
};

struct LayerScratch {
  float *x1, *x2;       // [M,256]
  float *qp;            // [M,256] projected cross-attn queries
  float *att, *proj;    // [M,256]
  float *qk;            // [M,512]
  float *vv;            // [M,256]
  float *ffh;           // [M,ffn]
  float *parts;         // [FFN_SPLITS][M,256]
  void* mha_ws; size_t mha_ws_bytes;
};
constexpr int FFN_SPLITS = 8;

static size_t layer_scratch_take(Workspace& ws, LayerScratch& s, int M, int ffn, size_t mha_bytes) {
  s.x1 = ws.take<float>((size_t)M * D);
  s.x2 = ws.take<float>((size_t)M * D);
  s.qp = ws.take<float>((size_t)M * D);
  s.att = ws.take<float>((size_t)M * D);
  s.proj = ws.take<float>((size_t)M * D);
  s.qk = ws.take<float>((size_t)M * 2 * D);
  s.vv = ws.take<float>((size_t)M * D);
  s.ffh = ws.take<float>((size_t)M * ffn);
  s.parts = ws.take<float>((size_t)FFN_SPLITS * M * D);
  s.mha_ws = ws.take<char>(mha_bytes);
  s.mha_ws_bytes = mha_bytes;
  return ws.off;
}

// One mmcv BaseTransformerLayer (cross_attn, norm, self_attn, norm, ffn, norm), post-norm residuals.
//   x [B*Nq,256] in/out (updated in place), xpos = x + qpos in/out.
//   kproj/vproj: already projected cross-attention keys/values [B,Nk,256].
static int decoder_layer(const PnDecoderLayer& L, int ffn, float* x, float* xpos, const float* qpos, int B, int Nq,
                         const float* kproj, const float* vproj, int Nk, const uint32_t* bits, int words,
                         const int* rowany, const PnNorm* post_norm, float* xn, LayerScratch& s, cudaStream_t st,
                         Side* sd = nullptr, cudaEvent_t* kv_free = nullptr, const FaKV* fa = nullptr) {
  const int M = B * Nq;
  PN_REQUIRE(ffn % (FFN_SPLITS * 32) == 0, PN_ERR_UNSUPPORTED, "ffn_dims=%d must be a multiple of %d", ffn,
             FFN_SPLITS * 32);
  // ---- cross attention: q = (x + qpos) Wq^T + bq
  {
    PnLinear q{L.cross_attn.in_proj_w, L.cross_attn.in_proj_b};
    PN_TRY(linear1(xpos, D, q, s.qp, D, M, D, D, 0, st));
    if (fa) {
      // tcgen05 flash attention: q scaled + split hi/lo (s.qk doubles as the two [M,256] halves)
      float* q_hi = s.qk;
      float* q_lo = s.qk + (size_t)M * D;
      PN_TRY(launch_split_tf32_scaled(s.qp, q_hi, q_lo, (size_t)M * D, ATTN_QSCALE, st));
      FaArgs a{q_hi, q_lo, fa->k_hi, fa->k_lo, fa->vt_hi, fa->vt_lo, fa->ldv, bits, words, rowany, s.att, B, Nq, Nk};
      a.ldk = fa->ldk; a.k_col0 = fa->k_col0; a.vt_img_rows = fa->vt_img_rows; a.vt_row0 = fa->vt_row0;
      PN_TRY(launch_fa_umma(a, s.mha_ws, s.mha_ws_bytes, st));
    } else {
      MhaArgs a{s.qp, D, kproj, D, vproj, D, bits, words, rowany, s.att, B, Nq, Nk};
      PN_TRY(launch_mha(a, s.mha_ws, s.mha_ws_bytes, st));
    }
    if (sd && kv_free) *kv_free = side_record(sd, st);  // K/V of this layer may be overwritten from here on
    PnLinear o{L.cross_attn.out_proj_w, L.cross_attn.out_proj_b};
    PN_TRY(linear1(s.att, D, o, s.proj, D, M, D, D, 0, st));
    LnArgs n{};
    n.x = s.proj; n.nparts = 1; n.resid = x; n.gamma = L.norm[0].gamma; n.beta = L.norm[0].beta;
    n.y = s.x1; n.pos = qpos; n.pos_mod = Nq; n.ypos = xpos; n.M = M;
    n.zero_rows = const_cast<int*>(rowany);  // the attention above was the last reader: clear for the next layer's bits
    PN_TRY(launch_layernorm(n, st));
  }
  // ---- self attention: q = k = (x1 + qpos) W{q,k}^T, v = x1 Wv^T
  {
    GemmBatch g{};
    g.p[0] = make_linear(xpos, D, L.self_attn.in_proj_w, L.self_attn.in_proj_b, s.qk, 2 * D, M, 2 * D, D);
    g.p[0].w_static = 1;
    g.p[1] = make_linear(s.x1, D, L.self_attn.in_proj_w + (size_t)2 * D * D, L.self_attn.in_proj_b + 2 * D, s.vv, D,
                         M, D, D);
    g.p[1].w_static = 1;
    g.count = 2;
    PN_TRY(launch_gemm(g, st));
    MhaArgs a{s.qk, 2 * D, s.qk + D, 2 * D, s.vv, D, nullptr, 0, nullptr, s.att, B, Nq, Nq};
    PN_TRY(launch_mha(a, s.mha_ws, s.mha_ws_bytes, st));
    PnLinear o{L.self_attn.out_proj_w, L.self_attn.out_proj_b};
    PN_TRY(linear1(s.att, D, o, s.proj, D, M, D, D, 0, st));
    LnArgs n{};
    n.x = s.proj; n.nparts = 1; n.resid = s.x1; n.gamma = L.norm[1].gamma; n.beta = L.norm[1].beta;
    n.y = s.x2; n.M = M;
    PN_TRY(launch_layernorm(n, st));
  }
  // ---- FFN: x3 = LN(x2 + relu(x2 W1^T + b1) W2^T + b2), split-K partials reduced inside the LN
  {
    PN_TRY(linear1(s.x2, D, L.ffn1, s.ffh, ffn, M, ffn, D, 1, st));
    GemmBatch g{};
    g.p[0] = make_linear(s.ffh, ffn, L.ffn2.w, nullptr, s.parts, D, M, D, ffn);
    g.p[0].w_static = 1;
    g.p[0].splits = FFN_SPLITS;
    g.p[0].split_stride = (long long)M * D;
    g.count = 1;
    PN_TRY(launch_gemm(g, st));
    LnArgs n{};
    n.x = s.parts; n.nparts = FFN_SPLITS; n.part_stride = (long long)M * D; n.bias = L.ffn2.b; n.resid = s.x2;
    n.gamma = L.norm[2].gamma; n.beta = L.norm[2].beta;
    n.y = x; n.pos = qpos; n.pos_mod = Nq; n.ypos = xpos; n.M = M;
    if (post_norm) { n.gamma2 = post_norm->gamma; n.beta2 = post_norm->beta; n.y2 = xn; }
    PN_TRY(launch_layernorm(n, st));
  }
  return 0;
}

// ---- mask einsums on the tcgen05 GEMM (rows 2 of SURVEY §8a).  Orientation: keys / pixels are the 128-row M tiles
// (streamed raw, split hi/lo in the SM), the <= 256 queries of one image are the N extent (mask embeddings,
// pre-split), so one accumulator column holds one query and one TMEM lane one key.
//   bits : sign + ballot epilogue -> packed attention mask (one 32-bit word per warp and query)
//   pred : transposed store -> mask_pred[b][q][pixel]
static int maskbits_tc(const float* e, float* eh, float* el, const float* Ftok_l, uint32_t* bits, int* rowany, int B,
                       int N, int hw, int words, cudaStream_t st) {
  PN_REQUIRE(N <= 256, PN_ERR_UNSUPPORTED, "mask bits on tensor cores: at most 256 queries (got %d)", N);
  PN_TRY(launch_split_tf32(e, eh, el, (size_t)B * N * D, st));
  for (int b0 = 0; b0 < B; b0 += 4) {
    UmmaOperand o[4];
    int nb = 0;
    for (int bi = b0; bi < B && nb < 4; ++bi, ++nb) {
      o[nb] = UmmaOperand{Ftok_l + (size_t)bi * hw * D, nullptr, D, eh + (size_t)bi * N * D, el + (size_t)bi * N * D, D,
                          nullptr, nullptr, 0, hw, N, D};
      o[nb].a_is_raw = 1;
      o[nb].bits = bits + (size_t)bi * N * words;
      o[nb].rowany = rowany + (size_t)bi * N;
      o[nb].bits_words = words;
    }
    PN_TRY(launch_umma_gemm(o, nb, 3, st));
  }
  return 0;
}
static int mask_pred_tc(const float* e, float* eh, float* el, const float* Ftok, float* mask_pred, int B, int N, int HW,
                        cudaStream_t st) {
  PN_TRY(launch_split_tf32(e, eh, el, (size_t)B * N * D, st));
  for (int b0 = 0; b0 < B; b0 += 4) {
    UmmaOperand o[4];
    int nb = 0;
    for (int bi = b0; bi < B && nb < 4; ++bi, ++nb) {
      o[nb] = UmmaOperand{Ftok + (size_t)bi * HW * D, nullptr, D, eh + (size_t)bi * N * D, el + (size_t)bi * N * D, D,
                          nullptr, mask_pred + (size_t)bi * N * HW, HW, HW, N, D};
      o[nb].a_is_raw = 1;
      o[nb].t_rows = HW;
    }
    PN_TRY(launch_umma_gemm(o, nb, 3, st));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Prepared weights (pn_rel_prepare / pn_m2f_prepare): static TF32 hi/lo splits, carved from a caller-owned blob.
struct PreparedLayer {
  float *cin_hi, *cin_lo, *co_hi, *co_lo, *sin_hi, *sin_lo, *so_hi, *so_lo, *f1_hi, *f1_lo, *f2_hi, *f2_lo;
};
static void carve_layer(Workspace& ws, int ffn, PreparedLayer& L) {
  L.cin_hi = ws.take<float>((size_t)3 * D * D); L.cin_lo = ws.take<float>((size_t)3 * D * D);
  L.co_hi = ws.take<float>((size_t)D * D); L.co_lo = ws.take<float>((size_t)D * D);
  L.sin_hi = ws.take<float>((size_t)3 * D * D); L.sin_lo = ws.take<float>((size_t)3 * D * D);
  L.so_hi = ws.take<float>((size_t)D * D); L.so_lo = ws.take<float>((size_t)D * D);
  L.f1_hi = ws.take<float>((size_t)ffn * D); L.f1_lo = ws.take<float>((size_t)ffn * D);
  L.f2_hi = ws.take<float>((size_t)ffn * D); L.f2_lo = ws.take<float>((size_t)ffn * D);
}
static int prepare_layer(const PnDecoderLayer& w, int ffn, const PreparedLayer& L, cudaStream_t st) {
  PN_TRY(launch_split_tf32(w.cross_attn.in_proj_w, L.cin_hi, L.cin_lo, (size_t)3 * D * D, st));
  PN_TRY(launch_split_tf32(w.cross_attn.out_proj_w, L.co_hi, L.co_lo, (size_t)D * D, st));
  PN_TRY(launch_split_tf32(w.self_attn.in_proj_w, L.sin_hi, L.sin_lo, (size_t)3 * D * D, st));
  PN_TRY(launch_split_tf32(w.self_attn.out_proj_w, L.so_hi, L.so_lo, (size_t)D * D, st));
  PN_TRY(launch_split_tf32(w.ffn1.w, L.f1_hi, L.f1_lo, (size_t)ffn * D, st));
  return launch_split_tf32(w.ffn2.w, L.f2_hi, L.f2_lo, (size_t)ffn * D, st);
}
static void chain_layer_fill(ChainLayer& c, const PnDecoderLayer& w, const PreparedLayer& L) {
  c.cin_hi = L.cin_hi; c.cin_lo = L.cin_lo; c.co_hi = L.co_hi; c.co_lo = L.co_lo; c.sin_hi = L.sin_hi; c.sin_lo = L.sin_lo;
  c.so_hi = L.so_hi; c.so_lo = L.so_lo; c.f1_hi = L.f1_hi; c.f1_lo = L.f1_lo; c.f2_hi = L.f2_hi; c.f2_lo = L.f2_lo;
  c.cin_b = w.cross_attn.in_proj_b; c.co_b = w.cross_attn.out_proj_b; c.sin_b = w.self_attn.in_proj_b;
  c.so_b = w.self_attn.out_proj_b; c.f1_b = w.ffn1.b; c.f2_b = w.ffn2.b;
  for (int i = 0; i < 3; ++i) { c.gamma[i] = w.norm[i].gamma; c.beta[i] = w.norm[i].beta; }
}

struct PreparedM2F {
  PreparedLayer L[PN_MAX_LAYERS];
  float *me_hi[3], *me_lo[3];  // mask_embed MLP [256,256] x3
};
static void carve_m2f(Workspace& ws, const PnM2FWeights* w, PreparedM2F& p) {
  for (int l = 0; l < w->num_layers; ++l) carve_layer(ws, w->ffn_dims, p.L[l]);
  for (int i = 0; i < 3; ++i) {
    p.me_hi[i] = ws.take<float>((size_t)D * D);
    p.me_lo[i] = ws.take<float>((size_t)D * D);
  }
}

// ------------------------------------------------------------------------------------------------
struct M2FPlan {
  int B, N, M, L, nl, ffn;
  int hw[PN_MAX_LEVELS], ldf[PN_MAX_LEVELS];
  int maxhw, maxldf;
  size_t mha_bytes;
  bool tc_mask;    // mask einsums (attention-mask bits, final mask_pred) on the tcgen05 GEMM, token-major operands
  bool need_ftok;  // ... and mask_features arrives NCHW: a token-major copy is made first
};

static int m2f_plan(const PnM2FWeights* w, const PnM2FInputs* in, M2FPlan& p) {
  PN_REQUIRE(w && in, PN_ERR_BAD_ARG, "m2f: null weights/inputs");
  p.B = in->B; p.N = w->num_queries; p.M = p.B * p.N; p.L = w->num_levels; p.nl = w->num_layers; p.ffn = w->ffn_dims;
  PN_REQUIRE(p.B > 0 && p.N > 0, PN_ERR_BAD_ARG, "m2f: bad B/N");
  PN_REQUIRE(p.L >= 1 && p.L <= PN_MAX_LEVELS, PN_ERR_BAD_ARG, "m2f: num_levels out of range");
  PN_REQUIRE(p.nl >= 1 && p.nl <= PN_MAX_LAYERS, PN_ERR_BAD_ARG, "m2f: num_layers out of range");
  PN_REQUIRE(in->H4 > 0 && in->W4 > 0 && in->mask_features, PN_ERR_BAD_ARG, "m2f: bad mask_features");
  p.tc_mask = get_option(OPT_TENSOR_CORES) && get_option(OPT_MASK_TC);
  p.need_ftok = p.tc_mask && !in->mask_features_token_major;
  PN_REQUIRE(p.tc_mask || !in->mask_features_token_major, PN_ERR_UNSUPPORTED,
             "m2f: token-major mask_features need the tensor-core mask path (PN_OPT_TENSOR_CORES / PN_OPT_MASK_TC)");
  p.maxhw = 0; p.maxldf = 0; p.mha_bytes = mha_workspace_bytes(p.B, p.N, p.N);
  for (int l = 0; l < p.L; ++l) {
    PN_REQUIRE(in->h[l] > 0 && in->w[l] > 0 && in->memory[l], PN_ERR_BAD_ARG, "m2f: bad level %d", l);
    p.hw[l] = in->h[l] * in->w[l];
    p.ldf[l] = (int)round_up(p.hw[l], 64);
    p.maxhw = p.hw[l] > p.maxhw ? p.hw[l] : p.maxhw;
    p.maxldf = p.ldf[l] > p.maxldf ? p.ldf[l] : p.maxldf;
    size_t mb = mha_workspace_bytes(p.B, p.N, p.hw[l]);
    p.mha_bytes = mb > p.mha_bytes ? mb : p.mha_bytes;
    mb = fa_workspace_bytes(p.B, p.N, p.hw[l]);
    p.mha_bytes = mb > p.mha_bytes ? mb : p.mha_bytes;
  }
  return 0;
}

static size_t in_hw4(const PnM2FInputs* in, const M2FPlan&) { return (size_t)in->H4 * in->W4; }
constexpr int TC_MIN_ROWS = 1024;  // memory levels with fewer tokens stay on the FFMA GEMM
constexpr size_t PPN_L2_CHUNK_BYTES = 48u << 20;  // pair matrices of one PPN chunk (L2 is 126 MB)
struct M2FBuffers {
  float *X[PN_MAX_LEVELS], *XP[PN_MAX_LEVELS], *Fl[PN_MAX_LEVELS], *pos[PN_MAX_LEVELS];
  float *Xlo[PN_MAX_LEVELS], *XPlo[PN_MAX_LEVELS];  // 3xTF32 low parts (tensor-core K/V projection)
  float *Whi, *Wlo;                                  // split [Wk;Wv] per layer [nl][512,256]
  float *K2[2], *V2[2];                              // K/V double buffer (layer i uses set i & 1)
  float *Klo2[2], *Vlo2[2];                          // tensor-core attention: K lo, and V2/Vlo2 hold V^T hi/lo
  float *K, *V;
  uint32_t* bits; int* rowany;
  float *x, *xpos, *xn, *e1, *e2, *e;
  float *eh, *el;   // mask embeddings split hi/lo (B operand of the tensor-core mask GEMMs)
  float *teh, *tel; // same for the (possibly deferred) output head
  float* Ftok;      // token-major copy of NCHW mask_features (tc_mask only)
  LayerScratch ls;
};

static void m2f_take(Workspace& ws, const M2FPlan& p, const PnM2FInputs* in, M2FBuffers& b, bool own_tail = true) {
  for (int l = 0; l < p.L; ++l) {
    b.X[l] = ws.take<float>((size_t)p.B * p.hw[l] * D);
    b.XP[l] = ws.take<float>((size_t)p.B * p.hw[l] * D);
    b.Xlo[l] = ws.take<float>((size_t)p.B * p.hw[l] * D);
    b.XPlo[l] = ws.take<float>((size_t)p.B * p.hw[l] * D);
    b.Fl[l] = ws.take<float>((size_t)p.B * D * p.ldf[l]);
    b.pos[l] = (in && in->pos[l]) ? nullptr : ws.take<float>((size_t)p.hw[l] * D);
  }
  b.Whi = ws.take<float>((size_t)p.nl * 2 * D * D);
  b.Wlo = ws.take<float>((size_t)p.nl * 2 * D * D);
  for (int t = 0; t < 2; ++t) {
    b.K2[t] = ws.take<float>((size_t)p.B * p.maxhw * D);
    b.V2[t] = ws.take<float>((size_t)p.B * (p.maxhw + 4) * D);
    b.Klo2[t] = ws.take<float>((size_t)p.B * p.maxhw * D);
    b.Vlo2[t] = ws.take<float>((size_t)p.B * (p.maxhw + 4) * D);
  }
  b.K = b.K2[0];
  b.V = b.V2[0];
  b.bits = ws.take<uint32_t>((size_t)p.M * (p.maxldf / 32));
  b.rowany = ws.take<int>((size_t)p.M);
  b.x = ws.take<float>((size_t)p.M * D);
  b.xpos = ws.take<float>((size_t)p.M * D);
  b.xn = ws.take<float>((size_t)p.M * D);
  b.e1 = ws.take<float>((size_t)p.M * D);
  b.e2 = ws.take<float>((size_t)p.M * D);
  b.e = ws.take<float>((size_t)p.M * D);
  b.eh = ws.take<float>((size_t)p.M * D);
  b.el = ws.take<float>((size_t)p.M * D);
  if (own_tail) {  // otherwise the caller provides buffers that outlive this stage (deferred output head)
    b.teh = ws.take<float>((size_t)p.M * D);
    b.tel = ws.take<float>((size_t)p.M * D);
    b.Ftok = p.need_ftok ? ws.take<float>((size_t)p.B * in_hw4(in, p) * D) : nullptr;
  }
  layer_scratch_take(ws, b.ls, p.M, p.ffn, p.mha_bytes);
}

// Deferred output head: when given, cls_pred / mask_pred are produced on the side stream from buffers that
// outlive this stage's scratch; the caller joins on `done` before reading them.
struct TailCtx {
  float *xn, *e1, *e2, *e;   // [M,256] each, caller-owned
  float *eh, *el;            // [M,256] split mask embeddings (tensor-core mask_pred)
  float* ftok;               // [B,HW4,256] token-major mask_features copy, or null when the input already is
  cudaEvent_t done;
  bool deferred;
};

static int m2f_forward(const PnM2FWeights* w, const PnM2FInputs* in, const PnM2FOutputs* out, Workspace& ws,
                       cudaStream_t st, TailCtx* tail = nullptr) {
  NvtxRange nvtx("pn::m2f_decoder (pairnet_head.py:268-320)");
  M2FPlan p;
  PN_TRY(m2f_plan(w, in, p));
  PN_REQUIRE(out && out->cls_pred && out->mask_pred, PN_ERR_BAD_ARG, "m2f: null outputs");
  M2FBuffers b{};
  m2f_take(ws, p, in, b, tail == nullptr);
  PN_REQUIRE(ws.ok() && !ws.dry, PN_ERR_WORKSPACE, "m2f: workspace too small (%zu needed so far, %zu given)", ws.off,
             ws.cap);
  const int HW4 = in->H4 * in->W4;
  Side* sd = get_option(OPT_OVERLAP) ? get_side() : nullptr;
  cudaStream_t s2 = sd ? sd->s2 : st;
  if (tail) {
    b.xn = tail->xn; b.e1 = tail->e1; b.e2 = tail->e2; b.e = tail->e; tail->deferred = false;
    b.teh = tail->eh; b.tel = tail->el; b.Ftok = tail->ftok;
  }
  const bool tcm = p.tc_mask;
  PreparedM2F prep{};
  const bool have_prep = w->prepared != nullptr;
  if (have_prep) {
    Workspace pw(const_cast<void*>(w->prepared), (size_t)1 << 60);
    carve_m2f(pw, w, prep);
  }
  const float* Ftok = nullptr;  // mask_features as [B,HW4,256]
  if (tcm) {
    Ftok = in->mask_features;
    if (p.need_ftok) {
      PN_REQUIRE(b.Ftok, PN_ERR_WORKSPACE, "m2f: no buffer for the token-major mask_features copy");
      PN_TRY(launch_nchw_to_tokens(in->mask_features, b.Ftok, p.B, HW4, st));
      Ftok = b.Ftok;
    }
  }

  // ---- row 1 + the linear half of row 2 that does not depend on the queries
  for (int l = 0; l < p.L; ++l) {
    const float* pos = in->pos[l];
    if (!pos) {
      PN_TRY(launch_sine_posenc(b.pos[l], in->h[l], in->w[l], st));
      pos = b.pos[l];
    }
    const bool tc = get_option(OPT_TENSOR_CORES) && p.B * p.hw[l] >= TC_MIN_ROWS;
    const bool raw_a = tc && get_option(OPT_UMMA_RAW_A);
    // split forms needed by the tensor-core consumers: X as the B operand of V^T = Wv . X^T (always pre-split);
    // X / XP as A operands only when the GEMM does not split A in-kernel
    const bool x_split = tc && (get_option(OPT_FA_TC) || !raw_a);
    const bool xp_split = tc && !raw_a;
    if (in->memory_token_major[l])
      PN_TRY(launch_level_prep_tokens(in->memory[l], in->memory_batch_stride[l], w->level_embed + (size_t)l * D, pos,
                                      b.X[l], b.XP[l], p.B, p.hw[l], st, x_split ? b.Xlo[l] : nullptr,
                                      xp_split ? b.XPlo[l] : nullptr));
    else
      PN_TRY(launch_level_prep(in->memory[l], w->level_embed + (size_t)l * D, pos, b.X[l], b.XP[l], p.B, p.hw[l], st,
                               x_split ? b.Xlo[l] : nullptr, xp_split ? b.XPlo[l] : nullptr));
    if (tcm)
      PN_TRY(launch_mask_feature_resize_tokens(Ftok, b.Fl[l], p.B, in->H4, in->W4, in->h[l], in->w[l], st));
    else
      PN_TRY(launch_mask_feature_resize(in->mask_features, b.Fl[l], p.B, in->H4, in->W4, in->h[l], in->w[l], p.ldf[l],
                                        st));
  }
  if (sd) PN_TRY(side_wait(s2, side_record(sd, st)));  // fork: the side stream sees the prepared memories
  cudaEvent_t kv_free[PN_MAX_LAYERS];
  // ---- learned queries (pairnet_head.py:290-291) and post_norm for the first head call
  PN_TRY(launch_bcast_rows(w->query_feat, w->query_embed, b.x, b.xpos, p.B, p.N, st));
  {
    LnArgs n{};
    n.x = b.x; n.nparts = 1; n.gamma = w->post_norm.gamma; n.beta = w->post_norm.beta; n.y = b.xn; n.M = p.M;
    n.zero_rows = b.rowany;  // every later clear rides on the cross-attention LayerNorm of the layer that consumed it
    PN_TRY(launch_layernorm(n, st));
  }
  for (int i = 0; i < p.nl; ++i) {
    const int l = i % p.L;
    NvtxRange nvtx_layer("pn::m2f_layer");
    const PnDecoderLayer& Lw = w->layers[i];
    const int words = p.ldf[l] / 32;
    float* Kc = b.K2[i & 1];
    float* Vc = b.V2[i & 1];
    // ---- memory side (side stream): K/V projections of this layer's level; independent of the queries, so it
    //      overlaps the query-side chain below.   k = (mem + lvl + pos) Wk^T + bk,  v = (mem + lvl) Wv^T + bv
    if (sd && i >= 2) PN_TRY(side_wait(s2, kv_free[i - 2]));  // attention of layer i-2 has released this K/V set
    const int Mk = p.B * p.hw[l];
    const bool tc = get_option(OPT_TENSOR_CORES) && Mk >= TC_MIN_ROWS;
    const bool fa_tc = tc && get_option(OPT_FA_TC);
    const bool raw_a = tc && get_option(OPT_UMMA_RAW_A);
    FaKV fa{Kc, b.Klo2[i & 1], Vc, b.Vlo2[i & 1], (int)round_up(p.hw[l], 4)};
    if (tc) {
      // tcgen05 path: TMA-staged tiles, UMMA kind::tf32 with hi/lo split operands (fp32 parity)
      float* Whi = b.Whi + (size_t)i * 2 * D * D;
      float* Wlo = b.Wlo + (size_t)i * 2 * D * D;
      if (have_prep) {  // static [Wk;Wv] splits from the prepared blob (rows 256..767 of in_proj)
        Whi = prep.L[i].cin_hi + (size_t)D * D;
        Wlo = prep.L[i].cin_lo + (size_t)D * D;
      } else {
        PN_TRY(launch_split_tf32(Lw.cross_attn.in_proj_w + (size_t)D * D, Whi, Wlo, (size_t)2 * D * D, s2));
      }
      UmmaOperand o[2] = {
          {b.XP[l], b.XPlo[l], D, Whi, Wlo, D, Lw.cross_attn.in_proj_b + D, Kc, D, Mk, D, D},
          {b.X[l], b.Xlo[l], D, Whi + (size_t)D * D, Wlo + (size_t)D * D, D, Lw.cross_attn.in_proj_b + 2 * D, Vc, D, Mk, D,
           D}};
      if (raw_a) {  // A operands enter raw; the kernel splits them through TMEM
        o[0].a_lo = nullptr; o[0].a_is_raw = 1;
        if (!fa_tc) { o[1].a_lo = nullptr; o[1].a_is_raw = 1; }
      }
      if (fa_tc) {
        // operands for fa_umma_kernel: K split hi/lo; V^T per image (keys contiguous) computed directly as
        // Wv . X_b^T (weights as the A operand, per-row bias) so its stores stay row-contiguous.
        o[0].C_lo = b.Klo2[i & 1];
        UmmaOperand batch[4];
        int nb = 0;
        if (raw_a) {
          PN_TRY(launch_umma_gemm(&o[0], 1, 3, s2));  // raw-A and split-A problems use different kernel variants
        } else {
          batch[nb++] = o[0];
        }
        for (int bi = 0; bi < p.B; ++bi) {
          UmmaOperand v{Whi + (size_t)D * D, Wlo + (size_t)D * D, D,
                        b.X[l] + (size_t)bi * p.hw[l] * D, b.Xlo[l] + (size_t)bi * p.hw[l] * D, D,
                        Lw.cross_attn.in_proj_b + 2 * D, Vc + (size_t)bi * D * fa.ldv, fa.ldv, D, p.hw[l], D};
          v.C_lo = b.Vlo2[i & 1] + (size_t)bi * D * fa.ldv;
          v.bias_per_row = 1;
          batch[nb++] = v;
          if (nb == 4 || bi + 1 == p.B) {
            PN_TRY(launch_umma_gemm(batch, nb, 3, s2));
            nb = 0;
          }
        }
      } else {
        PN_TRY(launch_umma_gemm(o, 2, 3, s2));
      }
    } else {
      GemmBatch g{};
      g.p[0] = make_linear(b.XP[l], D, Lw.cross_attn.in_proj_w + (size_t)D * D, Lw.cross_attn.in_proj_b + D, Kc, D, Mk,
                           D, D);
      g.p[1] = make_linear(b.X[l], D, Lw.cross_attn.in_proj_w + (size_t)2 * D * D, Lw.cross_attn.in_proj_b + 2 * D, Vc,
                           D, Mk, D, D);
      g.count = 2;
      PN_TRY(launch_gemm(g, s2));
    }
    cudaEvent_t kv_ready = sd ? side_record(sd, s2) : nullptr;
    // ---- query side: forward_head (mask branch only): attn_mask = (mask_embed(post_norm(x)) . resize(F) < 0)
    PN_TRY(mlp3(b.xn, w->mask_embed, b.e1, b.e2, b.e, p.M, st));
    if (tcm)
      PN_TRY(maskbits_tc(b.e, b.eh, b.el, b.Fl[l], b.bits, b.rowany, p.B, p.N, p.hw[l], words, st));
    else
      PN_TRY(launch_gemm_nmajor_maskbits(b.e, b.Fl[l], b.bits, b.rowany, p.B, p.N, p.hw[l], p.ldf[l], st));
    if (out->mask_trace) {
      PN_REQUIRE(out->trace_words >= words, PN_ERR_BAD_ARG, "m2f: trace_words too small");
      ++g_launches;
      cudaError_t e = cudaMemcpy2DAsync(out->mask_trace + (size_t)i * p.M * out->trace_words,
                                        (size_t)out->trace_words * 4, b.bits, (size_t)words * 4, (size_t)words * 4,
                                        p.M, cudaMemcpyDeviceToDevice, st);
      PN_REQUIRE(e == cudaSuccess, (int)e, "mask trace copy: %s", cudaGetErrorString(e));
    }
    if (sd) PN_TRY(side_wait(st, kv_ready));  // join: attention needs this layer's K/V
    PN_TRY(decoder_layer(Lw, p.ffn, b.x, b.xpos, w->query_embed, p.B, p.N, Kc, Vc, p.hw[l], b.bits, words, b.rowany,
                         &w->post_norm, b.xn, b.ls, st, sd, &kv_free[i], fa_tc ? &fa : nullptr));
    if (out->query_trace) PN_TRY(copy_async(out->query_trace + (size_t)i * p.M * D, b.x, sizeof(float) * p.M * D, st));
  }
  if (out->query_out) PN_TRY(copy_async(out->query_out, b.x, sizeof(float) * p.M * D, st));
  // ---- last forward_head: cls_pred + full-resolution mask_pred (the only ones that are outputs).  With a
  //      TailCtx it runs on the side stream, overlapping the PPN / Relation Fusion stages of the caller.
  const bool defer = sd && tail;
  cudaStream_t ts = defer ? s2 : st;
  if (defer) PN_TRY(side_wait(s2, side_record(sd, st)));
  PN_TRY(linear1(b.xn, D, w->cls_embed, out->cls_pred, w->num_cls, p.M, w->num_cls, D, 0, ts));
  PN_TRY(mlp3(b.xn, w->mask_embed, b.e1, b.e2, b.e, p.M, ts));
  if (tcm) {
    PN_TRY(mask_pred_tc(b.e, b.teh, b.tel, Ftok, out->mask_pred, p.B, p.N, HW4, ts));
  } else {
    GemmProb g = make_linear(b.e, D, in->mask_features, nullptr, out->mask_pred, HW4, p.N, HW4, D);
    g.ldw = HW4;
    g.nb = p.B; g.sA = (long long)p.N * D; g.sW = (long long)D * HW4; g.sC = (long long)p.N * HW4;
    PN_TRY(launch_gemm_nmajor_store(g, ts));
  }
  if (defer) {
    tail->done = side_record(sd, s2);
    tail->deferred = true;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
static int ppn_forward(const float* query, const float* query_obj, const PnMlp3* sub_mlp, const PnMlp3* obj_mlp,
                       const PnConvTiny* conv, float* importance_raw, float* importance, int64_t* topk_idx,
                       int64_t* sub_pos, int64_t* obj_pos, float* pair_feat, int B, int N, int K, Workspace& ws,
                       cudaStream_t st) {
  NvtxRange nvtx("pn::ppn (pairnet_head.py:322-351)");
  PN_REQUIRE(query && importance && sub_pos && obj_pos, PN_ERR_BAD_ARG, "ppn: null pointer");
  PN_REQUIRE((sub_mlp != nullptr) == (obj_mlp != nullptr), PN_ERR_BAD_ARG, "ppn: sub/obj MLP must come together");
  PN_REQUIRE(sub_mlp ? query_obj == nullptr : query_obj != nullptr, PN_ERR_BAD_ARG,
             "ppn: query_obj is only for the MLP-less microbenchmark mode");
  const int M = B * N;
  const float *S = query, *O = query_obj;
  float* h1 = nullptr; float* h2 = nullptr; float* emb = nullptr; float* nrm = nullptr;
  if (sub_mlp) {
    h1 = ws.take<float>((size_t)2 * M * D);
    h2 = ws.take<float>((size_t)2 * M * D);
    emb = ws.take<float>((size_t)2 * M * D);
    nrm = ws.take<float>((size_t)2 * M * D);
  }
  float* raw = importance_raw;
  if (!raw) raw = conv ? ws.take<float>((size_t)B * N * N) : importance;
  const size_t conv_bytes = conv ? conv_tiny_workspace_bytes(B, N, conv->mid_channels) : 0;
  char* conv_ws = conv ? ws.take<char>(conv_bytes) : nullptr;
  int* redo = ws.take<int>((size_t)B);
  PN_REQUIRE(ws.ok() && !ws.dry, PN_ERR_WORKSPACE, "ppn: workspace too small");

  if (sub_mlp) {
    // rows 4: only the last decoder layer's output feeds the pair matrix (pairnet_head.py:325-326)
    const float* in[2] = {query, query};
    float* bufs[3] = {h1, h2, emb};
    for (int s = 0; s < 3; ++s) {
      GemmBatch g{};
      for (int t = 0; t < 2; ++t) {
        const PnMlp3* m = t == 0 ? sub_mlp : obj_mlp;
        g.p[t] = make_linear(in[t], D, m->l[s].w, m->l[s].b, bufs[s] + (size_t)t * M * D, D, M, D, D, s < 2 ? 1 : 0);
        g.p[t].w_static = 1;
      }
      g.count = 2;
      PN_TRY(launch_gemm(g, st));
      in[0] = bufs[s];
      in[1] = bufs[s] + (size_t)M * D;
    }
    PN_TRY(launch_l2norm(emb, nrm, 2 * M, st));
    S = nrm;
    O = nrm + (size_t)M * D;
  }
  // row 5: importance_raw[b] = S[b] O[b]^T.  >= 1024 embedding rows: tcgen05 3xTF32 pair-matrix kernel (operands split
  // in the SM); fewer: exact-fp32 FFMA.
  const bool tc = get_option(OPT_TENSOR_CORES) && get_option(OPT_PPN_TC) && (long long)B * N >= TC_MIN_ROWS;
  auto pair_matrix = [&](const float* s, const float* o, float* c, int nb) -> int {
    if (tc) return launch_pair_matrix_tc(s, o, c, nb, N, D, st);
    GemmBatch g{};
    g.p[0] = make_linear(s, D, o, nullptr, c, N, N, N, D);
    g.p[0].nb = nb; g.p[0].sA = (long long)N * D; g.p[0].sW = (long long)N * D; g.p[0].sC = (long long)N * N;
    g.count = 1;
    return launch_gemm(g, st);
  };
  const size_t img_bytes = sizeof(float) * (size_t)N * N;
  if (tc && !conv && raw == importance && !pair_feat && get_option(OPT_PPN_FUSED_TOPK) &&
      pair_topk_fused_supported(N, D, K)) {
    // micro-benchmark 5a at scale: ONE pass over HBM -- the top-k runs on the tensor-memory accumulator inside the
    // pair-matrix kernel; images it flags (candidate overflow: constant / adversarial matrices) go to the exact kernel
    PN_TRY(launch_pair_topk_fused(S, O, false, importance, topk_idx, sub_pos, obj_pos, redo, B, N, D, K, st));
    return launch_topk_pairs(importance, topk_idx, sub_pos, obj_pos, query, nullptr, B, N, K, st, redo);
  }
  if (!conv && raw == importance && (size_t)B * img_bytes > PPN_L2_CHUNK_BYTES) {
    // large batches (micro-benchmark 5a): walk the batch in chunks whose pair matrices stay resident in the 126 MB
    // L2 between the kernel that writes them and the top-k kernel that reads them back
    int cb = (int)(PPN_L2_CHUNK_BYTES / img_bytes);
    // the top-k kernel runs one CTA per image: never hand it less than one wave of SMs (N = 400: 75 images fit the L2
    // budget, which left half of the 148 SMs idle), even if part of the chunk then spills to HBM
    const int num_sms = sm_count();
    cb = cb < num_sms ? num_sms : cb;
    for (int b0 = 0; b0 < B; b0 += cb) {
      const int nb = B - b0 < cb ? B - b0 : cb;
      PN_TRY(pair_matrix(S + (size_t)b0 * N * D, O + (size_t)b0 * N * D, importance + (size_t)b0 * N * N, nb));
      PN_TRY(launch_topk_pairs(importance + (size_t)b0 * N * N, topk_idx ? topk_idx + (size_t)b0 * K : nullptr,
                               sub_pos + (size_t)b0 * K, obj_pos + (size_t)b0 * K, query + (size_t)b0 * N * D,
                               pair_feat ? pair_feat + (size_t)b0 * 2 * K * D : nullptr, nb, N, K, st));
    }
    return 0;
  }
  PN_TRY(pair_matrix(S, O, raw, B));
  if (conv) {
    PN_TRY(launch_conv_tiny(raw, conv, importance, B, N, conv_ws, conv_bytes, st));
  } else if (raw != importance) {
    PN_TRY(copy_async(importance, raw, sizeof(float) * B * N * N, st));
  }
  return launch_topk_pairs(importance, topk_idx, sub_pos, obj_pos, query, pair_feat, B, N, K, st);
}

static size_t ppn_bytes(int B, int N, int K, int mid) {
  Workspace ws(nullptr, 0);
  const int M = B * N;
  for (int i = 0; i < 4; ++i) ws.take<float>((size_t)2 * M * D);
  ws.take<float>((size_t)B * N * N);
  if (mid > 0) ws.take<char>(conv_tiny_workspace_bytes(B, N, mid));  // mid_channels <= 0: no ConvTiny (config 5a)
  ws.take<int>((size_t)B);
  return ws.off + 1024;
}

// ------------------------------------------------------------------------------------------------
struct PreparedRel {
  PreparedLayer L[PN_MAX_LAYERS];
  float *ck_hi, *ck_lo, *cv_hi, *cv_lo;  // [nl*256,256]: cross-attention Wk / Wv of all layers, concatenated
  float *bk, *bv;                        // [nl*256]
  float *cls_hi, *cls_lo;                // [num_rel_cls,256]
};
static void carve_rel(Workspace& ws, const PnRelWeights* w, PreparedRel& p) {
  const int nl = w->num_layers;
  for (int l = 0; l < nl; ++l) carve_layer(ws, w->ffn_dims, p.L[l]);
  p.ck_hi = ws.take<float>((size_t)nl * D * D); p.ck_lo = ws.take<float>((size_t)nl * D * D);
  p.cv_hi = ws.take<float>((size_t)nl * D * D); p.cv_lo = ws.take<float>((size_t)nl * D * D);
  p.bk = ws.take<float>((size_t)nl * D); p.bv = ws.take<float>((size_t)nl * D);
  p.cls_hi = ws.take<float>((size_t)w->num_rel_cls * D); p.cls_lo = ws.take<float>((size_t)w->num_rel_cls * D);
}
constexpr int CHAIN_AUTO_MIN_BATCH = 8;  // PN_OPT_FUSED_CHAIN = 1: cluster chain from this many images per call (measured)
// tensor-core key side available: prepared blob + tcgen05 enabled
static bool rel_prepared_ok(const PnRelWeights* w) {
  return w->prepared && get_option(OPT_TENSOR_CORES) && get_option(OPT_FUSED_CHAIN) != 0;
}
static bool rel_chain_ok(const PnRelWeights* w, int B) {
  const int opt = get_option(OPT_FUSED_CHAIN);
  return rel_prepared_ok(w) && (opt >= 2 || (opt == 1 && B >= CHAIN_AUTO_MIN_BATCH)) && w->num_layers <= CHAIN_MAX_LAYERS &&
         w->ffn_dims % 1024 == 0 && w->num_rel_cls <= 128 && w->num_rel_cls % 4 == 0;
}

// key-side operands of ALL relation layers (pair features do not change across layers, pairnet_head.py:365-376)
struct RelKeyBufs {
  float *pk, *p_hi, *p_lo, *kc_hi, *kc_lo, *vtc_hi, *vtc_lo;
};
static void rel_take_keys(Workspace& ws, int B, int K2, int nl, RelKeyBufs& b) {
  const size_t Mk = (size_t)B * K2;
  const size_t ldvc = (size_t)round_up(K2, 4);
  b.pk = ws.take<float>(Mk * D); b.p_hi = ws.take<float>(Mk * D); b.p_lo = ws.take<float>(Mk * D);
  b.kc_hi = ws.take<float>(Mk * nl * D); b.kc_lo = ws.take<float>(Mk * nl * D);
  b.vtc_hi = ws.take<float>((size_t)B * nl * D * ldvc); b.vtc_lo = ws.take<float>((size_t)B * nl * D * ldvc);
}
// 4 launches on the tcgen05 GEMM: K [B*K2, nl*256] and V^T [B*nl*256, ldvc] of every layer, emitted as TF32 hi/lo pairs
static int rel_key_side(const PnRelWeights* w, const PreparedRel& P, const float* pair_feat, const RelKeyBufs& b, int B,
                        int K2, cudaStream_t st) {
  const int nl = w->num_layers;
  const int Mk = B * K2, ldk = nl * D;
  const int ldvc = (int)round_up(K2, 4);
  PN_TRY(launch_add_rows(pair_feat, w->rel_query_embed2, b.pk, B, K2, st));
  PN_TRY(launch_split_tf32(pair_feat, b.p_hi, b.p_lo, (size_t)Mk * D, st));
  {  // K = (pair + key_pos) [Wk_0; ..; Wk_nl-1]^T + bk   (column chunks <= 1024: bias staging limit of the GEMM)
    UmmaOperand o[4];
    int n = 0;
    for (int n0 = 0; n0 < ldk; n0 += 1024) {
      const int nc = ldk - n0 < 1024 ? ldk - n0 : 1024;
      UmmaOperand k{b.pk, nullptr, D, P.ck_hi + (size_t)n0 * D, P.ck_lo + (size_t)n0 * D, D, P.bk + n0, b.kc_hi + n0, ldk, Mk,
                    nc, D};
      k.C_lo = b.kc_lo + n0;
      k.a_is_raw = 1;
      o[n++] = k;
      if (n == 4 || n0 + 1024 >= ldk) { PN_TRY(launch_umma_gemm(o, n, 3, st)); n = 0; }
    }
  }
  {  // V^T per image = [Wv_0; ..] pair_b^T + bv (weights as the A operand, per-row bias, row chunks <= 1024)
    UmmaOperand o[4];
    int n = 0;
    for (int bi = 0; bi < B; ++bi)
      for (int m0 = 0; m0 < ldk; m0 += 1024) {
        const int mc = ldk - m0 < 1024 ? ldk - m0 : 1024;
        UmmaOperand v{P.cv_hi + (size_t)m0 * D, P.cv_lo + (size_t)m0 * D, D, b.p_hi + (size_t)bi * K2 * D,
                      b.p_lo + (size_t)bi * K2 * D, D, P.bv + m0, b.vtc_hi + ((size_t)bi * ldk + m0) * ldvc, ldvc, mc, K2, D};
        v.C_lo = b.vtc_lo + ((size_t)bi * ldk + m0) * ldvc;
        v.bias_per_row = 1;
        o[n++] = v;
        const bool last = bi == B - 1 && m0 + 1024 >= ldk;
        if (n == 4 || last) { PN_TRY(launch_umma_gemm(o, n, 3, st)); n = 0; }
      }
  }
  return 0;
}

struct RelBufs {
  float *x, *xpos, *pk, *Kall, *Vall, *chain_scratch;
  RelKeyBufs keys;
  LayerScratch ls;
};
// one carve-up for all three variants (per-op FFMA keys, per-op tensor-core keys, cluster chain): sized for the largest
static void rel_take(Workspace& ws, int B, int R, int K2, int nl, int ffn, RelBufs& b) {
  const int M = B * R, Mk = B * K2;
  b.x = ws.take<float>((size_t)M * D);
  b.xpos = ws.take<float>((size_t)M * D);
  b.pk = ws.take<float>((size_t)Mk * D);
  b.Kall = ws.take<float>((size_t)nl * Mk * D);
  b.Vall = ws.take<float>((size_t)nl * Mk * D);
  rel_take_keys(ws, B, K2, nl, b.keys);
  b.chain_scratch = ws.take<float>(chain_scratch_floats(B, R, ffn));
  size_t mb = mha_workspace_bytes(B, R, K2);
  size_t mb2 = mha_workspace_bytes(B, R, R);
  size_t mb3 = fa_workspace_bytes(B, R, K2);
  mb = mb > mb2 ? mb : mb2;
  layer_scratch_take(ws, b.ls, M, ffn, mb > mb3 ? mb : mb3);
}

// Relation Fusion decoder + relation classifier (pairnet_head.py:353-378).
//   prepared weights + tensor cores: the key side of all layers runs on the tcgen05 GEMM (4 launches), then either
//     * ONE cluster launch of the fused chain kernel (chain.cu) for the six layers + classifier  (PN_OPT_FUSED_CHAIN = 2,
//       or = 1 with >= 8 images: the chain uses 8 SMs per image, so it wins once a batch fills the GPU), or
//     * per-op kernels on all SMs with the cross attention on tcgen05 (fa_umma.cu) -- the fastest path at bs = 2;
//   otherwise: exact-fp32 / warp-MMA per-op kernels as in round 1.
static int rel_forward(const PnRelWeights* w, const float* pair_feat, float* rel_preds, float* rel_feat_out, int B,
                       int K2, Workspace& ws, cudaStream_t st) {
  NvtxRange nvtx("pn::relation_fusion (pairnet_head.py:353-378)");
  PN_REQUIRE(w && pair_feat && rel_preds, PN_ERR_BAD_ARG, "relation_fusion: null pointer");
  const int R = w->num_rel_queries, nl = w->num_layers, ffn = w->ffn_dims;
  PN_REQUIRE(R > 0 && K2 > 0 && B > 0 && nl >= 1 && nl <= PN_MAX_LAYERS, PN_ERR_BAD_ARG, "relation_fusion: bad sizes");
  const int M = B * R, Mk = B * K2;
  RelBufs b;
  rel_take(ws, B, R, K2, nl, ffn, b);
  PN_REQUIRE(ws.ok() && !ws.dry, PN_ERR_WORKSPACE, "relation_fusion: workspace too small");
  const bool prep = rel_prepared_ok(w);
  PreparedRel P{};
  if (prep) {
    Workspace pw(const_cast<void*>(w->prepared), (size_t)1 << 60);
    carve_rel(pw, w, P);
    PN_TRY(rel_key_side(w, P, pair_feat, b.keys, B, K2, st));
  }
  const int ldvc = (int)round_up(K2, 4);
  if (rel_chain_ok(w, B)) {
    ChainArgs a{};
    for (int l = 0; l < nl; ++l) chain_layer_fill(a.layers[l], w->layers[l], P.L[l]);
    a.x = b.x; a.scratch = b.chain_scratch;
    a.kc_hi = b.keys.kc_hi; a.kc_lo = b.keys.kc_lo; a.vtc_hi = b.keys.vtc_hi; a.vtc_lo = b.keys.vtc_lo;
    a.init_feat = w->rel_query_feat; a.qpos = w->rel_query_embed;
    a.cls_hi = P.cls_hi; a.cls_lo = P.cls_lo; a.cls_b = w->rel_cls_embed.b; a.cls_out = rel_preds;
    a.B = B; a.R = R; a.Nk = K2; a.nl = nl; a.ffn = ffn; a.ncls = w->num_rel_cls; a.ldvc = ldvc; a.has_cross_attn = 1;
    PN_TRY(launch_decoder_chain(a, st));
    if (rel_feat_out) PN_TRY(copy_async(rel_feat_out, b.x, sizeof(float) * (size_t)M * D, st));
    return 0;
  }

  PN_TRY(launch_bcast_rows(w->rel_query_feat, w->rel_query_embed, b.x, b.xpos, B, R, st));
  if (!prep) {
    PN_TRY(launch_add_rows(pair_feat, w->rel_query_embed2, b.pk, B, K2, st));
    // K/V of every layer up front: pair_feat does not change across layers (pairnet_head.py:365-376)
    for (int l0 = 0; l0 < nl; l0 += GEMM_MAX_PROBS / 2) {
      GemmBatch g{};
      int c = 0;
      for (int l = l0; l < nl && c + 2 <= GEMM_MAX_PROBS; ++l) {
        const PnMHA& a = w->layers[l].cross_attn;
        g.p[c++] = make_linear(b.pk, D, a.in_proj_w + (size_t)D * D, a.in_proj_b + D, b.Kall + (size_t)l * Mk * D, D, Mk, D, D);
        g.p[c++] = make_linear(pair_feat, D, a.in_proj_w + (size_t)2 * D * D, a.in_proj_b + 2 * D,
                               b.Vall + (size_t)l * Mk * D, D, Mk, D, D);
      }
      g.count = c;
      PN_TRY(launch_gemm(g, st));
    }
  }
  for (int l = 0; l < nl; ++l) {
    // tensor-core cross attention reads layer l's slice of the key-side operands in place
    FaKV fa{b.keys.kc_hi, b.keys.kc_lo, b.keys.vtc_hi, b.keys.vtc_lo, ldvc};
    fa.ldk = nl * D; fa.k_col0 = l * D; fa.vt_img_rows = nl * D; fa.vt_row0 = l * D;
    PN_TRY(decoder_layer(w->layers[l], ffn, b.x, b.xpos, w->rel_query_embed, B, R, b.Kall + (size_t)l * Mk * D,
                         b.Vall + (size_t)l * Mk * D, K2, nullptr, 0, nullptr, nullptr, nullptr, b.ls, st, nullptr, nullptr,
                         prep ? &fa : nullptr));
  }
  PN_TRY(linear1(b.x, D, w->rel_cls_embed, rel_preds, w->num_rel_cls, M, w->num_rel_cls, D, 0, st));
  if (rel_feat_out) PN_TRY(copy_async(rel_feat_out, b.x, sizeof(float) * M * D, st));
  return 0;
}

}  // namespace pn

// ================================================================================================
// extern "C"
// ================================================================================================
using namespace pn;

extern "C" {

int pn_version(void) { return PN_VERSION; }
int pn_set_option(int key, int value) {
  PN_REQUIRE(key >= 0 && key < OPT_COUNT, PN_ERR_BAD_ARG, "pn_set_option: unknown key %d", key);
  g_options[key] = value;
  return 0;
}
int pn_get_option(int key) { return get_option(key); }
const char* pn_last_error_string(void) { return g_err; }
int pn_last_launch_count(void) { return g_launches; }
int pn_debug_chain_timing(unsigned long long* device_buf, int capacity) {
  chain_set_timing(device_buf, capacity);
  return 0;
}

int pn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { set_error("pn_device_info: %s", cudaGetErrorString(e)); return (int)e; }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}

int pn_sine_posenc(float* pos, int h, int w, pn_stream_t stream) { return launch_sine_posenc(pos, h, w, as_stream(stream)); }

int pn_level_prep(const float* mem, const float* level_embed, const float* pos, float* x, float* xp, int B, int hw,
                  pn_stream_t stream) {
  return launch_level_prep(mem, level_embed, pos, x, xp, B, hw, as_stream(stream));
}

int pn_level_prep_tokens(const float* mem, long long batch_stride, const float* level_embed, const float* pos, float* x,
                         float* xp, int B, int hw, pn_stream_t stream) {
  return launch_level_prep_tokens(mem, batch_stride, level_embed, pos, x, xp, B, hw, as_stream(stream));
}

int pn_mask_feature_resize(const float* mask_feature, float* out, int B, int H, int W, int h, int w, int ldo,
                           pn_stream_t stream) {
  return launch_mask_feature_resize(mask_feature, out, B, H, W, h, w, ldo, as_stream(stream));
}

int pn_attn_mask_bits(const float* E, const float* F, uint32_t* bits, int* rowany, int B, int N, int hw, int ldf,
                      pn_stream_t stream) {
  return launch_gemm_nmajor_maskbits(E, F, bits, rowany, B, N, hw, ldf, as_stream(stream));
}

int pn_mask_feature_resize_tokens(const float* mask_feature, float* out, int B, int H, int W, int h, int w,
                                  pn_stream_t stream) {
  return launch_mask_feature_resize_tokens(mask_feature, out, B, H, W, h, w, as_stream(stream));
}
int pn_nchw_to_tokens(const float* src, float* dst, int B, int HW, pn_stream_t stream) {
  return launch_nchw_to_tokens(src, dst, B, HW, as_stream(stream));
}
size_t pn_mask_tc_workspace_bytes(int B, int N) { return 2 * (((size_t)B * N * D * sizeof(float) + 255) & ~size_t(255)) + 256; }
int pn_attn_mask_bits_tc(const float* E, const float* F_tokens, uint32_t* bits, int* rowany, int B, int N, int hw,
                         int words, void* wsp, size_t ws_bytes, pn_stream_t stream) {
  PN_REQUIRE(E && F_tokens && bits && rowany && wsp, PN_ERR_BAD_ARG, "attn_mask_bits_tc: null pointer");
  PN_REQUIRE(words * 32 >= hw && hw > 0, PN_ERR_BAD_ARG, "attn_mask_bits_tc: words*32 must cover hw");
  Workspace ws(wsp, ws_bytes);
  float* eh = ws.take<float>((size_t)B * N * D);
  float* el = ws.take<float>((size_t)B * N * D);
  PN_REQUIRE(ws.ok() && eh && el, PN_ERR_WORKSPACE, "attn_mask_bits_tc: workspace too small");
  return maskbits_tc(E, eh, el, F_tokens, bits, rowany, B, N, hw, words, as_stream(stream));
}
int pn_mask_pred_tc(const float* E, const float* F_tokens, float* mask_pred, int B, int N, int HW, void* wsp,
                    size_t ws_bytes, pn_stream_t stream) {
  PN_REQUIRE(E && F_tokens && mask_pred && wsp && B > 0 && N > 0 && HW > 0, PN_ERR_BAD_ARG, "mask_pred_tc: bad args");
  Workspace ws(wsp, ws_bytes);
  float* eh = ws.take<float>((size_t)B * N * D);
  float* el = ws.take<float>((size_t)B * N * D);
  PN_REQUIRE(ws.ok() && eh && el, PN_ERR_WORKSPACE, "mask_pred_tc: workspace too small");
  return mask_pred_tc(E, eh, el, F_tokens, mask_pred, B, N, HW, as_stream(stream));
}

int pn_mask_pred(const float* E, const float* F, float* mask_pred, int B, int N, int HW, pn_stream_t stream) {
  PN_REQUIRE(E && F && mask_pred && B > 0 && N > 0 && HW > 0, PN_ERR_BAD_ARG, "mask_pred: bad args");
  GemmProb g = make_linear(E, D, F, nullptr, mask_pred, HW, N, HW, D);
  g.ldw = HW;
  g.nb = B; g.sA = (long long)N * D; g.sW = (long long)D * HW; g.sC = (long long)N * HW;
  return launch_gemm_nmajor_store(g, as_stream(stream));
}

int pn_linear(const float* x, int ldx, const float* w, const float* b, const float* resid, float* y, int ldy, int M,
              int N, int K, int relu, pn_stream_t stream) {
  GemmBatch g{};
  g.p[0] = make_linear(x, ldx, w, b, y, ldy, M, N, K, relu, resid, ldy);
  g.count = 1;
  return launch_gemm(g, as_stream(stream));
}

size_t pn_linear_tc_workspace_bytes(int M, int N, int K) {
  Workspace ws(nullptr, 0);
  ws.take<float>((size_t)M * K); ws.take<float>((size_t)M * K);
  ws.take<float>((size_t)N * K); ws.take<float>((size_t)N * K);
  return ws.off + 256;
}

int pn_linear_tc(const float* x, int ldx, const float* w, const float* b, float* y, int ldy, int M, int N, int K,
                 int passes, void* ws, size_t ws_bytes, pn_stream_t stream) {
  PN_REQUIRE(x && w && y && ws, PN_ERR_BAD_ARG, "linear_tc: null pointer");
  PN_REQUIRE(ldx == K, PN_ERR_UNSUPPORTED, "linear_tc: x must be dense (ldx == K)");
  Workspace W(ws, ws_bytes);
  float* xh = W.take<float>((size_t)M * K); float* xl = W.take<float>((size_t)M * K);
  float* wh = W.take<float>((size_t)N * K); float* wl = W.take<float>((size_t)N * K);
  PN_REQUIRE(W.ok() && xh && xl && wh && wl, PN_ERR_WORKSPACE, "linear_tc: workspace too small");
  cudaStream_t st = as_stream(stream);
  PN_TRY(launch_split_tf32(w, wh, wl, (size_t)N * K, st));
  if (passes == 3 && get_option(OPT_UMMA_RAW_A)) {  // A split in-kernel (TMEM), no pre-pass over x
    UmmaOperand o{x, nullptr, K, wh, wl, K, b, y, ldy, M, N, K};
    o.a_is_raw = 1;
    return launch_umma_gemm(&o, 1, passes, st);
  }
  PN_TRY(launch_split_tf32(x, xh, xl, (size_t)M * K, st));
  UmmaOperand o{xh, xl, K, wh, wl, K, b, y, ldy, M, N, K};
  return launch_umma_gemm(&o, 1, passes, st);
}

int pn_split_tf32(const float* x, float* hi, float* lo, size_t n, pn_stream_t stream) {
  return launch_split_tf32(x, hi, lo, n, as_stream(stream));
}

int pn_linear_tc_rawa(const float* x, const float* w_hi, const float* w_lo, const float* b, float* y, int ldy, int M,
                      int N, int K, pn_stream_t stream) {
  UmmaOperand o{x, nullptr, K, w_hi, w_lo, K, b, y, ldy, M, N, K};
  o.a_is_raw = 1;
  return launch_umma_gemm(&o, 1, 3, as_stream(stream));
}

int pn_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, size_t n, pn_stream_t stream) {
  return launch_split_bf16(x, hi, lo, n, as_stream(stream));
}

int pn_linear_tc_bf16x3(const float* x, const uint16_t* w_hi, const uint16_t* w_lo, const float* b, float* y, int ldy,
                        int M, int N, int K, pn_stream_t stream) {
  UmmaOperand o{x, nullptr, K, reinterpret_cast<const float*>(w_hi), reinterpret_cast<const float*>(w_lo), K, b, y, ldy, M, N, K};
  o.a_is_raw = 1;
  o.w_bf16 = 1;
  return launch_umma_gemm(&o, 1, 3, as_stream(stream));
}

int pn_linear_tc_presplit(const float* x_hi, const float* x_lo, const float* w_hi, const float* w_lo, const float* b,
                          float* y, int ldy, int M, int N, int K, int passes, pn_stream_t stream) {
  UmmaOperand o{x_hi, x_lo, K, w_hi, w_lo, K, b, y, ldy, M, N, K};
  return launch_umma_gemm(&o, 1, passes, as_stream(stream));
}

int pn_add_layernorm(const float* x, const float* resid, const float* gamma, const float* beta, float* y, int M,
                     pn_stream_t stream) {
  LnArgs n{};
  n.x = x; n.nparts = 1; n.resid = resid; n.gamma = gamma; n.beta = beta; n.y = y; n.M = M;
  return launch_layernorm(n, as_stream(stream));
}

size_t pn_mha_workspace_bytes(int B, int Nq, int Nk) { return mha_workspace_bytes(B, Nq, Nk); }

int pn_mha_core(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, const uint32_t* mask_bits,
                int mask_words, const int* rowany, float* out, int B, int Nq, int Nk, void* ws, size_t ws_bytes,
                pn_stream_t stream) {
  MhaArgs a{q, ldq, k, ldk, v, ldv, mask_bits, mask_words, rowany, out, B, Nq, Nk};
  return launch_mha(a, ws, ws_bytes, as_stream(stream));
}

static void fa_test_take(Workspace& ws, int B, int Nq, int Nk, float** p) {
  const int ldv = (int)round_up(Nk, 4);
  p[0] = ws.take<float>((size_t)B * Nq * D); p[1] = ws.take<float>((size_t)B * Nq * D);
  p[2] = ws.take<float>((size_t)B * Nk * D); p[3] = ws.take<float>((size_t)B * Nk * D);
  p[4] = ws.take<float>((size_t)B * D * ldv); p[5] = ws.take<float>((size_t)B * D * ldv);
  p[6] = reinterpret_cast<float*>(ws.take<char>(fa_workspace_bytes(B, Nq, Nk)));
}
size_t pn_mha_core_tc_workspace_bytes(int B, int Nq, int Nk) {
  Workspace ws(nullptr, 0);
  float* p[7];
  fa_test_take(ws, B, Nq, Nk, p);
  return ws.off + 256;
}
int pn_mha_core_tc(const float* q, const float* k, const float* v, const uint32_t* mask_bits, int mask_words,
                   const int* rowany, float* out, int B, int Nq, int Nk, void* wsp, size_t ws_bytes,
                   pn_stream_t stream) {
  PN_REQUIRE(q && k && v && out && wsp, PN_ERR_BAD_ARG, "mha_core_tc: null pointer");
  Workspace ws(wsp, ws_bytes);
  float* p[7];
  fa_test_take(ws, B, Nq, Nk, p);
  PN_REQUIRE(ws.ok(), PN_ERR_WORKSPACE, "mha_core_tc: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int ldv = (int)round_up(Nk, 4);
  PN_TRY(launch_split_tf32_scaled(q, p[0], p[1], (size_t)B * Nq * D, ATTN_QSCALE, st));
  PN_TRY(launch_split_tf32(k, p[2], p[3], (size_t)B * Nk * D, st));
  PN_TRY(launch_split_transpose(v, p[4], p[5], B, Nk, ldv, st));
  FaArgs a{p[0], p[1], p[2], p[3], p[4], p[5], ldv, mask_bits, mask_words, rowany, out, B, Nq, Nk};
  return launch_fa_umma(a, p[6], fa_workspace_bytes(B, Nq, Nk), st);
}

size_t pn_m2f_decoder_workspace_bytes(const PnM2FWeights* w, const PnM2FInputs* in) {
  M2FPlan p;
  if (m2f_plan(w, in, p) != 0) return 0;
  Workspace ws(nullptr, 0);
  M2FBuffers b{};
  PnM2FInputs nopos = *in;  // size for the worst case: position tables computed into the workspace
  for (int l = 0; l < PN_MAX_LEVELS; ++l) nopos.pos[l] = nullptr;
  m2f_take(ws, p, &nopos, b);
  return ws.off + 1024;
}

int pn_m2f_decoder_forward(const PnM2FWeights* w, const PnM2FInputs* in, const PnM2FOutputs* out, void* ws,
                           size_t ws_bytes, pn_stream_t stream) {
  g_launches = 0;
  PN_REQUIRE(ws, PN_ERR_WORKSPACE, "m2f: null workspace");
  Workspace W(ws, ws_bytes);
  return m2f_forward(w, in, out, W, as_stream(stream));
}

size_t pn_ppn_workspace_bytes(int B, int N, int K, int mid_channels) { return ppn_bytes(B, N, K, mid_channels); }

int pn_ppn_forward(const float* query, const float* query_obj, const PnMlp3* sub_mlp, const PnMlp3* obj_mlp,
                   const PnConvTiny* conv, float* importance_raw, float* importance, int64_t* topk_idx,
                   int64_t* sub_pos, int64_t* obj_pos, float* pair_feat, int B, int N, int K, void* ws,
                   size_t ws_bytes, pn_stream_t stream) {
  g_launches = 0;
  PN_REQUIRE(ws, PN_ERR_WORKSPACE, "ppn: null workspace");
  Workspace W(ws, ws_bytes);
  return ppn_forward(query, query_obj, sub_mlp, obj_mlp, conv, importance_raw, importance, topk_idx, sub_pos, obj_pos,
                     pair_feat, B, N, K, W, as_stream(stream));
}

size_t pn_ppn_pair_topk_bf16_workspace_bytes(int B) { return sizeof(int) * (size_t)(B > 0 ? B : 0) + 1024; }

int pn_ppn_pair_topk_bf16(const uint16_t* sub_embed, const uint16_t* obj_embed, float* importance, int64_t* topk_idx,
                          int64_t* sub_pos, int64_t* obj_pos, int B, int N, int K, void* ws, size_t ws_bytes,
                          pn_stream_t stream) {
  g_launches = 0;
  PN_REQUIRE(sub_embed && obj_embed && importance && sub_pos && obj_pos && B > 0, PN_ERR_BAD_ARG, "ppn bf16: bad args");
  PN_REQUIRE(ws, PN_ERR_WORKSPACE, "ppn bf16: null workspace");
  PN_REQUIRE(get_option(OPT_TENSOR_CORES), PN_ERR_UNSUPPORTED, "ppn bf16: the bf16 path exists on tcgen05 only");
  PN_REQUIRE(pair_topk_fused_supported(N, D, K, true), PN_ERR_UNSUPPORTED, "ppn bf16: N=%d K=%d unsupported", N, K);
  Workspace W(ws, ws_bytes);
  int* redo = W.take<int>((size_t)B);
  PN_REQUIRE(W.ok() && !W.dry, PN_ERR_WORKSPACE, "ppn bf16: workspace too small");
  cudaStream_t st = as_stream(stream);
  PN_TRY(launch_pair_topk_fused(sub_embed, obj_embed, true, importance, topk_idx, sub_pos, obj_pos, redo, B, N, D, K, st));
  // images the fused kernel flagged (candidate overflow: constant / adversarial matrices) go to the exact kernel
  return launch_topk_pairs(importance, topk_idx, sub_pos, obj_pos, nullptr, nullptr, B, N, K, st, redo);
}

int pn_conv_tiny(const float* x, const PnConvTiny* conv, float* y, int B, int N, void* ws, size_t ws_bytes,
                 pn_stream_t stream) {
  return launch_conv_tiny(x, conv, y, B, N, ws, ws_bytes, as_stream(stream));
}

int pn_topk_pairs(const float* importance, int64_t* topk_idx, int64_t* sub_pos, int64_t* obj_pos, const float* query,
                  float* pair_feat, int B, int N, int K, pn_stream_t stream) {
  return launch_topk_pairs(importance, topk_idx, sub_pos, obj_pos, query, pair_feat, B, N, K, as_stream(stream));
}

size_t pn_relation_fusion_workspace_bytes(int B, int R, int K2, int ffn_dims) {
  Workspace ws(nullptr, 0);
  RelBufs rb;
  rel_take(ws, B, R, K2, PN_MAX_LAYERS, ffn_dims, rb);
  return ws.off + 1024;
}

size_t pn_m2f_prepared_bytes(const PnM2FWeights* w) {
  if (!w || w->num_layers < 1 || w->num_layers > PN_MAX_LAYERS) return 0;
  Workspace ws(nullptr, 0);
  PreparedM2F p;
  carve_m2f(ws, w, p);
  return ws.off + 256;
}

int pn_m2f_prepare(const PnM2FWeights* w, void* prepared, size_t bytes, pn_stream_t stream) {
  PN_REQUIRE(w && prepared, PN_ERR_BAD_ARG, "m2f_prepare: null argument");
  PN_REQUIRE(w->num_layers >= 1 && w->num_layers <= PN_MAX_LAYERS, PN_ERR_BAD_ARG, "m2f_prepare: bad layer count");
  PN_REQUIRE(((uintptr_t)prepared & 255) == 0, PN_ERR_BAD_ARG, "m2f_prepare: buffer must be 256-byte aligned");
  Workspace ws(prepared, bytes);
  PreparedM2F p;
  carve_m2f(ws, w, p);
  PN_REQUIRE(ws.ok(), PN_ERR_WORKSPACE, "m2f_prepare: buffer too small (%zu < %zu)", bytes, ws.off);
  cudaStream_t st = as_stream(stream);
  for (int l = 0; l < w->num_layers; ++l) PN_TRY(prepare_layer(w->layers[l], w->ffn_dims, p.L[l], st));
  for (int i = 0; i < 3; ++i)
    PN_TRY(launch_split_tf32(w->mask_embed.l[i].w, p.me_hi[i], p.me_lo[i], (size_t)D * D, st));
  return 0;
}

size_t pn_rel_prepared_bytes(const PnRelWeights* w) {
  if (!w || w->num_layers < 1 || w->num_layers > PN_MAX_LAYERS) return 0;
  Workspace ws(nullptr, 0);
  PreparedRel p;
  carve_rel(ws, w, p);
  return ws.off + 256;
}

int pn_rel_prepare(const PnRelWeights* w, void* prepared, size_t bytes, pn_stream_t stream) {
  PN_REQUIRE(w && prepared, PN_ERR_BAD_ARG, "rel_prepare: null argument");
  PN_REQUIRE(w->num_layers >= 1 && w->num_layers <= PN_MAX_LAYERS, PN_ERR_BAD_ARG, "rel_prepare: bad layer count");
  PN_REQUIRE(((uintptr_t)prepared & 255) == 0, PN_ERR_BAD_ARG, "rel_prepare: buffer must be 256-byte aligned");
  Workspace ws(prepared, bytes);
  PreparedRel p;
  carve_rel(ws, w, p);
  PN_REQUIRE(ws.ok(), PN_ERR_WORKSPACE, "rel_prepare: buffer too small (%zu < %zu)", bytes, ws.off);
  cudaStream_t st = as_stream(stream);
  for (int l = 0; l < w->num_layers; ++l) {
    PN_TRY(prepare_layer(w->layers[l], w->ffn_dims, p.L[l], st));
    // [Wk;Wv] = in_proj rows 256..767: concatenate per kind across layers (already split)
    const size_t blk = (size_t)D * D;
    PN_TRY(copy_async(p.ck_hi + l * blk, p.L[l].cin_hi + blk, blk * 4, st));
    PN_TRY(copy_async(p.ck_lo + l * blk, p.L[l].cin_lo + blk, blk * 4, st));
    PN_TRY(copy_async(p.cv_hi + l * blk, p.L[l].cin_hi + 2 * blk, blk * 4, st));
    PN_TRY(copy_async(p.cv_lo + l * blk, p.L[l].cin_lo + 2 * blk, blk * 4, st));
    PN_TRY(copy_async(p.bk + (size_t)l * D, w->layers[l].cross_attn.in_proj_b + D, D * 4, st));
    PN_TRY(copy_async(p.bv + (size_t)l * D, w->layers[l].cross_attn.in_proj_b + 2 * D, D * 4, st));
  }
  return launch_split_tf32(w->rel_cls_embed.w, p.cls_hi, p.cls_lo, (size_t)w->num_rel_cls * D, st);
}

int pn_relation_fusion_forward(const PnRelWeights* w, const float* pair_feat, float* rel_preds, float* rel_feat_out,
                               int B, int K2, void* ws, size_t ws_bytes, pn_stream_t stream) {
  g_launches = 0;
  PN_REQUIRE(ws, PN_ERR_WORKSPACE, "relation_fusion: null workspace");
  Workspace W(ws, ws_bytes);
  return rel_forward(w, pair_feat, rel_preds, rel_feat_out, B, K2, W, as_stream(stream));
}

int pn_gather_rows(const float* src, const int64_t* idx, float* dst, int B, int Nsrc, int R, long long L,
                   pn_stream_t stream) {
  return launch_gather_rows(src, idx, dst, B, Nsrc, R, L, as_stream(stream));
}

size_t pn_head_workspace_bytes(const PnHeadWeights* w, const PnM2FInputs* in) {
  if (!w || !in) return 0;
  const size_t a = pn_m2f_decoder_workspace_bytes(&w->m2f, in);
  if (a == 0) return 0;
  const int B = in->B, N = w->m2f.num_queries, R = w->rel.num_rel_queries;
  const size_t b = ppn_bytes(B, N, R, w->update_importance.mid_channels);
  const size_t c = pn_relation_fusion_workspace_bytes(B, R, 2 * R, w->rel.ffn_dims);
  // stage scratch is reused (max), persistent taps are extra
  size_t stage = a > b ? a : b;
  stage = stage > c ? stage : c;
  const bool need_ftok = get_option(OPT_TENSOR_CORES) && get_option(OPT_MASK_TC) && !in->mask_features_token_major;
  const size_t persist = ((size_t)B * N * D * 4 + 255) * 7 + ((size_t)B * 2 * R * D * 4 + 255) + 1024 +
                         (need_ftok ? ((size_t)B * in->H4 * in->W4 * D * 4 + 255) : 0);  // token-major mask_features copy
  return stage + persist;
}

int pn_head_forward(const PnHeadWeights* w, const PnM2FInputs* in, const PnHeadOutputs* out, void* ws,
                    size_t ws_bytes, pn_stream_t stream) {
  g_launches = 0;
  PN_REQUIRE(w && in && out && ws, PN_ERR_BAD_ARG, "head: null argument");
  PN_REQUIRE(out->cls && out->mask && out->importance && out->rel && out->sub_pos && out->obj_pos, PN_ERR_BAD_ARG,
             "head: required output pointer is null");
  cudaStream_t st = as_stream(stream);
  const int B = in->B, N = w->m2f.num_queries, R = w->rel.num_rel_queries, K = R;
  const int HW4 = in->H4 * in->W4;
  PN_REQUIRE(ws_bytes >= pn_head_workspace_bytes(w, in), PN_ERR_WORKSPACE, "head: workspace too small (%zu < %zu)",
             ws_bytes, pn_head_workspace_bytes(w, in));
  // persistent (cross-stage) buffers first, then per-stage scratch that each stage re-carves
  Workspace P(ws, ws_bytes);
  float* query = out->query_out ? out->query_out : P.take<float>((size_t)B * N * D);
  float* pair = out->pair_feat ? out->pair_feat : P.take<float>((size_t)B * 2 * K * D);
  TailCtx tail{};
  tail.xn = P.take<float>((size_t)B * N * D);
  tail.e1 = P.take<float>((size_t)B * N * D);
  tail.e2 = P.take<float>((size_t)B * N * D);
  tail.e = P.take<float>((size_t)B * N * D);
  tail.eh = P.take<float>((size_t)B * N * D);
  tail.el = P.take<float>((size_t)B * N * D);
  const bool need_ftok = get_option(OPT_TENSOR_CORES) && get_option(OPT_MASK_TC) && !in->mask_features_token_major;
  tail.ftok = need_ftok ? P.take<float>((size_t)B * HW4 * D) : nullptr;
  char* stage = (char*)ws + P.off;
  const size_t stage_bytes = ws_bytes - P.off;

  {
    PnM2FOutputs mo{};
    mo.query_out = query; mo.cls_pred = out->cls; mo.mask_pred = out->mask;
    mo.query_trace = out->query_trace; mo.mask_trace = out->mask_trace; mo.trace_words = out->trace_words;
    Workspace W(stage, stage_bytes);
    PN_TRY(m2f_forward(&w->m2f, in, &mo, W, st, &tail));
  }
  {
    Workspace W(stage, stage_bytes);
    PN_TRY(ppn_forward(query, nullptr, &w->sub_query_update, &w->obj_query_update, &w->update_importance,
                       out->importance_raw, out->importance, nullptr, out->sub_pos, out->obj_pos, pair, B, N, K, W, st));
  }
  {
    Workspace W(stage, stage_bytes);
    PN_TRY(rel_forward(&w->rel, pair, out->rel, out->rel_feat, B, 2 * K, W, st));
  }
  if (tail.deferred) PN_TRY(side_wait(st, tail.done));  // join: cls / mask are complete
  // row 11 (pairnet_head.py:380-403)
  NvtxRange nvtx("pn::output_gathers (pairnet_head.py:380-403)");
  if (out->sub) PN_TRY(launch_gather_rows(out->cls, out->sub_pos, out->sub, B, N, K, w->m2f.num_cls, st));
  if (out->obj) PN_TRY(launch_gather_rows(out->cls, out->obj_pos, out->obj, B, N, K, w->m2f.num_cls, st));
  if (out->sub_seg) PN_TRY(launch_gather_rows(out->mask, out->sub_pos, out->sub_seg, B, N, K, HW4, st));
  if (out->obj_seg) PN_TRY(launch_gather_rows(out->mask, out->obj_pos, out->obj_seg, B, N, K, HW4, st));
  return 0;
}

}  // extern "C"
