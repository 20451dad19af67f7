// Shared helpers for the pairnet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pairnet_b200.h"

namespace pn {

constexpr int D = PN_EMBED_DIMS;  // 256
constexpr int HD = PN_HEAD_DIM;   // 32
constexpr int NH = D / HD;        // 8

// ---- error plumbing (thread-local last error string + launch counter) -------------------------
void set_error(const char* fmt, ...);
void count_launch();
int check_launch(const char* what);  // cudaGetLastError -> 0 / cudaError_t, records the string

#define PN_REQUIRE(cond, code, ...)            \
  do {                                         \
    if (!(cond)) {                             \
      pn::set_error(__VA_ARGS__);              \
      return (code);                           \
    }                                          \
  } while (0)

#define PN_TRY(expr)                 \
  do {                               \
    int _pn_rc = (expr);             \
    if (_pn_rc != 0) return _pn_rc;  \
  } while (0)

static inline cudaStream_t as_stream(pn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
// per-DEVICE caches (a process may drive several GPUs from several threads): SM count, and "function attributes set" flags
constexpr int PN_MAX_DEVICES = 64;
static inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < PN_MAX_DEVICES) ? dev : 0;
}
int sm_count();  // SMs of the current device (api.cu)
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------
// The query-side chain of the head is ~230 dependent launches of 3-10 us kernels: launch latency, not work, bounds it.
// Kernels of that chain are launched with cudaLaunchAttributeProgrammaticStreamSerialization (under stream capture: a
// programmatic graph edge), so kernel N+1 is scheduled and runs its prologue while kernel N is still executing.
// Contract: every kernel launched through `launch_pdl` executes `pdl_wait()` before it touches anything another kernel
// may have written or may still read (it blocks until the preceding grid has completed and flushed), and only THEN
// `pdl_trigger()` -- so when a dependent starts early, everything older than its immediate predecessor is complete.
// Code before `pdl_wait()` may only read data that no kernel of the forward writes (model weights, prepared blobs).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args);
#endif

// ---- bump allocator over the caller's workspace ------------------------------------------------
struct Workspace {
  char* base;
  size_t cap;
  size_t off;
  bool dry;  // dry run: only measure
  Workspace(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0), dry(p == nullptr) {}
  template <typename T>
  T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    size_t o = off;
    off += bytes;
    if (dry) return nullptr;
    if (off > cap) return nullptr;
    return reinterpret_cast<T*>(base + o);
  }
  bool ok() const { return dry || off <= cap; }
};

// ---- grouped GEMM descriptors ------------------------------------------------------------------
// C[b][m,n] = act( sum_k A[b][m,k] * Bop[b] + bias[n] ) + resid[m,n]
//   B_KMAJOR  (linear):  Bop = W[n,k]   (W row-major [N,K], torch Linear layout)
//   B_NMAJOR  (einsum):  Bop = F[k,n]   (F row-major [K,N])
struct GemmProb {
  const float* A;
  const float* W;
  const float* bias;
  const float* resid;
  float* C;
  int M, N, K;
  int lda, ldw, ldc, ldr;
  int relu;
  int nb;                       // strided batch count (>=1)
  long long sA, sW, sC, sR;     // batch strides (elements)
  int splits;                   // split-K (>=1); partial s goes to C + s*split_stride, no epilogue
  long long split_stride;
  int w_static;                 // W is a model weight (no kernel of the forward writes it): may be fetched before pdl_wait()
};
constexpr int GEMM_MAX_PROBS = 12;
struct GemmBatch {
  GemmProb p[GEMM_MAX_PROBS];
  int count;
};

GemmProb make_linear(const float* A, int lda, const float* W, const float* bias, float* C, int ldc,
                     int M, int N, int K, int relu = 0, const float* resid = nullptr, int ldr = 0);

// launchers (all return 0 / error code)
int launch_gemm(const GemmBatch& batch, cudaStream_t st);                      // B K-major
int launch_gemm_nmajor_store(const GemmProb& p, cudaStream_t st);              // einsum, fp32 store
bool skinny_gemm_ok(const GemmBatch& batch);                                   // skinny.cu
int launch_skinny_gemm(const GemmBatch& batch, int maxM, int maxN, int nz, cudaStream_t st);
int launch_gemm_nmajor_maskbits(const float* E, const float* F, uint32_t* bits, int* rowany, int B,
                                int N, int hw, int ldf, cudaStream_t st);

// tcgen05 (TMA + UMMA kind::tf32) GEMM: C = A . W^T + bias with pre-split hi/lo operands (umma_gemm.cu)
struct UmmaOperand {
  const float* a_hi; const float* a_lo; int lda;   // A [M,K]
  const float* w_hi; const float* w_lo; int ldw;   // W [N,K]
  const float* bias;
  float* C; int ldc;
  int M, N, K;
  float* C_lo;   // optional: emit the result pre-split (C = hi, C_lo = lo) for a chained 3xTF32 GEMM
  int relu;
  int t_rows;    // > 0: transposed store C[(m / t_rows) * N + n][m % t_rows] with row pitch ldc (V^T per image)
  int bias_per_row;  // bias indexed by output row (weights as the A operand, e.g. V^T = Wv . X^T)
  int a_is_raw;      // a_hi points to the RAW fp32 A (a_lo unused): the kernel splits it in the SM through TMEM
  int w_bf16;        // w_hi / w_lo point to bf16 hi / lo planes (launch_split_bf16): the "3xBF16" variant (raw A only)
  // sign/bit-pack epilogue instead of a store (C may be null): see umma::Problem
  uint32_t* bits; int* rowany; int bits_words;
};
int launch_split_tf32(const float* x, float* hi, float* lo, size_t n, cudaStream_t st);
int launch_split_bf16(const float* x, void* hi, void* lo, size_t n, cudaStream_t st);  // n elements -> bf16 planes
int launch_umma_gemm(const UmmaOperand* ops, int count, int passes, cudaStream_t st);

struct LnArgs {
  const float* x;          // [nparts][M][256]
  int nparts;
  long long part_stride;
  const float* bias;       // [256] or null
  const float* resid;      // [M][256] or null
  const float* gamma;
  const float* beta;
  float* y;                // LN output
  const float* pos;        // [pos_mod][256] or null
  int pos_mod;
  float* ypos;             // y + pos[m % pos_mod] or null
  float* y_hi;             // optional 3xTF32 split copies of y and ypos (tensor-core consumers)
  float* y_lo;
  float* ypos_hi;
  float* ypos_lo;
  const float* gamma2;     // optional chained second LN (post_norm)
  const float* beta2;
  float* y2;
  int* zero_rows;          // optional [M] int32 cleared by this launch (attention-mask `rowany` flags, consumed upstream)
  int M;
};
int launch_layernorm(const LnArgs& a, cudaStream_t st);

struct MhaArgs {
  const float* q; int ldq;
  const float* k; int ldk;
  const float* v; int ldv;
  const uint32_t* mask_bits; int mask_words;
  const int* rowany;
  float* out;       // [B,Nq,256]
  int B, Nq, Nk;
};
size_t mha_workspace_bytes(int B, int Nq, int Nk);
int launch_mha(const MhaArgs& a, void* ws, size_t ws_bytes, cudaStream_t st);

int launch_mha_combine(const float* opart, const float2* ml, float* out, int B, int Nq, int S, cudaStream_t st);

// tensor-core flash attention (fa_umma.cu): operands pre-split hi/lo; q pre-scaled by (1/sqrt(32))*log2(e)
struct FaArgs {
  const float* q_hi; const float* q_lo;      // [B*Nq, 256]
  const float* k_hi; const float* k_lo;      // [B*Nk, 256]
  const float* vt_hi; const float* vt_lo;    // [B*256, ldv]  V^T per image (row = head*32 + dim, col = key)
  int ldv;
  const uint32_t* mask_bits; int mask_words;
  const int* rowany;
  float* out;                                // [B,Nq,256]
  int B, Nq, Nk;
  // optional strided views (0 = dense defaults): K rows have leading dimension ldk (columns k_col0 .. k_col0+255 are
  // used); V^T of image b starts at row b * vt_img_rows + vt_row0
  int ldk, k_col0, vt_img_rows, vt_row0;
};
size_t fa_workspace_bytes(int B, int Nq, int Nk);
int launch_fa_umma(const FaArgs& a, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_split_tf32_scaled(const float* x, float* hi, float* lo, size_t n, float scale, cudaStream_t st);
int launch_split_transpose(const float* v, float* vt_hi, float* vt_lo, int B, int Nk, int ldv, cudaStream_t st);
constexpr float ATTN_QSCALE = 0.17677669529663687f * 1.4426950408889634f;  // (1/sqrt(32)) * log2(e)

// fused decoder chain (chain.cu): whole BaseTransformerLayers in one cluster launch, operands from the prepared blob
constexpr int CHAIN_MAX_LAYERS = 6;
struct ChainLayer {  // hi/lo TF32 splits of one layer's weights (row-major, torch layout) + fp32 biases / norms
  const float *cin_hi, *cin_lo, *co_hi, *co_lo, *sin_hi, *sin_lo, *so_hi, *so_lo, *f1_hi, *f1_lo, *f2_hi, *f2_lo;
  const float *cin_b, *co_b, *sin_b, *so_b, *f1_b, *f2_b;
  const float* gamma[3];
  const float* beta[3];
};
struct ChainArgs {
  ChainLayer layers[CHAIN_MAX_LAYERS];
  float* x;                                          // [B*R,256] fp32 layer input / output (written from init_feat when given)
  float* scratch;                                    // chain_scratch_floats(B, R, ffn) floats
  // activations exchanged with other kernels as TF32 hi/lo pairs (null = internal scratch):
  float *xpos_hi, *xpos_lo;                          //   x + query_pos   [B*R,256]  (in when init_feat is null; out)
  float *att_hi, *att_lo;                            //   cross-attention output [B*R,256]  (in when !has_cross_attn)
  float *q_hi, *q_lo;                                //   scaled query projection of the NEXT layer's cross attention (out)
  const float *kc_hi, *kc_lo;                        // cross keys of all layers [B*Nk, nl*256]   (has_cross_attn)
  const float *vtc_hi, *vtc_lo;                      // cross V^T [B*nl*256, ldvc]
  const float *init_feat, *qpos;                     // [R,256]
  const float *cls_hi, *cls_lo, *cls_b;              // final classifier [ncls,256] (cls_out null = none)
  float* cls_out;
  // Mask2Former tail: xn = post_norm(x) (fp32 out), e = mask_embed(xn) split hi/lo, q of the next layer's cross-attention
  float *xn, *e_hi, *e_lo;
  const float *pn_gamma, *pn_beta;
  const float* me_hi[3]; const float* me_lo[3]; const float* me_b[3];
  const float *nq_hi, *nq_lo, *nq_b;                 // next layer's cross in_proj [768,256] split (null = last layer)
  float* trace;                                      // optional [nl,B*R,256]
  int* zero_rows;                                    // optional [B*R]
  int B, R, Nk, nl, ffn, ncls, ldvc, has_cross_attn, m2f_tail;
};
int launch_decoder_chain(const ChainArgs& a, cudaStream_t st);
size_t chain_scratch_floats(int B, int R, int ffn);
void chain_set_timing(unsigned long long* buf, int cap);

int launch_sine_posenc(float* pos, int h, int w, cudaStream_t st);
int launch_level_prep(const float* mem, const float* level_embed, const float* pos, float* x, float* xp,
                      int B, int hw, cudaStream_t st, float* x_lo = nullptr, float* xp_lo = nullptr);
int launch_level_prep_tokens(const float* mem, long long bstride, const float* level_embed, const float* pos, float* x,
                             float* xp, int B, int hw, cudaStream_t st, float* x_lo = nullptr, float* xp_lo = nullptr);
// runtime options (pn_set_option)
enum { OPT_TENSOR_CORES = 0, OPT_UMMA_WIDE = 1, OPT_UMMA_EPI8 = 2, OPT_OVERLAP = 3, OPT_FA_TC = 4, OPT_UMMA_RAW_A = 5, OPT_TOPK_RADIX = 6, OPT_PPN_TC = 7, OPT_SKINNY = 8, OPT_MASK_TC = 9, OPT_FUSED_CHAIN = 10, OPT_PPN_FUSED_TOPK = 11, OPT_CONV_TC = 12, OPT_UMMA_TMA_STORE = 13, OPT_SINGLE_PASS = 14, OPT_NVTX = 15, OPT_PDL = 16, OPT_ENC_BF16X3 = 17, OPT_PPN_EPI2 = 18, OPT_PPN_HALF_KB = 19, OPT_PPN_SPECULATE = 20, OPT_COUNT = 21 };
int get_option(int key);
int launch_mask_feature_resize(const float* F, float* out, int B, int H, int W, int h, int w, int ldo,
                               cudaStream_t st);
int launch_mask_feature_resize_tokens(const float* F, float* out, int B, int H, int W, int h, int w, cudaStream_t st);
int launch_nchw_to_tokens(const float* src, float* dst, int B, int HW, cudaStream_t st);
int launch_bcast_rows(const float* a, const float* b, float* out, float* out_sum, int B, int N,
                      cudaStream_t st);
int launch_add_rows(const float* x, const float* pos, float* out, int B, int N, cudaStream_t st);
int launch_l2norm(const float* x, float* y, int M, cudaStream_t st);
int launch_gather_rows(const float* src, const int64_t* idx, float* dst, int B, int Nsrc, int R,
                       long long L, cudaStream_t st);
int launch_conv_tiny(const float* x, const PnConvTiny* conv, float* y, int B, int N, void* ws, size_t ws_bytes,
                     cudaStream_t st);
size_t conv_tiny_workspace_bytes(int B, int N, int mid);
int launch_pair_matrix_tc(const float* S, const float* O, float* C, int B, int N, int K, cudaStream_t st);
int launch_topk_pairs(const float* imp, int64_t* topk_idx, int64_t* sub_pos, int64_t* obj_pos,
                      const float* query, float* pair_feat, int B, int N, int K, cudaStream_t st,
                      const int* only_if = nullptr);
size_t conv_tiny_tc_workspace_bytes(int B, int N);
int launch_conv_tiny_tc(const float* x, const PnConvTiny* cv, float* y, int B, int N, void* ws, size_t ws_bytes,
                        cudaStream_t st);
bool pair_topk_fused_supported(int N, int K, int topk, bool bf16 = false);
// S, O: fp32 [B,N,K] (bf16 = false, 3xTF32) or __nv_bfloat16 [B,N,K] (bf16 = true, one kind::f16 pass)
int launch_pair_topk_fused(const void* S, const void* O, bool bf16, float* C, int64_t* topk_idx, int64_t* sub_pos,
                           int64_t* obj_pos, int* redo, int B, int N, int K, int topk, cudaStream_t st);

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = get_option(OPT_PDL) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

}  // namespace pn
