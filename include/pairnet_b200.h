/*
 * pairnet_b200 -- C-ABI of the B200 (sm_100a) implementation of Pair-Net's relation-head hot path.
 *
 * The reference (king159/Pair-Net) is 100 % Python and has NO FFI of its own: every entry point
 * below replaces a span of eager PyTorch calls in
 *     pairnet/models/relation_heads/pairnet_head.py   (class CrossHead2)
 *     pairnet/models/frameworks/cnn_factory.py        (class ConvTiny)
 * and is bound from Python with ctypes (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types.
 *   - every pointer is a DEVICE pointer valid on `stream` (a cudaStream_t passed as void*).
 *   - caller owns all buffers; nothing is allocated behind the caller's back; scratch memory is
 *     passed in as (ws, ws_bytes) and sized with the matching *_workspace_bytes() query.
 *   - return value: 0 = ok; <0 = PN_ERR_* ; >0 = a cudaError_t raised by a launch.
 *     pn_last_error_string() describes the last failure on the calling thread.  Never throws.
 *   - all activations are fp32, row-major, BATCH-MAJOR ([B, seq, d]); the reference's internal
 *     [seq, B, d] layout is only a view choice (outputs of CrossHead2.forward are batch-first).
 *   - asynchronous: kernels are enqueued on `stream`; CUDA-graph capturable (no sync, no malloc).
 */
#ifndef PAIRNET_B200_H_
#define PAIRNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN_VERSION 100          /* 0.1.0 */
#define PN_MAX_LAYERS 16
#define PN_MAX_LEVELS 4
#define PN_EMBED_DIMS 256       /* compiled-in embed dim (all shipped Pair-Net configs use 256) */
#define PN_HEAD_DIM 32          /* 256 / 8 heads */

#define PN_ERR_BAD_ARG      (-1)
#define PN_ERR_UNSUPPORTED  (-2)
#define PN_ERR_WORKSPACE    (-3)

#if defined(__GNUC__)
#define PN_API __attribute__((visibility("default")))
#else
#define PN_API
#endif

typedef void* pn_stream_t;

/* nn.Linear: y = x W^T + b ; w is [out, in] row-major (torch layout), b may be NULL */
typedef struct { const float* w; const float* b; } PnLinear;
/* nn.LayerNorm(256), eps 1e-5 */
typedef struct { const float* gamma; const float* beta; } PnNorm;
/* nn.MultiheadAttention(256, 8): in_proj_weight [768,256] = [Wq;Wk;Wv], in_proj_bias [768] */
typedef struct {
  const float* in_proj_w; const float* in_proj_b;
  const float* out_proj_w; const float* out_proj_b;
} PnMHA;
/* mmcv BaseTransformerLayer, operation_order (cross_attn, norm, self_attn, norm, ffn, norm)
 * -- configs/mask2former/pairnet.py:72-139 */
typedef struct {
  PnMHA cross_attn;      /* attentions.0 */
  PnMHA self_attn;       /* attentions.1 */
  PnLinear ffn1;         /* ffns.0.layers.0.0  [ffn, 256] */
  PnLinear ffn2;         /* ffns.0.layers.1    [256, ffn] */
  PnNorm norm[3];        /* norms.0..2 */
} PnDecoderLayer;
/* Linear-ReLU-Linear-ReLU-Linear (mask_embed, sub_query_update, obj_query_update) */
typedef struct { PnLinear l[3]; } PnMlp3;
/* ConvTiny: conv_layers.{0,1,2}.0  w0 [mid,1,7,7]  w1 [mid,mid,7,7]  w2 [1,mid,7,7] */
typedef struct { const float* w[3]; const float* b[3]; int mid_channels; } PnConvTiny;

/* ------------------------------------------------------------------ library */
PN_API int pn_version(void);
PN_API const char* pn_last_error_string(void);
/* process-wide options. PN_OPT_TENSOR_CORES (default 1): memory-side GEMMs with >= 1024 rows run on the
 * tcgen05 3xTF32 kernel; 0 = exact-fp32 FFMA kernels everywhere (A/B and parity studies). */
#define PN_OPT_TENSOR_CORES 0
#define PN_OPT_UMMA_WIDE 1      /* default 0 (measured slower: only 2 smem stages fit); 1 = 128x256 tcgen05 tiles when N % 256 == 0 */
PN_API int pn_set_option(int key, int value);
PN_API int pn_get_option(int key);
#define PN_OPT_UMMA_EPI8 2      /* default 0: 4 epilogue warps in the tcgen05 GEMM (1 = 8; measured slower in the encoder) */
#define PN_OPT_OVERLAP 3        /* default 1: memory-side K/V projections and the output head run on an internal side
                                 * stream (fork/join with events; parallel branches under graph capture); 0 = one stream */
#define PN_OPT_FA_TC 4          /* default 1: masked cross-attention of levels with >= 1024 tokens on tcgen05 (0 = FFMA) */
#define PN_OPT_UMMA_RAW_A 5     /* default 1: activations enter the tcgen05 GEMM raw and are split hi/lo inside the SM
                                 * (through TMEM) instead of being materialised pre-split by their producers */
#define PN_OPT_TOPK_RADIX 6     /* default 0: top-k by local-maxima threshold + candidate ranking; 1 = always the exact
                                 * 4-pass radix select (the fallback of the default path; parity studies) */
#define PN_OPT_PPN_TC 7         /* default 1: pair matrix S.O^T of batches with >= 1024 embedding rows on tcgen05 (3xTF32,
                                 * operands split in the SM); 0 = exact-fp32 FFMA */
#define PN_OPT_MASK_TC 9        /* default 1: mask einsums (attention-mask bits, final mask_pred) on the tcgen05 GEMM with
                                   token-major operands; 0 = FFMA kernels on NCHW / N-major operands */
#define PN_OPT_FUSED_CHAIN 10    /* Relation Fusion decoder with prepared weights (pn_rel_prepare): key side of all layers on the
                                   tcgen05 GEMM, then  2 = always ONE cluster launch of the fused tcgen05 chain kernel (chain.cu:
                                   6 layers + classifier);  1 (default) = that chain from 8 images per call (it runs 8 SMs per
                                   image), below per-op kernels on all SMs with the cross attention on tcgen05 (fa_umma.cu);
                                   0 = round-1 per-op kernels (warp-MMA linears, FFMA attention), prepared blob unused */
#define PN_OPT_PPN_FUSED_TOPK 11 /* default 1: micro-benchmark mode of pn_ppn_forward (no MLPs, no ConvTiny, no pair_feat) with >= 1024
                                    embedding rows: pair matrix and top-k in ONE tcgen05 kernel (pair_topk.cu), the matrix is written
                                    once and never read back; 0 = pair-matrix kernel followed by the stand-alone top-k kernel */
#define PN_OPT_CONV_TC 12       /* default 1: ConvTiny conv2 (64 -> 64, 7x7) as a tcgen05 implicit GEMM (3xTF32, shifted-window A
                                    descriptors over TMA-loaded slabs, conv_umma.cu); 0 = FFMA kernels of ppn.cu */
#define PN_OPT_UMMA_TMA_STORE 13 /* default 1: row-major epilogue of the tcgen05 GEMM through swizzled staging tiles + TMA stores
                                    (0 = one 128-bit store per thread and row: measured store-issue bound) */
#define PN_OPT_SINGLE_PASS 14    /* default 0 = fp32 parity (3xTF32).  1 = reduced-precision mode for the bf16-class configs: the
                                    memory-side tcgen05 GEMMs (K/V projections, mask einsums, pixel-decoder encoder) and the masked
                                    cross-attention run ONE kind::tf32 pass on the raw fp32 operands (10-bit mantissa >= bf16's 8);
                                    the pair matrix, ConvTiny and the top-k stay at fp32 parity */
#define PN_OPT_NVTX 15           /* default 1: NVTX v3 ranges around the stages of the hot path (pn::m2f_decoder, pn::m2f_layer,
                                    pn::ppn, pn::relation_fusion, pn::output_gathers); no-ops unless a profiler is attached */
#define PN_OPT_PDL 16            /* default 1: the small kernels of the query-side chain (skinny GEMM, LayerNorm, small attention, row
                                    ops) are launched with programmatic stream serialization (under graph capture: programmatic
                                    edges); each waits with griddepcontrol.wait before touching activations, the skinny GEMM
                                    fetches its (static) weights ahead of the wait.  0 = plain serialized launches */
#define PN_OPT_ENC_BF16X3 17     /* default 1: the pixel-decoder encoder's GEMMs (upstream of the hot path, fed by TF32 cuDNN convolutions)
                                    run as "3xBF16": raw fp32 activations split into bf16 hi / lo pairs in the SM, weights as
                                    prepared bf16 hi / lo planes, three tcgen05.mma kind::f16 products (twice the TF32 rate),
                                    ~1e-5 of the output scale.  0 = 3xTF32 (~1e-6), as everywhere in the head */
#define PN_OPT_PPN_EPI2 18       /* default 1: pn_ppn_pair_topk_bf16 on images of one or two tiles (N <= 224) runs TWO epilogue / top-k groups
                                    of 8 warps, each owning every other image, so the per-image top-k chains of two images overlap; 0 = one;
                                    2 = the fp32 kernel too where shared memory allows (A/B only: measured slower, 0.57 vs 0.62 at N = 100 --
                                    26 warps and a 2-stage operand ring) */
#define PN_OPT_PPN_HALF_KB 19    /* default 1: the fp32 pair-matrix + top-k kernel uses 16-channel k-blocks (64-byte rows, SWIZZLE_64B) when
                                    fewer than four 32-channel raw stages fit shared memory (N >= ~128): twice the ring depth; 0 = always 32 */
#define PN_OPT_PPN_SPECULATE 20  /* default 1: fused pair-matrix + top-k kernel, single-tile images: the ~1.5 K-th largest value of the previous
                                    image of the CTA is the speculative candidate threshold of the next one (one pass over the accumulator);
                                    K <= candidates <= 1408 proves the top-K is among them, otherwise the image is redone by the exact
                                    kernel.  0 = always the exact local-maxima threshold + second pass */
#define PN_OPT_SKINNY 8         /* default 1: query-side linears (< 1024 rows) on the latency-optimised warp-MMA kernel
                                   (3xTF32, no smem staging, one exposed memory round trip); 0 = k-tiled FFMA kernel */
/* fills SM count and compute capability of the current device */
PN_API int pn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------ row 1: level prep
 * replaces pairnet_head.py:268-288 (flatten/permute + level_embed, SinePositionalEncoding). */
/* pos [h*w, 256] token-major sine encoding of an all-false mask (input independent). */
PN_API int pn_sine_posenc(float* pos, int h, int w, pn_stream_t stream);
/* mem [B,256,hw] -> x [B,hw,256] = mem^T + level_embed ; xp = x + pos */
PN_API int pn_level_prep(const float* mem, const float* level_embed, const float* pos,
                  float* x, float* xp, int B, int hw, pn_stream_t stream);
/* same for a token-major memory: element (b, p, c) at mem[b*batch_stride + p*256 + c] (the pixel decoder's encoder
 * output sliced per level) -- no NCHW round trip */
PN_API int pn_level_prep_tokens(const float* mem, long long batch_stride, const float* level_embed, const float* pos,
                                float* x, float* xp, int B, int hw, pn_stream_t stream);

/* ------------------------------------------------------------------ row 2: forward_head
 * replaces pairnet_head.py:216-258 (post_norm, cls_embed, mask_embed, einsum, bilinear, threshold) */
/* bilinear (align_corners=False) resize of mask_features [B,256,H,W] to [B,256,ldo] rows of
 * h*w valid pixels (ldo >= h*w, multiple of 32, zero padded).  Linear, so it commutes with the
 * einsum over channels: sign(E . resize(F)) == sign(resize(E . F)) up to fp32 rounding. */
PN_API int pn_mask_feature_resize(const float* mask_feature, float* out, int B, int H, int W,
                           int h, int w, int ldo, pn_stream_t stream);
/* attention-mask bits: bit p of word [b][q][p/32] = 1 (blocked) iff sum_c E[b,q,c] F[b,c,p] < 0.
 * rowany[b*N+q] is OR-ed with 1 when the row has at least one unblocked key (must be zeroed by
 * the caller).  E [B,N,256], F [B,256,ldf], bits [B,N,words] with words = ldf/32. */
PN_API int pn_attn_mask_bits(const float* E, const float* F, uint32_t* bits, int* rowany,
                      int B, int N, int hw, int ldf, pn_stream_t stream);
/* mask_pred [B,N,HW] = E [B,N,256] . F [B,256,HW]   (the einsum "bqc,bchw->bqhw") */
PN_API int pn_mask_pred(const float* E, const float* F, float* mask_pred, int B, int N, int HW,
                 pn_stream_t stream);
/* tensor-core variants on token-major operands (F_tokens [B,hw,256]; E [B,N,256], N <= 256):
 * tcgen05 3xTF32 GEMM with keys / pixels as the M tiles; sign + warp-ballot epilogue (bits) or transposed store (pred).
 * bits / rowany as in pn_attn_mask_bits (rowany must be zeroed by the caller), words >= ceil(hw/32). */
PN_API int pn_mask_feature_resize_tokens(const float* mask_feature_tokens /* [B,H*W,256] */, float* out /* [B,h*w,256] */,
                                         int B, int H, int W, int h, int w, pn_stream_t stream);
PN_API int pn_nchw_to_tokens(const float* src /* [B,256,HW] */, float* dst /* [B,HW,256] */, int B, int HW,
                             pn_stream_t stream);
PN_API size_t pn_mask_tc_workspace_bytes(int B, int N);
PN_API int pn_attn_mask_bits_tc(const float* E, const float* F_tokens, uint32_t* bits, int* rowany, int B, int N, int hw,
                                int words, void* ws, size_t ws_bytes, pn_stream_t stream);
PN_API int pn_mask_pred_tc(const float* E, const float* F_tokens, float* mask_pred /* [B,N,HW] */, int B, int N, int HW,
                           void* ws, size_t ws_bytes, pn_stream_t stream);

/* ------------------------------------------------------------------ generic bricks */
/* y[M,N] = act(x[M,K] W[N,K]^T + b) (+ resid[M,N]);  relu: 0/1 */
PN_API int pn_linear(const float* x, int ldx, const float* w, const float* b, const float* resid,
              float* y, int ldy, int M, int N, int K, int relu, pn_stream_t stream);
/* Tensor-core variant of pn_linear for large M (tcgen05.mma kind::tf32, TMA-staged tiles, TMEM accumulator).
 * passes = 3: operands are split hi/lo ("3xTF32") so the result matches fp32 FFMA to ~1e-6;
 * passes = 1: plain TF32.  K % 32 == 0, ldx/ldy % 4 == 0.  ws >= pn_linear_tc_workspace_bytes(M,N,K). */
PN_API size_t pn_linear_tc_workspace_bytes(int M, int N, int K);
PN_API int pn_linear_tc(const float* x, int ldx, const float* w, const float* b, float* y, int ldy,
                        int M, int N, int K, int passes, void* ws, size_t ws_bytes, pn_stream_t stream);
/* the two halves of pn_linear_tc, for callers that keep operands pre-split (and for timing the GEMM alone):
 * hi = rna_tf32(x), lo = rna_tf32(x - hi) over n floats (n % 4 == 0); then the tcgen05 GEMM on split operands. */
PN_API int pn_split_tf32(const float* x, float* hi, float* lo, size_t n, pn_stream_t stream);
/* raw-A variant (default in production): x stays plain fp32 and is split inside the SM through TMEM */
PN_API int pn_linear_tc_rawa(const float* x, const float* w_hi, const float* w_lo, const float* b, float* y,
                             int ldy, int M, int N, int K, pn_stream_t stream);
/* "3xBF16" form of pn_linear_tc_rawa: y = x W^T + b with fp32 activations in and out, on tcgen05.mma kind::f16 at twice the
 * TF32 rate.  w_hi / w_lo: bf16 planes of W [N,K] from pn_split_bf16 (hi = bf16(w), lo = bf16(w - hi), bit patterns as
 * uint16_t); x is split into bf16 hi / lo pairs inside the SM.  Three products (hi*hi + hi*lo + lo*hi): ~1e-5 of the output
 * scale (3xTF32: ~1e-6).  K % 64 == 0.  Used by the pixel-decoder encoder (PN_OPT_ENC_BF16X3). */
PN_API int pn_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, size_t n, pn_stream_t stream);
PN_API int pn_linear_tc_bf16x3(const float* x, const uint16_t* w_hi, const uint16_t* w_lo, const float* b, float* y,
                        int ldy, int M, int N, int K, pn_stream_t stream);
PN_API int pn_linear_tc_presplit(const float* x_hi, const float* x_lo, const float* w_hi, const float* w_lo,
                                 const float* b, float* y, int ldy, int M, int N, int K, int passes,
                                 pn_stream_t stream);
/* y = LayerNorm(x + resid) over 256 channels (resid may be NULL) */
PN_API int pn_add_layernorm(const float* x, const float* resid, const float* gamma, const float* beta,
                     float* y, int M, pn_stream_t stream);
/* scaled-dot-product attention core of nn.MultiheadAttention for already projected q/k/v.
 * q [B,Nq,*] (row stride ldq), k,v [B,Nk,*]; 8 heads x 32; mask_bits (nullable) [B,Nq,words],
 * rowany (nullable) [B*Nq]: rows whose flag is 0 ignore the mask (pairnet_head.py:300).
 * out [B,Nq,256]. ws >= pn_mha_workspace_bytes(B,Nq,Nk). */
PN_API size_t pn_mha_workspace_bytes(int B, int Nq, int Nk);
PN_API int pn_mha_core(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                const uint32_t* mask_bits, int mask_words, const int* rowany,
                float* out, int B, int Nq, int Nk, void* ws, size_t ws_bytes, pn_stream_t stream);

/* Tensor-core variant of pn_mha_core (fa_umma.cu): QK^T and PV on tcgen05 (3xTF32 operands, P staged through
 * TMEM); same arguments, dense q/k/v with row stride 256.  Used by the M2F decoder for memory levels with
 * >= 1024 tokens (there the K/V projection emits the split / transposed operands directly). */
PN_API size_t pn_mha_core_tc_workspace_bytes(int B, int Nq, int Nk);
PN_API int pn_mha_core_tc(const float* q, const float* k, const float* v, const uint32_t* mask_bits,
                          int mask_words, const int* rowany, float* out, int B, int Nq, int Nk, void* ws,
                          size_t ws_bytes, pn_stream_t stream);

/* ------------------------------------------------------------------ rows 3 (+1,2): M2F decoder
 * replaces pairnet_head.py:262-320 minus the pixel decoder call. */
typedef struct {
  int num_queries;            /* N  (num_obj_query) */
  int num_layers;             /* 9 */
  int num_levels;             /* 3 */
  int ffn_dims;               /* 2048 */
  int num_cls;                /* num_classes + 1 = 134 */
  const float* query_feat;    /* query_feat.weight  [N,256] */
  const float* query_embed;   /* query_embed.weight [N,256] */
  const float* level_embed;   /* level_embed.weight [levels,256] */
  PnNorm post_norm;           /* transformer_decoder.post_norm */
  PnLinear cls_embed;         /* [num_cls,256] */
  PnMlp3 mask_embed;
  PnDecoderLayer layers[PN_MAX_LAYERS];
  const void* prepared;       /* optional: blob built by pn_m2f_prepare (static hi/lo weight splits); NULL = split per forward */
} PnM2FWeights;

typedef struct {
  int B;
  int H4, W4;                          /* mask_features spatial size */
  const float* mask_features;          /* [B,256,H4,W4] */
  int h[PN_MAX_LEVELS], w[PN_MAX_LEVELS];
  const float* memory[PN_MAX_LEVELS];  /* multi_scale_memorys[l] [B,256,h,w], low -> high res */
  const float* pos[PN_MAX_LEVELS];     /* optional precomputed pn_sine_posenc tables (NULL = compute) */
  /* memory_token_major[l] != 0: memory[l] is token-major, element (b, p, c) at memory[l][b*stride + p*256 + c] with
   * stride = memory_batch_stride[l] (a level slice of the pixel decoder's encoder output, no NCHW copy needed) */
  int memory_token_major[PN_MAX_LEVELS];
  long long memory_batch_stride[PN_MAX_LEVELS];
  /* != 0: mask_features is token-major [B,H4*W4,256] (a channels_last map); needs PN_OPT_MASK_TC */
  int mask_features_token_major;
} PnM2FInputs;

typedef struct {
  float* query_out;     /* [B,N,256]   last layer query_feat (pre post_norm) -- PPN input */
  float* cls_pred;      /* [B,N,num_cls] */
  float* mask_pred;     /* [B,N,H4*W4] */
  float* query_trace;   /* optional [layers,B,N,256] per-layer query_feat (NULL = skip) */
  uint32_t* mask_trace; /* optional [layers,B,N,trace_words] attention-mask bits (NULL = skip) */
  int trace_words;
} PnM2FOutputs;

PN_API size_t pn_m2f_decoder_workspace_bytes(const PnM2FWeights* w, const PnM2FInputs* in);
PN_API int pn_m2f_decoder_forward(const PnM2FWeights* w, const PnM2FInputs* in, const PnM2FOutputs* out,
                           void* ws, size_t ws_bytes, pn_stream_t stream);

/* ------------------------------------------------------------------ rows 4-8: Pair Proposal Network
 * replaces pairnet_head.py:322-351,365 and cnn_factory.py:49-53.
 * query [B,N,256] -> sub/obj MLP -> L2 normalise -> importance_raw = S O^T -> ConvTiny ->
 * top-k (k=K, descending, ties: lower flat index first) -> sub_pos = idx / N, obj_pos = idx % N ->
 * pair_feat [B,2K,256] = [query[sub_pos] ; query[obj_pos]].
 * conv == NULL selects the "pair matrix + top-k only" microbenchmark mode (BASELINE config 5a).
 * sub_mlp == NULL: `query` is taken as already-normalised sub embeddings and `query_obj` as obj
 * embeddings (microbenchmark inputs); otherwise query_obj must be NULL. */
/* mid_channels <= 0: workspace of the ConvTiny-less microbenchmark mode (conv == NULL) only. */
PN_API size_t pn_ppn_workspace_bytes(int B, int N, int K, int mid_channels);
PN_API int pn_ppn_forward(const float* query, const float* query_obj, const PnMlp3* sub_mlp,
                   const PnMlp3* obj_mlp, const PnConvTiny* conv,
                   float* importance_raw /* [B,N,N] nullable */, float* importance /* [B,N,N] */,
                   int64_t* topk_idx /* [B,K] nullable */, int64_t* sub_pos /* [B,K] */,
                   int64_t* obj_pos /* [B,K] */, float* pair_feat /* [B,2K,256] nullable */,
                   int B, int N, int K, void* ws, size_t ws_bytes, pn_stream_t stream);
/* bf16 variant of the "pair matrix + top-k only" mode ("bf16 variants take __nv_bfloat16* for activations,
 * fp32 for importance and logits"; BASELINE config 5, "fp32 (and a bf16 run)"; pairnet_head.py:327-340 with the embeddings
 * stored as bf16).  sub_embed / obj_embed: [B,N,256] bf16 bit patterns (`__nv_bfloat16`, passed as uint16_t so that this
 * header stays plain C).  One tcgen05.mma kind::f16 pass (bf16 x bf16 products are exact in the fp32 accumulator), the
 * matrix is written once as fp32 and the top-k runs on the tensor-memory accumulator.  N % 4 == 0, K <= 256. */
PN_API size_t pn_ppn_pair_topk_bf16_workspace_bytes(int B);
PN_API int pn_ppn_pair_topk_bf16(const uint16_t* sub_embed, const uint16_t* obj_embed,
                   float* importance /* [B,N,N] */, int64_t* topk_idx /* [B,K] nullable */,
                   int64_t* sub_pos /* [B,K] */, int64_t* obj_pos /* [B,K] */,
                   int B, int N, int K, void* ws, size_t ws_bytes, pn_stream_t stream);
/* stand-alone pieces (stage-wise parity tests) */
PN_API int pn_conv_tiny(const float* x /* [B,N,N] */, const PnConvTiny* conv, float* y /* [B,N,N] */,
                 int B, int N, void* ws, size_t ws_bytes, pn_stream_t stream);
PN_API int pn_topk_pairs(const float* importance /* [B,N*N] */, int64_t* topk_idx, int64_t* sub_pos,
                  int64_t* obj_pos, const float* query /* nullable */, float* pair_feat /* nullable */,
                  int B, int N, int K, pn_stream_t stream);

/* ------------------------------------------------------------------ rows 9-10: Relation Fusion
 * replaces pairnet_head.py:353-378. */
typedef struct {
  int num_rel_queries;           /* R */
  int num_layers;                /* 6 */
  int ffn_dims;                  /* 2048 */
  int num_rel_cls;               /* 56 */
  const float* rel_query_feat;   /* [R,256] */
  const float* rel_query_embed;  /* [R,256]  query_pos */
  const float* rel_query_embed2; /* [2K,256] key_pos   (rel_query_embed3 is dead in the reference) */
  PnLinear rel_cls_embed;        /* [num_rel_cls,256] */
  PnDecoderLayer layers[PN_MAX_LAYERS];
  const void* prepared;          /* optional: blob built by pn_rel_prepare; NULL = per-op kernels, nothing cached */
} PnRelWeights;

/* Prepared weights.  The tensor-core kernels consume every weight matrix as a TF32 hi/lo pair (3xTF32, fp32 parity).
 * Weights are static between optimiser steps, so the pairs (plus the per-stage concatenations the fused kernels
 * read) are built ONCE into a caller-owned device buffer and referenced from PnRelWeights.prepared /
 * PnM2FWeights.prepared.  Re-run after the weights change.  `w->prepared` itself is ignored by these calls. */
PN_API size_t pn_rel_prepared_bytes(const PnRelWeights* w);
PN_API int pn_rel_prepare(const PnRelWeights* w, void* prepared, size_t bytes, pn_stream_t stream);
PN_API size_t pn_m2f_prepared_bytes(const PnM2FWeights* w);
PN_API int pn_m2f_prepare(const PnM2FWeights* w, void* prepared, size_t bytes, pn_stream_t stream);

PN_API size_t pn_relation_fusion_workspace_bytes(int B, int R, int K2, int ffn_dims);
PN_API int pn_relation_fusion_forward(const PnRelWeights* w, const float* pair_feat /* [B,K2,256] */,
                               float* rel_preds /* [B,R,num_rel_cls] */,
                               float* rel_feat_out /* [B,R,256] nullable */,
                               int B, int K2, void* ws, size_t ws_bytes, pn_stream_t stream);

/* ------------------------------------------------------------------ row 11: output gathers
 * replaces pairnet_head.py:380-403: dst[b,r,:] = src[b, idx[b,r], :] (row length L floats) */
PN_API int pn_gather_rows(const float* src, const int64_t* idx, float* dst, int B, int Nsrc, int R,
                   long long L, pn_stream_t stream);

/* ------------------------------------------------------------------ inference post-processing
 * replaces the heavy part of CrossHead2._get_bboxes_single, pairnet_head.py:826-905: the three full-image
 * F.interpolate(bilinear, align_corners=False) + sigmoid > 0.5 / softmax-argmax / per-mask area counts.  Both read the
 * quarter-resolution logits [N,h,w] of ONE image and evaluate the interpolation on the fly.
 *   pn_upsample_threshold: out[r] = (up(mask[idx[r]]) > 0) as uint8 [R,H,W]; idx NULL = rows 0..R-1 (then R <= N).
 *   pn_panoptic_merge    : m = argmax_k up(mask[keep_idx[k]]) (first maximum), id = remap[m] (stuff de-duplication),
 *                          pan[y][x] = id * instance_offset + labels[id] (int64), area[id] += 1 (area is zeroed here). */
PN_API int pn_upsample_threshold(const float* mask, const int64_t* idx, uint8_t* out, int N, int R, int h, int w, int H,
                                 int W, pn_stream_t stream);
PN_API int pn_panoptic_merge(const float* mask, const int* keep_idx, const int* remap, const int64_t* labels, int n_keep,
                             int h, int w, int H, int W, long long instance_offset, int64_t* pan, int* area,
                             pn_stream_t stream);

/* ------------------------------------------------------------------ upstream "next" row
 * The 6-layer multi-scale deformable-attention encoder of mmdet's MSDeformAttnPixelDecoder
 * (cfg configs/mask2former/pairnet.py:38-66; called from pairnet_head.py:262).  8 heads x 32. */
typedef struct {
  PnLinear sampling_offsets;   /* [8*L*P*2, 256] */
  PnLinear attention_weights;  /* [8*L*P, 256] */
  PnLinear value_proj, output_proj;
  PnLinear ffn1, ffn2;         /* [ffn,256], [256,ffn] */
  PnNorm norm[2];
} PnMsdaEncoderLayer;
typedef struct {
  int num_layers, num_levels, num_points, ffn_dims;
  PnMsdaEncoderLayer layers[PN_MAX_LAYERS];
  const void* prepared; /* NULL, or the blob filled by pn_msda_encoder_prepare (static TF32 weight splits) */
} PnMsdaEncoderWeights;
/* x_in/x_out [B,nq,256] token-major (levels concatenated, low -> high res order of `h`,`w`);
 * pos [nq,256] = sine position + level encoding.  Linears run on the tcgen05 3xTF32 GEMM. */
PN_API size_t pn_msda_encoder_prepared_bytes(const PnMsdaEncoderWeights* w);
PN_API int pn_msda_encoder_prepare(const PnMsdaEncoderWeights* w, void* blob, size_t blob_bytes, pn_stream_t stream);
PN_API size_t pn_msda_encoder_workspace_bytes(int B, int nq, int ffn_dims, int num_levels, int num_points);
PN_API int pn_msda_encoder_forward(const PnMsdaEncoderWeights* w, const float* x_in, const float* pos,
                                   const int* h, const int* w_, float* x_out, int B, void* ws, size_t ws_bytes,
                                   pn_stream_t stream);
/* GroupNorm(groups, 256) (+ReLU) of the pixel decoder's ConvModules; x/y [B,256,HW] (NCHW) or [B,HW,256]
 * (channels_last = 1); y may alias x. */
PN_API size_t pn_group_norm_workspace_bytes(int B, int HW, int groups);
PN_API int pn_group_norm(const float* x, const float* gamma, const float* beta, float* y, int B, int HW,
                         int groups, int relu, int channels_last, float eps, void* ws, size_t ws_bytes,
                         pn_stream_t stream);
/* ResNet stem MaxPool2d(3, stride 2, padding 1) on a channels_last map: x [B,H,W,C] -> y [B,(H-1)/2+1,(W-1)/2+1,C] */
PN_API int pn_maxpool3x3s2_nhwc(const float* x, float* y, int B, int H, int W, int C, pn_stream_t stream);
/* GroupNorm fused with the FPN top-down merge of MSDeformAttnPixelDecoder.forward:
 *   y = GN(x) + F.interpolate(top, size=(H,W), mode="bilinear", align_corners=False)
 * x, y channels_last [B,H*W,256] (y may alias x); top token-major: element (b, ty, tx, c) at
 * top[b*top_batch_stride + (ty*w + tx)*256 + c] (a level slice of the encoder output). */
PN_API int pn_gn_upsample_add(const float* x, const float* gamma, const float* beta, const float* top,
                              long long top_batch_stride, float* y, int B, int H, int W, int h, int w, int groups,
                              float eps, void* ws, size_t ws_bytes, pn_stream_t stream);
/* 1x1 convolution of a channels_last map [B,HW,256] into an NCHW-contiguous map [B,cout,HW] (mask_feature):
 * tcgen05 3xTF32 GEMM with the activations split in the SM and a transposed store. */
PN_API size_t pn_conv1x1_nhwc_to_nchw_workspace_bytes(int cout);
PN_API int pn_conv1x1_nhwc_to_nchw(const float* x, const float* w, const float* bias, float* y, int B, int HW,
                                   int cout, void* ws, size_t ws_bytes, pn_stream_t stream);
/* sampling core only: value [B,nq,256], ol [B*nq, 8*L*P*3] (offsets then attention logits) -> out [B*nq,256] */
PN_API int pn_msda_sample(const float* value, const float* ol, float* out, const int* h, const int* w_,
                          int num_levels, int num_points, int B, pn_stream_t stream);

/* ------------------------------------------------------------------ whole hot path
 * CrossHead2.forward minus the pixel decoder (pairnet_head.py:264-417) in one call. */
typedef struct {
  PnM2FWeights m2f;
  PnMlp3 sub_query_update, obj_query_update;
  PnConvTiny update_importance;
  PnRelWeights rel;
} PnHeadWeights;

typedef struct {
  float* cls;          /* [B,N,num_cls] */
  float* mask;         /* [B,N,H4*W4] */
  float* importance;   /* [B,N,N] */
  float* rel;          /* [B,R,num_rel_cls] */
  int64_t* sub_pos;    /* [B,K] */
  int64_t* obj_pos;    /* [B,K] */
  float* sub;          /* [B,K,num_cls]  nullable */
  float* obj;          /* [B,K,num_cls]  nullable */
  float* sub_seg;      /* [B,K,H4*W4]    nullable */
  float* obj_seg;      /* [B,K,H4*W4]    nullable */
  /* optional taps for stage-wise parity tests (all nullable) */
  float* query_out;       /* [B,N,256] */
  float* importance_raw;  /* [B,N,N] */
  float* pair_feat;       /* [B,2K,256] */
  float* rel_feat;        /* [B,R,256] */
  float* query_trace;     /* [layers,B,N,256] */
  uint32_t* mask_trace;   /* [layers,B,N,trace_words] */
  int trace_words;
} PnHeadOutputs;

PN_API size_t pn_head_workspace_bytes(const PnHeadWeights* w, const PnM2FInputs* in);
PN_API int pn_head_forward(const PnHeadWeights* w, const PnM2FInputs* in, const PnHeadOutputs* out,
                    void* ws, size_t ws_bytes, pn_stream_t stream);
/* number of kernels pn_head_forward enqueued on its last call on this thread */
PN_API int pn_last_launch_count(void);
/* profiling hook of the fused decoder-chain kernel: when set (device buffer of `capacity` u64), cluster 0 / rank 0 writes
 * %globaltimer (ns) at kernel start and after every phase barrier.  NULL switches it off (default). */
PN_API int pn_debug_chain_timing(unsigned long long* device_buf, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* PAIRNET_B200_H_ */
