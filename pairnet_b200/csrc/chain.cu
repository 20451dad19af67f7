// Fused query-side decoder chain on tcgen05: ONE launch runs whole mmcv BaseTransformerLayers
// (cross_attn, norm, self_attn, norm, ffn, norm -- pairnet_head.py:353-376 Relation Fusion; :297-320 the
// query side of the Mask2Former decoder) instead of ~17 latency-bound launches per layer.
//
// Mapping.  One thread-block CLUSTER of 8 CTAs per image; CTA rank r of the cluster is attention head r.
//   * every linear is N-split across the cluster: CTA r computes 32 (projections), 2x128 (FFN1) output features,
//     or the K-slice r of FFN2 (split-K partials reduced in fixed order inside the following LayerNorm);
//   * the activations ([rows <= 128 per m-tile, 256] fp32, L2 resident) are the M = 128 operand: TMA stages the RAW
//     fp32 tile, four splitter warps split it hi/lo (3xTF32, fp32 parity) into TENSOR MEMORY and the MMAs take A
//     from TMEM and the pre-split weights (B) from shared memory (tcgen05.mma kind::tf32, TS form);
//   * q.K^T and P.V of head r run on the same tensor pipe (S, P hi/lo and O in TMEM, online softmax by 4 warps);
//   * LayerNorm rows are dealt round-robin to the 8 CTAs (one warp per row);
//   * phases are separated by barrier.cluster (release/acquire) -- activations are exchanged through L2, every
//     intra-kernel read of them bypasses L1 (ld.global.cg / TMA).
// Warp roles (320 threads): warp0 TMA producer, warp1 MMA issuer + TMEM owner, warps2-5 epilogue / softmax,
// warps6-9 A splitter.  All ten warps do LayerNorm rows.
#include <string.h>

#include "umma_ptx.cuh"

namespace pn {
namespace chain {

using namespace umma;

constexpr int CL = 8;                       // CTAs per cluster = attention heads
constexpr int BM = 128, BK = 32;
constexpr int TILE = BM * BK * 4;           // 16 KiB: 128 rows x 128 B
constexpr int STAGES = 4;
constexpr int STAGE_BYTES = 3 * TILE;       // A raw | W hi | W lo   (W slab <= 128 rows)
constexpr int ASETS = 4;                    // TMEM A staging sets (hi 32 + lo 32 columns)
constexpr int TM_A = 256;                   // accumulators [0,128) / [128,256); A sets [256,512)
constexpr int TM_S = 0, TM_PHI = 128, TM_PLO = 256, TM_O = 384;  // attention phase layout
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 320;
constexpr int VT_ATOM = 4096;               // 32 dims x 128 B (32 keys)
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 512;
constexpr float NEG_BIG = -1.0e30f;

enum { EPI_RAW = 0, EPI_SPLIT = 1, EPI_SPLIT_T = 2 };

struct LayerW {                 // one decoder layer, weights pre-split hi/lo (prepared blob), box = 32 x slab rows
  CUtensorMap cin_hi, cin_lo;   // cross in_proj [768,256]   (box rows 32)
  CUtensorMap co_hi, co_lo;     // cross out_proj [256,256]  (32)
  CUtensorMap sin_hi, sin_lo;   // self in_proj [768,256]    (32)
  CUtensorMap so_hi, so_lo;     // self out_proj [256,256]   (32)
  CUtensorMap f1_hi, f1_lo;     // ffn1 [ffn,256]            (128)
  CUtensorMap f2_hi, f2_lo;     // ffn2 [256,ffn]            (128)
  const float *cin_b, *co_b, *sin_b, *so_b, *f1_b, *f2_b;
  const float *gamma[3], *beta[3];
};

struct Params {
  LayerW L[CHAIN_MAX_LAYERS];
  // raw fp32 activations [B*R,256] (h: [B*R,ffn]), box 32 x 128
  CUtensorMap m_x, m_xpos, m_att, m_x1, m_x1pos, m_x2, m_h, m_xn, m_e1, m_e2;
  CUtensorMap m_q_hi, m_q_lo;       // scaled queries [B*R,256]
  CUtensorMap m_kc_hi, m_kc_lo;     // cross keys [B*Nk, nl*256]
  CUtensorMap m_vtc_hi, m_vtc_lo;   // cross V^T [B*nl*256, Nk]  (box 32 x 32)
  CUtensorMap m_ks_hi, m_ks_lo;     // self keys [B*R,256]
  CUtensorMap m_vts_hi, m_vts_lo;   // self V^T [B*256, R]       (box 32 x 32)
  CUtensorMap m_cls_hi, m_cls_lo;   // final classifier [ncls,256] (box 32 x 16)
  CUtensorMap m_me_hi[3], m_me_lo[3];   // mask_embed MLP [256,256] x3 (box 32 x 32)      (m2f mode)
  CUtensorMap m_nq_hi, m_nq_lo;         // next layer's cross in_proj (q rows) [768,256]   (m2f mode)
  float *x, *xpos, *pre, *x1, *x1pos, *x2, *att, *h, *parts, *xn, *e1, *e2;
  float *q_hi, *q_lo, *ks_hi, *ks_lo, *vts_hi, *vts_lo, *e_hi, *e_lo;
  const float *init_feat, *qpos;    // [R,256] learned queries (init_feat null = x/xpos are inputs), query_pos
  const float *cls_b;
  float* cls_out;
  const float *pn_gamma, *pn_beta;  // post_norm (m2f mode)
  const float* me_b[3];
  const float* nq_b;
  float* trace;                     // optional [nl,B*R,256] per-layer x
  int* zero_rows;                   // optional [B*R] cleared by the first LayerNorm (attention-mask rowany flags)
  int B, R, Nk, nl, ffn, ldvs, ncls, ldvc;
  int has_cross_attn;               // 1: relation mode (q-proj + cross attention in-kernel); 0: att is an input
  int m2f_tail;                     // 1: post_norm + mask-embed MLP (+ next layer's q projection) after the layer
  int has_next_q;
};

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// phase boundary: generic-proxy writes (global) become visible to the async proxy (TMA) of every CTA of the cluster
__device__ __forceinline__ void cluster_sync_all() {
  tc_fence_before();
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
  asm volatile("fence.proxy.async;" ::: "memory");
  tc_fence_after();
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

struct Ctx {
  uint8_t* smem;
  uint64_t *full, *empty, *a_ready, *a_free, *tm_full, *tm_empty;
  uint64_t *q_full, *kv_full, *kv_empty, *s_full, *p_ready, *o_full;
  uint32_t tmem;
  int warp, lane, rank, b;
  uint32_t it, tile_it;      // ring k-block counter, accumulator (sub)tile counter
  uint32_t qn, kvn, sn;      // attention: q loads, k/v tiles, S tiles
};

struct EpiArgs {
  int mode;            // EPI_*
  float scale;
  int relu;
  const float* bias;   // indexed by the weight row
  int out_sub;         // output column = weight row - out_sub
  float* dst;          // EPI_RAW / EPI_SPLIT (hi)
  float* dst_lo;
  int ld;
  const float* resid;  // EPI_RAW only (intra-kernel data: read with ld.global.cg)
  int ldr;
  int n_valid;         // weight rows >= n_valid are not stored
};

struct LinArgs {
  const CUtensorMap* a;          // raw activation map
  int a_row0;                    // first row of this image in the activation arrays
  int k0;                        // K offset (columns of A and of W)
  int num_kb;
  int mtiles, rows;              // rows valid per image
  int nsub;                      // subtiles; the weight row offset advances by sub_stride
  int nsrc;
  const CUtensorMap* w_hi[2];
  const CUtensorMap* w_lo[2];
  int w_row0[2], sub_stride[2], ncols[2];
  EpiArgs epi[2];
  int t_row0;                    // EPI_SPLIT_T: first row of this image's V^T block
};

// One N-split linear:  acc[128 x BN] = A[128 x K] . W_slab^T   (3xTF32), then the per-slab epilogues.
__device__ __noinline__ void lin_phase(Ctx& c, const LinArgs& a) {
  const int BN = a.ncols[0] + (a.nsrc > 1 ? a.ncols[1] : 0);
  if (c.warp == 0) {
    if (c.lane == 0) {
      const uint32_t bytes = (uint32_t)TILE + 2u * (uint32_t)BN * 128u;
      for (int mt = 0; mt < a.mtiles; ++mt)
        for (int sub = 0; sub < a.nsub; ++sub)
          for (int kb = 0; kb < a.num_kb; ++kb, ++c.it) {
            const int s = c.it % STAGES;
            mbar_wait(&c.empty[s], ((c.it / STAGES) & 1) ^ 1);
            uint8_t* st = c.smem + (size_t)s * STAGE_BYTES;
            mbar_expect_tx(&c.full[s], bytes);
            tma_load_2d(st, a.a, &c.full[s], a.k0 + kb * BK, a.a_row0 + mt * BM);
            int ofs = 0;
            for (int j = 0; j < a.nsrc; ++j) {
              const int wr = a.w_row0[j] + sub * a.sub_stride[j];
              tma_load_2d(st + TILE + ofs * 128, a.w_hi[j], &c.full[s], a.k0 + kb * BK, wr);
              tma_load_2d(st + 2 * TILE + ofs * 128, a.w_lo[j], &c.full[s], a.k0 + kb * BK, wr);
              ofs += a.ncols[j];
            }
          }
    }
  } else if (c.warp == 1) {
    if (c.lane == 0) {
      const uint32_t idesc = make_idesc(BN);
      for (int mt = 0; mt < a.mtiles; ++mt)
        for (int sub = 0; sub < a.nsub; ++sub, ++c.tile_it) {
          const uint32_t acc = c.tile_it & 1;
          mbar_wait(&c.tm_empty[acc], ((c.tile_it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = c.tmem + acc * 128;
          for (int kb = 0; kb < a.num_kb; ++kb, ++c.it) {
            const int s = c.it % STAGES;
            mbar_wait(&c.full[s], (c.it / STAGES) & 1);
            const uint32_t st = smem_u32(c.smem + (size_t)s * STAGE_BYTES);
            const uint64_t b_hi = make_smem_desc(st + TILE), b_lo = make_smem_desc(st + 2 * TILE);
            const uint32_t set = c.it % ASETS;
            mbar_wait(&c.a_ready[set], (c.it / ASETS) & 1);
            tc_fence_after();
            const uint32_t a_hi_t = c.tmem + TM_A + set * 64, a_lo_t = a_hi_t + 32;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
              const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
              umma_tf32_ts(d_tmem, a_lo_t + k * UMMA_K, b_hi + koff, idesc, first);
              umma_tf32_ts(d_tmem, a_hi_t + k * UMMA_K, b_lo + koff, idesc, 1u);
              umma_tf32_ts(d_tmem, a_hi_t + k * UMMA_K, b_hi + koff, idesc, 1u);
            }
            umma_commit(&c.empty[s]);
            umma_commit(&c.a_free[set]);
          }
          umma_commit(&c.tm_full[acc]);
        }
    }
  } else if (c.warp < 6) {
    // ===== epilogue: TMEM lane quadrant = warp % 4, one output row per thread
    const int quad = c.warp & 3;
    for (int mt = 0; mt < a.mtiles; ++mt)
      for (int sub = 0; sub < a.nsub; ++sub, ++c.tile_it) {
        const uint32_t acc = c.tile_it & 1;
        mbar_wait(&c.tm_full[acc], (c.tile_it >> 1) & 1);
        tc_fence_after();
        const int row = mt * BM + quad * 32 + c.lane;
        const bool row_ok = row < a.rows;
        const size_t grow = (size_t)(a.a_row0 + row);
        int ofs = 0;
        for (int j = 0; j < a.nsrc; ++j) {
          const EpiArgs& e = a.epi[j];
          const int wr0 = a.w_row0[j] + sub * a.sub_stride[j];
          for (int c0 = 0; c0 < a.ncols[j]; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(c.tmem + ((uint32_t)(quad * 32) << 16) + acc * 128 + (uint32_t)(ofs + c0), v);
            if (j == a.nsrc - 1 && c0 + 32 >= a.ncols[j]) {  // last TMEM read of this accumulator
              tc_fence_before();
              __syncwarp();
              if (c.lane == 0) mbar_arrive(&c.tm_empty[acc]);
            }
            if (!row_ok) continue;
            const int wr = wr0 + c0;  // weight row of v[0]
            float y[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) {
              float t = __uint_as_float(v[u]);
              if (e.bias && wr + u < e.n_valid) t += __ldg(e.bias + wr + u);
              t *= e.scale;
              y[u] = e.relu ? fmaxf(t, 0.f) : t;
            }
            const int oc = wr - e.out_sub;  // output column of y[0]
            if (e.mode == EPI_RAW) {
              float* dst = e.dst + grow * e.ld + oc;
              const float* rs = e.resid ? e.resid + grow * e.ldr + oc : nullptr;
#pragma unroll
              for (int u = 0; u < 32; u += 4) {
                if (wr + u >= e.n_valid || c0 + u >= a.ncols[j]) break;
                float4 o = make_float4(y[u], y[u + 1], y[u + 2], y[u + 3]);
                if (rs) {
                  const float4 r4 = ldcg4(rs + u);
                  o.x = r4.x + o.x; o.y = r4.y + o.y; o.z = r4.z + o.z; o.w = r4.w + o.w;
                }
                if (wr + u + 3 < e.n_valid && (e.ld & 3) == 0) {
                  *reinterpret_cast<float4*>(dst + u) = o;
                } else {
                  const float ov[4] = {o.x, o.y, o.z, o.w};
                  for (int t = 0; t < 4; ++t)
                    if (wr + u + t < e.n_valid) dst[u + t] = ov[t];
                }
              }
            } else if (e.mode == EPI_SPLIT) {
              float* dh = e.dst + grow * e.ld + oc;
              float* dl = e.dst_lo + grow * e.ld + oc;
#pragma unroll
              for (int u = 0; u < 32; u += 4) {
                if (c0 + u >= a.ncols[j]) break;
                float hh[4], ll[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  hh[t] = rna_tf32(y[u + t]);
                  ll[t] = rna_tf32(y[u + t] - hh[t]);
                }
                *reinterpret_cast<float4*>(dh + u) = make_float4(hh[0], hh[1], hh[2], hh[3]);
                *reinterpret_cast<float4*>(dl + u) = make_float4(ll[0], ll[1], ll[2], ll[3]);
              }
            } else {  // EPI_SPLIT_T: V^T[(t_row0 + oc + u)][row]; lanes = consecutive rows -> coalesced
#pragma unroll
              for (int u = 0; u < 32; ++u) {
                if (c0 + u >= a.ncols[j]) break;
                const float hh = rna_tf32(y[u]);
                const size_t o = (size_t)(a.t_row0 + oc + u) * e.ld + row;
                e.dst[o] = hh;
                e.dst_lo[o] = rna_tf32(y[u] - hh);
              }
            }
          }
          ofs += a.ncols[j];
        }
      }
  } else {
    // ===== A splitter: raw fp32 tile (smem, SWIZZLE_128B) -> hi / lo in tensor memory, one row per thread
    const int quad = c.warp & 3;
    const int r = quad * 32 + c.lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    for (int mt = 0; mt < a.mtiles; ++mt)
      for (int sub = 0; sub < a.nsub; ++sub)
        for (int kb = 0; kb < a.num_kb; ++kb, ++c.it) {
          const int s = c.it % STAGES;
          const uint32_t set = c.it % ASETS;
          mbar_wait(&c.full[s], (c.it / STAGES) & 1);
          mbar_wait(&c.a_free[set], ((c.it / ASETS) & 1) ^ 1);
          tc_fence_after();
          const uint8_t* arow = c.smem + (size_t)s * STAGE_BYTES + r * 128;
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const float4 v = *reinterpret_cast<const float4*>(arow + ((ch ^ (r & 7)) << 4));
            const float xx[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float hh = rna_tf32(xx[u]);
              hi[ch * 4 + u] = __float_as_uint(hh);
              lo[ch * 4 + u] = __float_as_uint(rna_tf32(xx[u] - hh));
            }
          }
          tmem_st_32x32b_x32(c.tmem + lane_addr + TM_A + set * 64, hi);
          tmem_st_32x32b_x32(c.tmem + lane_addr + TM_A + set * 64 + 32, lo);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (c.lane == 0) {
            mbar_arrive(&c.a_ready[set]);
            mbar_arrive(&c.empty[s]);
          }
        }
  }
  cluster_sync_all();
}

struct AttnArgs {
  const CUtensorMap *q_hi, *q_lo, *k_hi, *k_lo, *vt_hi, *vt_lo;
  int q_row0, q_col;     // queries: rows of this image, column = head * 32
  int k_row0, k_col;     // keys
  int vt_row;            // V^T row of (image, head, dim 0)
  int Nk, rows, mtiles;
  float* out;            // att [B*R,256]; this CTA writes columns [q_col, q_col + 32)
};

// softmax(q_h K_h^T) V_h of head = cluster rank (no mask), 128-key tiles, online softmax; see fa_umma.cu
__device__ __noinline__ void attn_phase(Ctx& c, const AttnArgs& a) {
  uint8_t* q_hi_s = c.smem;
  uint8_t* q_lo_s = c.smem + TILE;
  uint8_t* stage0 = c.smem + 2 * TILE;
  constexpr int KV_STAGE = 4 * TILE;
  const int ntiles = (a.Nk + 127) / 128;
  if (c.warp == 0) {
    if (c.lane == 0) {
      for (int mt = 0; mt < a.mtiles; ++mt) {
        if (mt > 0) {  // the MMAs that read the previous q tile have completed
          const uint32_t prev = c.kvn - 1;
          mbar_wait(&c.kv_empty[prev & 1], (prev >> 1) & 1);
        }
        mbar_expect_tx(c.q_full, 2 * TILE);
        tma_load_2d(q_hi_s, a.q_hi, c.q_full, a.q_col, a.q_row0 + mt * BM);
        tma_load_2d(q_lo_s, a.q_lo, c.q_full, a.q_col, a.q_row0 + mt * BM);
        for (int t = 0; t < ntiles; ++t, ++c.kvn) {
          const int s = c.kvn & 1;
          mbar_wait(&c.kv_empty[s], ((c.kvn >> 1) & 1) ^ 1);
          uint8_t* st = stage0 + (size_t)s * KV_STAGE;
          mbar_expect_tx(&c.kv_full[s], KV_STAGE);
          tma_load_2d(st, a.k_hi, &c.kv_full[s], a.k_col, a.k_row0 + t * 128);
          tma_load_2d(st + TILE, a.k_lo, &c.kv_full[s], a.k_col, a.k_row0 + t * 128);
#pragma unroll
          for (int at = 0; at < 4; ++at) {
            tma_load_2d(st + 2 * TILE + at * VT_ATOM, a.vt_hi, &c.kv_full[s], t * 128 + at * 32, a.vt_row);
            tma_load_2d(st + 3 * TILE + at * VT_ATOM, a.vt_lo, &c.kv_full[s], t * 128 + at * 32, a.vt_row);
          }
        }
      }
    }
  } else if (c.warp == 1) {
    if (c.lane == 0) {
      const uint32_t idesc_s = make_idesc(128);
      const uint32_t idesc_o = make_idesc(HD);
      const uint64_t dq_hi = make_smem_desc(smem_u32(q_hi_s)), dq_lo = make_smem_desc(smem_u32(q_lo_s));
      for (int mt = 0; mt < a.mtiles; ++mt) {
        mbar_wait(c.q_full, c.qn & 1);
        ++c.qn;
        for (int t = 0; t < ntiles; ++t, ++c.kvn, ++c.sn) {
          const int s = c.kvn & 1;
          mbar_wait(&c.kv_full[s], (c.kvn >> 1) & 1);
          tc_fence_after();
          const uint32_t st = smem_u32(stage0 + (size_t)s * KV_STAGE);
          const uint64_t dk_hi = make_smem_desc(st), dk_lo = make_smem_desc(st + TILE);
#pragma unroll
          for (int k = 0; k < HD / UMMA_K; ++k) {
            const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
            umma_tf32(c.tmem + TM_S, dq_lo + koff, dk_hi + koff, idesc_s, k == 0 ? 0u : 1u);
            umma_tf32(c.tmem + TM_S, dq_hi + koff, dk_lo + koff, idesc_s, 1u);
            umma_tf32(c.tmem + TM_S, dq_hi + koff, dk_hi + koff, idesc_s, 1u);
          }
          umma_commit(c.s_full);
          mbar_wait(c.p_ready, c.sn & 1);
          tc_fence_after();
#pragma unroll
          for (int at = 0; at < 4; ++at) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t col = (uint32_t)(at * 32 + k * UMMA_K);
              const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
              const uint64_t dv_hi = make_smem_desc(st + 2 * TILE + at * VT_ATOM) + koff;
              const uint64_t dv_lo = make_smem_desc(st + 3 * TILE + at * VT_ATOM) + koff;
              umma_tf32_ts(c.tmem + TM_O, c.tmem + TM_PLO + col, dv_hi, idesc_o, (at == 0 && k == 0) ? 0u : 1u);
              umma_tf32_ts(c.tmem + TM_O, c.tmem + TM_PHI + col, dv_lo, idesc_o, 1u);
              umma_tf32_ts(c.tmem + TM_O, c.tmem + TM_PHI + col, dv_hi, idesc_o, 1u);
            }
          }
          umma_commit(c.o_full);
          umma_commit(&c.kv_empty[s]);
        }
      }
    }
  } else if (c.warp < 6) {
    const int quad = c.warp & 3;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    for (int mt = 0; mt < a.mtiles; ++mt) {
      const int row = mt * BM + quad * 32 + c.lane;
      float o[HD];
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] = 0.f;
      float m_run = NEG_BIG, l_run = 0.f;
      for (int t = 0; t < ntiles; ++t, ++c.sn) {
        const int nvalid = min(128, a.Nk - t * 128);
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rem = nvalid - i * 32;
          w[i] = rem <= 0 ? 0xffffffffu : (rem < 32 ? (0xffffffffu << rem) : 0u);
        }
        mbar_wait(c.s_full, c.sn & 1);
        tc_fence_after();
        float cmax = NEG_BIG;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(c.tmem + lane_addr + TM_S + ch * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (!((w[ch] >> j) & 1u)) cmax = fmaxf(cmax, __uint_as_float(v[j]));
        }
        const float m_new = fmaxf(m_run, cmax);
        const float corr = fast_exp2(m_run - m_new);
        float lsum = 0.f;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t v[32], ph[32], pl[32];
          tmem_ld_32x32b_x32(c.tmem + lane_addr + TM_S + ch * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float p = ((w[ch] >> j) & 1u) ? 0.f : fast_exp2(__uint_as_float(v[j]) - m_new);
            lsum += p;
            const float hh = rna_tf32(p);
            ph[j] = __float_as_uint(hh);
            pl[j] = __float_as_uint(rna_tf32(p - hh));
          }
          tmem_st_32x32b_x32(c.tmem + lane_addr + TM_PHI + ch * 32, ph);
          tmem_st_32x32b_x32(c.tmem + lane_addr + TM_PLO + ch * 32, pl);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (c.lane == 0) mbar_arrive(c.p_ready);
        l_run = l_run * corr + lsum;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] *= corr;
        m_run = m_new;
        mbar_wait(c.o_full, c.sn & 1);
        tc_fence_after();
        {
          uint32_t v[32];
          tmem_ld_32x32b_x32(c.tmem + lane_addr + TM_O, v);
#pragma unroll
          for (int d = 0; d < HD; ++d) o[d] += __uint_as_float(v[d]);
        }
      }
      if (row < a.rows) {
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        float4* op = reinterpret_cast<float4*>(a.out + (size_t)(a.q_row0 + row) * D + a.q_col);
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4)
          op[d4] = make_float4(o[d4 * 4] * inv, o[d4 * 4 + 1] * inv, o[d4 * 4 + 2] * inv, o[d4 * 4 + 3] * inv);
      }
    }
  }
  cluster_sync_all();
}

struct LnPhaseArgs {
  const float* x;          // [nparts][B*R][256] (intra-kernel data)
  int nparts;
  long long part_stride;
  const float* bias;
  const float* resid;      // intra-kernel data
  const float *gamma, *beta;
  float* y;
  const float* pos;        // [R,256] or null
  float* ypos;
  const float *gamma2, *beta2;
  float* y2;
  float* trace;            // optional copy of y
  int* zero_rows;
  int row0, rows;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void ld_row_cg(float (&v)[8], const float* p, int lane) {
  const float4 a = ldcg4(p + 4 * lane), b = ldcg4(p + 128 + 4 * lane);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld_row_const(float (&v)[8], const float* p, int lane) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p) + lane), b = __ldg(reinterpret_cast<const float4*>(p) + 32 + lane);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st_row(const float (&v)[8], float* p, int lane) {
  reinterpret_cast<float4*>(p)[lane] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[32 + lane] = make_float4(v[4], v[5], v[6], v[7]);
}
// same arithmetic as rowops.cu ln_row: two-pass mean / variance, eps 1e-5
__device__ __forceinline__ void ln_row8(float (&v)[8], const float* gamma, const float* beta, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float d = v[i] - mean;
    q += d * d;
  }
  const float var = warp_sum(q) * (1.f / D);
  const float rstd = 1.f / sqrtf(var + 1e-5f);
  float g[8], bb[8];
  ld_row_const(g, gamma, lane);
  ld_row_const(bb, beta, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * g[i] + bb[i];
}

// rows of the image are dealt round-robin to the CTAs of the cluster, one warp per row
__device__ __noinline__ void ln_phase(Ctx& c, const LnPhaseArgs& a) {
  for (int row = c.rank + CL * c.warp; row < a.rows; row += CL * (NUM_THREADS / 32)) {
    const size_t m = (size_t)(a.row0 + row);
    if (a.zero_rows && c.lane == 0) a.zero_rows[m] = 0;
    float v[8];
    ld_row_cg(v, a.x + m * D, c.lane);
    for (int s = 1; s < a.nparts; ++s) {
      float t[8];
      ld_row_cg(t, a.x + (size_t)s * a.part_stride + m * D, c.lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
    if (a.bias) {
      float t[8];
      ld_row_const(t, a.bias, c.lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
    if (a.resid) {
      float t[8];
      ld_row_cg(t, a.resid + m * D, c.lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = t[i] + v[i];
    }
    ln_row8(v, a.gamma, a.beta, c.lane);
    st_row(v, a.y + m * D, c.lane);
    if (a.trace) st_row(v, a.trace + m * D, c.lane);
    if (a.ypos) {
      float t[8];
      ld_row_const(t, a.pos + (size_t)row * D, c.lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = v[i] + t[i];
      st_row(t, a.ypos + m * D, c.lane);
    }
    if (a.y2) {
      ln_row8(v, a.gamma2, a.beta2, c.lane);
      st_row(v, a.y2 + m * D, c.lane);
    }
  }
  cluster_sync_all();
}

__device__ __forceinline__ EpiArgs epi_raw(const float* bias, float* dst, int ld, int relu, const float* resid, int ldr,
                                           int n_valid, int out_sub = 0) {
  EpiArgs e;
  e.mode = EPI_RAW; e.scale = 1.f; e.relu = relu; e.bias = bias; e.out_sub = out_sub; e.dst = dst; e.dst_lo = nullptr;
  e.ld = ld; e.resid = resid; e.ldr = ldr; e.n_valid = n_valid;
  return e;
}
__device__ __forceinline__ EpiArgs epi_split(int mode, const float* bias, float* hi, float* lo, int ld, float scale,
                                             int out_sub, int n_valid) {
  EpiArgs e;
  e.mode = mode; e.scale = scale; e.relu = 0; e.bias = bias; e.out_sub = out_sub; e.dst = hi; e.dst_lo = lo; e.ld = ld;
  e.resid = nullptr; e.ldr = 0; e.n_valid = n_valid;
  return e;
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NUM_THREADS, 1)
decoder_chain_kernel(const __grid_constant__ Params prm) {
  extern __shared__ uint8_t smem_raw[];
  Ctx c;
  c.smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(c.smem + STAGES * STAGE_BYTES);
  c.full = bars;                      // [STAGES]
  c.empty = c.full + STAGES;          // [STAGES]
  c.a_ready = c.empty + STAGES;       // [ASETS]
  c.a_free = c.a_ready + ASETS;       // [ASETS]
  c.tm_full = c.a_free + ASETS;       // [2]
  c.tm_empty = c.tm_full + 2;         // [2]
  c.q_full = c.tm_empty + 2;
  c.kv_full = c.q_full + 1;           // [2]
  c.kv_empty = c.kv_full + 2;         // [2]
  c.s_full = c.kv_empty + 2;
  c.p_ready = c.s_full + 1;
  c.o_full = c.p_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c.o_full + 1);
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.rank = (int)cluster_rank();
  c.b = blockIdx.x / CL;
  c.it = c.tile_it = c.qn = c.kvn = c.sn = 0;

  if (c.warp == 0 && c.lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&c.full[s], 1); mbar_init(&c.empty[s], 5); }
    for (int s = 0; s < ASETS; ++s) { mbar_init(&c.a_ready[s], 4); mbar_init(&c.a_free[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&c.tm_full[s], 1); mbar_init(&c.tm_empty[s], 4);
      mbar_init(&c.kv_full[s], 1); mbar_init(&c.kv_empty[s], 1);
    }
    mbar_init(c.q_full, 1);
    mbar_init(c.s_full, 1);
    mbar_init(c.p_ready, 4);
    mbar_init(c.o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (c.warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *tmem_slot;

  const int R = prm.R, b = c.b, h = c.rank;
  const int row0 = b * R;
  const int mtiles = (R + BM - 1) / BM;
  const size_t MD = (size_t)prm.B * R * D;

  // ---- learned queries (pairnet_head.py:353-364): x = feat, xpos = feat + query_pos
  if (prm.init_feat) {
    for (int row = c.rank + CL * c.warp; row < R; row += CL * (NUM_THREADS / 32)) {
      float v[8], t[8];
      ld_row_const(v, prm.init_feat + (size_t)row * D, c.lane);
      ld_row_const(t, prm.qpos + (size_t)row * D, c.lane);
      st_row(v, prm.x + (size_t)(row0 + row) * D, c.lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = v[i] + t[i];
      st_row(t, prm.xpos + (size_t)(row0 + row) * D, c.lane);
    }
    cluster_sync_all();
  }

  for (int l = 0; l < prm.nl; ++l) {
    const LayerW& W = prm.L[l];
    if (prm.has_cross_attn) {
      // ---- q = ((x + qpos) Wq^T + bq) * scale, head h -> q_hi / q_lo
      LinArgs a{};
      a.a = &prm.m_xpos; a.a_row0 = row0; a.k0 = 0; a.num_kb = D / BK; a.mtiles = mtiles; a.rows = R; a.nsub = 1; a.nsrc = 1;
      a.w_hi[0] = &W.cin_hi; a.w_lo[0] = &W.cin_lo; a.w_row0[0] = h * HD; a.sub_stride[0] = 0; a.ncols[0] = HD;
      a.epi[0] = epi_split(EPI_SPLIT, W.cin_b, prm.q_hi, prm.q_lo, D, ATTN_QSCALE, 0, D);
      lin_phase(c, a);
      AttnArgs t{};
      t.q_hi = &prm.m_q_hi; t.q_lo = &prm.m_q_lo; t.k_hi = &prm.m_kc_hi; t.k_lo = &prm.m_kc_lo;
      t.vt_hi = &prm.m_vtc_hi; t.vt_lo = &prm.m_vtc_lo;
      t.q_row0 = row0; t.q_col = h * HD; t.k_row0 = b * prm.Nk; t.k_col = l * D + h * HD;
      t.vt_row = (b * prm.nl + l) * D + h * HD; t.Nk = prm.Nk; t.rows = R; t.mtiles = mtiles; t.out = prm.att;
      attn_phase(c, t);
    }
    {  // ---- cross-attention output projection + residual -> pre ; LN0 -> x1, x1pos
      LinArgs a{};
      a.a = &prm.m_att; a.a_row0 = row0; a.k0 = 0; a.num_kb = D / BK; a.mtiles = mtiles; a.rows = R; a.nsub = 1; a.nsrc = 1;
      a.w_hi[0] = &W.co_hi; a.w_lo[0] = &W.co_lo; a.w_row0[0] = h * HD; a.sub_stride[0] = 0; a.ncols[0] = HD;
      a.epi[0] = epi_raw(W.co_b, prm.pre, D, 0, prm.x, D, D);
      lin_phase(c, a);
      LnPhaseArgs n{};
      n.x = prm.pre; n.nparts = 1; n.gamma = W.gamma[0]; n.beta = W.beta[0]; n.y = prm.x1; n.pos = prm.qpos; n.ypos = prm.x1pos;
      n.row0 = row0; n.rows = R; n.zero_rows = prm.zero_rows;
      ln_phase(c, n);
    }
    {  // ---- self attention: q, k = (x1 + qpos) W{q,k}^T ; v = x1 Wv^T (stored transposed)
      LinArgs a{};
      a.a = &prm.m_x1pos; a.a_row0 = row0; a.k0 = 0; a.num_kb = D / BK; a.mtiles = mtiles; a.rows = R; a.nsub = 1; a.nsrc = 2;
      a.w_hi[0] = &W.sin_hi; a.w_lo[0] = &W.sin_lo; a.w_row0[0] = h * HD; a.sub_stride[0] = 0; a.ncols[0] = HD;
      a.w_hi[1] = &W.sin_hi; a.w_lo[1] = &W.sin_lo; a.w_row0[1] = D + h * HD; a.sub_stride[1] = 0; a.ncols[1] = HD;
      a.epi[0] = epi_split(EPI_SPLIT, W.sin_b, prm.q_hi, prm.q_lo, D, ATTN_QSCALE, 0, 3 * D);
      a.epi[1] = epi_split(EPI_SPLIT, W.sin_b, prm.ks_hi, prm.ks_lo, D, 1.f, D, 3 * D);
      lin_phase(c, a);
      LinArgs v{};
      v.a = &prm.m_x1; v.a_row0 = row0; v.k0 = 0; v.num_kb = D / BK; v.mtiles = mtiles; v.rows = R; v.nsub = 1; v.nsrc = 1;
      v.w_hi[0] = &W.sin_hi; v.w_lo[0] = &W.sin_lo; v.w_row0[0] = 2 * D + h * HD; v.sub_stride[0] = 0; v.ncols[0] = HD;
      v.epi[0] = epi_split(EPI_SPLIT_T, W.sin_b, prm.vts_hi, prm.vts_lo, prm.ldvs, 1.f, 2 * D, 3 * D);
      v.t_row0 = b * D;
      lin_phase(c, v);
      AttnArgs t{};
      t.q_hi = &prm.m_q_hi; t.q_lo = &prm.m_q_lo; t.k_hi = &prm.m_ks_hi; t.k_lo = &prm.m_ks_lo;
      t.vt_hi = &prm.m_vts_hi; t.vt_lo = &prm.m_vts_lo;
      t.q_row0 = row0; t.q_col = h * HD; t.k_row0 = row0; t.k_col = h * HD; t.vt_row = b * D + h * HD;
      t.Nk = R; t.rows = R; t.mtiles = mtiles; t.out = prm.att;
      attn_phase(c, t);
      LinArgs o{};
      o.a = &prm.m_att; o.a_row0 = row0; o.k0 = 0; o.num_kb = D / BK; o.mtiles = mtiles; o.rows = R; o.nsub = 1; o.nsrc = 1;
      o.w_hi[0] = &W.so_hi; o.w_lo[0] = &W.so_lo; o.w_row0[0] = h * HD; o.sub_stride[0] = 0; o.ncols[0] = HD;
      o.epi[0] = epi_raw(W.so_b, prm.pre, D, 0, prm.x1, D, D);
      lin_phase(c, o);
      LnPhaseArgs n{};
      n.x = prm.pre; n.nparts = 1; n.gamma = W.gamma[1]; n.beta = W.beta[1]; n.y = prm.x2; n.row0 = row0; n.rows = R;
      ln_phase(c, n);
    }
    {  // ---- FFN: hidden slice of this CTA (ffn / 8 features), then its split-K partial of the output
      const int fs = prm.ffn / CL;  // hidden features per CTA (multiple of 128)
      LinArgs a{};
      a.a = &prm.m_x2; a.a_row0 = row0; a.k0 = 0; a.num_kb = D / BK; a.mtiles = mtiles; a.rows = R; a.nsub = fs / 128; a.nsrc = 1;
      a.w_hi[0] = &W.f1_hi; a.w_lo[0] = &W.f1_lo; a.w_row0[0] = h * fs; a.sub_stride[0] = 128; a.ncols[0] = 128;
      a.epi[0] = epi_raw(W.f1_b, prm.h, prm.ffn, 1, nullptr, 0, prm.ffn);
      lin_phase(c, a);
      LinArgs f{};
      f.a = &prm.m_h; f.a_row0 = row0; f.k0 = h * fs; f.num_kb = fs / BK; f.mtiles = mtiles; f.rows = R; f.nsub = D / 128; f.nsrc = 1;
      f.w_hi[0] = &W.f2_hi; f.w_lo[0] = &W.f2_lo; f.w_row0[0] = 0; f.sub_stride[0] = 128; f.ncols[0] = 128;
      f.epi[0] = epi_raw(nullptr, prm.parts + (size_t)h * MD, D, 0, nullptr, 0, D);
      lin_phase(c, f);
      LnPhaseArgs n{};
      n.x = prm.parts; n.nparts = CL; n.part_stride = (long long)MD; n.bias = W.f2_b; n.resid = prm.x2;
      n.gamma = W.gamma[2]; n.beta = W.beta[2]; n.y = prm.x; n.pos = prm.qpos; n.ypos = prm.xpos; n.row0 = row0; n.rows = R;
      if (prm.trace) n.trace = prm.trace + (size_t)l * MD;
      if (prm.m2f_tail) { n.gamma2 = prm.pn_gamma; n.beta2 = prm.pn_beta; n.y2 = prm.xn; }
      ln_phase(c, n);
    }
  }

  if (prm.m2f_tail) {
    // ---- forward_head, mask branch (pairnet_head.py:236-243): e = mask_embed(post_norm(x)), emitted pre-split
    //      for the attention-mask GEMM; then the next layer's cross-attention query projection
    const CUtensorMap* src[3] = {&prm.m_xn, &prm.m_e1, &prm.m_e2};
    float* dst[2] = {prm.e1, prm.e2};
    for (int s = 0; s < 3; ++s) {
      LinArgs a{};
      a.a = src[s]; a.a_row0 = row0; a.k0 = 0; a.num_kb = D / BK; a.mtiles = mtiles; a.rows = R; a.nsub = 1; a.nsrc = 1;
      a.w_hi[0] = &prm.m_me_hi[s]; a.w_lo[0] = &prm.m_me_lo[s]; a.w_row0[0] = h * HD; a.sub_stride[0] = 0; a.ncols[0] = HD;
      if (s < 2) a.epi[0] = epi_raw(prm.me_b[s], dst[s], D, 1, nullptr, 0, D);
      else a.epi[0] = epi_split(EPI_SPLIT, prm.me_b[2], prm.e_hi, prm.e_lo, D, 1.f, 0, D);
      lin_phase(c, a);
    }
    if (prm.has_next_q) {
      LinArgs a{};
      a.a = &prm.m_xpos; a.a_row0 = row0; a.k0 = 0; a.num_kb = D / BK; a.mtiles = mtiles; a.rows = R; a.nsub = 1; a.nsrc = 1;
      a.w_hi[0] = &prm.m_nq_hi; a.w_lo[0] = &prm.m_nq_lo; a.w_row0[0] = h * HD; a.sub_stride[0] = 0; a.ncols[0] = HD;
      a.epi[0] = epi_split(EPI_SPLIT, prm.nq_b, prm.q_hi, prm.q_lo, D, ATTN_QSCALE, 0, D);
      lin_phase(c, a);
    }
  }

  if (prm.cls_out) {
    // ---- relation classifier (pairnet_head.py:377-378): 16 classes per CTA
    LinArgs a{};
    a.a = &prm.m_x; a.a_row0 = row0; a.k0 = 0; a.num_kb = D / BK; a.mtiles = mtiles; a.rows = R; a.nsub = 1; a.nsrc = 1;
    a.w_hi[0] = &prm.m_cls_hi; a.w_lo[0] = &prm.m_cls_lo; a.w_row0[0] = h * 16; a.sub_stride[0] = 0; a.ncols[0] = 16;
    a.epi[0] = epi_raw(prm.cls_b, prm.cls_out, prm.ncls, 0, nullptr, 0, prm.ncls);
    if (h * 16 < prm.ncls) lin_phase(c, a);
    else cluster_sync_all();
  }

  tc_fence_before();
  __syncthreads();
  if (c.warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace chain

// ================================================================================================ host side
static int chain_map(CUtensorMap* m, const float* p, long long rows, long long cols, long long ld, int box_rows) {
  return umma::make_tmap_2d(m, p, rows, cols, ld, 32, box_rows);
}

int launch_decoder_chain(const ChainArgs& g, cudaStream_t st) {
  using namespace chain;
  PN_REQUIRE(g.nl >= 1 && g.nl <= CHAIN_MAX_LAYERS, PN_ERR_BAD_ARG, "chain: 1..%d layers per launch", CHAIN_MAX_LAYERS);
  PN_REQUIRE(g.B > 0 && g.R > 0 && g.R <= 1024, PN_ERR_BAD_ARG, "chain: bad B/R");
  PN_REQUIRE(g.ffn % (CL * 128) == 0, PN_ERR_UNSUPPORTED, "chain: ffn_dims=%d must be a multiple of %d", g.ffn, CL * 128);
  PN_REQUIRE(!g.cls_out || g.ncls <= CL * 16, PN_ERR_UNSUPPORTED, "chain: at most %d classes in the fused classifier", CL * 16);
  PN_REQUIRE(g.ldvs % 4 == 0 && g.ldvs >= g.R, PN_ERR_BAD_ARG, "chain: bad ldvs");
  Params prm;  // ~16 KB kernel parameter block
  memset(&prm, 0, sizeof(prm));
  const long long M = (long long)g.B * g.R;
  for (int l = 0; l < g.nl; ++l) {
    const ChainLayer& s = g.layers[l];
    LayerW& W = prm.L[l];
    if (g.has_cross_attn) {
      PN_TRY(chain_map(&W.cin_hi, s.cin_hi, 3 * D, D, D, 32));
      PN_TRY(chain_map(&W.cin_lo, s.cin_lo, 3 * D, D, D, 32));
    }
    PN_TRY(chain_map(&W.co_hi, s.co_hi, D, D, D, 32));
    PN_TRY(chain_map(&W.co_lo, s.co_lo, D, D, D, 32));
    PN_TRY(chain_map(&W.sin_hi, s.sin_hi, 3 * D, D, D, 32));
    PN_TRY(chain_map(&W.sin_lo, s.sin_lo, 3 * D, D, D, 32));
    PN_TRY(chain_map(&W.so_hi, s.so_hi, D, D, D, 32));
    PN_TRY(chain_map(&W.so_lo, s.so_lo, D, D, D, 32));
    PN_TRY(chain_map(&W.f1_hi, s.f1_hi, g.ffn, D, D, 128));
    PN_TRY(chain_map(&W.f1_lo, s.f1_lo, g.ffn, D, D, 128));
    PN_TRY(chain_map(&W.f2_hi, s.f2_hi, D, g.ffn, g.ffn, 128));
    PN_TRY(chain_map(&W.f2_lo, s.f2_lo, D, g.ffn, g.ffn, 128));
    W.cin_b = s.cin_b; W.co_b = s.co_b; W.sin_b = s.sin_b; W.so_b = s.so_b; W.f1_b = s.f1_b; W.f2_b = s.f2_b;
    for (int i = 0; i < 3; ++i) { W.gamma[i] = s.gamma[i]; W.beta[i] = s.beta[i]; }
  }
  PN_TRY(chain_map(&prm.m_x, g.x, M, D, D, 128));
  PN_TRY(chain_map(&prm.m_xpos, g.xpos, M, D, D, 128));
  PN_TRY(chain_map(&prm.m_att, g.att, M, D, D, 128));
  PN_TRY(chain_map(&prm.m_x1, g.x1, M, D, D, 128));
  PN_TRY(chain_map(&prm.m_x1pos, g.x1pos, M, D, D, 128));
  PN_TRY(chain_map(&prm.m_x2, g.x2, M, D, D, 128));
  PN_TRY(chain_map(&prm.m_h, g.h, M, g.ffn, g.ffn, 128));
  PN_TRY(chain_map(&prm.m_q_hi, g.q_hi, M, D, D, 128));
  PN_TRY(chain_map(&prm.m_q_lo, g.q_lo, M, D, D, 128));
  PN_TRY(chain_map(&prm.m_ks_hi, g.ks_hi, M, D, D, 128));
  PN_TRY(chain_map(&prm.m_ks_lo, g.ks_lo, M, D, D, 128));
  PN_TRY(chain_map(&prm.m_vts_hi, g.vts_hi, (long long)g.B * D, g.R, g.ldvs, 32));
  PN_TRY(chain_map(&prm.m_vts_lo, g.vts_lo, (long long)g.B * D, g.R, g.ldvs, 32));
  if (g.has_cross_attn) {
    PN_REQUIRE(g.kc_hi && g.kc_lo && g.vtc_hi && g.vtc_lo && g.Nk > 0 && g.ldvc % 4 == 0 && g.ldvc >= g.Nk, PN_ERR_BAD_ARG,
               "chain: cross-attention operands missing");
    PN_TRY(chain_map(&prm.m_kc_hi, g.kc_hi, (long long)g.B * g.Nk, (long long)g.nl * D, (long long)g.nl * D, 128));
    PN_TRY(chain_map(&prm.m_kc_lo, g.kc_lo, (long long)g.B * g.Nk, (long long)g.nl * D, (long long)g.nl * D, 128));
    PN_TRY(chain_map(&prm.m_vtc_hi, g.vtc_hi, (long long)g.B * g.nl * D, g.Nk, g.ldvc, 32));
    PN_TRY(chain_map(&prm.m_vtc_lo, g.vtc_lo, (long long)g.B * g.nl * D, g.Nk, g.ldvc, 32));
  }
  if (g.cls_out) {
    PN_TRY(chain_map(&prm.m_cls_hi, g.cls_hi, g.ncls, D, D, 16));
    PN_TRY(chain_map(&prm.m_cls_lo, g.cls_lo, g.ncls, D, D, 16));
  }
  if (g.m2f_tail) {
    PN_REQUIRE(g.xn && g.e1 && g.e2 && g.e_hi && g.e_lo, PN_ERR_BAD_ARG, "chain: m2f tail buffers missing");
    PN_TRY(chain_map(&prm.m_xn, g.xn, M, D, D, 128));
    PN_TRY(chain_map(&prm.m_e1, g.e1, M, D, D, 128));
    PN_TRY(chain_map(&prm.m_e2, g.e2, M, D, D, 128));
    for (int i = 0; i < 3; ++i) {
      PN_TRY(chain_map(&prm.m_me_hi[i], g.me_hi[i], D, D, D, 32));
      PN_TRY(chain_map(&prm.m_me_lo[i], g.me_lo[i], D, D, D, 32));
      prm.me_b[i] = g.me_b[i];
    }
    if (g.nq_hi) {
      PN_TRY(chain_map(&prm.m_nq_hi, g.nq_hi, 3 * D, D, D, 32));
      PN_TRY(chain_map(&prm.m_nq_lo, g.nq_lo, 3 * D, D, D, 32));
      prm.nq_b = g.nq_b;
      prm.has_next_q = 1;
    }
  }
  prm.x = g.x; prm.xpos = g.xpos; prm.pre = g.pre; prm.x1 = g.x1; prm.x1pos = g.x1pos; prm.x2 = g.x2; prm.att = g.att;
  prm.h = g.h; prm.parts = g.parts; prm.xn = g.xn; prm.e1 = g.e1; prm.e2 = g.e2;
  prm.q_hi = g.q_hi; prm.q_lo = g.q_lo; prm.ks_hi = g.ks_hi; prm.ks_lo = g.ks_lo; prm.vts_hi = g.vts_hi; prm.vts_lo = g.vts_lo;
  prm.e_hi = g.e_hi; prm.e_lo = g.e_lo;
  prm.init_feat = g.init_feat; prm.qpos = g.qpos; prm.cls_b = g.cls_b; prm.cls_out = g.cls_out;
  prm.pn_gamma = g.pn_gamma; prm.pn_beta = g.pn_beta; prm.trace = g.trace; prm.zero_rows = g.zero_rows;
  prm.B = g.B; prm.R = g.R; prm.Nk = g.Nk; prm.nl = g.nl; prm.ffn = g.ffn; prm.ldvs = g.ldvs; prm.ncls = g.ncls;
  prm.ldvc = g.ldvc; prm.has_cross_attn = g.has_cross_attn; prm.m2f_tail = g.m2f_tail;
  cudaError_t e = cudaFuncSetAttribute(decoder_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  PN_REQUIRE(e == cudaSuccess, (int)e, "chain: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  decoder_chain_kernel<<<dim3(CL * g.B), NUM_THREADS, SMEM_BYTES, st>>>(prm);
  return check_launch("decoder_chain_kernel");
}

size_t chain_scratch_floats(int B, int R, int ffn) {
  const size_t M = (size_t)B * R;
  const size_t ldvs = (size_t)round_up(R, 4);
  // x1, x1pos, x2, pre, att (5 x [M,256]); h [M,ffn]; parts [8][M,256]; q/ks hi+lo (4 x [M,256]); vts hi+lo
  return M * D * 5 + M * ffn + (size_t)chain::CL * M * D + M * D * 4 + 2 * (size_t)B * D * ldvs + 64 * 16;
}

}  // namespace pn
