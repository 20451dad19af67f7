// Pair Proposal Network at scale (BASELINE config 5a): pair matrix + top-k pair select in ONE pass over HBM.
//
//   importance[b] = S[b] . O[b]^T              (pairnet_head.py:327)
//   idx = topk(importance[b].flatten(), K)     (pairnet_head.py:334-336; descending, ties by ascending flat index)
//   sub_pos = idx / N, obj_pos = idx % N       (pairnet_head.py:337-340)
//
//   S, O : [B, N, 256] fp32 (L2-normalised subject / object embeddings);  importance [B, N, N] fp32;  int64 indices.
//
// One persistent CTA per SM owns whole images.  The matrix of an image is produced tile by tile (128 rows x bn <= 256
// columns) on tcgen05 as 3xTF32 and is written to HBM exactly once; the top-k never reads it back:
//
//   warp 0      TMA producer: raw fp32 k-blocks (32 channels = one 128-byte swizzle row) of S and O through 3-D maps
//               [B][N][256]; rows past N are zero-filled by the TMA unit (no padded copy, no bytes of the next image).
//               The raw ring is as deep as shared memory allows (5 stages at N = 100): the kernel is bound by bytes in
//               flight per SM, so raw tiles are kept apart from the short-lived lo tiles (2 slots).
//   warps 6-13  splitters: the tensor pipe reads fp32 containers as TF32 by IGNORING the low 13 mantissa bits, so the
//               raw tile IS the hi operand (hi = trunc(x)); only lo = x - trunc(x) (exact in fp32, nudged by half a
//               TF32 ulp so that the hardware truncation rounds it to nearest) is written, to a twin tile with the
//               identical swizzled layout.
//   warp 1      MMA issuer: lo*hi + hi*lo + hi*hi, tcgen05.mma kind::tf32, M = 128, N = bn, two ping-pong TMEM
//               accumulators.
//   warps 2-5   epilogue (one matrix row per thread): tcgen05.ld -> store the valid corner -> threshold top-k on the
//               accumulator itself:
//                 first tile of an image: two order-key maxima per row (left / right half of the tile) -> t0 = K-th
//                 largest of the 256 local maxima.  They are K distinct matrix elements >= t0, so { x >= t0 } contains
//                 the whole top-K; a second read of the tile from TMEM compacts those candidates (key << 32 | ~flat
//                 index) into shared memory.  Later tiles of the image push candidates in their single pass.
//                 last tile: rank the candidates by counting (the composite order IS the output order), write int64s.
//               Images whose candidate set overflows (adversarial: constant matrices) or that cannot be thresholded
//               set redo[b] = 1; the caller runs the stand-alone exact kernel (`launch_topk_pairs`) on those.
//
// HBM traffic per image is the algorithmic 2*N*256*4 B in + N*N*4 B + 2(3)*K*8 B out; operand re-reads of images with
// more than one tile (N > 128) come from L2.
#include "umma_ptx.cuh"

namespace pn {
namespace pairtopk {

using namespace umma;

constexpr int BM = 128, BK = 32;
constexpr int NUM_SPLIT_WARPS = 8;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS + 32 * NUM_SPLIT_WARPS;  // warp0 TMA, warp1 MMA, 2-9 epilogue, 10-17 splitters
constexpr int LO_SLOTS = 2;
constexpr int TOPK_MAX = 256, CAND_MAX = 1408;
constexpr int SMEM_LIMIT = 232448;  // 227 KB opt-in maximum per CTA
constexpr int TAIL_PAD = 2048;      // MMA row over-reads past the last tile (< 16 rows >= N: results never stored) stay inside

struct alignas(16) TopkSmem {  // 16-byte aligned: the ranking loop reads the candidates as ulonglong2, and there may be two of these
  unsigned long long cand[CAND_MAX];
  unsigned long long win[TOPK_MAX];
  uint32_t lm[2 * BM];
  uint32_t sl[8][32];  // the local maxima as 8 sorted lists
  uint32_t t0;
  unsigned ncand;
  uint32_t tguess;  // order key of the ~1.5 K-th largest element of the image just ranked: speculative threshold of the next one
};
constexpr int CTRL_BYTES = 256;  // mbarriers + TMEM base slot
constexpr int STAGE_TILE = 32 * 128;  // one 32-row x 32-column fp32 output tile (SWIZZLE_128B)

struct Params {
  CUtensorMap s_map, o_map;  // [B][N][K]; boxes 32 x s_box x 1 and 32 x o_box x 1
  CUtensorMap c_map;         // [B][N][N]; box 32 x 32 x 1 (store)
  float* C;                  // [B, N, N]
  int64_t *topk_idx, *sub_pos, *obj_pos;  // [B, topk] (topk_idx may be null)
  int* redo;                 // [B]
  int B, N, K, topk;
  int mtiles, ntiles;
  int n_step, bn;            // n-tile origin = nt * n_step (multiple of 8); bn = MMA N extent (n_step rounded up to 16, <= 256)
  int s_tile, o_tile;        // bytes of one raw k-block tile of S / O (multiples of 1024)
  int raw_stages;            // depth of the raw ring
  int acc_stride;            // TMEM columns between the two accumulators
  int tmem_cols;
  int bk;                    // fp32 kernel: channels per k-block, 32 (128-byte rows, SWIZZLE_128B) or 16 (64-byte rows, SWIZZLE_64B)
  int speculate;             // single-tile images: previous image's ~1.5 K-th value as the first-pass threshold (PN_OPT_PPN_SPECULATE)
  int epi_groups;            // 1, or 2 (bf16, single-tile images): two epilogue groups of 8 warps, one per TMEM accumulator
};

__device__ __forceinline__ uint32_t order_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
__device__ __forceinline__ unsigned long long composite(uint32_t key, uint32_t idx) {
  return ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - idx);
}
// the 8 warps of one epilogue group (named barrier 1 + group)
__device__ __forceinline__ void epi_sync(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory"); }

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// L2 cache policies: images with several tiles re-read their operand k-blocks once per tile.  At N = 400 the operands of
// the 148 images in flight (0.8 MB each) fill the 126 MB L2 exactly while 0.64 MB of matrix per image streams through it:
// without hints ncu shows 2.5 GB of DRAM traffic for 1.5 GB of algorithmic bytes.  Operand loads are marked evict_last,
// matrix stores evict_first.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                 int c2, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], "
      "[%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2,
                                                  uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// lo part of the 3xTF32 split when the hi part is the hardware's own truncation of x
__device__ __forceinline__ float lo_of(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  return __uint_as_float(__float_as_uint(x - hi) + 0x1000u);  // + half a TF32 ulp: truncation then rounds to nearest
}

// Candidates of one 32-column chunk held in v (one matrix row per thread): a 32-bit pass mask is built with two
// instructions per element; the (rare) passing elements are then fetched by dynamic index from the thread's own row of
// the swizzled staging tile `srow` (registers cannot be indexed dynamically) and pushed as composites.
__device__ __forceinline__ void push_candidates(const uint32_t (&v)[32], int nv, float t0f, const uint8_t* srow, int lane,
                                                uint32_t flat, TopkSmem& tk) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) m |= (!(__uint_as_float(v[j]) < t0f)) ? (1u << j) : 0u;  // NaN passes; its key ranks it
  if (nv < 32) m &= (1u << nv) - 1u;
  while (m) {
    const int j = __ffs(m) - 1;
    m &= m - 1;
    const float x = *reinterpret_cast<const float*>(srow + ((((j >> 2) ^ (lane & 7)) << 4) | ((j & 3) << 2)));
    const unsigned slot = atomicAdd(&tk.ncand, 1u);
    if (slot < CAND_MAX) tk.cand[slot] = composite(order_key(x), flat + (uint32_t)j);
  }
}
// descending bitonic sort of one key per lane
__device__ __forceinline__ uint32_t warp_sort_desc(uint32_t x, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
      const bool keep_max = ((lane & j) == 0) == ((lane & k) == 0);
      x = keep_max ? max(x, y) : min(x, y);
    }
  }
  return x;
}

// K-major operand tile with 64-byte rows (16 fp32 channels), SWIZZLE_64B: 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;  // SWIZZLE_64B
  return d;
}

struct Ring {  // (slot, phase) walker over a ring of runtime depth
  int slot, depth;
  uint32_t phase;
  __device__ __forceinline__ Ring(int d) : slot(0), depth(d), phase(0) {}
  __device__ __forceinline__ void next() {
    if (++slot == depth) { slot = 0; phase ^= 1u; }
  }
};

// BF16 = true: S and O are bf16 (`__nv_bfloat16` [B, N, 256]); one tcgen05.mma kind::f16 pass per 16 channels replaces the
// three TF32 passes, a k-block is 64 channels (the same 128-byte swizzle row), and there is nothing to split: the MMA
// warp takes the raw ring directly and the splitter warps are not launched.  The matrix and the top-k stay fp32 / int64.
template <bool BF16>
__global__ void __launch_bounds__(NUM_THREADS + 32 * NUM_EPI_WARPS, 1) pair_topk_kernel(const __grid_constant__ Params prm) {
  // channels per k-block.  fp32: 32, or 16 when the operand tiles are large (N >= ~128): with 32 only two raw stages fit
  // beside the lo ring and the staging tiles, and two stages do not cover TMA latency + split + MMA (measured 1.35 us per
  // k-block against 0.9 us of MMAs at N = 200); half-size k-blocks double the ring depth
  const int BKE = BF16 ? 64 : prm.bk;
  const bool sw64 = !BF16 && prm.bk == 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  const int stage_bytes = prm.s_tile + prm.o_tile;
  const int RS = prm.raw_stages;
  uint8_t* raw_ring = smem;
  uint8_t* lo_ring = smem + (size_t)RS * stage_bytes;
  uint8_t* out_stage = lo_ring + (size_t)(BF16 ? 0 : LO_SLOTS) * stage_bytes + TAIL_PAD;  // [8 warps][32 rows x 128 B], swizzled
  uint8_t* ctrl = out_stage + prm.epi_groups * NUM_EPI_WARPS * STAGE_TILE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);  // [RS<=8] TMA -> splitters
  uint64_t* empty_bar = full_bar + 8;                       // [RS<=8] MMA -> TMA
  uint64_t* split_bar = empty_bar + 8;                      // [2] splitters -> MMA
  uint64_t* lo_empty_bar = split_bar + 2;                   // [2] MMA -> splitters
  // MMA -> epilogue, [accumulator + 2 * epilogue group]: with two groups a group would otherwise skip the phases of the
  // other group's tiles, and a parity wait cannot tell phase n from phase n + 2
  uint64_t* tmem_full_bar = lo_empty_bar + 2;               // [4]
  uint64_t* tmem_empty_bar = tmem_full_bar + 4;             // [2] epilogue -> MMA
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  TopkSmem* tk_base = reinterpret_cast<TopkSmem*>(ctrl + CTRL_BYTES);  // one per epilogue group

  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < RS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&split_bar[a], NUM_SPLIT_WARPS);
      mbar_init(&lo_empty_bar[a], 1);
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_full_bar[a + 2], 1);
      mbar_init(&tmem_empty_bar[a], NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "r"(prm.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_base_slot);
  const int num_kb = prm.K / BKE;
  const int tiles_per_img = prm.mtiles * prm.ntiles;
  const int N = prm.N;

  if (warp == 0) {
    // ===== TMA producer (warp-uniform loop, one elected lane issues)
    Ring r(RS);
    const uint32_t stage_tx = (uint32_t)stage_bytes;
    const bool reuse = tiles_per_img > 1;
    const uint64_t pol_keep = l2_policy_evict_last();
    for (int b = blockIdx.x; b < prm.B; b += gridDim.x) {
      for (int t = 0; t < tiles_per_img; ++t) {
        const int m0 = (t / prm.ntiles) * BM, n0 = (t % prm.ntiles) * prm.n_step;
        for (int kb = 0; kb < num_kb; ++kb, r.next()) {
          mbar_wait(&empty_bar[r.slot], r.phase ^ 1u);
          uint8_t* st = raw_ring + (size_t)r.slot * stage_bytes;
          if (elect_one()) {
            mbar_expect_tx(&full_bar[r.slot], stage_tx);
            if (reuse) {
              tma_load_3d_hint(st, &prm.s_map, &full_bar[r.slot], kb * BKE, m0, b, pol_keep);
              tma_load_3d_hint(st + prm.s_tile, &prm.o_map, &full_bar[r.slot], kb * BKE, n0, b, pol_keep);
            } else {
              tma_load_3d(st, &prm.s_map, &full_bar[r.slot], kb * BKE, m0, b);
              tma_load_3d(st + prm.s_tile, &prm.o_map, &full_bar[r.slot], kb * BKE, n0, b);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues)
    const uint32_t idesc = BF16 ? make_idesc_bf16(prm.bn) : make_idesc(prm.bn);
    Ring r(RS), l(LO_SLOTS);
    uint32_t tile_it = 0, img_it = 0;
    for (int b = blockIdx.x; b < prm.B; b += gridDim.x, ++img_it) {
      const uint32_t full_grp = prm.epi_groups == 2 ? 2u * (img_it & 1u) : 0u;  // the epilogue group that owns this image
      for (int t = 0; t < tiles_per_img; ++t, ++tile_it) {
        const uint32_t acc = tile_it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((tile_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)prm.acc_stride;
        for (int kb = 0; kb < num_kb; ++kb, r.next(), l.next()) {
          if (BF16) {
            mbar_wait(&full_bar[r.slot], r.phase);  // bf16 tiles are MMA operands as they land
            tc_fence_after();
            const uint32_t op = smem_u32(raw_ring + (size_t)r.slot * stage_bytes);
            const uint64_t a = make_smem_desc(op), bd = make_smem_desc(op + prm.s_tile);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BKE / UMMA_K_BF16; ++k) {
                const uint64_t koff = (uint64_t)((k * UMMA_K_BF16 * 2) >> 4);
                umma_bf16(d_tmem, a + koff, bd + koff, idesc, (kb == 0 && k == 0) ? 0u : 1u);
              }
              umma_commit(&empty_bar[r.slot]);
            }
            __syncwarp();
            continue;
          }
          mbar_wait(&split_bar[l.slot], l.phase);  // lo tiles written (and, transitively, the raw tiles landed)
          tc_fence_after();
          const uint32_t hi = smem_u32(raw_ring + (size_t)r.slot * stage_bytes);
          const uint32_t lo = smem_u32(lo_ring + (size_t)l.slot * stage_bytes);
          const uint64_t a_hi = sw64 ? make_smem_desc_sw64(hi) : make_smem_desc(hi);
          const uint64_t b_hi = sw64 ? make_smem_desc_sw64(hi + prm.s_tile) : make_smem_desc(hi + prm.s_tile);
          const uint64_t a_lo = sw64 ? make_smem_desc_sw64(lo) : make_smem_desc(lo);
          const uint64_t b_lo = sw64 ? make_smem_desc_sw64(lo + prm.s_tile) : make_smem_desc(lo + prm.s_tile);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (k * UMMA_K >= BKE) break;
              const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
              const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
              umma_tf32(d_tmem, a_lo + koff, b_hi + koff, idesc, first);
              umma_tf32(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
              umma_tf32(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
            }
            umma_commit(&empty_bar[r.slot]);
            umma_commit(&lo_empty_bar[l.slot]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&tmem_full_bar[acc + full_grp]);
        __syncwarp();
      }
    }
  } else if (warp < 2 + NUM_EPI_WARPS * prm.epi_groups) {
    // ===== epilogue + top-k: 8 warps per group.  With two groups (bf16 entry point, images of one or two tiles: N <= 224)
    // group g owns every other image of the CTA: the per-image latency chain of the top-k (~6 us, what bounds the bf16
    // kernel there) runs for two images at once; the bf16 kernel has no splitter warps, so the CTA still has 18 warps.  TMEM lane quadrant = warp % 4 (one matrix row per thread); the two warps of a
    // quadrant take alternate 32-column chunks (par).  Everything below is latency bound at one warp per scheduler, so
    // the work per element is kept to one or two instructions: float compares against the threshold, keys only for
    // the (rare) candidates.
    const int quad = warp & 3;
    const int grp = (warp - 2) / NUM_EPI_WARPS;
    const int ew = (warp - 2) % NUM_EPI_WARPS;   // warp index inside the group
    const int par = ew >> 2;
    const int et = par * BM + quad * 32 + lane;  // 0 .. 255
    TopkSmem& tk = tk_base[grp];
    const int G = prm.epi_groups;
    const int K = prm.topk;
    uint8_t* sbuf = out_stage + (size_t)(warp - 2) * STAGE_TILE;  // one 32 x 32 fp32 staging tile per epilogue warp
    const float NEG_INF = __uint_as_float(0xff800000u);
    const uint32_t KEY_NEG_INF = order_key(NEG_INF);
    const uint64_t pol_stream = l2_policy_evict_first();
    uint32_t img_it = (uint32_t)grp;
    // Speculative threshold (single-tile images): the ~1.5 K-th largest value of the PREVIOUS image of this group is used
    // as the candidate threshold of the next one, in its first and only pass over the accumulator.  If that yields between
    // K and CAND_MAX candidates they contain the exact top-K (everything left out is smaller than every candidate) and the
    // local-maxima threshold + second TMEM read (~half of the per-image chain) are skipped; otherwise the image is flagged
    // and redone by the exact stand-alone kernel, and the next image takes the exact path again.
    const bool speculate = tiles_per_img == 1 && prm.speculate != 0;
    bool guess_ok = false;
    uint32_t guess = 0;
    if (et == 0) tk.ncand = 0;  // later resets happen at the end of each image, before its last epi_sync
    epi_sync(grp);
    for (int b = blockIdx.x + grp * (int)gridDim.x; b < prm.B; b += G * (int)gridDim.x, img_it += (uint32_t)G) {
      bool have_t0 = speculate && guess_ok;
      uint32_t t0 = have_t0 ? guess : 0u;
      float t0f = have_t0 ? key_to_float(guess) : 0.f;
      guess_ok = false;
      uint32_t tile_it = img_it * (uint32_t)tiles_per_img;  // same numbering as the MMA warp (G = 2: one tile per image)
      for (int t = 0; t < tiles_per_img; ++t, ++tile_it) {
        const int m0 = (t / prm.ntiles) * BM, n0 = (t % prm.ntiles) * prm.n_step;
        const uint32_t acc = tile_it & 1;
        // two groups (1 or 2 tiles per image): this group's barrier for `acc` completes once per image of the group
        if (G == 2) mbar_wait(&tmem_full_bar[acc + 2 * grp], (img_it >> 1) & 1);
        else mbar_wait(&tmem_full_bar[acc], (tile_it >> 1) & 1);
        tc_fence_after();
        const int row = m0 + quad * 32 + lane;
        const bool rvalid = row < N;
        const bool wvalid = m0 + quad * 32 < N;  // this warp's 32-row slab holds at least one matrix row
        const int ncols = min(prm.n_step, N - n0);  // valid columns of this n-tile
        const uint32_t flat0 = (uint32_t)row * (uint32_t)N + (uint32_t)n0;
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (uint32_t)prm.acc_stride;
        float lm = NEG_INF;
#pragma unroll 1
        for (int c0 = 32 * par; c0 < ncols; c0 += 64) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tbase + (uint32_t)c0, v);
          // ---- store: 32 rows x 32 columns through this warp's swizzled staging tile and ONE TMA store (rows / columns
          // past N are clipped by the tensor map); a thread-per-row STG would touch 32 lines per instruction
          if (wvalid) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile's previous store
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                              __uint_as_float(v[4 * j + 3]));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              if (tiles_per_img > 1) tma_store_3d_hint(&prm.c_map, sbuf, n0 + c0, m0 + quad * 32, b, pol_stream);
              else tma_store_3d(&prm.c_map, sbuf, n0 + c0, m0 + quad * 32, b);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
          if (rvalid) {
            const int nv = ncols - c0;  // valid columns of this chunk (>= 32: all)
            if (have_t0) {
              push_candidates(v, nv, t0f, sbuf + lane * 128, lane, flat0 + (uint32_t)c0, tk);
            } else if (nv >= 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) lm = fmaxf(lm, __uint_as_float(v[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) lm = fmaxf(lm, j < nv ? __uint_as_float(v[j]) : NEG_INF);
            }
          }
        }
        if (!have_t0) {
          // ---- threshold: t0 = K-th largest of the 256 local maxima (row x chunk parity) of the image's first tile; they
          // are distinct matrix elements, so { x >= t0 } contains the whole top-K.  The low 8 key bits are replaced by the
          // slot number (all entries distinct; t0 only moves down, by < 256 ulps).  The entries are dealt round-robin
          // to the 8 warps, each sorts its 32 with shuffles, and every thread ranks its entry by binary search in the
          // other seven sorted lists (an all-pairs count cost 2.5 us per image here, latency bound).
          tk.lm[et] = (order_key(lm) & 0xffffff00u) | (uint32_t)(255 - et);
          epi_sync(grp);
          const uint32_t mine = warp_sort_desc(tk.lm[lane * NUM_EPI_WARPS + ew], lane);
          tk.sl[ew][lane] = mine;
          epi_sync(grp);
          {
            int rank = lane;  // entries of the own list above this one
#pragma unroll
            for (int w = 1; w < NUM_EPI_WARPS; ++w) {
              const uint32_t* l = tk.sl[(ew + w) & (NUM_EPI_WARPS - 1)];
              int c = 0;  // entries of list l greater than mine
#pragma unroll
              for (int st = 16; st > 0; st >>= 1) c += (l[c + st - 1] > mine) ? st : 0;
              c += l[c] > mine;
              rank += c;
            }
            if (rank == K - 1) tk.t0 = mine & 0xffffff00u;
          }
          epi_sync(grp);
          t0 = tk.t0;
          t0f = key_to_float(t0);
          have_t0 = true;
          // ---- candidates of the first tile: second read of the accumulator
          if (t0 > (KEY_NEG_INF | 0xffu)) {  // else: fewer than K valid local maxima -> the image goes to the exact kernel
#pragma unroll 1
            for (int c0 = 32 * par; c0 < ncols; c0 += 64) {
              uint32_t v[32];
              tmem_ld_32x32b_x32(tbase + (uint32_t)c0, v);
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging tile is free again
              __syncwarp();
              if (rvalid) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  *reinterpret_cast<float4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                      make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                  __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                push_candidates(v, ncols - c0, t0f, sbuf + lane * 128, lane, flat0 + (uint32_t)c0, tk);
              }
            }
          }
        }
        // accumulator drained: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        const bool last = t == tiles_per_img - 1;
        if (last || (t & 3) == 0) {
          // ---- rank the candidates by counting: win[r] = candidate with r larger candidates (the composite order IS the
          // output order).  Last tile: emit.  Every 4th tile of a multi-tile image: keep the K best so far and raise
          // t0 to the K-th of them, so that the candidate list stays short however many tiles follow.
          epi_sync(grp);
          const unsigned nc = tk.ncand;
          const bool bad = t0 <= (KEY_NEG_INF | 0xffu) || nc > (unsigned)CAND_MAX || nc < (unsigned)K;
          if (last && et == 0) prm.redo[b] = bad ? 1 : 0;
          if (!bad) {
            const unsigned tgt = min(nc, (unsigned)(K + K / 2)) - 1u;  // rank whose key becomes the next image's guess
            for (unsigned i = et; i < nc; i += 32 * NUM_EPI_WARPS) {
              const unsigned long long me = tk.cand[i];
              unsigned r0 = 0, r1 = 0;
              unsigned j = 0;
#pragma unroll 4
              for (; j + 1 < nc; j += 2) {
                const ulonglong2 c = *reinterpret_cast<const ulonglong2*>(&tk.cand[j]);
                r0 += c.x > me; r1 += c.y > me;
              }
              if (j < nc) r0 += tk.cand[j] > me;
              if (r0 + r1 < (unsigned)K) tk.win[r0 + r1] = me;
              if (r0 + r1 == tgt) tk.tguess = (uint32_t)(me >> 32);
            }
            epi_sync(grp);
            if (last && speculate) { guess = tk.tguess; guess_ok = true; }
            if (last) {
              for (int r = et; r < K; r += 32 * NUM_EPI_WARPS) {
                const uint32_t idx = 0xffffffffu - (uint32_t)(tk.win[r] & 0xffffffffull);
                const size_t o = (size_t)b * K + r;
                if (prm.topk_idx) prm.topk_idx[o] = (long long)idx;
                prm.sub_pos[o] = (long long)(idx / (uint32_t)N);
                prm.obj_pos[o] = (long long)(idx % (uint32_t)N);
              }
            } else {
              for (int r = et; r < K; r += 32 * NUM_EPI_WARPS) tk.cand[r] = tk.win[r];
              if (et == 0) tk.ncand = (unsigned)K;
              t0 = (uint32_t)(tk.win[K - 1] >> 32);
              t0f = key_to_float(t0);
            }
          }
          if (last && et == 0) tk.ncand = 0;  // the next image may push in its first pass (speculative threshold)
          epi_sync(grp);  // last: smem of this image is recycled by the next one; else: the shortened list is published
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging tiles outlive their stores
    tc_fence_before();
  } else if (!BF16) {
    // ===== splitters: lo tiles of S and O (16-byte chunks, swizzled layout preserved); the raw tile is the hi operand
    const int sid = threadIdx.x - (64 + 32 * NUM_EPI_WARPS * prm.epi_groups);  // 0 .. 255
    constexpr int NSPLIT = 32 * NUM_SPLIT_WARPS;
    Ring r(RS), l(LO_SLOTS);
    for (int b = blockIdx.x; b < prm.B; b += gridDim.x) {
      for (int t = 0; t < tiles_per_img; ++t) {
        // only rows that exist are split (8 chunks of 16 B per row; the swizzle permutes chunks inside a row only).
        // Rows past N are zero in the raw tile (TMA fill) or lie past it; whatever the lo tile holds there only
        // reaches output rows / columns >= N, which are never stored nor ranked.
        const int m0 = (t / prm.ntiles) * BM, n0 = (t % prm.ntiles) * prm.n_step;
        const int cpr = BKE / 4;  // 16-byte chunks per row of a k-block tile
        const int chunks_s = ((min(BM, N - m0) + 7) & ~7) * cpr;
        const int chunks_o = ((min(prm.n_step, N - n0) + 7) & ~7) * cpr;
        const int o_shift = prm.s_tile - chunks_s * 16;  // chunk index -> byte offset jump from the S to the O tile
        for (int kb = 0; kb < num_kb; ++kb, r.next(), l.next()) {
          mbar_wait(&full_bar[r.slot], r.phase);
          mbar_wait(&lo_empty_bar[l.slot], l.phase ^ 1u);
          const uint8_t* src = raw_ring + (size_t)r.slot * stage_bytes;
          uint8_t* dst = lo_ring + (size_t)l.slot * stage_bytes;
#pragma unroll 4
          for (int c = sid; c < chunks_s + chunks_o; c += NSPLIT) {
            const int off = c * 16 + (c >= chunks_s ? o_shift : 0);
            const float4 x = *reinterpret_cast<const float4*>(src + off);
            float4 y;
            y.x = lo_of(x.x); y.y = lo_of(x.y); y.z = lo_of(x.z); y.w = lo_of(x.w);
            *reinterpret_cast<float4*>(dst + off) = y;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to tcgen05.mma
          __syncwarp();
          if (lane == 0) mbar_arrive(&split_bar[l.slot]);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(prm.tmem_cols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map_3d(CUtensorMap* map, const void* ptr, int B, int N, int K, int box_rows, int box_cols = BK,
                       bool bf16 = false, bool swizzle64 = false) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  PN_REQUIRE(fn, PN_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  PN_REQUIRE(((uintptr_t)ptr & 15) == 0, PN_ERR_UNSUPPORTED, "pair top-k: embeddings must be 16B aligned");
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)B};
  const cuuint64_t es = bf16 ? 2 : 4;
  cuuint64_t strides[2] = {(cuuint64_t)K * es, (cuuint64_t)N * K * es};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                  const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PN_REQUIRE(r == CUDA_SUCCESS, PN_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
  return 0;
}

}  // namespace pairtopk

// the fused kernel thresholds on 2 local maxima (even / odd 32-column chunks) per row of the image's first tile
bool pair_topk_fused_supported(int N, int K, int topk, bool bf16) {
  const int rows0 = N < pairtopk::BM ? N : pairtopk::BM;
  const int bke = bf16 ? 64 : pairtopk::BK;
  return topk >= 1 && topk <= pairtopk::TOPK_MAX && topk <= (N > 32 ? 2 : 1) * rows0 && N >= 2 && (N & 3) == 0 &&
         K % bke == 0 && K >= bke && (long long)N * N < (1ll << 31);
}

// importance[b] = S[b] . O[b]^T (3xTF32 on tcgen05) with the top-k pair select fused into the epilogue.
// redo [B] int: set to 1 for images the caller must pass to `launch_topk_pairs` (exact kernel), 0 otherwise.
int launch_pair_topk_fused(const void* S, const void* O, bool bf16, float* C, int64_t* topk_idx, int64_t* sub_pos,
                           int64_t* obj_pos, int* redo, int B, int N, int K, int topk, cudaStream_t st) {
  using namespace pairtopk;
  PN_REQUIRE(S && O && C && sub_pos && obj_pos && redo && B > 0, PN_ERR_BAD_ARG, "pair top-k: bad args");
  PN_REQUIRE(pair_topk_fused_supported(N, K, topk, bf16), PN_ERR_UNSUPPORTED,
             "pair top-k: N=%d K=%d topk=%d unsupported", N, K, topk);
  Params prm{};
  prm.C = C; prm.topk_idx = topk_idx; prm.sub_pos = sub_pos; prm.obj_pos = obj_pos; prm.redo = redo;
  prm.B = B; prm.N = N; prm.K = K; prm.topk = topk;
  prm.mtiles = cdiv(N, BM);
  // n-tiles of one image start at multiples of 32 columns: the epilogue stores 32-column boxes
  prm.ntiles = (int)round_up(N, 8) <= 224 ? 1 : cdiv(N, 224);  // <= 224 columns: two raw stages + lo + staging fit smem
  prm.n_step = prm.ntiles == 1 ? (int)round_up(N, 8) : (int)round_up(cdiv(N, prm.ntiles), 32);
  prm.bn = (int)round_up(prm.n_step, 16);
  const int s_box = (int)round_up(N < BM ? N : BM, 8);
  const int o_box = prm.n_step;
  // fp32: half-size k-blocks (16 channels, SWIZZLE_64B) when fewer than 4 raw stages of 32 channels would fit
  prm.bk = BK;
  if (!bf16 && get_option(OPT_PPN_HALF_KB)) {
    const int fixed1 = 1024 + TAIL_PAD + NUM_EPI_WARPS * STAGE_TILE + (int)sizeof(TopkSmem) + CTRL_BYTES + 64;
    if ((SMEM_LIMIT - fixed1) / ((s_box + o_box) * BK * 4) - LO_SLOTS < 4) prm.bk = 16;
  }
  prm.s_tile = s_box * prm.bk * 4;
  prm.o_tile = o_box * prm.bk * 4;
  PN_TRY(make_map_3d(&prm.s_map, S, B, N, K, s_box, bf16 ? 64 : prm.bk, bf16, prm.bk == 16));
  PN_TRY(make_map_3d(&prm.o_map, O, B, N, K, o_box, bf16 ? 64 : prm.bk, bf16, prm.bk == 16));
  PN_TRY(make_map_3d(&prm.c_map, C, B, N, N, 32, 32));
  prm.speculate = get_option(OPT_PPN_SPECULATE) != 0;
  const int stage = prm.s_tile + prm.o_tile;
  // PN_OPT_PPN_EPI2: 1 = two groups for the bf16 entry point; 2 = for the fp32 kernel too (A/B studies: 26 warps)
  prm.epi_groups = (prm.mtiles * prm.ntiles <= 2 && (bf16 ? get_option(OPT_PPN_EPI2) >= 1 : get_option(OPT_PPN_EPI2) >= 2)) ? 2 : 1;
  const int lo_slots = bf16 ? 0 : LO_SLOTS;
  auto fixed_bytes = [&](int groups) {
    return 1024 + TAIL_PAD + groups * (NUM_EPI_WARPS * STAGE_TILE + (int)sizeof(TopkSmem)) + CTRL_BYTES + 64;
  };
  // the second group's staging tiles and top-k state cost 47 KiB of the operand ring: keep at least 3 raw stages
  if (prm.epi_groups == 2 && (SMEM_LIMIT - fixed_bytes(2)) / stage - lo_slots < (bf16 ? 3 : 2)) prm.epi_groups = 1;
  const int fixed = fixed_bytes(prm.epi_groups);
  int rs = (SMEM_LIMIT - fixed) / stage - lo_slots;
  rs = rs > 8 ? 8 : rs;
  PN_REQUIRE(rs >= 2, PN_ERR_UNSUPPORTED, "pair top-k: tiles of N=%d do not fit shared memory", N);
  prm.raw_stages = rs;
  prm.acc_stride = prm.bn <= 128 ? 128 : 256;
  prm.tmem_cols = 2 * prm.acc_stride;
  const size_t smem = (size_t)fixed + (size_t)(rs + lo_slots) * stage;
  static bool attr_done[PN_MAX_DEVICES] = {false};  // the attribute is per device
  bool& attr_set = attr_done[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pair_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(pair_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    PN_REQUIRE(e == cudaSuccess, (int)e, "pair top-k: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int num_sms = sm_count();
  const int grid = B < num_sms ? B : num_sms;
  if (bf16)  // no splitter warps; one or two epilogue groups
    pair_topk_kernel<true><<<grid, 64 + 32 * NUM_EPI_WARPS * prm.epi_groups, smem, st>>>(prm);
  else pair_topk_kernel<false><<<grid, NUM_THREADS + 32 * NUM_EPI_WARPS * (prm.epi_groups - 1), smem, st>>>(prm);
  return check_launch("pair_topk_kernel");
}

}  // namespace pn
