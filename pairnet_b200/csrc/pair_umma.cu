// Pair Proposal Network pair matrix on tcgen05:  importance_raw[b] = S[b] . O[b]^T   (pairnet_head.py:327)
//
//   S, O : [B, N, K] row-major fp32 (L2-normalised subject / object embeddings, K = 256),  C : [B, N, N] fp32.
//
// One persistent CTA per SM walks the (image, m-tile, n-tile) list.  Both operands arrive RAW: a 3-D TMA map
// [B][N][K] (box 1 x 128 x 32, SWIZZLE_128B) stages 128-row k-blocks in shared memory -- rows past N are
// out-of-bounds for the map and are zero-filled by the TMA unit, so no padded copy of the embeddings ever exists
// and no byte of the next image is fetched.  Eight splitter warps turn each raw tile into its 3xTF32 pair IN
// PLACE (hi = rna_tf32(x) overwrites the tile, lo = rna_tf32(x - hi) goes to a twin tile with the identical
// swizzled layout), fence the generic->async proxy and hand the stage to the single MMA thread, which issues
// lo*hi + hi*lo + hi*hi tcgen05.mma kind::tf32 (M = 128, N = BN <= 256) into one of two ping-pong TMEM
// accumulators; four epilogue warps drain the other accumulator with tcgen05.ld and store the valid N x N corner.
//
// HBM traffic per image is exactly the algorithmic 2*N*K*4 B in + N*N*4 B out (operand re-reads for N > 128
// come from L2).
#include "umma_ptx.cuh"

namespace pn {
namespace pairmm {

using namespace umma;

constexpr int BM = 128, BK = 32;
constexpr int TILE16K = BM * BK * 4;  // one 128-row x 128-byte operand tile
constexpr int NUM_SPLIT_WARPS = 8;
constexpr int NUM_EPI_WARPS = 4;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS + 32 * NUM_SPLIT_WARPS;  // warp0 TMA, warp1 MMA, 2-5 epi, 6-13 split

// BNMAX = 128: N <= 128 (one 128-row O box per k-block), 4 stages.  BNMAX = 256: two O boxes, 2 stages.
template <int BNMAX>
struct Cfg {
  static constexpr int OBOXES = BNMAX / 128;
  static constexpr int STAGE_BYTES = 2 * TILE16K + 2 * OBOXES * TILE16K;  // S hi(raw), S lo, O hi(raw), O lo
  static constexpr int STAGES = BNMAX == 128 ? 3 : 2;
  static constexpr int OFF_S_HI = 0, OFF_S_LO = TILE16K, OFF_O_HI = 2 * TILE16K, OFF_O_LO = (2 + OBOXES) * TILE16K;
  static constexpr int TMEM_COLS = 2 * BNMAX;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
};

struct Params {
  CUtensorMap s_map, o_map;  // [B][N][K], box 1 x 128 x 32
  float* C;                  // [B, N, N]
  int B, N, K;
  int mtiles, ntiles, bn;    // bn = MMA N extent of one n-tile (multiple of 16, <= BNMAX); n-tile origin = nt * bn
  int total_tiles;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
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

struct TileCoord { int b, m0, n0; };
__device__ __forceinline__ TileCoord tile_coord(const Params& p, int t) {
  const int per_img = p.mtiles * p.ntiles;
  const int b = t / per_img, r = t - b * per_img;
  return TileCoord{b, (r / p.ntiles) * BM, (r % p.ntiles) * p.bn};
}

template <int BNMAX>
__global__ void __launch_bounds__(NUM_THREADS, 1) pair_umma_kernel(const __grid_constant__ Params prm) {
  using C = Cfg<BNMAX>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);  // TMA -> splitters
  uint64_t* split_bar = full_bar + STAGES;                                           // splitters -> MMA
  uint64_t* empty_bar = split_bar + STAGES;                                          // MMA -> TMA
  uint64_t* tmem_full_bar = empty_bar + STAGES;                                      // [2] MMA -> epilogue
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                                      // [2] epilogue -> MMA
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], NUM_SPLIT_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "r"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_base_slot);
  const int num_kb = prm.K / BK;
  // rows of O actually needed by one n-tile, in 128-row boxes
  const int oboxes = (prm.bn + 127) / 128;
  const uint32_t stage_tx = (uint32_t)(TILE16K + oboxes * TILE16K);

  if (warp == 0) {
    // ===== TMA producer (warp-uniform loop, one elected lane issues)
    uint32_t it = 0;
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
      const TileCoord tc = tile_coord(prm, t);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        uint8_t* st = smem + (size_t)s * C::STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[s], stage_tx);
          tma_load_3d(st + C::OFF_S_HI, &prm.s_map, &full_bar[s], kb * BK, tc.m0, tc.b);
          for (int h = 0; h < oboxes; ++h)
            tma_load_3d(st + C::OFF_O_HI + h * TILE16K, &prm.o_map, &full_bar[s], kb * BK, tc.n0 + h * 128, tc.b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues)
    const uint32_t idesc = make_idesc(prm.bn);
    uint32_t it = 0, tile_it = 0;
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x, ++tile_it) {
      const uint32_t acc = tile_it & 1;
      mbar_wait(&tmem_empty_bar[acc], ((tile_it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BNMAX;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&split_bar[s], (it / STAGES) & 1);  // hi/lo tiles of this k-block are in smem
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)s * C::STAGE_BYTES);
        const uint64_t a_hi = make_smem_desc(st + C::OFF_S_HI), a_lo = make_smem_desc(st + C::OFF_S_LO);
        const uint64_t b_hi = make_smem_desc(st + C::OFF_O_HI), b_lo = make_smem_desc(st + C::OFF_O_LO);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
            const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
            umma_tf32(d_tmem, a_lo + koff, b_hi + koff, idesc, first);
            umma_tf32(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
            umma_tf32(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&tmem_full_bar[acc]);
      __syncwarp();
    }
  } else if (warp < 2 + NUM_EPI_WARPS) {
    // ===== epilogue: TMEM lane quadrant = warp % 4; one output row per thread
    const int quad = warp & 3;
    const int N = prm.N;
    const bool vec = (N & 3) == 0;
    uint32_t tile_it = 0;
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x, ++tile_it) {
      const TileCoord tc = tile_coord(prm, t);
      const uint32_t acc = tile_it & 1;
      mbar_wait(&tmem_full_bar[acc], (tile_it >> 1) & 1);
      tc_fence_after();
      const int row = tc.m0 + quad * 32 + lane;
      float* crow = prm.C + ((size_t)tc.b * N + row) * N;
      const int ncols = min(prm.bn, N - tc.n0);  // valid columns of this n-tile
#pragma unroll 1
      for (int c0 = 0; c0 < prm.bn; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BNMAX + (uint32_t)c0, v);
        if (c0 + 16 >= prm.bn) {  // last TMEM read of the tile: hand the accumulator back before storing
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (row < N && c0 < ncols) {
          float* dst = crow + tc.n0 + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (vec && c0 + j + 3 < ncols) {
              *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (c0 + j + u < ncols) dst[j + u] = __uint_as_float(v[j + u]);
            }
          }
        }
      }
    }
    tc_fence_before();
  } else {
    // ===== splitters: raw fp32 tile -> hi (in place) + lo (twin tile); 16-byte chunks, layout preserved
    const int sid = threadIdx.x - (64 + 32 * NUM_EPI_WARPS);  // 0 .. 255
    constexpr int NSPLIT = 32 * NUM_SPLIT_WARPS;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
      // only rows that exist are split (8 chunks of 16 B per row; the swizzle permutes chunks inside a row only).
      // Rows past N are zero in the hi tile (TMA fill); whatever the lo tile holds there only reaches output
      // rows / columns >= N, which are never stored.
      const TileCoord tc = tile_coord(prm, t);
      const int chunks_s = ((min(BM, prm.N - tc.m0) + 7) & ~7) * 8;
      const int chunks_o = ((min(prm.bn, prm.N - tc.n0) + 7) & ~7) * 8;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full_bar[s], (it / STAGES) & 1);
        uint8_t* st = smem + (size_t)s * C::STAGE_BYTES;
#pragma unroll 2
        for (int c = sid; c < chunks_s + chunks_o; c += NSPLIT) {
          const bool is_s = c < chunks_s;
          const int off = (is_s ? c : c - chunks_s) * 16;
          float4* ph = reinterpret_cast<float4*>(st + (is_s ? C::OFF_S_HI : C::OFF_O_HI) + off);
          float4* pl = reinterpret_cast<float4*>(st + (is_s ? C::OFF_S_LO : C::OFF_O_LO) + off);
          const float4 x = *ph;
          float4 h, l;
          h.x = rna_tf32(x.x); h.y = rna_tf32(x.y); h.z = rna_tf32(x.z); h.w = rna_tf32(x.w);
          l.x = rna_tf32(x.x - h.x); l.y = rna_tf32(x.y - h.y); l.z = rna_tf32(x.z - h.z); l.w = rna_tf32(x.w - h.w);
          *ph = h;
          *pl = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[s]);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map_3d(CUtensorMap* map, const float* ptr, int B, int N, int K) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  PN_REQUIRE(fn, PN_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  PN_REQUIRE(((uintptr_t)ptr & 15) == 0, PN_ERR_UNSUPPORTED, "pair matrix: embeddings must be 16B aligned");
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)K * 4, (cuuint64_t)N * K * 4};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PN_REQUIRE(r == CUDA_SUCCESS, PN_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
  return 0;
}

}  // namespace pairmm

// C[b] = S[b] . O[b]^T for b < B;  S, O [B,N,K] fp32 (K % 32 == 0), C [B,N,N] fp32.  3xTF32 on tcgen05.
int launch_pair_matrix_tc(const float* S, const float* O, float* C, int B, int N, int K, cudaStream_t st) {
  using namespace pairmm;
  PN_REQUIRE(S && O && C && B > 0 && N > 0, PN_ERR_BAD_ARG, "pair matrix: bad args");
  PN_REQUIRE(K % BK == 0 && K >= BK, PN_ERR_UNSUPPORTED, "pair matrix: K=%d must be a multiple of %d", K, BK);
  Params prm{};
  PN_TRY(make_map_3d(&prm.s_map, S, B, N, K));
  PN_TRY(make_map_3d(&prm.o_map, O, B, N, K));
  prm.C = C; prm.B = B; prm.N = N; prm.K = K;
  prm.mtiles = cdiv(N, BM);
  prm.ntiles = cdiv(N, 256);
  prm.bn = (int)round_up(cdiv(N, prm.ntiles), 16);
  const long long tiles = (long long)B * prm.mtiles * prm.ntiles;
  PN_REQUIRE(tiles < (1ll << 31), PN_ERR_UNSUPPORTED, "pair matrix: too many tiles");
  prm.total_tiles = (int)tiles;
  static bool attr_done[PN_MAX_DEVICES] = {false};  // the attribute is per device
  bool& attr_set = attr_done[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pair_umma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg<128>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(pair_umma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)Cfg<256>::SMEM_BYTES);
    PN_REQUIRE(e == cudaSuccess, (int)e, "pair matrix: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int num_sms = sm_count();
  const int grid = prm.total_tiles < num_sms ? prm.total_tiles : num_sms;
  if (prm.bn <= 128)
    pair_umma_kernel<128><<<grid, NUM_THREADS, Cfg<128>::SMEM_BYTES, st>>>(prm);
  else
    pair_umma_kernel<256><<<grid, NUM_THREADS, Cfg<256>::SMEM_BYTES, st>>>(prm);
  return check_launch("pair_umma_kernel");
}

}  // namespace pn
