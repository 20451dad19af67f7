// tcgen05 tensor-core GEMM for the memory-side projections:  C[M,N] = A[M,K] . W[N,K]^T + bias
//
// Blackwell-native structure (sm_100a): TMA (cp.async.bulk.tensor, SWIZZLE_128B) stages 128 x 32-float
// operand tiles in shared memory through a 3-deep mbarrier ring; one elected thread issues
// tcgen05.mma.kind::tf32 (M=128, N=128, K=8) with the fp32 accumulator in TMEM; four epilogue warps read the
// accumulator back with tcgen05.ld, add the bias and store 128-bit rows.
//
// fp32 parity ("3xTF32"): every operand is pre-split into hi = rna_tf32(x) and lo = rna_tf32(x - hi);
// the accumulator receives lo*hi + hi*lo + hi*hi, i.e. all product bits down to ~2^-22 relative, so the
// result matches an fp32 FFMA GEMM to ~1e-6 while running on the tensor pipe.  `passes = 1` (hi*hi only)
// is plain TF32.
#include "umma_ptx.cuh"

namespace pn {
namespace umma {

constexpr int BM = 128, BK = 32;            // BK fp32 = 128 B = one SWIZZLE_128B row
constexpr int A_TILE_BYTES = BM * BK * 4;   // 16 KiB
// epilogue warps: 4 (one per TMEM lane quadrant) or 8 (two per quadrant, each takes half the columns)
constexpr int BIAS_MAX = 1024;              // per-problem bias staged in smem (epilogue reads it with LDS, not LDG)
// TMA-store epilogue: one swizzled 32-row x 32-column fp32 staging tile per epilogue warp (4 warps).  A thread-per-row
// STG touches 32 lines per instruction (~6.5 cycles per 16-byte request): the ncu role profile showed the four
// epilogue warps 100 % busy and every other role waiting on them -- the kernel was store-issue bound, not MMA bound.
constexpr int EPI_TILE_BYTES = 32 * 128;
constexpr int EPI_STAGE_BYTES = 4 * EPI_TILE_BYTES;
// BN = 128: 3 stages x 64 KiB, 2 x 128 TMEM columns.  BN = 256 (N % 256 == 0): A tiles are re-read half as often;
// 2 stages x 96 KiB, 2 x 256 TMEM columns (the whole TMEM).
template <int BN>
struct Cfg {
  static constexpr int STAGES = BN == 128 ? 3 : 2;
  static constexpr int B_TILE_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;  // a_hi, a_lo, b_hi, b_lo
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr size_t SMEM_BYTES =
      (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + EPI_STAGE_BYTES + 256 /*barriers*/ + 4 * BIAS_MAX * sizeof(float);
};

// raw-A variant: 4 stages x (A raw 16 KiB + B hi/lo 32 KiB)
constexpr size_t RAW_SMEM_BYTES =
    (size_t)4 * (A_TILE_BYTES + 2 * 128 * BK * 4) + 1024 + EPI_STAGE_BYTES + 256 + 4 * BIAS_MAX * sizeof(float);

struct Problem {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;  // A [M,K], W [N,K]; box 32 x 128, SWIZZLE_128B
  CUtensorMap c_map, clo_map;          // C / C_lo [M,N] (row pitch ldc); box 32 x 32: TMA-store epilogue
  int tma_store;                       // row-major store through the staging tile + TMA (else per-thread STG)
  const float* bias;
  float* C;
  float* C_lo;  // non-null: write the result pre-split for a following 3xTF32 GEMM (C = hi, C_lo = lo)
  int M, N, K, ldc;
  int relu;
  int t_rows;   // > 0: transposed store  C[(m / t_rows) * N + n][m % t_rows]  (row pitch ldc), e.g. V^T per image
  int bias_per_row;  // bias indexed by the output row m instead of the column n (weights as the A operand)
  // sign/bit-pack epilogue (boolean attention mask, pairnet_head.py:244-256): rows = keys, columns = queries;
  // bits[q * bits_words + m / 32] bit (m % 32) = (acc[m][q] < 0) | (m >= M); rowany[q] |= 1 if any key is open
  uint32_t* bits;
  int* rowany;
  int bits_words;
};
constexpr int MAX_PROBLEMS = 4;
struct Params {
  Problem p[MAX_PROBLEMS];
  int passes;  // 3 = 3xTF32 (fp32 parity), 1 = plain TF32
  int count;
  int total_tiles;
};

// Tile scheduler: linear tile id -> (problem, m-tile, n-tile), n fastest so that CTAs working at the same time
// share the A rows in L2.
struct TileCoord { int p, m0, n0; };
template <int BN>
__device__ __forceinline__ TileCoord tile_coord(const Params& prm, int t) {
  TileCoord c{0, 0, 0};
#pragma unroll
  for (int i = 0; i < MAX_PROBLEMS; ++i) {
    const int nt = (prm.p[i].N + BN - 1) / BN;
    const int tiles = ((prm.p[i].M + BM - 1) / BM) * nt;
    if (i == prm.count - 1 || t < tiles) {
      c.p = i; c.m0 = (t / nt) * BM; c.n0 = (t % nt) * BN;
      return c;
    }
    t -= tiles;
  }
  return c;
}

// Persistent, warp-specialised: warp0 = TMA producer, warp1 = MMA issuer (+ TMEM owner), warps2-5 = epilogue,
// (A_RAW only) warps6-9 = A splitter.  Two 128-column TMEM accumulators ping-pong so the epilogue of tile i
// overlaps the MMAs of tile i+1.
//
// A_RAW = true: the A operand is loaded RAW (plain fp32, one 16 KiB tile per k-block instead of a hi and a lo
// tile) and split in the SM: four splitter warps read the swizzled tile (one row per thread), compute
// hi = rna_tf32(a), lo = rna_tf32(a - hi) and park both in tensor memory with tcgen05.st; the MMAs then take A
// from TMEM (tcgen05.mma with a TMEM A operand) and only B (weights, pre-split) from shared memory.  This halves
// the A bytes moved through HBM/L2 and frees smem for a 4th pipeline stage.
//
// W16 = true ("3xBF16", raw-A only): the same fp32-in / fp32-out GEMM on the kind::f16 pipe, which runs at TWICE the TF32
// rate.  Weights arrive as bf16 hi / lo planes (hi = bf16(w), lo = bf16(w - hi)); a k-block is 64 channels: two raw fp32
// A sub-tiles (2 x 16 KiB) + one 16 KiB bf16 tile per weight plane = 64 KiB per stage, 3 stages.  The splitter warps turn
// each raw row into packed bf16 hi / lo pairs (cvt.rn.bf16x2, two elements per 32-bit TMEM column: the same 32 + 32 columns
// per k-block as the TF32 variant) and the MMA warp issues lo*hi + hi*lo + hi*hi with K = 16 per instruction -- 12
// instructions per 64 channels instead of 24.  Dropped terms (lo*lo and the third bf16 digit) are 2^-16..2^-17 relative
// per product: ~1e-5 of the result's scale, used where the INPUTS already carry TF32-level error (the pixel-decoder
// encoder, fed by cuDNN TF32 convolutions); the head keeps 3xTF32.
template <int BN, int NUM_EPI_WARPS, bool A_RAW, bool W16 = false>
__global__ void __launch_bounds__(64 + 32 * NUM_EPI_WARPS + (A_RAW ? (W16 ? 288 : 128) : 0), 1)
umma_gemm_kernel(const __grid_constant__ Params prm) {
  static_assert(!W16 || A_RAW, "the bf16-split variant takes a raw fp32 A operand");
  constexpr int STAGES = W16 ? 3 : (A_RAW ? 4 : Cfg<BN>::STAGES);
  constexpr int B_TILE = Cfg<BN>::B_TILE_BYTES;
  constexpr int STAGE_BYTES = W16 ? (2 * A_TILE_BYTES + 2 * B_TILE) : (A_RAW ? (A_TILE_BYTES + 2 * B_TILE) : Cfg<BN>::STAGE_BYTES);
  constexpr int TMEM_COLS = A_RAW ? 512 : Cfg<BN>::TMEM_COLS;
  constexpr int OFF_A_HI = 0, OFF_A_LO = A_TILE_BYTES;
  constexpr int OFF_B_HI = W16 ? 2 * A_TILE_BYTES : (A_RAW ? A_TILE_BYTES : 2 * A_TILE_BYTES);
  constexpr int OFF_B_LO = OFF_B_HI + B_TILE;
  constexpr int KB = W16 ? 64 : BK;  // channels per k-block
  // TMA-store epilogue: one staging tile per epilogue warp.  W16 runs EIGHT epilogue warps (two per TMEM lane quadrant, 64
  // columns each -- the ncu role profile of the 4-warp version showed them 78 % busy with the splitters starved behind
  // them on the K = 256 problems); their four extra tiles take the place of the bias staging area, the bias is read
  // from global memory instead (one broadcast 16-byte request per 4 columns).
  constexpr bool TMA_EPI = (BN == 128 && (NUM_EPI_WARPS == 4 || W16));
  constexpr int EPI_BYTES = W16 ? NUM_EPI_WARPS * EPI_TILE_BYTES : EPI_STAGE_BYTES;
  static_assert(!W16 || NUM_EPI_WARPS == 8, "the bf16-split variant is built for 8 epilogue warps");
  constexpr int TM_A = 2 * BN;  // A_RAW: TMEM columns [TM_A + set*64, +32) = hi, [+32, +64) = lo  (4 sets -> 512 total)
  static_assert(!A_RAW || BN == 128, "raw-A variant is built for 128x128 tiles");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space (LDS/STS)
  uint8_t* epi_stage = smem + STAGES * STAGE_BYTES;  // [4 warps][32 rows x 128 B], 1024-aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  constexpr int ASETS = 4;                        // TMEM A staging sets (hi 32 + lo 32 columns each)
  uint64_t* a_ready_bar = tmem_empty_bar + 2;     // [ASETS]  splitter -> MMA   (A_RAW)
  uint64_t* a_free_bar = a_ready_bar + ASETS;     // [ASETS]  MMA -> splitter   (A_RAW)
  // W16: the weight tiles of a stage have their own barrier pair.  The raw A tiles are released by the SPLITTERS (they are
  // the only readers), the weight tiles by the MMA commit, and a second producer warp refills them independently: the A
  // round trip (TMA latency + split) no longer contains the MMA time, which is what kept the 3-deep ring from covering
  // the L2 latency (measured 0.9 us per 64-channel k-block against 0.56 us of MMAs).
  uint64_t* fullB_bar = a_free_bar + ASETS;       // [STAGES]
  uint64_t* emptyB_bar = fullB_bar + STAGES;      // [STAGES]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(emptyB_bar + STAGES);
  static_assert((2 * 4 + 4 + 2 * ASETS + 2 * 4) * 8 + 4 <= 256, "barrier block");
  float* bias_s = reinterpret_cast<float*>(epi_stage + EPI_BYTES + 256);  // [MAX_PROBLEMS][BIAS_MAX]  (not W16)

  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const bool is_epi = warp >= 2 && warp < 2 + NUM_EPI_WARPS;
  // (Not a PDL kernel: launching the persistent tcgen05 kernels with programmatic serialization, or letting them trigger
  // their dependents early, made the pixel-decoder encoder 0.2 ms SLOWER on B200 -- measured, scratch/pdl_ab.py; the
  // attribute is kept for the microsecond-scale kernels of the query-side chain only.)
  if (is_epi && !W16) {  // epilogue warps stage the bias vectors (zeros when absent / beyond N)
    for (int i = threadIdx.x - 64; i < MAX_PROBLEMS * BIAS_MAX; i += 32 * NUM_EPI_WARPS) {
      const int pi = i / BIAS_MAX, n = i % BIAS_MAX;
      float bv = 0.f;
      if (pi < prm.count && prm.p[pi].bias && n < (prm.p[pi].bias_per_row ? prm.p[pi].M : prm.p[pi].N))
        bv = __ldg(prm.p[pi].bias + n);
      bias_s[i] = bv;
    }
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], A_RAW ? (W16 ? 8 : 5) : 1);  // MMA commit (+ the 4 splitter warps that read the raw A tile); W16: the 8 splitter warps only
      mbar_init(&fullB_bar[s], 1);
      mbar_init(&emptyB_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], NUM_EPI_WARPS);  // one arrive per epilogue warp
    }
    for (int a = 0; a < ASETS; ++a) {
      mbar_init(&a_ready_bar[a], W16 ? 8 : 4);
      mbar_init(&a_free_bar[a], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // whole warp allocates the accumulators (+ the A staging columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = uniform_u32(*tmem_base_slot);

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the loop (warp-uniform control flow), one elected lane issues
    uint32_t it = 0;  // global k-block counter -> smem ring position
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
      const TileCoord tc = tile_coord<BN>(prm, t);
      const Problem& P = prm.p[tc.p];
      const int num_kb = P.K / KB;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * STAGE_BYTES;
        if (W16) {
          if (elect_one()) {
            mbar_expect_tx(&full_bar[s], (uint32_t)(2 * A_TILE_BYTES));
            tma_load_2d(st, &P.a_hi, &full_bar[s], kb * KB, tc.m0);                       // raw A, channels [0, 32)
            tma_load_2d(st + A_TILE_BYTES, &P.a_hi, &full_bar[s], kb * KB + BK, tc.m0);   // raw A, channels [32, 64)
          }
          __syncwarp();
          continue;
        }
        const uint32_t bytes = A_RAW ? (uint32_t)STAGE_BYTES
                                     : ((prm.passes == 3) ? (uint32_t)STAGE_BYTES : (uint32_t)STAGE_BYTES / 2);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[s], bytes);
          // B tiles are loaded as BN/128 boxes of 128 rows (the tensor-map box is 32 x 128)
          tma_load_2d(st + OFF_A_HI, &P.a_hi, &full_bar[s], kb * BK, tc.m0);  // A_RAW: a_hi holds the raw A map
#pragma unroll
          for (int h = 0; h < BN / 128; ++h)
            tma_load_2d(st + OFF_B_HI + h * A_TILE_BYTES, &P.b_hi, &full_bar[s], kb * BK, tc.n0 + h * 128);
          if (A_RAW || prm.passes == 3) {
            if (!A_RAW) tma_load_2d(st + OFF_A_LO, &P.a_lo, &full_bar[s], kb * BK, tc.m0);
#pragma unroll
            for (int h = 0; h < BN / 128; ++h)
              tma_load_2d(st + OFF_B_LO + h * A_TILE_BYTES, &P.b_lo, &full_bar[s], kb * BK, tc.n0 + h * 128);
          }
        }
        __syncwarp();
      }
    }
  } else if (W16 && warp == 2 + NUM_EPI_WARPS + 8) {
    // ===== weight-tile producer (W16): bf16 hi / lo planes, 64 channels = 128-byte rows
    uint32_t it = 0;
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
      const TileCoord tc = tile_coord<BN>(prm, t);
      const Problem& P = prm.p[tc.p];
      const int num_kb = P.K / KB;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&emptyB_bar[s], ((it / STAGES) & 1) ^ 1);
        uint8_t* st = smem + (size_t)s * STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&fullB_bar[s], (uint32_t)(2 * B_TILE));
          tma_load_2d(st + OFF_B_HI, &P.b_hi, &fullB_bar[s], kb * KB, tc.n0);
          tma_load_2d(st + OFF_B_LO, &P.b_lo, &fullB_bar[s], kb * KB, tc.n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues tcgen05.mma / commit
    const uint32_t idesc = make_idesc(BN);
    uint32_t it = 0, tile_it = 0;
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x, ++tile_it) {
      const TileCoord tc = tile_coord<BN>(prm, t);
      const int num_kb = prm.p[tc.p].K / KB;
      const uint32_t acc = tile_it & 1;
      mbar_wait(&tmem_empty_bar[acc], ((tile_it >> 1) & 1) ^ 1);  // epilogue drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(W16 ? &fullB_bar[s] : &full_bar[s], ph);
        const uint32_t st = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint64_t b_hi = make_smem_desc(st + OFF_B_HI), b_lo = make_smem_desc(st + OFF_B_LO);
        if (A_RAW) {
          const uint32_t set = it % ASETS;
          mbar_wait(&a_ready_bar[set], (it / ASETS) & 1);  // splitter has parked hi/lo of this k-block in TMEM
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi_t = tmem_base + TM_A + set * 64, a_lo_t = a_hi_t + 32;
          if (W16) {
            if (elect_one()) {
              const uint32_t idesc16 = make_idesc_bf16(BN);
#pragma unroll
              for (int k = 0; k < KB / UMMA_K_BF16; ++k) {  // 16 channels = 8 packed TMEM columns / 32 bytes of the B row
                const uint64_t koff = (uint64_t)((k * UMMA_K_BF16 * 2) >> 4);
                const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
                umma_bf16_ts(d_tmem, a_lo_t + k * 8, b_hi + koff, idesc16, first);
                umma_bf16_ts(d_tmem, a_hi_t + k * 8, b_lo + koff, idesc16, 1u);
                umma_bf16_ts(d_tmem, a_hi_t + k * 8, b_hi + koff, idesc16, 1u);
              }
              umma_commit(&emptyB_bar[s]);     // weight tiles of this stage
              umma_commit(&a_free_bar[set]);   // TMEM A set
            }
          } else if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
              const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
              umma_tf32_ts(d_tmem, a_lo_t + k * UMMA_K, b_hi + koff, idesc, first);
              umma_tf32_ts(d_tmem, a_hi_t + k * UMMA_K, b_lo + koff, idesc, 1u);
              umma_tf32_ts(d_tmem, a_hi_t + k * UMMA_K, b_hi + koff, idesc, 1u);
            }
            umma_commit(&empty_bar[s]);      // B tiles of this stage are free once these MMAs complete
            umma_commit(&a_free_bar[set]);   // ... and so is the TMEM A set
          }
        } else {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t a_hi = make_smem_desc(st + OFF_A_HI), a_lo = make_smem_desc(st + OFF_A_LO);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);  // advance inside the 128 B swizzle row
              const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
              if (prm.passes == 3) {
                umma_tf32(d_tmem, a_lo + koff, b_hi + koff, idesc, first);
                umma_tf32(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
                umma_tf32(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
              } else {
                umma_tf32(d_tmem, a_hi + koff, b_hi + koff, idesc, first);
              }
            }
            umma_commit(&empty_bar[s]);  // frees the smem stage once the MMAs above have read it
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
      __syncwarp();
    }
  } else if (is_epi) {
    // ===== epilogue: TMEM lane quadrant (warp % 4), column slice ((warp - 2) / 4)
    const int quad = warp & 3;
    const int chalf = (warp - 2) >> 2;
    constexpr int CH = BN / (NUM_EPI_WARPS / 4);  // columns per epilogue warp
    uint32_t tile_it = 0;
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x, ++tile_it) {
      const TileCoord tc = tile_coord<BN>(prm, t);
      const Problem& P = prm.p[tc.p];
      const uint32_t acc = tile_it & 1;
      mbar_wait(&tmem_full_bar[acc], (tile_it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = tc.m0 + quad * 32 + lane;
      const int n0 = tc.n0;
      const float* bias_t = bias_s + tc.p * BIAS_MAX + n0;  // n0 + 127 < BIAS_MAX (checked on the host)
      const float bias_row = P.bias_per_row ? bias_s[tc.p * BIAS_MAX + min(row, BIAS_MAX - 1)] : 0.f;
      if (P.bias_per_row) bias_t = bias_s + tc.p * BIAS_MAX;  // (unused columns; keep the address in range)
#pragma unroll 1
      for (int c0 = chalf * CH; c0 < (chalf + 1) * CH; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + (uint32_t)c0, v);
        if (c0 + 32 >= (chalf + 1) * CH) {  // last TMEM read of this tile: hand the accumulator back before storing
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc])) : "memory");
        }
        if (P.bits) {
          if (n0 + c0 >= P.N) continue;
          // one ballot per query column: the 32 lanes of this warp are 32 consecutive keys = one mask word
          const bool key_ok = row < P.M;
          uint32_t myword = 0xffffffffu;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const uint32_t wd = __ballot_sync(0xffffffffu, !key_ok || __uint_as_float(v[j]) < 0.f);
            if (lane == j) myword = wd;
          }
          const int q = n0 + c0 + lane;
          const int wi = (tc.m0 >> 5) + quad;
          if (q < P.N && wi < P.bits_words) {
            P.bits[(size_t)q * P.bits_words + wi] = myword;
            if (myword != 0xffffffffu) atomicOr(&P.rowany[q], 1);
          }
        } else if (row < P.M && n0 + c0 < P.N && P.t_rows > 0) {
          // transposed store: lanes hold consecutive rows -> each column is one coalesced 128-byte store
          const int bi = row / P.t_rows, r = row - bi * P.t_rows;
          const float* bias_c = bias_t + c0;
          if (!P.C_lo && !P.relu && !P.bias_per_row) {
            // fast path (plain store): one FADD + one STG per element, pointer bumped by the row pitch
            float* dst = P.C + ((size_t)bi * P.N + n0 + c0) * P.ldc + r;
            const size_t pitch = (size_t)P.ldc;
            const int nvalid = P.N - (n0 + c0);  // >= 1
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(bias_c + j);
              if (j + 3 < nvalid) {
                dst[0] = __uint_as_float(v[j]) + bb.x;
                dst[pitch] = __uint_as_float(v[j + 1]) + bb.y;
                dst[2 * pitch] = __uint_as_float(v[j + 2]) + bb.z;
                dst[3 * pitch] = __uint_as_float(v[j + 3]) + bb.w;
              } else {
                if (j < nvalid) dst[0] = __uint_as_float(v[j]) + bb.x;
                if (j + 1 < nvalid) dst[pitch] = __uint_as_float(v[j + 1]) + bb.y;
                if (j + 2 < nvalid) dst[2 * pitch] = __uint_as_float(v[j + 2]) + bb.z;
              }
              dst += 4 * pitch;
            }
            continue;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + c0 + j;
            if (n < P.N) {
              float x = __uint_as_float(v[j]) + bias_c[j];
              x = P.relu ? fmaxf(x, 0.f) : x;
              const size_t o = ((size_t)bi * P.N + n) * P.ldc + r;
              if (P.C_lo) {
                const float h = rna_tf32(x);
                P.C[o] = h;
                P.C_lo[o] = rna_tf32(x - h);
              } else {
                P.C[o] = x;
              }
            }
          }
        } else if (TMA_EPI && P.tma_store) {
          // row-major store through this warp's swizzled staging tile: ONE TMA store per 32 x 32 block (rows >= M and
          // columns >= N are clipped by the tensor map); the hi/lo split store sends two tiles
          if (tc.m0 + quad * 32 >= P.M || n0 + c0 >= P.N) continue;  // warp-uniform
          uint8_t* sb = epi_stage + (size_t)(warp - 2) * EPI_TILE_BYTES;
          float o[32];
          if (W16) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = (P.bias && n0 + c0 + j < P.N) ? __ldg(reinterpret_cast<const float4*>(P.bias + n0 + c0 + j))
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
              o[j] = __uint_as_float(v[j]) + bb.x; o[j + 1] = __uint_as_float(v[j + 1]) + bb.y;
              o[j + 2] = __uint_as_float(v[j + 2]) + bb.z; o[j + 3] = __uint_as_float(v[j + 3]) + bb.w;
            }
          } else if (P.bias_per_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]) + bias_row;
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(bias_t + c0 + j);
              o[j] = __uint_as_float(v[j]) + bb.x; o[j + 1] = __uint_as_float(v[j + 1]) + bb.y;
              o[j + 2] = __uint_as_float(v[j + 2]) + bb.z; o[j + 3] = __uint_as_float(v[j + 3]) + bb.w;
            }
          }
          if (P.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
          }
          const uint32_t sb_row = smem_u32(sb) + lane * 128;
          const uint32_t xr = (uint32_t)(lane & 7);
          auto store_tile = [&](const float (&t)[32], const CUtensorMap* map) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile's previous store
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb_row + (((uint32_t)j ^ xr) << 4)),
                           "f"(t[4 * j]), "f"(t[4 * j + 1]), "f"(t[4 * j + 2]), "f"(t[4 * j + 3])
                           : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                           "r"(smem_u32(sb)), "r"(n0 + c0), "r"(tc.m0 + quad * 32)
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          };
          if (!P.C_lo) {
            store_tile(o, &P.c_map);
          } else {
            float h[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) h[j] = rna_tf32(o[j]);
            store_tile(h, &P.c_map);
#pragma unroll
            for (int j = 0; j < 32; ++j) h[j] = rna_tf32(o[j] - h[j]);
            store_tile(h, &P.clo_map);
          }
        } else if (row < P.M && n0 + c0 < P.N) {
          float* dst = P.C + (size_t)row * P.ldc + n0 + c0;
          float* dlo = P.C_lo ? P.C_lo + (size_t)row * P.ldc + n0 + c0 : nullptr;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int n = n0 + c0 + j;
            float o[4];
            float4 bb = *reinterpret_cast<const float4*>(bias_t + c0 + j);
            if (P.bias_per_row) bb = make_float4(bias_row, bias_row, bias_row, bias_row);
            const float bq[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float x = __uint_as_float(v[j + u]) + bq[u];
              o[u] = P.relu ? fmaxf(x, 0.f) : x;
            }
            if (dlo) {
              float h[4], l[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                h[u] = rna_tf32(o[u]);
                l[u] = rna_tf32(o[u] - h[u]);
              }
              if (n + 3 < P.N) {
                *reinterpret_cast<float4*>(dst + j) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(dlo + j) = make_float4(l[0], l[1], l[2], l[3]);
              } else {
                for (int u = 0; u < 4; ++u)
                  if (n + u < P.N) { dst[j + u] = h[u]; dlo[j + u] = l[u]; }
              }
            } else if (n + 3 < P.N) {
              *reinterpret_cast<float4*>(dst + j) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
              for (int u = 0; u < 4; ++u)
                if (n + u < P.N) dst[j + u] = o[u];
            }
          }
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // outstanding TMA stores (smem + global)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else if (A_RAW) {
    // ===== A splitter: raw fp32 tile (smem, SWIZZLE_128B) -> hi / lo in tensor memory, one row per thread
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    // W16: TWO splitter groups of four warps, each takes one 32-channel half of every k-block.  One warp per scheduler
    // turns a 64-channel k-block around in ~0.9 us (wait -> LDS -> ~200 ALU -> tcgen05.st -> wait::st -> arrive: a
    // latency chain, measured) while its MMAs take 0.56 us; halving the work per warp puts the splitters back under the
    // tensor pipe.  (Alternating k-blocks between the groups instead would make each group skip mbarrier phases of the
    // 3-deep ring: a parity wait cannot tell phase n from phase n + 2.)
    const int grp = (warp - (2 + NUM_EPI_WARPS)) >> 2;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
      const TileCoord tc = tile_coord<BN>(prm, t);
      const int num_kb = prm.p[tc.p].K / KB;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t set = it % ASETS;
        mbar_wait(&full_bar[s], (it / STAGES) & 1);               // raw tile has landed
        mbar_wait(&a_free_bar[set], ((it / ASETS) & 1) ^ 1);      // MMAs of k-block it-ASETS are done with this set
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint8_t* arow = smem + (size_t)s * STAGE_BYTES + OFF_A_HI + r * 128;
        uint32_t hi[32], lo[32];
        if (W16) {
          // 64 channels of this row (two swizzled sub-tiles) -> 32 packed bf16 hi pairs + 32 packed lo pairs; the lower
          // channel of a pair sits in the lower half of the 32-bit TMEM column
          // this group's half: channels [32 grp, 32 grp + 32) = sub-tile grp -> 16 hi + 16 lo columns
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(arow + grp * A_TILE_BYTES + ((c ^ (r & 7)) << 4));
            const uint32_t h0 = pack_bf16x2(v.x, v.y), h1 = pack_bf16x2(v.z, v.w);
            hi[c * 2] = h0;
            hi[c * 2 + 1] = h1;
            lo[c * 2] = pack_bf16x2(v.x - __uint_as_float(h0 << 16), v.y - __uint_as_float(h0 & 0xffff0000u));
            lo[c * 2 + 1] = pack_bf16x2(v.z - __uint_as_float(h1 << 16), v.w - __uint_as_float(h1 & 0xffff0000u));
          }
          tmem_st_32x32b_x16(tmem_base + lane_addr + TM_A + set * 64 + grp * 16, hi);
          tmem_st_32x32b_x16(tmem_base + lane_addr + TM_A + set * 64 + 32 + grp * 16, lo);
          tmem_st_wait();
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&a_ready_bar[set]);
            mbar_arrive(&empty_bar[s]);
          }
          continue;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // logical 16-byte chunk c of the row lives at physical chunk c ^ (r & 7)
          const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (r & 7)) << 4));
          const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            // rna_tf32 in two integer ops: round to nearest, ties away from zero == add half an ulp to the magnitude and
            // clear the low 13 bits (cvt.rna.tf32.f32 compiles to ~9 instructions with its NaN / Inf selects); the lo
            // part only gets the half-ulp nudge, the tensor pipe's own truncation does the rest
            const uint32_t h = (__float_as_uint(x[u]) + 0x1000u) & 0xffffe000u;
            hi[c * 4 + u] = h;
            lo[c * 4 + u] = __float_as_uint(x[u] - __uint_as_float(h)) + 0x1000u;
          }
        }
        tmem_st_32x32b_x32(tmem_base + lane_addr + TM_A + set * 64, hi);
        tmem_st_32x32b_x32(tmem_base + lane_addr + TM_A + set * 64 + 32, lo);
        tmem_st_wait();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a_ready_bar[set]);
          mbar_arrive(&empty_bar[s]);  // this warp is done with the raw tile in smem
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- hi/lo split (round-to-nearest tf32) ---------------------------------------------------------------
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                          float* __restrict__ lo, size_t n4, float scale) {
  pdl_wait();     // PDL contract (common.cuh)
  pdl_trigger();
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    float4 h, l;
    h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
    l.x = rna_tf32(v.x - h.x); l.y = rna_tf32(v.y - h.y); l.z = rna_tf32(v.z - h.z); l.w = rna_tf32(v.w - h.w);
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
}

// weights -> bf16 hi / lo planes for the W16 variant: hi = bf16(w), lo = bf16(w - hi)   (round to nearest even)
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, uint32_t* __restrict__ hi,
                                                          uint32_t* __restrict__ lo, size_t n2) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n2; i += (size_t)gridDim.x * 256) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(x) + i);
    const uint32_t h = pack_bf16x2(v.x, v.y);
    hi[i] = h;
    lo[i] = pack_bf16x2(v.x - __uint_as_float(h << 16), v.y - __uint_as_float(h & 0xffff0000u));
  }
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* map, const float* ptr, long long rows, long long cols, long long ld, int box_cols,
                 int box_rows) {
  EncodeTiledFn fn = encode_fn();
  PN_REQUIRE(fn, PN_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  PN_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld * 4) % 16 == 0, PN_ERR_UNSUPPORTED, "umma: operand must be 16B aligned");
  PN_REQUIRE(box_cols * 4 == 128 && box_rows >= 1 && box_rows <= 256, PN_ERR_BAD_ARG, "umma: bad TMA box");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PN_REQUIRE(r == CUDA_SUCCESS, PN_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}
static int make_map(CUtensorMap* map, const float* ptr, int rows, int cols, int ld) {
  return make_tmap_2d(map, ptr, rows, cols, ld, BK, BM);
}
// bf16 plane [rows, cols] (ld elements): box 64 channels (= 128 B) x 128 rows, SWIZZLE_128B
static int make_map_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld) {
  EncodeTiledFn fn = encode_fn();
  PN_REQUIRE(fn, PN_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  PN_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld * 2) % 16 == 0, PN_ERR_UNSUPPORTED, "umma: bf16 operand must be 16B aligned");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PN_REQUIRE(r == CUDA_SUCCESS, PN_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled(bf16) failed (%d)", (int)r);
  return 0;
}

}  // namespace umma

// v [B,Nk,256] -> V^T per image: vt[(b*256 + c) * ldv + key] split hi/lo (32x32 smem transpose tiles)
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ v, float* __restrict__ vt_hi,
                                                               float* __restrict__ vt_lo, int Nk, int ldv) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, k0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int key = k0 + ty + r * 8;
    tile[ty + r * 8][tx] = key < Nk ? __ldg(v + ((size_t)b * Nk + key) * D + c0 + tx) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = c0 + ty + r * 8, key = k0 + tx;
    if (key < Nk) {
      const float x = tile[tx][ty + r * 8];
      const float h = umma::rna_tf32(x);
      const size_t o = ((size_t)b * D + c) * ldv + key;
      vt_hi[o] = h;
      vt_lo[o] = umma::rna_tf32(x - h);
    }
  }
}

int launch_split_transpose(const float* v, float* vt_hi, float* vt_lo, int B, int Nk, int ldv, cudaStream_t st) {
  dim3 grid(cdiv(Nk, 32), D / 32, B);
  split_transpose_kernel<<<grid, 256, 0, st>>>(v, vt_hi, vt_lo, Nk, ldv);
  return check_launch("split_transpose_kernel");
}

int launch_split_tf32(const float* x, float* hi, float* lo, size_t n, cudaStream_t st) {
  return launch_split_tf32_scaled(x, hi, lo, n, 1.0f, st);
}

int launch_split_bf16(const float* x, void* hi, void* lo, size_t n, cudaStream_t st) {
  PN_REQUIRE(x && hi && lo && n % 2 == 0, PN_ERR_BAD_ARG, "split_bf16: bad args");
  PN_REQUIRE((((uintptr_t)x & 7) | ((uintptr_t)hi & 3) | ((uintptr_t)lo & 3)) == 0, PN_ERR_UNSUPPORTED, "split_bf16: alignment");
  const size_t n2 = n / 2;
  int blocks = (int)((n2 + 255) / 256);
  blocks = blocks > 148 * 16 ? 148 * 16 : (blocks < 1 ? 1 : blocks);
  umma::split_bf16_kernel<<<blocks, 256, 0, st>>>(x, reinterpret_cast<uint32_t*>(hi), reinterpret_cast<uint32_t*>(lo), n2);
  return check_launch("split_bf16_kernel");
}

int launch_split_tf32_scaled(const float* x, float* hi, float* lo, size_t n, float scale, cudaStream_t st) {
  PN_REQUIRE(x && hi && lo && n % 4 == 0, PN_ERR_BAD_ARG, "split_tf32: bad args");
  PN_REQUIRE((((uintptr_t)x | (uintptr_t)hi | (uintptr_t)lo) & 15) == 0, PN_ERR_UNSUPPORTED, "split_tf32: alignment");
  const size_t n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  blocks = blocks > 148 * 16 ? 148 * 16 : (blocks < 1 ? 1 : blocks);
  launch_pdl(umma::split_tf32_kernel, dim3(blocks), dim3(256), 0, st, x, hi, lo, n4, scale);
  return check_launch("split_tf32_kernel");
}

int launch_umma_gemm(const UmmaOperand* ops, int count, int passes, cudaStream_t st) {
  using namespace umma;
  PN_REQUIRE(count >= 1 && count <= MAX_PROBLEMS, PN_ERR_BAD_ARG, "umma: bad problem count");
  PN_REQUIRE(passes == 1 || passes == 3, PN_ERR_BAD_ARG, "umma: passes must be 1 or 3");
  // PN_OPT_SINGLE_PASS: reduced-precision mode -- every tensor-core GEMM runs ONE kind::tf32 pass on the raw fp32
  // operands (the tensor pipe truncates them to TF32: 10-bit mantissa, >= bf16's 8); no hi/lo planes are read or written
  const bool single = get_option(OPT_SINGLE_PASS) != 0;
  if (single) passes = 1;
  Params prm{};
  prm.passes = passes;
  int maxM = 0, maxN = 0;
  for (int i = 0; i < count; ++i) {
    const UmmaOperand& o = ops[i];
    PN_REQUIRE(o.a_hi && o.w_hi && (o.C || o.bits) && (passes == 1 || ((o.a_lo || o.a_is_raw) && o.w_lo)), PN_ERR_BAD_ARG,
               "umma: null operand");
    PN_REQUIRE(!o.a_is_raw || passes == 3 || single, PN_ERR_BAD_ARG, "umma: raw A operands need passes == 3");
    PN_REQUIRE(o.K % BK == 0 && o.K >= BK, PN_ERR_UNSUPPORTED, "umma: K=%d must be a multiple of %d", o.K, BK);
    PN_REQUIRE(o.bias_per_row || cdiv(o.N, 256) * 256 <= BIAS_MAX, PN_ERR_UNSUPPORTED, "umma: N=%d exceeds %d", o.N,
               BIAS_MAX);
    PN_REQUIRE(o.bits || o.t_rows > 0 || (o.ldc % 4 == 0 && ((uintptr_t)o.C & 15) == 0), PN_ERR_UNSUPPORTED,
               "umma: C must be 16B aligned");
    Problem& p = prm.p[i];
    PN_TRY(make_map(&p.a_hi, o.a_hi, o.M, o.K, o.lda));
    if (o.w_bf16) {
      // 3xBF16: raw fp32 A, weights as bf16 hi / lo planes (same pointers, 2-byte elements)
      PN_REQUIRE(o.a_is_raw && passes == 3 && !single && o.K % 64 == 0 && !o.C_lo, PN_ERR_UNSUPPORTED,
                 "umma: bf16-split weights need a raw A operand, 3 passes and K %% 64 == 0");
      PN_TRY(make_map(&p.a_lo, o.a_hi, o.M, o.K, o.lda));
      PN_TRY(make_map_bf16(&p.b_hi, o.w_hi, o.N, o.K, o.ldw));
      PN_TRY(make_map_bf16(&p.b_lo, o.w_lo, o.N, o.K, o.ldw));
    } else {
      PN_TRY(make_map(&p.b_hi, o.w_hi, o.N, o.K, o.ldw));
      PN_TRY(make_map(&p.a_lo, (passes == 3 && !o.a_is_raw) ? o.a_lo : o.a_hi, o.M, o.K, o.lda));
      PN_TRY(make_map(&p.b_lo, passes == 3 ? o.w_lo : o.w_hi, o.N, o.K, o.ldw));
    }
    p.tma_store = 0;
    if (!o.bits && o.t_rows <= 0 && get_option(OPT_UMMA_TMA_STORE) && o.ldc % 4 == 0 && ((uintptr_t)o.C & 15) == 0) {
      PN_TRY(make_tmap_2d(&p.c_map, o.C, o.M, o.N, o.ldc, 32, 32));
      if (o.C_lo && !single) PN_TRY(make_tmap_2d(&p.clo_map, o.C_lo, o.M, o.N, o.ldc, 32, 32));
      p.tma_store = 1;
    }
    // the W16 kernel has no bias staging in smem and only the TMA-store epilogue
    PN_REQUIRE(!o.w_bf16 || (p.tma_store && !o.bias_per_row && !o.bits && o.t_rows <= 0 && ((uintptr_t)o.bias & 15) == 0),
               PN_ERR_UNSUPPORTED, "umma: the bf16-split variant needs the TMA-store epilogue (16B-aligned row-major C and "
               "bias, PN_OPT_UMMA_TMA_STORE = 1)");
    p.bias = o.bias; p.C = o.C; p.M = o.M; p.N = o.N; p.K = o.K; p.ldc = o.ldc;
    p.C_lo = single ? nullptr : o.C_lo; p.relu = o.relu; p.t_rows = o.t_rows; p.bias_per_row = o.bias_per_row;
    p.bits = o.bits; p.rowany = o.rowany; p.bits_words = o.bits_words;
    PN_REQUIRE(!o.bits || (o.rowany && o.bits_words > 0), PN_ERR_BAD_ARG, "umma: bit-pack epilogue needs rowany/words");
    PN_REQUIRE(!o.bias_per_row || o.M <= BIAS_MAX, PN_ERR_UNSUPPORTED, "umma: per-row bias needs M <= %d", BIAS_MAX);
    PN_REQUIRE(single || !o.C_lo || ((uintptr_t)o.C_lo & 15) == 0, PN_ERR_UNSUPPORTED, "umma: C_lo must be 16B aligned");
    maxM = o.M > maxM ? o.M : maxM;
    maxN = o.N > maxN ? o.N : maxN;
  }
  static bool attr_done[PN_MAX_DEVICES] = {false};  // the attribute is per device
  bool& attr_set = attr_done[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm_kernel<128, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg<128>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(umma_gemm_kernel<128, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)Cfg<128>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(umma_gemm_kernel<256, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)Cfg<256>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(umma_gemm_kernel<128, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)RAW_SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(umma_gemm_kernel<128, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)RAW_SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(umma_gemm_kernel<128, 8, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)RAW_SMEM_BYTES);
    PN_REQUIRE(e == cudaSuccess, (int)e, "umma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  bool raw = passes == 3;
  for (int i = 0; i < count; ++i) raw = raw && ops[i].a_is_raw;
  bool w16 = raw;
  for (int i = 0; i < count; ++i) w16 = w16 && ops[i].w_bf16;
  for (int i = 0; i < count; ++i)
    PN_REQUIRE(w16 || !ops[i].w_bf16, PN_ERR_BAD_ARG, "umma: bf16-split and TF32-split problems cannot share a launch");
  bool wide = !raw && get_option(OPT_UMMA_WIDE) != 0;
  for (int i = 0; i < count; ++i) wide = wide && (ops[i].N % 256 == 0);
  const int BN = wide ? 256 : 128;
  prm.count = count;
  prm.total_tiles = 0;
  for (int i = 0; i < count; ++i) prm.total_tiles += cdiv(ops[i].M, BM) * cdiv(ops[i].N, BN);
  const int num_sms = sm_count();
  const int grid = prm.total_tiles < num_sms ? prm.total_tiles : num_sms;
  if (w16)
    umma_gemm_kernel<128, 8, true, true><<<grid, 64 + 32 * 8 + 288, RAW_SMEM_BYTES, st>>>(prm);
  else if (raw && get_option(OPT_UMMA_EPI8))
    umma_gemm_kernel<128, 8, true><<<grid, 64 + 32 * 8 + 128, RAW_SMEM_BYTES, st>>>(prm);
  else if (raw)
    umma_gemm_kernel<128, 4, true><<<grid, 64 + 32 * 4 + 128, RAW_SMEM_BYTES, st>>>(prm);
  else if (wide)
    umma_gemm_kernel<256, 8, false><<<grid, 64 + 32 * 8, Cfg<256>::SMEM_BYTES, st>>>(prm);
  else if (get_option(OPT_UMMA_EPI8))
    umma_gemm_kernel<128, 8, false><<<grid, 64 + 32 * 8, Cfg<128>::SMEM_BYTES, st>>>(prm);
  else
    umma_gemm_kernel<128, 4, false><<<grid, 64 + 32 * 4, Cfg<128>::SMEM_BYTES, st>>>(prm);
  return check_launch("umma_gemm_kernel");
}

}  // namespace pn
