// Query-side linears (M = B*100 rows): latency-optimised 3xTF32 GEMM, C = act(A . W^T + bias) + resid.
//
// These GEMMs are 13 MFLOP each and sit on the critical path ~130 times per forward
// (pairnet_head.py:236-243 mask/cls MLPs, :322-326 PPN MLPs, every q/k/v/out/FFN projection of the
// 9 + 6 transformer layers).  The k-tiled shared-memory GEMM took 9-10 us per launch even with every operand
// L2-resident: 8 dependent (global load -> smem -> barrier -> 32 k of FFMA) rounds with one warp per scheduler.
//
// Here a CTA owns one 16 x (8*NT) output tile and its 4 warps split K.  There is NO shared-memory staging and no
// k-loop dependency: every lane requests its whole operand slice (float4 along K) up front -- one exposed L2/HBM
// round trip per kernel -- and feeds it to warp-level tensor-core MMAs (mma.sync m16n8k8 tf32; the 16-row tile
// granule is what a 200-row problem needs to fill 148 SMs -- tcgen05's 128-row granule would leave 2..16 CTAs).
// fp32 parity comes from the same hi/lo split as the tcgen05 kernels: acc += lo*hi + hi*lo + hi*hi.
//
// K permutation: inside a 16-k chunk lane t (= lane & 3) holds physical k = 4t..4t+3 of rows g / g+8 (A) and of
// weight row g (W); MMA step 0 consumes (.x, .y) as logical k = (t, t+4), step 1 consumes (.z, .w).  A and W use
// the same permutation, so the dot products are unchanged.
#include "common.cuh"

namespace pn {

// hi = x rounded to tf32 (nearest, ties away: add half an ulp of the 10-bit mantissa, clear the 13 low bits --
// cvt.rna.tf32.f32 without its NaN/Inf selects); the residual x - hi is exact in fp32 and is truncated to tf32.
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x, uint32_t hi) {
  return __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int SK_WARPS = 4;
constexpr int SK_CHUNKS = 4;  // 16-k chunks requested per lane per round (64 k per warp)

template <int MT, int NT>
__global__ void __launch_bounds__(32 * SK_WARPS) skinny_gemm_kernel(const __grid_constant__ GemmBatch batch) {
  constexpr int BM = 16 * MT, BN = 8 * NT;
  __shared__ __align__(16) float red[SK_WARPS][BM][BN + 4];

  int z = blockIdx.z, pi = 0;
  for (; pi < batch.count - 1; ++pi) {
    const int nz = batch.p[pi].nb * batch.p[pi].splits;
    if (z < nz) break;
    z -= nz;
  }
  const GemmProb& P = batch.p[pi];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= P.M || n0 >= P.N) return;
  const int b = z / P.splits, s = z % P.splits;
  const int kper = P.K / P.splits;
  const int kw = kper / SK_WARPS;  // k extent of one warp (multiple of 16, checked on the host)
  const float* A = P.A + (size_t)b * P.sA;
  const float* W = P.W + (size_t)b * P.sW;
  float* C = P.C + (size_t)b * P.sC + (size_t)s * P.split_stride;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int kbase = s * kper + warp * kw + 4 * t;
  const float* ap[2 * MT];  // rows g + 8 h of the tile
  bool a_ok[2 * MT];
#pragma unroll
  for (int h = 0; h < 2 * MT; ++h) {
    a_ok[h] = m0 + g + 8 * h < P.M;
    ap[h] = A + (size_t)(m0 + g + 8 * h) * P.lda + kbase;
  }
  const float* wp[NT];
  bool w_ok[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    w_ok[j] = n0 + j * 8 + g < P.N;
    wp[j] = W + (size_t)(n0 + j * 8 + g) * P.ldw + kbase;
  }

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[i][j][u] = 0.f;

  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int nch = kw / 16;
  // PDL: model weights of the first round are requested BEFORE waiting for the producer of A (the previous kernel of the
  // chain) -- the weight round trip to L2 overlaps that kernel's tail; activations only after the wait
  const bool w_early = P.w_static != 0;
  for (int c0 = 0; c0 < nch; c0 += SK_CHUNKS) {
    float4 ra[SK_CHUNKS][2 * MT], rw[SK_CHUNKS][NT];
    if (c0 > 0 || w_early) {
#pragma unroll
      for (int c = 0; c < SK_CHUNKS; ++c) {
        const bool live = c0 + c < nch;
        const int ko = (c0 + c) * 16;
#pragma unroll
        for (int j = 0; j < NT; ++j)
          rw[c][j] = (live && w_ok[j]) ? __ldg(reinterpret_cast<const float4*>(wp[j] + ko)) : zero4;
      }
    }
    if (c0 == 0) {
      pdl_wait();
      pdl_trigger();
      if (!w_early) {
#pragma unroll
        for (int c = 0; c < SK_CHUNKS; ++c) {
          const bool live = c < nch;
          const int ko = c * 16;
#pragma unroll
          for (int j = 0; j < NT; ++j)
            rw[c][j] = (live && w_ok[j]) ? __ldg(reinterpret_cast<const float4*>(wp[j] + ko)) : zero4;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < SK_CHUNKS; ++c) {
      const bool live = c0 + c < nch;
      const int ko = (c0 + c) * 16;
#pragma unroll
      for (int h = 0; h < 2 * MT; ++h)
        ra[c][h] = (live && a_ok[h]) ? __ldg(reinterpret_cast<const float4*>(ap[h] + ko)) : zero4;
    }
#pragma unroll
    for (int c = 0; c < SK_CHUNKS; ++c) {
      if (c0 + c >= nch) break;
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const float lo4[4] = {ra[c][2 * i].x, ra[c][2 * i].y, ra[c][2 * i].z, ra[c][2 * i].w};
          const float hi4[4] = {ra[c][2 * i + 1].x, ra[c][2 * i + 1].y, ra[c][2 * i + 1].z, ra[c][2 * i + 1].w};
          // fragment order: a0 (g, t), a1 (g+8, t), a2 (g, t+4), a3 (g+8, t+4)
          const float ax[4] = {lo4[2 * st], hi4[2 * st], lo4[2 * st + 1], hi4[2 * st + 1]};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            ah[i][u] = tf32_hi(ax[u]);
            al[i][u] = tf32_lo(ax[u], ah[i][u]);
          }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float wv[4] = {rw[c][j].x, rw[c][j].y, rw[c][j].z, rw[c][j].w};
          const float b0 = wv[2 * st], b1 = wv[2 * st + 1];
          const uint32_t bh0 = tf32_hi(b0), bh1 = tf32_hi(b1);
          const uint32_t bl0 = tf32_lo(b0, bh0), bl1 = tf32_lo(b1, bh1);
#pragma unroll
          for (int i = 0; i < MT; ++i) {
            mma_tf32(acc[i][j], al[i], bh0, bh1);
            mma_tf32(acc[i][j], ah[i], bl0, bl1);
            mma_tf32(acc[i][j], ah[i], bh0, bh1);
          }
        }
      }
    }
  }

  // ---- reduce the 4 k-slices through shared memory; C fragment: c0 (g, 2t), c1 (g, 2t+1), c2/c3 rows g+8
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      *reinterpret_cast<float2*>(&red[warp][16 * i + g][j * 8 + 2 * t]) = make_float2(acc[i][j][0], acc[i][j][1]);
      *reinterpret_cast<float2*>(&red[warp][16 * i + g + 8][j * 8 + 2 * t]) = make_float2(acc[i][j][2], acc[i][j][3]);
    }
  __syncthreads();
  const bool epi = (P.splits == 1);
  const float* resid = (epi && P.resid) ? P.resid + (size_t)b * P.sR : nullptr;
  for (int o = threadIdx.x; o < BM * BN; o += 32 * SK_WARPS) {
    const int r = o / BN, cidx = o % BN;
    const int m = m0 + r, n = n0 + cidx;
    if (m >= P.M || n >= P.N) continue;
    float x = (red[0][r][cidx] + red[1][r][cidx]) + (red[2][r][cidx] + red[3][r][cidx]);
    if (epi) {
      if (P.bias) x += __ldg(P.bias + n);
      if (P.relu) x = fmaxf(x, 0.f);
      if (resid) x += __ldg(resid + (size_t)m * P.ldr + n);
    }
    C[(size_t)m * P.ldc + n] = x;
  }
}

// true when every problem of the batch fits the skinny kernel (k extent per split a multiple of 64)
bool skinny_gemm_ok(const GemmBatch& batch) {
  for (int i = 0; i < batch.count; ++i) {
    const GemmProb& p = batch.p[i];
    if (p.splits < 1 || p.K % p.splits != 0) return false;
    const int kper = p.K / p.splits;
    // longer reductions stay on the exact-fp32 FFMA kernel (the product splits its K = 2048 GEMM 8 ways)
    if (kper % (16 * SK_WARPS) != 0 || kper > 512) return false;
  }
  return true;
}

int launch_skinny_gemm(const GemmBatch& batch, int maxM, int maxN, int nz, cudaStream_t st) {
  const int variant = get_option(OPT_SKINNY);  // 1 = auto; 11.. = forced tile (A/B studies)
#define PN_SKINNY(MT, NT)                                                    \
  do {                                                                       \
    dim3 grid(cdiv(maxN, 8 * NT), cdiv(maxM, 16 * MT), nz);                  \
    launch_pdl(skinny_gemm_kernel<MT, NT>, grid, dim3(32 * SK_WARPS), 0, st, batch); \
  } while (0)
  const long long tiles11 = (long long)cdiv(maxN, 8) * cdiv(maxM, 16) * nz;
  int v = variant;
  if (v >= 100) v = tiles11 > 600 ? v - 100 : 11;  // A/B studies: auto rule with another tile for the wide problems
  else if (v < 10) v = tiles11 > 600 ? 12 : 11;    // measured on B200 (scratch/skinny_ab.py)
  switch (v) {
    case 11: PN_SKINNY(1, 1); break;
    case 12: PN_SKINNY(1, 2); break;
    case 14: PN_SKINNY(1, 4); break;
    case 22: PN_SKINNY(2, 2); break;
    default: PN_SKINNY(2, 4); break;
  }
#undef PN_SKINNY
  return check_launch("skinny_gemm_kernel");
}

}  // namespace pn
