// fp32 SIMT tiled GEMMs for the query-side (M = B*100 rows) and memory-side (M = B*hw rows)
// projections, and the "bqc,bchw->bqhw" einsum with either an fp32 store or a fused
// sign -> bit-pack epilogue (the boolean attention mask of pairnet_head.py:244-256, shared by the
// 8 heads instead of being repeated 8x).
//
// Exact-fp32 (FFMA) path: parity config 2 is fp32 and the attention-mask threshold is
// sign-sensitive, so products are never rounded to tf32/bf16 here.
#include "common.cuh"

namespace pn {

template <int BM, int BN, int BK, int TM, int TN, bool B_NMAJOR>
struct GemmCore {
  static constexpr int NTX = BN / TN;
  static constexpr int NTY = BM / TM;
  static constexpr int NT = NTX * NTY;
  static constexpr int LDA_S = BM + 4;
  static constexpr int LDB_S = BN + 4;
  static constexpr int A_F4 = BM * BK / 4;
  static constexpr int B_F4 = BN * BK / 4;
  static constexpr int A_PER = A_F4 / NT;
  static constexpr int B_PER = B_F4 / NT;
  static_assert(A_F4 % NT == 0 && B_F4 % NT == 0, "tile/threads mismatch");
  static_assert(TM == 2 || TM == 4 || TM == 8, "TM");
  static_assert(TN == 4 || TN == 8, "TN");

  struct Smem {
    float a[2][BK][LDA_S];
    float b[2][BK][LDB_S];
  };

  __device__ static __forceinline__ int row_of(int ty, int i) {
    if (TM == 2) return ty * 2 + i;
    return (i < 4) ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4);
  }
  __device__ static __forceinline__ int col_of(int tx, int j) {
    return (j < 4) ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4);
  }

  // A: [M,K] row-major (lda).  B: K-major W[N,K] (ldw) or N-major F[K,N] (ldw).
  __device__ static __forceinline__ void load_a(float4 (&r)[A_PER], const float* __restrict__ A, int lda, int M,
                                                int m0, int k0, int tid) {
#pragma unroll
    for (int t = 0; t < A_PER; ++t) {
      int i = tid + t * NT;
      int row = i / (BK / 4), kq = i % (BK / 4);
      int m = m0 + row;
      r[t] = (m < M) ? __ldg(reinterpret_cast<const float4*>(A + (size_t)m * lda + k0 + kq * 4))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ static __forceinline__ void store_a(const float4 (&r)[A_PER], float (*sa)[LDA_S], int tid) {
#pragma unroll
    for (int t = 0; t < A_PER; ++t) {
      int i = tid + t * NT;
      int row = i / (BK / 4), kq = i % (BK / 4);
      sa[kq * 4 + 0][row] = r[t].x;
      sa[kq * 4 + 1][row] = r[t].y;
      sa[kq * 4 + 2][row] = r[t].z;
      sa[kq * 4 + 3][row] = r[t].w;
    }
  }
  __device__ static __forceinline__ void load_b(float4 (&r)[B_PER], const float* __restrict__ W, int ldw, int N,
                                                int n0, int k0, int tid, bool vec_ok) {
#pragma unroll
    for (int t = 0; t < B_PER; ++t) {
      int i = tid + t * NT;
      if (!B_NMAJOR) {
        int row = i / (BK / 4), kq = i % (BK / 4);
        int n = n0 + row;
        r[t] = (n < N) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)n * ldw + k0 + kq * 4))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        int k = i / (BN / 4), nq = i % (BN / 4);
        int n = n0 + nq * 4;
        const float* src = W + (size_t)(k0 + k) * ldw + n;
        if (vec_ok && n + 3 < N) {
          r[t] = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n + 0 < N) v.x = __ldg(src + 0);
          if (n + 1 < N) v.y = __ldg(src + 1);
          if (n + 2 < N) v.z = __ldg(src + 2);
          if (n + 3 < N) v.w = __ldg(src + 3);
          r[t] = v;
        }
      }
    }
  }
  __device__ static __forceinline__ void store_b(const float4 (&r)[B_PER], float (*sb)[LDB_S], int tid) {
#pragma unroll
    for (int t = 0; t < B_PER; ++t) {
      int i = tid + t * NT;
      if (!B_NMAJOR) {
        int row = i / (BK / 4), kq = i % (BK / 4);
        sb[kq * 4 + 0][row] = r[t].x;
        sb[kq * 4 + 1][row] = r[t].y;
        sb[kq * 4 + 2][row] = r[t].z;
        sb[kq * 4 + 3][row] = r[t].w;
      } else {
        int k = i / (BN / 4), nq = i % (BN / 4);
        *reinterpret_cast<float4*>(&sb[k][nq * 4]) = r[t];
      }
    }
  }

  __device__ static __forceinline__ void mainloop(float (&acc)[TM][TN], Smem& sm, const float* __restrict__ A,
                                                  int lda, int M, const float* __restrict__ W, int ldw, int N,
                                                  int m0, int n0, int k_begin, int k_end, bool vec_ok) {
    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[A_PER], rb[B_PER];
    load_a(ra, A, lda, M, m0, k_begin, tid);
    load_b(rb, W, ldw, N, n0, k_begin, tid, vec_ok);
    store_a(ra, sm.a[0], tid);
    store_b(rb, sm.b[0], tid);
    __syncthreads();
    int buf = 0;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
      const bool has_next = (k0 + BK) < k_end;
      if (has_next) {
        load_a(ra, A, lda, M, m0, k0 + BK, tid);
        load_b(rb, W, ldw, N, n0, k0 + BK, tid, vec_ok);
      }
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM < 4 ? 4 : TM], b[TN];
        if (TM == 2)
          *reinterpret_cast<float2*>(&a[0]) = *reinterpret_cast<const float2*>(&sm.a[buf][kk][ty * 2]);
        else
          *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&sm.a[buf][kk][ty * 4]);
        if (TM == 8)
          *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&sm.a[buf][kk][BM / 2 + ty * 4]);
        *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&sm.b[buf][kk][tx * 4]);
        if (TN == 8)
          *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&sm.b[buf][kk][BN / 2 + tx * 4]);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (has_next) {
        store_a(ra, sm.a[buf ^ 1], tid);
        store_b(rb, sm.b[buf ^ 1], tid);
        __syncthreads();
        buf ^= 1;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// grouped / batched / split-K GEMM with bias + ReLU + residual epilogue
// ------------------------------------------------------------------------------------------------
template <int BM, int BN, int BK, int TM, int TN, bool B_NMAJOR>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_store_kernel(const __grid_constant__ GemmBatch batch) {
  using Core = GemmCore<BM, BN, BK, TM, TN, B_NMAJOR>;
  __shared__ __align__(16) typename Core::Smem sm;

  int z = blockIdx.z, pi = 0;
  for (; pi < batch.count - 1; ++pi) {
    int nz = batch.p[pi].nb * batch.p[pi].splits;
    if (z < nz) break;
    z -= nz;
  }
  const GemmProb& P = batch.p[pi];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= P.M || n0 >= P.N) return;
  const int b = z / P.splits, s = z % P.splits;
  const int kper = P.K / P.splits;
  const float* A = P.A + (size_t)b * P.sA;
  const float* W = P.W + (size_t)b * P.sW;
  float* C = P.C + (size_t)b * P.sC + (size_t)s * P.split_stride;
  const bool wvec = B_NMAJOR ? ((P.ldw & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) & 15) == 0)) : true;

  float acc[TM][TN];
  Core::mainloop(acc, sm, A, P.lda, P.M, W, P.ldw, P.N, m0, n0, s * kper, (s + 1) * kper, wvec);

  const int tid = threadIdx.x;
  const int tx = tid % Core::NTX, ty = tid / Core::NTX;
  const bool epi = (P.splits == 1);
  const float* resid = (epi && P.resid) ? P.resid + (size_t)b * P.sR : nullptr;
  const bool cvec = ((P.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                    (!resid || (((P.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(resid) & 15) == 0)));
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + Core::row_of(ty, i);
    if (m >= P.M) continue;
#pragma unroll
    for (int jg = 0; jg < TN / 4; ++jg) {
      const int n = n0 + Core::col_of(tx, jg * 4);
      if (n >= P.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = acc[i][jg * 4 + j];
        if (epi) {
          if (P.bias && n + j < P.N) x += __ldg(P.bias + n + j);
          if (P.relu) x = fmaxf(x, 0.f);
          if (resid && n + j < P.N) x += __ldg(resid + (size_t)m * P.ldr + n + j);
        }
        v[j] = x;
      }
      float* dst = C + (size_t)m * P.ldc + n;
      if (cvec && n + 3 < P.N) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < P.N) dst[j] = v[j];
      }
    }
  }
}


GemmProb make_linear(const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int M,
                     int N, int K, int relu, const float* resid, int ldr) {
  GemmProb p{};
  p.A = A; p.W = W; p.bias = bias; p.resid = resid; p.C = C;
  p.M = M; p.N = N; p.K = K;
  p.lda = lda; p.ldw = K; p.ldc = ldc; p.ldr = ldr;
  p.relu = relu;
  p.nb = 1; p.sA = p.sW = p.sC = p.sR = 0;
  p.splits = 1; p.split_stride = 0;
  return p;
}

static int validate(const GemmProb& p, int bk) {
  PN_REQUIRE(p.A && p.W && p.C, PN_ERR_BAD_ARG, "gemm: null pointer");
  PN_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, PN_ERR_BAD_ARG, "gemm: bad shape %d %d %d", p.M, p.N, p.K);
  PN_REQUIRE(p.splits >= 1 && p.nb >= 1, PN_ERR_BAD_ARG, "gemm: bad splits/nb");
  PN_REQUIRE(p.K % (p.splits * bk) == 0, PN_ERR_UNSUPPORTED, "gemm: K=%d not a multiple of splits*BK=%d", p.K,
             p.splits * bk);
  PN_REQUIRE((p.lda & 3) == 0 && ((uintptr_t)p.A & 15) == 0, PN_ERR_UNSUPPORTED, "gemm: A not 16B aligned");
  PN_REQUIRE(p.splits == 1 || (!p.bias && !p.resid && !p.relu), PN_ERR_BAD_ARG,
             "gemm: split-K problems cannot carry an epilogue");
  return 0;
}

int launch_gemm(const GemmBatch& batch, cudaStream_t st) {
  PN_REQUIRE(batch.count >= 1 && batch.count <= GEMM_MAX_PROBS, PN_ERR_BAD_ARG, "gemm: bad problem count");
  int maxM = 0, maxN = 0, nz = 0;
  for (int i = 0; i < batch.count; ++i) {
    maxM = batch.p[i].M > maxM ? batch.p[i].M : maxM;
    maxN = batch.p[i].N > maxN ? batch.p[i].N : maxN;
    nz += batch.p[i].nb * batch.p[i].splits;
  }
  const bool big = maxM >= 1024;
  const int bk = big ? 16 : 32;
  for (int i = 0; i < batch.count; ++i) {
    PN_TRY(validate(batch.p[i], bk));
    PN_REQUIRE((batch.p[i].ldw & 3) == 0 && ((uintptr_t)batch.p[i].W & 15) == 0, PN_ERR_UNSUPPORTED,
               "gemm: W not 16B aligned");
  }
  if (!big && get_option(OPT_TENSOR_CORES) && get_option(OPT_SKINNY) && skinny_gemm_ok(batch))
    return launch_skinny_gemm(batch, maxM, maxN, nz, st);
  if (big) {
    dim3 grid(cdiv(maxN, 128), cdiv(maxM, 128), nz);
    gemm_store_kernel<128, 128, 16, 8, 8, false><<<grid, 256, 0, st>>>(batch);
  } else {
    dim3 grid(cdiv(maxN, 64), cdiv(maxM, 32), nz);
    gemm_store_kernel<32, 64, 32, 4, 4, false><<<grid, 128, 0, st>>>(batch);
  }
  return check_launch("gemm_store_kernel");
}

int launch_gemm_nmajor_store(const GemmProb& p, cudaStream_t st) {
  PN_TRY(validate(p, 16));
  GemmBatch batch{};
  batch.p[0] = p;
  batch.count = 1;
  dim3 grid(cdiv(p.N, 128), cdiv(p.M, 128), p.nb * p.splits);
  gemm_store_kernel<128, 128, 16, 8, 8, true><<<grid, 256, 0, st>>>(batch);
  return check_launch("gemm_store_kernel<nmajor>");
}

// ------------------------------------------------------------------------------------------------
// einsum + sign + bit-pack: bits[b][q][p/32] bit (p%32) = (sum_c E[b,q,c] F[b,c,p] < 0)
// ------------------------------------------------------------------------------------------------
template <int BM, int BN, int BK, int TM>
__global__ void __launch_bounds__((BM / TM) * (BN / 4))
gemm_maskbits_kernel(const float* __restrict__ E, const float* __restrict__ F, uint32_t* __restrict__ bits,
                     int* __restrict__ rowany, int Nq, int hw, int ldf) {
  using Core = GemmCore<BM, BN, BK, TM, 4, true>;
  static_assert(Core::NTX == 32, "one warp must span the tile width");
  __shared__ __align__(16) typename Core::Smem sm;
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const float* A = E + (size_t)b * Nq * D;
  const float* W = F + (size_t)b * D * ldf;
  float acc[TM][4];
  Core::mainloop(acc, sm, A, D, Nq, W, ldf, ldf, m0, n0, 0, D, true);

  const int tid = threadIdx.x;
  const int tx = tid % 32, ty = tid / 32;
  const int words = ldf / 32;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + Core::row_of(ty, i);
    uint32_t nib = 0;
    bool any_open = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < hw) {
        const bool blocked = acc[i][j] < 0.f;
        nib |= (blocked ? 1u : 0u) << j;
        any_open |= !blocked;
      } else {
        nib |= 1u << j;
      }
    }
    uint32_t word = nib << (4 * (tx & 7));
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    word |= __shfl_xor_sync(0xffffffffu, word, 4);
    const bool row_open = __any_sync(0xffffffffu, any_open);
    if (m < Nq) {
      const int wi = n0 / 32 + tx / 8;
      if ((tx & 7) == 0 && wi < words) bits[((size_t)b * Nq + m) * words + wi] = word;
      if (tx == 0 && row_open) atomicOr(&rowany[b * Nq + m], 1);
    }
  }
}

int launch_gemm_nmajor_maskbits(const float* E, const float* F, uint32_t* bits, int* rowany, int B, int N,
                                int hw, int ldf, cudaStream_t st) {
  PN_REQUIRE(E && F && bits && rowany, PN_ERR_BAD_ARG, "attn_mask_bits: null pointer");
  PN_REQUIRE(ldf % 32 == 0 && ldf >= hw && hw > 0, PN_ERR_BAD_ARG, "attn_mask_bits: ldf must be a multiple of 32 >= hw");
  PN_REQUIRE(((uintptr_t)F & 15) == 0 && ((uintptr_t)E & 15) == 0, PN_ERR_UNSUPPORTED, "attn_mask_bits: unaligned");
  dim3 grid(cdiv(ldf, 128), cdiv(N, 128), B);
  gemm_maskbits_kernel<128, 128, 16, 8><<<grid, 512, 0, st>>>(E, F, bits, rowany, N, hw, ldf);
  return check_launch("gemm_maskbits_kernel");
}

}  // namespace pn
