// Pair Proposal Network kernels:
//   * ConvTiny "Matrix Learner" filter on the N x N pair matrix: 3 x conv7x7(pad 3), 1->C->C->1,
//     ReLU between (cnn_factory.py:6-53).  conv2 (C->C, 98 % of the flops) is an smem-tiled
//     implicit GEMM on FFMA; conv1/conv3 are direct.
//   * top-k pair select: exact radix select on order-preserving keys + bitonic sort of the k
//     survivors; contract = descending value, ties by ascending flat index; emits
//     sub_pos = idx / N, obj_pos = idx % N as int64 (pairnet_head.py:334-340) and, fused, the
//     subject (+) object query gather/concat (pairnet_head.py:342-351,365).
#include "common.cuh"

namespace pn {

// ------------------------------------------------------------------------------------------ conv1
// x [B,N,N] -> y [B,C,N,N] = relu(conv7x7(x) + b) ; w [C,1,7,7]
template <int C>
__global__ void __launch_bounds__(256) conv1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ y, int N) {
  __shared__ float in_s[22][24];
  __shared__ __align__(16) float w_s[C][52];
  __shared__ float b_s[C];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * 16, y0 = blockIdx.y * 16;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* src = x + (size_t)b * N * N;
  for (int i = tid; i < 22 * 22; i += 256) {
    const int r = i / 22, c = i % 22;
    const int gy = y0 + r - 3, gx = x0 + c - 3;
    in_s[r][c] = (gy >= 0 && gy < N && gx >= 0 && gx < N) ? __ldg(src + (size_t)gy * N + gx) : 0.f;
  }
  for (int i = tid; i < C * 52; i += 256) {
    const int co = i / 52, k = i % 52;
    w_s[co][k] = (k < 49) ? __ldg(w + co * 49 + k) : 0.f;
  }
  for (int i = tid; i < C; i += 256) b_s[i] = __ldg(bias + i);
  __syncthreads();
  float v[52];
#pragma unroll
  for (int ky = 0; ky < 7; ++ky)
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) v[ky * 7 + kx] = in_s[ty + ky][tx + kx];
  v[49] = v[50] = v[51] = 0.f;
  const int gy = y0 + ty, gx = x0 + tx;
  const bool valid = gy < N && gx < N;
  for (int co = 0; co < C; ++co) {
    float acc = b_s[co];
#pragma unroll
    for (int k4 = 0; k4 < 13; ++k4) {
      const float4 ww = *reinterpret_cast<const float4*>(&w_s[co][k4 * 4]);
      acc = fmaf(v[k4 * 4 + 0], ww.x, acc);
      acc = fmaf(v[k4 * 4 + 1], ww.y, acc);
      acc = fmaf(v[k4 * 4 + 2], ww.z, acc);
      acc = fmaf(v[k4 * 4 + 3], ww.w, acc);
    }
    if (valid) y[(((size_t)b * C + co) * N + gy) * N + gx] = fmaxf(acc, 0.f);
  }
}

// ------------------------------------------------------------------------------------------ conv2
// w [C,C,7,7] (torch) -> wp [ci][49][co]
__global__ void pack_conv2_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int C) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= C * C * 49) return;
  const int co = i % C, k = (i / C) % 49, ci = i / (C * 49);
  wp[i] = __ldg(w + ((size_t)co * C + ci) * 49 + k);
}

// x [B,C,N,N] -> y [B,C,N,N] = relu(conv7x7 + b).  CTA = 8 x 16 output pixels x all C channels.
// thread = (pixel group of 4 consecutive x, channel group of 8): acc[4][8].
constexpr int C2_CI = 4;  // input channels per smem stage
template <int C>
__global__ void __launch_bounds__(32 * (C / 8)) conv2_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ wp,
                                                              const float* __restrict__ bias,
                                                              float* __restrict__ y, int N) {
  constexpr int NT = 32 * (C / 8);
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;                          // [C2_CI][49][C]
  float* in_s = smem + C2_CI * 49 * C;        // [C2_CI][14][24]
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * 16, y0 = blockIdx.y * 8;
  const int tid = threadIdx.x;
  const int pg = tid & 31, cg = tid >> 5;
  const int prow = pg >> 2, xseg = (pg & 3) * 4;
  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[p][c] = 0.f;

  for (int ci0 = 0; ci0 < C; ci0 += C2_CI) {
    __syncthreads();
    {  // weights: contiguous C2_CI*49*C floats
      const float4* src = reinterpret_cast<const float4*>(wp + (size_t)ci0 * 49 * C);
      float4* dst = reinterpret_cast<float4*>(w_s);
      for (int i = tid; i < C2_CI * 49 * C / 4; i += NT) dst[i] = __ldg(src + i);
    }
    for (int i = tid; i < C2_CI * 14 * 22; i += NT) {
      const int ci = i / (14 * 22), r = (i / 22) % 14, c = i % 22;
      const int gy = y0 + r - 3, gx = x0 + c - 3;
      in_s[(ci * 14 + r) * 24 + c] =
          (gy >= 0 && gy < N && gx >= 0 && gx < N) ? __ldg(x + (((size_t)b * C + ci0 + ci) * N + gy) * N + gx) : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < C2_CI; ++ci) {
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        const float* rowp = &in_s[(ci * 14 + prow + ky) * 24 + xseg];
        float r[10];
        *reinterpret_cast<float4*>(&r[0]) = *reinterpret_cast<const float4*>(rowp);
        *reinterpret_cast<float4*>(&r[4]) = *reinterpret_cast<const float4*>(rowp + 4);
        *reinterpret_cast<float2*>(&r[8]) = *reinterpret_cast<const float2*>(rowp + 8);
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float* wk = &w_s[((ci * 49) + ky * 7 + kx) * C + cg * 8];
          const float4 wa = *reinterpret_cast<const float4*>(wk);
          const float4 wb = *reinterpret_cast<const float4*>(wk + 4);
          const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[p][c] = fmaf(r[p + kx], wv[c], acc[p][c]);
        }
      }
    }
  }
  const int gy = y0 + prow;
  if (gy < N) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int co = cg * 8 + c;
      const float bv = __ldg(bias + co);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int gx = x0 + xseg + p;
        if (gx < N) y[(((size_t)b * C + co) * N + gy) * N + gx] = fmaxf(acc[p][c] + bv, 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ conv3
// x [B,C,N,N] -> y [B,N,N] = conv7x7 + b ; w [1,C,7,7]
template <int C>
__global__ void __launch_bounds__(256) conv3_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ y, int N) {
  constexpr int CI = 8;
  __shared__ float in_s[CI][22][24];
  __shared__ float w_s[C * 49];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * 16, y0 = blockIdx.y * 16;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int i = tid; i < C * 49; i += 256) w_s[i] = __ldg(w + i);
  float acc = 0.f;
  for (int ci0 = 0; ci0 < C; ci0 += CI) {
    __syncthreads();
    for (int i = tid; i < CI * 22 * 22; i += 256) {
      const int ci = i / (22 * 22), r = (i / 22) % 22, c = i % 22;
      const int gy = y0 + r - 3, gx = x0 + c - 3;
      in_s[ci][r][c] =
          (gy >= 0 && gy < N && gx >= 0 && gx < N) ? __ldg(x + (((size_t)b * C + ci0 + ci) * N + gy) * N + gx) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int ci = 0; ci < CI; ++ci) {
      const float* wk = &w_s[(ci0 + ci) * 49];
#pragma unroll
      for (int ky = 0; ky < 7; ++ky)
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) acc = fmaf(in_s[ci][ty + ky][tx + kx], wk[ky * 7 + kx], acc);
    }
  }
  const int gy = y0 + ty, gx = x0 + tx;
  if (gy < N && gx < N) y[((size_t)b * N + gy) * N + gx] = acc + __ldg(bias);
}

size_t conv_tiny_workspace_bytes(int B, int N, int mid) {
  size_t act = ((size_t)B * mid * N * N * sizeof(float) + 255) & ~size_t(255);
  size_t wp = ((size_t)mid * mid * 49 * sizeof(float) + 255) & ~size_t(255);
  const size_t ffma = 2 * act + wp;
  const size_t tc = mid == 64 ? conv_tiny_tc_workspace_bytes(B, N) : 0;  // conv_umma.cu
  return ffma > tc ? ffma : tc;
}

template <int C>
static int conv_tiny_impl(const float* x, const PnConvTiny* cv, float* y, int B, int N, Workspace& ws,
                          cudaStream_t st) {
  float* a1 = ws.take<float>((size_t)B * C * N * N);
  float* a2 = ws.take<float>((size_t)B * C * N * N);
  float* wp = ws.take<float>((size_t)C * C * 49);
  PN_REQUIRE(a1 && a2 && wp, PN_ERR_WORKSPACE, "conv_tiny: workspace too small");
  pack_conv2_weights_kernel<<<cdiv(C * C * 49, 256), 256, 0, st>>>(cv->w[1], wp, C);
  PN_TRY(check_launch("pack_conv2_weights_kernel"));
  dim3 g1(cdiv(N, 16), cdiv(N, 16), B);
  conv1_kernel<C><<<g1, 256, 0, st>>>(x, cv->w[0], cv->b[0], a1, N);
  PN_TRY(check_launch("conv1_kernel"));
  const size_t smem2 = (size_t)(C2_CI * 49 * C + C2_CI * 14 * 24) * sizeof(float);
  static bool attr_done[PN_MAX_DEVICES] = {false};  // the attribute is per device
  bool& attr_set = attr_done[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv2_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    PN_REQUIRE(e == cudaSuccess, (int)e, "conv2: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 g2(cdiv(N, 16), cdiv(N, 8), B);
  conv2_kernel<C><<<g2, 32 * (C / 8), smem2, st>>>(a1, wp, cv->b[1], a2, N);
  PN_TRY(check_launch("conv2_kernel"));
  conv3_kernel<C><<<g1, 256, 0, st>>>(a2, cv->w[2], cv->b[2], y, N);
  return check_launch("conv3_kernel");
}

int launch_conv_tiny(const float* x, const PnConvTiny* cv, float* y, int B, int N, void* wsp, size_t ws_bytes,
                     cudaStream_t st) {
  PN_REQUIRE(x && cv && y && B > 0 && N > 0, PN_ERR_BAD_ARG, "conv_tiny: bad args");
  for (int i = 0; i < 3; ++i) PN_REQUIRE(cv->w[i] && cv->b[i], PN_ERR_BAD_ARG, "conv_tiny: null weights");
  Workspace ws(wsp, ws_bytes);
  PN_REQUIRE(wsp, PN_ERR_WORKSPACE, "conv_tiny: null workspace");
  // conv2 (98 % of the flops) as a tcgen05 implicit GEMM (conv_umma.cu); PN_OPT_CONV_TC = 0 keeps the FFMA kernels
  if (cv->mid_channels == 64 && get_option(OPT_TENSOR_CORES) && get_option(OPT_CONV_TC))
    return launch_conv_tiny_tc(x, cv, y, B, N, wsp, ws_bytes, st);
  if (cv->mid_channels == 64) return conv_tiny_impl<64>(x, cv, y, B, N, ws, st);
  if (cv->mid_channels == 16) return conv_tiny_impl<16>(x, cv, y, B, N, ws, st);
  set_error("conv_tiny: mid_channels=%d (supported: 16, 64)", cv->mid_channels);
  return PN_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------ top-k
// Row 7/8 of the hot path (pairnet_head.py:334-351): per image, the K largest entries of the N x N matrix in
// descending order (ties: ascending flat index), idx -> (idx / N, idx % N) as int64, and the fused pair gather.
//
// Throughput design (one CTA per image, several CTAs per SM):
//   1. every thread scans a strided slice of the matrix (128-bit loads) and keeps its local maximum;
//   2. t0 = K-th largest of the THREADS local maxima.  The K largest local maxima are K distinct matrix elements
//      >= t0, so t0 is a lower bound of the K-th largest element: { x >= t0 } contains the whole top-K.  For
//      exchangeable data it holds only ~1.05-1.25 K elements (N = 400..100);
//   3. the candidates are compacted into shared memory (second scan, L1/L2 hits) as 64-bit composites
//      (order-preserving key << 32 | ~index) and ranked by counting -- the composite order IS the output order;
//   4. adversarial inputs (more than CAND_MAX candidates, e.g. a constant matrix) fall back to an exact 4-pass
//      radix select over the whole matrix.
constexpr int TOPK_MAXK = 1024;
constexpr int CAND_MAX = 2048;

__device__ __forceinline__ uint32_t order_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long composite(uint32_t key, uint32_t idx) {
  return ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - idx);
}

// exact radix select (4 passes x 8 bits, MSB first) + ordered tie handling + bitonic sort; result in win[0..K)
template <int THREADS>
__device__ void radix_select_topk(const float* __restrict__ v, int NN, int K, unsigned long long* win, unsigned* hist,
                                  unsigned* warp_cnt, unsigned* scal /* [4] */) {
  const int tid = threadIdx.x;
  unsigned& s_prefix = scal[0];
  unsigned& s_remaining = scal[1];
  unsigned& s_count = scal[2];
  unsigned& s_eq_taken = scal[3];
  if (tid == 0) { s_prefix = 0; s_remaining = (unsigned)K; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += THREADS) hist[i] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < NN; i += THREADS) {
      const uint32_t key = order_key(__ldg(v + i));
      if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned rem = s_remaining;
      int bin = 255;
      for (; bin > 0; --bin) {
        if (hist[bin] >= rem) break;
        rem -= hist[bin];
      }
      s_prefix = prefix | ((unsigned)bin << shift);
      s_remaining = rem;  // how many to take among keys equal to the (growing) prefix
    }
    __syncthreads();
  }
  const uint32_t T = s_prefix;            // key of the K-th largest element
  const unsigned need_eq = s_remaining;   // how many elements with key == T belong to the top-K
  if (tid == 0) { s_count = 0; s_eq_taken = 0; }
  __syncthreads();
  // collect: keys > T in any order (sorted below), keys == T by ascending index
  for (int base = 0; base < NN; base += THREADS) {
    const int i = base + tid;
    uint32_t key = 0;
    bool gt = false, eq = false;
    if (i < NN) {
      key = order_key(__ldg(v + i));
      gt = key > T;
      eq = key == T;
    }
    if (gt) win[atomicAdd(&s_count, 1u)] = composite(key, (uint32_t)i);
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    const int lane = tid & 31, wid = tid >> 5;
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    unsigned before = s_eq_taken;
    for (int w2 = 0; w2 < wid; ++w2) before += warp_cnt[w2];
    const unsigned rank = before + __popc(bal & ((1u << lane) - 1u));
    if (eq && rank < need_eq) win[atomicAdd(&s_count, 1u)] = composite(key, (uint32_t)i);
    __syncthreads();
    if (tid == 0) {
      unsigned tot = 0;
      for (int w2 = 0; w2 < THREADS / 32; ++w2) tot += warp_cnt[w2];
      s_eq_taken += tot;
    }
    __syncthreads();
  }
  // bitonic sort (descending) of the K winners, padded with 0 (smallest)
  int P = 1;
  while (P < K) P <<= 1;
  for (int i = K + tid; i < P; i += THREADS) win[i] = 0ull;
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < P / 2; i += THREADS) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = win[lo], c = win[hi];
        if ((a < c) == desc) { win[lo] = c; win[hi] = a; }
      }
      __syncthreads();
    }
  }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) topk_pairs_kernel(const float* __restrict__ imp,
                                                             int64_t* __restrict__ topk_idx,
                                                             int64_t* __restrict__ sub_pos,
                                                             int64_t* __restrict__ obj_pos,
                                                             const float* __restrict__ query,
                                                             float* __restrict__ pair_feat, int N, int K,
                                                             int force_radix,
                                                             const int* __restrict__ only_if) {
  __shared__ uint32_t maxk[THREADS];
  __shared__ unsigned long long cand[CAND_MAX];
  __shared__ unsigned long long win[TOPK_MAXK];
  __shared__ unsigned hist[256];
  __shared__ unsigned warp_cnt[32];
  __shared__ unsigned scal[4];
  __shared__ unsigned s_t0, s_ncand;
  const int b = blockIdx.x;
  if (only_if && !only_if[b]) return;  // clean-up pass after the fused pair-matrix + top-k kernel (pair_topk.cu)
  const int tid = threadIdx.x, lane = tid & 31;
  const int NN = N * N;
  const float* v = imp + (size_t)b * NN;
  const bool vec = (NN & 3) == 0 && ((reinterpret_cast<uintptr_t>(imp) & 15) == 0);
  const int W = vec ? 4 : 1;  // elements per thread per sweep step

  // ---- 1. local maxima over a strided slice
  uint32_t lmax = 0;
  if (vec) {
    for (int i = tid * 4; i < NN; i += THREADS * 4) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(v + i));
      lmax = max(max(lmax, order_key(x.x)), max(order_key(x.y), max(order_key(x.z), order_key(x.w))));
    }
  } else {
    for (int i = tid; i < NN; i += THREADS) lmax = max(lmax, order_key(__ldg(v + i)));
  }
  maxk[tid] = lmax;
  if (tid == 0) { s_t0 = 0; s_ncand = 0; }
  __syncthreads();
  // ---- 2. t0 = K-th largest local maximum (rank by counting; ties broken by thread id)
  if (K <= THREADS) {
    unsigned cnt = 0;
    for (int j = 0; j < THREADS; j += 4) {
      const uint4 m = *reinterpret_cast<const uint4*>(&maxk[j]);
      cnt += (m.x > lmax || (m.x == lmax && j < tid)) + (m.y > lmax || (m.y == lmax && j + 1 < tid)) +
             (m.z > lmax || (m.z == lmax && j + 2 < tid)) + (m.w > lmax || (m.w == lmax && j + 3 < tid));
    }
    if (cnt == (unsigned)(K - 1)) s_t0 = lmax;
  }
  __syncthreads();
  const uint32_t t0 = s_t0;
  // ---- 3. compact { key >= t0 } (warp-aggregated slots; order is irrelevant, the composite carries it)
  if (!force_radix) {
    for (int i0 = 0; i0 < NN; i0 += THREADS * W) {
      const int i = i0 + tid * W;
      uint32_t key[4] = {0, 0, 0, 0};
      bool ok[4] = {false, false, false, false};
      if (vec) {
        if (i < NN) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(v + i));
          key[0] = order_key(x.x); key[1] = order_key(x.y); key[2] = order_key(x.z); key[3] = order_key(x.w);
#pragma unroll
          for (int u = 0; u < 4; ++u) ok[u] = key[u] >= t0;
        }
      } else if (i < NN) {
        key[0] = order_key(__ldg(v + i));
        ok[0] = key[0] >= t0;
      }
      const int mine = (int)ok[0] + (int)ok[1] + (int)ok[2] + (int)ok[3];
      // warp-inclusive scan of the per-lane counts -> one atomic per warp per step
      int incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      unsigned base = 0;
      if (total > 0) {
        if (lane == 31) base = atomicAdd(&s_ncand, (unsigned)total);
        base = __shfl_sync(0xffffffffu, base, 31);
        unsigned slot = base + (unsigned)(incl - mine);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (ok[u]) {
            if (slot < CAND_MAX) cand[slot] = composite(key[u], (uint32_t)(i + u));
            ++slot;
          }
      }
    }
  }
  __syncthreads();
  const unsigned nc = s_ncand;
  if (force_radix || nc > CAND_MAX || nc < (unsigned)K) {
    radix_select_topk<THREADS>(v, NN, K, win, hist, warp_cnt, scal);   // ends with a barrier
  } else {
    // ---- 4. rank by counting: win[r] = candidate with r larger candidates
    for (unsigned i = tid; i < nc; i += THREADS) {
      const unsigned long long me = cand[i];
      unsigned r = 0;
      for (unsigned j = 0; j < nc; ++j) r += cand[j] > me;
      if (r < (unsigned)K) win[r] = me;
    }
    __syncthreads();
  }
  // ---- emit indices (pairnet_head.py:337-340)
  for (int r = tid; r < K; r += THREADS) {
    const long long idx = (long long)(0xffffffffu - (uint32_t)(win[r] & 0xffffffffull));
    if (topk_idx) topk_idx[(size_t)b * K + r] = idx;
    sub_pos[(size_t)b * K + r] = idx / N;
    obj_pos[(size_t)b * K + r] = idx % N;
  }
  // ---- fused pair gather: pair_feat[b] = [query[b, sub_pos] ; query[b, obj_pos]]   ([2K,256])
  if (pair_feat) {
    const float4* q4 = reinterpret_cast<const float4*>(query + (size_t)b * N * D);
    float4* p4 = reinterpret_cast<float4*>(pair_feat + (size_t)b * 2 * K * D);
    for (int i = tid; i < 2 * K * (D / 4); i += THREADS) {
      const int r = i / (D / 4), c = i % (D / 4);
      const uint32_t idx = 0xffffffffu - (uint32_t)(win[r < K ? r : r - K] & 0xffffffffull);
      const int row = (r < K) ? (int)(idx / (uint32_t)N) : (int)(idx % (uint32_t)N);
      p4[(size_t)r * (D / 4) + c] = __ldg(q4 + (size_t)row * (D / 4) + c);
    }
  }
}

int launch_topk_pairs(const float* imp, int64_t* topk_idx, int64_t* sub_pos, int64_t* obj_pos, const float* query,
                      float* pair_feat, int B, int N, int K, cudaStream_t st, const int* only_if) {
  PN_REQUIRE(imp && sub_pos && obj_pos && B > 0 && N > 0, PN_ERR_BAD_ARG, "topk_pairs: bad args");
  PN_REQUIRE(K >= 1 && K <= TOPK_MAXK && (long long)K <= (long long)N * N, PN_ERR_UNSUPPORTED,
             "topk_pairs: K=%d unsupported (1..%d, <= N*N)", K, TOPK_MAXK);
  PN_REQUIRE(!pair_feat || query, PN_ERR_BAD_ARG, "topk_pairs: pair_feat needs query");
  const long long NN = (long long)N * N;
  const int force_radix = get_option(OPT_TOPK_RADIX) != 0;
  if (NN <= 16384)
    topk_pairs_kernel<256><<<B, 256, 0, st>>>(imp, topk_idx, sub_pos, obj_pos, query, pair_feat, N, K, force_radix,
                                       only_if);
  else if (NN <= 65536)
    topk_pairs_kernel<512><<<B, 512, 0, st>>>(imp, topk_idx, sub_pos, obj_pos, query, pair_feat, N, K, force_radix,
                                       only_if);
  else
    topk_pairs_kernel<1024><<<B, 1024, 0, st>>>(imp, topk_idx, sub_pos, obj_pos, query, pair_feat, N, K, force_radix,
                                       only_if);
  return check_launch("topk_pairs_kernel");
}

}  // namespace pn
