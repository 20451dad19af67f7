// ConvTiny "Matrix Learner" filter (cnn_factory.py:6-53; call pairnet_head.py:333) with conv2 -- 64 -> 64 channels, 7x7,
// 98 % of the filter's flops -- as an implicit GEMM on tcgen05:
//
//   per image:  M = N*N pixels,  N_mma = 64 output channels,  K = 49 taps x 64 input channels = 3136
//
// Layout: activations are channels-last [B, N, N, 64] fp32 so that a pixel is one 256-byte row = two 128-byte swizzle
// rows (k-blocks of 32 channels).  3xTF32 with the truncation split of pair_topk.cu: the tensor pipe ignores the low 13
// mantissa bits, so the raw tensor IS the hi operand; conv1 writes the lo plane (x - trunc(x), nudged by half a TF32
// ulp) next to the raw one, the weights' lo plane is packed once per call.
//
// One CTA = one 8 (x) x 16 (y) patch of output pixels = the 128 rows of the MMA.  For a tap row dy and a k-block the TMA
// unit loads ONE slab [16 y][16 x][32 ch] (raw + lo, 64 KB) with a 4-D box whose out-of-image elements are zero-filled
// (= the convolution's zero padding; coordinates may be negative): slab row y*16 + x.  The seven taps dx = 0..6 of that
// row are SHIFTED WINDOWS of the same slab: the A descriptor starts dx rows (dx * 128 B) into the slab, 8-row groups
// (8 consecutive x of one y) 2048 B apart (the swizzle phase follows the absolute smem address, see make_desc).  So the activations cross L2 -> SM 7x less often than with one tile per tap, and
// never get re-laid-out (no im2col buffer).  Weight tiles [64 oc][32 ic] (hi + lo, 16 KB per tap and k-block) stream
// through a 4-stage ring.
//
//   warp 0   TMA producer (slabs + weight tiles)     warp 1   MMA issuer: per (dy, kb, dx) 4 k-steps x (lo*hi + hi*lo +
//   warps 2-5 epilogue: fp32 sum of the per-unit partial          hi*hi) into one of 8 TMEM accumulators (64 columns) per
//            accumulators, + bias, ReLU, channels-last store      (tap row, k-block) unit: short accumulation chains
//
// conv1 (1 -> 64) and conv3 (64 -> 1) are small FFMA kernels on the same channels-last layout.
#include "umma_ptx.cuh"

namespace pn {
namespace convtc {

using namespace umma;

constexpr int C = 64;                    // mid channels
constexpr int PW = 8, PH = 16;           // output patch: 8 x * 16 y = 128 pixels
constexpr int SLAB_ROWS = 16 * 16;       // [16 y][16 x] pixels of one tap row (x0-3 .. x0+12)
constexpr int SLAB_BYTES = SLAB_ROWS * 128;   // 32 KB (one k-block of 32 channels)
constexpr int SLAB_STAGE = 2 * SLAB_BYTES;    // raw + lo
constexpr int SLAB_STAGES = 2;
constexpr int WT_BYTES = C * 128;             // [64 oc][32 ic] = 8 KB
constexpr int WT_STAGE = 2 * WT_BYTES;        // hi + lo
constexpr int WT_STAGES = 4;
constexpr int NUM_THREADS = 192;
constexpr int ACC_SLOTS = 8;                   // TMEM accumulators of 64 columns: one per (tap row, k-block) chain in flight
constexpr int TMEM_COLS = ACC_SLOTS * C;
constexpr size_t SMEM_BYTES = (size_t)SLAB_STAGES * SLAB_STAGE + (size_t)WT_STAGES * WT_STAGE + 1024 + 256;

struct Params {
  CUtensorMap a_hi, a_lo;  // [B][N][N][64] fp32, box 32 ch x 16 x x 16 y x 1
  CUtensorMap w_hi, w_lo;  // [(tap*2 + kb)*64 + oc][32 ic], box 32 x 64
  const float* bias;       // [64]
  float* out;              // [B][N][N][64] = relu(conv + bias)
  int B, N, tiles_x, tiles_y, total_tiles;
};

__device__ __forceinline__ float lo_of(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  return __uint_as_float(__float_as_uint(x - hi) + 0x1000u);
}
// K-major SWIZZLE_128B operand: 8-row groups `sbo` bytes apart.  The start address may be ANY multiple of 128 bytes: the
// tensor pipe derives the swizzle phase of a row from its absolute shared-memory address (bits 7-9), exactly as the TMA
// unit did when it wrote the slab, so a window shifted by dx rows needs no base offset (measured on B200: base offset 0
// is exact for all 49 taps, base offset = dx or 8 - dx is wrong for every dx != 0).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1) conv2_umma_kernel(const __grid_constant__ Params prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* slabs = smem;
  uint8_t* wts = smem + SLAB_STAGES * SLAB_STAGE;
  uint64_t* slab_full = reinterpret_cast<uint64_t*>(wts + WT_STAGES * WT_STAGE);
  uint64_t* slab_empty = slab_full + SLAB_STAGES;
  uint64_t* w_full = slab_empty + SLAB_STAGES;
  uint64_t* w_empty = w_full + WT_STAGES;
  uint64_t* tmem_full = w_empty + WT_STAGES;      // [ACC_SLOTS] MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + ACC_SLOTS;   // [ACC_SLOTS] epilogue -> MMA
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC_SLOTS);

  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < SLAB_STAGES; ++s) { mbar_init(&slab_full[s], 1); mbar_init(&slab_empty[s], 1); }
    for (int s = 0; s < WT_STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int a = 0; a < ACC_SLOTS; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_base_slot);
  const int per_img = prm.tiles_x * prm.tiles_y;

  if (warp == 0) {
    // ===== TMA producer
    uint32_t su = 0, wu = 0;  // slab / weight-tile sequence numbers
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
      const int b = t / per_img, r = t - b * per_img;
      const int x0 = (r % prm.tiles_x) * PW, y0 = (r / prm.tiles_x) * PH;
      for (int u = 0; u < 14; ++u, ++su) {
        const int dy = u >> 1, kb = u & 1;
        const int ss = su % SLAB_STAGES;
        mbar_wait(&slab_empty[ss], ((su / SLAB_STAGES) & 1) ^ 1);
        uint8_t* sl = slabs + (size_t)ss * SLAB_STAGE;
        if (elect_one()) {
          mbar_expect_tx(&slab_full[ss], (uint32_t)SLAB_STAGE);
          tma_load_4d(sl, &prm.a_hi, &slab_full[ss], kb * 32, x0 - 3, y0 - 3 + dy, b);
          tma_load_4d(sl + SLAB_BYTES, &prm.a_lo, &slab_full[ss], kb * 32, x0 - 3, y0 - 3 + dy, b);
        }
        __syncwarp();
        for (int dx = 0; dx < 7; ++dx, ++wu) {
          const int ws = wu % WT_STAGES;
          mbar_wait(&w_empty[ws], ((wu / WT_STAGES) & 1) ^ 1);
          uint8_t* wt = wts + (size_t)ws * WT_STAGE;
          if (elect_one()) {
            const int row = ((dy * 7 + dx) * 2 + kb) * C;
            mbar_expect_tx(&w_full[ws], (uint32_t)WT_STAGE);
            tma_load_2d(wt, &prm.w_hi, &w_full[ws], 0, row);
            tma_load_2d(wt + WT_BYTES, &prm.w_lo, &w_full[ws], 0, row);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer.  The tensor pipe accumulates with truncation: one chain over all 1176 MMAs of a tile drifts to
    // 2e-5 of the output scale (measured; the FFMA path is at 1.6e-6).  So every (tap row, k-block) unit -- 84 MMAs --
    // accumulates into its OWN tensor-memory slot, and the epilogue warps add the 14 partial sums in fp32 (round to
    // nearest) while later units are still running.
    const uint32_t idesc = make_idesc(C);
    uint32_t su = 0, wu = 0;
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
      for (int u = 0; u < 14; ++u, ++su) {
        const int ss = su % SLAB_STAGES;
        const uint32_t slot = su % ACC_SLOTS;
        mbar_wait(&tmem_empty[slot], ((su / ACC_SLOTS) & 1) ^ 1);
        mbar_wait(&slab_full[ss], (su / SLAB_STAGES) & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + slot * C;
        const uint32_t sl = smem_u32(slabs + (size_t)ss * SLAB_STAGE);
        for (int dx = 0; dx < 7; ++dx, ++wu) {
          const int ws = wu % WT_STAGES;
          mbar_wait(&w_full[ws], (wu / WT_STAGES) & 1);
          tc_fence_after();
          const uint32_t wt = smem_u32(wts + (size_t)ws * WT_STAGE);
          // window of the slab shifted by dx pixels: rows y*16 + x + dx
          const uint64_t a_hi = make_desc(sl + dx * 128, 2048), a_lo = make_desc(sl + SLAB_BYTES + dx * 128, 2048);
          const uint64_t b_hi = make_desc(wt, 1024), b_lo = make_desc(wt + WT_BYTES, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
              const uint32_t first = (dx == 0 && k == 0) ? 0u : 1u;
              umma_tf32(d_tmem, a_lo + koff, b_hi + koff, idesc, first);
              umma_tf32(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
              umma_tf32(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
            }
            umma_commit(&w_empty[ws]);
            if (dx == 6) {
              umma_commit(&slab_empty[ss]);
              umma_commit(&tmem_full[slot]);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue: TMEM lane quadrant = warp % 4; thread = output pixel (y_l = row / 8, x_l = row % 8).  Drains the 14
    // partial accumulators of a tile as they complete and sums them in registers.
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    uint32_t su = 0;
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
      const int b = t / per_img, r = t - b * per_img;
      const int x = (r % prm.tiles_x) * PW + (m & 7), y = (r / prm.tiles_x) * PH + (m >> 3);
      float sum[C];
#pragma unroll
      for (int j = 0; j < C; ++j) sum[j] = 0.f;
#pragma unroll 1
      for (int u = 0; u < 14; ++u, ++su) {
        const uint32_t slot = su % ACC_SLOTS;
        mbar_wait(&tmem_full[slot], (su / ACC_SLOTS) & 1);
        tc_fence_after();
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + slot * C;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(tbase, v0);
        tmem_ld_32x32b_x32(tbase + 32, v1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[slot]);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sum[j] += __uint_as_float(v0[j]);
          sum[32 + j] += __uint_as_float(v1[j]);
        }
      }
      if (x < prm.N && y < prm.N) {
        float4* dst = reinterpret_cast<float4*>(prm.out + (((size_t)b * prm.N + y) * prm.N + x) * C);
        const float4* b4 = reinterpret_cast<const float4*>(prm.bias);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 bb = __ldg(b4 + j);
          dst[j] = make_float4(fmaxf(sum[4 * j] + bb.x, 0.f), fmaxf(sum[4 * j + 1] + bb.y, 0.f),
                               fmaxf(sum[4 * j + 2] + bb.z, 0.f), fmaxf(sum[4 * j + 3] + bb.w, 0.f));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ conv1 (1 -> 64), channels-last out
// x [B,N,N] -> h1 [B,N,N,64] = relu(conv7x7(x) + b) and its TF32 lo plane.  Thread = (pixel, 4 channels); the sum runs
// bias first, then the taps in (ky, kx) order -- the order of the FFMA path (ppn.cu conv1_kernel): bit-identical h1.
__global__ void __launch_bounds__(256) conv1_cl_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ h1,
                                                        float* __restrict__ h1_lo, int N, int total_px) {
  __shared__ __align__(16) float w_s[49][C];
  for (int i = threadIdx.x; i < 49 * C; i += 256) w_s[i / C][i % C] = __ldg(w + (i % C) * 49 + i / C);
  __syncthreads();
  const int p = blockIdx.x * 16 + (threadIdx.x >> 4);
  const int cq = (threadIdx.x & 15) * 4;
  if (p >= total_px) return;
  const int b = p / (N * N), r = p - b * N * N, py = r / N, px = r - py * N;
  const float* src = x + (size_t)b * N * N;
  float4 acc = __ldg(reinterpret_cast<const float4*>(bias + cq));
#pragma unroll
  for (int ky = 0; ky < 7; ++ky) {
    const int gy = py + ky - 3;
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      const int gx = px + kx - 3;
      const float v = (gy >= 0 && gy < N && gx >= 0 && gx < N) ? __ldg(src + (size_t)gy * N + gx) : 0.f;
      const float4 ww = *reinterpret_cast<const float4*>(&w_s[ky * 7 + kx][cq]);
      acc.x = fmaf(v, ww.x, acc.x); acc.y = fmaf(v, ww.y, acc.y); acc.z = fmaf(v, ww.z, acc.z); acc.w = fmaf(v, ww.w, acc.w);
    }
  }
  acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
  *reinterpret_cast<float4*>(h1 + (size_t)p * C + cq) = acc;
  *reinterpret_cast<float4*>(h1_lo + (size_t)p * C + cq) = make_float4(lo_of(acc.x), lo_of(acc.y), lo_of(acc.z), lo_of(acc.w));
}

// w2 [oc][ic][7][7] (torch) -> hi / lo tiles [(tap*2 + kb)*64 + oc][32 ic_l], ic = kb*32 + ic_l
__global__ void pack_conv2_tc_weights_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= 49 * 2 * C * 32) return;
  const int icl = i & 31, oc = (i >> 5) & 63, kb = (i >> 11) & 1, tap = i >> 12;
  const float v = __ldg(w + ((size_t)oc * C + kb * 32 + icl) * 49 + tap);
  hi[i] = v;
  lo[i] = lo_of(v);
}

// ------------------------------------------------------------------------------------------ conv3 (64 -> 1), channels-last in
// h2 [B,N,N,64] -> y [B,N,N] = conv7x7 + b.  One warp = 8 consecutive outputs of a row; lane = channel pair (a warp reads
// one pixel = 256 contiguous bytes per load); the 14-pixel input window of a tap row is read once and feeds up to 7
// outputs each; 8 accumulators per lane, reduced over the lanes at the end.
__global__ void __launch_bounds__(256) conv3_cl_kernel(const float* __restrict__ h2, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y, int B,
                                                        int N) {
  __shared__ __align__(8) float w_s[49][C];
  for (int i = threadIdx.x; i < 49 * C; i += 256) w_s[i / C][i % C] = __ldg(w + (i % C) * 49 + i / C);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int segs = (N + 7) / 8;
  const int task = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (task >= B * N * segs) return;
  const int b = task / (N * segs), r = task - b * N * segs, py = r / segs, x0 = (r - py * segs) * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 1
  for (int ky = 0; ky < 7; ++ky) {
    const int gy = py + ky - 3;
    if (gy < 0 || gy >= N) continue;  // warp-uniform
    float2 wk[7];
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) wk[kx] = *reinterpret_cast<const float2*>(&w_s[ky * 7 + kx][2 * lane]);
    const float* rowp = h2 + (((size_t)b * N + gy) * N) * C + 2 * lane;
#pragma unroll
    for (int xs = 0; xs < 14; ++xs) {
      const int gx = x0 + xs - 3;
      float2 v = make_float2(0.f, 0.f);
      if (gx >= 0 && gx < N) v = __ldg(reinterpret_cast<const float2*>(rowp + (size_t)gx * C));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kx = xs - j;
        if (kx >= 0 && kx < 7) acc[j] = fmaf(v.y, wk[kx].y, fmaf(v.x, wk[kx].x, acc[j]));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a = acc[j];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    acc[j] = a;
  }
  if (lane < 8 && x0 + lane < N) {
    float a = acc[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) a = lane == j ? acc[j] : a;
    y[((size_t)b * N + py) * N + x0 + lane] = a + __ldg(bias);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map_act(CUtensorMap* map, const float* ptr, int B, int N) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  PN_REQUIRE(fn, PN_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)N * C * 4, (cuuint64_t)N * N * C * 4};
  cuuint32_t box[4] = {32, 16, 16, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PN_REQUIRE(r == CUDA_SUCCESS, PN_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled(4d) failed (%d)", (int)r);
  return 0;
}

}  // namespace convtc

size_t conv_tiny_tc_workspace_bytes(int B, int N) {
  const size_t act = ((size_t)B * convtc::C * N * N * sizeof(float) + 1023) & ~size_t(1023);
  const size_t wp = ((size_t)49 * 2 * convtc::C * 32 * sizeof(float) + 1023) & ~size_t(1023);
  return 3 * act + 2 * wp + 1024;
}

// y = conv3(relu(conv2(relu(conv1(x)))))  with conv2 on tcgen05 (mid_channels = 64)
int launch_conv_tiny_tc(const float* x, const PnConvTiny* cv, float* y, int B, int N, void* wsp, size_t ws_bytes,
                        cudaStream_t st) {
  using namespace convtc;
  PN_REQUIRE(x && cv && y && B > 0 && N > 0 && cv->mid_channels == C, PN_ERR_BAD_ARG, "conv_tiny_tc: bad args");
  Workspace ws(wsp, ws_bytes);
  const size_t act = (size_t)B * C * N * N;
  float* h1 = ws.take<float>(act);
  float* h1_lo = ws.take<float>(act);
  float* h2 = ws.take<float>(act);
  float* w_hi = ws.take<float>((size_t)49 * 2 * C * 32);
  float* w_lo = ws.take<float>((size_t)49 * 2 * C * 32);
  PN_REQUIRE(wsp && ws.ok() && h1 && h1_lo && h2 && w_hi && w_lo, PN_ERR_WORKSPACE, "conv_tiny_tc: workspace too small");
  pack_conv2_tc_weights_kernel<<<cdiv(49 * 2 * C * 32, 256), 256, 0, st>>>(cv->w[1], w_hi, w_lo);
  PN_TRY(check_launch("pack_conv2_tc_weights_kernel"));
  const int total_px = B * N * N;
  conv1_cl_kernel<<<cdiv(total_px, 16), 256, 0, st>>>(x, cv->w[0], cv->b[0], h1, h1_lo, N, total_px);
  PN_TRY(check_launch("conv1_cl_kernel"));
  Params prm{};
  PN_TRY(make_map_act(&prm.a_hi, h1, B, N));
  PN_TRY(make_map_act(&prm.a_lo, h1_lo, B, N));
  PN_TRY(make_tmap_2d(&prm.w_hi, w_hi, 49 * 2 * C, 32, 32, 32, C));
  PN_TRY(make_tmap_2d(&prm.w_lo, w_lo, 49 * 2 * C, 32, 32, 32, C));
  prm.bias = cv->b[1];
  prm.out = h2;
  prm.B = B; prm.N = N;
  prm.tiles_x = cdiv(N, PW); prm.tiles_y = cdiv(N, PH);
  prm.total_tiles = B * prm.tiles_x * prm.tiles_y;
  static bool attr_done[PN_MAX_DEVICES] = {false};
  bool& attr_set = attr_done[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv2_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    PN_REQUIRE(e == cudaSuccess, (int)e, "conv2_umma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int num_sms = sm_count();
  const int grid = prm.total_tiles < num_sms ? prm.total_tiles : num_sms;
  conv2_umma_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(prm);
  PN_TRY(check_launch("conv2_umma_kernel"));
  const int segs = cdiv(N, 8);
  conv3_cl_kernel<<<cdiv(B * N * segs, 8), 256, 0, st>>>(h2, cv->w[2], cv->b[2], y, B, N);
  return check_launch("conv3_cl_kernel");
}

}  // namespace pn
