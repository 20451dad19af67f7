// Upstream plumbing: GroupNorm(32, 256) (+ optional ReLU) of the pixel decoder's ConvModules
// (mmdet MSDeformAttnPixelDecoder input/lateral/output convs).  PyTorch's channels_last GroupNorm takes
// ~0.75 ms on the [2,256,200,334] FPN map; this two-pass version is HBM-bound (read twice, write once).
// Supports NCHW-contiguous and channels_last (NHWC) storage; output uses the same storage as the input.
#include "common.cuh"

namespace pn {

constexpr int GN_THREADS = 256;

// ---- pass 1: per-CTA partial (sum, sumsq) per (b, group), fp32 per thread -> double per CTA ---------------
// NHWC: x[b][p][c]; CTA covers `ppc` pixels; thread t: channel quad (t % 64) * 4, pixel lane t / 64.
__global__ void __launch_bounds__(GN_THREADS) gn_stats_nhwc_kernel(const float* __restrict__ x, double* __restrict__ part,
                                                                    int HW, int ppc, int groups, int nchunk) {
  __shared__ float s_sum[GN_THREADS], s_sq[GN_THREADS];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int t = threadIdx.x, cq = t & 63, pl = t >> 6;
  const int p0 = chunk * ppc, p1 = min(HW, p0 + ppc);
  float s = 0.f, q = 0.f;
  const float4* xb = reinterpret_cast<const float4*>(x + (size_t)b * HW * D);
  for (int p = p0 + pl; p < p1; p += 4) {
    const float4 v = __ldg(xb + (size_t)p * (D / 4) + cq);
    s += (v.x + v.y) + (v.z + v.w);
    q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  s_sum[t] = s; s_sq[t] = q;
  __syncthreads();
  // group g owns channels [g*cpg, (g+1)*cpg): channel quads cq with cq*4/cpg == g
  const int cpg = D / groups;
  if (t < groups) {
    double ds = 0.0, dq = 0.0;
    const int q0 = t * cpg / 4, q1 = (t + 1) * cpg / 4;
    for (int c = q0; c < q1; ++c)
      for (int l = 0; l < 4; ++l) { ds += (double)s_sum[l * 64 + c]; dq += (double)s_sq[l * 64 + c]; }
    part[(((size_t)b * groups + t) * nchunk + chunk) * 2 + 0] = ds;
    part[(((size_t)b * groups + t) * nchunk + chunk) * 2 + 1] = dq;
  }
}
// NCHW: x[b][c][p]; one CTA per (chunk, b*groups+g): the group's cpg*HW elements are contiguous.
__global__ void __launch_bounds__(GN_THREADS) gn_stats_nchw_kernel(const float* __restrict__ x, double* __restrict__ part,
                                                                    long long n_per_group, int nchunk) {
  __shared__ double s_sum[GN_THREADS / 32], s_sq[GN_THREADS / 32];
  const int bg = blockIdx.y, chunk = blockIdx.x;
  const long long per = (n_per_group + nchunk - 1) / nchunk;
  const long long i0 = chunk * per, i1 = min(n_per_group, i0 + per);
  const float* xg = x + (size_t)bg * n_per_group;
  float s = 0.f, q = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += GN_THREADS) {
    const float v = __ldg(xg + i);
    s += v; q += v * v;
  }
  double ds = s, dq = q;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { ds += __shfl_xor_sync(0xffffffffu, ds, o); dq += __shfl_xor_sync(0xffffffffu, dq, o); }
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = ds; s_sq[threadIdx.x >> 5] = dq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < GN_THREADS / 32; ++w) { ds += s_sum[w]; dq += s_sq[w]; }
    part[((size_t)bg * nchunk + chunk) * 2 + 0] = ds;
    part[((size_t)bg * nchunk + chunk) * 2 + 1] = dq;
  }
}
// ---- pass 1b: reduce partials -> per (b, c) scale/shift:  y = x * scale + shift -----------------------------
__global__ void gn_finalize_kernel(const double* __restrict__ part, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ scale_shift, int groups, int nchunk,
                                   double count, float eps) {
  const int b = blockIdx.x, c = threadIdx.x;  // 256 threads
  const int cpg = D / groups, g = c / cpg;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < nchunk; ++k) {
    s += part[(((size_t)b * groups + g) * nchunk + k) * 2 + 0];
    q += part[(((size_t)b * groups + g) * nchunk + k) * 2 + 1];
  }
  const double mean = s / count;
  double var = q / count - mean * mean;
  var = var < 0.0 ? 0.0 : var;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = rstd * __ldg(gamma + c);
  scale_shift[((size_t)b * D + c) * 2 + 0] = sc;
  scale_shift[((size_t)b * D + c) * 2 + 1] = __ldg(beta + c) - (float)mean * sc;
}
// ---- pass 2: apply --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_apply_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ ss,
                                                             float* __restrict__ y, long long n4, int HW, int relu) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;  // float4 index over [B,HW,64]
  if (i >= n4) return;
  const int cq = (int)(i & 63);
  const int b = (int)(i / ((long long)HW * 64));
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(ss) + ((size_t)b * D + cq * 4) / 2);
  const float4 s1 = __ldg(reinterpret_cast<const float4*>(ss) + ((size_t)b * D + cq * 4) / 2 + 1);
  float4 o;
  o.x = fmaf(v.x, s0.x, s0.y); o.y = fmaf(v.y, s0.z, s0.w); o.z = fmaf(v.z, s1.x, s1.y); o.w = fmaf(v.w, s1.z, s1.w);
  if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
  reinterpret_cast<float4*>(y)[i] = o;
}
__global__ void __launch_bounds__(256) gn_apply_nchw_kernel(const float* __restrict__ x, const float* __restrict__ ss,
                                                             float* __restrict__ y, int HW, int relu) {
  const int bc = blockIdx.y;  // b*256 + c
  const float sc = __ldg(ss + (size_t)bc * 2), sh = __ldg(ss + (size_t)bc * 2 + 1);
  const float* xp = x + (size_t)bc * HW;
  float* yp = y + (size_t)bc * HW;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
    float o = fmaf(__ldg(xp + p), sc, sh);
    yp[p] = relu ? fmaxf(o, 0.f) : o;
  }
}

}  // namespace pn

using namespace pn;

extern "C" {

size_t pn_group_norm_workspace_bytes(int B, int HW, int groups) {
  const int nchunk = 64;
  return (size_t)B * groups * nchunk * 2 * sizeof(double) + (size_t)B * D * 2 * sizeof(float) + 512;
}

int pn_group_norm(const float* x, const float* gamma, const float* beta, float* y, int B, int HW, int groups, int relu,
                  int channels_last, float eps, void* wsp, size_t ws_bytes, pn_stream_t stream) {
  PN_REQUIRE(x && gamma && beta && y && wsp, PN_ERR_BAD_ARG, "group_norm: null argument");
  PN_REQUIRE(groups > 0 && groups <= 64 && D % groups == 0 && (D / groups) % 4 == 0, PN_ERR_UNSUPPORTED,
             "group_norm: 256 channels, groups must divide 64");
  PN_REQUIRE(ws_bytes >= pn_group_norm_workspace_bytes(B, HW, groups), PN_ERR_WORKSPACE, "group_norm: workspace too small");
  PN_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, PN_ERR_UNSUPPORTED, "group_norm: 16B alignment");
  cudaStream_t st = as_stream(stream);
  const int nchunk = 64;
  double* part = reinterpret_cast<double*>(wsp);
  float* ss = reinterpret_cast<float*>(part + (size_t)B * groups * nchunk * 2);
  const double count = (double)HW * (D / groups);
  if (channels_last) {
    const int ppc = cdiv(HW, nchunk);
    gn_stats_nhwc_kernel<<<dim3(nchunk, B), GN_THREADS, 0, st>>>(x, part, HW, ppc, groups, nchunk);
    PN_TRY(check_launch("gn_stats_nhwc_kernel"));
  } else {
    gn_stats_nchw_kernel<<<dim3(nchunk, B * groups), GN_THREADS, 0, st>>>(x, part, (long long)HW * (D / groups), nchunk);
    PN_TRY(check_launch("gn_stats_nchw_kernel"));
  }
  gn_finalize_kernel<<<B, D, 0, st>>>(part, gamma, beta, ss, groups, nchunk, count, eps);
  PN_TRY(check_launch("gn_finalize_kernel"));
  if (channels_last) {
    const long long n4 = (long long)B * HW * (D / 4);
    gn_apply_nhwc_kernel<<<cdiv(n4, 256), 256, 0, st>>>(x, ss, y, n4, HW, relu);
    return check_launch("gn_apply_nhwc_kernel");
  }
  int gx = cdiv(HW, 256 * 4);
  gx = gx < 1 ? 1 : gx;
  gn_apply_nchw_kernel<<<dim3(gx, B * D), 256, 0, st>>>(x, ss, y, HW, relu);
  return check_launch("gn_apply_nchw_kernel");
}

}  // extern "C"
