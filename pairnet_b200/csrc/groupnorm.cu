// Upstream plumbing: GroupNorm(32, 256) (+ optional ReLU) of the pixel decoder's ConvModules
// (mmdet MSDeformAttnPixelDecoder input/lateral/output convs).  PyTorch's channels_last GroupNorm takes
// ~0.75 ms on the [2,256,200,334] FPN map; this two-pass version is HBM-bound (read twice, write once).
// Supports NCHW-contiguous and channels_last (NHWC) storage; output uses the same storage as the input.
#include "common.cuh"

namespace pn {

constexpr int GN_THREADS = 256;
constexpr int GN_NCHUNK = 148;  // statistics CTAs per image (x B images: >= one wave of the 148 SMs)

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
// one warp per (image, group): lanes stride over the per-CTA partials, shuffle-reduce in double, then the group's
// channels (cpg <= 32) get their scale / shift
__global__ void __launch_bounds__(256) gn_finalize_kernel(const double* __restrict__ part, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ scale_shift,
                                                          int B, int groups, int nchunk, double count, float eps) {
  const int wid = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wid >= B * groups) return;
  const int b = wid / groups, g = wid - b * groups;
  const double* p = part + (size_t)wid * nchunk * 2;
  double s = 0.0, q = 0.0;
  for (int k = lane; k < nchunk; k += 32) { s += p[2 * k]; q += p[2 * k + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  const double mean = s / count;
  double var = q / count - mean * mean;
  var = var < 0.0 ? 0.0 : var;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const int cpg = D / groups;
  for (int j = lane; j < cpg; j += 32) {
    const int c = g * cpg + j;
    const float sc = rstd * __ldg(gamma + c);
    scale_shift[((size_t)b * D + c) * 2 + 0] = sc;
    scale_shift[((size_t)b * D + c) * 2 + 1] = __ldg(beta + c) - (float)mean * sc;
  }
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
// FPN top-down merge fused into the apply pass (mmdet MSDeformAttnPixelDecoder.forward:
//   y = lateral_conv(x) [conv -> GN]  +  F.interpolate(top, size=(H,W), mode="bilinear", align_corners=False)):
// x, y NHWC [B,H*W,256]; top token-major [b * top_bstride + (ty*w + tx) * 256 + c].  Same arithmetic order as
// ATen's upsample_bilinear2d (h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11)).
__global__ void __launch_bounds__(256) gn_apply_upadd_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ ss,
                                                                   const float* __restrict__ top, long long top_bstride,
                                                                   float* __restrict__ y, long long n4, int H, int W, int h,
                                                                   int w, float sh, float sw) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;  // float4 index over [B,H*W,64]
  if (i >= n4) return;
  const int cq = (int)(i & 63);
  const long long pix = i >> 6;
  const int HW = H * W;
  const int b = (int)(pix / HW), p = (int)(pix - (long long)b * HW);
  const int oy = p / W, ox = p - oy * W;
  float sy = sh * ((float)oy + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  float sx = sw * ((float)ox + 0.5f) - 0.5f;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = (int)sy, x0 = (int)sx;
  const int yp = (y0 < h - 1) ? 1 : 0, xq = (x0 < w - 1) ? 1 : 0;
  const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
  const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
  const float4* tb = reinterpret_cast<const float4*>(top + (size_t)b * top_bstride) + cq;
  const float4 v00 = __ldg(tb + (size_t)(y0 * w + x0) * 64), v01 = __ldg(tb + (size_t)(y0 * w + x0 + xq) * 64);
  const float4 v10 = __ldg(tb + (size_t)((y0 + yp) * w + x0) * 64), v11 = __ldg(tb + (size_t)((y0 + yp) * w + x0 + xq) * 64);
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(ss) + ((size_t)b * D + cq * 4) / 2);
  const float4 s1 = __ldg(reinterpret_cast<const float4*>(ss) + ((size_t)b * D + cq * 4) / 2 + 1);
  float4 o;
  o.x = fmaf(v.x, s0.x, s0.y) + (ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x));
  o.y = fmaf(v.y, s0.z, s0.w) + (ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y));
  o.z = fmaf(v.z, s1.x, s1.y) + (ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z));
  o.w = fmaf(v.w, s1.z, s1.w) + (ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w));
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

// ---- upstream plumbing: the ResNet stem's MaxPool2d(3, stride 2, pad 1) on a channels_last map.
// ATen's max_pool_forward_nhwc takes 164 us on [2,64,400,667]; this is a plain 128-bit HBM-bound sweep (~35 us).
__global__ void __launch_bounds__(256) maxpool3x3s2_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int H,
                                                                 int W, int Ho, int Wo, int C4, long long n4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;  // float4 index over [B,Ho,Wo,C/4]
  if (i >= n4) return;
  const int c = (int)(i % C4);
  long long t = i / C4;
  const int ox = (int)(t % Wo);
  t /= Wo;
  const int oy = (int)(t % Ho), b = (int)(t / Ho);
  const float4* src = reinterpret_cast<const float4*>(x) + (size_t)b * H * W * C4 + c;
  const float NEG = -__int_as_float(0x7f800000);
  float4 m = make_float4(NEG, NEG, NEG, NEG);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = oy * 2 - 1 + dy;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = ox * 2 - 1 + dx;
      if (ix < 0 || ix >= W) continue;
      const float4 v = __ldg(src + ((size_t)iy * W + ix) * C4);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  reinterpret_cast<float4*>(y)[i] = m;
}

}  // namespace pn

using namespace pn;

extern "C" {

size_t pn_group_norm_workspace_bytes(int B, int HW, int groups) {
  const int nchunk = GN_NCHUNK;
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
  const int nchunk = GN_NCHUNK;
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
  gn_finalize_kernel<<<cdiv(B * groups, 8), 256, 0, st>>>(part, gamma, beta, ss, B, groups, nchunk, count, eps);
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


/* MaxPool2d(kernel 3, stride 2, padding 1) on a channels_last map x [B,H,W,C] -> y [B,Ho,Wo,C], C % 4 == 0 */
int pn_maxpool3x3s2_nhwc(const float* x, float* y, int B, int H, int W, int C, pn_stream_t stream) {
  PN_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, PN_ERR_BAD_ARG, "maxpool: bad args");
  PN_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, PN_ERR_UNSUPPORTED, "maxpool: 16B alignment");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long n4 = (long long)B * Ho * Wo * (C / 4);
  maxpool3x3s2_nhwc_kernel<<<cdiv(n4, 256), 256, 0, as_stream(stream)>>>(x, y, H, W, Ho, Wo, C / 4, n4);
  return check_launch("maxpool3x3s2_nhwc_kernel");
}

/* GroupNorm fused with the FPN top-down merge:  y = GN(x) + bilinear_upsample(top)  (channels_last x / y) */
int pn_gn_upsample_add(const float* x, const float* gamma, const float* beta, const float* top,
                       long long top_batch_stride, float* y, int B, int H, int W, int h, int w, int groups, float eps,
                       void* wsp, size_t ws_bytes, pn_stream_t stream) {
  PN_REQUIRE(x && gamma && beta && top && y && wsp, PN_ERR_BAD_ARG, "gn_upsample_add: null argument");
  PN_REQUIRE(B > 0 && H > 0 && W > 0 && h > 0 && w > 0, PN_ERR_BAD_ARG, "gn_upsample_add: bad sizes");
  PN_REQUIRE(groups > 0 && groups <= 64 && D % groups == 0 && (D / groups) % 4 == 0, PN_ERR_UNSUPPORTED,
             "gn_upsample_add: 256 channels, groups must divide 64");
  const int HW = H * W;
  PN_REQUIRE(ws_bytes >= pn_group_norm_workspace_bytes(B, HW, groups), PN_ERR_WORKSPACE,
             "gn_upsample_add: workspace too small");
  PN_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)top) & 15) == 0 && top_batch_stride % 4 == 0,
             PN_ERR_UNSUPPORTED, "gn_upsample_add: 16B alignment");
  cudaStream_t st = as_stream(stream);
  const int nchunk = GN_NCHUNK;
  double* part = reinterpret_cast<double*>(wsp);
  float* ss = reinterpret_cast<float*>(part + (size_t)B * groups * nchunk * 2);
  gn_stats_nhwc_kernel<<<dim3(nchunk, B), GN_THREADS, 0, st>>>(x, part, HW, cdiv(HW, nchunk), groups, nchunk);
  PN_TRY(check_launch("gn_stats_nhwc_kernel"));
  gn_finalize_kernel<<<cdiv(B * groups, 8), 256, 0, st>>>(part, gamma, beta, ss, B, groups, nchunk, (double)HW * (D / groups), eps);
  PN_TRY(check_launch("gn_finalize_kernel"));
  const long long n4 = (long long)B * HW * (D / 4);
  gn_apply_upadd_nhwc_kernel<<<cdiv(n4, 256), 256, 0, st>>>(x, ss, top, top_batch_stride, y, n4, H, W, h, w,
                                                            (float)h / (float)H, (float)w / (float)W);
  return check_launch("gn_apply_upadd_nhwc_kernel");
}

/* 1x1 convolution of a channels_last map into an NCHW-contiguous map (mask_feature of the pixel decoder):
 * y[b][n][p] = sum_k x[b][p][k] w[n][k] + bias[n];  tcgen05 3xTF32 GEMM, activations split in the SM, transposed store. */
size_t pn_conv1x1_nhwc_to_nchw_workspace_bytes(int cout) { return (size_t)2 * cout * D * sizeof(float) + 512; }
int pn_conv1x1_nhwc_to_nchw(const float* x, const float* w, const float* bias, float* y, int B, int HW, int cout,
                            void* wsp, size_t ws_bytes, pn_stream_t stream) {
  PN_REQUIRE(x && w && y && wsp && B > 0 && HW > 0 && cout > 0, PN_ERR_BAD_ARG, "conv1x1: bad args");
  PN_REQUIRE(ws_bytes >= pn_conv1x1_nhwc_to_nchw_workspace_bytes(cout), PN_ERR_WORKSPACE, "conv1x1: workspace too small");
  cudaStream_t st = as_stream(stream);
  Workspace ws(wsp, ws_bytes);
  float* wh = ws.take<float>((size_t)cout * D);
  float* wl = ws.take<float>((size_t)cout * D);
  PN_TRY(launch_split_tf32(w, wh, wl, (size_t)cout * D, st));
  UmmaOperand o{x, nullptr, D, wh, wl, D, bias, y, HW, B * HW, cout, D};
  o.a_is_raw = 1;
  o.t_rows = HW;
  return launch_umma_gemm(&o, 1, 3, st);
}

}  // extern "C"
