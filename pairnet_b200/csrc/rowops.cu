// Row-wise / bandwidth-bound kernels: fused (split-K reduce + bias + residual + LayerNorm
// [+ query_pos add] [+ chained post_norm]), L2 normalise, learned-query broadcast, level prep
// (transpose + level_embed + sine positional encoding), bilinear mask-feature resize, row gathers.
// All are HBM/L2-bound: one warp per 256-channel row, 128-bit accesses, grid sized to the rows.
#include "common.cuh"

namespace pn {

__device__ __forceinline__ float rna_tf32f(float v) {  // == cvt.rna.tf32.f32 for finite inputs, 2 integer ops (umma_ptx.cuh)
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void ln_row(float (&v)[8], const float* __restrict__ gamma,
                                       const float* __restrict__ beta, int lane) {
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
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 32 + lane);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + lane);
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 32 + lane);
  const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * g[i] + bb[i];
}

__device__ __forceinline__ void load_row(float (&v)[8], const float* __restrict__ p, int lane) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p) + lane);
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 32 + lane);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store_row(const float (&v)[8], float* __restrict__ p, int lane) {
  reinterpret_cast<float4*>(p)[lane] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[32 + lane] = make_float4(v[4], v[5], v[6], v[7]);
}

__device__ __forceinline__ void store_split(const float (&v)[8], float* __restrict__ hi, float* __restrict__ lo,
                                            int lane) {
  float h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = rna_tf32f(v[i]);
    l[i] = rna_tf32f(v[i] - h[i]);
  }
  store_row(h, hi, lane);
  store_row(l, lo, lane);
}

// lane owns channels [4*lane, 4*lane+4) and [128+4*lane, 128+4*lane+4)
__global__ void __launch_bounds__(256) layernorm_kernel(const LnArgs a) {
  pdl_wait();     // PDL contract (common.cuh): nothing is read or written before the preceding grid has completed
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= a.M) return;
  if (a.zero_rows && lane == 0) a.zero_rows[m] = 0;
  float v[8];
  load_row(v, a.x + (size_t)m * D, lane);
  for (int s = 1; s < a.nparts; ++s) {  // fixed order -> deterministic split-K reduction
    float t[8];
    load_row(t, a.x + (size_t)s * a.part_stride + (size_t)m * D, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += t[i];
  }
  if (a.bias) {
    float t[8];
    load_row(t, a.bias, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += t[i];
  }
  if (a.resid) {
    float t[8];
    load_row(t, a.resid + (size_t)m * D, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = t[i] + v[i];
  }
  ln_row(v, a.gamma, a.beta, lane);
  store_row(v, a.y + (size_t)m * D, lane);
  if (a.y_hi) store_split(v, a.y_hi + (size_t)m * D, a.y_lo + (size_t)m * D, lane);
  if (a.ypos || a.ypos_hi) {
    float t[8];
    load_row(t, a.pos + (size_t)(m % a.pos_mod) * D, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = v[i] + t[i];
    if (a.ypos) store_row(t, a.ypos + (size_t)m * D, lane);
    if (a.ypos_hi) store_split(t, a.ypos_hi + (size_t)m * D, a.ypos_lo + (size_t)m * D, lane);
  }
  if (a.y2) {
    ln_row(v, a.gamma2, a.beta2, lane);
    store_row(v, a.y2 + (size_t)m * D, lane);
  }
}

int launch_layernorm(const LnArgs& a, cudaStream_t st) {
  PN_REQUIRE(a.x && a.gamma && a.beta && a.y && a.M > 0 && a.nparts >= 1, PN_ERR_BAD_ARG, "layernorm: bad args");
  PN_REQUIRE(!(a.ypos || a.ypos_hi) || (a.pos && a.pos_mod > 0), PN_ERR_BAD_ARG, "layernorm: ypos needs pos");
  PN_REQUIRE((a.y_hi == nullptr) == (a.y_lo == nullptr) && (a.ypos_hi == nullptr) == (a.ypos_lo == nullptr),
             PN_ERR_BAD_ARG, "layernorm: split outputs come in hi/lo pairs");
  PN_REQUIRE(!a.y2 || (a.gamma2 && a.beta2), PN_ERR_BAD_ARG, "layernorm: y2 needs gamma2/beta2");
  // PDL only for the query-side chain (a few hundred rows, latency bound); the 43 900-token LayerNorms of the pixel-decoder
  // encoder measured slightly slower with it (thousands of CTAs made resident early)
  if (a.M <= 4096) launch_pdl(layernorm_kernel, dim3(cdiv(a.M, 8)), dim3(256), 0, st, a);
  else layernorm_kernel<<<cdiv(a.M, 8), 256, 0, st>>>(a);
  return check_launch("layernorm_kernel");
}

// F.normalize(x, p=2, dim=-1, eps=1e-12): x / max(||x||_2, eps)      (pairnet_head.py:325-326)
__global__ void __launch_bounds__(256) l2norm_kernel(const float* __restrict__ x, float* __restrict__ y, int M) {
  pdl_wait();     // PDL contract (common.cuh)
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  float v[8];
  load_row(v, x + (size_t)m * D, lane);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) q += v[i] * v[i];
  const float nrm = fmaxf(sqrtf(warp_sum(q)), 1e-12f);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = v[i] / nrm;
  store_row(v, y + (size_t)m * D, lane);
}
int launch_l2norm(const float* x, float* y, int M, cudaStream_t st) {
  launch_pdl(l2norm_kernel, dim3(cdiv(M, 8)), dim3(256), 0, st, x, y, M);
  return check_launch("l2norm_kernel");
}

// out[b,n,:] = a[n,:] ; out_sum[b,n,:] = a[n,:] + b[n,:]   (learned queries repeated per image)
__global__ void __launch_bounds__(256) bcast_rows_kernel(const float* __restrict__ a, const float* __restrict__ bpos,
                                                          float* __restrict__ out, float* __restrict__ out_sum,
                                                          int B, int N) {
  pdl_wait();     // PDL contract (common.cuh)
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= B * N) return;
  const int n = m % N;
  float v[8];
  load_row(v, a + (size_t)n * D, lane);
  if (out) store_row(v, out + (size_t)m * D, lane);
  if (out_sum) {
    float t[8];
    load_row(t, bpos + (size_t)n * D, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = v[i] + t[i];
    store_row(t, out_sum + (size_t)m * D, lane);
  }
}
int launch_bcast_rows(const float* a, const float* b, float* out, float* out_sum, int B, int N, cudaStream_t st) {
  launch_pdl(bcast_rows_kernel, dim3(cdiv((long long)B * N, 8)), dim3(256), 0, st, a, b, out, out_sum, B, N);
  return check_launch("bcast_rows_kernel");
}

// out[b,n,:] = x[b,n,:] + pos[n,:]
__global__ void __launch_bounds__(256) add_rows_kernel(const float* __restrict__ x, const float* __restrict__ pos,
                                                        float* __restrict__ out, int B, int N) {
  pdl_wait();     // PDL contract (common.cuh)
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= B * N) return;
  float v[8], t[8];
  load_row(v, x + (size_t)m * D, lane);
  load_row(t, pos + (size_t)(m % N) * D, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = v[i] + t[i];
  store_row(v, out + (size_t)m * D, lane);
}
int launch_add_rows(const float* x, const float* pos, float* out, int B, int N, cudaStream_t st) {
  launch_pdl(add_rows_kernel, dim3(cdiv((long long)B * N, 8)), dim3(256), 0, st, x, pos, out, B, N);
  return check_launch("add_rows_kernel");
}

// ---------------------------------------------------------------------------------------------
// mmdet SinePositionalEncoding(num_feats=128, normalize=True, temperature=1e4, scale=2pi,
// eps=1e-6, offset=0) of an all-false mask, written token-major [h*w, 256]:
// channels [0,128) = pos_y, [128,256) = pos_x; even channel -> sin, odd -> cos.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sine_posenc_kernel(float* __restrict__ pos, int h, int w) {
  const int p = blockIdx.x;
  const int c = threadIdx.x;
  const int y = p / w, x = p % w;
  const float scale = 6.283185307179586f;  // float32(2*pi)
  const bool is_x = c >= 128;
  const int i = c & 127;
  const float embed = is_x ? ((float)(x + 1) / ((float)w + 1e-6f)) * scale
                           : ((float)(y + 1) / ((float)h + 1e-6f)) * scale;
  const float expo = (float)(2 * (i / 2)) / 128.f;
  const float dim_t = (float)pow(10000.0, (double)expo);
  const float arg = embed / dim_t;
  pos[(size_t)p * D + c] = (i & 1) ? cosf(arg) : sinf(arg);
}
int launch_sine_posenc(float* pos, int h, int w, cudaStream_t st) {
  PN_REQUIRE(pos && h > 0 && w > 0, PN_ERR_BAD_ARG, "sine_posenc: bad args");
  sine_posenc_kernel<<<h * w, 256, 0, st>>>(pos, h, w);
  return check_launch("sine_posenc_kernel");
}

// mem [B,256,hw] -> x [B,hw,256] (+ level_embed) ; xp = x + pos.   32x32 smem transpose tiles.
// x_lo (xp_lo) != null: 3xTF32 pre-split form -- x (xp) receives hi = rna_tf32(v), x_lo (xp_lo) receives
// rna_tf32(v - hi); otherwise the raw fp32 value is stored.
__global__ void __launch_bounds__(256) level_prep_kernel(const float* __restrict__ mem,
                                                          const float* __restrict__ level_embed,
                                                          const float* __restrict__ pos, float* __restrict__ x,
                                                          float* __restrict__ xp, float* __restrict__ x_lo,
                                                          float* __restrict__ xp_lo, int hw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* src = mem + (size_t)b * D * hw;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = c0 + ty + r * 8, p = p0 + tx;
    tile[ty + r * 8][tx] = (p < hw) ? __ldg(src + (size_t)c * hw + p) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int p = p0 + ty + r * 8, c = c0 + tx;
    if (p < hw) {
      const float v = tile[tx][ty + r * 8] + __ldg(level_embed + c);
      const size_t o = ((size_t)b * hw + p) * D + c;
      const float vp = v + __ldg(pos + (size_t)p * D + c);
      if (x_lo) {
        const float h = rna_tf32f(v);
        x[o] = h;
        x_lo[o] = rna_tf32f(v - h);
      } else {
        x[o] = v;
      }
      if (xp_lo) {
        const float hp = rna_tf32f(vp);
        xp[o] = hp;
        xp_lo[o] = rna_tf32f(vp - hp);
      } else {
        xp[o] = vp;
      }
    }
  }
}
// token-major source (a level slice of the pixel decoder's encoder output): no transpose, 128-bit elementwise
__global__ void __launch_bounds__(256) level_prep_tokens_kernel(const float* __restrict__ mem, long long bstride,
                                                                 const float* __restrict__ level_embed,
                                                                 const float* __restrict__ pos, float* __restrict__ x,
                                                                 float* __restrict__ xp, float* __restrict__ x_lo,
                                                                 float* __restrict__ xp_lo, int hw, long long n4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;  // float4 index over [B,hw,64]
  if (i >= n4) return;
  const int cq = (int)(i & 63);
  const long long tok = i >> 6;
  const int b = (int)(tok / hw), p = (int)(tok - (long long)b * hw);
  const float4 m = __ldg(reinterpret_cast<const float4*>(mem + (size_t)b * bstride + (size_t)p * D) + cq);
  const float4 le = __ldg(reinterpret_cast<const float4*>(level_embed) + cq);
  const float4 ps = __ldg(reinterpret_cast<const float4*>(pos + (size_t)p * D) + cq);
  const float v[4] = {m.x + le.x, m.y + le.y, m.z + le.z, m.w + le.w};
  const float vp[4] = {v[0] + ps.x, v[1] + ps.y, v[2] + ps.z, v[3] + ps.w};
  float a[4], al[4], c[4], cl[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    a[u] = x_lo ? rna_tf32f(v[u]) : v[u];
    al[u] = x_lo ? rna_tf32f(v[u] - a[u]) : 0.f;
    c[u] = xp_lo ? rna_tf32f(vp[u]) : vp[u];
    cl[u] = xp_lo ? rna_tf32f(vp[u] - c[u]) : 0.f;
  }
  reinterpret_cast<float4*>(x)[i] = make_float4(a[0], a[1], a[2], a[3]);
  reinterpret_cast<float4*>(xp)[i] = make_float4(c[0], c[1], c[2], c[3]);
  if (x_lo) reinterpret_cast<float4*>(x_lo)[i] = make_float4(al[0], al[1], al[2], al[3]);
  if (xp_lo) reinterpret_cast<float4*>(xp_lo)[i] = make_float4(cl[0], cl[1], cl[2], cl[3]);
}
int launch_level_prep_tokens(const float* mem, long long bstride, const float* level_embed, const float* pos, float* x,
                             float* xp, int B, int hw, cudaStream_t st, float* x_lo, float* xp_lo) {
  PN_REQUIRE(mem && level_embed && pos && x && xp && B > 0 && hw > 0, PN_ERR_BAD_ARG, "level_prep: bad args");
  PN_REQUIRE(((uintptr_t)mem & 15) == 0 && bstride % 4 == 0 && bstride >= (long long)hw * D, PN_ERR_UNSUPPORTED,
             "level_prep: token-major memory must be 16B aligned with a batch stride >= hw*256");
  const long long n4 = (long long)B * hw * (D / 4);
  level_prep_tokens_kernel<<<cdiv(n4, 256), 256, 0, st>>>(mem, bstride, level_embed, pos, x, xp, x_lo, xp_lo, hw, n4);
  return check_launch("level_prep_tokens_kernel");
}

int launch_level_prep(const float* mem, const float* level_embed, const float* pos, float* x, float* xp, int B,
                      int hw, cudaStream_t st, float* x_lo, float* xp_lo) {
  PN_REQUIRE(mem && level_embed && pos && x && xp && B > 0 && hw > 0, PN_ERR_BAD_ARG, "level_prep: bad args");
  dim3 grid(cdiv(hw, 32), D / 32, B);
  level_prep_kernel<<<grid, 256, 0, st>>>(mem, level_embed, pos, x, xp, x_lo, xp_lo, hw);
  return check_launch("level_prep_kernel");
}

// F.interpolate(mode="bilinear", align_corners=False) of [B,256,H,W] to (h,w); rows padded to ldo.
__global__ void __launch_bounds__(256) mask_feature_resize_kernel(const float* __restrict__ F,
                                                                   float* __restrict__ out, int H, int W, int h,
                                                                   int w, int ldo, float sh, float sw) {
  const int o = blockIdx.x * 256 + threadIdx.x;
  const size_t plane = blockIdx.y;  // b*256 + c
  if (o >= ldo) return;
  float r = 0.f;
  if (o < h * w) {
    const int oy = o / w, ox = o % w;
    float sy = sh * ((float)oy + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    float sx = sw * ((float)ox + 0.5f) - 0.5f;
    sx = sx < 0.f ? 0.f : sx;
    const int y0 = (int)sy, x0 = (int)sx;
    const int yp = (y0 < H - 1) ? 1 : 0, xq = (x0 < W - 1) ? 1 : 0;
    const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
    const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
    const float* src = F + plane * (size_t)H * W;
    const float v00 = __ldg(src + (size_t)y0 * W + x0), v01 = __ldg(src + (size_t)y0 * W + x0 + xq);
    const float v10 = __ldg(src + (size_t)(y0 + yp) * W + x0), v11 = __ldg(src + (size_t)(y0 + yp) * W + x0 + xq);
    r = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
  }
  out[plane * (size_t)ldo + o] = r;
}
int launch_mask_feature_resize(const float* F, float* out, int B, int H, int W, int h, int w, int ldo,
                               cudaStream_t st) {
  PN_REQUIRE(F && out && ldo >= h * w && ldo % 32 == 0, PN_ERR_BAD_ARG, "mask_feature_resize: bad args");
  dim3 grid(cdiv(ldo, 256), B * D);
  mask_feature_resize_kernel<<<grid, 256, 0, st>>>(F, out, H, W, h, w, ldo, (float)H / (float)h,
                                                   (float)W / (float)w);
  return check_launch("mask_feature_resize_kernel");
}

// token-major variant: F [B,H*W,256] (channels_last mask_features) -> out [B,h*w,256]; same arithmetic
__global__ void __launch_bounds__(256) mask_feature_resize_tokens_kernel(const float* __restrict__ F,
                                                                          float* __restrict__ out, int H, int W, int h,
                                                                          int w, float sh, float sw, long long n4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;  // float4 index over [B,h*w,64]
  if (i >= n4) return;
  const int cq = (int)(i & 63);
  const long long tok = i >> 6;
  const int hw = h * w;
  const int b = (int)(tok / hw), o = (int)(tok - (long long)b * hw);
  const int oy = o / w, ox = o - oy * w;
  float sy = sh * ((float)oy + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  float sx = sw * ((float)ox + 0.5f) - 0.5f;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = (int)sy, x0 = (int)sx;
  const int yp = (y0 < H - 1) ? 1 : 0, xq = (x0 < W - 1) ? 1 : 0;
  const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
  const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
  const float4* src = reinterpret_cast<const float4*>(F + (size_t)b * H * W * D) + cq;
  const float4 v00 = __ldg(src + (size_t)(y0 * W + x0) * 64), v01 = __ldg(src + (size_t)(y0 * W + x0 + xq) * 64);
  const float4 v10 = __ldg(src + (size_t)((y0 + yp) * W + x0) * 64);
  const float4 v11 = __ldg(src + (size_t)((y0 + yp) * W + x0 + xq) * 64);
  float4 r;
  r.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
  r.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
  r.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
  r.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
  reinterpret_cast<float4*>(out)[i] = r;
}
int launch_mask_feature_resize_tokens(const float* F, float* out, int B, int H, int W, int h, int w, cudaStream_t st) {
  PN_REQUIRE(F && out && B > 0 && H > 0 && W > 0 && h > 0 && w > 0, PN_ERR_BAD_ARG, "mask_feature_resize_tokens: bad args");
  PN_REQUIRE((((uintptr_t)F | (uintptr_t)out) & 15) == 0, PN_ERR_UNSUPPORTED, "mask_feature_resize_tokens: alignment");
  const long long n4 = (long long)B * h * w * (D / 4);
  mask_feature_resize_tokens_kernel<<<cdiv(n4, 256), 256, 0, st>>>(F, out, H, W, h, w, (float)H / (float)h,
                                                                   (float)W / (float)w, n4);
  return check_launch("mask_feature_resize_tokens_kernel");
}

// [B,256,HW] (NCHW) -> [B,HW,256] (token-major); 32x32 smem transpose tiles
__global__ void __launch_bounds__(256) nchw_to_tokens_kernel(const float* __restrict__ src, float* __restrict__ dst, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* s = src + (size_t)b * D * HW;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = c0 + ty + r * 8, p = p0 + tx;
    tile[ty + r * 8][tx] = (p < HW) ? __ldg(s + (size_t)c * HW + p) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int p = p0 + ty + r * 8, c = c0 + tx;
    if (p < HW) dst[((size_t)b * HW + p) * D + c] = tile[tx][ty + r * 8];
  }
}
int launch_nchw_to_tokens(const float* src, float* dst, int B, int HW, cudaStream_t st) {
  PN_REQUIRE(src && dst && B > 0 && HW > 0, PN_ERR_BAD_ARG, "nchw_to_tokens: bad args");
  dim3 grid(cdiv(HW, 32), D / 32, B);
  nchw_to_tokens_kernel<<<grid, 256, 0, st>>>(src, dst, HW);
  return check_launch("nchw_to_tokens_kernel");
}

// dst[b,r,:] = src[b, idx[b,r], :]  (row length L floats); grid.x = B*R rows, grid.y = chunks
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src,
                                                           const int64_t* __restrict__ idx,
                                                           float* __restrict__ dst, int Nsrc, int R, long long L,
                                                           int vec) {
  pdl_wait();     // PDL contract (common.cuh)
  pdl_trigger();
  const int br = blockIdx.x;
  const int b = br / R;
  long long row = idx[br];
  row = row < 0 ? 0 : (row >= Nsrc ? Nsrc - 1 : row);
  const float* s = src + ((size_t)b * Nsrc + row) * L;
  float* d = dst + (size_t)br * L;
  if (vec) {
    const long long n4 = L / 4;
    for (long long i = (long long)blockIdx.y * 256 + threadIdx.x; i < n4; i += (long long)gridDim.y * 256)
      reinterpret_cast<float4*>(d)[i] = __ldg(reinterpret_cast<const float4*>(s) + i);
  } else {
    for (long long i = (long long)blockIdx.y * 256 + threadIdx.x; i < L; i += (long long)gridDim.y * 256)
      d[i] = __ldg(s + i);
  }
}
int launch_gather_rows(const float* src, const int64_t* idx, float* dst, int B, int Nsrc, int R, long long L,
                       cudaStream_t st) {
  PN_REQUIRE(src && idx && dst && B > 0 && R > 0 && L > 0, PN_ERR_BAD_ARG, "gather_rows: bad args");
  const int vec = (L % 4 == 0) && (((uintptr_t)src & 15) == 0) && (((uintptr_t)dst & 15) == 0);
  int chunks = cdiv(vec ? L / 4 : L, 256 * 8);
  chunks = chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks);
  dim3 grid(B * R, chunks);
  launch_pdl(gather_rows_kernel, grid, dim3(256), 0, st, src, idx, dst, Nsrc, R, L, vec);
  return check_launch("gather_rows_kernel");
}

}  // namespace pn
