// Inference post-processing of CrossHead2 (pairnet_head.py:788-924, `_get_bboxes_single`).
//
// The reference upsamples three [100, 200, 334] logit stacks to the image size in fp32 (3 x 427 MB at 800x1333),
// thresholds two of them and runs softmax / argmax over the third, then counts areas with one `.item()` per mask.
// Here both consumers read the quarter-resolution logits directly (L2 resident, 26.7 MB per image) and evaluate the
// bilinear interpolation (ATen upsample_bilinear2d, align_corners = False) on the fly:
//   * upsample_threshold_kernel : masks[r] = sigmoid(up(mask[idx[r]])) > 0.5   (evaluated as up(.) > 0)  -> uint8
//   * panoptic_merge_kernel     : m_id = argmax_k up(mask[keep[k]]) (first maximum, = argmax of the softmax),
//                                 stuff de-duplication through a remap table, pan = m_id * OFFSET + label[m_id],
//                                 per-segment pixel areas (block histogram + atomics)
// HBM-bound byte / index work: one pass over the quarter-resolution logits from L2, one write of the result.
#include "common.cuh"

namespace pn {

struct Bilin {
  int o00, o01, o10, o11;  // offsets of the four taps inside one [h, w] plane
  float ly0, ly1, lx0, lx1;
};
__device__ __forceinline__ Bilin bilin_setup(int oy, int ox, int h, int w, float sh, float sw) {
  float sy = sh * ((float)oy + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  float sx = sw * ((float)ox + 0.5f) - 0.5f;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = (int)sy, x0 = (int)sx;
  const int yp = (y0 < h - 1) ? 1 : 0, xq = (x0 < w - 1) ? 1 : 0;
  Bilin b;
  b.ly1 = sy - (float)y0; b.ly0 = 1.f - b.ly1;
  b.lx1 = sx - (float)x0; b.lx0 = 1.f - b.lx1;
  b.o00 = y0 * w + x0; b.o01 = b.o00 + xq; b.o10 = (y0 + yp) * w + x0; b.o11 = b.o10 + xq;
  return b;
}
__device__ __forceinline__ float bilin_eval(const float* __restrict__ p, const Bilin& b) {
  return b.ly0 * (b.lx0 * __ldg(p + b.o00) + b.lx1 * __ldg(p + b.o01)) +
         b.ly1 * (b.lx0 * __ldg(p + b.o10) + b.lx1 * __ldg(p + b.o11));
}

// grid: (ceil(W/256), H, R)
__global__ void __launch_bounds__(256) upsample_threshold_kernel(const float* __restrict__ mask,
                                                                  const int64_t* __restrict__ idx,
                                                                  uint8_t* __restrict__ out, int N, int h, int w, int H,
                                                                  int W, float sh, float sw) {
  const int ox = blockIdx.x * 256 + threadIdx.x, oy = blockIdx.y, r = blockIdx.z;
  if (ox >= W) return;
  long long row = idx ? idx[r] : (long long)r;
  row = row < 0 ? 0 : (row >= N ? N - 1 : row);
  const Bilin b = bilin_setup(oy, ox, h, w, sh, sw);
  const float v = bilin_eval(mask + (size_t)row * h * w, b);
  out[((size_t)r * H + oy) * W + ox] = v > 0.f ? 1 : 0;
}

constexpr int PAN_MAX_KEEP = 1024;
// grid: (ceil(W/256), H)
__global__ void __launch_bounds__(256) panoptic_merge_kernel(const float* __restrict__ mask,
                                                              const int* __restrict__ keep_idx,
                                                              const int* __restrict__ remap,
                                                              const int64_t* __restrict__ labels, int n_keep, int h,
                                                              int w, int H, int W, float sh, float sw, long long offset,
                                                              int64_t* __restrict__ pan, int* __restrict__ area) {
  __shared__ int s_area[PAN_MAX_KEEP];
  __shared__ int s_keep[PAN_MAX_KEEP];
  for (int i = threadIdx.x; i < n_keep; i += 256) { s_area[i] = 0; s_keep[i] = keep_idx[i]; }
  __syncthreads();
  const int ox = blockIdx.x * 256 + threadIdx.x, oy = blockIdx.y;
  if (ox < W) {
    const Bilin b = bilin_setup(oy, ox, h, w, sh, sw);
    const size_t plane = (size_t)h * w;
    float best = bilin_eval(mask + (size_t)s_keep[0] * plane, b);
    int arg = 0;
    for (int k = 1; k < n_keep; ++k) {
      const float v = bilin_eval(mask + (size_t)s_keep[k] * plane, b);
      if (v > best) { best = v; arg = k; }   // strict: first maximum wins, like torch.argmax
    }
    const int id = remap[arg];
    pan[(size_t)oy * W + ox] = (long long)id * offset + labels[id];
    atomicAdd(&s_area[id], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_keep; i += 256)
    if (s_area[i]) atomicAdd(&area[i], s_area[i]);
}

}  // namespace pn

using namespace pn;

extern "C" {

int pn_upsample_threshold(const float* mask, const int64_t* idx, uint8_t* out, int N, int R, int h, int w, int H, int W,
                          pn_stream_t stream) {
  PN_REQUIRE(mask && out && N > 0 && R > 0 && h > 0 && w > 0 && H > 0 && W > 0, PN_ERR_BAD_ARG,
             "upsample_threshold: bad args");
  PN_REQUIRE(R <= 65535 && H <= 65535, PN_ERR_UNSUPPORTED, "upsample_threshold: R and H must fit a grid dimension");
  dim3 grid(cdiv(W, 256), H, R);
  upsample_threshold_kernel<<<grid, 256, 0, as_stream(stream)>>>(mask, idx, out, N, h, w, H, W, (float)h / (float)H,
                                                                 (float)w / (float)W);
  return check_launch("upsample_threshold_kernel");
}

int pn_panoptic_merge(const float* mask, const int* keep_idx, const int* remap, const int64_t* labels, int n_keep, int h,
                      int w, int H, int W, long long instance_offset, int64_t* pan, int* area, pn_stream_t stream) {
  PN_REQUIRE(mask && keep_idx && remap && labels && pan && area, PN_ERR_BAD_ARG, "panoptic_merge: null pointer");
  PN_REQUIRE(n_keep >= 1 && n_keep <= PAN_MAX_KEEP, PN_ERR_UNSUPPORTED, "panoptic_merge: n_keep=%d (1..%d)", n_keep,
             PAN_MAX_KEEP);
  PN_REQUIRE(h > 0 && w > 0 && H > 0 && W > 0 && H <= 65535, PN_ERR_BAD_ARG, "panoptic_merge: bad sizes");
  cudaStream_t st = as_stream(stream);
  count_launch();
  cudaError_t e = cudaMemsetAsync(area, 0, sizeof(int) * n_keep, st);
  PN_REQUIRE(e == cudaSuccess, (int)e, "panoptic_merge: cudaMemsetAsync: %s", cudaGetErrorString(e));
  dim3 grid(cdiv(W, 256), H);
  panoptic_merge_kernel<<<grid, 256, 0, st>>>(mask, keep_idx, remap, labels, n_keep, h, w, H, W, (float)h / (float)H,
                                              (float)w / (float)W, instance_offset, pan, area);
  return check_launch("panoptic_merge_kernel");
}

}  // extern "C"
