// Scaled-dot-product attention core of nn.MultiheadAttention (8 heads x 32) for the masked
// cross-attention of the Mask2Former decoder (keys = 1 050 / 4 200 / 16 700 pixel tokens), the
// 100x100 query self-attention and the Relation Fusion cross-attention (200 pair keys).
//
// Flash-style: keys are streamed through shared memory in 64-key tiles (cp.async double buffer),
// scores never touch HBM, softmax is online in the log2 domain.  The key range is split across
// CTAs (split-KV) so that B*8 (image, head) pairs still fill 148 SMs; partial (m, l, o) triples are
// merged by a small deterministic combine kernel.  The boolean attention mask is read as packed
// bits shared by all 8 heads (the reference materialises a [B*8, N, hw] float mask).
// Exact fp32 FFMA arithmetic (see gemm.cu for why).
#include "common.cuh"

namespace pn {

constexpr int ATT_TK = 64;       // keys per smem tile
constexpr int ATT_THREADS = 128; // one query per thread
constexpr int ATT_CH = 8;        // keys per online-softmax chunk
constexpr float NEG_BIG = -1.0e30f;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

struct MhaKernelArgs {
  const float* q; int ldq;
  const float* k; int ldk;
  const float* v; int ldv;
  const uint32_t* mask_bits; int mask_words;
  const int* rowany;
  float* out;        // splits == 1: [B,Nq,256] normalised
  float* opart;      // splits > 1 : [S,B,Nq,256] un-normalised
  float2* ml;        // splits > 1 : [S,B,NH,Nq] (running max (log2 domain), running sum)
  int B, Nq, Nk;
  int splits, keys_per_split;  // keys_per_split multiple of ATT_TK
  float qscale;                // (1/sqrt(32)) * log2(e)
};

__global__ void __launch_bounds__(ATT_THREADS) mha_kernel(const MhaKernelArgs a) {
  pdl_wait();     // PDL contract (common.cuh): nothing is read or written before the preceding grid has completed
  pdl_trigger();
  __shared__ __align__(16) float ks[2][ATT_TK][HD];
  __shared__ __align__(16) float vs[2][ATT_TK][HD];
  const int split = blockIdx.x, h = blockIdx.y;
  const int qtiles = (a.Nq + ATT_THREADS - 1) / ATT_THREADS;
  const int b = blockIdx.z / qtiles, qt = blockIdx.z % qtiles;
  const int tid = threadIdx.x;
  const int qi = qt * ATT_THREADS + tid;
  const bool qvalid = qi < a.Nq;

  const int k_begin = split * a.keys_per_split;
  const int k_end = min(a.Nk, k_begin + a.keys_per_split);
  const int ntiles = (k_end - k_begin + ATT_TK - 1) / ATT_TK;

  float q[HD], o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  if (qvalid) {
    const float4* qp = reinterpret_cast<const float4*>(a.q + ((size_t)b * a.Nq + qi) * a.ldq + h * HD);
#pragma unroll
    for (int d4 = 0; d4 < HD / 4; ++d4) {
      const float4 t = __ldg(qp + d4);
      q[d4 * 4 + 0] = t.x * a.qscale; q[d4 * 4 + 1] = t.y * a.qscale;
      q[d4 * 4 + 2] = t.z * a.qscale; q[d4 * 4 + 3] = t.w * a.qscale;
    }
  } else {
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] = 0.f;
  }
  float m_run = NEG_BIG, l_run = 0.f;

  const bool use_mask = a.mask_bits != nullptr && qvalid && (a.rowany == nullptr || a.rowany[b * a.Nq + qi] != 0);
  const uint32_t* mrow = use_mask ? a.mask_bits + ((size_t)b * a.Nq + qi) * a.mask_words : nullptr;

  const float* kbase = a.k + (size_t)b * a.Nk * a.ldk + h * HD;
  const float* vbase = a.v + (size_t)b * a.Nk * a.ldv + h * HD;

  auto issue_tile = [&](int t, int buf) {
    const int key0 = k_begin + t * ATT_TK;
    // 64 keys x 32 floats = 512 float4 per operand; 128 threads x 4
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = tid + r * ATT_THREADS;
      const int row = i >> 3, c4 = i & 7;
      int key = key0 + row;
      key = key < a.Nk ? key : a.Nk - 1;  // clamp: rows past the end are never consumed
      cp_async16(&ks[buf][row][c4 * 4], kbase + (size_t)key * a.ldk + c4 * 4);
      cp_async16(&vs[buf][row][c4 * 4], vbase + (size_t)key * a.ldv + c4 * 4);
    }
    cp_async_commit();
  };

  if (ntiles > 0) issue_tile(0, 0);
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      issue_tile(t + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int key0 = k_begin + t * ATT_TK;
    const int nvalid = min(ATT_TK, k_end - key0);
    uint32_t w0 = 0, w1 = 0;
    if (use_mask) {
      const int wi = key0 >> 5;
      w0 = __ldg(mrow + wi);
      w1 = (wi + 1 < a.mask_words) ? __ldg(mrow + wi + 1) : 0xffffffffu;
    }
    if (qvalid) {
      for (int c0 = 0; c0 < nvalid; c0 += ATT_CH) {
        float s[ATT_CH];
        float cmax = NEG_BIG;
#pragma unroll
        for (int j = 0; j < ATT_CH; ++j) {
          const int kk = c0 + j;
          float acc = 0.f;
          const float4* kp = reinterpret_cast<const float4*>(&ks[buf][kk < ATT_TK ? kk : 0][0]);
#pragma unroll
          for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 kv = kp[d4];
            acc = fmaf(q[d4 * 4 + 0], kv.x, acc);
            acc = fmaf(q[d4 * 4 + 1], kv.y, acc);
            acc = fmaf(q[d4 * 4 + 2], kv.z, acc);
            acc = fmaf(q[d4 * 4 + 3], kv.w, acc);
          }
          const uint32_t w = (kk < 32) ? w0 : w1;
          const bool blocked = (kk >= nvalid) || ((w >> (kk & 31)) & 1u);
          s[j] = blocked ? NEG_BIG : acc;
          cmax = fmaxf(cmax, s[j]);
        }
        if (cmax > m_run) {
          const float corr = exp2f(m_run - cmax);  // m_run == NEG_BIG -> 0
          l_run *= corr;
#pragma unroll
          for (int d = 0; d < HD; ++d) o[d] *= corr;
          m_run = cmax;
        }
        if (m_run > NEG_BIG) {
#pragma unroll
          for (int j = 0; j < ATT_CH; ++j) {
            const int kk = c0 + j;
            const float p = (s[j] > NEG_BIG) ? exp2f(s[j] - m_run) : 0.f;
            l_run += p;
            const float4* vp = reinterpret_cast<const float4*>(&vs[buf][kk < ATT_TK ? kk : 0][0]);
#pragma unroll
            for (int d4 = 0; d4 < HD / 4; ++d4) {
              const float4 vv = vp[d4];
              o[d4 * 4 + 0] = fmaf(p, vv.x, o[d4 * 4 + 0]);
              o[d4 * 4 + 1] = fmaf(p, vv.y, o[d4 * 4 + 1]);
              o[d4 * 4 + 2] = fmaf(p, vv.z, o[d4 * 4 + 2]);
              o[d4 * 4 + 3] = fmaf(p, vv.w, o[d4 * 4 + 3]);
            }
          }
        }
      }
    }
    __syncthreads();  // tile buffer is re-filled two iterations later
  }

  if (!qvalid) return;
  if (a.splits == 1) {
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    float4* op = reinterpret_cast<float4*>(a.out + ((size_t)b * a.Nq + qi) * D + h * HD);
#pragma unroll
    for (int d4 = 0; d4 < HD / 4; ++d4)
      op[d4] = make_float4(o[d4 * 4 + 0] * inv, o[d4 * 4 + 1] * inv, o[d4 * 4 + 2] * inv, o[d4 * 4 + 3] * inv);
  } else {
    float4* op = reinterpret_cast<float4*>(a.opart + (((size_t)split * a.B + b) * a.Nq + qi) * D + h * HD);
#pragma unroll
    for (int d4 = 0; d4 < HD / 4; ++d4)
      op[d4] = make_float4(o[d4 * 4 + 0], o[d4 * 4 + 1], o[d4 * 4 + 2], o[d4 * 4 + 3]);
    a.ml[(((size_t)split * a.B + b) * NH + h) * a.Nq + qi] = make_float2(m_run, l_run);
  }
}

// out[b,q,c] = sum_s w_s o_s[c] / sum_s w_s l_s,  w_s = 2^(m_s - max_s m_s); fixed order over s.
// The loads of a chunk of 8 splits are issued back to back before anything consumes them: with one load per loop
// iteration the kernel was a chain of 3 S exposed L2 round trips (8.9 us for S = 7-9 in the replayed graph, 30 launches per
// forward on the critical chain of the head).
__global__ void __launch_bounds__(256) mha_combine_kernel(const float* __restrict__ opart,
                                                           const float2* __restrict__ ml, float* __restrict__ out,
                                                           int B, int Nq, int S) {
  pdl_wait();     // PDL contract (common.cuh)
  pdl_trigger();
  const int bq = blockIdx.x;
  const int b = bq / Nq, qi = bq % Nq;
  const int c = threadIdx.x, h = c / HD;
  constexpr int CHUNK = 8;
  const size_t ml_stride = (size_t)B * NH * Nq, op_stride = (size_t)B * Nq * D;
  const float2* mlp = ml + ((size_t)b * NH + h) * Nq + qi;
  const float* opp = opart + ((size_t)b * Nq + qi) * D + c;
  float mmax = NEG_BIG;
  for (int s0 = 0; s0 < S; s0 += CHUNK) {
    float m[CHUNK];
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) m[j] = (s0 + j < S) ? mlp[(size_t)(s0 + j) * ml_stride].x : NEG_BIG;
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) mmax = fmaxf(mmax, m[j]);
  }
  float num = 0.f, den = 0.f;
  for (int s0 = 0; s0 < S; s0 += CHUNK) {
    float2 t[CHUNK];
    float o[CHUNK];
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) {
      const bool live = s0 + j < S;
      t[j] = live ? mlp[(size_t)(s0 + j) * ml_stride] : make_float2(NEG_BIG, 0.f);
      o[j] = live ? opp[(size_t)(s0 + j) * op_stride] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) {   // same order and arithmetic as a plain loop over s
      const float w = (t[j].y > 0.f) ? exp2f(t[j].x - mmax) : 0.f;
      num = fmaf(w, o[j], num);
      den = fmaf(w, t[j].y, den);
    }
  }
  out[((size_t)b * Nq + qi) * D + c] = den > 0.f ? num / den : 0.f;
}

// split granularity: whole 64-key tiles when the mask is present (bit words are addressed per tile); 16 keys
// otherwise, so the 100-key self-attention and 200-key relation attention still spread over >100 CTAs.
int launch_mha_combine(const float* opart, const float2* ml, float* out, int B, int Nq, int S, cudaStream_t st) {
  launch_pdl(mha_combine_kernel, dim3(B * Nq), dim3(256), 0, st, opart, ml, out, B, Nq, S);
  return check_launch("mha_combine_kernel");
}

static void pick_splits(int B, int Nq, int Nk, bool masked, int* splits, int* keys_per_split) {
  const int qtiles = cdiv(Nq, ATT_THREADS);
  const int base = B * NH * qtiles;
  if (!masked && Nk <= 512) {
    int want = cdiv(148, base);
    int kps = (int)round_up(cdiv(Nk, want), 16);
    kps = kps < 16 ? 16 : kps;
    *keys_per_split = kps;
    *splits = cdiv(Nk, kps);
    return;
  }
  const int tiles = cdiv(Nk, ATT_TK);
  int want = cdiv(4 * 148, base);  // ~4 CTAs (16 warps) per SM
  if (want > tiles) want = tiles;
  if (want > 64) want = 64;
  if (want < 1) want = 1;
  int tps = cdiv(tiles, want);
  *keys_per_split = tps * ATT_TK;
  *splits = cdiv(tiles, tps);
}

size_t mha_workspace_bytes(int B, int Nq, int Nk) {
  int S, kps, S2, kps2;
  pick_splits(B, Nq, Nk, true, &S, &kps);
  pick_splits(B, Nq, Nk, false, &S2, &kps2);
  S = S > S2 ? S : S2;
  if (S == 1) return 256;
  size_t o = ((size_t)S * B * Nq * D * sizeof(float) + 255) & ~size_t(255);
  size_t m = ((size_t)S * B * NH * Nq * sizeof(float2) + 255) & ~size_t(255);
  return o + m;
}

int launch_mha(const MhaArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  PN_REQUIRE(a.q && a.k && a.v && a.out, PN_ERR_BAD_ARG, "mha: null pointer");
  PN_REQUIRE(a.B > 0 && a.Nq > 0 && a.Nk > 0, PN_ERR_BAD_ARG, "mha: bad shape");
  PN_REQUIRE((a.ldq & 3) == 0 && (a.ldk & 3) == 0 && (a.ldv & 3) == 0, PN_ERR_UNSUPPORTED, "mha: strides");
  PN_REQUIRE((((uintptr_t)a.q | (uintptr_t)a.k | (uintptr_t)a.v | (uintptr_t)a.out) & 15) == 0, PN_ERR_UNSUPPORTED,
             "mha: pointers must be 16B aligned");
  PN_REQUIRE(!a.mask_bits || a.mask_words * 32 >= a.Nk, PN_ERR_BAD_ARG, "mha: mask_words too small");
  MhaKernelArgs k{};
  k.q = a.q; k.ldq = a.ldq; k.k = a.k; k.ldk = a.ldk; k.v = a.v; k.ldv = a.ldv;
  k.mask_bits = a.mask_bits; k.mask_words = a.mask_words; k.rowany = a.rowany;
  k.out = a.out; k.B = a.B; k.Nq = a.Nq; k.Nk = a.Nk;
  k.qscale = 0.17677669529663687f * 1.4426950408889634f;
  pick_splits(a.B, a.Nq, a.Nk, a.mask_bits != nullptr, &k.splits, &k.keys_per_split);
  if (k.splits > 1) {
    PN_REQUIRE(ws && ws_bytes >= mha_workspace_bytes(a.B, a.Nq, a.Nk), PN_ERR_WORKSPACE, "mha: workspace too small");
    size_t o = ((size_t)k.splits * a.B * a.Nq * D * sizeof(float) + 255) & ~size_t(255);
    k.opart = reinterpret_cast<float*>(ws);
    k.ml = reinterpret_cast<float2*>(reinterpret_cast<char*>(ws) + o);
  }
  dim3 grid(k.splits, NH, a.B * cdiv(a.Nq, ATT_THREADS));
  launch_pdl(mha_kernel, grid, dim3(ATT_THREADS), 0, st, k);
  PN_TRY(check_launch("mha_kernel"));
  if (k.splits > 1) {
    launch_pdl(mha_combine_kernel, dim3(a.B * a.Nq), dim3(256), 0, st, k.opart, k.ml, a.out, a.B, a.Nq, k.splits);
    PN_TRY(check_launch("mha_combine_kernel"));
  }
  return 0;
}

}  // namespace pn
