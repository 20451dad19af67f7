// Tensor-core masked cross-attention (Mask2Former decoder, keys = pixel tokens of one memory level) on tcgen05.
//
//   per CTA = (image b, head h, key split):   for each 128-key tile
//     S[128 q x 128 keys]  = Q_h K_h^T          tcgen05.mma kind::tf32, Q/K tiles staged by TMA (SWIZZLE_128B), S in TMEM
//     softmax (online, log2 domain, packed-bit mask) by 4 warps, one query row per thread, straight out of TMEM
//     P (hi/lo tf32 split) written back to TMEM with tcgen05.st and used as the A operand of
//     O_tile[128 q x 32]   = P V_h               tcgen05.mma with A from TMEM, B = V_h^T tile (keys contiguous) from smem
//     O accumulated in registers with the usual running-max rescale
//   partial (m, l, o) per split -> mha_combine_kernel (same merge as the FFMA kernel).
//
// fp32 parity: every product is 3xTF32 (lo*hi + hi*lo + hi*hi) -- Q, K, V^T arrive pre-split from their
// producers (K/V projection epilogue of umma_gemm.cu), P is split in registers.
#include "umma_ptx.cuh"

namespace pn {
namespace fa {

using namespace umma;

constexpr int TQ = 128, TKEYS = 128;
constexpr int TILE16K = 16384;                 // 128 rows x 128 B
constexpr int VT_ATOM = 4096;                  // 32 dims x 128 B (32 keys)
constexpr int STAGE_BYTES = 4 * TILE16K;       // K_hi, K_lo, VT_hi, VT_lo
constexpr int KV_STAGES = 2;
constexpr int NUM_THREADS = 192;               // warp0 TMA, warp1 MMA + TMEM, warps2-5 softmax
constexpr int TM_S = 0, TM_PHI = 128, TM_PLO = 256, TM_O = 384, TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = 2 * TILE16K + (size_t)KV_STAGES * STAGE_BYTES + 1024 + 256;
constexpr float NEG_BIG = -1.0e30f;

__device__ __forceinline__ float fast_exp2(float x) {  // ex2.approx: 2 ulp, flushes tiny results to 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Params {
  CUtensorMap q_hi, q_lo, k_hi, k_lo, vt_hi, vt_lo;
  const uint32_t* mask_bits; int mask_words;
  const int* rowany;
  float* out;      // splits == 1
  float* opart;    // [S,B,Nq,256]
  float2* ml;      // [S,B,NH,Nq]
  int B, Nq, Nk, splits, tiles_per_split;
  int k_col0, vt_img_rows, vt_row0;
  int passes;  // 3 = 3xTF32 (fp32 parity); 1 = single-pass TF32 (PN_OPT_SINGLE_PASS): hi operands only
};

__global__ void __launch_bounds__(NUM_THREADS, 1) fa_umma_kernel(const __grid_constant__ Params prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_hi_s = smem;
  uint8_t* q_lo_s = smem + TILE16K;
  uint8_t* stage0 = smem + 2 * TILE16K;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + KV_STAGES * STAGE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int split = blockIdx.x, h = blockIdx.y;
  const int qtiles = (prm.Nq + TQ - 1) / TQ;
  const int b = blockIdx.z / qtiles, qt = blockIdx.z % qtiles;
  const int total_tiles = (prm.Nk + TKEYS - 1) / TKEYS;
  const int tile0 = split * prm.tiles_per_split;
  const int ntiles = min(prm.tiles_per_split, total_tiles - tile0);

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = uniform_u32(*tmem_slot);

  if (warp == 0) {
    // warp-uniform loop, one elected lane issues the TMA (see elect_one)
    const bool three = prm.passes == 3;
    if (elect_one()) {
      mbar_expect_tx(q_full, three ? 2 * TILE16K : TILE16K);
      tma_load_2d(q_hi_s, &prm.q_hi, q_full, h * HD, b * prm.Nq + qt * TQ);
      if (three) tma_load_2d(q_lo_s, &prm.q_lo, q_full, h * HD, b * prm.Nq + qt * TQ);
    }
    __syncwarp();
    for (int t = 0; t < ntiles; ++t) {
      const int s = t & 1;
      const uint32_t ph = (t >> 1) & 1;
      mbar_wait(&kv_empty[s], ph ^ 1);
      uint8_t* st = stage0 + (size_t)s * STAGE_BYTES;
      const int key0 = (tile0 + t) * TKEYS;
      if (elect_one()) {
        mbar_expect_tx(&kv_full[s], three ? STAGE_BYTES : STAGE_BYTES / 2);
        tma_load_2d(st, &prm.k_hi, &kv_full[s], prm.k_col0 + h * HD, b * prm.Nk + key0);
        if (three) tma_load_2d(st + TILE16K, &prm.k_lo, &kv_full[s], prm.k_col0 + h * HD, b * prm.Nk + key0);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          tma_load_2d(st + 2 * TILE16K + a * VT_ATOM, &prm.vt_hi, &kv_full[s], key0 + a * 32, b * prm.vt_img_rows + prm.vt_row0 + h * HD);
          if (three)
            tma_load_2d(st + 3 * TILE16K + a * VT_ATOM, &prm.vt_lo, &kv_full[s], key0 + a * 32, b * prm.vt_img_rows + prm.vt_row0 + h * HD);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = make_idesc(TKEYS);  // M=128, N=128 keys
    const uint32_t idesc_o = make_idesc(HD);     // M=128, N=32 dims
    mbar_wait(q_full, 0);
    const uint64_t dq_hi = make_smem_desc(smem_u32(q_hi_s)), dq_lo = make_smem_desc(smem_u32(q_lo_s));
    for (int t = 0; t < ntiles; ++t) {
      const int s = t & 1;
      const uint32_t ph = (t >> 1) & 1;
      mbar_wait(&kv_full[s], ph);
      tc_fence_after();
      const uint32_t st = smem_u32(stage0 + (size_t)s * STAGE_BYTES);
      const uint64_t dk_hi = make_smem_desc(st), dk_lo = make_smem_desc(st + TILE16K);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / UMMA_K; ++k) {
          const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
          if (prm.passes == 3) {
            umma_tf32(tmem + TM_S, dq_lo + koff, dk_hi + koff, idesc_s, k == 0 ? 0u : 1u);
            umma_tf32(tmem + TM_S, dq_hi + koff, dk_lo + koff, idesc_s, 1u);
            umma_tf32(tmem + TM_S, dq_hi + koff, dk_hi + koff, idesc_s, 1u);
          } else {
            umma_tf32(tmem + TM_S, dq_hi + koff, dk_hi + koff, idesc_s, k == 0 ? 0u : 1u);
          }
        }
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_ready, t & 1);  // softmax warps have consumed S and written P (hi, lo) to TMEM
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t col = (uint32_t)(a * 32 + k * UMMA_K);
            const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
            const uint64_t dv_hi = make_smem_desc(st + 2 * TILE16K + a * VT_ATOM) + koff;
            const uint64_t dv_lo = make_smem_desc(st + 3 * TILE16K + a * VT_ATOM) + koff;
            if (prm.passes == 3) {
              umma_tf32_ts(tmem + TM_O, tmem + TM_PLO + col, dv_hi, idesc_o, (a == 0 && k == 0) ? 0u : 1u);
              umma_tf32_ts(tmem + TM_O, tmem + TM_PHI + col, dv_lo, idesc_o, 1u);
              umma_tf32_ts(tmem + TM_O, tmem + TM_PHI + col, dv_hi, idesc_o, 1u);
            } else {
              umma_tf32_ts(tmem + TM_O, tmem + TM_PHI + col, dv_hi, idesc_o, (a == 0 && k == 0) ? 0u : 1u);
            }
          }
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[s]);
      }
      __syncwarp();
    }
  } else {
    // ===== softmax / accumulate: one query row per thread
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int qi = qt * TQ + r;
    const bool qvalid = qi < prm.Nq;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const bool use_mask =
        prm.mask_bits != nullptr && qvalid && (prm.rowany == nullptr || prm.rowany[b * prm.Nq + qi] != 0);
    const uint32_t* mrow = use_mask ? prm.mask_bits + ((size_t)b * prm.Nq + qi) * prm.mask_words : nullptr;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    float m_run = NEG_BIG, l_run = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int key0 = (tile0 + t) * TKEYS;
      const int nvalid = min(TKEYS, prm.Nk - key0);
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (use_mask) {
        const int wi = key0 >> 5;
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = (wi + i < prm.mask_words) ? __ldg(mrow + wi + i) : 0xffffffffu;
      }
      // keys past the end of the level behave like blocked keys
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rem = nvalid - i * 32;
        if (rem <= 0) w[i] = 0xffffffffu;
        else if (rem < 32) w[i] |= 0xffffffffu << rem;
      }
      mbar_wait(s_full, t & 1);
      tc_fence_after();
      // pass 1: row max over the open keys
      float cmax = NEG_BIG;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem + lane_addr + TM_S + c * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (!((w[c] >> j) & 1u)) cmax = fmaxf(cmax, __uint_as_float(v[j]));
      }
      const float m_new = fmaxf(m_run, cmax);
      const float corr = fast_exp2(m_run - m_new);
      // pass 2: p = 2^(s - m_new), split hi/lo, store to TMEM as the A operand of P V
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32], ph[32], pl[32];
        tmem_ld_32x32b_x32(tmem + lane_addr + TM_S + c * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float p = ((w[c] >> j) & 1u) ? 0.f : fast_exp2(__uint_as_float(v[j]) - m_new);
          lsum += p;
          // truncation-hi split (see pair_topk.cu): the tensor pipe ignores the low 13 mantissa bits, so p itself is the
          // hi operand; lo = p - trunc(p), nudged by half a TF32 ulp so that the hardware truncation rounds it to
          // nearest.  3 instructions per element instead of ~18 for two cvt.rna.tf32 (this loop is the kernel's
          // critical path: the ncu role profile shows the softmax warps 88 % busy, MMA and TMA waiting on them)
          ph[j] = __float_as_uint(p);
          pl[j] = __float_as_uint(p - __uint_as_float(__float_as_uint(p) & 0xffffe000u)) + 0x1000u;
        }
        tmem_st_32x32b_x32(tmem + lane_addr + TM_PHI + c * 32, ph);
        if (prm.passes == 3) tmem_st_32x32b_x32(tmem + lane_addr + TM_PLO + c * 32, pl);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      l_run = l_run * corr + lsum;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] *= corr;
      m_run = m_new;
      mbar_wait(o_full, t & 1);
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem + lane_addr + TM_O, v);
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] += __uint_as_float(v[d]);
      }
    }
    tc_fence_before();
    if (qvalid) {
      if (prm.splits == 1) {
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        float4* op = reinterpret_cast<float4*>(prm.out + ((size_t)b * prm.Nq + qi) * D + h * HD);
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4)
          op[d4] = make_float4(o[d4 * 4] * inv, o[d4 * 4 + 1] * inv, o[d4 * 4 + 2] * inv, o[d4 * 4 + 3] * inv);
      } else {
        float4* op =
            reinterpret_cast<float4*>(prm.opart + (((size_t)split * prm.B + b) * prm.Nq + qi) * D + h * HD);
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4) op[d4] = make_float4(o[d4 * 4], o[d4 * 4 + 1], o[d4 * 4 + 2], o[d4 * 4 + 3]);
        prm.ml[(((size_t)split * prm.B + b) * NH + h) * prm.Nq + qi] = make_float2(m_run, l_run);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace fa

// host launcher ---------------------------------------------------------------------------------------------------
size_t fa_workspace_bytes(int B, int Nq, int Nk) {
  // same layout as the FFMA kernel's split-KV partials; splits <= 64
  const int S = 64;
  size_t o = ((size_t)S * B * Nq * D * sizeof(float) + 255) & ~size_t(255);
  size_t m = ((size_t)S * B * NH * Nq * sizeof(float2) + 255) & ~size_t(255);
  (void)Nk;
  return o + m;
}

int launch_fa_umma(const FaArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace fa;
  PN_REQUIRE(a.q_hi && a.q_lo && a.k_hi && a.k_lo && a.vt_hi && a.vt_lo && a.out, PN_ERR_BAD_ARG, "fa: null pointer");
  PN_REQUIRE(a.B > 0 && a.Nq > 0 && a.Nk > 0 && a.ldv % 4 == 0 && a.ldv >= a.Nk, PN_ERR_BAD_ARG, "fa: bad shape");
  PN_REQUIRE(!a.mask_bits || a.mask_words * 32 >= a.Nk, PN_ERR_BAD_ARG, "fa: mask_words too small");
  Params prm{};
  PN_TRY(umma::make_tmap_2d(&prm.q_hi, a.q_hi, (long long)a.B * a.Nq, D, D, 32, TQ));
  PN_TRY(umma::make_tmap_2d(&prm.q_lo, a.q_lo, (long long)a.B * a.Nq, D, D, 32, TQ));
  const int ldk = a.ldk > 0 ? a.ldk : D;
  const int vt_img_rows = a.vt_img_rows > 0 ? a.vt_img_rows : D;
  PN_REQUIRE(ldk % 4 == 0 && a.k_col0 % 32 == 0 && a.k_col0 + D <= ldk && a.vt_row0 + D <= vt_img_rows, PN_ERR_BAD_ARG,
             "fa: bad K / V^T view");
  PN_TRY(umma::make_tmap_2d(&prm.k_hi, a.k_hi, (long long)a.B * a.Nk, ldk, ldk, 32, TKEYS));
  PN_TRY(umma::make_tmap_2d(&prm.k_lo, a.k_lo, (long long)a.B * a.Nk, ldk, ldk, 32, TKEYS));
  PN_TRY(umma::make_tmap_2d(&prm.vt_hi, a.vt_hi, (long long)a.B * vt_img_rows, a.Nk, a.ldv, 32, HD));
  PN_TRY(umma::make_tmap_2d(&prm.vt_lo, a.vt_lo, (long long)a.B * vt_img_rows, a.Nk, a.ldv, 32, HD));
  prm.k_col0 = a.k_col0; prm.vt_img_rows = vt_img_rows; prm.vt_row0 = a.vt_row0;
  prm.mask_bits = a.mask_bits; prm.mask_words = a.mask_words; prm.rowany = a.rowany;
  prm.out = a.out; prm.B = a.B; prm.Nq = a.Nq; prm.Nk = a.Nk;
  prm.passes = get_option(OPT_SINGLE_PASS) ? 1 : 3;
  const int qtiles = cdiv(a.Nq, TQ);
  const int base = a.B * NH * qtiles;
  const int total_tiles = cdiv(a.Nk, TKEYS);
  // one CTA per SM (TMEM: 512 columns; smem: 160 KB): keep the grid within ONE wave -- B*8*qtiles*splits <= #SMs --
  // a second, nearly empty wave doubles the kernel time (measured: 160 CTAs 103 us -> 144 CTAs 56 us at hw = 16 700)
  const int num_sms = sm_count();
  int want = num_sms / base;
  want = want < 1 ? 1 : (want > 64 ? 64 : want);
  want = want > total_tiles ? total_tiles : want;
  prm.tiles_per_split = cdiv(total_tiles, want);
  prm.splits = cdiv(total_tiles, prm.tiles_per_split);
  if (prm.splits > 1) {
    PN_REQUIRE(ws && ws_bytes >= fa_workspace_bytes(a.B, a.Nq, a.Nk), PN_ERR_WORKSPACE, "fa: workspace too small");
    size_t o = ((size_t)prm.splits * a.B * a.Nq * D * sizeof(float) + 255) & ~size_t(255);
    prm.opart = reinterpret_cast<float*>(ws);
    prm.ml = reinterpret_cast<float2*>(reinterpret_cast<char*>(ws) + o);
  }
  static bool attr_done[PN_MAX_DEVICES] = {false};  // the attribute is per device
  bool& attr_set = attr_done[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fa_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    PN_REQUIRE(e == cudaSuccess, (int)e, "fa: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid(prm.splits, NH, a.B * qtiles);
  fa_umma_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(prm);
  PN_TRY(check_launch("fa_umma_kernel"));
  if (prm.splits > 1) PN_TRY(launch_mha_combine(prm.opart, prm.ml, a.out, a.B, a.Nq, prm.splits, st));
  return 0;
}

}  // namespace pn
