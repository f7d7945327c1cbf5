// Streamed skinny tensor-core products of the DFMF iteration:
//
//   kTransX = false :  C[m, q] = sum_c X[m, c] * Gs[c, t*kp + q]   summed over the split terms t
//                      (A_ij = R_ij * G_j       -- reference _dfmf.py:237,254, regrouped F6)
//   kTransX = true  :  C[m, q] = sum_r X[r, m] * Gs[r, t*kp + q]
//                      (B_ij = R_ij^T * G_i     -- reference _dfmf.py:266)
//
// X is a relation matrix stored row-major in bf16.  Gs is the factor of the *other* type in its
// tensor-core operand form: row-major [n][N] bf16, N = terms*kp, column block t holding the t-th
// term of the bf16 split G = G^(0) + G^(1) + ...  (each term the bf16 rounding of the residual).
// The terms are multiplied in one UMMA of width N and summed in the epilogue, so C carries fp32
// accumulation of an (almost) fp32 factor although the tensor cores only see bf16.
//
// Pipeline (one CTA = one 128-row block of C, optional split along the reduction axis):
//   warp 0      : TMA producer   (cp.async.bulk.tensor, 128B swizzle, mbarrier expect_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer, tcgen05.commit -> mbarriers
//   warps 2..5  : epilogue       (tcgen05.ld 32x32b, sum the split terms, store / red.add to C)
// The X tile is the MMA "A" operand: K-major when kTransX == false, MN-major (transposed view of
// the same bytes TMA wrote) when kTransX == true.  Gs is always the MN-major "B" operand.
#pragma once
#include "sm100_ptx.cuh"

namespace fz {

struct SkinnyParams {
  float* C;            // output base
  long long ldc;       // elements between consecutive rows of C
  int M;               // rows of C  (rows of X, or columns of X when transposed)
  int K;               // reduction length (columns of X, or rows of X when transposed)
  int k;               // valid output columns (<= kp)
  int kp;              // padded width of one split term (64 here)
  int terms;           // number of split terms (N = terms * kp)
  int k_per_split;     // reduction elements handled per blockIdx.y (multiple of 64)
  int atomic;          // 1: red.add into C (split-K), 0: plain store
  int g_row0;          // first row of Gs that pairs with reduction index 0 (row-sharded factors)
  int g_col0 = 0;      // first column of Gs read as term 0 (64 = the residual term of a two-term operand: error probes;
                       // c * 64 = the c-th 64-column block of a factor of rank > 64)
  int g_term_stride = 64;   // columns of Gs between consecutive terms of the operand (the padded rank of the factor)
  int chunks_per_term = 1;  // 64-column chunks of Gs per term in this launch (kp / 64): 2 = a 128-column block of a wide factor
};

constexpr int kSkBM = 128;   // rows of C per CTA == UMMA M
constexpr int kSkBK = 64;    // reduction elements per pipeline stage (one 128B swizzle row of bf16)
constexpr int kSkThreads = 192;

template <int N>
struct SkinnyCfg {
  static constexpr int kStageX = kSkBM * kSkBK * 2;          // 16 KB
  static constexpr int kStageG = (N / 64) * (64 * 64 * 2);   // N/64 boxes of 8 KB
  static constexpr int kStageBytes = kStageX + kStageG;
  static constexpr int kStages = (N <= 128) ? 6 : 4;
  static constexpr int kTmemCols = (N <= 64) ? 64 : (N <= 128 ? 128 : 256);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int N, bool kTransX>
__global__ void __launch_bounds__(kSkThreads, 1)
umma_skinny_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                   const SkinnyParams p) {
  using Cfg = SkinnyCfg<N>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms are 1024 B: align the stage ring.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kSkBM;
  const int k_begin = blockIdx.y * p.k_per_split;
  const int k_end = min(p.K, k_begin + p.k_per_split);
  const int num_kb = (k_end - k_begin + kSkBK - 1) / kSkBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmG);
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Both issuing roles run as WHOLE, converged warps with one elected lane per step: under a divergent `if (lane == 0)` ptxas
  // wraps every TMA / tcgen05 instruction in a ~100-cycle waterfall loop (csrc/dev/mma_pace.cu).
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % Cfg::kStages;
      const uint32_t ph = (kb / Cfg::kStages) & 1;
      ptx::mbar_wait(&empty_bar[s], ph ^ 1);
      if (ptx::elect_one()) {
        uint8_t* xs = smem + s * Cfg::kStageBytes;
        uint8_t* gs = xs + Cfg::kStageX;
        const int k0 = k_begin + kb * kSkBK;
        ptx::mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
        if (!kTransX) {
          // box {64 cols (K), 128 rows (M)}
          ptx::tma_load_2d(xs, &tmX, &full_bar[s], k0, m0, ptx::kEvictFirst);
        } else {
          // two boxes {64 cols (M), 64 rows (K)}
          ptx::tma_load_2d(xs, &tmX, &full_bar[s], m0, k0, ptx::kEvictFirst);
          ptx::tma_load_2d(xs + 8192, &tmX, &full_bar[s], m0 + 64, k0, ptx::kEvictFirst);
        }
#pragma unroll
        for (int ch = 0; ch < N / 64; ++ch)
          ptx::tma_load_2d(gs + ch * 8192, &tmG, &full_bar[s],
                           p.g_col0 + (ch / p.chunks_per_term) * p.g_term_stride + (ch % p.chunks_per_term) * 64, p.g_row0 + k0,
                           ptx::kEvictLast);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::idesc_bf16_f32(kSkBM, N, kTransX, true);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % Cfg::kStages;
      const uint32_t ph = (kb / Cfg::kStages) & 1;
      ptx::mbar_wait(&full_bar[s], ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t xs = ptx::smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t gs = xs + Cfg::kStageX;
#pragma unroll
        for (int ks = 0; ks < kSkBK / 16; ++ks) {
          uint64_t adesc, bdesc;
          if (!kTransX) adesc = ptx::smem_desc_sw128(xs + ks * 32, 16, 1024);          // K-major
          else          adesc = ptx::smem_desc_sw128(xs + ks * 2048, 8192, 1024);      // MN-major
          bdesc = ptx::smem_desc_sw128(gs + ks * 2048, 8192, 1024);                    // MN-major
          ptx::umma_bf16(tmem_base, adesc, bdesc, idesc, (kb | ks) != 0);
        }
        ptx::umma_commit(&empty_bar[s]);  // smem slot reusable once these MMAs retire
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::umma_commit(tmem_full_bar);    // accumulator complete
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
    const int row = m0 + quarter * 32 + lane;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    float* crow = nullptr;
    if (row < p.M) crow = p.C + (long long)row * p.ldc;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    for (int q0 = 0; q0 < p.kp; q0 += 32) {
      float acc[32], v[32];
      if (num_kb > 0) {
        ptx::tmem_ld32(lane_addr + q0, acc);
        ptx::tmem_ld_wait();
        for (int t = 1; t < p.terms; ++t) {
          ptx::tmem_ld32(lane_addr + t * p.kp + q0, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] += v[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
      }
      if (crow != nullptr) {
        if (p.atomic) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (q0 + i < p.k) atomicAdd(crow + q0 + i, acc[i]);
        } else if (((p.ldc & 3) == 0) && (q0 + 32 <= p.k) &&
                   ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(crow + q0 + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (q0 + i < p.k) crow[q0 + i] = acc[i];
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace fz
