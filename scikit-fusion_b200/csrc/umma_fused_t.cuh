// Fused single-pass streamed products, TRANSPOSED accumulator form ("v4"): same contract as umma_fused.cuh
//        A_ij = R_ij G_j ,   B_ij = R_ij^T G_i          from ONE stream of the bf16 relation
// but both products are computed with the FACTOR as the M operand (M = 128 = 2 split terms x 64 latent columns) and
// the RELATION as the N operand, so that per 32 KB of relation the shared-memory port moves 160 KB instead of the
// 208 KB of the v3 kernel (v3 is bound by the 128 B/clk shared-memory port, not by HBM):
//
//   A^T-product (SS)  D_A[128 x 256 rows]  += GsT_j[128 x 16 cols] * R[256 rows x 16 cols]^T      M128 N256 K16
//                     both operands K-major; the factor slice (4 KB) is read once per 256 relation rows
//   B^T-product (TS)  D_B[128 x 64 cols]   += GsT_i[128 x 16 rows] * R[16 rows x 64 cols]         M128 N64  K16
//                     the A operand GsT_i (the CTA's 256 rows of the transposed factor) is RESIDENT IN TMEM, so the
//                     only shared-memory read is the 2 KB relation slice (MN-major B operand: the same bytes the
//                     A^T-product reads K-major)
//
// CTA = 256 rows of R x a range of 64-column chunks.  Per chunk (32 KB of R + 16 KB of GsT_j in one ring stage):
//   TMA   : R[r0 .. r0+255, c0 .. c0+63]  (box {64,256}, 128B swizzle)   +   GsT_j[0..127, c0 .. c0+63] (box {64,128})
//   MMA   : 16 x TS UMMA into D_B[chunk & 1]  -> commit bacc_full ;  4 x SS UMMA into D_A -> commit stage empty
//   epilog: 4 warps drain D_B[chunk & 1] (64 columns), add the hi/lo split terms (they sit 16 lanes apart inside each
//           warp's TMEM quarter, see row order below, so one shuffle per pair of values), stage 64 x 16 fp32 per warp
//           and hand it to the TMA unit as cp.reduce.async.bulk.tensor .add into B (reduction happens in L2)
//           while the tensor pipe works on the A^T-product of this chunk and the B^T-product of the next one.
// TMEM (512 columns): D_A 0..255 | GsT_i (bf16 pairs, 256 rows) 256..383 | D_B[0] 384..447 | D_B[1] 448..511.
// SMEM (210 KB)     : 4 stages x (32 KB R + 16 KB GsT_j) | 4 x 4 KB flush staging.
//
// Row order of the transposed operand form GsT_t [128][ldt] (bf16; built by split_factor_t):
//   row m = 32 * (k / 16) + 16 * term + (k % 16)
// i.e. TMEM lane 32 q + t holds split term t / 16 of latent column 16 q + t % 16: the two terms of a column live in the
// same warp's lane quarter, 16 lanes apart.
#pragma once
#include "sm100_ptx.cuh"

namespace fz {

struct FusedTParams {
  float* A;                       // [n_rows][lda]   (+)= R Gs_j
  float* B;                       // [n_cols][ldb]   += R^T Gs_i       (always reduced into; caller zeroes B)
  long long lda, ldb;
  const __nv_bfloat16* GiT;       // transposed operand form of type i  [128][ldt]
  long long ldt;
  int n_rows;                     // local rows of R (rows of A)
  int n_cols;                     // columns of R (rows of B)
  int k_a;                        // valid columns of A  (rank of type j)
  int k_b;                        // valid columns of B  (rank of type i)
  int gi_row0;                    // column of GiT that pairs with local row 0 of R (row-sharded factors)
  int tiles_per_split;            // 128-column tiles handled per blockIdx.y (same unit as the v3 kernel)
  int a_atomic;                   // 1: several column splits add into A (caller zeroes A), 0: plain store
  int tma_flush;                  // B partials: 2 = ONE TMA reduce-add of 64 rows x 256 B per chunk and CTA (tmB box {64, 64}),
                                  // 1 = one of 64 rows x 64 B per warp (tmB box {16, 64}), 0 = red.global fallback
  int variant;                    // developer probe: bit0 = swap the bf16 halves of the TMEM A-operand words,
                                  //                  bit1 = skip B^T-product, bit2 = skip A^T-product, bit3 = skip flush,
                                  //                  bit4 = staggered sweep, bit5 = no GsT_j reloads (wrong results: L2->SM traffic study)
};

constexpr int kFtThreads = 192;   // warp 0: TMA producer | 1: MMA issuer | 2..5: epilogue
constexpr int kFtRows = 256;      // relation rows per CTA
constexpr int kFtChunk = 64;      // relation columns per ring stage
constexpr int kFtStages = 4;
constexpr int kFtRBytes = kFtRows * kFtChunk * 2;         // 32 KB
constexpr int kFtGBytes = 128 * kFtChunk * 2;             // 16 KB
constexpr int kFtStageBytes = kFtRBytes + kFtGBytes;      // 48 KB
constexpr int kFtFlushBytes = 2 * 16384;                  // two staging buffers of 64 rows x 64 fp32 (flush mode 2), or 4 warps x 4 KB (mode 1)
constexpr int kFtSmemBytes = kFtStages * kFtStageBytes + kFtFlushBytes + 1024 + 256;

namespace ptx {
// D[tmem] (+)= A[tmem] * B[smem desc]   (A operand resident in tensor memory: lane = row, two bf16 per 32-bit column)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns <- 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0],"
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16,"
      " %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
}  // namespace ptx

__global__ void __launch_bounds__(kFtThreads, 1)
umma_fused_t_kernel(const __grid_constant__ CUtensorMap tmR,    // relation, bf16, box {64 cols, 256 rows}, 128B swizzle
                    const __grid_constant__ CUtensorMap tmGjT,  // GsT_j [128][ldt], bf16, box {64 cols, 128 rows}
                    const __grid_constant__ CUtensorMap tmB,    // B, fp32, box {16 cols, 64 rows}, no swizzle
                    const FusedTParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* st_base = smem;                                        // 4 x 48 KB
  uint8_t* fl_st = st_base + kFtStages * kFtStageBytes;           // 4 x 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(fl_st + kFtFlushBytes);
  uint64_t* full = bars;                        // [4]  TMA -> MMA
  uint64_t* empty = full + kFtStages;           // [4]  MMA -> TMA
  uint64_t* bacc_full = empty + kFtStages;      // [2]  MMA -> epilogue
  uint64_t* bacc_empty = bacc_full + 2;         // [2]  epilogue -> MMA
  uint64_t* git_ready = bacc_empty + 2;         // [1]  epilogue -> MMA (GsT_i resident in TMEM)
  uint64_t* aacc_full = git_ready + 1;          // [1]  MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aacc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * kFtRows;
  const int total_chunks = (p.n_cols + kFtChunk - 1) / kFtChunk;
  const int chunk_begin = blockIdx.y * p.tiles_per_split * 2;
  const int chunk_end = min(total_chunks, chunk_begin + p.tiles_per_split * 2);
  const int n_chunks = max(0, chunk_end - chunk_begin);
  // Optional staggered sweep (variant bit4): CTA x starts its column sweep at a different chunk and wraps around, so that
  // at any moment the CTAs reduce their B partials into different rows of B.  +1.6 % in a burst, but the GsT_j tiles are
  // then no longer shared in L2 at the same time (DRAM traffic 1.09x instead of 1.02x) and under the power cap it is
  // 2-3 % slower (profiles/r01b_sustained_v4_flush_study.log): off by default.
  const int shift = !(p.variant & 16) || n_chunks == 0 ? 0 : (int)(((long long)blockIdx.x * n_chunks) / gridDim.x);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmR);
    ptx::prefetch_tmap(&tmGjT);
    if (p.tma_flush) ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < kFtStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&bacc_full[s], 1);
      ptx::mbar_init(&bacc_empty[s], 128);      // every epilogue thread arrives
    }
    ptx::mbar_init(git_ready, 128);
    ptx::mbar_init(aacc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColGiT = 256, kColBacc = 384;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (whole warp, one elected lane issues)
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c % kFtStages;
      const int cc = c + shift < n_chunks ? c + shift : c + shift - n_chunks;
      const int col0 = (chunk_begin + cc) * kFtChunk;
      ptx::mbar_wait_wd(&empty[s], ((c / kFtStages) & 1) ^ 1);
      if (ptx::elect_one()) {
        const bool load_g = !(p.variant & 32) || c < kFtStages;     // probe bit5: factor chunks loaded once, then reused
        ptx::mbar_expect_tx(&full[s], load_g ? kFtStageBytes : kFtRBytes);
        uint8_t* dst = st_base + s * kFtStageBytes;
        ptx::tma_load_2d(dst, &tmR, &full[s], col0, r0, ptx::kEvictFirst);
        if (load_g) ptx::tma_load_2d(dst + kFtRBytes, &tmGjT, &full[s], col0, 0, ptx::kEvictLast);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    // The whole warp runs this loop convergently and ONE elected lane issues (elect.sync): with warp-uniform control
    // flow ptxas keeps descriptors and TMEM addresses in uniform registers and emits bare UTCHMMA instructions.
    // Under a divergent `if (lane == 0)` it wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY
    // waterfall loop that costs ~100 cycles per instruction (measured: csrc/dev/mma_pace.cu) -- more than the
    // 32..128 cycles the tensor pipe needs for the instruction itself.
    if (n_chunks > 0) {
      constexpr uint32_t idesc_a = ptx::idesc_bf16_f32(128, 256, false, false);  // GsT_j K-major x R K-major
      constexpr uint32_t idesc_b = ptx::idesc_bf16_f32(128, 64, false, true);    // GsT_i (TMEM) x R MN-major
      const bool do_b = !(p.variant & 2), do_a = !(p.variant & 4);
      const uint32_t st0 = ptx::smem_u32(st_base);
      ptx::mbar_wait_wd(git_ready, 0);
      ptx::tc_fence_after();
      for (int c = 0; c < n_chunks; ++c) {
        const int s = c % kFtStages;
        const int h = c & 1;
        const uint32_t rt = st0 + s * kFtStageBytes;
        const uint32_t gt = rt + kFtRBytes;
        ptx::mbar_wait_wd(&full[s], (c / kFtStages) & 1);
        ptx::mbar_wait_wd(&bacc_empty[h], ((c >> 1) & 1) ^ 1);     // epilogue has drained this D_B buffer
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          if (do_b)
#pragma unroll
            for (int ks = 0; ks < 16; ++ks)
              ptx::umma_bf16_ts(tmem_base + kColBacc + h * 64, tmem_base + kColGiT + ks * 8,
                                ptx::smem_desc_sw128(rt + ks * 2048, 32768, 1024), idesc_b, ks != 0);
          ptx::umma_commit(&bacc_full[h]);
          if (do_a)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_bf16(tmem_base, ptx::smem_desc_sw128(gt + ks * 32, 16, 1024),
                             ptx::smem_desc_sw128(rt + ks * 32, 16, 1024), idesc_a, (c | ks) != 0);
          ptx::umma_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (ptx::elect_one()) ptx::umma_commit(aacc_full);
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- epilogue (warps 2..5)
    const int quarter = warp & 3;                                   // TMEM lane quarter this warp may access
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int half = lane >> 4;                                     // 0: hi-term lane, 1: lo-term lane
    const int kcol = quarter * 16 + (lane & 15);                    // latent column of this lane (after the pair sum)
    float* my_stage = reinterpret_cast<float*>(fl_st + quarter * 4096);
    if (n_chunks > 0) {
      // resident A operand of the B^T-product: row m = 32 q + lane of GsT_i, the CTA's 256 relation rows -> 128 columns
      const __nv_bfloat16* src = p.GiT + (long long)(quarter * 32 + lane) * p.ldt + p.gi_row0 + r0;
      const bool vec = ((p.gi_row0 & 7) == 0) && ((p.ldt & 7) == 0);
#pragma unroll 1
      for (int b = 0; b < 4; ++b) {
        uint32_t w[32];
        if (vec) {
          const uint4* s4 = reinterpret_cast<const uint4*>(src + b * 64);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 v = __ldg(s4 + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
          }
        } else {
          const unsigned short* s2 = reinterpret_cast<const unsigned short*>(src + b * 64);
#pragma unroll
          for (int i = 0; i < 32; ++i) w[i] = (uint32_t)__ldg(s2 + 2 * i) | ((uint32_t)__ldg(s2 + 2 * i + 1) << 16);
        }
        if (p.variant & 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) w[i] = (w[i] >> 16) | (w[i] << 16);
        }
        ptx::tmem_st32(lane_addr + kColGiT + b * 32, w);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(git_ready);
    }
    const bool skip_flush = (p.variant & 8) != 0;
    bool staged = false;
    for (int c = 0; c < n_chunks; ++c) {
      const int h = c & 1;
      ptx::mbar_wait_wd(&bacc_full[h], (c >> 1) & 1);
      ptx::tc_fence_after();
      float x[64];
      ptx::tmem_ld32(lane_addr + kColBacc + h * 64, x);
      ptx::tmem_ld32(lane_addr + kColBacc + h * 64 + 32, x + 32);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bacc_empty[h]);                             // D_B[h] may be overwritten
      if (skip_flush || (p.tma_flush != 2 && quarter * 16 >= p.k_b)) continue;   // (mode 2: every warp joins the barriers)
      // pair sum: the hi lane keeps even columns, the lo lane odd ones; y[q] belongs to column 2 q + half
      float y[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const float send = half ? x[2 * q] : x[2 * q + 1];
        const float mine = half ? x[2 * q + 1] : x[2 * q];
        y[q] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
      }
      const int cc = c + shift < n_chunks ? c + shift : c + shift - n_chunks;
      const int col0 = (chunk_begin + cc) * kFtChunk;               // first B row (column of R) of this chunk
      if (p.tma_flush == 2) {
        // CTA-wide staging buffer [64 cols][64 k] fp32 (double-buffered), one 16 KB reduce-add per chunk: 256-byte rows
        // instead of four 64-byte-row operations (fewer TMA / L2 requests per flushed byte)
        float* stage = reinterpret_cast<float*>(fl_st + (c & 1) * 16384);
        if (warp == 2 && ptx::elect_one()) ptx::tma_wait_read<1>();   // the reduce issued two chunks ago has read this buffer
        ptx::named_barrier_sync(1, 128);
#pragma unroll
        for (int q = 0; q < 32; ++q) stage[(2 * q + half) * 64 + kcol] = y[q];
        ptx::fence_proxy_async();
        ptx::named_barrier_sync(2, 128);
        if (warp == 2 && ptx::elect_one()) {
          ptx::tma_reduce_add_2d(&tmB, stage, 0, col0);             // rows / columns beyond the tensor are clipped
          ptx::tma_commit_group();
        }
      } else if (p.tma_flush) {
        if (staged) {                                               // previous reduce has read the staging buffer
          if (ptx::elect_one()) ptx::tma_wait_read_all();           // (elect.sync is deterministic: always the same lane)
          __syncwarp();
        }
        // stage[col][16]: one warp store covers columns 2q, 2q+1 = 128 contiguous bytes
#pragma unroll
        for (int q = 0; q < 32; ++q) my_stage[q * 32 + lane] = y[q];
        ptx::fence_proxy_async();
        __syncwarp();
        if (ptx::elect_one()) {
          ptx::tma_reduce_add_2d(&tmB, my_stage, quarter * 16, col0);   // rows / columns beyond the tensor are clipped
          ptx::tma_commit_group();
        }
        __syncwarp();
        staged = true;
      } else if (kcol < p.k_b) {
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int bcol = col0 + 2 * q + half;
          if (bcol < p.n_cols) atomicAdd(p.B + (long long)bcol * p.ldb + kcol, y[q]);
        }
      }
    }
    // final A accumulators: D_A[m][row] -> A[r0 + row][k]
    if (n_chunks > 0) {
      ptx::mbar_wait_wd(aacc_full, 0);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int b = 0; b < 8; ++b) {
        float x[32];
        ptx::tmem_ld32(lane_addr + b * 32, x);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float send = half ? x[2 * q] : x[2 * q + 1];
          const float mine = half ? x[2 * q + 1] : x[2 * q];
          const float v = mine + __shfl_xor_sync(0xffffffffu, send, 16);
          const int arow = r0 + b * 32 + 2 * q + half;
          if (arow < p.n_rows && kcol < p.k_a) {
            float* out = p.A + (long long)arow * p.lda + kcol;
            if (p.a_atomic) atomicAdd(out, v);
            else *out = v;
          }
        }
      }
    }
    __syncwarp();
    if (p.tma_flush && ptx::elect_one()) ptx::tma_wait_all();       // reductions performed before the CTA retires
    __syncwarp();
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// Transposed operand form of a factor for the kernel above:  GsT[m][r], m = 32 (k/16) + 16 term + k%16, two bf16 split
// terms (hi = bf16(G), lo = bf16(G - hi)).  Block = 64 factor rows; coalesced reads of G, 128-byte row segments out.
// Rows >= n_valid and latent columns >= k are written as zeros.
template <class T>
__global__ void __launch_bounds__(256)
split_factor_t(const T* __restrict__ G, long long ldg, __nv_bfloat16* __restrict__ GsT, long long ldt, long long n_valid,
               long long n_rows, int k) {
  __shared__ uint16_t tile[128][66];
  const long long row0 = (long long)blockIdx.x * 64;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int r = i >> 6, q = i & 63;
    float v = 0.f;
    if (row0 + r < n_valid && q < k) v = (float)G[(row0 + r) * ldg + q];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const int m = 32 * (q >> 4) + (q & 15);
    tile[m][r] = __bfloat16_as_ushort(hi);
    tile[m + 16][r] = __bfloat16_as_ushort(lo);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * 32; i += 256) {
    const int m = i >> 5, pr = i & 31;
    if (row0 + 2 * pr < n_rows) {
      const uint32_t w = (uint32_t)tile[m][2 * pr] | ((uint32_t)tile[m][2 * pr + 1] << 16);
      *reinterpret_cast<uint32_t*>(GsT + (long long)m * ldt + row0 + 2 * pr) = w;
    }
  }
}

}  // namespace fz
