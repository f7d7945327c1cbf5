// Fused single-pass streamed products: ONE read of a bf16 relation tile feeds both
//        A_ij = R_ij G_j          (rows of type i;   accumulators persistent in TMEM)
//        B_ij = R_ij^T G_i        (rows of type j;   accumulated over the CTA's row pair, then reduced into B)
// so the relation is streamed from HBM once per iteration instead of once per product.
//
// CTA = a PAIR of 128-row blocks (256 rows of R) x a range of 128-column tiles.  Per column tile c:
//   TMA   : R[r0+128t .. , c] for t = 0,1 (32 KB each, 128B swizzle; warp 0) and the two 64-row halves of the Gs_j
//           tile (warp 6, a separate producer so factor operands are prefetched independently of relation stages)
//   MMA   : A_acc[t] += R_tile (K-major A)  * Gs_j[c]   (MN-major B)     8 x UMMA 128x128x16 per row block
//           B_acc[c&1] (+)= R_tile^T (MN-major A view of the same bytes) * Gs_i[t] (resident, MN-major B)
//           issue order  A(t0) B(t0) | A(t1) B(t1): a relation stage is held for just its own 16 MMAs
//   epilog: once both row blocks of tile c are multiplied, 4 warps drain B_acc[c&1] from TMEM, add the split terms,
//           stage the 128 x k fp32 partial in shared memory (swizzled) and hand it to the TMA unit as
//           cp.reduce.async.bulk.tensor ... .add  -- the reduction happens in L2, the SM's LSU never sees it --
//           while the tensor pipe already works on tile c+1 with the other B_acc buffer.
//           (A red.global fallback exists for ranks that are not a multiple of 4; it is LSU-bound at ~8 B/clk/SM,
//           which capped the first version of this kernel at 54 % of the HBM roofline, profiles/r01_fused_*.)
// TMEM (512 columns): A_acc[0] 0..127 | A_acc[1] 128..255 | B_acc[0] 256..383 | B_acc[1] 384..511.
// SMEM (224 KB)     : 3 R stages x 32 KB | ring of 3 Gs_j halves x 16 KB | 16 KB flush staging | 2 resident Gs_i x 32 KB.
// Requires N = terms*64 == 128 (two split terms); other term counts use the two-pass kernels.
#pragma once
#include "sm100_ptx.cuh"

namespace fz {

struct FusedParams {
  float* A;            // [M_rows][lda]  (+)= R Gs_j
  float* B;            // [N_cols][ldb]  += R^T Gs_i       (always reduced into; caller zeroes B)
  long long lda, ldb;
  int n_rows;          // local rows of R (rows of A)
  int n_cols;          // columns of R (rows of B)
  int k_a;             // valid columns of A  (rank of type j)
  int k_b;             // valid columns of B  (rank of type i)
  int gi_row0;         // row of Gs_i that pairs with local row 0 of R (row-sharded factors)
  int tiles_per_split; // column tiles handled per blockIdx.y
  int a_atomic;        // 1: several column splits add into A (caller zeroes A), 0: plain store
  int tma_flush;       // 1: B partials go out as TMA reduce-add (needs tmB); 0: red.global fallback
  int b_terms;         // split terms multiplied in the B-product: 2 (default) or 1 = first term only (N = 64; validated, unused by
                       // the engine: see DESIGN.md section 4).  An fp16 first term is NOT possible: kind::f16 with a bf16 A operand
                       // and an fp16 B operand raises an illegal-instruction exception on sm_100a (profiles/r01b_mixed_format_probe.log)
  // PAIR mode (two restarts batched into one pass over R; reference fan-out over n_run: dfmf.py:87-95):
  // the two 64-column halves of the operands are the single-term forms of RUN 0 and RUN 1 instead of the two split terms of
  // one factor, and the halves of every accumulator go to separate outputs (A / A2, tmB / tmB2, cj / cj2) instead of being added.
  int pair = 0;
  float* A2 = nullptr;
  float* B2 = nullptr;             // red.global fallback target of run 1
  const float* cj2 = nullptr;
  const float* rowsum = nullptr;   // mean-centred operand form (G = 1 c^T + D, the split terms represent D): row sums of R, and
  const float* cj = nullptr;       // the centre of the column factor; A += rowsum c_j^T in the epilogue (nullptr: plain form)
  int probe_skip_flush;// developer probe only (wrong results): bit0 = no reductions, bit1 = no B-product MMAs, bit2 = no A-product MMAs,
                       // bit4 = B-product with the first split term only (N = 64), bit5 = A-product likewise: what a 1.5- / 1-term
                       // operand form would cost (profiles/r01b_sustained_term_count_study.log)
};

constexpr int kFuThreads = 224;   // warp 0: R producer | 1: MMA | 2..5: epilogue | 6: Gs producer
constexpr int kFuTile = 128;
constexpr int kFuRStages = 3;
constexpr int kFuGjSlots = 3;
constexpr int kFuTileBytes = kFuTile * kFuTile * 2;     // 32 KB: an R tile, or a 128-row x 128-col Gs tile
constexpr int kFuHalfBytes = kFuTileBytes / 2;          // 16 KB: 64 rows x 128 cols of Gs
constexpr int kFuStageBytes = 16384;                    // flush staging: 4 warps x (32 rows x 32 fp32)
constexpr int kFuSmemBytes = kFuRStages * kFuTileBytes + kFuGjSlots * kFuHalfBytes + kFuStageBytes + 2 * kFuTileBytes + 1024 + 256;

__global__ void __launch_bounds__(kFuThreads, 1)
umma_fused_kernel(const __grid_constant__ CUtensorMap tmR,    // relation, bf16, box {64 cols, 128 rows}
                  const __grid_constant__ CUtensorMap tmGj,   // Gs_j,     bf16, box {64 cols, 64 rows}
                  const __grid_constant__ CUtensorMap tmGi,   // Gs_i,     bf16, box {64 cols, 128 rows}
                  const __grid_constant__ CUtensorMap tmB,    // B,        fp32, box {32 cols, 32 rows} (reduce target)
                  const __grid_constant__ CUtensorMap tmB2,   // B of run 1 in pair mode (any valid map otherwise)
                  const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* r_st = smem;                                          // 3 x 32 KB
  uint8_t* gj_st = r_st + kFuRStages * kFuTileBytes;             // 3 x 16 KB
  uint8_t* fl_st = gj_st + kFuGjSlots * kFuHalfBytes;            // 16 KB
  uint8_t* gi_st = fl_st + kFuStageBytes;                        // 2 x 32 KB (resident)
  uint64_t* bars = reinterpret_cast<uint64_t*>(gi_st + 2 * kFuTileBytes);
  uint64_t* r_full = bars;                     // [3]
  uint64_t* r_empty = r_full + kFuRStages;     // [3]
  uint64_t* gj_full = r_empty + kFuRStages;    // [3]
  uint64_t* gj_empty = gj_full + kFuGjSlots;   // [3]
  uint64_t* bacc_full = gj_empty + kFuGjSlots; // [2]
  uint64_t* bacc_empty = bacc_full + 2;        // [2]
  uint64_t* gi_full = bacc_empty + 2;          // [1]
  uint64_t* aacc_full = gi_full + 1;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aacc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 2 * kFuTile;                       // first local row of the pair
  const int total_tiles = (p.n_cols + kFuTile - 1) / kFuTile;
  const int tile_begin = blockIdx.y * p.tiles_per_split;
  const int tile_end = min(total_tiles, tile_begin + p.tiles_per_split);
  const int n_tiles = max(0, tile_end - tile_begin);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmR);
    ptx::prefetch_tmap(&tmGj);
    ptx::prefetch_tmap(&tmGi);
    if (p.tma_flush) ptx::prefetch_tmap(&tmB);
    if (p.tma_flush && p.pair) ptx::prefetch_tmap(&tmB2);
    for (int s = 0; s < kFuRStages; ++s) { ptx::mbar_init(&r_full[s], 1); ptx::mbar_init(&r_empty[s], 1); }
    for (int s = 0; s < kFuGjSlots; ++s) { ptx::mbar_init(&gj_full[s], 1); ptx::mbar_init(&gj_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&bacc_full[s], 1);
      ptx::mbar_init(&bacc_empty[s], 128);   // every epilogue thread arrives
    }
    ptx::mbar_init(gi_full, 1);
    ptx::mbar_init(aacc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer: relation tiles
    // (whole warp, one elected lane issues: under a divergent `if (lane == 0)` ptxas wraps every TMA / tcgen05
    //  instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop of ~100 cycles, csrc/dev/mma_pace.cu)
    int it = 0;
    for (int c = 0; c < n_tiles; ++c) {
      const int col0 = (tile_begin + c) * kFuTile;
      for (int t = 0; t < 2; ++t, ++it) {
        const int s = it % kFuRStages;
        ptx::mbar_wait(&r_empty[s], ((it / kFuRStages) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&r_full[s], kFuTileBytes);
          for (int ch = 0; ch < 2; ++ch)
            ptx::tma_load_2d(r_st + s * kFuTileBytes + ch * 16384, &tmR, &r_full[s], col0 + ch * 64, r0 + t * kFuTile,
                             ptx::kEvictFirst);
        }
        __syncwarp();
      }
    }
  } else if (warp == 6) {
    // ---------------------------------------------------------------- TMA producer: factor operands
    // (own warp so that a Gs_j half is requested the moment its ring slot frees up, ~1.25 tiles ahead of use,
    //  instead of queueing behind relation tiles that wait for a free stage)
    if (n_tiles > 0) {
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(gi_full, 2 * kFuTileBytes);          // resident Gs_i tiles of the two row blocks
        for (int t = 0; t < 2; ++t)
          for (int ch = 0; ch < 2; ++ch)
            ptx::tma_load_2d(gi_st + t * kFuTileBytes + ch * 16384, &tmGi, gi_full, ch * 64, p.gi_row0 + r0 + t * kFuTile,
                             ptx::kEvictLast);
      }
      __syncwarp();
      for (int i = 0; i < 2 * n_tiles; ++i) {
        const int col0 = (tile_begin + (i >> 1)) * kFuTile + (i & 1) * 64;
        const int slot = i % kFuGjSlots;
        ptx::mbar_wait(&gj_empty[slot], ((i / kFuGjSlots) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&gj_full[slot], kFuHalfBytes);
          for (int ch = 0; ch < 2; ++ch)
            ptx::tma_load_2d(gj_st + slot * kFuHalfBytes + ch * 8192, &tmGj, &gj_full[slot], ch * 64, col0, ptx::kEvictLast);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    // (whole warp converged, one elected lane issues -- see the producer note above: this removes ~100 cycles of
    //  waterfall-loop overhead per tcgen05.mma, which at 16 MMAs per tile was the kernel's actual bound)
    if (n_tiles > 0) {
      const uint32_t idesc_a = ptx::idesc_bf16_f32(128, (p.probe_skip_flush & 32) ? 64 : 128, false, true);   // R K-major  x Gs MN-major
      const uint32_t idesc_b = ptx::idesc_bf16_f32(128, ((p.probe_skip_flush & 16) || p.b_terms == 1) ? 64 : 128, true, true);   // R^T MN-major x Gs MN-major
      const bool do_a = !(p.probe_skip_flush & 4), do_b = !(p.probe_skip_flush & 2);
      const uint32_t r_base = ptx::smem_u32(r_st), gj_base = ptx::smem_u32(gj_st);
      const uint32_t gi0 = ptx::smem_u32(gi_st), gi1 = gi0 + kFuTileBytes;
      ptx::mbar_wait(gi_full, 0);
      int it = 0;
      for (int c = 0; c < n_tiles; ++c) {
        const int gs = c & 1;
        const int i0 = 2 * c, i1 = 2 * c + 1;
        const int slot0 = i0 % kFuGjSlots, slot1 = i1 % kFuGjSlots;
        const int s0 = it % kFuRStages, s1 = (it + 1) % kFuRStages;
        const uint32_t ph0 = (it / kFuRStages) & 1, ph1 = ((it + 1) / kFuRStages) & 1;
        const uint32_t rt0 = r_base + s0 * kFuTileBytes, rt1 = r_base + s1 * kFuTileBytes;
        const uint32_t g0 = gj_base + slot0 * kFuHalfBytes, g1 = gj_base + slot1 * kFuHalfBytes;
        const uint32_t bacc = tmem_base + 256 + gs * 128;
        it += 2;
        // ---- row block 0: A-product over both Gs_j halves, then the transposed product; the relation stage is
        //      released as soon as its 16 MMAs retire (short stage hold time keeps two tiles in flight from HBM)
        ptx::mbar_wait(&gj_full[slot0], (i0 / kFuGjSlots) & 1);
        ptx::mbar_wait(&r_full[s0], ph0);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          if (do_a)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_bf16(tmem_base, ptx::smem_desc_sw128(rt0 + ks * 32, 16, 1024),
                             ptx::smem_desc_sw128(g0 + ks * 2048, 8192, 1024), idesc_a, (c | ks) != 0);
        }
        __syncwarp();
        ptx::mbar_wait(&gj_full[slot1], (i1 / kFuGjSlots) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          if (do_a)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_bf16(tmem_base, ptx::smem_desc_sw128(rt0 + 16384 + ks * 32, 16, 1024),
                             ptx::smem_desc_sw128(g1 + ks * 2048, 8192, 1024), idesc_a, 1);
        }
        __syncwarp();
        ptx::mbar_wait(&bacc_empty[gs], ((c >> 1) & 1) ^ 1);       // epilogue has drained this B_acc buffer
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          if (do_b)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              ptx::umma_bf16(bacc, ptx::smem_desc_sw128(rt0 + ks * 2048, 16384, 1024),
                             ptx::smem_desc_sw128(gi0 + ks * 2048, 16384, 1024), idesc_b, ks != 0);
          ptx::umma_commit(&r_empty[s0]);
        }
        __syncwarp();
        // ---- row block 1; each Gs_j half goes back to the ring right after its last use
        ptx::mbar_wait(&r_full[s1], ph1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          if (do_a)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_bf16(tmem_base + 128, ptx::smem_desc_sw128(rt1 + ks * 32, 16, 1024),
                             ptx::smem_desc_sw128(g0 + ks * 2048, 8192, 1024), idesc_a, (c | ks) != 0);
          ptx::umma_commit(&gj_empty[slot0]);
          if (do_a)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_bf16(tmem_base + 128, ptx::smem_desc_sw128(rt1 + 16384 + ks * 32, 16, 1024),
                             ptx::smem_desc_sw128(g1 + ks * 2048, 8192, 1024), idesc_a, 1);
          ptx::umma_commit(&gj_empty[slot1]);
          if (do_b)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              ptx::umma_bf16(bacc, ptx::smem_desc_sw128(rt1 + ks * 2048, 16384, 1024),
                             ptx::smem_desc_sw128(gi1 + ks * 2048, 16384, 1024), idesc_b, 1);
          ptx::umma_commit(&r_empty[s1]);
          ptx::umma_commit(&bacc_full[gs]);
        }
        __syncwarp();
      }
      if (ptx::elect_one()) ptx::umma_commit(aacc_full);
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- epilogue (warps 2..5)
    const int quarter = warp & 3;
    const int lrow = quarter * 32 + lane;                         // TMEM lane = row of the accumulator tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint8_t* my_stage = fl_st + quarter * 4096;                   // 32 rows x 128 B, 128B-swizzled like tmB's box
    const bool skip_red = (p.probe_skip_flush & 1) != 0;
    for (int c = 0; c < n_tiles; ++c) {
      const int gs = c & 1;
      ptx::mbar_wait(&bacc_full[gs], (c >> 1) & 1);
      ptx::tc_fence_after();
      const int brow0 = (tile_begin + c) * kFuTile + quarter * 32;   // first B row (column of R) of this warp
#pragma unroll
      for (int q0 = 0; q0 < 64; q0 += 32) {
        float hi[32], lo[32];
        ptx::tmem_ld32(lane_addr + 256 + gs * 128 + q0, hi);
        ptx::tmem_ld32(lane_addr + 256 + gs * 128 + 64 + q0, lo);
        ptx::tmem_ld_wait();
        if (p.b_terms == 1) {                                      // single-term B-product: the lo half was never written
#pragma unroll
          for (int i = 0; i < 32; ++i) lo[i] = 0.f;
        }
        if (q0 == 32) {             // all TMEM reads of this buffer are done: hand it back to the MMA warp
          ptx::tc_fence_before();
          ptx::mbar_arrive(&bacc_empty[gs]);
        }
        if (q0 >= p.k_b || skip_red) continue;
        if (p.tma_flush && p.pair) {
          // the halves are the partials of two different runs: one reduce each, into each run's own B
          for (int run = 0; run < 2; ++run) {
            const float* src = run == 0 ? hi : lo;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(my_stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_float4(src[4 * j], src[4 * j + 1], src[4 * j + 2], src[4 * j + 3]);
            ptx::fence_proxy_async();
            __syncwarp();
            if (ptx::elect_one()) {
              ptx::tma_reduce_add_2d(run == 0 ? &tmB : &tmB2, my_stage, q0, brow0);
              ptx::tma_commit_group();
              ptx::tma_wait_read_all();
            }
            __syncwarp();
          }
        } else if (p.tma_flush) {
          // row `lane` of the warp's 32 x 32 box: 8 chunks of 16 B, chunk j stored at j ^ (lane & 7)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = make_float4(hi[4 * j] + lo[4 * j], hi[4 * j + 1] + lo[4 * j + 1], hi[4 * j + 2] + lo[4 * j + 2],
                                         hi[4 * j + 3] + lo[4 * j + 3]);
            *reinterpret_cast<float4*>(my_stage + lane * 128 + ((j ^ (lane & 7)) << 4)) = v;
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (ptx::elect_one()) {                                 // deterministic: always the same lane
            ptx::tma_reduce_add_2d(&tmB, my_stage, q0, brow0);    // rows / columns beyond the tensor are clipped
            ptx::tma_commit_group();
            ptx::tma_wait_read_all();                             // staging reusable
          }
          __syncwarp();
        } else if (p.pair) {
          const int bcol = brow0 + lane;
          if (bcol < p.n_cols) {
            float* brow = p.B + (long long)bcol * p.ldb + q0;
            float* brow2 = p.B2 + (long long)bcol * p.ldb + q0;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (q0 + i < p.k_b) { atomicAdd(brow + i, hi[i]); atomicAdd(brow2 + i, lo[i]); }
          }
        } else {
          const int bcol = brow0 + lane;
          if (bcol < p.n_cols) {
            float* brow = p.B + (long long)bcol * p.ldb + q0;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (q0 + i < p.k_b) atomicAdd(brow + i, hi[i] + lo[i]);
          }
        }
      }
    }
    // final A accumulators of the two row blocks
    if (n_tiles > 0) {
      ptx::mbar_wait(aacc_full, 0);
      ptx::tc_fence_after();
      const bool rank1 = (p.rowsum != nullptr) && (blockIdx.y == 0);
      for (int t = 0; t < 2; ++t) {
        const int arow = r0 + t * kFuTile + lrow;
        float* out = (arow < p.n_rows) ? p.A + (long long)arow * p.lda : nullptr;
        const float rs = (rank1 && out != nullptr) ? p.rowsum[arow] : 0.f;
#pragma unroll
        for (int q0 = 0; q0 < 64; q0 += 32) {
          float hi[32], lo[32];
          ptx::tmem_ld32(lane_addr + t * 128 + q0, hi);
          ptx::tmem_ld32(lane_addr + t * 128 + 64 + q0, lo);
          ptx::tmem_ld_wait();
          if (out != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (q0 + i < p.k_a && p.pair) {      // two runs: separate outputs, each with its own centre
                float v = hi[i], w = lo[i];
                if (rank1) { v = fmaf(rs, __ldg(p.cj + q0 + i), v); w = fmaf(rs, __ldg(p.cj2 + q0 + i), w); }
                float* out2 = p.A2 + (long long)arow * p.lda;
                if (p.a_atomic) { atomicAdd(out + q0 + i, v); atomicAdd(out2 + q0 + i, w); }
                else { out[q0 + i] = v; out2[q0 + i] = w; }
              } else if (q0 + i < p.k_a) {
                float v = hi[i] + lo[i];
                if (rank1) v = fmaf(rs, __ldg(p.cj + q0 + i), v);
                if (p.a_atomic) atomicAdd(out + q0 + i, v);
                else out[q0 + i] = v;
              }
            }
          }
        }
      }
    }
    __syncwarp();
    if (p.tma_flush && ptx::elect_one()) ptx::tma_wait_all();     // reductions performed before the CTA retires
    __syncwarp();
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace fz
