// Fused single-pass streamed products: ONE read of a bf16 relation tile feeds both
//        A_ij = R_ij G_j          (rows of type i;   accumulators persistent in TMEM)
//        B_ij = R_ij^T G_i        (rows of type j;   accumulated over the CTA's row pair, flushed with red.add)
// so the relation is streamed from HBM once per iteration instead of once per product.
//
// CTA = a PAIR of 128-row blocks (256 rows of R) x a range of 128-column tiles.  Per column tile c:
//   TMA   : R[r0+128t .. , c] for t = 0,1 (32 KB each, 128B swizzle) and the Gs_j tile of tile c (32 KB)
//   MMA   : A_acc[t] += R_tile (K-major A)  * Gs_j[c]   (MN-major B)     8 x UMMA 128x128x16
//           B_acc[c&1] (+)= R_tile^T (MN-major A view of the same bytes) * Gs_i[t] (resident, MN-major B)
//   epilog: once both row blocks of tile c are multiplied, 4 warps drain B_acc[c&1] from TMEM, add the split
//           terms and red.add the 128 x k fp32 partial into B (L2-resident), while the tensor pipe already
//           works on tile c+1 with the other B_acc buffer.
// TMEM (512 columns): A_acc[0] 0..127 | A_acc[1] 128..255 | B_acc[0] 256..383 | B_acc[1] 384..511.
// SMEM (224 KB)     : 3 R stages x 32 KB | 2 Gs_j stages x 32 KB | 2 resident Gs_i tiles x 32 KB.
// Requires N = terms*64 == 128 (two split terms); other term counts use the two-pass kernels.
#pragma once
#include "sm100_ptx.cuh"

namespace fz {

struct FusedParams {
  float* A;            // [M_rows][lda]  (+)= R Gs_j
  float* B;            // [N_cols][ldb]  += R^T Gs_i       (always red.add; caller zeroes B)
  long long lda, ldb;
  int n_rows;          // local rows of R (rows of A)
  int n_cols;          // columns of R (rows of B)
  int k_a;             // valid columns of A  (rank of type j)
  int k_b;             // valid columns of B  (rank of type i)
  int gi_row0;         // row of Gs_i that pairs with local row 0 of R (row-sharded factors)
  int tiles_per_split; // column tiles handled per blockIdx.y
  int a_atomic;        // 1: several column splits add into A (caller zeroes A), 0: plain store
};

constexpr int kFuThreads = 192;
constexpr int kFuTile = 128;
constexpr int kFuRStages = 3;
constexpr int kFuTileBytes = kFuTile * kFuTile * 2;     // 32 KB: an R tile, or a 128-row x 128-col Gs tile
constexpr int kFuSmemBytes = (kFuRStages + 2 + 2) * kFuTileBytes + 1024 + 256;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kFuThreads, 1)
umma_fused_kernel(const __grid_constant__ CUtensorMap tmR,    // relation, box {64 cols, 128 rows}
                  const __grid_constant__ CUtensorMap tmGj,   // Gs_j,     box {64 cols, 128 rows}
                  const __grid_constant__ CUtensorMap tmGi,   // Gs_i,     box {64 cols, 128 rows}
                  const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* r_st = smem;                                         // kFuRStages x 32 KB
  uint8_t* gj_st = smem + kFuRStages * kFuTileBytes;            // 2 x 32 KB
  uint8_t* gi_st = gj_st + 2 * kFuTileBytes;                    // 2 x 32 KB (resident)
  uint64_t* bars = reinterpret_cast<uint64_t*>(gi_st + 2 * kFuTileBytes);
  uint64_t* r_full = bars;                   // [3]
  uint64_t* r_empty = r_full + kFuRStages;   // [3]
  uint64_t* gj_full = r_empty + kFuRStages;  // [2]
  uint64_t* gj_empty = gj_full + 2;          // [2]
  uint64_t* bacc_full = gj_empty + 2;        // [2]
  uint64_t* bacc_empty = bacc_full + 2;      // [2]
  uint64_t* gi_full = bacc_empty + 2;        // [1]
  uint64_t* aacc_full = gi_full + 1;         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aacc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 2 * kFuTile;                      // first local row of the pair
  const int total_tiles = (p.n_cols + kFuTile - 1) / kFuTile;
  const int tile_begin = blockIdx.y * p.tiles_per_split;
  const int tile_end = min(total_tiles, tile_begin + p.tiles_per_split);
  const int n_tiles = max(0, tile_end - tile_begin);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmR);
    ptx::prefetch_tmap(&tmGj);
    ptx::prefetch_tmap(&tmGi);
    for (int s = 0; s < kFuRStages; ++s) { ptx::mbar_init(&r_full[s], 1); ptx::mbar_init(&r_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&gj_full[s], 1);
      ptx::mbar_init(&gj_empty[s], 1);
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
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0 && n_tiles > 0) {
      // resident Gs_i tiles of the two row blocks
      ptx::mbar_expect_tx(gi_full, 2 * kFuTileBytes);
      for (int t = 0; t < 2; ++t)
        for (int ch = 0; ch < 2; ++ch)
          ptx::tma_load_2d(gi_st + t * kFuTileBytes + ch * 16384, &tmGi, gi_full, ch * 64, p.gi_row0 + r0 + t * kFuTile,
                           ptx::kEvictLast);
      int it = 0;
      for (int c = 0; c < n_tiles; ++c) {
        const int col0 = (tile_begin + c) * kFuTile;
        const int gs = c & 1;
        ptx::mbar_wait(&gj_empty[gs], ((c >> 1) & 1) ^ 1);
        ptx::mbar_expect_tx(&gj_full[gs], kFuTileBytes);
        for (int ch = 0; ch < 2; ++ch)
          ptx::tma_load_2d(gj_st + gs * kFuTileBytes + ch * 16384, &tmGj, &gj_full[gs], ch * 64, col0, ptx::kEvictLast);
        for (int t = 0; t < 2; ++t, ++it) {
          const int s = it % kFuRStages;
          ptx::mbar_wait(&r_empty[s], ((it / kFuRStages) & 1) ^ 1);
          ptx::mbar_expect_tx(&r_full[s], kFuTileBytes);
          for (int ch = 0; ch < 2; ++ch)
            ptx::tma_load_2d(r_st + s * kFuTileBytes + ch * 16384, &tmR, &r_full[s], col0 + ch * 64, r0 + t * kFuTile,
                             ptx::kEvictFirst);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idesc_a = ptx::idesc_bf16_f32(128, 128, false, true);   // R K-major  x Gs MN-major
      constexpr uint32_t idesc_b = ptx::idesc_bf16_f32(128, 128, true, true);    // R^T MN-major x Gs MN-major
      ptx::mbar_wait(gi_full, 0);
      int it = 0;
      for (int c = 0; c < n_tiles; ++c) {
        const int gs = c & 1;
        ptx::mbar_wait(&gj_full[gs], (c >> 1) & 1);
        ptx::mbar_wait(&bacc_empty[gs], ((c >> 1) & 1) ^ 1);      // epilogue has drained this B_acc buffer
        ptx::tc_fence_after();
        const uint32_t gj = ptx::smem_u32(gj_st + gs * kFuTileBytes);
        const uint32_t bacc = tmem_base + 256 + gs * 128;
        for (int t = 0; t < 2; ++t, ++it) {
          const int s = it % kFuRStages;
          ptx::mbar_wait(&r_full[s], (it / kFuRStages) & 1);
          ptx::tc_fence_after();
          const uint32_t rt = ptx::smem_u32(r_st + s * kFuTileBytes);
          const uint32_t gi = ptx::smem_u32(gi_st + t * kFuTileBytes);
          const uint32_t aacc = tmem_base + t * 128;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {   // A_acc[t] += R_tile * Gs_j[c]        (reduction over the tile's columns)
            const uint64_t ad = ptx::smem_desc_sw128(rt + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
            const uint64_t bd = ptx::smem_desc_sw128(gj + ks * 2048, 16384, 1024);
            ptx::umma_bf16(aacc, ad, bd, idesc_a, (c | ks) != 0);
          }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {   // B_acc += R_tile^T * Gs_i[t]         (reduction over the tile's rows)
            const uint64_t ad = ptx::smem_desc_sw128(rt + ks * 2048, 16384, 1024);
            const uint64_t bd = ptx::smem_desc_sw128(gi + ks * 2048, 16384, 1024);
            ptx::umma_bf16(bacc, ad, bd, idesc_b, (t | ks) != 0);
          }
          ptx::umma_commit(&r_empty[s]);
        }
        ptx::umma_commit(&gj_empty[gs]);
        ptx::umma_commit(&bacc_full[gs]);
      }
      ptx::umma_commit(aacc_full);
    }
  } else {
    // ---------------------------------------------------------------- epilogue (warps 2..5)
    const int quarter = warp & 3;
    const int lrow = quarter * 32 + lane;                         // TMEM lane = row of the accumulator tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const bool vec_b = ((p.ldb & 3) == 0) && ((p.k_b & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0);
    for (int c = 0; c < n_tiles; ++c) {
      const int gs = c & 1;
      ptx::mbar_wait(&bacc_full[gs], (c >> 1) & 1);
      ptx::tc_fence_after();
      const int bcol = (tile_begin + c) * kFuTile + lrow;          // column of R == row of B
      float* brow = (bcol < p.n_cols) ? p.B + (long long)bcol * p.ldb : nullptr;
#pragma unroll
      for (int q0 = 0; q0 < 64; q0 += 32) {
        float hi[32], lo[32];
        ptx::tmem_ld32(lane_addr + 256 + gs * 128 + q0, hi);
        ptx::tmem_ld32(lane_addr + 256 + gs * 128 + 64 + q0, lo);
        ptx::tmem_ld_wait();
        if (q0 == 32) {             // all TMEM reads of this buffer are done: hand it back to the MMA warp
          ptx::tc_fence_before();
          ptx::mbar_arrive(&bacc_empty[gs]);
        }
        if (brow != nullptr) {
          if (vec_b) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              if (q0 + i < p.k_b)
                red_add_v4(brow + q0 + i, hi[i] + lo[i], hi[i + 1] + lo[i + 1], hi[i + 2] + lo[i + 2], hi[i + 3] + lo[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (q0 + i < p.k_b) atomicAdd(brow + q0 + i, hi[i] + lo[i]);
          }
        }
      }
    }
    // final A accumulators of the two row blocks
    if (n_tiles > 0) {
      ptx::mbar_wait(aacc_full, 0);
      ptx::tc_fence_after();
      for (int t = 0; t < 2; ++t) {
        const int arow = r0 + t * kFuTile + lrow;
        float* out = (arow < p.n_rows) ? p.A + (long long)arow * p.lda : nullptr;
#pragma unroll
        for (int q0 = 0; q0 < 64; q0 += 32) {
          float hi[32], lo[32];
          ptx::tmem_ld32(lane_addr + t * 128 + q0, hi);
          ptx::tmem_ld32(lane_addr + t * 128 + 64 + q0, lo);
          ptx::tmem_ld_wait();
          if (out != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (q0 + i < p.k_a) {
                const float v = hi[i] + lo[i];
                if (p.a_atomic) atomicAdd(out + q0 + i, v);
                else out[q0 + i] = v;
              }
            }
          }
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace fz
