// Fused single-pass streamed products, SINGLE-TERM operand form (fused kernel v5).
//
// Same contract as umma_fused.cuh -- one read of a bf16 relation tile feeds both
//        A_ij = R_ij G_j          B_ij = R_ij^T G_i
// -- but the factor enters as ONE bf16 term of its mean-centred form:  G = 1 c^T + D,  Gs = bf16(D).
// The rank-1 part is exact algebra outside the MMAs (rowsum(R) c_j^T is added in the A epilogue here, colsum(R) c_i^T is
// the initial value of B), and the first-order effect of the dropped residual lo = D - Gs on the backbone solve is
// restored in fp64 from B itself (G_i^T R lo_j = B^T lo_j, fz_engine.cu: corr_M).  What that buys:
//   * half the executed tensor work per relation byte (128 instead of 256 flop/B): the v3 kernel is bound by the
//     1000 W power cap through its MMAs (Finding 2);
//   * 64-column accumulators, so a CTA holds FOUR 128-row blocks of A in TMEM (v3: two) and the B partial of a
//     column tile is flushed once per 512 relation rows: half the L2 reduce traffic per relation byte.
// Whether this operand form is accurate enough for a given graph is decided by the engine from a measured error
// estimate (fz_engine.cu: choose_terms); ill-conditioned small problems keep the two-term kernel.
//
// Persistent grid (one CTA per SM); a CTA walks a contiguous range of (512-row group, 128-column tile) units (F1Segments).
// Per column tile c of a row group:
//   TMA   : R[r0+128t .., c] for every row block t (32 KB each, 128B swizzle; warp 0); Gs_j[c] (128 x 64, 16 KB; warp 6)
//   MMA   : A_acc[t] += R_tile (K-major A) * Gs_j[c] (MN-major B)            8 x UMMA 128x64x16
//           B_acc[c&1] (+)= R_tile^T (MN-major view of the same bytes) * Gs_i[t] (resident)   8 x UMMA 128x64x16
//   epilog: after the last row block of tile c, 4 warps drain B_acc[c&1], stage the 128 x 64 fp32 partial (swizzled)
//           and hand it to the TMA unit as cp.reduce.async.bulk.tensor .add into B (reduction in L2).
// TMEM (512 columns): A_acc[t] at 64 t (t < 4) | B_acc[0] 256..319 | B_acc[1] 320..383.
// SMEM (~212 KB)    : 3 R stages x 32 KB | 2 Gs_j slots x 16 KB | 32 KB flush staging | 4 resident Gs_i x 16 KB.
#pragma once
#include "sm100_ptx.cuh"

namespace fz {

struct Fused1Params {
  float* A;             // [n_rows][lda]  += R Gs_j + rowsum c_j^T     (always reduced into; caller zeroes A)
  float* B;             // [n_cols][ldb]  += R^T Gs_i                  (always reduced into; caller initialises B with colsum c_i^T)
  long long lda, ldb;
  const float* rowsum;  // [n_rows] row sums of R (fp32), or nullptr: no rank-1 term
  const float* cj;      // [k_a] centre of the column factor
  int n_rows, n_cols;
  int k_a, k_b;
  int gi_row0;          // row of Gs_i that pairs with local row 0 of R (row-sharded factors)
  int* work_counter;    // dynamic tail of the schedule: a device int, zero at launch (nullptr / dyn_chunk == 0: static only)
  int dyn_chunk;        // units per dynamic chunk (0 = the whole launch is statically partitioned)
  int tma_flush;        // bit0: B partials go out as TMA reduce-add (needs tmB); bit1: A likewise (needs tmA); else red.global
  int probe;            // developer probe only (wrong results): bit0 no reductions, bit1 no B MMAs, bit2 no A MMAs
};

// Shared-memory geometry (compile-time; the probe csrc/dev/umma_probe.cu is built with each candidate):
//   FZ_F1_BLOCKS   128-row blocks per row group = A accumulators in TMEM = resident Gs_i tiles; the B partial of a column
//                  tile is flushed once per FZ_F1_BLOCKS relation tiles
//   FZ_F1_RSTAGES  relation tiles (32 KB each) in flight per SM
//   FZ_F1_STAGING  flush staging: 32768 = both 32-column halves of a 128 x 64 fp32 partial at once, 16384 = one half
//                  after the other through the same 4 KB per warp
#ifndef FZ_F1_BLOCKS
#define FZ_F1_BLOCKS 4
#endif
#ifndef FZ_F1_RSTAGES
#define FZ_F1_RSTAGES 3
#endif
#ifndef FZ_F1_STAGING
#define FZ_F1_STAGING 32768
#endif
constexpr int kF1Threads = 224;   // warp 0: R producer | 1: MMA | 2..5: epilogue | 6: Gs producer
constexpr int kF1Blocks = FZ_F1_BLOCKS;      // 128-row blocks per row group
constexpr int kF1Tile = 128;
constexpr int kF1RStages = FZ_F1_RSTAGES;
constexpr int kF1GjSlots = 2;
constexpr int kF1TileBytes = kF1Tile * kF1Tile * 2;   // 32 KB relation tile
constexpr int kF1GBytes = kF1Tile * 64 * 2;           // 16 KB: 128 rows x 64 columns of Gs (one term)
constexpr int kF1StageBytes = FZ_F1_STAGING;          // flush staging: 4 warps x (2 or 1) x (32 rows x 32 fp32)
constexpr bool kF1HalfStaging = kF1StageBytes < 32768;
constexpr int kF1SmemBytes = kF1RStages * kF1TileBytes + kF1GjSlots * kF1GBytes + kF1StageBytes + kF1Blocks * kF1GBytes + 1024 + 256;
static_assert(kF1StageBytes == 32768 || kF1StageBytes == 16384, "staging holds one or both halves of a partial");
static_assert(kF1Blocks >= 1 && kF1Blocks <= 4, "A accumulators occupy TMEM columns [0, 64 * blocks), B from 256");
static_assert(kF1SmemBytes <= 227 * 1024, "shared memory budget of one CTA per SM");

// Work of one launch = (row groups of 512 rows) x (128-column tiles), flattened row-group-major into "units" of one column
// tile of one row group.  The first three quarters of the units are partitioned STATICALLY: CTA b of the persistent grid owns
// a contiguous range, walked as at most a few SEGMENTS (maximal runs inside one row group, each with its own resident Gs_i
// and A accumulators) -- no wave quantisation, no per-unit overhead.  The last quarter is handed out DYNAMICALLY in small
// chunks (an atomic counter; the relation-producer warp fetches, the other roles follow through a shared-memory ring), so an
// SM that shares its cycles with another kernel -- the fp64 reductions of the previous relation, an NCCL reduce-scatter --
// simply takes fewer chunks instead of holding the whole launch back.
struct F1Segments {
  long long u, u_end, units, dyn_begin;
  int tiles, chunk, fetched;
  __device__ F1Segments(int n_rows, int n_cols, int dyn_chunk) {
    tiles = (n_cols + kF1Tile - 1) / kF1Tile;
    const long long groups = (n_rows + kF1Blocks * kF1Tile - 1) / (kF1Blocks * kF1Tile);
    units = groups * tiles;
    chunk = dyn_chunk;
    fetched = 0;
    dyn_begin = chunk > 0 ? units - units / 4 : units;
    u = dyn_begin * blockIdx.x / gridDim.x;
    u_end = dyn_begin * (blockIdx.x + 1) / gridDim.x;
  }
  // next segment: row group, first tile, number of tiles.  fetch(k) returns the k-th dynamic chunk index of this CTA.
  template <class Fetch>
  __device__ bool next(int& group, int& tile0, int& n, Fetch&& fetch) {
    while (u >= u_end) {
      if (chunk <= 0) return false;
      const long long c = fetch(fetched++);
      u = dyn_begin + c * chunk;
      u_end = min(units, u + chunk);
      if (u >= units) { chunk = 0; return false; }
    }
    group = (int)(u / tiles);
    tile0 = (int)(u % tiles);
    n = (int)min((long long)(tiles - tile0), u_end - u);
    u += n;
    return true;
  }
};

__global__ void __launch_bounds__(kF1Threads, 1)
umma_fused1_kernel(const __grid_constant__ CUtensorMap tmR,    // relation, bf16, box {64 cols, 128 rows}
                   const __grid_constant__ CUtensorMap tmGj,   // Gs_j,     bf16, box {64 cols, 128 rows}
                   const __grid_constant__ CUtensorMap tmGi,   // Gs_i,     bf16, box {64 cols, 128 rows}
                   const __grid_constant__ CUtensorMap tmB,    // B,        fp32, box {32 cols, 32 rows} (reduce target)
                   const __grid_constant__ CUtensorMap tmA,    // A,        fp32, box {32 cols, 32 rows} (reduce target)
                   const Fused1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* r_st = smem;                                          // 3 x 32 KB
  uint8_t* gj_st = r_st + kF1RStages * kF1TileBytes;             // 2 x 16 KB
  uint8_t* fl_st = gj_st + kF1GjSlots * kF1GBytes;               // 32 KB
  uint8_t* gi_st = fl_st + kF1StageBytes;                        // 4 x 16 KB (resident per segment)
  uint64_t* bars = reinterpret_cast<uint64_t*>(gi_st + kF1Blocks * kF1GBytes);
  uint64_t* r_full = bars;                     // [3]
  uint64_t* r_empty = r_full + kF1RStages;     // [3]
  uint64_t* gj_full = r_empty + kF1RStages;    // [2]
  uint64_t* gj_empty = gj_full + kF1GjSlots;   // [2]
  uint64_t* bacc_full = gj_empty + kF1GjSlots; // [2]
  uint64_t* bacc_empty = bacc_full + 2;        // [2]
  uint64_t* gi_full = bacc_empty + 2;          // [1]  per segment
  uint64_t* gi_empty = gi_full + 1;            // [1]  per segment: the segment's B-product MMAs have read Gs_i
  uint64_t* aacc_full = gi_empty + 1;          // [1]  per segment
  uint64_t* aacc_empty = aacc_full + 1;        // [1]  per segment: the epilogue has drained A_acc
  uint64_t* q_full = aacc_empty + 1;           // [4]  dynamic-chunk ring: entry published by the relation producer
  uint64_t* q_empty = q_full + 4;              // [4]  ... and read by the MMA warp, the Gs producer and the 4 epilogue warps
  int* q_val = reinterpret_cast<int*>(q_empty + 4);   // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_val + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmR);
    ptx::prefetch_tmap(&tmGj);
    ptx::prefetch_tmap(&tmGi);
    if (p.tma_flush & 1) ptx::prefetch_tmap(&tmB);
    if (p.tma_flush & 2) ptx::prefetch_tmap(&tmA);
    for (int s = 0; s < kF1RStages; ++s) { ptx::mbar_init(&r_full[s], 1); ptx::mbar_init(&r_empty[s], 1); }
    for (int s = 0; s < kF1GjSlots; ++s) { ptx::mbar_init(&gj_full[s], 1); ptx::mbar_init(&gj_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&bacc_full[s], 1);
      ptx::mbar_init(&bacc_empty[s], 128);   // every epilogue thread arrives
    }
    ptx::mbar_init(gi_full, 1);
    ptx::mbar_init(gi_empty, 1);
    ptx::mbar_init(aacc_full, 1);
    ptx::mbar_init(aacc_empty, 128);
    for (int s = 0; s < 4; ++s) { ptx::mbar_init(&q_full[s], 1); ptx::mbar_init(&q_empty[s], 6); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  F1Segments segs(p.n_rows, p.n_cols, p.work_counter != nullptr ? p.dyn_chunk : 0);
  int group, tile0, n_tiles;
  // k-th dynamic chunk of this CTA: fetched from the global counter by the relation producer, followed by everyone else
  auto fetch_lead = [&](int k) -> long long {
    const int slot = k & 3;
    ptx::mbar_wait(&q_empty[slot], ((k >> 2) & 1) ^ 1);
    if (ptx::elect_one()) {
      q_val[slot] = atomicAdd(p.work_counter, 1);
      ptx::mbar_arrive(&q_full[slot]);
    }
    __syncwarp();
    return (long long)q_val[slot];
  };
  auto fetch_follow = [&](int k) -> long long {
    const int slot = k & 3;
    ptx::mbar_wait(&q_full[slot], (k >> 2) & 1);
    const int v = q_val[slot];
    __syncwarp();
    if (ptx::elect_one()) ptx::mbar_arrive(&q_empty[slot]);
    __syncwarp();
    return (long long)v;
  };

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer: relation tiles
    // (whole warp, one elected lane issues: a divergent `if (lane == 0)` makes ptxas wrap every TMA / tcgen05
    //  instruction in a ~100-cycle waterfall loop, csrc/dev/mma_pace.cu)
    int it = 0;
    while (segs.next(group, tile0, n_tiles, fetch_lead)) {
      const int r0 = group * kF1Blocks * kF1Tile;
      const int nb = min(kF1Blocks, (p.n_rows - r0 + kF1Tile - 1) / kF1Tile);
      for (int c = 0; c < n_tiles; ++c) {
        const int col0 = (tile0 + c) * kF1Tile;
        for (int t = 0; t < nb; ++t, ++it) {
          const int s = it % kF1RStages;
          ptx::mbar_wait(&r_empty[s], ((it / kF1RStages) & 1) ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&r_full[s], kF1TileBytes);
            for (int ch = 0; ch < 2; ++ch)
              ptx::tma_load_2d(r_st + s * kF1TileBytes + ch * 16384, &tmR, &r_full[s], col0 + ch * 64, r0 + t * kF1Tile,
                               ptx::kEvictFirst);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 6) {
    // ---------------------------------------------------------------- TMA producer: factor operands
    int ct = 0, seg = 0;
    while (segs.next(group, tile0, n_tiles, fetch_follow)) {
      const int r0 = group * kF1Blocks * kF1Tile;
      const int nb = min(kF1Blocks, (p.n_rows - r0 + kF1Tile - 1) / kF1Tile);
      ptx::mbar_wait(gi_empty, (seg & 1) ^ 1);                     // previous segment's MMAs are done with Gs_i
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(gi_full, nb * kF1GBytes);             // resident Gs_i tiles of the segment's row blocks
        for (int t = 0; t < nb; ++t)
          ptx::tma_load_2d(gi_st + t * kF1GBytes, &tmGi, gi_full, 0, p.gi_row0 + r0 + t * kF1Tile, ptx::kEvictLast);
      }
      __syncwarp();
      for (int c = 0; c < n_tiles; ++c, ++ct) {
        const int slot = ct % kF1GjSlots;
        ptx::mbar_wait(&gj_empty[slot], ((ct / kF1GjSlots) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&gj_full[slot], kF1GBytes);
          ptx::tma_load_2d(gj_st + slot * kF1GBytes, &tmGj, &gj_full[slot], 0, (tile0 + c) * kF1Tile, ptx::kEvictLast);
        }
        __syncwarp();
      }
      ++seg;
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (whole warp converged, elected lane)
    const uint32_t idesc_a = ptx::idesc_bf16_f32(128, 64, false, true);   // R K-major      x Gs MN-major
    const uint32_t idesc_b = ptx::idesc_bf16_f32(128, 64, true, true);    // R^T (MN-major) x Gs MN-major
    const bool do_a = !(p.probe & 4), do_b = !(p.probe & 2);
    const uint32_t r_base = ptx::smem_u32(r_st), gj_base = ptx::smem_u32(gj_st), gi_base = ptx::smem_u32(gi_st);
    int it = 0, ct = 0, seg = 0;
    while (segs.next(group, tile0, n_tiles, fetch_follow)) {
      const int r0 = group * kF1Blocks * kF1Tile;
      const int nb = min(kF1Blocks, (p.n_rows - r0 + kF1Tile - 1) / kF1Tile);
      ptx::mbar_wait(aacc_empty, (seg & 1) ^ 1);                   // the previous segment's A accumulators are drained
      ptx::mbar_wait(gi_full, seg & 1);
      ptx::tc_fence_after();
      for (int c = 0; c < n_tiles; ++c, ++ct) {
        const int gs = ct & 1;
        const int slot = ct % kF1GjSlots;
        const uint32_t g = gj_base + slot * kF1GBytes;
        const uint32_t bacc = tmem_base + 256 + gs * 64;
        ptx::mbar_wait(&gj_full[slot], (ct / kF1GjSlots) & 1);
        for (int t = 0; t < nb; ++t, ++it) {
          const int s = it % kF1RStages;
          const uint32_t rt = r_base + s * kF1TileBytes;
          ptx::mbar_wait(&r_full[s], (it / kF1RStages) & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            if (do_a)
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                ptx::umma_bf16(tmem_base + t * 64, ptx::smem_desc_sw128(rt + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                               ptx::smem_desc_sw128(g + ks * 2048, 16, 1024), idesc_a, (c | ks) != 0);
            if (t == nb - 1) ptx::umma_commit(&gj_empty[slot]);     // last use of this Gs_j tile
          }
          __syncwarp();
          if (t == 0) {
            ptx::mbar_wait(&bacc_empty[gs], ((ct >> 1) & 1) ^ 1);   // epilogue has drained this B_acc buffer
            ptx::tc_fence_after();
          }
          if (ptx::elect_one()) {
            if (do_b)
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                ptx::umma_bf16(bacc, ptx::smem_desc_sw128(rt + ks * 2048, 16384, 1024),
                               ptx::smem_desc_sw128(gi_base + t * kF1GBytes + ks * 2048, 16, 1024), idesc_b, (t | ks) != 0);
            ptx::umma_commit(&r_empty[s]);
            if (t == nb - 1) ptx::umma_commit(&bacc_full[gs]);
          }
          __syncwarp();
        }
      }
      if (ptx::elect_one()) {
        ptx::umma_commit(gi_empty);
        ptx::umma_commit(aacc_full);
      }
      __syncwarp();
      ++seg;
    }
  } else {
    // ---------------------------------------------------------------- epilogue (warps 2..5)
    const int quarter = warp & 3;
    const int lrow = quarter * 32 + lane;                         // TMEM lane = row of the accumulator tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint8_t* my_stage = fl_st + quarter * (kF1StageBytes / 4);    // 2 (or 1) x (32 rows x 128 B), 128B-swizzled like the reduce boxes
    const bool skip_red = (p.probe & 1) != 0;
    const bool b_tma = (p.tma_flush & 1) != 0, a_tma = (p.tma_flush & 2) != 0;
    bool staged = false;                                          // this warp has reduces in flight that read its staging
    int ct = 0, seg = 0;
    while (segs.next(group, tile0, n_tiles, fetch_follow)) {
      const int r0 = group * kF1Blocks * kF1Tile;
      const int nb = min(kF1Blocks, (p.n_rows - r0 + kF1Tile - 1) / kF1Tile);
      for (int c = 0; c < n_tiles; ++c, ++ct) {
        const int gs = ct & 1;
        ptx::mbar_wait(&bacc_full[gs], (ct >> 1) & 1);
        ptx::tc_fence_after();
        const int brow0 = (tile0 + c) * kF1Tile + quarter * 32;     // first B row (column of R) of this warp
        float v0[32], v1[32];
        ptx::tmem_ld32(lane_addr + 256 + gs * 64, v0);
        ptx::tmem_ld32(lane_addr + 256 + gs * 64 + 32, v1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bacc_empty[gs]);                          // TMEM reads done: hand the buffer back to the MMA warp
        if (skip_red) continue;
        if (b_tma && kF1HalfStaging) {
          // one 32-column half after the other through the warp's single 4 KB box
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && p.k_b <= 32) break;
            if (staged) {
              if (ptx::elect_one()) ptx::tma_wait_read_all();       // earlier reduces have read the staging
              __syncwarp();
            }
            const float* v = h == 0 ? v0 : v1;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(my_stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            ptx::fence_proxy_async();
            __syncwarp();
            if (ptx::elect_one()) {
              ptx::tma_reduce_add_2d(&tmB, my_stage, 32 * h, brow0);
              ptx::tma_commit_group();
            }
            __syncwarp();
            staged = true;
          }
        } else if (b_tma) {
          if (staged) {
            if (ptx::elect_one()) ptx::tma_wait_read_all();         // earlier reduces have read the staging
            __syncwarp();
          }
          // row `lane` of each 32 x 32 box: 8 chunks of 16 B, chunk j stored at j ^ (lane & 7)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            *reinterpret_cast<float4*>(my_stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]);
            *reinterpret_cast<float4*>(my_stage + 4096 + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(v1[4 * j], v1[4 * j + 1], v1[4 * j + 2], v1[4 * j + 3]);
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (ptx::elect_one()) {                                   // deterministic: always the same lane
            ptx::tma_reduce_add_2d(&tmB, my_stage, 0, brow0);       // rows / columns beyond the tensor are clipped
            if (p.k_b > 32) ptx::tma_reduce_add_2d(&tmB, my_stage + 4096, 32, brow0);
            ptx::tma_commit_group();
          }
          __syncwarp();
          staged = true;
        } else {
          const int bcol = brow0 + lane;
          if (bcol < p.n_cols) {
            float* brow = p.B + (long long)bcol * p.ldb;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i < p.k_b) atomicAdd(brow + i, v0[i]);
              if (32 + i < p.k_b) atomicAdd(brow + 32 + i, v1[i]);
            }
          }
        }
      }
      // ---- the segment's A accumulators (+ the rank-1 part of the centred operand form, once per row: with tile 0)
      ptx::mbar_wait(aacc_full, seg & 1);
      ptx::tc_fence_after();
      const bool rank1 = (p.rowsum != nullptr) && (tile0 == 0);
      for (int t = 0; t < nb; ++t) {
        const int arow = r0 + t * kF1Tile + lrow;
        const bool live = arow < p.n_rows;
        const float rs = (rank1 && live) ? p.rowsum[arow] : 0.f;
        float v0[32], v1[32];
        ptx::tmem_ld32(lane_addr + t * 64, v0);
        ptx::tmem_ld32(lane_addr + t * 64 + 32, v1);
        ptx::tmem_ld_wait();
        if (t == nb - 1) {                                          // all TMEM reads of the segment are done
          ptx::tc_fence_before();
          ptx::mbar_arrive(aacc_empty);
        }
        if (skip_red) continue;
        if (rank1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v0[i] = (i < p.k_a) ? fmaf(rs, __ldg(p.cj + i), v0[i]) : v0[i];
            v1[i] = (32 + i < p.k_a) ? fmaf(rs, __ldg(p.cj + 32 + i), v1[i]) : v1[i];
          }
        }
        if (a_tma && kF1HalfStaging) {
          const int arow0 = r0 + t * kF1Tile + quarter * 32;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && p.k_a <= 32) break;
            if (staged) {
              if (ptx::elect_one()) ptx::tma_wait_read_all();
              __syncwarp();
            }
            const float* v = h == 0 ? v0 : v1;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(my_stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            ptx::fence_proxy_async();
            __syncwarp();
            if (ptx::elect_one()) {
              ptx::tma_reduce_add_2d(&tmA, my_stage, 32 * h, arow0);   // rows beyond n_rows / columns beyond k_a are clipped
              ptx::tma_commit_group();
            }
            __syncwarp();
            staged = true;
          }
        } else if (a_tma) {
          if (staged) {
            if (ptx::elect_one()) ptx::tma_wait_read_all();
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            *reinterpret_cast<float4*>(my_stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]);
            *reinterpret_cast<float4*>(my_stage + 4096 + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(v1[4 * j], v1[4 * j + 1], v1[4 * j + 2], v1[4 * j + 3]);
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (ptx::elect_one()) {
            const int arow0 = r0 + t * kF1Tile + quarter * 32;
            ptx::tma_reduce_add_2d(&tmA, my_stage, 0, arow0);       // rows beyond n_rows / columns beyond k_a are clipped
            if (p.k_a > 32) ptx::tma_reduce_add_2d(&tmA, my_stage + 4096, 32, arow0);
            ptx::tma_commit_group();
          }
          __syncwarp();
          staged = true;
        } else if (live) {
          float* out = p.A + (long long)arow * p.lda;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < p.k_a) atomicAdd(out + i, v0[i]);
            if (32 + i < p.k_a) atomicAdd(out + 32 + i, v1[i]);
          }
        }
      }
      ++seg;
    }
    __syncwarp();
    if (staged && ptx::elect_one()) ptx::tma_wait_all();          // reductions performed before the CTA retires
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
