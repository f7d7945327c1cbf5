// Rank-k profile product on the tensor cores:  C[M x N] = X[M x k] * Y[N x k]^T  with k <= 64 -- the shape of
//   complete()      G_i (S_ij G_j^T)                  (reference skfusion/fusion/base/base.py:119-167)
//   chain profiles  G_i (S_ab S_bc ...) G_j^T          (reference examples/dicty_chaining.py:40-53)
// The product is OUTPUT-bound (M N fp32 written for 2 M N k flop), so the roofline is HBM write bandwidth and the
// tensor work is free: both operands enter as two bf16 terms (x = xh + xl, y = yh + yl) and the kernel multiplies the
// K-concatenated forms  X3 = [xh | xh | xl],  Y3 = [yh | yl | yh]  (K = 192), i.e. xh yh + xh yl + xl yh -- everything
// except the 2^-18 term xl yl, so the result carries fp32-level accuracy although the MMAs only see bf16.
//
// Persistent grid; work = (128-row blocks of X) x (128-row tiles of Y) flattened row-block-major, each CTA a contiguous
// range walked as segments (one resident X3 tile per segment).  Per output tile: TMA loads the Y3 tile (3 boxes of
// 128 x 64, 128B swizzle), 12 x UMMA 128x128x16 into one of two TMEM accumulators, and the four epilogue warps drain
// the other accumulator through swizzled shared-memory boxes into TMA tensor stores (full-line writes, clipped at the
// matrix edge by the tensor map).
#pragma once
#include "sm100_ptx.cuh"

namespace fz {

struct OuterParams {
  int M, N;      // rows of X (rows of C), rows of Y (columns of C)
};

constexpr int kOuThreads = 192;   // warp 0: TMA producer | 1: MMA | 2..5: epilogue
constexpr int kOuTile = 128;
constexpr int kOuK = 192;                                 // 3 K-blocks of 64
constexpr int kOuOpBytes = kOuTile * kOuK * 2;            // 48 KB: one operand tile (3 boxes of 16 KB)
constexpr int kOuYStages = 2;
constexpr int kOuStageBytes = 4 * 2 * 4096;               // store staging: 4 warps x 2 x (32 rows x 32 fp32)
constexpr int kOuSmemBytes = kOuOpBytes + kOuYStages * kOuOpBytes + kOuStageBytes + 1024 + 256;

namespace ptx {
// 2-D tiled store: the smem box (swizzled as the tensor map says) is written to global memory by the TMA unit.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
}  // namespace ptx

struct OuSegments {
  long long u, u_end;
  int tiles;
  __device__ OuSegments(int M, int N) {
    tiles = (N + kOuTile - 1) / kOuTile;
    const long long units = (long long)((M + kOuTile - 1) / kOuTile) * tiles;
    u = units * blockIdx.x / gridDim.x;
    u_end = units * (blockIdx.x + 1) / gridDim.x;
  }
  __device__ bool next(int& block, int& tile0, int& n) {
    if (u >= u_end) return false;
    block = (int)(u / tiles);
    tile0 = (int)(u % tiles);
    n = (int)min((long long)(tiles - tile0), u_end - u);
    u += n;
    return true;
  }
};

__global__ void __launch_bounds__(kOuThreads, 1)
umma_outer_kernel(const __grid_constant__ CUtensorMap tmX,   // X3, bf16 [M][192], box {64 cols, 128 rows}
                  const __grid_constant__ CUtensorMap tmY,   // Y3, bf16 [N][192], box {64 cols, 128 rows}
                  const __grid_constant__ CUtensorMap tmC,   // C,  fp32 [M][N],   box {32 cols, 32 rows} (store target)
                  const OuterParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* x_st = smem;                                   // 48 KB (resident per segment)
  uint8_t* y_st = x_st + kOuOpBytes;                      // 2 x 48 KB
  uint8_t* c_st = y_st + kOuYStages * kOuOpBytes;         // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(c_st + kOuStageBytes);
  uint64_t* y_full = bars;                   // [2]
  uint64_t* y_empty = y_full + kOuYStages;   // [2]
  uint64_t* acc_full = y_empty + kOuYStages; // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint64_t* x_full = acc_empty + 2;          // [1] per segment
  uint64_t* x_empty = x_full + 1;            // [1] per segment
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmY);
    ptx::prefetch_tmap(&tmC);
    for (int s = 0; s < kOuYStages; ++s) { ptx::mbar_init(&y_full[s], 1); ptx::mbar_init(&y_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&acc_full[s], 1); ptx::mbar_init(&acc_empty[s], 128); }
    ptx::mbar_init(x_full, 1);
    ptx::mbar_init(x_empty, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  OuSegments segs(p.M, p.N);
  int block, tile0, n_tiles;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (elected lane of a converged warp)
    int ct = 0, seg = 0;
    while (segs.next(block, tile0, n_tiles)) {
      ptx::mbar_wait(x_empty, (seg & 1) ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(x_full, kOuOpBytes);
        for (int kb = 0; kb < 3; ++kb)
          ptx::tma_load_2d(x_st + kb * 16384, &tmX, x_full, kb * 64, block * kOuTile, ptx::kEvictNormal);
      }
      __syncwarp();
      for (int c = 0; c < n_tiles; ++c, ++ct) {
        const int s = ct % kOuYStages;
        ptx::mbar_wait(&y_empty[s], ((ct / kOuYStages) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&y_full[s], kOuOpBytes);
          for (int kb = 0; kb < 3; ++kb)
            ptx::tma_load_2d(y_st + s * kOuOpBytes + kb * 16384, &tmY, &y_full[s], kb * 64, (tile0 + c) * kOuTile, ptx::kEvictLast);
        }
        __syncwarp();
      }
      ++seg;
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = ptx::idesc_bf16_f32(128, 128, false, false);    // both operands K-major
    const uint32_t xb = ptx::smem_u32(x_st), yb = ptx::smem_u32(y_st);
    int ct = 0, seg = 0;
    while (segs.next(block, tile0, n_tiles)) {
      ptx::mbar_wait(x_full, seg & 1);
      ptx::tc_fence_after();
      for (int c = 0; c < n_tiles; ++c, ++ct) {
        const int s = ct % kOuYStages, a = ct & 1;
        ptx::mbar_wait(&y_full[s], (ct / kOuYStages) & 1);
        ptx::mbar_wait(&acc_empty[a], ((ct >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t yt = yb + s * kOuOpBytes;
#pragma unroll
          for (int ks = 0; ks < 12; ++ks)
            ptx::umma_bf16(tmem_base + a * 128, ptx::smem_desc_sw128(xb + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                           ptx::smem_desc_sw128(yt + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), idesc, ks != 0);
          ptx::umma_commit(&y_empty[s]);
          ptx::umma_commit(&acc_full[a]);
          if (c == n_tiles - 1) ptx::umma_commit(x_empty);
        }
        __syncwarp();
      }
      ++seg;
    }
  } else {
    // ---------------------------------------------------------------- epilogue (warps 2..5): TMEM -> smem boxes -> TMA store
    const int quarter = warp & 3;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint8_t* my_stage = c_st + quarter * 8192;        // two 32 x 32 fp32 boxes, used alternately
    int ct = 0, box = 0;
    bool staged = false;
    while (segs.next(block, tile0, n_tiles)) {
      const int crow0 = block * kOuTile + quarter * 32;
      for (int c = 0; c < n_tiles; ++c, ++ct) {
        const int a = ct & 1;
        ptx::mbar_wait(&acc_full[a], (ct >> 1) & 1);
        ptx::tc_fence_after();
        const int ccol0 = (tile0 + c) * kOuTile;
#pragma unroll
        for (int q0 = 0; q0 < 128; q0 += 32) {
          float v[32];
          ptx::tmem_ld32(lane_addr + a * 128 + q0, v);
          ptx::tmem_ld_wait();
          if (q0 == 96) {
            ptx::tc_fence_before();
            ptx::mbar_arrive(&acc_empty[a]);
          }
          if (ccol0 + q0 >= p.N || crow0 >= p.M) continue;
          uint8_t* st = my_stage + (box & 1) * 4096;
          if (staged) {                                 // the store issued two boxes ago has read this buffer
            if (ptx::elect_one()) ptx::tma_wait_read<1>();
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(st + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          ptx::fence_proxy_async();
          __syncwarp();
          if (ptx::elect_one()) {
            ptx::tma_store_2d(&tmC, st, ccol0 + q0, crow0);
            ptx::tma_commit_group();
          }
          __syncwarp();
          staged = true;
          ++box;
        }
      }
    }
    __syncwarp();
    if (staged && ptx::elect_one()) ptx::tma_wait_all();
    __syncwarp();
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<256>(tmem_base);
  }
}

// Operand forms of the profile product: Z3[r] = [hi | hi | lo] (lo_second = false) or [hi | lo | hi] (lo_second = true),
// hi = bf16(z), lo = bf16(z - hi); columns >= k and rows >= n_valid are zero.  grid over n_pad * 64 elements.
template <class T>
__global__ void outer_operand(const T* __restrict__ Z, long long ldz, __nv_bfloat16* __restrict__ Z3, long long n_valid, long long n_pad,
                              int k, int lo_second) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pad * 64) return;
  const long long r = idx / 64;
  const int q = (int)(idx % 64);
  float v = 0.f;
  if (r < n_valid && q < k) v = (float)Z[r * ldz + q];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  __nv_bfloat16* out = Z3 + r * kOuK + q;
  out[0] = hi;
  out[64] = lo_second ? lo : hi;
  out[128] = lo_second ? hi : lo;
}

}  // namespace fz
