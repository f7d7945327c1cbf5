// Developer microbenchmark (not part of the library): issue pacing of tcgen05.mma for the operand forms the fused
// kernels use, with operands already resident (no TMA, no epilogue).  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_pace mma_pace.cu && ./mma_pace
// Prints cycles per MMA and the shared-memory operand bytes per cycle each form needs.
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../sm100_ptx.cuh"
#include "../umma_fused_t.cuh"

using namespace fz;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

struct PaceParams {
  int a_mode;      // 0: SS, A K-major   1: SS, A MN-major   2: TS (A in TMEM)
  int b_mn;        // 0: B K-major       1: B MN-major
  int n;           // N of the MMA
  int groups;      // groups of 8 MMAs issued
  int nbuf;        // distinct 48 KB operand buffers cycled through (1..4)
  int ld_warps;    // 0..4 extra warps that keep draining 64 TMEM columns per round (epilogue traffic)
  int st_bytes;    // bytes of st.shared traffic per group issued by a 6th warp (0 = none), mimics TMA fill pressure
  int n_acc;       // accumulators the MMAs rotate over (1 = one dependent chain; 2, 4 = independent chains)
  int fixed_desc;  // 1: every MMA uses the same precomputed descriptors (pure issue loop)
};

constexpr int kPaceSmem = 4 * 49152 + 16384 + 1024 + 256;

__global__ void __launch_bounds__(224, 1) pace_kernel(PaceParams p, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * 49152 + 16384);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(bar + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (4 * 49152 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c3c3c3cu;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar[0], 1);
    ptx::fence_barrier_init();
    *stop = 0;
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::idesc_bf16_f32(128, p.n, p.a_mode == 1, p.b_mn != 0);
      const uint32_t base0 = ptx::smem_u32(smem);
      const int acc_stride = p.n;                      // accumulators side by side
      const uint32_t tmem_a = tmem_base + 448;          // 64 columns of A operand (TS)
      const long long t0 = clock64();
      for (int g = 0; g < p.groups; ++g) {
        const uint32_t base = ptx::smem_u32(smem + (g % p.nbuf) * 49152);
        const uint32_t a_base = base + 32768;   // 16 KB region for the A operand
        if (p.fixed_desc) {
          const uint64_t bd = p.b_mn ? ptx::smem_desc_sw128(base0, 16384, 1024) : ptx::smem_desc_sw128(base0, 16, 1024);
          const uint64_t ad = p.a_mode == 1 ? ptx::smem_desc_sw128(base0 + 32768, 8192, 1024) : ptx::smem_desc_sw128(base0 + 32768, 16, 1024);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t d = tmem_base + (uint32_t)((ks % p.n_acc) * acc_stride);
            if (p.a_mode == 2) ptx::umma_bf16_ts(d, tmem_a, bd, idesc, 1);
            else ptx::umma_bf16(d, ad, bd, idesc, 1);
          }
          continue;
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t dacc = tmem_base + (uint32_t)((ks % p.n_acc) * acc_stride);
          // B operand: K-major: 32 B steps inside 128 B rows (4 per 64-column chunk);  MN-major: 2048 B steps
          const uint64_t bdesc = p.b_mn ? ptx::smem_desc_sw128(base + ks * 2048, 16384, 1024)
                                        : ptx::smem_desc_sw128(base + (ks & 3) * 32, 16, 1024);
          if (p.a_mode == 2) {
            ptx::umma_bf16_ts(dacc, tmem_a + ks * 8, bdesc, idesc, 1);
          } else {
            const uint64_t adesc = p.a_mode == 1 ? ptx::smem_desc_sw128(a_base + (ks & 3) * 2048, 8192, 1024)
                                                 : ptx::smem_desc_sw128(a_base + (ks & 3) * 32, 16, 1024);
            ptx::umma_bf16(dacc, adesc, bdesc, idesc, 1);
          }
        }
      }
      ptx::umma_commit(&bar[0]);
      ptx::mbar_wait_wd(&bar[0], 0);
      const long long t1 = clock64();
      *stop = 1;
      cycles[blockIdx.x] = t1 - t0;
    }
  } else if (warp >= 2 && warp < 6) {
    if (warp - 2 < p.ld_warps) {
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
      float acc = 0.f;
      while (!*stop) {
        float x[32];
        ptx::tmem_ld32(lane_addr + 384, x);
        ptx::tmem_ld_wait();
        acc += x[lane];
        ptx::tmem_ld32(lane_addr + 416, x);
        ptx::tmem_ld_wait();
        acc += x[lane];
      }
      if (acc == 123.456f) cycles[0] = 0;
    }
  } else if (warp == 6 && p.st_bytes > 0) {
    // shared-memory write pressure in the flush staging region (16 KB at the end), 512 B per warp store
    uint4* dst = reinterpret_cast<uint4*>(smem + 4 * 49152);
    uint4 v = make_uint4(lane, 1, 2, 3);
    int i = 0;
    while (!*stop) {
      dst[(i & 31) * 32 + lane] = v;
      ++i;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

static void run(const char* label, PaceParams p, int grid) {
  long long* d;
  CK(cudaMalloc(&d, grid * sizeof(long long)));
  CK(cudaFuncSetAttribute(pace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPaceSmem));
  pace_kernel<<<grid, 224, kPaceSmem>>>(p, d);   // warm-up
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  pace_kernel<<<grid, 224, kPaceSmem>>>(p, d);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h(grid);
  CK(cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (long long c : h) avg += (double)c;
  avg /= grid;
  const double mmas = 8.0 * p.groups;
  const double smem_bytes = (p.a_mode == 2 ? 0.0 : 128 * 32.0) + p.n * 32.0;
  const double flop = 2.0 * 128 * p.n * 16 * mmas * grid;
  printf("%-46s N=%3d grid=%3d ld_warps=%d st=%d : %7.1f clk/MMA (floor %3d)  operand smem %.0f B/clk  %.0f TFLOP/s  (%.3f ms)\n", label, p.n, grid,
         p.ld_warps, p.st_bytes ? 1 : 0, avg / mmas, 128 * p.n / 256, smem_bytes / (avg / mmas), flop / (ms * 1e-3) / 1e12, ms);
  cudaFree(d);
}

int main(int argc, char** argv) {
  const int groups = argc > 1 ? atoi(argv[1]) : 5000;
  const int grid = 148;
  struct Form { const char* label; int a_mode, b_mn; };
  const Form forms[] = {{"SS K x MN", 0, 1}, {"SS MN x MN", 1, 1}, {"SS K x K", 0, 0}, {"TS x MN", 2, 1}};
  for (int fixed = 0; fixed < 2; ++fixed)
    for (const Form& f : forms)
      for (int n : {64, 128, 256})
        for (int n_acc : {1, 2, 4}) {
          if (n * n_acc > 384) continue;
          char label[96];
          snprintf(label, sizeof label, "%s%s n_acc=%d", f.label, fixed ? " [fixed desc]" : "", n_acc);
          run(label, PaceParams{f.a_mode, f.b_mn, n, groups, 4, 0, 0, n_acc, fixed}, grid);
        }
  return 0;
}
