// Reduce-scatter of the B partials over NVLink PEER MEMORY, fused with what follows it.
//
// Sharded DFMF needs, per relation, B_ij = sum over ranks of the partial R_ij[p]^T G_i[p] on the rows of type j each rank
// owns.  An NCCL reduce-scatter does that with a few long-running CTAs; beside the persistent streamed
// kernel (one CTA per SM, 225 KB of shared memory -- nothing can co-reside) those CTAs take SMs hostage for ~100 us, and the
// streamed kernel's static partition turns every late SM into a late launch.  Here every rank instead
// PULLS its rows straight out of the peers' partial buffers (P2P loads through NVSwitch; the buffers are mapped by CUDA IPC or,
// inside one process, by peer access), sums them in rank order -- deterministic given the partials -- and adds the rank-1 part
// of the mean-centred operand form in the same pass: one short, wide kernel (hundreds of CTAs, tens of microseconds).
//
// Synchronisation is by epoch flags in each rank's own memory, written by the peers with system-scope release stores:
//   arrive[p][rel]    rank p's partial of relation `rel` for iteration `epoch` is complete        (signal_arrive, after the product)
//   consumed[p][rel]  rank p has pulled its rows of my partial of iteration `epoch`              (last block of pull_reduce)
// A partial buffer is only zeroed for the next iteration once every peer has consumed it (zero_when_consumed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fz {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
  const float* B[kMaxPeers];          // every rank's partial buffer of ONE relation (full height), as mapped here
};
struct PeerFlags {
  unsigned long long* arrive[kMaxPeers];     // every rank's arrive array   [world][n_rel]
  unsigned long long* consumed[kMaxPeers];   // every rank's consumed array [world][n_rel]
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spin (with back-off and a ~60 s watchdog: a flag that never comes is a protocol bug or a dead peer -- trap, do not hang)
__device__ __forceinline__ void wait_flag(const unsigned long long* p, unsigned long long epoch) {
  if (ld_acquire_sys(p) >= epoch) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (unsigned spins = 1;; ++spins) {
    if (ld_acquire_sys(p) >= epoch) return;
    __nanosleep(200);
    if ((spins & 0xFFFu) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 60000000000ull) __trap();
    }
  }
}

// "my partial of relation `rel` is complete for `epoch`": one thread per peer (launched behind the product on its stream)
__global__ void signal_arrive(PeerFlags f, int world, int rank, int n_rel, int rel, unsigned long long epoch) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();
  st_release_sys(f.arrive[p] + (size_t)rank * n_rel + rel, epoch);
}

// Wait (one tiny block: it co-resides with anything, so no SM is held while the ranks' skew is being waited out) until every
// rank's partial of relation `rel` for `epoch` has arrived; the wide pull kernel is launched behind it on the same stream.
__global__ void wait_arrive(const unsigned long long* __restrict__ arrive_mine, int world, int n_rel, int rel, unsigned long long epoch) {
  if (threadIdx.x < world) wait_flag(arrive_mine + (size_t)threadIdx.x * n_rel + rel, epoch);
}

// B <- 0 once every peer has consumed the partial of the previous iteration (epoch_prev; 0 on the first)
__global__ void __launch_bounds__(256)
zero_when_consumed(float4* __restrict__ B, long long n_vec, const unsigned long long* __restrict__ consumed_mine, int world, int n_rel,
                   int rel, unsigned long long epoch_prev) {
  if (threadIdx.x < world) wait_flag(consumed_mine + (size_t)threadIdx.x * n_rel + rel, epoch_prev);
  __syncthreads();
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) B[i] = z;
}

// Bloc[r][q] = sum_p B_p[row0 + r][q]  (+ colsum[row0 + r] * centre[q]),  r < rows, q < k (k % 4 == 0), in rank order.
// The last block to finish tells every peer that this rank is done with their partial.
__global__ void __launch_bounds__(256)
pull_reduce(PeerPtrs src, PeerFlags f, float* __restrict__ Bloc, long long row0, long long rows, int k, const float* __restrict__ colsum,
            const float* __restrict__ centre, const unsigned long long* __restrict__ arrive_mine, unsigned int* __restrict__ done_ctr,
            int world, int rank, int n_rel, int rel, unsigned long long epoch) {
  // (the arrivals were waited for by wait_arrive, launched in front of this kernel; the check below is then a single load)
  if (threadIdx.x < world) wait_flag(arrive_mine + (size_t)threadIdx.x * n_rel + rel, epoch);
  __syncthreads();
  const int kv = k >> 2;
  const long long n_vec = rows * kv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / kv;
    const int q4 = (int)(i % kv) * 4;
    const long long off = (row0 + r) * k + q4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < world; ++p) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(src.B[p] + off));      // L2 only: peer lines must not linger in L1
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (colsum != nullptr) {
      const float cs = colsum[row0 + r];
      acc.x = fmaf(cs, centre[q4], acc.x);
      acc.y = fmaf(cs, centre[q4 + 1], acc.y);
      acc.z = fmaf(cs, centre[q4 + 2], acc.z);
      acc.w = fmaf(cs, centre[q4 + 3], acc.w);
    }
    *reinterpret_cast<float4*>(Bloc + r * k + q4) = acc;
  }
  // ---- last block: release the peers' buffers
  __shared__ unsigned int last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(done_ctr, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *done_ctr = 0;
    if (threadIdx.x < world) {
      __threadfence_system();
      st_release_sys(f.consumed[threadIdx.x] + (size_t)rank * n_rel + rel, epoch);
    }
  }
}

}  // namespace fz
