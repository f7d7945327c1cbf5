// The fp64 "small chain" of one DFMF iteration: everything that is k x k.
//   pinv_spd        P_t = pinv(nan_to_num(G_t^T G_t))                         (reference _dfmf.py:228-232)
//   backbone_chain  S_ij = P_i (G_i^T R_ij G_j) P_j ; S G^T G S^T terms       (reference _dfmf.py:236-239, 260, 272)
//   type_sums       per-type sums of the +/- parts of those k x k terms       (reference _dfmf.py:278-282)
// The reference gets P from scipy.linalg.pinv (SVD, cutoff max(M,N)*eps*sigma_max).  Here: Cholesky
// inverse when the Gram matrix is safely positive definite, otherwise a one-sided Jacobi eigen-solve
// with the same cutoff rule -- both in fp64, one CTA per matrix, matrices in global/L2 memory so any
// rank works.  SURVEY F7 explains why this chain cannot be fp32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fz_kernels.cuh"

namespace fz {

constexpr int kChainThreads = 512;

constexpr int kChainSmemDim = 64;                                         // operands up to 64 x 64 are staged in smem
constexpr int kChainSmemBytes = 2 * kChainSmemDim * (kChainSmemDim + 1) * 8;   // two padded operand tiles

// C[m x n] = op(A) * op(B), all row-major fp64; executed by the whole block; caller syncs afterwards.
// When both operands fit 64 x 64 they are first copied (already transposed as needed) into shared memory `stage`
// (kChainSmemBytes), which turns ~2*kk dependent global loads per output into conflict-free shared loads.
__device__ __forceinline__ void block_mm(double* C, int ldc, const double* A, int lda, bool ta, const double* B, int ldb,
                                         bool tb, int m, int n, int kk, double* stage = nullptr) {
  if (stage != nullptr && m <= kChainSmemDim && n <= kChainSmemDim && kk <= kChainSmemDim) {
    double(*sA)[kChainSmemDim + 1] = reinterpret_cast<double(*)[kChainSmemDim + 1]>(stage);                          // [m][kk]
    double(*sB)[kChainSmemDim + 1] = reinterpret_cast<double(*)[kChainSmemDim + 1]>(stage + kChainSmemDim * (kChainSmemDim + 1));  // [kk][n]
    __syncthreads();   // previous users of the staging area are done
    for (int o = threadIdx.x; o < m * kk; o += blockDim.x) {
      const int i = o / kk, c = o % kk;
      sA[i][c] = ta ? A[(long long)c * lda + i] : A[(long long)i * lda + c];
    }
    for (int o = threadIdx.x; o < kk * n; o += blockDim.x) {
      const int c = o / n, j = o % n;
      sB[c][j] = tb ? B[(long long)j * ldb + c] : B[(long long)c * ldb + j];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < m * n; o += blockDim.x) {
      const int i = o / n, j = o % n;
      double s = 0.0;
#pragma unroll 8
      for (int c = 0; c < kk; ++c) s += sA[i][c] * sB[c][j];
      C[(long long)i * ldc + j] = s;
    }
    return;
  }
  if (stage != nullptr && blockDim.x == kChainThreads) {
    // Ranks above 64: the same staging, tile by tile -- 64 x 64 output tiles, the reduction in chunks of 64 (ascending, so
    // every output is summed in the order of the plain loop below), 8 outputs per thread.  The naive form below pays two
    // dependent global loads per multiply-add and put the chain at 2.3 ms per iteration at rank 128.
    constexpr int kT = kChainSmemDim;
    constexpr int kPer = kT * kT / kChainThreads;
    double(*sA)[kT + 1] = reinterpret_cast<double(*)[kT + 1]>(stage);
    double(*sB)[kT + 1] = reinterpret_cast<double(*)[kT + 1]>(stage + kT * (kT + 1));
    for (int i0 = 0; i0 < m; i0 += kT)
      for (int j0 = 0; j0 < n; j0 += kT) {
        double acc[kPer];
#pragma unroll
        for (int e = 0; e < kPer; ++e) acc[e] = 0.0;
        for (int c0 = 0; c0 < kk; c0 += kT) {
          __syncthreads();   // previous users of the staging area are done
          for (int o = threadIdx.x; o < kT * kT; o += kChainThreads) {
            const int i = o / kT, c = o % kT;
            double a = 0.0, b = 0.0;
            if (i0 + i < m && c0 + c < kk) a = ta ? A[(long long)(c0 + c) * lda + (i0 + i)] : A[(long long)(i0 + i) * lda + (c0 + c)];
            // (same index pair read as B's [c][j]: i plays the reduction index, c the output column)
            if (c0 + i < kk && j0 + c < n) b = tb ? B[(long long)(j0 + c) * ldb + (c0 + i)] : B[(long long)(c0 + i) * ldb + (j0 + c)];
            sA[i][c] = a;
            sB[i][c] = b;
          }
          __syncthreads();
#pragma unroll
          for (int e = 0; e < kPer; ++e) {
            const int o = threadIdx.x + e * kChainThreads;
            const int i = o / kT, j = o % kT;
            double s = acc[e];
#pragma unroll 8
            for (int c = 0; c < kT; ++c) s += sA[i][c] * sB[c][j];
            acc[e] = s;
          }
        }
#pragma unroll
        for (int e = 0; e < kPer; ++e) {
          const int o = threadIdx.x + e * kChainThreads;
          const int i = i0 + o / kT, j = j0 + o % kT;
          if (i < m && j < n) C[(long long)i * ldc + j] = acc[e];
        }
      }
    return;
  }
  for (int o = threadIdx.x; o < m * n; o += blockDim.x) {
    const int i = o / n, j = o % n;
    double s = 0.0;
    for (int c = 0; c < kk; ++c) {
      const double a = ta ? A[(long long)c * lda + i] : A[(long long)i * lda + c];
      const double b = tb ? B[(long long)j * ldb + c] : B[(long long)c * ldb + j];
      s += a * b;
    }
    C[(long long)i * ldc + j] = s;
  }
}

struct PinvJob {
  const double* gram_raw;  // k x k (sum over chunks / ranks), not yet scrubbed
  double* gram;            // k x k scrubbed copy kept for the backbone chain
  double* P;               // k x k output
  double* work;            // 3 * k * k doubles
  int* info;               // [0] = 0 Cholesky, 1 Jacobi ; [1] = numerical rank
  double* cond;            // [1] estimate of cond_2(gram): |gram|_F |P|_F / sqrt(k) (Cholesky path; between cond_2 / sqrt(k) and
                           //     sqrt(k) cond_2, ~cond_2 for the usual one-dominant-direction Gram matrices) or
                           //     sigma_max / sigma_min over the kept singular values (Jacobi; 1e300 when rank-deficient)
  int k;
};

__global__ void __launch_bounds__(kChainThreads)
pinv_spd(const PinvJob* __restrict__ jobs) {
  const PinvJob job = jobs[blockIdx.x];
  const int k = job.k;
  const int tid = threadIdx.x, nth = blockDim.x;
  extern __shared__ double chain_stage[];
  const bool in_smem = (k <= kChainSmemDim);    // small ranks: factor and its inverse live in shared memory
  double* L = in_smem ? chain_stage : job.work;                              // k x k
  double* X = in_smem ? chain_stage + k * k : job.work + (long long)k * k;  // k x k
  double* V = job.work + 2ll * k * k;           // k x k (Jacobi)
  __shared__ double s_maxdiag;
  __shared__ int s_fail;
  __shared__ int s_rot;
  __shared__ double s_smax;

  for (int o = tid; o < k * k; o += nth) {
    const double v = scrub(job.gram_raw[o]);
    job.gram[o] = v;
    L[o] = v;
  }
  if (tid == 0) { s_fail = 0; s_maxdiag = 0.0; }
  __syncthreads();
  if (tid == 0) {
    double m = 0.0;
    for (int i = 0; i < k; ++i) m = fmax(m, fabs(L[(long long)i * k + i]));
    s_maxdiag = m;
    if (!(m > 0.0) || !(m < 1e300)) s_fail = 1;
  }
  __syncthreads();

  // ---------------------------------------------------------------- Cholesky (right-looking)
  if (!s_fail) {
    const double tol = 1e-10 * s_maxdiag;
    for (int j = 0; j < k; ++j) {
      if (tid == 0) {
        const double d = L[(long long)j * k + j];
        if (!(d > tol)) s_fail = 1;
        else L[(long long)j * k + j] = sqrt(d);
      }
      __syncthreads();
      if (s_fail) break;
      const double dj = L[(long long)j * k + j];
      for (int i = j + 1 + tid; i < k; i += nth) L[(long long)i * k + j] /= dj;
      __syncthreads();
      const int rem = k - j - 1;
      for (int o = tid; o < rem * rem; o += nth) {
        const int i = j + 1 + o / rem, c = j + 1 + o % rem;
        if (c <= i) L[(long long)i * k + c] -= L[(long long)i * k + j] * L[(long long)c * k + j];
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // A factorisation that went through says little about conditioning: pivots above 1e-10 max(diag) still allow
  // singular values below scipy's cut-off (k eps sigma_max), which pinv would DROP rather than invert.  The diagonal of
  // L bounds the condition number from below, cond_2 >= (max L_ii / min L_ii)^2: beyond 1 / (k eps) hand the matrix to
  // the eigen-solve, which applies the cut-off rule exactly.
  if (!s_fail) {
    if (tid == 0) {
      double lo = L[0], hi = L[0];
      for (int i = 1; i < k; ++i) { lo = fmin(lo, L[(long long)i * k + i]); hi = fmax(hi, L[(long long)i * k + i]); }
      const double c = (hi / lo) * (hi / lo);
      if (c * (double)k * 2.220446049250313e-16 > 1.0) s_fail = 1;
    }
    __syncthreads();
  }
  if (!s_fail) {
    // X = L^{-1} (lower triangular), one column per thread
    for (int j = tid; j < k; j += nth) {
      for (int i = 0; i < j; ++i) X[(long long)i * k + j] = 0.0;
      for (int i = j; i < k; ++i) {
        double s = (i == j) ? 1.0 : 0.0;
        for (int c = j; c < i; ++c) s -= L[(long long)i * k + c] * X[(long long)c * k + j];
        X[(long long)i * k + j] = s / L[(long long)i * k + i];
      }
    }
    __syncthreads();
    // P = X^T X
    for (int o = tid; o < k * k; o += nth) {
      const int a = o / k, b = o % k;
      double s = 0.0;
      for (int i = max(a, b); i < k; ++i) s += X[(long long)i * k + a] * X[(long long)i * k + b];
      job.P[o] = s;
    }
    if (tid == 0) { job.info[0] = 0; job.info[1] = k; }
    if (job.cond != nullptr) {
      // |gram|_F |P|_F / sqrt(k), summed by warp 0 in a fixed order (the estimate feeds a decision: keep it reproducible)
      __syncthreads();
      if (tid < 32) {
        double sg = 0.0, sp = 0.0;
        for (int o = tid; o < k * k; o += 32) {
          const double g = job.gram[o], q = job.P[o];
          sg += g * g;
          sp += q * q;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          sg += __shfl_xor_sync(0xffffffffu, sg, off);
          sp += __shfl_xor_sync(0xffffffffu, sp, off);
        }
        if (tid == 0) job.cond[0] = sqrt(sg * sp / (double)k);
      }
    }
    return;
  }

  // ---------------------------------------------------------------- one-sided Jacobi on the rows of W = A
  // (A symmetric: rows == columns).  After convergence row p of W is (A v_p)^T with v_p = row p of V,
  // sigma_p = |row p|, and pinv(A) = sum_{sigma_p > cutoff} v_p (A v_p)^T / sigma_p^2.
  double* W = job.work;                 // the eigen-solve always runs out of the global workspace
  X = job.work + (long long)k * k;
  for (int o = tid; o < k * k; o += nth) {
    W[o] = job.gram[o];
    V[o] = (o / k == o % k) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
  const int kk = (k + 1) & ~1;            // players (dummy when k is odd)
  const int npairs = kk / 2;
  for (int sweep = 0; sweep < 40; ++sweep) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int round = 0; round < kk - 1; ++round) {
      for (int pi = warp; pi < npairs; pi += nwarps) {
        int p, q;
        if (pi == 0) { p = kk - 1; q = round; }
        else {
          p = (round + pi) % (kk - 1);
          q = (round - pi + (kk - 1)) % (kk - 1);
        }
        if (p >= k || q >= k) continue;
        if (p > q) { const int t = p; p = q; q = t; }
        double* wp = W + (long long)p * k;
        double* wq = W + (long long)q * k;
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int c = lane; c < k; c += 32) {
          const double a = wp[c], b = wq[c];
          al += a * a; be += b * b; ga += a * b;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          al += __shfl_xor_sync(0xffffffffu, al, off);
          be += __shfl_xor_sync(0xffffffffu, be, off);
          ga += __shfl_xor_sync(0xffffffffu, ga, off);
        }
        if (fabs(ga) > 1e-15 * sqrt(al * be) && fabs(ga) > 0.0) {
          const double zeta = (be - al) / (2.0 * ga);
          const double t = ((zeta >= 0.0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
          double* vp = V + (long long)p * k;
          double* vq = V + (long long)q * k;
          for (int c = lane; c < k; c += 32) {
            const double a = wp[c], b = wq[c];
            wp[c] = cs * a - sn * b;
            wq[c] = sn * a + cs * b;
            const double x = vp[c], y = vq[c];
            vp[c] = cs * x - sn * y;
            vq[c] = sn * x + cs * y;
          }
          if (lane == 0) atomicAdd(&s_rot, 1);
        }
      }
      __syncthreads();
    }
    if (s_rot == 0) break;
    __syncthreads();
  }
  // sigma_p^2 into X[p], sigma_max
  for (int p = tid; p < k; p += nth) {
    double s = 0.0;
    for (int c = 0; c < k; ++c) s += W[(long long)p * k + c] * W[(long long)p * k + c];
    X[p] = s;
  }
  __syncthreads();
  if (tid == 0) {
    double m = 0.0;
    for (int p = 0; p < k; ++p) m = fmax(m, X[p]);
    s_smax = sqrt(m);
    const double cut = (double)k * 2.220446049250313e-16 * s_smax;
    int rank = 0;
    double smin = s_smax;
    for (int p = 0; p < k; ++p) {
      if (sqrt(X[p]) > cut) { ++rank; smin = fmin(smin, sqrt(X[p])); }
    }
    job.info[0] = 1;
    job.info[1] = rank;
    if (job.cond) job.cond[0] = (rank == k && smin > 0.0) ? s_smax / smin : 1e300;
  }
  __syncthreads();
  const double cutoff = (double)k * 2.220446049250313e-16 * s_smax;
  for (int o = tid; o < k * k; o += nth) {
    const int a = o / k, b = o % k;
    double s = 0.0;
    for (int p = 0; p < k; ++p) {
      const double s2 = X[p];
      if (sqrt(s2) > cutoff) s += V[(long long)p * k + a] * W[(long long)p * k + b] / s2;
    }
    job.P[o] = s;
  }
}

template <class T>
struct BackboneJob {
  const double* M_raw;   // k_i x k_j  = G_i^T R_ij G_j  (summed over chunks / ranks)
  const double* P_i;
  const double* P_j;
  const double* gram_i;
  const double* gram_j;
  double* S;             // k_i x k_j
  double* t2;            // k_i x k_i  = S G_j^T G_j S^T
  double* t5;            // k_j x k_j  = S^T G_i^T G_i S
  T* W1;                 // k_j x k_i  = S^T in the compute dtype  (tmp1 = A S^T)
  T* W4;                 // k_i x k_j  = S                         (tmp4 = B S)
  double* work;          // 6 * max(k_i,k_j)^2 (three matrices per CTA of the relation's pair)
  int ki, kj;
  int solve;             // 1: compute S from M_raw; 0: S is given (transform)
  int scrub;             // dfmf: nan_to_num on every k x k product
};

// grid = (relations, 2): both CTAs of a relation solve S (cheap, and it saves a grid-wide dependency); CTA 0 publishes S and
// the compute-dtype copies and forms t2 = S gram_j S^T, CTA 1 forms t5 = S^T gram_i S -- the chain sits on the critical path
// of every iteration (sharded runs: right behind the all-reduce), so its two independent halves run side by side.
template <class T>
__global__ void __launch_bounds__(kChainThreads)
backbone_chain(const BackboneJob<T>* __restrict__ jobs) {
  const BackboneJob<T> job = jobs[blockIdx.x];
  const int half = blockIdx.y;
  const int ki = job.ki, kj = job.kj;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int km = max(ki, kj);
  double* U = job.work + (long long)half * 3 * km * km;
  double* Vw = U + (long long)km * km;
  double* Sl = Vw + (long long)km * km;                      // CTA 1's private copy of S
  double* S = (half == 0 || !job.solve) ? job.S : Sl;
  extern __shared__ double chain_stage[];
  double* stage = chain_stage;       // operands beyond 64 x 64 are tiled through it
  if (job.solve) {
    // Vw = scrub(M) ; U = Vw * P_j ; S = scrub(P_i * U)
    for (int o = tid; o < ki * kj; o += nth) Vw[o] = scrub(job.M_raw[o]);
    __syncthreads();
    block_mm(U, kj, Vw, kj, false, job.P_j, kj, false, ki, kj, kj, stage);
    __syncthreads();
    block_mm(S, kj, job.P_i, ki, false, U, kj, false, ki, kj, ki, stage);
    __syncthreads();
    for (int o = tid; o < ki * kj; o += nth) S[o] = scrub(S[o]);
    __syncthreads();
  }
  if (half == 0) {
    // t2 = S gram_j S^T
    block_mm(U, kj, S, kj, false, job.gram_j, kj, false, ki, kj, kj, stage);
    __syncthreads();
    block_mm(job.t2, ki, U, kj, false, S, kj, true, ki, ki, kj, stage);
    __syncthreads();
    if (job.scrub)
      for (int o = tid; o < ki * ki; o += nth) job.t2[o] = scrub(job.t2[o]);
    for (int o = tid; o < ki * kj; o += nth) {
      const int a = o / kj, b = o % kj;
      const double s = S[o];
      job.W4[o] = (T)s;
      job.W1[(long long)b * ki + a] = (T)s;
    }
  } else {
    // t5 = S^T gram_i S
    block_mm(U, ki, S, kj, true, job.gram_i, ki, false, kj, ki, ki, stage);
    __syncthreads();
    block_mm(job.t5, kj, U, ki, false, S, kj, false, kj, kj, ki, stage);
    __syncthreads();
    if (job.scrub)
      for (int o = tid; o < kj * kj; o += nth) job.t5[o] = scrub(job.t5[o]);
  }
}

// Squared Frobenius residual of one relation from k x k quantities only (reference objective, _dfmf.py:306-319, without
// forming the n_i x n_j reconstruction):
//     ||R - G_i S G_j^T||_F^2 = ||R||_F^2 - 2 tr(S^T (G_i^T R G_j)) + tr(S^T (G_i^T G_i) S (G_j^T G_j))
// with M = G_i^T R G_j and the Gram matrices of the CURRENT factors (the raw sums the products phase leaves in the all-reduce
// buffer -- on a sharded handle they are already summed over the ranks).  One CTA per relation, fp64.
struct TraceJob {
  const double* M;        // k_i x k_j
  const double* gram_i;   // k_i x k_i
  const double* gram_j;   // k_j x k_j
  const double* S;        // k_i x k_j
  const double* rnorm2;   // [1]  ||R||_F^2
  double* work;           // 2 * max(k_i, k_j)^2
  double* out;            // [1]
  int ki, kj;
};
__global__ void __launch_bounds__(kChainThreads)
trace_objective(const TraceJob* __restrict__ jobs) {
  const TraceJob job = jobs[blockIdx.x];
  const int ki = job.ki, kj = job.kj, tid = threadIdx.x, nth = blockDim.x;
  const int km = max(ki, kj);
  double* U = job.work;
  double* V = job.work + (long long)km * km;
  extern __shared__ double chain_stage[];
  double* stage = chain_stage;
  block_mm(U, kj, job.gram_i, ki, false, job.S, kj, false, ki, kj, ki, stage);      // U = Gram_i S
  __syncthreads();
  block_mm(V, kj, U, kj, false, job.gram_j, kj, false, ki, kj, kj, stage);          // V = U Gram_j
  __syncthreads();
  __shared__ double red[kChainThreads];
  double acc = 0.0;
  for (int o = tid; o < ki * kj; o += nth) acc += job.S[o] * (V[o] - 2.0 * job.M[o]);
  red[tid] = acc;
  __syncthreads();
  for (int w = nth >> 1; w > 0; w >>= 1) {
    if (tid < w) red[tid] += red[tid + w];
    __syncthreads();
  }
  if (tid == 0) job.out[0] = job.rnorm2[0] + red[0];
}

template <class T>
struct TypeSumJob {
  const double* const* mats;  // n_mats pointers to k x k matrices (t2 of row-role relations, t5 of column-role ones)
  T* Nsum;                    // k x k : sum of magnitudes of negative parts  (feeds the numerator)
  T* Dsum;                    // k x k : sum of positive parts                (feeds the denominator)
  int n_mats;
  int k;
};

template <class T>
__global__ void __launch_bounds__(256)
type_sums(const TypeSumJob<T>* __restrict__ jobs) {
  const TypeSumJob<T> job = jobs[blockIdx.x];
  for (int o = threadIdx.x; o < job.k * job.k; o += blockDim.x) {
    double n = 0.0, d = 0.0;
    for (int m = 0; m < job.n_mats; ++m) {
      double p, q;
      sign_split(job.mats[m][o], p, q);
      d += p;
      n += q;
    }
    job.Nsum[o] = (T)n;
    job.Dsum[o] = (T)d;
  }
}

}  // namespace fz
