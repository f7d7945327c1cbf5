// CUDA-core kernels of the fusion engine (everything that is not the streamed tensor-core product):
//   gemm_simt        exact fp32 / fp64 tiled products  R*G, R^T*G, Theta(+/-)*G, G*S      (a6, a7, a8)
//   gram_partial     fp64-accumulated k x k reductions  G^T G  and  G_i^T A_ij            (a5, a6)
//   fused_update     per-type multiplicative update                                         (a7, a9)
//   split_factor     fp32 factor -> bf16 split terms (tensor-core operand form)
//   impute / mask    dfmc re-imputation of unknown entries                                  (a11)
//   recon_err        ||R - G_i S G_j^T||_F^2                                               (a10)
// Row references are to SURVEY.md §8(a); file:line citations of the reference are in the engine.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fz {

// numpy.nan_to_num in the compute dtype: NaN -> 0, +-inf -> +-max finite (_dfmf.py:27,255).
template <class T> struct Lim;
template <> struct Lim<float> { static __device__ __forceinline__ float max() { return 3.402823466e+38f; } };
template <> struct Lim<double> { static __device__ __forceinline__ double max() { return 1.7976931348623157e+308; } };

template <class T>
__device__ __forceinline__ T scrub(T x) {
  if (x != x) return T(0);
  if (x > Lim<T>::max()) return Lim<T>::max();
  if (x < -Lim<T>::max()) return -Lim<T>::max();
  return x;
}
// the reference's literal sign split  t = x > 0;  pos = t*x;  neg = (t-1)*x   (_dfmf.py:256-258)
template <class T>
__device__ __forceinline__ void sign_split(T x, T& pos, T& neg) {
  const T t = (x > T(0)) ? T(1) : T(0);
  pos = t * x;
  neg = (t - T(1)) * x;
}

template <class XT, class T> __device__ __forceinline__ T load_as(const XT* p) { return static_cast<T>(*p); }
template <> __device__ __forceinline__ float load_as<__nv_bfloat16, float>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ double load_as<__nv_bfloat16, double>(const __nv_bfloat16* p) { return (double)__bfloat162float(*p); }

// ------------------------------------------------------------------------------------------------
// C[M x N] (+)= op(X) * Y      op(X) = X (M x K, row-major)  or  X^T (X is K x M, row-major)
// kSplit: two outputs  C += max(X,0)*Y,  C2 += max(-X,0)*Y   (constraint matrices, _dfmf.py:203-208)
// ------------------------------------------------------------------------------------------------
constexpr int kGemmBM = 64, kGemmBN = 64, kGemmBK = 16;

template <class T, class XT, bool kTrans, bool kSplit>
__global__ void __launch_bounds__(256)
gemm_simt(const XT* __restrict__ X, long long ldx, const T* __restrict__ Y, long long ldy, T* __restrict__ C,
          T* __restrict__ C2, long long ldc, int M, int N, int K, int accumulate) {
  __shared__ T Xs[kGemmBK][kGemmBM + 4];
  __shared__ T Xn[kSplit ? kGemmBK : 1][kSplit ? kGemmBM + 4 : 1];
  __shared__ T Ys[kGemmBK][kGemmBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * kGemmBM, n0 = blockIdx.y * kGemmBN;
  const int ty = tid / 16, tx = tid % 16;
  T acc[4][4], acc2[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = T(0); acc2[i][j] = T(0); }

  for (int k0 = 0; k0 < K; k0 += kGemmBK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int mm, kk;
      if (!kTrans) { mm = idx / kGemmBK; kk = idx % kGemmBK; }
      else         { kk = idx / kGemmBM; mm = idx % kGemmBM; }
      T v = T(0);
      if (m0 + mm < M && k0 + kk < K)
        v = kTrans ? load_as<XT, T>(X + (long long)(k0 + kk) * ldx + (m0 + mm))
                   : load_as<XT, T>(X + (long long)(m0 + mm) * ldx + (k0 + kk));
      if (kSplit) { T p, n; sign_split(v, p, n); Xs[kk][mm] = p; Xn[kk][mm] = n; }
      else Xs[kk][mm] = v;
      const int yk = idx / kGemmBN, yn = idx % kGemmBN;
      T w = T(0);
      if (k0 + yk < K && n0 + yn < N) w = Y[(long long)(k0 + yk) * ldy + (n0 + yn)];
      Ys[yk][yn] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGemmBK; ++kk) {
      T a[4], b[4], a2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = Xs[kk][ty * 4 + i]; if (kSplit) a2[i] = Xn[kk][ty * 4 + i]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ys[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] += a[i] * b[j];
          if (kSplit) acc2[i][j] += a2[i] * b[j];
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const long long o = (long long)m * ldc + n;
      C[o] = accumulate ? C[o] + acc[i][j] : acc[i][j];
      if (kSplit) C2[o] = accumulate ? C2[o] + acc2[i][j] : acc2[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// part[chunk][ka][kb] = sum over the chunk's rows of X[r][a] * Y[r][b], accumulated in fp64 from the
// exact products of the stored values (F7: the k x k chain must not see fp32 reduction error).
// grid = (chunks, ceil(ka/64), ceil(kb/64)), 256 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
gram_partial(const T* __restrict__ X, long long ldx, const T* __restrict__ Y, long long ldy, double* __restrict__ part,
             long long n_rows, int ka, int kb, int rows_per_chunk, int scrub_inputs) {
  // 16-row slabs, double-buffered in shared memory; the next slab's global loads are issued into registers
  // before the current slab is multiplied, so DRAM/L2 latency overlaps the fp64 FMAs.
  __shared__ double Xs[2][16][64 + 2];
  __shared__ double Ys[2][16][64 + 2];
  const int tid = threadIdx.x;
  const int a0 = blockIdx.y * 64, b0 = blockIdx.z * 64;
  const long long r_begin = (long long)blockIdx.x * rows_per_chunk;
  const long long r_end = min(n_rows, r_begin + rows_per_chunk);
  const int ty = tid / 16, tx = tid % 16;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  double px[4], py[4];
  auto fetch = [&](long long r0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      const int rr = idx / 64, cc = idx % 64;
      double xv = 0.0, yv = 0.0;
      if (r0 + rr < r_end) {
        if (a0 + cc < ka) xv = (double)X[(r0 + rr) * ldx + a0 + cc];
        if (b0 + cc < kb) yv = (double)Y[(r0 + rr) * ldy + b0 + cc];
        if (scrub_inputs) { xv = scrub(xv); yv = scrub(yv); }
      }
      px[e] = xv;
      py[e] = yv;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      Xs[buf][idx / 64][idx % 64] = px[e];
      Ys[buf][idx / 64][idx % 64] = py[e];
    }
  };
  int buf = 0;
  if (r_begin < r_end) fetch(r_begin);
  for (long long r0 = r_begin; r0 < r_end; r0 += 16) {
    stash(buf);
    __syncthreads();
    if (r0 + 16 < r_end) fetch(r0 + 16);
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Xs[buf][rr][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ys[buf][rr][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    buf ^= 1;   // the other buffer was last read two iterations ago, behind the barrier above
  }
  double* out = part + (long long)blockIdx.x * ka * kb;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + ty * 4 + i;
    if (a >= ka) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = b0 + tx * 4 + j;
      if (b < kb) out[(long long)a * kb + b] = acc[i][j];
    }
  }
}

// The same reduction on the fp64 tensor cores (mma.sync.m8n8k4.f64: DMMA).  Same grid, same partial layout, same 16-row slabs
// and register prefetch; each of the 8 warps owns 8 rows of the 64 x 64 output tile and sweeps its 8 column tiles, so one
// shared-memory load of the A fragment feeds 8 MMAs (the CUDA-core version reads 8 doubles per 16 FMAs and runs at ~7 TFLOP/s).
// Products are exact, accumulation is IEEE fp64: the result equals the CUDA-core kernel's up to the order of the fp64 adds.
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <class T>
__global__ void __launch_bounds__(256)
gram_partial_dmma(const T* __restrict__ X, long long ldx, const T* __restrict__ Y, long long ldy, double* __restrict__ part,
                  long long n_rows, int ka, int kb, int rows_per_chunk) {
  __shared__ double Xs[2][16][64 + 4];      // row pitch 68 doubles: the 4 k-rows of a fragment land 8 banks apart
  __shared__ double Ys[2][16][64 + 4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a0 = blockIdx.y * 64, b0 = blockIdx.z * 64;
  const long long r_begin = (long long)blockIdx.x * rows_per_chunk;
  const long long r_end = min(n_rows, r_begin + rows_per_chunk);
  double acc[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
  double px[4], py[4];
  auto fetch = [&](long long r0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      const int rr = idx / 64, cc = idx % 64;
      double xv = 0.0, yv = 0.0;
      if (r0 + rr < r_end) {
        if (a0 + cc < ka) xv = (double)X[(r0 + rr) * ldx + a0 + cc];
        if (b0 + cc < kb) yv = (double)Y[(r0 + rr) * ldy + b0 + cc];
      }
      px[e] = xv;
      py[e] = yv;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      Xs[buf][idx / 64][idx % 64] = px[e];
      Ys[buf][idx / 64][idx % 64] = py[e];
    }
  };
  int buf = 0;
  if (r_begin < r_end) fetch(r_begin);
  const int kr = lane & 3, mc = lane >> 2;      // fragment coordinates: k-row inside the step, row (A) / column (B) of the 8 x 8 tile
  for (long long r0 = r_begin; r0 < r_end; r0 += 16) {
    stash(buf);
    __syncthreads();
    if (r0 + 16 < r_end) fetch(r0 + 16);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const double a = Xs[buf][4 * ks + kr][8 * warp + mc];
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma_m8n8k4(acc[j], a, Ys[buf][4 * ks + kr][8 * j + mc]);
    }
    buf ^= 1;
  }
  double* out = part + (long long)blockIdx.x * ka * kb;
  const int a = a0 + 8 * warp + mc;
  if (a < ka) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = b0 + 8 * j + 2 * kr;
      if (b < kb) out[(long long)a * kb + b] = acc[j][0];
      if (b + 1 < kb) out[(long long)a * kb + b + 1] = acc[j][1];
    }
  }
}

// out[i] = sum_c part[c][i], deterministic: block = 32 consecutive elements x 32 chunk lanes (warp w sums the chunks
// w, w+32, ... in order, four independent loads in flight), then the 32 lane sums are added in fixed order.
// grid = ceil(elems / 32), 1024 threads.
constexpr int kRedThreads = 1024;
__global__ void __launch_bounds__(kRedThreads)
reduce_partials(const double* __restrict__ part, double* __restrict__ out, int n_chunks, long long elems) {
  __shared__ double lanes[32][33];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * 32 + l;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (i < elems) {
    int c = w;
    for (; c + 96 < n_chunks; c += 128) {
      const double v0 = part[(long long)c * elems + i], v1 = part[(long long)(c + 32) * elems + i];
      const double v2 = part[(long long)(c + 64) * elems + i], v3 = part[(long long)(c + 96) * elems + i];
      s0 += v0; s1 += v1; s2 += v2; s3 += v3;
    }
    for (; c < n_chunks; c += 32) s0 += part[(long long)c * elems + i];
  }
  lanes[w][l] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && i < elems) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) s += lanes[k][l];
    out[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// Fused per-type multiplicative update (reference _dfmf.py:246-296):
//   num = sum_terms pos(X_r W_r) + G Nsum + sum_add addN ;  den = sum_terms neg(X_r W_r) + G Dsum + sum_add addD
//   Gnew = G * sqrt(num / max(den, eps64))
// A "term" is one incident relation: X_r = A_ij (rows of this type) with W_r = S_ij^T, or X_r = B_ij with
// W_r = S_ij.  Nsum/Dsum are the per-type sums of the negative/positive parts of S G^T G S^T (k x k).
// Additive pairs carry the constraint products (Theta^- G, Theta^+ G) or transform's frozen terms.
// ------------------------------------------------------------------------------------------------
template <class T>
struct UpdTerm {
  const T* X;      // rows x kx
  long long ldx;
  const T* W;      // kx x kt, row-major
  int kx;
  int pad_;
};
template <class T>
struct UpdAdd {
  const T* num;    // rows x kt (ld = kt)
  const T* den;
};
template <class T>
struct UpdArgs {
  const T* G;      // rows x kt (old factor, local rows)
  T* Gnew;
  long long ldg;
  long long rows;
  int kt;
  int n_terms;
  int n_adds;
  int scrub_terms;  // dfmf: nan_to_num on tmp1/tmp4 (and on the k x k terms upstream); dfmc/transform: no
  const UpdTerm<T>* terms;
  const UpdAdd<T>* adds;
  const T* Nsum;   // kt x kt
  const T* Dsum;
};

constexpr int kUpdRows = 64;    // rows of the factor per block
// reduction slab staged in shared memory (double-buffered): 32 for fp32, 16 for fp64 (48 KB of static shared memory)
template <class T> struct UpdCfg { static constexpr int kSlab = sizeof(T) == 4 ? 32 : 16; };

// One (64 rows x 64 cols) output tile per block and q-tile; 256 threads, 4 x 4 outputs per thread:
// per reduction step 4 broadcast reads of X and one 16-byte read of W feed 16 FMAs.
template <class T>
__device__ __forceinline__ void upd_accumulate(T (&acc)[4][4], const T (*Xs)[UpdCfg<T>::kSlab + 1], const T (*Ws)[64 + 4], int ty, int tx,
                                               int cmax) {
  for (int cc = 0; cc < cmax; ++cc) {
    T x[4], w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = Xs[ty * 4 + i][cc];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = Ws[cc][tx * 4 + j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += x[i] * w[j];
  }
}

template <class T>
__global__ void __launch_bounds__(256)
fused_update(const UpdArgs<T> a) {
  // The work of a block is a flat list of SLAB JOBS -- (operand X, weight W, reduction offset) -- over the relation terms and
  // the two G * sum passes.  The next job's global loads are issued into registers before the current slab is multiplied
  // (double-buffered shared memory), so L2 / HBM latency overlaps the FMAs instead of being paid once per slab.
  constexpr int kUpdSlab = UpdCfg<T>::kSlab;
  constexpr int kE = kUpdSlab * 64 / 256;     // elements of each tile a thread moves
  __shared__ T Xs[2][kUpdRows][kUpdSlab + 1];
  __shared__ __align__(16) T Ws[2][kUpdSlab][64 + 4];
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * kUpdRows;
  const int ty = tid / 16, tx = tid % 16;
  const T eps = (T)2.220446049250313e-16;

  for (int q0 = 0; q0 < a.kt; q0 += 64) {
    T num[4][4], den[4][4], acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { num[i][j] = T(0); den[i][j] = T(0); acc[i][j] = T(0); }

    const int slabs_g = (a.kt + kUpdSlab - 1) / kUpdSlab;
    // job cursor: term index (n_terms = G*Nsum, n_terms + 1 = G*Dsum) and slab inside it
    int jt = 0, jc = 0;
    auto job_valid = [&](int t) { return t < a.n_terms + 2; };
    auto job_slabs = [&](int t) { return t < a.n_terms ? (a.terms[t].kx + kUpdSlab - 1) / kUpdSlab : slabs_g; };
    // skip empty terms (kx == 0 cannot happen; kept cheap)
    T px[kE], pw[kE];
    auto fetch = [&](int t, int c) {
      const T* X; const T* W; long long ldx; int kx;
      if (t < a.n_terms) { const UpdTerm<T> term = a.terms[t]; X = term.X; W = term.W; ldx = term.ldx; kx = term.kx; }
      else { X = a.G; W = (t == a.n_terms) ? a.Nsum : a.Dsum; ldx = a.ldg; kx = a.kt; }
      const int c0 = c * kUpdSlab;
#pragma unroll
      for (int e = 0; e < kE; ++e) {
        const int idx = tid + e * 256;
        const int rr = idx / kUpdSlab, cc = idx % kUpdSlab;
        px[e] = (r0 + rr < a.rows && c0 + cc < kx) ? X[(r0 + rr) * ldx + c0 + cc] : T(0);
        const int wc = idx / 64, wq = idx % 64;
        pw[e] = (c0 + wc < kx && q0 + wq < a.kt) ? W[(long long)(c0 + wc) * a.kt + q0 + wq] : T(0);
      }
    };
    auto stash = [&](int buf) {
#pragma unroll
      for (int e = 0; e < kE; ++e) {
        const int idx = tid + e * 256;
        Xs[buf][idx / kUpdSlab][idx % kUpdSlab] = px[e];
        Ws[buf][idx / 64][idx % 64] = pw[e];
      }
    };
    int buf = 0;
    fetch(jt, jc);
    while (job_valid(jt)) {
      stash(buf);
      __syncthreads();
      // advance the cursor and prefetch the next job
      int nt = jt, nc = jc + 1;
      if (nc >= job_slabs(jt)) { nt = jt + 1; nc = 0; }
      if (job_valid(nt)) fetch(nt, nc);
      upd_accumulate<T>(acc, Xs[buf], Ws[buf], ty, tx, kUpdSlab);     // zero padding beyond kx adds nothing
      if (nt != jt) {            // the term is complete: fold it into num / den
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            T v = acc[i][j];
            acc[i][j] = T(0);
            if (jt < a.n_terms) {
              if (a.scrub_terms) v = scrub(v);
              T pp, nn;
              sign_split(v, pp, nn);
              num[i][j] += pp;
              den[i][j] += nn;
            } else if (jt == a.n_terms) num[i][j] += v;
            else den[i][j] += v;
          }
      }
      jt = nt;
      jc = nc;
      buf ^= 1;     // the other buffer was last read one job ago, behind this job's barrier
    }
    // ---- additive pairs + the update itself
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long row = r0 + ty * 4 + i;
      if (row >= a.rows) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = q0 + tx * 4 + j;
        if (q >= a.kt) continue;
        T nu = num[i][j], de = den[i][j];
        for (int s = 0; s < a.n_adds; ++s) {
          nu += a.adds[s].num[row * a.kt + q];
          de += a.adds[s].den[row * a.kt + q];
        }
        const T g = a.G[row * a.ldg + q];
        a.Gnew[row * a.ldg + q] = g * sqrt(nu / max(de, eps));
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// replacing unknown (non-finite) entries of a device-resident relation before the fit
// (reference fill_mean / fill_row / fill_col / fill_const, skfusion/fusion/base/fusion_graph.py:464-510).
// numpy.nanmean semantics: NaN entries are skipped, +-inf take part in the mean.
// ------------------------------------------------------------------------------------------------
template <class XT> __device__ __forceinline__ XT store_as(double v) { return static_cast<XT>(v); }
template <> __device__ __forceinline__ __nv_bfloat16 store_as<__nv_bfloat16>(double v) { return __float2bfloat16_rn((float)v); }

// per row: sum and count of the non-NaN entries (one warp per row, fixed-order shuffle reduction)
template <class XT>
__global__ void row_nan_stats(const XT* __restrict__ X, long long ld, long long rows, long long cols, double* __restrict__ sum,
                              double* __restrict__ cnt) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0, c = 0.0;
  for (long long j = lane; j < cols; j += 32) {
    const double v = load_as<XT, double>(X + r * ld + j);
    if (v == v) { s += v; c += 1.0; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if (lane == 0) { sum[r] = s; cnt[r] = c; }
}
// per column and row chunk: part_sum[chunk][col], part_cnt[chunk][col]
template <class XT>
__global__ void col_nan_stats_partial(const XT* __restrict__ X, long long ld, long long rows, long long cols, long long rows_per_chunk,
                                      double* __restrict__ part_sum, double* __restrict__ part_cnt) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  double s = 0.0, n = 0.0;
  for (long long r = r0; r < r1; ++r) {
    const double v = load_as<XT, double>(X + r * ld + c);
    if (v == v) { s += v; n += 1.0; }
  }
  part_sum[(long long)blockIdx.y * cols + c] = s;
  part_cnt[(long long)blockIdx.y * cols + c] = n;
}
__global__ void sum_chunks(const double* __restrict__ part, double* __restrict__ out, int chunks, long long cols) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double s = 0.0;
  for (int k = 0; k < chunks; ++k) s += part[(long long)k * cols + c];
  out[c] = s;
}
// one block: total[0] = sum(sum[i]), total[1] = sum(cnt[i]) in a fixed order; then mean[i] = sum[i]/cnt[i], or the overall
// mean where a row / column has no known entry (NaN mean)
__global__ void finish_means(double* __restrict__ sum, const double* __restrict__ cnt, long long n, double* __restrict__ total) {
  __shared__ double ss[1024], sc[1024];
  double s = 0.0, c = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) { s += sum[i]; c += cnt[i]; }
  ss[threadIdx.x] = s; sc[threadIdx.x] = c;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { ss[threadIdx.x] += ss[threadIdx.x + o]; sc[threadIdx.x] += sc[threadIdx.x + o]; }
    __syncthreads();
  }
  const double overall = ss[0] / sc[0];
  if (threadIdx.x == 0) { total[0] = overall; }
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double m = sum[i] / cnt[i];
    sum[i] = (m == m) ? m : overall;
  }
}
// mask[r][c] = 1 where X[r][c] is not finite (NaN / +-inf), else 0: the completion mask of a device-resident relation
// (the reference takes it from numpy masked arrays, decomposition/dfmc.py:69-94)
template <class XT>
__global__ void unknown_mask(const XT* __restrict__ X, long long ld, long long rows, long long cols, uint8_t* __restrict__ mask,
                             long long mld) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols, c = idx % cols;
  const double v = load_as<XT, double>(X + r * ld + c);
  mask[r * mld + c] = (v == v && v - v == 0.0) ? 0 : 1;
}
// mode 0: every unknown <- scalar[0];  1: <- per_axis[row];  2: <- per_axis[col];  3: <- value
template <class XT>
__global__ void replace_unknown(XT* __restrict__ X, long long ld, long long rows, long long cols, int mode, const double* __restrict__ scalar,
                                const double* __restrict__ per_axis, double value) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols, c = idx % cols;
  const double v = load_as<XT, double>(X + r * ld + c);
  if (v == v && v - v == 0.0) return;            // finite
  double w = value;
  if (mode == 0) w = scalar[0];
  else if (mode == 1) w = per_axis[r];
  else if (mode == 2) w = per_axis[c];
  X[r * ld + c] = store_as<XT>(w);
}

// ------------------------------------------------------------------------------------------------
// factor initialisation on the device (reference _init.py:20-61)
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void fill_value(T* __restrict__ dst, T value, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = value;
}
// W[idx[c][s]][c] = 1   (W is n_other x k, leading dimension ldw)
template <class T>
__global__ void scatter_ones(T* __restrict__ W, long long ldw, const int32_t* __restrict__ idx, int k, int p_c, long long n_other) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)k * p_c) return;
  const int c = (int)(i / p_c);
  const long long row = idx[i];
  if (row >= 0 && row < n_other) W[row * ldw + c] = T(1);
}
// G += | P / count |      (numpy: mean = sum / count; count == 0 gives NaN like the mean of an empty slice)
template <class T>
__global__ void add_abs_mean(const T* __restrict__ P, T* __restrict__ G, long long n, T count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T m = P[i] / count;
  G[i] += (m < T(0)) ? -m : m;
}
// partial sums of squares of the columns over a chunk of rows: part[chunk][col]
template <class XT>
__global__ void col_sumsq_partial(const XT* __restrict__ X, long long ld, long long rows, long long cols, long long rows_per_chunk,
                                  double* __restrict__ part) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  double s = 0.0;
  for (long long r = r0; r < r1; ++r) {
    const double v = load_as<XT, double>(X + r * ld + c);
    s += v * v;
  }
  part[(long long)blockIdx.y * cols + c] = s;
}
// out[0] = sum of v[0..n) in a fixed order (one block)
__global__ void __launch_bounds__(1024)
total_sum(const double* __restrict__ v, long long n, double* __restrict__ out) {
  __shared__ double red[1024];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) s += v[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}
__global__ void sqrt_of_chunk_sums(const double* __restrict__ part, double* __restrict__ out, int chunks, long long cols) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double s = 0.0;
  for (int k = 0; k < chunks; ++k) s += part[(long long)k * cols + c];
  out[c] = sqrt(s);
}
__global__ void sqrt_inplace(double* __restrict__ v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = sqrt(v[i]);
}
// one warp per row, lanes stride over the columns, fixed-order shuffle reduction
template <class XT>
__global__ void row_norms(const XT* __restrict__ X, long long ld, long long rows, long long cols, double* __restrict__ out) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (long long c = lane; c < cols; c += 32) {
    const double v = load_as<XT, double>(X + r * ld + c);
    s += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = sqrt(s);
}

// ------------------------------------------------------------------------------------------------
// Factor -> tensor-core operand form: Gs[r][t*kp + q] = bf16 term t of the running residual of G[r][q].
// Rows >= n_valid and columns >= k are zero.  (Own design: no counterpart in the reference.)
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void split_factor(const T* __restrict__ G, long long ldg, __nv_bfloat16* __restrict__ Gs, long long n_valid,
                             long long n_pad, int k, int kp, int terms, const float* __restrict__ centre = nullptr) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pad * kp) return;
  const long long r = idx / kp;
  const int q = (int)(idx % kp);
  float v = 0.f;
  if (r < n_valid && q < k) {
    v = (float)G[r * ldg + q];
    if (centre != nullptr) v -= centre[q];      // mean-centred operand form: G = 1 c^T + D, the terms represent D
  }
  __nv_bfloat16* out = Gs + r * (long long)(kp * terms) + q;
  for (int t = 0; t < terms; ++t) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[(long long)t * kp] = h;
    v -= __bfloat162float(h);
  }
}

// First term only of the (centred) operand form, into a buffer with its own leading dimension: the two restarts of a batched
// pair sit side by side in one [n_pad][128] operand (umma_fused.cuh, pair mode).
template <class T>
__global__ void split_factor_hi(const T* __restrict__ G, long long ldg, __nv_bfloat16* __restrict__ out, long long ld_out, long long n_valid,
                                long long n_pad, int k, int kp, const float* __restrict__ centre) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pad * kp) return;
  const long long r = idx / kp;
  const int q = (int)(idx % kp);
  float v = 0.f;
  if (r < n_valid && q < k) {
    v = (float)G[r * ldg + q];
    if (centre != nullptr) v -= centre[q];
  }
  out[r * ld_out + q] = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------------
// Mean-centred single-term operand form (umma_fused1.cuh): centre of a factor, rank-1 parts, first-order correction.
// ------------------------------------------------------------------------------------------------
// part[chunk][q] = sum of G[r][q] over the chunk's rows (fp64); grid = chunks, 256 threads (4 row lanes x 64 columns)
template <class T>
__global__ void __launch_bounds__(256)
col_sum_partial(const T* __restrict__ G, long long ldg, long long n, int k, long long rows_per_chunk, double* __restrict__ part) {
  __shared__ double red[4][64];
  const int q = threadIdx.x & 63, lane = threadIdx.x >> 6;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = r0 + rows_per_chunk < n ? r0 + rows_per_chunk : n;
  double s = 0.0;
  if (q < k)
    for (long long r = r0 + lane; r < r1; r += 4) s += (double)G[r * ldg + q];
  red[lane][q] = s;
  __syncthreads();
  if (lane == 0 && q < k) part[(long long)blockIdx.x * k + q] = red[0][q] + red[1][q] + red[2][q] + red[3][q];
}
// centre[q] = (sum_chunks part[chunk][q]) / n; one block of 1024 threads = 16 chunk lanes x 64 columns, fixed order
__global__ void __launch_bounds__(1024)
finish_centre(const double* __restrict__ part, int chunks, int k, long long n, float* __restrict__ centre) {
  __shared__ double red[16][64];
  const int q = threadIdx.x & 63, lane = threadIdx.x >> 6;
  double s = 0.0;
  if (q < k)
    for (int c = lane; c < chunks; c += 16) s += part[(long long)c * k + q];
  red[lane][q] = s;
  __syncthreads();
  if (lane == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += red[i][q];
    centre[q] = (q < k && n > 0) ? (float)(t / (double)n) : 0.f;
  }
}
// row sums (one warp per row) and partial column sums of a stored relation, fp64 accumulation -> fp32
template <class XT>
__global__ void row_sums(const XT* __restrict__ X, long long ld, long long rows, long long cols, float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (long long c = lane; c < cols; c += 32) s += load_as<XT, double>(X + r * ld + c);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = (float)s;
}
template <class XT>
__global__ void col_sums_partial(const XT* __restrict__ X, long long ld, long long rows, long long cols, long long rows_per_chunk,
                                 double* __restrict__ part) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  double s = 0.0;
  for (long long r = r0; r < r1; ++r) s += load_as<XT, double>(X + r * ld + c);
  part[(long long)blockIdx.y * cols + c] = s;
}
__global__ void finish_col_sums(const double* __restrict__ part, float* __restrict__ out, int chunks, long long cols) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double s = 0.0;
  for (int k = 0; k < chunks; ++k) s += part[(long long)k * cols + c];
  out[c] = (float)s;
}
// out[0] = sum of squares of X (rows x k, leading dimension ldx), out[1] = the same of Y: one block, fixed order.
// (single-term gate: X = slab product with the residual term, Y = the matching rows of A or B)
__global__ void __launch_bounds__(256)
slab_sumsq(const float* __restrict__ X, long long ldx, const float* __restrict__ Y, long long ldy, int rows, int k, double* __restrict__ out,
           const float* __restrict__ y_rowscale = nullptr, const float* __restrict__ y_centre = nullptr) {
  // y_rowscale / y_centre: Y is taken as Y + y_rowscale[r] * y_centre[q] (a rank-1 part that is kept apart from the stored Y)
  __shared__ double rx[256], ry[256];
  double sx = 0.0, sy = 0.0;
  for (int o = threadIdx.x; o < rows * k; o += 256) {
    const int r = o / k, q = o % k;
    const double x = X[(long long)r * ldx + q];
    double y = Y[(long long)r * ldy + q];
    if (y_rowscale != nullptr) y += (double)y_rowscale[r] * (double)y_centre[q];
    sx += x * x;
    sy += y * y;
  }
  rx[threadIdx.x] = sx;
  ry[threadIdx.x] = sy;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) { rx[threadIdx.x] += rx[threadIdx.x + w]; ry[threadIdx.x] += ry[threadIdx.x + w]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = rx[0]; out[1] = ry[0]; }
}
// B[r][q] = colsum[r] * centre[q]  (rows >= n_valid: 0) -- the rank-1 part of R^T G_i, initial value of the reduce target
__global__ void rank1_init(float* __restrict__ B, long long ldb, long long n_rows, long long n_valid, int k,
                           const float* __restrict__ colsum, const float* __restrict__ centre) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * k) return;
  const long long r = idx / k;
  const int q = (int)(idx % k);
  B[r * ldb + q] = (r < n_valid) ? colsum[r] * centre[q] : 0.f;
}
// B[r][q] += colsum[r] * centre[q] on a block of rows (the rank-1 part, added after a reduce-scatter)
__global__ void rank1_add(float* __restrict__ B, long long ldb, long long n_rows, int k, const float* __restrict__ colsum,
                          const float* __restrict__ centre) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * k) return;
  const long long r = idx / k;
  const int q = (int)(idx % k);
  B[r * ldb + q] += colsum[r] * centre[q];
}
// First-order restoration of the residual term in the backbone solve: part[chunk][a][b] = sum_r B[r][a] * lo[r][b],
// lo[r][b] = (G[r][b] - centre[b]) - Gs[r][b]  (exact in fp32).  The correction is 2^-9 of M, so fp32 accumulation per
// chunk (and fp64 across chunks) is ample.  grid = (chunks, ceil(ka/64), ceil(kb/64)), 256 threads, 4x4 per thread.
__global__ void __launch_bounds__(256)
corr_partial(const float* __restrict__ B, long long ldb, const float* __restrict__ G, long long ldg,
             const __nv_bfloat16* __restrict__ Gs, long long ldgs, const float* __restrict__ centre, double* __restrict__ part,
             long long n_rows, int ka, int kb, int rows_per_chunk) {
  // 16-row slabs, double-buffered; the next slab's global loads are in flight while the current one is multiplied
  __shared__ float Xs[2][16][64 + 4];
  __shared__ float Ys[2][16][64 + 4];
  const int tid = threadIdx.x;
  const int a0 = blockIdx.y * 64, b0 = blockIdx.z * 64;
  const long long r_begin = (long long)blockIdx.x * rows_per_chunk;
  const long long r_end = min(n_rows, r_begin + rows_per_chunk);
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float px[4], py[4];
  auto fetch = [&](long long r0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      const int rr = idx / 64, cc = idx % 64;
      float xv = 0.f, yv = 0.f;
      if (r0 + rr < r_end) {
        if (a0 + cc < ka) xv = B[(r0 + rr) * ldb + a0 + cc];
        if (b0 + cc < kb) yv = (G[(r0 + rr) * ldg + b0 + cc] - centre[b0 + cc]) - __bfloat162float(Gs[(r0 + rr) * ldgs + b0 + cc]);
      }
      px[e] = xv;
      py[e] = yv;
    }
  };
  int buf = 0;
  if (r_begin < r_end) fetch(r_begin);
  for (long long r0 = r_begin; r0 < r_end; r0 += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      Xs[buf][idx / 64][idx % 64] = px[e];
      Ys[buf][idx / 64][idx % 64] = py[e];
    }
    __syncthreads();
    if (r0 + 16 < r_end) fetch(r0 + 16);
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Xs[buf][rr][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ys[buf][rr][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    buf ^= 1;
  }
  double* out = part + (long long)blockIdx.x * ka * kb;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + ty * 4 + i;
    if (a >= ka) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = b0 + tx * 4 + j;
      if (b < kb) out[(long long)a * kb + b] = (double)acc[i][j];
    }
  }
}

// dst[r][c] = (DT) src[r][c]   (dtype / layout conversion of uploaded and downloaded matrices)
template <class ST, class DT>
__global__ void convert_2d(const ST* __restrict__ src, long long lds, DT* __restrict__ dst, long long ldd, long long rows,
                           long long cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols, c = idx % cols;
  dst[r * ldd + c] = static_cast<DT>(src[r * lds + c]);
}
template <class ST>
__global__ void convert_2d_to_bf16(const ST* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst, long long ldd,
                                   long long rows, long long cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols, c = idx % cols;
  dst[r * ldd + c] = __float2bfloat16_rn((float)src[r * lds + c]);
}
template <class DT>
__global__ void convert_2d_from_bf16(const __nv_bfloat16* __restrict__ src, long long lds, DT* __restrict__ dst, long long ldd,
                                     long long rows, long long cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols, c = idx % cols;
  dst[r * ldd + c] = (DT)__bfloat162float(src[r * lds + c]);
}

// ------------------------------------------------------------------------------------------------
// fp32 relation -> bf16 planes (storage FZ_BF16X3): X = P0 + P1 + P2 EXACTLY, P_t = bf16 rounding of the running residual
// (three round-to-nearest 8-bit terms cover the 24-bit significand).  The planes feed the same tcgen05 kernels as a
// bf16-stored relation, one pass per non-zero plane, accumulating into the same outputs, so an fp32 relation reaches the
// tensor cores without losing a bit of R.  `part`: 0 the value itself, +1 max(x, 0), -1 max(-x, 0) (the two halves of a
// constraint matrix, _dfmf.py:203-208).  (Own design: no counterpart in the reference.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float plane_part(float v, int part) {
  if (part > 0) return v > 0.f ? v : 0.f;
  if (part < 0) return v < 0.f ? -v : 0.f;
  return v;
}
// need[t] = 1 when plane t of the split holds a non-zero somewhere (integers, ratings, 0/1 data fit plane 0 alone); need[3] = 1
// on a non-finite entry.  One block walks whole rows (coalesced, no index division); the verdict of a block is folded in
// shared memory and written once, so the four flags are not hammered by every thread.  grid = min(rows, a few waves).
__global__ void __launch_bounds__(256)
planes_needed(const float* __restrict__ X, long long ld, long long rows, long long cols, int part, unsigned int* __restrict__ need) {
  unsigned int f = 0u;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* row = X + r * ld;
    for (long long c = threadIdx.x; c < cols; c += blockDim.x) {
      float v = plane_part(row[c], part);
      if (v != v || fabsf(v) > 3.3e38f) { f |= 8u; continue; }      // non-finite entries cannot be split: refused by the caller
      if (v != 0.f) f |= 1u;
      v -= __bfloat162float(__float2bfloat16_rn(v));
      if (v != 0.f) f |= 2u;
      v -= __bfloat162float(__float2bfloat16_rn(v));
      if (v != 0.f) f |= 4u;
    }
  }
  __shared__ unsigned int sf;
  if (threadIdx.x == 0) sf = 0u;
  __syncthreads();
  f = __reduce_or_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && f) atomicOr(&sf, f);
  __syncthreads();
  if (threadIdx.x < 4 && ((sf >> threadIdx.x) & 1u)) need[threadIdx.x] = 1u;
}
// P[t][r][c] = term t (t < n_planes); plane t starts at P + t * plane_stride, rows pitched to ldp (pad columns stay zero).
// grid = min(rows, a few waves) blocks, each walking whole rows.
__global__ void __launch_bounds__(256)
split_planes(const float* __restrict__ X, long long ld, __nv_bfloat16* __restrict__ P, long long ldp, long long plane_stride,
             int n_planes, long long rows, long long cols, int part) {
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* row = X + r * ld;
    __nv_bfloat16* out = P + r * ldp;
    for (long long c = threadIdx.x; c < cols; c += blockDim.x) {
      float v = plane_part(row[c], part);
      for (int t = 0; t < n_planes; ++t) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        out[(long long)t * plane_stride + c] = h;
        v -= __bfloat162float(h);
      }
    }
  }
}
// the same for the masked entries only (dfmc re-imputes them every iteration; the known entries never change)
__global__ void __launch_bounds__(256)
split_planes_masked(const float* __restrict__ X, long long ld, const uint8_t* __restrict__ mask, long long mld,
                    __nv_bfloat16* __restrict__ P, long long ldp, long long plane_stride, int n_planes, long long rows,
                    long long cols) {
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* row = X + r * ld;
    const uint8_t* mrow = mask + r * mld;
    __nv_bfloat16* out = P + r * ldp;
    for (long long c = threadIdx.x; c < cols; c += blockDim.x) {
      if (!mrow[c]) continue;
      float v = row[c];
      for (int t = 0; t < n_planes; ++t) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        out[(long long)t * plane_stride + c] = h;
        v -= __bfloat162float(h);
      }
    }
  }
}
// out[r] = sum over the columns of one sign part of X (fp64 accumulation): the rank-1 part of Theta+- G in the centred form
__global__ void row_sums_part(const float* __restrict__ X, long long ld, long long rows, long long cols, int part,
                              float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (long long c = lane; c < cols; c += 32) s += (double)plane_part(X[r * ld + c], part);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = (float)s;
}

// ------------------------------------------------------------------------------------------------
// dfmc: R[mask] = 0 before the first iteration (_dfmc.py:287-292) and
//       R[mask] = (G_i S G_j^T)[mask] after every S-update (_dfmc.py:319-325);  T1 = G_i S is given.
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void mask_zero(T* __restrict__ R, long long ld, const uint8_t* __restrict__ mask, long long mld, long long rows,
                          long long cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols, c = idx % cols;
  if (mask[r * mld + c]) R[r * ld + c] = T(0);
}

template <class T>
__global__ void __launch_bounds__(256)
impute_masked(T* __restrict__ R, long long ld, const uint8_t* __restrict__ mask, long long mld, const T* __restrict__ T1,
              long long ldt, const T* __restrict__ Gj, long long ldg, long long rows, long long cols, int kj) {
  __shared__ T Ts[32][32 + 1];
  __shared__ T Gsm[32][32 + 1];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;  // ty 0..7, 4 rows each
  const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
  T acc[4] = {T(0), T(0), T(0), T(0)};
  for (int q0 = 0; q0 < kj; q0 += 32) {
    for (int e = 0; e < 4; ++e) {
      const int rr = ty * 4 + e;
      T a = T(0), b = T(0);
      if (r0 + rr < rows && q0 + tx < kj) a = T1[(r0 + rr) * ldt + q0 + tx];
      if (c0 + rr < cols && q0 + tx < kj) b = Gj[(c0 + rr) * ldg + q0 + tx];
      Ts[rr][tx] = a;
      Gsm[rr][tx] = b;
    }
    __syncthreads();
    const int qmax = min(32, kj - q0);
    for (int q = 0; q < qmax; ++q) {
      const T g = Gsm[tx][q];
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] += Ts[ty * 4 + e][q] * g;
    }
    __syncthreads();
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const long long r = r0 + ty * 4 + e, c = c0 + tx;
    if (r < rows && c < cols && mask[r * mld + c]) R[r * ld + c] = acc[e];
  }
}

// ------------------------------------------------------------------------------------------------
// out[0] += sum over the tile of (R - T1 Gj^T)^2   (fp64 accumulation; _dfmf.py:306-319)
// optionally also writes the reconstruction (complete(), base.py:119-146)
// ------------------------------------------------------------------------------------------------
template <class T, class XT>
__global__ void __launch_bounds__(256)
recon_err(const XT* __restrict__ R, long long ld, const T* __restrict__ T1, long long ldt, const T* __restrict__ Gj,
          long long ldg, long long rows, long long cols, int kj, double* __restrict__ out_sq, T* __restrict__ recon,
          long long ldr) {
  __shared__ T Ts[32][32 + 1];
  __shared__ T Gsm[32][32 + 1];
  __shared__ double red[256];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
  T acc[4] = {T(0), T(0), T(0), T(0)};
  for (int q0 = 0; q0 < kj; q0 += 32) {
    for (int e = 0; e < 4; ++e) {
      const int rr = ty * 4 + e;
      T a = T(0), b = T(0);
      if (r0 + rr < rows && q0 + tx < kj) a = T1[(r0 + rr) * ldt + q0 + tx];
      if (c0 + rr < cols && q0 + tx < kj) b = Gj[(c0 + rr) * ldg + q0 + tx];
      Ts[rr][tx] = a;
      Gsm[rr][tx] = b;
    }
    __syncthreads();
    const int qmax = min(32, kj - q0);
    for (int q = 0; q < qmax; ++q) {
      const T g = Gsm[tx][q];
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] += Ts[ty * 4 + e][q] * g;
    }
    __syncthreads();
  }
  double s = 0.0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const long long r = r0 + ty * 4 + e, c = c0 + tx;
    if (r < rows && c < cols) {
      if (recon) recon[r * ldr + c] = acc[e];
      if (R) {
        const double d = (double)load_as<XT, T>(R + r * ld + c) - (double)acc[e];
        s += d * d;
      }
    }
  }
  if (out_sq) {
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
      if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out_sq, red[0]);
  }
}

// num += pos(C), den += neg(C)  (transform: frozen relation terms, _dfmf.py:394-419; no scrub there)
template <class T>
__global__ void accum_sign_split(const T* __restrict__ C, T* __restrict__ num, T* __restrict__ den, long long elems) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= elems) return;
  T p, n;
  sign_split(C[i], p, n);
  num[i] += p;
  den[i] += n;
}

// Counter-based synthetic relation entries: value(r, c) = top 24 bits of
// splitmix64(seed * golden + r * n_cols + c) / 2^24, identical on any sharding and reproducible in
// numpy (oracle/fusion_oracle.py: hashed_uniform).  Written in the relation's storage dtype.
__device__ __forceinline__ float hashed_uniform(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed * 0x9E3779B97F4A7C15ull + idx;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}
template <class OT> __device__ __forceinline__ OT cast_out(float v) { return (OT)v; }
template <> __device__ __forceinline__ __nv_bfloat16 cast_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <class OT>
__global__ void fill_hashed_uniform(OT* __restrict__ dst, long long ld, long long rows, long long cols, long long row0,
                                    unsigned long long seed) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols, c = idx % cols;
  dst[r * ld + c] = cast_out<OT>(hashed_uniform(seed, (unsigned long long)((row0 + r) * cols + c)));
}

}  // namespace fz
