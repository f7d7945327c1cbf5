// Developer probe (not part of the library): validates the tcgen05 skinny products against a CPU
// double-precision reference for every operand-layout variant, then times them on a large matrix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>
#include "../umma_skinny.cuh"
#include "../umma_fused.cuh"
#include "../umma_fused_t.cuh"
#include "../umma_fused1.cuh"
#include "../fz_kernels.cuh"
#include "../tmap.h"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

using namespace fz;

static double g_sustain_s = 0.0;   // > 0: time the kernels in a loop of this many seconds (power-capped regime)
__global__ void fill_random_bf16(__nv_bfloat16* x, size_t n, uint32_t seed, float scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + seed;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    x[i] = __float2bfloat16_rn((h >> 8) * (1.0f / 16777216.0f) * scale);
  }
}
static uint32_t rng_state = 12345u;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return (rng_state >> 8) * (1.0f / 16777216.0f);
}

template <int N, bool T>
static void launch(const CUtensorMap& tx, const CUtensorMap& tg, SkinnyParams p, int ksplit, cudaStream_t st) {
  using Cfg = SkinnyCfg<N>;
  static bool attr = false;
  if (!attr) {
    CK(cudaFuncSetAttribute(umma_skinny_kernel<N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr = true;
  }
  dim3 grid((p.M + kSkBM - 1) / kSkBM, ksplit);
  umma_skinny_kernel<N, T><<<grid, kSkThreads, Cfg::kSmemBytes, st>>>(tx, tg, p);
}

static void dispatch(int N, bool T, const CUtensorMap& tx, const CUtensorMap& tg, SkinnyParams p, int ksplit,
                     cudaStream_t st) {
  if (N == 64 && !T) launch<64, false>(tx, tg, p, ksplit, st);
  else if (N == 64 && T) launch<64, true>(tx, tg, p, ksplit, st);
  else if (N == 128 && !T) launch<128, false>(tx, tg, p, ksplit, st);
  else if (N == 128 && T) launch<128, true>(tx, tg, p, ksplit, st);
  else if (N == 192 && !T) launch<192, false>(tx, tg, p, ksplit, st);
  else if (N == 192 && T) launch<192, true>(tx, tg, p, ksplit, st);
  else { printf("bad N\n"); exit(2); }
}

// returns max relative error (vs max |ref|)
static double run_case(int rows, int cols, int k, int terms, bool trans, int ksplit, bool verbose) {
  const int kp = 64, N = terms * kp;
  const int ld = (cols + 7) / 8 * 8;
  const int M = trans ? cols : rows;       // rows of C
  const int K = trans ? rows : cols;       // reduction
  const int ng = K;                        // rows of Gs
  std::vector<__nv_bfloat16> hX((size_t)rows * ld), hG((size_t)ng * N);
  std::vector<float> fX((size_t)rows * ld), fG((size_t)ng * N);
  for (size_t i = 0; i < hX.size(); ++i) { hX[i] = __float2bfloat16(frand() - 0.3f); fX[i] = __bfloat162float(hX[i]); }
  for (int r = 0; r < ng; ++r)
    for (int c = 0; c < N; ++c) {
      int q = c % kp;
      float v = (q < k) ? (frand() - 0.5f) * ((c / kp) == 0 ? 1.f : 0.004f) : 0.f;
      hG[(size_t)r * N + c] = __float2bfloat16(v);
      fG[(size_t)r * N + c] = __bfloat162float(hG[(size_t)r * N + c]);
    }
  __nv_bfloat16 *dX, *dG; float* dC;
  CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dG, hG.size() * 2));
  CK(cudaMalloc(&dC, (size_t)M * k * 4));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dG, hG.data(), hG.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0, (size_t)M * k * 4));
  CUtensorMap tx, tg; std::string err;
  bool ok = trans ? make_tmap_bf16_2d(&tx, dX, rows, cols, ld, 64, 64, &err)
                  : make_tmap_bf16_2d(&tx, dX, rows, cols, ld, 64, 128, &err);
  ok = ok && make_tmap_bf16_2d(&tg, dG, ng, N, N, 64, 64, &err);
  if (!ok) { printf("tmap error: %s\n", err.c_str()); exit(2); }
  SkinnyParams p;
  p.C = dC; p.ldc = k; p.g_row0 = 0; p.M = M; p.K = K; p.k = k; p.kp = kp;
  p.terms = terms;
  int kps = ((K + ksplit - 1) / ksplit + 63) / 64 * 64;
  int eff_split = (K + kps - 1) / kps;
  p.k_per_split = kps; p.atomic = eff_split > 1;
  dispatch(N, trans, tx, tg, p, eff_split, 0);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hC((size_t)M * k);
  CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
  double maxref = 0, maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int q = 0; q < k; ++q) {
      double s = 0;
      for (int kk = 0; kk < K; ++kk) {
        double x = trans ? fX[(size_t)kk * ld + m] : fX[(size_t)m * ld + kk];
        double g = 0;
        for (int t = 0; t < terms; ++t) g += fG[(size_t)kk * N + t * kp + q];
        s += x * g;
      }
      maxref = fmax(maxref, fabs(s));
      maxerr = fmax(maxerr, fabs(s - hC[(size_t)m * k + q]));
    }
  if (verbose)
    printf("case rows=%d cols=%d k=%d terms=%d trans=%d ksplit=%d : max|ref|=%.4g max err=%.3g rel=%.3g %s\n", rows,
           cols, k, terms, (int)trans, eff_split, maxref, maxerr, maxerr / maxref,
           (maxerr / maxref < 2e-5) ? "OK" : "FAIL");
  cudaFree(dX); cudaFree(dG); cudaFree(dC);
  return maxerr / maxref;
}

static void bench(int n, int terms, bool trans) {
  const int kp = 64, N = terms * kp, k = 64;
  size_t elems = (size_t)n * n;
  __nv_bfloat16 *dX, *dG; float* dC;
  CK(cudaMalloc(&dX, elems * 2)); CK(cudaMalloc(&dG, (size_t)n * N * 2)); CK(cudaMalloc(&dC, (size_t)n * k * 4));
  CK(cudaMemset(dX, 0x3c, elems * 2));  // bf16 0x3c3c ~ 0.0115
  CK(cudaMemset(dG, 0x3c, (size_t)n * N * 2));
  CUtensorMap tx, tg; std::string err;
  bool ok = trans ? make_tmap_bf16_2d(&tx, dX, n, n, n, 64, 64, &err) : make_tmap_bf16_2d(&tx, dX, n, n, n, 64, 128, &err);
  ok = ok && make_tmap_bf16_2d(&tg, dG, n, N, N, 64, 64, &err);
  if (!ok) { printf("tmap error: %s\n", err.c_str()); exit(2); }
  for (int ksplit : {1, 2, 4}) {
    SkinnyParams p;
    p.C = dC; p.ldc = k; p.g_row0 = 0; p.M = n; p.K = n; p.k = k; p.kp = kp;
    p.terms = terms;
    int kps = ((n + ksplit - 1) / ksplit + 63) / 64 * 64;
    p.k_per_split = kps; p.atomic = ksplit > 1;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) dispatch(N, trans, tx, tg, p, ksplit, 0);
    CK(cudaDeviceSynchronize());
    const int reps = 5;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) dispatch(N, trans, tx, tg, p, ksplit, 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    double gb = elems * 2.0 / 1e9;
    printf("bench n=%d terms=%d trans=%d ksplit=%d : %.3f ms  %.1f GB/s (R bytes only)  %.1f TFLOP/s\n", n, terms,
           (int)trans, ksplit, ms, gb / (ms * 1e-3), 2.0 * elems * N / (ms * 1e-3) / 1e12);
  }
  cudaFree(dX); cudaFree(dG); cudaFree(dC);
}

// ---- fused single-pass kernel: A = X Gj, B = X^T Gi from one stream of X
static int run_fused_case(int rows, int cols, int ka, int kb, int csplit, int tma_flush = 1, int b_terms = 2) {
  const int kp = 64, N = 128;
  const int ld = (cols + 7) / 8 * 8;
  std::vector<__nv_bfloat16> hX((size_t)rows * ld), hGj((size_t)cols * N), hGi((size_t)rows * N);
  std::vector<float> fX((size_t)rows * ld), fGj((size_t)cols * N), fGi((size_t)rows * N);
  for (size_t i = 0; i < hX.size(); ++i) { hX[i] = __float2bfloat16(frand() - 0.3f); fX[i] = __bfloat162float(hX[i]); }
  auto fillG = [&](std::vector<__nv_bfloat16>& h, std::vector<float>& f, int n, int k) {
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < N; ++c) {
        int q = c % kp;
        float v = (q < k) ? (frand() - 0.5f) * ((c / kp) == 0 ? 1.f : 0.004f) : 0.f;
        h[(size_t)r * N + c] = __float2bfloat16(v);
        f[(size_t)r * N + c] = __bfloat162float(h[(size_t)r * N + c]);
      }
  };
  fillG(hGj, fGj, cols, ka);
  fillG(hGi, fGi, rows, kb);
  if (b_terms == 1)
    for (int r = 0; r < rows; ++r)
      for (int q = 0; q < kp; ++q) fGi[(size_t)r * N + kp + q] = 0.f;   // the reference ignores the second term too
  __nv_bfloat16 *dX, *dGj, *dGi; float *dA, *dB;
  CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dGj, hGj.size() * 2)); CK(cudaMalloc(&dGi, hGi.size() * 2));
  CK(cudaMalloc(&dA, (size_t)rows * ka * 4)); CK(cudaMalloc(&dB, (size_t)cols * kb * 4));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dGj, hGj.data(), hGj.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dGi, hGi.data(), hGi.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dA, 0, (size_t)rows * ka * 4)); CK(cudaMemset(dB, 0, (size_t)cols * kb * 4));
  CUtensorMap tr, tgj, tgi, tb; std::string err;
  if (kb % 4 != 0 || kb < 32) tma_flush = 0;
  bool ok = make_tmap_bf16_2d(&tr, dX, rows, cols, ld, 64, 128, &err) && make_tmap_bf16_2d(&tgj, dGj, cols, N, N, 64, 64, &err) &&
            make_tmap_bf16_2d(&tgi, dGi, rows, N, N, 64, 128, &err);
  if (ok && tma_flush) ok = make_tmap_f32_2d(&tb, dB, cols, kb, kb, 32, 32, &err);
  if (!tma_flush) tb = tr;
  if (!ok) { printf("tmap error: %s\n", err.c_str()); exit(2); }
  FusedParams p;
  p.A = dA; p.B = dB; p.lda = ka; p.ldb = kb; p.n_rows = rows; p.n_cols = cols; p.k_a = ka; p.k_b = kb; p.gi_row0 = 0; p.probe_skip_flush = 0; p.tma_flush = tma_flush; p.b_terms = b_terms;
  const int tiles = (cols + 127) / 128;
  p.tiles_per_split = (tiles + csplit - 1) / csplit;
  const int splits = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.a_atomic = splits > 1;
  CK(cudaFuncSetAttribute(umma_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemBytes));
  dim3 grid((rows + 255) / 256, splits);
  umma_fused_kernel<<<grid, kFuThreads, kFuSmemBytes>>>(tr, tgj, tgi, tb, tb, p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hA((size_t)rows * ka), hB((size_t)cols * kb);
  CK(cudaMemcpy(hA.data(), dA, hA.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hB.data(), dB, hB.size() * 4, cudaMemcpyDeviceToHost));
  double ma = 0, ea = 0, mb = 0, eb = 0;
  for (int m = 0; m < rows; ++m)
    for (int q = 0; q < ka; ++q) {
      double s = 0;
      for (int c = 0; c < cols; ++c) s += (double)fX[(size_t)m * ld + c] * ((double)fGj[(size_t)c * N + q] + fGj[(size_t)c * N + kp + q]);
      ma = fmax(ma, fabs(s)); ea = fmax(ea, fabs(s - hA[(size_t)m * ka + q]));
    }
  for (int c = 0; c < cols; ++c)
    for (int q = 0; q < kb; ++q) {
      double s = 0;
      for (int r = 0; r < rows; ++r) s += (double)fX[(size_t)r * ld + c] * ((double)fGi[(size_t)r * N + q] + fGi[(size_t)r * N + kp + q]);
      mb = fmax(mb, fabs(s)); eb = fmax(eb, fabs(s - hB[(size_t)c * kb + q]));
    }
  const bool good = ea / ma < 2e-5 && eb / mb < 2e-5;
  printf("fused rows=%d cols=%d ka=%d kb=%d csplit=%d flush=%s b_terms=%d : A rel=%.3g  B rel=%.3g  %s\n", rows, cols, ka, kb, splits,
         tma_flush ? "tma" : "red", b_terms, ea / ma, eb / mb, good ? "OK" : "FAIL");
  cudaFree(dX); cudaFree(dGj); cudaFree(dGi); cudaFree(dA); cudaFree(dB);
  return good ? 0 : 1;
}

static void bench_fused(int n, int skip_flush) {
  const int N = 128, k = 64;
  size_t elems = (size_t)n * n;
  __nv_bfloat16 *dX, *dG; float *dA, *dB;
  CK(cudaMalloc(&dX, elems * 2)); CK(cudaMalloc(&dG, (size_t)n * N * 2));
  CK(cudaMalloc(&dA, (size_t)n * k * 4)); CK(cudaMalloc(&dB, (size_t)n * k * 4));
  CK(cudaMemset(dX, 0x3c, elems * 2)); CK(cudaMemset(dG, 0x3c, (size_t)n * N * 2));
  if (g_sustain_s > 0) {   // random operands: switching activity (power) like real data
    fill_random_bf16<<<1184, 256>>>(dX, elems, 1u, 1.0f);
    fill_random_bf16<<<1184, 256>>>(dG, (size_t)n * N, 2u, 1.0f);
  }
  CUtensorMap tr, tg, tg64, tb; std::string err;
  bool ok = make_tmap_bf16_2d(&tr, dX, n, n, n, 64, 128, &err) && make_tmap_bf16_2d(&tg, dG, n, N, N, 64, 128, &err) &&
            make_tmap_bf16_2d(&tg64, dG, n, N, N, 64, 64, &err) && make_tmap_f32_2d(&tb, dB, n, k, k, 32, 32, &err);
  if (!ok) { printf("tmap error: %s\n", err.c_str()); exit(2); }
  CK(cudaFuncSetAttribute(umma_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemBytes));
  for (int csplit : {1}) {
    FusedParams p;
    p.A = dA; p.B = dB; p.lda = k; p.ldb = k; p.n_rows = n; p.n_cols = n; p.k_a = k; p.k_b = k; p.gi_row0 = 0;
    p.probe_skip_flush = skip_flush & 0x37; p.tma_flush = (skip_flush & 8) ? 0 : 1; p.b_terms = 2;
    const int tiles = (n + 127) / 128;
    p.tiles_per_split = (tiles + csplit - 1) / csplit;
    const int splits = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
    p.a_atomic = splits > 1;
    dim3 grid((n + 255) / 256, splits);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) umma_fused_kernel<<<grid, kFuThreads, kFuSmemBytes>>>(tr, tg64, tg, tb, tb, p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    int reps = 40;
    float ms;
    if (g_sustain_s > 0) {
      // warm into the power-capped regime for half the time, then measure the second half
      reps = (int)(g_sustain_s * 0.5 / 0.6e-3);
      for (int r = 0; r < reps; ++r) umma_fused_kernel<<<grid, kFuThreads, kFuSmemBytes>>>(tr, tg64, tg, tb, tb, p);
    }
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) umma_fused_kernel<<<grid, kFuThreads, kFuSmemBytes>>>(tr, tg64, tg, tb, tb, p);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    printf("bench FUSED%s n=%d csplit=%d grid=%dx%d : %.3f ms  %.1f GB/s (one pass over R)  %.1f TFLOP/s\n", skip_flush == 0 ? " (tma flush)" : skip_flush == 8 ? " (red flush)" : (skip_flush == 1 ? " (probe: no RED)" : (skip_flush == 2 ? " (probe: no B MMA)" : (skip_flush == 4 ? " (probe: no A MMA)" : (skip_flush == 3 ? " (probe: no RED, no B MMA)" : (skip_flush == 6 ? " (probe: no MMA, flush on)" : (skip_flush == 16 ? " (probe: B-product 1 term)" : (skip_flush == 48 ? " (probe: both products 1 term)" : " (probe: TMA only)"))))))), n, splits, grid.x, grid.y,
           ms, elems * 2.0 / 1e9 / (ms * 1e-3), 2.0 * 2.0 * elems * N / (ms * 1e-3) / 1e12);
  }
  cudaFree(dX); cudaFree(dG); cudaFree(dA); cudaFree(dB);
}


// ---- v4 (transposed accumulators, TMEM-resident factor operand): same contract, operands in the GsT form
static inline int gst_row(int term, int q) { return 32 * (q >> 4) + 16 * term + (q & 15); }

static bool make_tmap_generic(CUtensorMap* out, CUtensorMapDataType dt, int esz, const void* base, uint64_t rows, uint64_t cols,
                              uint64_t ld, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle sw, std::string* err) {
  PFN_encodeTiled enc = get_encode_tiled(err);
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esz};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { if (err) *err = "encode failed " + std::to_string((int)r); return false; }
  return true;
}

static int run_fused_t_case(int rows, int cols, int ka, int kb, int csplit, int tma_flush, int variant, int gi_row0 = 0,
                            bool device_split = false) {
  const int ld = (cols + 7) / 8 * 8;
  const long long ldtj = ((cols + 255) / 256) * 256 + 256, ldti = ((gi_row0 + rows + 255) / 256) * 256 + 256;
  std::vector<__nv_bfloat16> hX((size_t)rows * ld), hGjT((size_t)128 * ldtj, __float2bfloat16(0.f)), hGiT((size_t)128 * ldti, __float2bfloat16(0.f));
  std::vector<float> fX((size_t)rows * ld), gj((size_t)cols * 64, 0.f), gi((size_t)rows * 64, 0.f);   // hi + lo as the kernel sees them
  std::vector<float> Gj32((size_t)cols * 64, 0.f), Gi32((size_t)(gi_row0 + rows) * 64, 0.f);
  for (size_t i = 0; i < hX.size(); ++i) { hX[i] = __float2bfloat16(frand() - 0.3f); fX[i] = __bfloat162float(hX[i]); }
  auto fill = [&](std::vector<__nv_bfloat16>& hT, long long ldt, std::vector<float>& sum, std::vector<float>& g32, int n, int k, int off) {
    for (int r = 0; r < n; ++r)
      for (int q = 0; q < k; ++q) {
        const float v = frand() - 0.5f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        hT[(size_t)gst_row(0, q) * ldt + off + r] = hi;
        hT[(size_t)gst_row(1, q) * ldt + off + r] = lo;
        sum[(size_t)r * 64 + q] = __bfloat162float(hi) + __bfloat162float(lo);
        g32[(size_t)(off + r) * 64 + q] = v;
      }
  };
  fill(hGjT, ldtj, gj, Gj32, cols, ka, 0);
  fill(hGiT, ldti, gi, Gi32, rows, kb, gi_row0);
  __nv_bfloat16 *dX, *dGjT, *dGiT; float *dA, *dB;
  CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dGjT, hGjT.size() * 2)); CK(cudaMalloc(&dGiT, hGiT.size() * 2));
  CK(cudaMalloc(&dA, (size_t)rows * ka * 4)); CK(cudaMalloc(&dB, (size_t)cols * kb * 4));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  if (device_split) {
    // operand forms built by split_factor_t from fp32 factors (k columns, ld = 64)
    float *dGj32, *dGi32;
    CK(cudaMalloc(&dGj32, Gj32.size() * 4)); CK(cudaMalloc(&dGi32, Gi32.size() * 4));
    CK(cudaMemcpy(dGj32, Gj32.data(), Gj32.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dGi32, Gi32.data(), Gi32.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dGjT, 0, hGjT.size() * 2)); CK(cudaMemset(dGiT, 0, hGiT.size() * 2));
    split_factor_t<float><<<(cols + 63) / 64, 256>>>(dGj32, 64, dGjT, ldtj, cols, ldtj, ka);
    split_factor_t<float><<<(gi_row0 + rows + 63) / 64, 256>>>(dGi32, 64, dGiT, ldti, gi_row0 + rows, ldti, kb);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<__nv_bfloat16> chk(hGjT.size());
    CK(cudaMemcpy(chk.data(), dGjT, chk.size() * 2, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < chk.size(); ++i) bad += (__bfloat16_as_ushort(chk[i]) != __bfloat16_as_ushort(hGjT[i]));
    std::vector<__nv_bfloat16> chk2(hGiT.size());
    CK(cudaMemcpy(chk2.data(), dGiT, chk2.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < chk2.size(); ++i) {
      // rows below gi_row0 hold other random data on the device side (Gi32 rows < gi_row0 are zero here) -> identical anyway
      bad += (__bfloat16_as_ushort(chk2[i]) != __bfloat16_as_ushort(hGiT[i]));
    }
    printf("split_factor_t vs host operand form: %zu mismatching elements %s\n", bad, bad ? "FAIL" : "OK");
    cudaFree(dGj32); cudaFree(dGi32);
    if (bad) return 1;
  } else {
    CK(cudaMemcpy(dGjT, hGjT.data(), hGjT.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dGiT, hGiT.data(), hGiT.size() * 2, cudaMemcpyHostToDevice));
  }
  CK(cudaMemset(dA, 0, (size_t)rows * ka * 4)); CK(cudaMemset(dB, 0, (size_t)cols * kb * 4));
  CUtensorMap tr, tgj, tb; std::string err;
  if (kb % 4 != 0) tma_flush = 0;
  bool ok = make_tmap_bf16_2d(&tr, dX, rows, cols, ld, 64, 256, &err) && make_tmap_bf16_2d(&tgj, dGjT, 128, ldtj, ldtj, 64, 128, &err);
  if (ok && tma_flush) ok = make_tmap_generic(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dB, cols, kb, kb, tma_flush == 2 ? 64 : 16, 64, CU_TENSOR_MAP_SWIZZLE_NONE, &err);
  if (!tma_flush) tb = tr;
  if (!ok) { printf("fusedT rows=%d cols=%d ka=%d kb=%d: tmap error: %s  FAIL\n", rows, cols, ka, kb, err.c_str()); return 1; }
  FusedTParams p;
  p.A = dA; p.B = dB; p.lda = ka; p.ldb = kb; p.GiT = dGiT; p.ldt = ldti; p.n_rows = rows; p.n_cols = cols; p.k_a = ka; p.k_b = kb;
  p.gi_row0 = gi_row0; p.tma_flush = tma_flush; p.variant = variant;
  const int tiles = (cols + 127) / 128;
  p.tiles_per_split = (tiles + csplit - 1) / csplit;
  const int splits = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.a_atomic = splits > 1;
  CK(cudaFuncSetAttribute(umma_fused_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFtSmemBytes));
  dim3 grid((rows + 255) / 256, splits);
  umma_fused_t_kernel<<<grid, kFtThreads, kFtSmemBytes>>>(tr, tgj, tb, p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hA((size_t)rows * ka), hB((size_t)cols * kb);
  CK(cudaMemcpy(hA.data(), dA, hA.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hB.data(), dB, hB.size() * 4, cudaMemcpyDeviceToHost));
  double ma = 0, ea = 0, mb = 0, eb = 0;
  for (int m = 0; m < rows; ++m)
    for (int q = 0; q < ka; ++q) {
      double s = 0;
      for (int c = 0; c < cols; ++c) s += (double)fX[(size_t)m * ld + c] * (double)gj[(size_t)c * 64 + q];
      ma = fmax(ma, fabs(s)); ea = fmax(ea, fabs(s - hA[(size_t)m * ka + q]));
    }
  for (int c = 0; c < cols; ++c)
    for (int q = 0; q < kb; ++q) {
      double s = 0;
      for (int r = 0; r < rows; ++r) s += (double)fX[(size_t)r * ld + c] * (double)gi[(size_t)r * 64 + q];
      mb = fmax(mb, fabs(s)); eb = fmax(eb, fabs(s - hB[(size_t)c * kb + q]));
    }
  const bool good = ea / ma < 2e-5 && eb / mb < 2e-5;
  printf("fusedT rows=%d cols=%d ka=%d kb=%d csplit=%d flush=%s variant=%d gi_row0=%d : A rel=%.3g  B rel=%.3g  %s\n", rows, cols, ka,
         kb, splits, tma_flush == 2 ? "tma64" : (tma_flush ? "tma16" : "red"), variant, gi_row0, ea / ma, eb / mb, good ? "OK" : "FAIL");
  cudaFree(dX); cudaFree(dGjT); cudaFree(dGiT); cudaFree(dA); cudaFree(dB);
  return good ? 0 : 1;
}

static void bench_fused_t(int n, int variant, int csplit, int flush_mode = 2) {
  const int k = 64;
  size_t elems = (size_t)n * n;
  const long long ldt = ((n + 255) / 256) * 256 + 256;
  __nv_bfloat16 *dX, *dGT; float *dA, *dB;
  CK(cudaMalloc(&dX, elems * 2)); CK(cudaMalloc(&dGT, (size_t)128 * ldt * 2));
  CK(cudaMalloc(&dA, (size_t)n * k * 4)); CK(cudaMalloc(&dB, (size_t)n * k * 4));
  CK(cudaMemset(dX, 0x3c, elems * 2)); CK(cudaMemset(dGT, 0x3c, (size_t)128 * ldt * 2));
  if (g_sustain_s > 0) {
    fill_random_bf16<<<1184, 256>>>(dX, elems, 1u, 1.0f);
    fill_random_bf16<<<1184, 256>>>(dGT, (size_t)128 * ldt, 2u, 1.0f);
  }
  CUtensorMap tr, tg, tb; std::string err;
  bool ok = make_tmap_bf16_2d(&tr, dX, n, n, n, 64, 256, &err) && make_tmap_bf16_2d(&tg, dGT, 128, ldt, ldt, 64, 128, &err) &&
            make_tmap_generic(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dB, n, k, k, flush_mode == 2 ? 64 : 16, 64, CU_TENSOR_MAP_SWIZZLE_NONE, &err);
  if (!ok) { printf("tmap error: %s\n", err.c_str()); exit(2); }
  CK(cudaFuncSetAttribute(umma_fused_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFtSmemBytes));
  FusedTParams p;
  p.A = dA; p.B = dB; p.lda = k; p.ldb = k; p.GiT = dGT; p.ldt = ldt; p.n_rows = n; p.n_cols = n; p.k_a = k; p.k_b = k; p.gi_row0 = 0;
  p.tma_flush = flush_mode; p.variant = variant;
  const int tiles = (n + 127) / 128;
  p.tiles_per_split = (tiles + csplit - 1) / csplit;
  const int splits = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.a_atomic = splits > 1;
  dim3 grid((n + 255) / 256, splits);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) umma_fused_t_kernel<<<grid, kFtThreads, kFtSmemBytes>>>(tr, tg, tb, p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  int reps = 40;
  float ms;
  if (g_sustain_s > 0) {
    reps = (int)(g_sustain_s * 0.5 / 0.6e-3);
    for (int r = 0; r < reps; ++r) umma_fused_t_kernel<<<grid, kFtThreads, kFtSmemBytes>>>(tr, tg, tb, p);
  }
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) umma_fused_t_kernel<<<grid, kFtThreads, kFtSmemBytes>>>(tr, tg, tb, p);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  printf("bench FUSED-T flush=%d variant=%d n=%d grid=%dx%d : %.3f ms  %.1f GB/s (one pass over R)  %.1f TFLOP/s\n", flush_mode, variant, n, grid.x, grid.y,
         ms, elems * 2.0 / 1e9 / (ms * 1e-3), 2.0 * 2.0 * elems * 128 / (ms * 1e-3) / 1e12);
  cudaFree(dX); cudaFree(dGT); cudaFree(dA); cudaFree(dB);
}

static int main_fused_t(int variant, int nbench) {
  int fails = 0;
  fails += run_fused_t_case(256, 64, 64, 64, 1, 2, variant);
  fails += run_fused_t_case(256, 256, 64, 64, 1, 2, variant);
  fails += run_fused_t_case(512, 384, 64, 64, 1, 2, variant);
  fails += run_fused_t_case(1000, 520, 64, 40, 1, 2, variant);
  fails += run_fused_t_case(520, 1000, 50, 64, 3, 2, variant);
  fails += run_fused_t_case(2048, 4096, 64, 64, 4, 2, variant);
  fails += run_fused_t_case(1000, 520, 64, 40, 1, 0, variant);
  fails += run_fused_t_case(3000, 2100, 33, 36, 2, 2, variant);
  fails += run_fused_t_case(777, 3001, 64, 64, 1, 2, variant);
  fails += run_fused_t_case(777, 1001, 64, 64, 2, 2, variant, 1024);     // sharded: factor rows offset, 16-byte aligned
  fails += run_fused_t_case(500, 1001, 64, 48, 1, 2, variant, 12500);    // sharded: offset not a multiple of 8 (scalar preload)
  fails += run_fused_t_case(1000, 520, 64, 40, 1, 2, variant, 0, true);  // operand forms from split_factor_t
  fails += run_fused_t_case(777, 1001, 20, 64, 2, 2, variant, 1024, true);
  fails += run_fused_t_case(130, 77, 7, 12, 1, 2, variant);              // boxes larger than the tensors
  fails += run_fused_t_case(300, 40, 64, 8, 1, 2, variant);
  fails += run_fused_t_case(1000, 520, 64, 40, 1, 1, variant);           // per-warp 64-byte-row flush (mode 1)
  fails += run_fused_t_case(2048, 4096, 64, 64, 4, 1, variant);
  printf("fusedT correctness (variant %d): %d failing cases\n", variant, fails);
  if (nbench > 0) {
    bench_fused_t(nbench, variant, 1);
    bench_fused_t(nbench, variant, 1, 1);     // per-warp flush
    bench_fused_t(nbench, variant | 16, 1);   // staggered sweep
    bench_fused_t(nbench, variant | 8, 1);    // no flush
    bench_fused_t(nbench, variant | 2, 1);    // no B^T-product
    bench_fused_t(nbench, variant | 4, 1);    // no A^T-product
    bench_fused_t(nbench, variant | 14, 1);   // TMA only
  }
  return fails ? 1 : 0;
}


static int g_dyn_chunk = 0;   // units per dynamic chunk of the fused1 schedule (0 = static only)
// ---- v5 (single-term, mean-centred operand form, 4 row blocks per CTA): A = X Gs_j + rowsum c_j^T, B = B0 + X^T Gs_i
static int run_fused1_case(int rows, int cols, int ka, int kb, int csplit, int tma_flush = 1, int gi_row0 = 0, bool rank1 = true) {
  const int N = 128;   // operand rows hold [hi | lo]; the kernel must read the hi half only
  const int ld = (cols + 7) / 8 * 8;
  const int gi_rows = gi_row0 + rows;
  std::vector<__nv_bfloat16> hX((size_t)rows * ld), hGj((size_t)cols * N), hGi((size_t)gi_rows * N);
  std::vector<float> fX((size_t)rows * ld), fGj((size_t)cols * 64, 0.f), fGi((size_t)gi_rows * 64, 0.f);
  for (size_t i = 0; i < hX.size(); ++i) { hX[i] = __float2bfloat16(frand() - 0.3f); fX[i] = __bfloat162float(hX[i]); }
  auto fillG = [&](std::vector<__nv_bfloat16>& h, std::vector<float>& f, int n, int k) {
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < N; ++c) {
        const int q = c % 64;
        float v = (q < k) ? (frand() - 0.5f) : 0.f;
        if (c >= 64) v = 7.f;                                      // poison: the second term must be ignored
        h[(size_t)r * N + c] = __float2bfloat16(v);
        if (c < 64) f[(size_t)r * 64 + q] = __bfloat162float(h[(size_t)r * N + c]);
      }
  };
  fillG(hGj, fGj, cols, ka);
  fillG(hGi, fGi, gi_rows, kb);
  std::vector<float> rowsum(rows, 0.f), colsum(cols, 0.f), cj(64, 0.f), ci(64, 0.f);
  for (int r = 0; r < rows; ++r) { double s = 0; for (int c = 0; c < cols; ++c) s += fX[(size_t)r * ld + c]; rowsum[r] = (float)s; }
  for (int c = 0; c < cols; ++c) { double s = 0; for (int r = 0; r < rows; ++r) s += fX[(size_t)r * ld + c]; colsum[c] = (float)s; }
  for (int q = 0; q < 64; ++q) { cj[q] = q < ka ? 0.25f + 0.01f * q : 0.f; ci[q] = q < kb ? 0.5f - 0.005f * q : 0.f; }
  __nv_bfloat16 *dX, *dGj, *dGi; float *dA, *dB, *dRs, *dCs, *dCj, *dCi;
  CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dGj, hGj.size() * 2)); CK(cudaMalloc(&dGi, hGi.size() * 2));
  CK(cudaMalloc(&dA, (size_t)rows * ka * 4)); CK(cudaMalloc(&dB, (size_t)cols * kb * 4));
  CK(cudaMalloc(&dRs, rows * 4)); CK(cudaMalloc(&dCs, cols * 4)); CK(cudaMalloc(&dCj, 256)); CK(cudaMalloc(&dCi, 256));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dGj, hGj.data(), hGj.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dGi, hGi.data(), hGi.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dRs, rowsum.data(), rows * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dCs, colsum.data(), cols * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dCj, cj.data(), 256, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dCi, ci.data(), 256, cudaMemcpyHostToDevice));
  CK(cudaMemset(dA, 0, (size_t)rows * ka * 4));
  if (rank1) rank1_init<<<(unsigned)(((size_t)cols * kb + 255) / 256), 256>>>(dB, kb, cols, cols, kb, dCs, dCi);
  else CK(cudaMemset(dB, 0, (size_t)cols * kb * 4));
  CUtensorMap tr, tgj, tgi, tb, ta; std::string err;
  int flush_b = tma_flush, flush_a = tma_flush;
  if (kb % 4 != 0 || kb < 32) flush_b = 0;
  if (ka % 4 != 0 || ka < 32 || rows < 32) flush_a = 0;
  bool ok = make_tmap_bf16_2d(&tr, dX, rows, cols, ld, 64, 128, &err) && make_tmap_bf16_2d(&tgj, dGj, cols, N, N, 64, 128, &err) &&
            make_tmap_bf16_2d(&tgi, dGi, gi_rows, N, N, 64, 128, &err);
  if (ok && flush_b) ok = make_tmap_f32_2d(&tb, dB, cols, kb, kb, 32, 32, &err);
  if (ok && flush_a) ok = make_tmap_f32_2d(&ta, dA, rows, ka, ka, 32, 32, &err);
  if (!flush_b) tb = tr;
  if (!flush_a) ta = tr;
  if (!ok) { printf("tmap error: %s\n", err.c_str()); exit(2); }
  Fused1Params p;
  p.A = dA; p.B = dB; p.lda = ka; p.ldb = kb; p.rowsum = rank1 ? dRs : nullptr; p.cj = dCj; p.n_rows = rows; p.n_cols = cols;
  p.k_a = ka; p.k_b = kb; p.gi_row0 = gi_row0; p.probe = 0; p.tma_flush = (flush_b ? 1 : 0) | (flush_a ? 2 : 0);
  int* dCtr; CK(cudaMalloc(&dCtr, 4)); CK(cudaMemset(dCtr, 0, 4));
  p.work_counter = g_dyn_chunk > 0 ? dCtr : nullptr; p.dyn_chunk = g_dyn_chunk;
  CK(cudaFuncSetAttribute(umma_fused1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kF1SmemBytes));
  const int splits = csplit;       // number of CTAs of the persistent grid
  umma_fused1_kernel<<<csplit, kF1Threads, kF1SmemBytes>>>(tr, tgj, tgi, tb, ta, p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hA((size_t)rows * ka), hB((size_t)cols * kb);
  CK(cudaMemcpy(hA.data(), dA, hA.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hB.data(), dB, hB.size() * 4, cudaMemcpyDeviceToHost));
  double ma = 0, ea = 0, mb = 0, eb = 0;
  for (int m = 0; m < rows; ++m)
    for (int q = 0; q < ka; ++q) {
      double s = rank1 ? (double)rowsum[m] * cj[q] : 0.0;
      for (int c = 0; c < cols; ++c) s += (double)fX[(size_t)m * ld + c] * (double)fGj[(size_t)c * 64 + q];
      ma = fmax(ma, fabs(s)); ea = fmax(ea, fabs(s - hA[(size_t)m * ka + q]));
    }
  for (int c = 0; c < cols; ++c)
    for (int q = 0; q < kb; ++q) {
      double s = rank1 ? (double)colsum[c] * ci[q] : 0.0;
      for (int r = 0; r < rows; ++r) s += (double)fX[(size_t)r * ld + c] * (double)fGi[(size_t)(gi_row0 + r) * 64 + q];
      mb = fmax(mb, fabs(s)); eb = fmax(eb, fabs(s - hB[(size_t)c * kb + q]));
    }
  const bool good = ea / ma < 2e-5 && eb / mb < 2e-5;
  printf("fused1 rows=%d cols=%d ka=%d kb=%d ctas=%d dyn_chunk=%d flush=%d gi_row0=%d rank1=%d : A rel=%.3g  B rel=%.3g  %s\n", rows, cols, ka, kb,
         splits, g_dyn_chunk, p.tma_flush, gi_row0, (int)rank1, ea / ma, eb / mb, good ? "OK" : "FAIL");
  cudaFree(dCtr);
  cudaFree(dX); cudaFree(dGj); cudaFree(dGi); cudaFree(dA); cudaFree(dB); cudaFree(dRs); cudaFree(dCs); cudaFree(dCj); cudaFree(dCi);
  return good ? 0 : 1;
}

static void bench_fused1(int n, int probe, int csplit = 1) {
  const int N = 128, k = 64;
  size_t elems = (size_t)n * n;
  __nv_bfloat16 *dX, *dG; float *dA, *dB, *dRs, *dC;
  CK(cudaMalloc(&dX, elems * 2)); CK(cudaMalloc(&dG, (size_t)n * N * 2));
  CK(cudaMalloc(&dA, (size_t)n * k * 4)); CK(cudaMalloc(&dB, (size_t)n * k * 4)); CK(cudaMalloc(&dRs, (size_t)n * 4)); CK(cudaMalloc(&dC, 256));
  CK(cudaMemset(dX, 0x3c, elems * 2)); CK(cudaMemset(dG, 0x3c, (size_t)n * N * 2)); CK(cudaMemset(dRs, 0, (size_t)n * 4)); CK(cudaMemset(dC, 0, 256));
  if (g_sustain_s > 0) {   // random operands: switching activity (power) like real data; centred operands have both signs
    fill_random_bf16<<<1184, 256>>>(dX, elems, 1u, 1.0f);
    fill_random_bf16<<<1184, 256>>>(dG, (size_t)n * N, 2u, 1.0f);
  }
  CUtensorMap tr, tg, tb, ta; std::string err;
  bool ok = make_tmap_bf16_2d(&tr, dX, n, n, n, 64, 128, &err) && make_tmap_bf16_2d(&tg, dG, n, N, N, 64, 128, &err) &&
            make_tmap_f32_2d(&tb, dB, n, k, k, 32, 32, &err) && make_tmap_f32_2d(&ta, dA, n, k, k, 32, 32, &err);
  if (!ok) { printf("tmap error: %s\n", err.c_str()); exit(2); }
  CK(cudaFuncSetAttribute(umma_fused1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kF1SmemBytes));
  Fused1Params p;
  p.A = dA; p.B = dB; p.lda = k; p.ldb = k; p.rowsum = dRs; p.cj = dC; p.n_rows = n; p.n_cols = n; p.k_a = k; p.k_b = k; p.gi_row0 = 0;
  p.probe = probe; p.tma_flush = 3;
  int* dCtr; CK(cudaMalloc(&dCtr, 4)); CK(cudaMemset(dCtr, 0, 4));
  p.work_counter = g_dyn_chunk > 0 ? dCtr : nullptr; p.dyn_chunk = g_dyn_chunk;
  dim3 grid(csplit > 0 ? csplit : 148);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) { if (p.work_counter) cudaMemsetAsync(p.work_counter, 0, 4); umma_fused1_kernel<<<grid, kF1Threads, kF1SmemBytes>>>(tr, tg, tg, tb, ta, p); }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  int reps = 40;
  float ms;
  if (g_sustain_s > 0) {
    reps = (int)(g_sustain_s * 0.5 / 0.6e-3);
    for (int r = 0; r < reps; ++r) { if (p.work_counter) cudaMemsetAsync(p.work_counter, 0, 4); umma_fused1_kernel<<<grid, kF1Threads, kF1SmemBytes>>>(tr, tg, tg, tb, ta, p); }
  }
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) { if (p.work_counter) cudaMemsetAsync(p.work_counter, 0, 4); umma_fused1_kernel<<<grid, kF1Threads, kF1SmemBytes>>>(tr, tg, tg, tb, ta, p); }
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  printf("bench FUSED1 (persistent) probe=%d n=%d grid=%dx%d : %.3f ms  %.1f GB/s (one pass over R)  %.1f TFLOP/s executed\n", probe, n, grid.x, grid.y,
         ms, elems * 2.0 / 1e9 / (ms * 1e-3), 2.0 * 2.0 * elems * 64 / (ms * 1e-3) / 1e12);
  cudaFree(dX); cudaFree(dG); cudaFree(dA); cudaFree(dB); cudaFree(dRs); cudaFree(dC);
}

static int main_fused1(int nbench, double sustain, int csplit) {
  int fails = 0;
  fails += run_fused1_case(128, 128, 64, 64, 1);
  fails += run_fused1_case(512, 384, 64, 64, 1);
  fails += run_fused1_case(640, 256, 64, 64, 1);                 // one CTA, two segments (second row group: one row block)
  fails += run_fused1_case(640, 256, 64, 64, 3);                 // 4 units over 3 CTAs
  fails += run_fused1_case(1000, 520, 64, 40, 2);
  fails += run_fused1_case(520, 1000, 50, 64, 5);
  fails += run_fused1_case(130, 77, 7, 12, 1);                    // red.global fallback, boxes larger than the tensors
  fails += run_fused1_case(2048, 4096, 64, 64, 7);                // row groups split between CTAs at odd tiles
  fails += run_fused1_case(1000, 520, 64, 40, 1, 0);
  fails += run_fused1_case(3000, 2100, 33, 36, 148);              // more CTAs than units: idle CTAs
  fails += run_fused1_case(777, 3001, 64, 64, 13);
  fails += run_fused1_case(777, 1001, 64, 64, 2, 1, 1024);        // sharded: factor rows offset
  fails += run_fused1_case(500, 1001, 64, 48, 1, 1, 12500);
  fails += run_fused1_case(1000, 520, 64, 40, 4, 1, 0, false);    // no rank-1 part
  fails += run_fused1_case(4096, 8192, 64, 64, 148);
  fails += run_fused1_case(5000, 3000, 64, 64, 37);
  for (int dc : {1, 3, 16}) {       // the same shapes with a dynamic tail
    g_dyn_chunk = dc;
    fails += run_fused1_case(640, 256, 64, 64, 3);
    fails += run_fused1_case(2048, 4096, 64, 64, 7);
    fails += run_fused1_case(777, 3001, 64, 64, 13);
    fails += run_fused1_case(4096, 8192, 64, 64, 148);
    fails += run_fused1_case(5000, 3000, 64, 64, 37);
    fails += run_fused1_case(3000, 2100, 33, 36, 148);
  }
  g_dyn_chunk = 0;
  printf("fused1 correctness: %d failing cases\n", fails);
  if (nbench > 0) {
    g_sustain_s = 0;
    for (int probe : {0, 1, 2, 4, 7}) bench_fused1(nbench, probe, csplit);   // full, no flush, no B MMA, no A MMA, TMA only
    for (int dc : {4, 16}) { g_dyn_chunk = dc; printf("dyn_chunk=%d: ", dc); bench_fused1(nbench, 0, csplit); }
    g_dyn_chunk = 0;
    if (sustain > 0) {
      g_sustain_s = sustain;
      for (int probe : {0, 1, 2, 4, 7}) bench_fused1(nbench, probe, csplit);
      g_dyn_chunk = 16; printf("dyn_chunk=16: "); bench_fused1(nbench, 0, csplit); g_dyn_chunk = 0;
      bench_fused(nbench, 0);                                        // v3 for reference on the same box
    }
  }
  return fails ? 1 : 0;
}

int main(int argc, char** argv) {
  setvbuf(stdout, nullptr, _IONBF, 0);
  if (argc > 1 && argv[1][0] == 'b') {   // timing only of the single-term kernel as compiled (geometry study):  b <n> <seconds> <ctas>
    const int n = argc > 2 ? atoi(argv[2]) : 36864;
    const double sustain = argc > 3 ? atof(argv[3]) : 2.0;
    const int ctas = argc > 4 ? atoi(argv[4]) : 148;
    printf("geometry: %d row blocks per group, %d relation stages, %d B of flush staging, %d B of shared memory\n", kF1Blocks, kF1RStages,
           kF1StageBytes, kF1SmemBytes);
    g_sustain_s = 0;
    for (int probe : {0, 1, 7}) bench_fused1(n, probe, ctas);
    g_sustain_s = sustain;
    for (int probe : {0, 1}) bench_fused1(n, probe, ctas);
    return 0;
  }
  if (argc > 1 && argv[1][0] == '1') return main_fused1(argc > 2 ? atoi(argv[2]) : 0, argc > 3 ? atof(argv[3]) : 0.0, argc > 4 ? atoi(argv[4]) : 148);
  if (argc > 1 && argv[1][0] == 'm') {   // single-term B-product (the fp16 variant of this probe hit an illegal instruction:
    int fails = 0;                       //  bf16 A x fp16 B is not a legal kind::f16 combination, profiles/r01b_mixed_format_probe.log)
    fails += run_fused_case(512, 384, 64, 64, 1, 1, 1);
    fails += run_fused_case(1000, 520, 64, 40, 1, 1, 1);
    fails += run_fused_case(2048, 4096, 64, 64, 4, 1, 1);
    printf("single-term B-product probe: %d failing cases\n", fails);
    return fails ? 1 : 0;
  }
  if (argc > 1 && argv[1][0] == 's') {   // sustained (power-capped) component study:  s <n> <seconds>
    const int n = argc > 2 ? atoi(argv[2]) : 37888;
    g_sustain_s = argc > 3 ? atof(argv[3]) : 4.0;
    printf("sustained mode: %.1f s per configuration (second half timed)\n", g_sustain_s);
    if (argc > 4) for (int mode : {0, 2, 4, 6, 7}) bench_fused(n, mode);   // v3: full, no B MMA, no A MMA, no MMA, TMA only
    for (int mode : {0, 16, 48, 1}) bench_fused(n, mode);                 // v3: full, B 1 term, A and B 1 term, no flush
    for (int v : {0, 16, 8}) bench_fused_t(n, v, 1, 2);                   // v4, one 16 KB reduce per chunk: plain, staggered sweep, no flush
    bench_fused_t(n, 0, 1, 1);                                           // v4, per-warp 64-byte-row reduces
    return 0;
  }
  if (argc > 1 && argv[1][0] == 't') return main_fused_t(argc > 2 ? atoi(argv[2]) : 0, argc > 3 ? atoi(argv[3]) : 0);
  int nbench = argc > 1 ? atoi(argv[1]) : 32768;
  int fails = 0;
  for (int trans = 0; trans < 2; ++trans)
    for (int terms = 1; terms <= 3; ++terms) {
      fails += run_case(256, 256, 64, terms, trans, 1, true) > 2e-5;
      fails += run_case(1000, 520, 64, terms, trans, 1, true) > 2e-5;
      fails += run_case(520, 1000, 50, terms, trans, 3, true) > 2e-5;
      fails += run_case(130, 77, 7, terms, trans, 1, true) > 2e-5;
    }
  fails += run_case(2048, 4096, 64, 2, false, 4, true) > 2e-5;
  fails += run_case(4096, 2048, 64, 2, true, 4, true) > 2e-5;
  fails += run_fused_case(256, 256, 64, 64, 1);
  fails += run_fused_case(512, 384, 64, 64, 1);
  fails += run_fused_case(1000, 520, 64, 40, 1);
  fails += run_fused_case(520, 1000, 50, 64, 3);
  fails += run_fused_case(130, 77, 7, 12, 1);
  fails += run_fused_case(2048, 4096, 64, 64, 4);
  fails += run_fused_case(1000, 520, 64, 40, 1, 0);
  fails += run_fused_case(2048, 4096, 64, 64, 4, 0);
  fails += run_fused_case(3000, 2100, 33, 36, 2);
  fails += run_fused_case(777, 3001, 64, 64, 1);
  printf("correctness: %d failing cases\n", fails);
  if (nbench > 0) {
    for (int mode : {0, 8, 1, 7}) bench_fused(nbench, mode);           // fused kernel: flush variants / component probes
    if (argc > 2) {                                                     // any second argument: also the two-pass kernels
      for (int trans = 0; trans < 2; ++trans)
        for (int terms = 1; terms <= 2; ++terms) bench(nbench, terms, trans);
    }
  }
  return fails ? 1 : 0;
}
