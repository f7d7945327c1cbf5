// B200-native collective matrix tri-factorization engine behind include/fz_fusion.h.
//
// One handle = one call of the reference's dfmf() / dfmc() / transform()
// (skfusion/fusion/decomposition/_dfmf.py:127, _dfmc.py:181, _dfmf.py:330).  The per-iteration work is
// regrouped around two streamed products per relation,
//        A_ij = R_ij G_j          B_ij = R_ij^T G_i ,
// from which everything the reference computes follows with k x k algebra:
//        G_i^T R_ij G_j = G_i^T A_ij                     (S-update,  _dfmf.py:236-239)
//        tmp1 = R_ij (G_j S^T) = A_ij S^T                 (_dfmf.py:254)
//        tmp4 = R_ij^T (G_i S) = B_ij S                   (_dfmf.py:266)
// bf16-stored relations take the tcgen05/TMA kernels (umma_fused*.cuh, umma_skinny.cuh), and so do fp32 relations kept as
// exact bf16 planes (FZ_BF16X3: constraint matrices, masked relations and ranks above 64 included); plain fp32 / fp64-stored
// ones the exact CUDA-core kernel.  The k x k chain is always fp64 (fz_chain.cuh).  There is no CPU fallback anywhere.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

#include "../../include/fz_fusion.h"
#include "fz_chain.cuh"
#include "fz_kernels.cuh"
#include "nccl_shim.h"
#include "peer_reduce.cuh"
#include "tmap.h"
#include "umma_fused.cuh"
#include "umma_fused1.cuh"
#include "umma_fused_t.cuh"
#include "umma_outer.cuh"
#include "umma_skinny.cuh"

namespace fz {

struct FzError {
  int status;
  std::string msg;
};
#define FZ_THROW(st, ...)                              \
  do {                                                 \
    char buf_[512];                                    \
    snprintf(buf_, sizeof(buf_), __VA_ARGS__);         \
    throw FzError{(st), std::string(buf_)};            \
  } while (0)
#define NCCL_OK(x)                                                                                       \
  do {                                                                                                   \
    ncclResult_t r_ = (x);                                                                               \
    if (r_ != ncclSuccess) FZ_THROW(FZ_ERR_CUDA, "%s failed: %s (%s:%d)", #x, nccl_api().GetErrorString(r_), __FILE__, __LINE__); \
  } while (0)
#define CUDA_OK(x)                                                                                 \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) FZ_THROW(FZ_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  void alloc(size_t n, bool zero = true) {
    if (n == 0) n = 16;
    if (p && bytes == n) {   // same size again (e.g. a second fz_transform_prepare): keep the allocation
      if (zero) zero_now(n);
      return;
    }
    release();
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {
      p = nullptr;
      FZ_THROW(FZ_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", n, cudaGetErrorString(e));
    }
    bytes = n;
    if (zero) zero_now(n);
  }
  // Zeroing runs on the legacy default stream, which non-blocking caller streams do not wait for: finish it here, so that
  // whatever stream the caller enqueues on next finds the buffer cleared (allocation is a set-up step, never in the loop).
  void zero_now(size_t n) {
    CUDA_OK(cudaMemset(p, 0, n));
    CUDA_OK(cudaStreamSynchronize(nullptr));
  }
  template <class U> U* as() const { return reinterpret_cast<U*>(p); }
};

static inline size_t dtype_size(int d) {
  switch (d) {
    case FZ_F64: return 8;
    case FZ_F32: return 4;
    case FZ_BF16: return 2;
    case FZ_U8: return 1;
  }
  return 0;
}
static inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

class EngineBase {
 public:
  virtual ~EngineBase() {
    for (auto& pr : prof_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  }
  std::string err;
  int64_t launches = 0;
  // optional per-launch timing of the streamed tensor-core products (bench.py's roofline leg)
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
  double prof_bytes = 0.0;     // relation bytes streamed by the timed launches (rows x cols x 2 each)
  double prof_alg_bytes = 0.0; // algorithmic share: a launch that yields only one of the two products of a
                               // relation is credited with half of the bytes it streams (one pass feeds both)
  void prof_begin(cudaStream_t st) {
    if (!profile) return;
    if (prof_used == prof_events.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      prof_events.push_back({a, b});
    }
    cudaEventRecord(prof_events[prof_used].first, st);
  }
  void prof_end(cudaStream_t st, double bytes, int products_per_pass = 2) {
    if (!profile) return;
    cudaEventRecord(prof_events[prof_used].second, st);
    ++prof_used;
    prof_bytes += bytes;
    prof_alg_bytes += bytes / products_per_pass;   // 2: single-product launch (half credit); 1: fused launch
  }
  int device = 0;              // every C-ABI entry makes this the calling thread's current device (DeviceGuard)
  virtual int compute_dtype() const = 0;
  virtual void set_shard(int world, int rank) = 0;
  virtual void comm_init(const void* unique_id) = 0;
  virtual int add_type(int64_t n, int k) = 0;
  virtual int add_relation(int ti, int tj, const void* data, int64_t ld, int src, int mem, int storage, int borrow,
                           const uint8_t* mask, int64_t mask_ld, int mask_mem) = 0;
  virtual void set_factor(int t, const void* G0, int64_t ld, int src, int mem) = 0;
  virtual void set_backbone(int rel, const void* S, int64_t ld, int src, int mem) = 0;
  virtual void set_split_terms(int terms) = 0;
  virtual void finalize() = 0;
  virtual void iterate(int algo, int n_iters, cudaStream_t st) = 0;
  virtual void phase_products(int algo, cudaStream_t st) = 0;
  virtual void phase_update(int algo, cudaStream_t st) = 0;
  virtual void phase_products_begin(int algo, cudaStream_t st) = 0;
  virtual void phase_product_relation(int algo, int rel, cudaStream_t st) = 0;
  virtual void phase_products_end(int algo, cudaStream_t st) = 0;
  virtual void comm_small(void** ptr, int64_t* count) = 0;
  virtual void comm_bpartial(int rel, void** full, void** local, int64_t* local_count, int* dtype) = 0;
  virtual void comm_factor(int t, void** full, int64_t* local_count, int* dtype) = 0;
  virtual void transform_prepare(int target, cudaStream_t st) = 0;
  virtual void transform_iterate(int n_iters, cudaStream_t st) = 0;
  virtual void get_factor(int t, void* dst, int64_t ld, int dd, int mem, cudaStream_t st) = 0;
  virtual void get_backbone(int rel, void* dst, int64_t ld, int dd, int mem, cudaStream_t st) = 0;
  virtual void objective(double* per_rel, double* total, cudaStream_t st) = 0;
  virtual void complete(int rel, void* dst, int64_t ld, int dd, int mem, cudaStream_t st) = 0;
  virtual void profile_product(int ti, int tj, const void* S, int64_t lds, int sd, int smem, void* dst, int64_t ld, int dd, int mem,
                               cudaStream_t st) = 0;
  virtual void init_fill(int t, double value, cudaStream_t st) = 0;
  virtual void relation_norms(int rel, int axis, double* dst_host, cudaStream_t st) = 0;
  virtual void init_add_sampled_means(int t, int rel, const int32_t* idx_host, int p_c, cudaStream_t st) = 0;
  virtual void init_end() = 0;
  virtual void pair_iterate(EngineBase* other, int n_iters, cudaStream_t st) = 0;
  virtual void relation_device_ptr(int rel, void** ptr, int64_t* ld, int* dtype) = 0;
  virtual void operand_stats(int64_t* single_iters, int64_t* two_term_iters, int64_t* paired_iters, double* err_estimate, double* cond_estimate) = 0;
};

// ------------------------------------------------------------------------------------------------
// dtype-generic copy-in / copy-out helpers
// ------------------------------------------------------------------------------------------------
template <class ST, class DT>
static void launch_convert(const void* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int64_t cols, cudaStream_t st) {
  if (rows * cols == 0) return;
  convert_2d<ST, DT><<<nblk(rows * cols, 256), 256, 0, st>>>((const ST*)src, lds, (DT*)dst, ldd, rows, cols);
}
// device src (dtype sd) -> device dst (dtype dd)
static void device_convert(const void* src, int64_t lds, int sd, void* dst, int64_t ldd, int dd, int64_t rows, int64_t cols,
                           cudaStream_t st) {
  if (rows * cols == 0) return;
  const unsigned g = nblk(rows * cols, 256);
  if (sd == FZ_BF16 && dd == FZ_BF16) {
    CUDA_OK(cudaMemcpy2DAsync(dst, ldd * 2, src, lds * 2, cols * 2, rows, cudaMemcpyDeviceToDevice, st));
  } else if (dd == FZ_BF16) {
    if (sd == FZ_F64) convert_2d_to_bf16<double><<<g, 256, 0, st>>>((const double*)src, lds, (__nv_bfloat16*)dst, ldd, rows, cols);
    else if (sd == FZ_F32) convert_2d_to_bf16<float><<<g, 256, 0, st>>>((const float*)src, lds, (__nv_bfloat16*)dst, ldd, rows, cols);
    else FZ_THROW(FZ_ERR_INVALID, "unsupported conversion %d -> bf16", sd);
  } else if (sd == FZ_BF16) {
    if (dd == FZ_F64) convert_2d_from_bf16<double><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, lds, (double*)dst, ldd, rows, cols);
    else if (dd == FZ_F32) convert_2d_from_bf16<float><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, lds, (float*)dst, ldd, rows, cols);
    else FZ_THROW(FZ_ERR_INVALID, "unsupported conversion bf16 -> %d", dd);
  } else if (sd == FZ_F64 && dd == FZ_F64) launch_convert<double, double>(src, lds, dst, ldd, rows, cols, st);
  else if (sd == FZ_F64 && dd == FZ_F32) launch_convert<double, float>(src, lds, dst, ldd, rows, cols, st);
  else if (sd == FZ_F32 && dd == FZ_F64) launch_convert<float, double>(src, lds, dst, ldd, rows, cols, st);
  else if (sd == FZ_F32 && dd == FZ_F32) launch_convert<float, float>(src, lds, dst, ldd, rows, cols, st);
  else if (sd == FZ_U8 && dd == FZ_U8) launch_convert<uint8_t, uint8_t>(src, lds, dst, ldd, rows, cols, st);
  else FZ_THROW(FZ_ERR_INVALID, "unsupported conversion %d -> %d", sd, dd);
  CUDA_OK(cudaGetLastError());
}
// caller buffer (host or device) -> engine device buffer, converting dtype
static void copy_in(const void* src, int64_t lds, int sd, int mem, void* dst, int64_t ldd, int dd, int64_t rows, int64_t cols,
                    cudaStream_t st) {
  if (rows * cols == 0) return;
  if (mem == FZ_DEVICE) {
    device_convert(src, lds, sd, dst, ldd, dd, rows, cols, st);
    return;
  }
  const size_t es = dtype_size(sd);
  if (sd == dd) {
    if (lds == cols && ldd == cols)   // contiguous on both sides: one linear DMA (the 2-D path is slower over PCIe)
      CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)rows * cols * es, cudaMemcpyHostToDevice, st));
    else
      CUDA_OK(cudaMemcpy2DAsync(dst, ldd * es, src, lds * es, cols * es, rows, cudaMemcpyHostToDevice, st));
    return;
  }
  DevBuf stage;
  stage.alloc((size_t)rows * cols * es, false);
  CUDA_OK(cudaMemcpy2DAsync(stage.p, cols * es, src, lds * es, cols * es, rows, cudaMemcpyHostToDevice, st));
  device_convert(stage.p, cols, sd, dst, ldd, dd, rows, cols, st);
  CUDA_OK(cudaStreamSynchronize(st));  // staging buffer dies here
}
// engine device buffer -> caller buffer (host or device)
static void copy_out(const void* src, int64_t lds, int sd, void* dst, int64_t ldd, int dd, int mem, int64_t rows, int64_t cols,
                     cudaStream_t st) {
  if (rows * cols == 0) return;
  if (mem == FZ_DEVICE) {
    device_convert(src, lds, sd, dst, ldd, dd, rows, cols, st);
    return;
  }
  const size_t es = dtype_size(dd);
  if (sd == dd) {
    CUDA_OK(cudaMemcpy2DAsync(dst, ldd * es, src, lds * es, cols * es, rows, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return;
  }
  DevBuf stage;
  stage.alloc((size_t)rows * cols * es, false);
  device_convert(src, lds, sd, stage.p, cols, dd, rows, cols, st);
  CUDA_OK(cudaMemcpy2DAsync(dst, ldd * es, stage.p, cols * es, cols * es, rows, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
}

template <class T> struct DtypeOf;
template <> struct DtypeOf<float> { static constexpr int value = FZ_F32; };
template <> struct DtypeOf<double> { static constexpr int value = FZ_F64; };

// ------------------------------------------------------------------------------------------------
template <class T>
class Engine : public EngineBase {
  static constexpr int kDT = DtypeOf<T>::value;
  static constexpr int kKp = 64;  // padded width of one bf16 split term on the tensor-core path

  struct TypeRec {
    int64_t n = 0, n_pad = 0, m_loc = 0, row0 = 0, rows_loc = 0;
    int k = 0;
    int kp = 64;             // rank padded to whole 64-column blocks: width of one term of the operand form
    DevBuf G[2];
    int cur = 0;
    bool has_factor = false;
    bool need_gs = false;
    bool gs_centred = false; // Gs currently holds the terms of G - 1 c^T (centred form) rather than of G
    DevBuf Gs;               // bf16 [n_pad][terms*kp]
    CUtensorMap tmG;         // box {64 cols, 64 rows}   (two-pass kernels)
    CUtensorMap tmG128;      // box {64 cols, 128 rows}  (fused kernel)
    DevBuf GsT;              // bf16 [128][ldt]: transposed operand form (umma_fused_t.cuh), fused kernel v4
    int64_t ldt = 0;
    CUtensorMap tmGT;        // box {64 cols, 128 rows}
    DevBuf pairGs;                // batched pair of restarts: [n_pad][128] bf16 = [run 0 term | run 1 term] (held by run 0's handle)
    CUtensorMap tmPair64, tmPair128;
    const __nv_bfloat16* hi_ptr = nullptr;   // where the first operand term of this factor currently lives, and its leading
    long long hi_ld = 0;                     // dimension (own Gs, or a half of the partner's pair operand)
    DevBuf centre, centre_part;   // mean-centred operand form: column means of the factor (fp32 [64]) and their partial sums
    int centre_chunks = 0;
    long long centre_rows_per_chunk = 0;
    DevBuf gram_part;
    int gram_chunks = 0, gram_rows_per_chunk = 0;
    double* gram_raw = nullptr;  // inside `small`
    DevBuf gram, P, pinv_work, info;
    DevBuf Nsum, Dsum;           // T k*k
    DevBuf thP, thN;             // T [m_loc][k]
    std::vector<int> thetas, row_rels, col_rels;
    DevBuf upd_terms, upd_adds, sum_ptrs;
    int n_terms = 0, n_adds = 0;
  };
  // bf16 planes of an fp32 matrix (FZ_BF16X3): P0 + P1 + ... = the matrix, exactly; each plane has its own TMA descriptors
  // and is streamed like a bf16-stored relation.  n = 0: the matrix (or this sign part of it) is all zero.
  struct PlaneSet {
    DevBuf buf;              // [n][rows][ld] bf16
    int n = 0;
    int64_t ld = 0, stride = 0;
    CUtensorMap tmX[3], tmXT[3];     // box {64 cols, 128 rows} / {64 cols, 64 rows}
    DevBuf rowsum;           // constraint parts in the centred operand form: row sums of this part (fp32)
  };
  struct RelRec {
    int ti = 0, tj = 0, storage = FZ_F32;
    bool theta = false, borrowed = false;
    bool tc = false;         // streamed products on the tensor cores: bf16 storage, or fp32 master + bf16 planes (x3)
    bool x3 = false;
    PlaneSet pl, pl_neg;     // the relation's planes (bf16 storage: the relation itself as plane 0); constraints: Theta+ / Theta-
    void* data = nullptr;
    int64_t ld = 0, rows_loc = 0, cols = 0;
    DevBuf own;
    uint8_t* mask = nullptr;
    int64_t mask_ld = 0;
    DevBuf mask_own;
    DevBuf A, B, Bloc, T1, E, Es, Cx;
    double* M_raw = nullptr;
    DevBuf M_part;
    int m_chunks = 0, m_rows_per_chunk = 0;
    DevBuf S, t2, t5, W1, W4, work;
    bool has_backbone = false;
    CUtensorMap tmX, tmXT, tmEs, tmB, tmA;
    bool has_tmB = false, has_tmA = false;
    DevBuf colsum_loc;           // sharded with a communicator: this rank's share of the column sums (colsum then holds the total)
    DevBuf rowsum, colsum;       // centred operand form: row / column sums of the stored relation (fp32), computed once
    bool sums_ready = false;
    int corr_chunks = 0, corr_rows_per_chunk = 0;   // first-order correction of M (single-term form): extra slots of M_part
    CUtensorMap tmX256, tmB16;   // fused kernel v4: relation box {64 cols, 256 rows}; B box {64 cols, 64 rows}, no swizzle
    bool has_tmB16 = false, v4_ok = false;
  };

  int device_;
  int world_ = 1, rank_ = 0;
  // ---- collectives inside the library (fz_comm_init): NCCL on its own stream, ordered against the products by events
  ncclComm_t comm_ = nullptr;
  cudaStream_t comm_stream_ = nullptr;
  cudaEvent_t ev_c0_ = nullptr, ev_c1_ = nullptr, ev_gram_ = nullptr, ev_gram_done_ = nullptr;
  std::vector<cudaEvent_t> ev_upd_, ev_gather_, ev_rs_;
  std::vector<char> gather_pending_;
  // ---- reduce-scatter of the B partials over NVLink peer memory (peer_reduce.cuh); NCCL's stays the fall-back
  int peer_mode_ = -1;          // -1 not set up yet, 0 unavailable / disabled (FZ_PEER_RS=0), 1 active
  unsigned long long peer_epoch_ = 0;
  DevBuf peer_flags_;           // [2][world][n_rel] u64: arrive | consumed, written by the peers
  DevBuf peer_done_;            // [n_rel] u32: finished-block counters of pull_reduce
  PeerFlags peer_flag_ptrs_;
  std::vector<PeerPtrs> peer_B_;          // per relation: every rank's partial buffer as mapped in this process
  std::vector<void*> peer_opened_;        // IPC mappings to close
  cudaEvent_t ev_sig_ = nullptr;
  int64_t gram_count_ = 0;      // leading doubles of small_ that hold the Gram sums (all-reduced ahead of the rest)
  bool pinv_done_ = false;      // this iteration's pseudo-inverses already ran beside the products
  // operand forms of the NEXT iteration are built per type right behind that type's update (and all-gather), beside the
  // updates of the other types, instead of in front of the next iteration's first product
  bool presplit_valid_ = false, no_presplit_ = false;
  // the products of the CURRENT factors are already in the buffers (fz_objective ran them for its trace form): the next
  // iteration starts from them instead of streaming the relations again
  bool products_valid_ = false;
  // FZ_TIMELINE=1 (studies): CUDA events at the phase boundaries of every sharded iteration on the caller's stream; the mean
  // duration of each phase is printed when the handle is destroyed
  bool timeline_ = false;
  std::vector<std::vector<cudaEvent_t>> tl_events_;     // per iteration: begin, products issued, reductions joined, exchanged, chain, updated
  void tl_mark(cudaStream_t st, size_t it, int slot) {
    if (!timeline_) return;
    if (tl_events_.size() <= it) tl_events_.resize(it + 1);
    if (tl_events_[it].size() <= (size_t)slot) tl_events_[it].resize((size_t)slot + 1, nullptr);
    if (!tl_events_[it][(size_t)slot]) cudaEventCreate(&tl_events_[it][(size_t)slot]);
    cudaEventRecord(tl_events_[it][(size_t)slot], st);
  }
  void tl_report() {
    if (!timeline_ || tl_events_.size() < 4) return;
    const char* names[] = {"products (issue .. last product done)", "join fp64 reductions", "exchange + all-reduce", "backbone chain", "updates + gathers + next operand forms", "to next iteration"};
    double sum[6] = {0, 0, 0, 0, 0, 0};
    int n = 0;
    cudaDeviceSynchronize();
    for (size_t it = 2; it + 1 < tl_events_.size(); ++it) {       // skip the first (two-term, set-up) iterations
      if (tl_events_[it].size() < 6 || tl_events_[it + 1].empty()) continue;
      bool ok = true;
      float ms[6];
      for (int s = 0; s < 5 && ok; ++s) ok = cudaEventElapsedTime(&ms[s], tl_events_[it][s], tl_events_[it][s + 1]) == cudaSuccess;
      ok = ok && cudaEventElapsedTime(&ms[5], tl_events_[it][5], tl_events_[it + 1][0]) == cudaSuccess;
      if (!ok) { cudaGetLastError(); continue; }
      for (int s = 0; s < 6; ++s) sum[s] += ms[s];
      ++n;
    }
    if (n == 0) return;
    fprintf(stderr, "[fz timeline] rank %d, mean over %d iterations (ms):", rank_, n);
    for (int s = 0; s < 6; ++s) fprintf(stderr, "  %s %.3f;", names[s], sum[s] / n);
    fprintf(stderr, "\n");
  }
  size_t tl_it_ = 0;
  bool obj_exact_ = false;      // FZ_OBJ_EXACT=1: always the n_i x n_j form of the objective
  DevBuf rnorm2_, trace_jobs_, trace_out_;
  bool rnorm2_ready_ = false;
  cudaEvent_t ev_prep_ = nullptr;
  int terms_ = 2;           // split terms of the factor operand: 1..3 plain form; FZ_TERMS_AUTO / FZ_TERMS_CENTRED1: centred form
  int gs_terms_ = 2;        // terms stored in Gs (the centred forms always keep [hi | lo])
  bool centred_ = false;    // mean-centred operand form for the fused dfmf products (terms_ <= 0)
  bool dmma_ = true;        // fp64 Gram / G_i^T A reductions on the fp64 tensor cores (FZ_NO_DMMA=1: CUDA-core kernel)
  int dyn_sched_ = 0;       // dynamic tail of the single-term kernel's schedule (FZ_DYN_SCHED=1).  Off by default: alone on the GPU
                            // the tail costs ~4 % (profiles/r02_fused1_probe_hybrid_schedule.log); beside NCCL at 4 GPUs the kernel
                            // itself gets faster (0.69 -> 0.73 of the roofline) but the step does not -- the work it was sharing
                            // the SMs with still has to run (profiles/r02_bench_n4_{static,dynamic_tail}.log)
  DevBuf sched_ctr_;        // chunk counter of its dynamic tail
  int reserve_sms_ = 0;     // SMs the persistent single-term kernel leaves free (FZ_RESERVE_SMS): the fp64 reductions, the
                            // correction and the exchange kernels then run BESIDE the streamed product instead of delaying the
                            // CTAs of the next launch (a one-CTA-per-SM kernel with 225 KB of shared memory shares its SMs with nothing)
  bool no_corr_ = false;    // FZ_NO_CORR=1 (studies / tests only): single-term kernel WITHOUT the first-order correction of M
  bool single_now_ = false; // this iteration's fused products use the single-term kernel (umma_fused1.cuh) + M correction
  // ---- FZ_TERMS_AUTO: which kernel may run is decided from measurements (gate_measure / gate_decide)
  int64_t it_count_ = 0;            // dfmf iterations run on this handle
  bool gate_enabled_ = false;
  bool gate_single_ = false;        // decision in force for the coming iterations
  bool gate_check_now_ = false;     // the current iteration measures
  double gate_e_ = -1.0, gate_cond_est_ = -1.0, gate_pred_ = -1.0;
  int64_t n_single_ = 0, n_two_ = 0;
  DevBuf gate_probe_;               // 2 x [128][64] fp32: slab products with the residual term
  DevBuf gate_cond_;                // one double per type (pinv_spd)
  double* gate_slots_ = nullptr;    // inside small_ (so they are summed over the ranks): per relation numA, denA, numB, denB
  int64_t gate_slot_count_ = 0;
  int sm_count_ = 148;
  // auxiliary stream: the fp64 reductions (Gram, G_i^T A) run beside the streamed products and fill their tails
  cudaStream_t aux_ = nullptr;
  cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr;
  std::vector<cudaEvent_t> ev_rel_;
  bool use_aux_ = true;      // FZ_NO_AUX=1 keeps everything on the caller's stream
  // CUDA-graph replay of the iteration for launch-bound (small) graphs; FZ_NO_GRAPH=1 disables
  bool use_graph_ = true;
  cudaStream_t gstream_ = nullptr;
  cudaEvent_t ev_gfork_ = nullptr, ev_gjoin_ = nullptr;
  cudaGraphExec_t graph_exec_[2] = {nullptr, nullptr};
  int graph_parity_[2] = {-1, -1};
  int64_t graph_launches_[2] = {0, 0};
  bool fused_ = true;       // single-pass A/B kernel for bf16 relations (needs terms_ == 2); FZ_NO_FUSED=1 disables
  int fused_csplit_ = 0;    // column splits of the fused kernel (0 = automatic); FZ_FUSED_CSPLIT overrides
  int fused_ver_ = 3;       // 3: umma_fused.cuh (row-form accumulators), 4: umma_fused_t.cuh (transposed, TMEM-resident
                            // factor operand); FZ_FUSED_VER overrides
  bool finalized_ = false;
  bool dfmc_started_ = false;
  std::vector<std::unique_ptr<TypeRec>> types_;
  std::vector<std::unique_ptr<RelRec>> rels_;
  DevBuf small_;
  int64_t small_count_ = 0;
  DevBuf pinv_jobs_, bb_jobs_, sum_jobs_;
  DevBuf err_acc_;
  // transform state
  int tf_target_ = -1;
  DevBuf tf_Cp_, tf_Cn_;
  std::vector<int> tf_sum_types_;

 public:
  explicit Engine(int device) : device_(device) { this->device = device; }
  ~Engine() override {
    tl_report();
    for (auto& v : tl_events_) for (auto e : v) if (e) cudaEventDestroy(e);
    for (void* q : peer_opened_) cudaIpcCloseMemHandle(q);
    if (ev_sig_) cudaEventDestroy(ev_sig_);
    if (comm_) nccl_api().CommDestroy(comm_);
    if (comm_stream_) cudaStreamDestroy(comm_stream_);
    for (auto e : {ev_c0_, ev_c1_, ev_gram_, ev_gram_done_}) if (e) cudaEventDestroy(e);
    for (auto e : ev_upd_) cudaEventDestroy(e);
    for (auto e : ev_gather_) cudaEventDestroy(e);
    for (auto e : ev_rs_) cudaEventDestroy(e);
    if (ev_prep_) cudaEventDestroy(ev_prep_);
    if (aux_) cudaStreamDestroy(aux_);
    if (ev_fork_) cudaEventDestroy(ev_fork_);
    if (ev_join_) cudaEventDestroy(ev_join_);
    for (auto e : ev_rel_) cudaEventDestroy(e);
    for (auto g : graph_exec_) if (g) cudaGraphExecDestroy(g);
    if (gstream_) cudaStreamDestroy(gstream_);
    if (ev_gfork_) cudaEventDestroy(ev_gfork_);
    if (ev_gjoin_) cudaEventDestroy(ev_gjoin_);
  }
  int compute_dtype() const override { return kDT; }

  void set_shard(int world, int rank) override {
    if (!types_.empty()) FZ_THROW(FZ_ERR_INVALID, "fz_set_shard must precede fz_add_type");
    if (world < 1 || rank < 0 || rank >= world) FZ_THROW(FZ_ERR_INVALID, "bad shard %d/%d", rank, world);
    world_ = world;
    rank_ = rank;
  }

  // One communicator per handle; every rank of the shard group calls this with the same id (fz_comm_unique_id), each from
  // its own process or thread.  From then on fz_iterate / fz_objective run the collectives themselves.
  void comm_init(const void* unique_id) override {
    if (world_ < 2) FZ_THROW(FZ_ERR_INVALID, "fz_comm_init needs a sharded handle (fz_set_shard with world > 1)");
    if (comm_) FZ_THROW(FZ_ERR_INVALID, "communicator already initialised");
    if (unique_id == nullptr) FZ_THROW(FZ_ERR_INVALID, "null unique id");
    const NcclApi& nc = nccl_api();
    if (!nc.ok) FZ_THROW(FZ_ERR_UNSUPPORTED, "NCCL unavailable: %s", nc.error.c_str());
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    NCCL_OK(nc.CommInitRank(&comm_, world_, id, rank_));
    CUDA_OK(cudaStreamCreateWithFlags(&comm_stream_, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&ev_c0_, &ev_c1_, &ev_gram_, &ev_gram_done_}) CUDA_OK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }

  int add_type(int64_t n, int k) override {
    if (finalized_) FZ_THROW(FZ_ERR_INVALID, "engine already finalized");
    if (n <= 0 || k <= 0) FZ_THROW(FZ_ERR_INVALID, "object type needs n > 0 and rank > 0 (got %lld, %d)", (long long)n, k);
    auto t = std::make_unique<TypeRec>();
    t->n = n;
    t->k = k;
    t->m_loc = (n + world_ - 1) / world_;
    t->n_pad = t->m_loc * world_;
    t->row0 = (int64_t)rank_ * t->m_loc;
    t->rows_loc = std::max<int64_t>(0, std::min<int64_t>(n, t->row0 + t->m_loc) - t->row0);
    t->G[0].alloc((size_t)t->n_pad * k * sizeof(T));
    t->G[1].alloc((size_t)t->n_pad * k * sizeof(T));
    types_.push_back(std::move(t));
    return (int)types_.size() - 1;
  }

  int add_relation(int ti, int tj, const void* data, int64_t ld, int src, int mem, int storage, int borrow,
                   const uint8_t* mask, int64_t mask_ld, int mask_mem) override {
    if (finalized_) FZ_THROW(FZ_ERR_INVALID, "engine already finalized");
    if (ti < 0 || tj < 0 || ti >= (int)types_.size() || tj >= (int)types_.size()) FZ_THROW(FZ_ERR_INVALID, "unknown type id");
    if (data == nullptr) FZ_THROW(FZ_ERR_INVALID, "relation data is NULL");
    auto r = std::make_unique<RelRec>();
    r->ti = ti;
    r->tj = tj;
    r->theta = (ti == tj);
    TypeRec& Ti = *types_[ti];
    TypeRec& Tj = *types_[tj];
    r->rows_loc = Ti.rows_loc;
    r->cols = Tj.n;
    // exact tensor-core form: fp32 master (every element-wise kernel keeps reading it) + bf16 planes built in fz_finalize
    r->x3 = (storage == FZ_BF16X3);
    if (r->x3 && kDT != FZ_F32) FZ_THROW(FZ_ERR_UNSUPPORTED, "bf16x3 relations need the fp32 engine");
    if (storage != FZ_BF16) storage = kDT;  // CUDA-core path (and the x3 master) keep the relation in the compute dtype
    if (storage == FZ_BF16 && (r->theta || mask != nullptr)) storage = kDT;  // constraints / completion stay exact
    if (storage == FZ_BF16 && kDT != FZ_F32) FZ_THROW(FZ_ERR_UNSUPPORTED, "bf16 relations need the fp32 engine");
    r->storage = storage;
    r->tc = (storage == FZ_BF16) || r->x3;
    cudaStream_t st = 0;
    if (borrow && mem == FZ_HOST) {
      // Out-of-core relation: PINNED host memory used in place.  Every kernel that touches the relation (the TMA loads of the
      // streamed products first of all) reads it over PCIe through the unified address space, once per iteration -- for graphs
      // whose relations exceed HBM on the GPUs at hand.  Slow by construction (PCIe, not HBM, bounds it).
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, data) != cudaSuccess || attr.type != cudaMemoryTypeHost || attr.devicePointer == nullptr) {
        cudaGetLastError();
        FZ_THROW(FZ_ERR_INVALID, "a borrowed host relation must be pinned (page-locked, device-mapped) memory");
      }
      data = attr.devicePointer;
      mem = FZ_DEVICE;
    }
    if (borrow) {
      if (mem != FZ_DEVICE || src != storage) FZ_THROW(FZ_ERR_INVALID, "borrowed relations must be device (or pinned host) memory in the storage dtype");
      if (mask != nullptr) FZ_THROW(FZ_ERR_INVALID, "masked relations are rewritten by dfmc and cannot be borrowed");
      if (storage == FZ_BF16 && ((ld % 8) != 0 || ((uintptr_t)data & 15) != 0))
        FZ_THROW(FZ_ERR_INVALID, "borrowed bf16 relation needs 16-byte alignment and ld %% 8 == 0");
      r->data = const_cast<void*>(data);
      r->ld = ld;
      r->borrowed = true;
    } else {
      const int64_t ldo = (storage == FZ_BF16) ? ((r->cols + 63) / 64) * 64 : r->cols;   // bf16 rows pitched to 128 bytes (one L2 line per TMA box row)
      r->own.alloc((size_t)std::max<int64_t>(1, r->rows_loc) * ldo * dtype_size(storage));
      copy_in(data, ld, src, mem, r->own.p, ldo, storage, r->rows_loc, r->cols, st);
      r->data = r->own.p;
      r->ld = ldo;
    }
    if (mask != nullptr) {
      if (r->theta) FZ_THROW(FZ_ERR_INVALID, "constraint matrices cannot be masked");
      r->mask_own.alloc((size_t)std::max<int64_t>(1, r->rows_loc) * r->cols);
      copy_in(mask, mask_ld, FZ_U8, mask_mem, r->mask_own.p, r->cols, FZ_U8, r->rows_loc, r->cols, st);
      r->mask = r->mask_own.template as<uint8_t>();
      r->mask_ld = r->cols;
    }
    CUDA_OK(cudaStreamSynchronize(st));
    const int id = (int)rels_.size();
    if (r->theta) Ti.thetas.push_back(id);
    else {
      Ti.row_rels.push_back(id);
      Tj.col_rels.push_back(id);
    }
    if (r->tc) { Ti.need_gs = true; Tj.need_gs = true; }
    rels_.push_back(std::move(r));
    return id;
  }

  void set_factor(int t, const void* G0, int64_t ld, int src, int mem) override {
    if (t < 0 || t >= (int)types_.size()) FZ_THROW(FZ_ERR_INVALID, "unknown type id %d", t);
    TypeRec& Tt = *types_[t];
    copy_in(G0, ld, src, mem, Tt.G[Tt.cur].p, Tt.k, kDT, Tt.n, Tt.k, 0);
    CUDA_OK(cudaStreamSynchronize(0));
    Tt.has_factor = true;
    presplit_valid_ = false;
    products_valid_ = false;
  }

  void set_backbone(int rel, const void* S, int64_t ld, int src, int mem) override {
    if (!finalized_) FZ_THROW(FZ_ERR_INVALID, "fz_set_backbone needs a finalized engine");
    RelRec& r = relation(rel);
    if (r.theta) FZ_THROW(FZ_ERR_INVALID, "constraint matrices have no backbone");
    copy_in(S, ld, src, mem, r.S.p, types_[r.tj]->k, FZ_F64, types_[r.ti]->k, types_[r.tj]->k, 0);
    CUDA_OK(cudaStreamSynchronize(0));
    r.has_backbone = true;
  }

  void set_split_terms(int terms) override {
    if (finalized_) FZ_THROW(FZ_ERR_INVALID, "engine already finalized");
    if (terms != FZ_TERMS_AUTO && terms != FZ_TERMS_CENTRED1 && (terms < 1 || terms > 3))
      FZ_THROW(FZ_ERR_INVALID, "split terms must be 1..3, FZ_TERMS_AUTO or FZ_TERMS_CENTRED1");
    terms_ = terms;
  }

  // ---------------------------------------------------------------------------------------------
  void finalize() override {
    if (finalized_) return;
    if (types_.empty()) FZ_THROW(FZ_ERR_INVALID, "no object types");
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device_));
    const int sms = prop.multiProcessorCount;
    sm_count_ = sms;
    if (const char* nf = getenv("FZ_NO_FUSED")) fused_ = !(nf[0] == '1');
    for (auto& t : types_) {
      t->kp = ((t->k + kKp - 1) / kKp) * kKp;
      // the fused kernels hold one 64-column accumulator per product: factors of rank > 64 take the two-pass kernels over
      // 64-column blocks of the operand (umma), with the plain operand form
      if (t->need_gs && t->k > kKp) fused_ = false;
    }
    centred_ = (terms_ <= 0) && fused_ && kDT == FZ_F32;
    if (terms_ <= 0 && !centred_) terms_ = 2;       // the centred forms exist for the fused fp32-engine products only
    gs_terms_ = terms_ <= 0 ? 2 : terms_;
    if (const char* fv = getenv("FZ_FUSED_VER")) fused_ver_ = atoi(fv);
    if (fused_ver_ != 3 && fused_ver_ != 4) FZ_THROW(FZ_ERR_INVALID, "FZ_FUSED_VER must be 3 or 4");
    // fp64 all-reduce buffer: [gram_t ...][M_r ...]
    small_count_ = 0;
    for (auto& t : types_) small_count_ += (int64_t)t->k * t->k;
    for (auto& r : rels_)
      if (!r->theta) small_count_ += (int64_t)types_[r->ti]->k * types_[r->tj]->k;
    for (auto& t : types_) gram_count_ += (int64_t)t->k * t->k;
    int64_t gate_off = small_count_;
    if (centred_ && terms_ == FZ_TERMS_AUTO) {
      for (auto& r : rels_)
        if (!r->theta && r->tc) gate_slot_count_ += 4;
      small_count_ += gate_slot_count_;
      gate_probe_.alloc((size_t)2 * 128 * 64 * sizeof(float));
    }
    gate_cond_.alloc(types_.size() * 8);
    small_.alloc((size_t)small_count_ * 8);
    gate_slots_ = small_.template as<double>() + gate_off;
    int64_t off = 0;
    for (auto& tp : types_) {
      TypeRec& t = *tp;
      t.gram_raw = small_.template as<double>() + off;
      off += (int64_t)t.k * t.k;
      t.gram_rows_per_chunk = (int)std::max<int64_t>(64, (t.m_loc + 2 * sms - 1) / (2 * sms));
      t.gram_rows_per_chunk = ((t.gram_rows_per_chunk + 15) / 16) * 16;
      t.gram_chunks = (int)std::max<int64_t>(1, (t.m_loc + t.gram_rows_per_chunk - 1) / t.gram_rows_per_chunk);
      t.gram_part.alloc((size_t)t.gram_chunks * t.k * t.k * 8);
      t.gram.alloc((size_t)t.k * t.k * 8);
      t.P.alloc((size_t)t.k * t.k * 8);
      t.pinv_work.alloc((size_t)3 * t.k * t.k * 8);
      t.info.alloc(2 * sizeof(int));
      t.Nsum.alloc((size_t)t.k * t.k * sizeof(T));
      t.Dsum.alloc((size_t)t.k * t.k * sizeof(T));
      if (!t.thetas.empty()) {
        t.thP.alloc((size_t)t.m_loc * t.k * sizeof(T));
        t.thN.alloc((size_t)t.m_loc * t.k * sizeof(T));
      }
      if (t.need_gs && centred_) {
        t.centre.alloc(64 * sizeof(float));
        t.centre_rows_per_chunk = std::max<long long>(256, (t.n + 4 * sms - 1) / (4 * sms));
        t.centre_chunks = (int)std::max<long long>(1, (t.n + t.centre_rows_per_chunk - 1) / t.centre_rows_per_chunk);
        t.centre_part.alloc((size_t)t.centre_chunks * t.k * 8);
      }
      if (t.need_gs && fused_ver_ == 4 && gs_terms_ == 2 && kDT == FZ_F32 && fused_) {
        t.ldt = ((t.n_pad + 255) / 256) * 256 + 256;   // a CTA preloads 256 rows from any local row offset
        t.GsT.alloc((size_t)128 * t.ldt * 2);
        CUDA_OK(cudaMemset(t.GsT.p, 0, t.GsT.bytes));
        std::string e;
        if (!make_tmap_bf16_2d(&t.tmGT, t.GsT.p, 128, (uint64_t)t.ldt, (uint64_t)t.ldt, 64, 128, &e))
          FZ_THROW(FZ_ERR_CUDA, "%s", e.c_str());
      }
      if (t.need_gs) {
        t.Gs.alloc((size_t)t.n_pad * gs_terms_ * t.kp * 2);
        std::string e;
        if (!make_tmap_bf16_2d(&t.tmG, t.Gs.p, (uint64_t)t.n_pad, (uint64_t)gs_terms_ * t.kp, (uint64_t)gs_terms_ * t.kp, 64, 64, &e) ||
            !make_tmap_bf16_2d(&t.tmG128, t.Gs.p, (uint64_t)t.n_pad, (uint64_t)gs_terms_ * t.kp, (uint64_t)gs_terms_ * t.kp, 64, 128, &e))
          FZ_THROW(FZ_ERR_CUDA, "%s", e.c_str());
      }
    }
    for (auto& rp : rels_) {
      RelRec& r = *rp;
      if (r.theta) {
        if (r.x3) {      // Theta+ and Theta- as separate plane sets (both products are non-negative sums, _dfmf.py:284-292)
          build_planes(r, r.pl, +1, false, false);
          build_planes(r, r.pl_neg, -1, false, false);
          if (centred_) {      // the rank-1 parts of Theta+- G in the centred operand form
            for (PlaneSet* ps : {&r.pl, &r.pl_neg}) {
              if (ps->n == 0 || r.rows_loc <= 0) continue;
              ps->rowsum.alloc((size_t)r.rows_loc * sizeof(float));
              row_sums_part<<<nblk(r.rows_loc, 8), 256>>>((const float*)r.data, r.ld, r.rows_loc, r.cols, ps == &r.pl ? +1 : -1,
                                                         ps->rowsum.template as<float>());
            }
            CUDA_OK(cudaDeviceSynchronize());
          }
        }
        continue;
      }
      TypeRec& Ti = *types_[r.ti];
      TypeRec& Tj = *types_[r.tj];
      r.M_raw = small_.template as<double>() + off;
      off += (int64_t)Ti.k * Tj.k;
      r.A.alloc((size_t)std::max<int64_t>(1, Ti.m_loc) * Tj.k * sizeof(T));
      r.B.alloc((size_t)Tj.n_pad * Ti.k * sizeof(T));
      if (world_ > 1) r.Bloc.alloc((size_t)Tj.m_loc * Ti.k * sizeof(T));
      r.m_rows_per_chunk = Ti.gram_rows_per_chunk;
      r.m_chunks = Ti.gram_chunks;
      if (centred_ && r.tc) {
        r.corr_rows_per_chunk = (int)std::max<int64_t>(64, (Tj.n_pad + 2 * sms - 1) / (2 * sms));
        r.corr_rows_per_chunk = ((r.corr_rows_per_chunk + 15) / 16) * 16;
        r.corr_chunks = (int)std::max<int64_t>(1, (Tj.n_pad + r.corr_rows_per_chunk - 1) / r.corr_rows_per_chunk);
        r.rowsum.alloc((size_t)std::max<int64_t>(1, r.rows_loc) * sizeof(float));
        r.colsum.alloc((size_t)r.cols * sizeof(float));
      }
      r.M_part.alloc((size_t)(r.m_chunks + r.corr_chunks) * Ti.k * Tj.k * 8);
      r.S.alloc((size_t)Ti.k * Tj.k * 8);
      r.t2.alloc((size_t)Ti.k * Ti.k * 8);
      r.t5.alloc((size_t)Tj.k * Tj.k * 8);
      r.W1.alloc((size_t)Ti.k * Tj.k * sizeof(T));
      r.W4.alloc((size_t)Ti.k * Tj.k * sizeof(T));
      const int km = std::max(Ti.k, Tj.k);
      r.work.alloc((size_t)6 * km * km * 8);
      if (r.x3) {
        // masked relations keep all three planes: dfmc writes arbitrary fp32 values into the unknown entries
        build_planes(r, r.pl, 0, /*all_three=*/r.mask != nullptr, /*keep_one=*/true);
        r.tmX = r.pl.tmX[0];
        r.tmXT = r.pl.tmXT[0];
      }
      if (r.tc) {
        std::string e;
        if (!r.x3) {
          bool ok = make_tmap_bf16_2d(&r.tmX, r.data, (uint64_t)r.rows_loc, (uint64_t)r.cols, (uint64_t)r.ld, 64, 128, &e) &&
                    make_tmap_bf16_2d(&r.tmXT, r.data, (uint64_t)r.rows_loc, (uint64_t)r.cols, (uint64_t)r.ld, 64, 64, &e);
          if (!ok) FZ_THROW(FZ_ERR_CUDA, "%s", e.c_str());
          r.pl.n = 1;            // the relation itself is its only plane
          r.pl.tmX[0] = r.tmX;
          r.pl.tmXT[0] = r.tmXT;
        }
        // reduce target of the fused kernel's TMA flush (fp32, box 32 x 32, 128B swizzle)
        if (kDT == FZ_F32 && (Ti.k % 4) == 0 && Ti.k >= 32)
          r.has_tmB = make_tmap_f32_2d(&r.tmB, r.B.p, (uint64_t)Tj.n_pad, (uint64_t)Ti.k, (uint64_t)Ti.k, 32, 32, &e);
        if (kDT == FZ_F32 && centred_ && (Tj.k % 4) == 0 && Tj.k >= 32 && r.rows_loc >= 32)   // A reduce target of the single-term kernel
          r.has_tmA = make_tmap_f32_2d(&r.tmA, r.A.p, (uint64_t)r.rows_loc, (uint64_t)Tj.k, (uint64_t)Tj.k, 32, 32, &e);
        // v4 boxes are {64, 256} on the relation and {16, 64} on B: relations smaller than a box keep the v3 kernel
        if (fused_ver_ == 4 && kDT == FZ_F32 && !r.x3 && r.rows_loc >= 256 && r.cols >= 64 && Ti.GsT.p != nullptr && Tj.GsT.p != nullptr) {
          r.v4_ok = make_tmap_2d(&r.tmX256, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, r.data, (uint64_t)r.rows_loc, (uint64_t)r.cols,
                                 (uint64_t)r.ld, 64, 256, CU_TENSOR_MAP_SWIZZLE_128B, &e);
          if (r.v4_ok && (Ti.k % 4) == 0 && Tj.n_pad >= 64)      // one {64 k, 64 rows} reduce per chunk (flush mode 2)
            r.has_tmB16 = make_tmap_2d(&r.tmB16, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, r.B.p, (uint64_t)Tj.n_pad, (uint64_t)Ti.k,
                                       (uint64_t)Ti.k, 64, 64, CU_TENSOR_MAP_SWIZZLE_NONE, &e);
        }
      }
    }
    err_acc_.alloc(8);
    if (const char* cs = getenv("FZ_FUSED_CSPLIT")) fused_csplit_ = atoi(cs);
    if (const char* nc = getenv("FZ_NO_CORR")) no_corr_ = (nc[0] == '1');
    if (const char* nm = getenv("FZ_NO_DMMA")) dmma_ = !(nm[0] == '1');
    if (const char* oe = getenv("FZ_OBJ_EXACT")) obj_exact_ = (oe[0] == '1');
    if (const char* tl = getenv("FZ_TIMELINE")) timeline_ = (tl[0] == '1');
    if (const char* nd = getenv("FZ_DYN_SCHED")) dyn_sched_ = (nd[0] == '1') ? 1 : 0;
    if (const char* rs = getenv("FZ_RESERVE_SMS")) reserve_sms_ = std::max(0, std::min(sms / 2, atoi(rs)));
    sched_ctr_.alloc(64);
    if (gs_terms_ != 2) fused_ = false;
    if (const char* na = getenv("FZ_NO_AUX")) use_aux_ = !(na[0] == '1');
    if (const char* ng = getenv("FZ_NO_GRAPH")) use_graph_ = !(ng[0] == '1');
    if (use_graph_) {
      CUDA_OK(cudaStreamCreateWithFlags(&gstream_, cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&ev_gfork_, cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&ev_gjoin_, cudaEventDisableTiming));
    }
    if (use_aux_) {
      CUDA_OK(cudaStreamCreateWithFlags(&aux_, cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
    }
    ev_rel_.resize(rels_.size());
    for (auto& e : ev_rel_) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ev_prep_, cudaEventDisableTiming));
    ev_rs_.resize(rels_.size());
    for (auto& e : ev_rs_) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ev_upd_.resize(types_.size());
    ev_gather_.resize(types_.size());
    gather_pending_.assign(types_.size(), 0);
    for (auto& e : ev_upd_) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : ev_gather_) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    gate_enabled_ = centred_ && terms_ == FZ_TERMS_AUTO && !graph_worthwhile();   // small graphs are launch-bound: nothing to gain
    build_job_tables();
    // opt in to the large dynamic shared memory of the tensor-core kernels
    set_umma_attrs();
    CUDA_OK(cudaDeviceSynchronize());
    finalized_ = true;
  }

  // ---------------------------------------------------------------------------------------------
  void iterate(int algo, int n_iters, cudaStream_t st) override {
    need_final();
    if (world_ != 1) {
      if (!comm_) FZ_THROW(FZ_ERR_INVALID, "fz_iterate on a sharded handle needs fz_comm_init (or drive the fz_phase_* calls)");
      for (int it = 0; it < n_iters; ++it) run_one_sharded(algo, st);
      wait_gathers(st);
      return;
    }
    int done = 0;
    // Small graphs are launch-latency-bound (tens of kernels of a few microseconds each): after one eager
    // iteration, two iterations (the factor double-buffer has period two) are captured into a CUDA graph and
    // replayed.  Large graphs gain nothing from it and keep the per-launch profiling hooks.
    if (use_graph_ && !profile && n_iters >= 5 && graph_worthwhile()) {
      run_one(algo, st);
      ++done;
      const int pairs = (n_iters - done) / 2;
      if (pairs > 0) {
        CUDA_OK(cudaEventRecord(ev_gfork_, st));
        CUDA_OK(cudaStreamWaitEvent(gstream_, ev_gfork_, 0));
        const int parity = types_[0]->cur;
        if (graph_exec_[algo] == nullptr || graph_parity_[algo] != parity) {
          if (graph_exec_[algo]) { cudaGraphExecDestroy(graph_exec_[algo]); graph_exec_[algo] = nullptr; }
          const int64_t before = launches;
          const int64_t c_single = n_single_, c_two = n_two_, c_it = it_count_;     // capture executes nothing: undo its counting
          cudaGraph_t graph = nullptr;
          CUDA_OK(cudaStreamBeginCapture(gstream_, cudaStreamCaptureModeThreadLocal));
          try {
            run_one(algo, gstream_);
            run_one(algo, gstream_);
          } catch (...) {
            cudaStreamEndCapture(gstream_, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
          }
          CUDA_OK(cudaStreamEndCapture(gstream_, &graph));
          cudaError_t ie = cudaGraphInstantiate(&graph_exec_[algo], graph, 0);
          cudaGraphDestroy(graph);
          if (ie != cudaSuccess) FZ_THROW(FZ_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
          graph_launches_[algo] = launches - before;
          launches = before;                      // capture does not execute anything
          n_single_ = c_single;
          n_two_ = c_two;
          it_count_ = c_it;
          graph_parity_[algo] = parity;
        }
        for (int g = 0; g < pairs; ++g) CUDA_OK(cudaGraphLaunch(graph_exec_[algo], gstream_));
        launches += graph_launches_[algo] * pairs;
        done += 2 * pairs;
        if (algo == FZ_DFMF) {
          (single_now_ ? n_single_ : n_two_) += 2 * pairs;
          it_count_ += 2 * pairs;
        }
        CUDA_OK(cudaEventRecord(ev_gjoin_, gstream_));
        CUDA_OK(cudaStreamWaitEvent(st, ev_gjoin_, 0));
      }
    }
    for (; done < n_iters; ++done) run_one(algo, st);
  }

  void run_one(int algo, cudaStream_t st) {
    if (!(products_valid_ && algo == FZ_DFMF)) phase_products(algo, st);
    products_valid_ = false;
    phase_update(algo, st);
  }
  // One sharded iteration with the collectives on the communicator's stream: the reduce-scatter of relation
  // r's B partial runs under the streamed products of relation r+1, the Gram all-reduce and the pseudo-inverses under the
  // first products, the all-gather of type t's new factor under the update of type t+1.
  void run_one_sharded(int algo, cudaStream_t st) {
    const NcclApi& nc = nccl_api();
    const ncclDataType_t dt = (kDT == FZ_F32) ? ncclFloat32 : ncclFloat64;
    if (algo == FZ_DFMC) {
      // completion: the imputation R[M] <- (G_i S G_j^T)[M] is row-local (rows of R and G_i local, G_j whole), so dfmc shards
      // like dfmf; its B partials can only be exchanged after the imputation, inside phase_update (_dfmc.py:319-345)
      phase_products(algo, st);
      CUDA_OK(cudaEventRecord(ev_c0_, st));
      CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_c0_, 0));
      NCCL_OK(nc.AllReduce(small_.p, small_.p, (size_t)small_count_, ncclFloat64, ncclSum, comm_, comm_stream_));
      CUDA_OK(cudaEventRecord(ev_c1_, comm_stream_));
      CUDA_OK(cudaStreamWaitEvent(st, ev_c1_, 0));
      phase_update(algo, st);
      return;
    }
    if (algo != FZ_DFMF) FZ_THROW(FZ_ERR_INVALID, "unknown algorithm");
    if (!products_valid_) sharded_products(st);
    products_valid_ = false;
    phase_update(algo, st);
    tl_mark(st, tl_it_, 5);
    ++tl_it_;
  }
  void sharded_products(cudaStream_t st) {
    const int algo = FZ_DFMF;
    const NcclApi& nc = nccl_api();
    const ncclDataType_t dt = (kDT == FZ_F32) ? ncclFloat32 : ncclFloat64;
    if (peer_mode_ < 0) peer_setup(st);
    ++peer_epoch_;
    tl_mark(st, tl_it_, 0);
    phase_products_begin(algo, st);
    for (size_t r = 0; r < rels_.size(); ++r) {
      phase_product_relation(algo, (int)r, st);
      RelRec& rel = *rels_[r];
      if (rel.theta) continue;
      if (peer_rs(rel)) {
        // every rank pulls its rows out of the peers' partials over NVLink (+ the rank-1 part), no NCCL kernel involved
        TypeRec& Ti = *types_[rel.ti];
        TypeRec& Tj = *types_[rel.tj];
        signal_arrive<<<1, 32, 0, st>>>(peer_flag_ptrs_, world_, rank_, (int)rels_.size(), (int)r, peer_epoch_);
        CUDA_OK(cudaEventRecord(ev_sig_, st));
        CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_sig_, 0));
        const long long n_vec = Tj.rows_loc * (Ti.k / 4);
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(4ll * sm_count_, (n_vec + 255) / 256));
        wait_arrive<<<1, 32, 0, comm_stream_>>>(peer_flag_ptrs_.arrive[rank_], world_, (int)rels_.size(), (int)r, peer_epoch_);
        pull_reduce<<<grid, 256, 0, comm_stream_>>>(peer_B_[r], peer_flag_ptrs_, rel.Bloc.template as<float>(), Tj.row0, Tj.rows_loc, Ti.k,
                                                    centred_ ? rel.colsum.template as<float>() : nullptr,
                                                    centred_ ? Ti.centre.template as<float>() : nullptr, peer_flag_ptrs_.arrive[rank_],
                                                    peer_done_.template as<unsigned int>() + r, world_, rank_, (int)rels_.size(), (int)r,
                                                    peer_epoch_);
        launches += 3;
        CUDA_OK(cudaGetLastError());
      } else {
        CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_rel_[r], 0));          // this relation's B partial is complete
        NCCL_OK(nc.ReduceScatter(rel.B.p, rel.Bloc.p, (size_t)types_[rel.tj]->m_loc * types_[rel.ti]->k, dt, ncclSum, comm_, comm_stream_));
        if (centred_ && rel.tc) rank1_add_local(rel, comm_stream_);
      }
      if (corr_deferred()) {      // the correction of M reads this rank's reduce-scattered rows of B
        cudaStream_t fin = use_aux_ ? aux_ : st;
        CUDA_OK(cudaEventRecord(ev_rs_[r], comm_stream_));
        CUDA_OK(cudaStreamWaitEvent(fin, ev_rs_[r], 0));
        finish_M(rel, fin, /*local_rows=*/true);
      }
    }
    tl_mark(st, tl_it_, 1);
    phase_products_end(algo, st);                                          // joins the fp64 reductions into st
    tl_mark(st, tl_it_, 2);
    CUDA_OK(cudaEventRecord(ev_c0_, st));
    CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_c0_, 0));
    const int64_t first = pinv_done_ ? gram_count_ : 0;                    // the Gram sums went ahead (phase_products_begin)
    if (small_count_ > first)
      NCCL_OK(nc.AllReduce(small_.template as<double>() + first, small_.template as<double>() + first, (size_t)(small_count_ - first),
                           ncclFloat64, ncclSum, comm_, comm_stream_));
    CUDA_OK(cudaEventRecord(ev_c1_, comm_stream_));
    CUDA_OK(cudaStreamWaitEvent(st, ev_c1_, 0));                           // every reduce-scatter and the all-reduce have landed
    tl_mark(st, tl_it_, 3);
  }
  bool peer_rs(const RelRec& r) const { return peer_mode_ == 1 && !r.theta && r.tc && fused_ && kDT == FZ_F32 && (types_[r.ti]->k % 4) == 0; }
  // Map every rank's B partial buffers and flag arrays into this rank (once, at the first sharded iteration; collective).
  // Ranks in other processes are reached through CUDA IPC handles, ranks in this process through peer access; the handles
  // travel through the communicator itself (one all-gather).  Any failure on any rank leaves every rank on NCCL's reduce-scatter.
  void peer_setup(cudaStream_t st) {
    peer_mode_ = 0;
    const char* env = getenv("FZ_PEER_RS");
    int want = (env == nullptr || env[0] != '0') ? 1 : 0;
    if (world_ > kMaxPeers || kDT != FZ_F32) want = 0;
    const int n_rel = (int)rels_.size();
    struct Entry {
      long long pid;
      int device, ok;
      unsigned long long ptr[2];
      cudaIpcMemHandle_t handle[2];
    };
    // slot 0: flags, slot 1 + r: relation r's partial
    const size_t per_rank = sizeof(Entry) * (size_t)(1 + n_rel);
    std::vector<Entry> mine((size_t)(1 + n_rel));
    memset(mine.data(), 0, per_rank);
    if (want) {
      peer_flags_.alloc((size_t)2 * world_ * n_rel * 8);
      peer_done_.alloc((size_t)std::max(1, n_rel) * 4);
      ev_sig_ = nullptr;
      CUDA_OK(cudaEventCreateWithFlags(&ev_sig_, cudaEventDisableTiming));
    }
    for (int s = 0; s <= n_rel; ++s) {
      Entry& e = mine[(size_t)s];
      e.pid = (long long)getpid();
      e.device = device_;
      void* q = (s == 0) ? peer_flags_.p : rels_[(size_t)(s - 1)]->B.p;
      e.ok = (want && q != nullptr) ? 1 : 0;
      e.ptr[0] = (unsigned long long)(uintptr_t)q;
      if (e.ok && cudaIpcGetMemHandle(&e.handle[0], q) != cudaSuccess) { cudaGetLastError(); e.ok = 0; }
    }
    DevBuf xchg;
    xchg.alloc(per_rank * (size_t)world_);
    CUDA_OK(cudaMemcpyAsync((char*)xchg.p + per_rank * (size_t)rank_, mine.data(), per_rank, cudaMemcpyHostToDevice, st));
    all_gather_on(xchg.p, per_rank, ncclChar, st);
    std::vector<Entry> all((size_t)(1 + n_rel) * (size_t)world_);
    CUDA_OK(cudaMemcpyAsync(all.data(), xchg.p, per_rank * (size_t)world_, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    int ok = want;
    peer_B_.assign((size_t)n_rel, PeerPtrs());
    for (int p = 0; p < world_ && ok; ++p) {
      for (int s = 0; s <= n_rel && ok; ++s) {
        const Entry& e = all[(size_t)p * (1 + n_rel) + s];
        if (s > 0 && rels_[(size_t)(s - 1)]->theta) continue;
        if (!e.ok) { ok = 0; break; }
        void* q = nullptr;
        if (p == rank_) q = (void*)(uintptr_t)e.ptr[0];
        else if (e.pid == (long long)getpid()) {          // same process: plain peer access to the other handle's device
          int can = 0;
          if (cudaDeviceCanAccessPeer(&can, device_, e.device) != cudaSuccess || !can) { ok = 0; break; }
          cudaError_t pe = cudaDeviceEnablePeerAccess(e.device, 0);
          if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { ok = 0; break; }
          cudaGetLastError();
          q = (void*)(uintptr_t)e.ptr[0];
        } else {
          if (cudaIpcOpenMemHandle(&q, e.handle[0], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
          peer_opened_.push_back(q);
        }
        if (s == 0) {
          peer_flag_ptrs_.arrive[p] = (unsigned long long*)q;
          peer_flag_ptrs_.consumed[p] = (unsigned long long*)q + (size_t)world_ * n_rel;
        } else {
          peer_B_[(size_t)(s - 1)].B[p] = (const float*)q;
        }
      }
    }
    // every rank must agree: the minimum of the local verdicts
    DevBuf flag;
    flag.alloc(8);
    const double mine_ok = ok ? 1.0 : 0.0;
    CUDA_OK(cudaMemcpyAsync(flag.p, &mine_ok, 8, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaEventRecord(ev_c0_, st));
    CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_c0_, 0));
    NCCL_OK(nccl_api().AllReduce(flag.p, flag.p, 1, ncclFloat64, ncclMin, comm_, comm_stream_));
    CUDA_OK(cudaEventRecord(ev_c1_, comm_stream_));
    CUDA_OK(cudaStreamWaitEvent(st, ev_c1_, 0));
    double all_ok = 0.0;
    CUDA_OK(cudaMemcpyAsync(&all_ok, flag.p, 8, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    peer_mode_ = (all_ok > 0.5) ? 1 : 0;
    if (const char* gl = getenv("FZ_GATE_LOG"))
      if (gl[0] == '1') fprintf(stderr, "[fz peer] rank %d: reduce-scatter over %s\n", rank_, peer_mode_ ? "NVLink peer memory (pull)" : "NCCL");
  }
  // the B partial of relation r is (re)started at zero: only once every peer has pulled last iteration's from it
  void zero_partial(RelRec& r, cudaStream_t st) {
    if (peer_rs(r) && peer_epoch_ > 0) {
      const int rel = rel_index(r);
      const long long n_vec = (long long)(r.B.bytes / 16);
      zero_when_consumed<<<(unsigned)std::min<long long>(2 * sm_count_, (n_vec + 255) / 256), 256, 0, st>>>(
          (float4*)r.B.p, n_vec, peer_flag_ptrs_.consumed[rank_], world_, (int)rels_.size(), rel, peer_epoch_ - 1);
      ++launches;
    } else {
      CUDA_OK(cudaMemsetAsync(r.B.p, 0, r.B.bytes, st));
    }
  }
  int rel_index(const RelRec& r) const {
    for (size_t i = 0; i < rels_.size(); ++i)
      if (rels_[i].get() == &r) return (int)i;
    return -1;
  }
  void wait_gathers(cudaStream_t st) {
    for (size_t t = 0; t < gather_pending_.size(); ++t)
      if (gather_pending_[t]) { CUDA_OK(cudaStreamWaitEvent(st, ev_gather_[t], 0)); gather_pending_[t] = 0; }
  }
  // worth a graph when the relations are small enough that launch latency, not bandwidth, sets the pace
  bool graph_worthwhile() const {
    double entries = 0.0;
    for (auto& r : rels_) entries += (double)r->rows_loc * (double)r->cols;
    return entries <= 6.4e7;
  }

  void phase_products(int algo, cudaStream_t st) override {
    need_final();
    check_factors();
    if (algo == FZ_DFMC) {
      if (world_ != 1 && !comm_) FZ_THROW(FZ_ERR_UNSUPPORTED, "dfmc on a sharded handle needs fz_comm_init");
      wait_gathers(st);
      if (!dfmc_started_) {
        for (auto& rp : rels_)                                   // _dfmc.py:287-292
          if (rp->mask) {
            mask_zero<T><<<nblk(rp->rows_loc * rp->cols, 256), 256, 0, st>>>((T*)rp->data, rp->ld, rp->mask, rp->mask_ld,
                                                                              rp->rows_loc, rp->cols);
            ++launches;
            resplit_masked(*rp, st);
          }
        dfmc_started_ = true;
      }
      grams(st);
      for (auto& rp : rels_)
        if (!rp->theta) { product_A(*rp, st); reduce_M(*rp, st); }
      return;  // the second half (imputation, A/B on the completed R) runs in phase_update
    }
    phase_products_begin(algo, st);
    for (size_t r = 0; r < rels_.size(); ++r) phase_product_relation(algo, (int)r, st);
    phase_products_end(algo, st);
  }

  // The same phase in pieces, so that a sharded caller can start the reduce-scatter of one relation's B partial
  // while the next relation is still being streamed (skfusion/fusion/distributed.py).
  void phase_products_begin(int algo, cudaStream_t st) override {
    need_final();
    check_factors();
    if (algo != FZ_DFMF) FZ_THROW(FZ_ERR_UNSUPPORTED, "piecewise products are for dfmf");
    begin_flags(st, centred_ && choose_single());
    if (!use_aux_) { grams(st, centred_); return; }
    if (!presplit_valid_)
      for (auto& tp : types_)
        if (tp->need_gs) split(*tp, st, centred_);     // operand forms first: every streamed product needs them
    presplit_valid_ = false;
    begin_reductions(st);
  }
  bool gate_checks_next() const { return gate_enabled_ && (it_count_ < 4 || (it_count_ % 8) == 0); }
  void begin_flags(cudaStream_t st, bool single) {
    single_now_ = single;
    gate_check_now_ = gate_checks_next();
    wait_gathers(st);                                  // the factors updated by the previous iteration are whole again
    if (gate_slot_count_ > 0) CUDA_OK(cudaMemsetAsync(gate_slots_, 0, (size_t)gate_slot_count_ * 8, st));
    pinv_done_ = false;
  }
  void begin_reductions(cudaStream_t st) {
    CUDA_OK(cudaEventRecord(ev_fork_, st));
    CUDA_OK(cudaStreamWaitEvent(aux_, ev_fork_, 0));
    for (auto& tp : types_) gram_of(*tp, aux_);        // Gram matrices beside the first streamed products
    if (world_ == 1 || comm_) {
      // The pseudo-inverses depend on the Gram sums only: run them now, under the streamed products, instead of in the
      // serial tail behind the last product (sharded: their all-reduce goes ahead of the relations' reduce-scatters).
      if (comm_) {
        CUDA_OK(cudaEventRecord(ev_gram_, aux_));
        CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_gram_, 0));
        NCCL_OK(nccl_api().AllReduce(small_.p, small_.p, (size_t)gram_count_, ncclFloat64, ncclSum, comm_, comm_stream_));
        CUDA_OK(cudaEventRecord(ev_gram_done_, comm_stream_));
        CUDA_OK(cudaStreamWaitEvent(aux_, ev_gram_done_, 0));
      }
      pinv_spd<<<(unsigned)types_.size(), kChainThreads, kChainSmemBytes, aux_>>>(pinv_jobs_.template as<PinvJob>());
      ++launches;
      pinv_done_ = true;
    }
  }
  void phase_product_relation(int algo, int rel, cudaStream_t st) override {
    (void)algo;
    RelRec& r = relation(rel);
    if (r.theta) return;
    if (centred_ && comm_ && r.tc) ensure_sums(r, st);     // holds a collective: every rank, rows or not
    if (!product_AB_fused(r, st)) {
      product_A(r, st);
      product_B(r, st);
    }
    after_product(r, rel, st);
  }
  void after_product(RelRec& r, int rel, cudaStream_t st) {
    CUDA_OK(cudaEventRecord(ev_rel_[rel], st));        // A_ij and the B partial are complete here
    if (!use_aux_) { reduce_M(r, st, !corr_deferred()); if (gate_check_now_) gate_measure(r, rel, st); return; }
    CUDA_OK(cudaStreamWaitEvent(aux_, ev_rel_[rel], 0));
    reduce_M(r, aux_, !corr_deferred());               // G_i^T A_ij overlaps the next relation's stream
    if (gate_check_now_) gate_measure(r, rel, aux_);
  }
  void phase_products_end(int algo, cudaStream_t st) override {
    (void)algo;
    theta_products(st);
    if (use_aux_) {
      CUDA_OK(cudaEventRecord(ev_join_, aux_));
      CUDA_OK(cudaStreamWaitEvent(st, ev_join_, 0));
    }
    CUDA_OK(cudaGetLastError());
  }

  void phase_update(int algo, cudaStream_t st) override {
    need_final();
    const bool dfmf = (algo == FZ_DFMF);
    run_chain(/*solve=*/true, /*scrub=*/dfmf, st);
    if (dfmf && world_ > 1) tl_mark(st, tl_it_, 4);
    if (dfmf) {
      if (single_now_) ++n_single_; else ++n_two_;
      if (gate_check_now_) gate_decide(st);
      gate_check_now_ = false;
      ++it_count_;
    }
    if (!dfmf) {
      for (auto& rp : rels_) {                                   // _dfmc.py:319-325
        RelRec& r = *rp;
        if (r.theta || !r.mask) continue;
        TypeRec& Ti = *types_[r.ti];
        TypeRec& Tj = *types_[r.tj];
        ensure_T1(r);
        gemm(cur(Ti) + Ti.row0 * Ti.k, Ti.k, r.W4.template as<T>(), Tj.k, r.T1.template as<T>(), Tj.k, (int)r.rows_loc, Tj.k, Ti.k, false, st);
        dim3 g(nblk(r.cols, 32), nblk(r.rows_loc, 32));
        impute_masked<T><<<g, 256, 0, st>>>((T*)r.data, r.ld, r.mask, r.mask_ld, r.T1.template as<T>(), Tj.k, cur(Tj), Tj.k,
                                            r.rows_loc, r.cols, Tj.k);
        ++launches;
        resplit_masked(r, st);
      }
      for (auto& rp : rels_) {
        if (rp->theta) continue;
        if (rp->mask) product_A(*rp, st);   // only completed relations changed
        product_B(*rp, st);
      }
      if (comm_) {      // the B partials of the completed relations go to the rows' owners before the update reads them
        CUDA_OK(cudaEventRecord(ev_c0_, st));
        CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_c0_, 0));
        for (auto& rp : rels_) {
          if (rp->theta) continue;
          NCCL_OK(nccl_api().ReduceScatter(rp->B.p, rp->Bloc.p, (size_t)types_[rp->tj]->m_loc * types_[rp->ti]->k,
                                           (kDT == FZ_F32) ? ncclFloat32 : ncclFloat64, ncclSum, comm_, comm_stream_));
        }
        CUDA_OK(cudaEventRecord(ev_c1_, comm_stream_));
        CUDA_OK(cudaStreamWaitEvent(st, ev_c1_, 0));
      }
      theta_products(st);
    }
    const bool presplit = dfmf && centred_ && use_aux_ && !no_presplit_ && (world_ == 1 || comm_) && tf_target_ < 0;
    for (size_t t = 0; t < types_.size(); ++t) {
      update_type((int)t, dfmf ? 1 : 0, st);
      if (presplit && !comm_) CUDA_OK(cudaEventRecord(ev_upd_[t], st));
      if (comm_) {       // rows of the new factor go round while the next type updates (in-place all-gather)
        TypeRec& Tt = *types_[t];
        CUDA_OK(cudaEventRecord(ev_upd_[t], st));
        CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_upd_[t], 0));
        NCCL_OK(nccl_api().AllGather(nxt(Tt) + Tt.row0 * Tt.k, nxt(Tt), (size_t)Tt.m_loc * Tt.k, (kDT == FZ_F32) ? ncclFloat32 : ncclFloat64,
                                     comm_, comm_stream_));
        CUDA_OK(cudaEventRecord(ev_gather_[t], comm_stream_));
        gather_pending_[t] = 1;
      }
      if (presplit && types_[t]->need_gs) {     // next iteration's operand form of this type, beside the other types' updates
        CUDA_OK(cudaStreamWaitEvent(aux_, comm_ ? ev_gather_[t] : ev_upd_[t], 0));
        split(*types_[t], aux_, true, nxt(*types_[t]));
      }
    }
    if (presplit) {
      CUDA_OK(cudaEventRecord(ev_prep_, aux_));
      CUDA_OK(cudaStreamWaitEvent(st, ev_prep_, 0));
      presplit_valid_ = true;
    }
    for (auto& tp : types_) tp->cur ^= 1;
    CUDA_OK(cudaGetLastError());
  }

  // ---------------------------------------------------------------------------------------------
  // Two restarts batched into one pass over the relations (the reference fans n_run restarts out over
  // joblib workers, dfmf.py:87-95).  `this` is run 0, `other` run 1: same graph, same device, the relations of run 1 borrowed
  // from run 0 (fz_relation_device_ptr).  Both runs use the single-term centred operand form; one launch of the two-row-block
  // fused kernel in pair mode multiplies each relation tile with both runs' operands (umma_fused.cuh), everything else runs
  // per handle.  Iterations in which either run's accuracy gate measures, or refuses the single-term form, run unpaired.
  void pair_iterate(EngineBase* other_base, int n_iters, cudaStream_t st) override {
    need_final();
    Engine<T>* other = dynamic_cast<Engine<T>*>(other_base);
    if (other == nullptr || other == this) FZ_THROW(FZ_ERR_INVALID, "pair needs two distinct handles of the same compute dtype");
    other->need_final();
    if (world_ != 1 || other->world_ != 1) FZ_THROW(FZ_ERR_UNSUPPORTED, "batched restarts run on unsharded handles");
    if (device_ != other->device_ || types_.size() != other->types_.size() || rels_.size() != other->rels_.size())
      FZ_THROW(FZ_ERR_INVALID, "the two handles of a pair must describe the same graph on the same device");
    bool can_pair = centred_ && other->centred_ && use_aux_ && other->use_aux_ && fused_ && kDT == FZ_F32;
    for (size_t t = 0; t < types_.size() && can_pair; ++t)
      can_pair = types_[t]->n == other->types_[t]->n && types_[t]->k == other->types_[t]->k;
    for (size_t r = 0; r < rels_.size() && can_pair; ++r) {
      RelRec& a = *rels_[r];
      RelRec& b = *other->rels_[r];
      // (constraints on the tensor cores read the per-handle operand form, which a paired iteration does not build)
      can_pair = a.ti == b.ti && a.tj == b.tj && a.theta == b.theta &&
                 (a.theta ? (!a.tc && !b.tc) : (a.storage == FZ_BF16 && a.data == b.data && a.ld == b.ld));
    }
    check_factors();
    other->check_factors();
    // the pair's operand forms live side by side in one buffer (split_pair): no per-handle pre-splitting while paired
    no_presplit_ = other->no_presplit_ = can_pair;
    presplit_valid_ = other->presplit_valid_ = false;
    products_valid_ = other->products_valid_ = false;
    struct Restore { Engine<T>*a, *b; ~Restore() { a->no_presplit_ = b->no_presplit_ = false; } } restore{this, other};
    for (int it = 0; it < n_iters; ++it) {
      const bool paired = can_pair && choose_single() && other->choose_single() && !gate_checks_next() && !other->gate_checks_next();
      if (!paired) {
        run_one(FZ_DFMF, st);
        other->run_one(FZ_DFMF, st);
        continue;
      }
      begin_flags(st, true);
      other->begin_flags(st, true);
      for (size_t t = 0; t < types_.size(); ++t)
        if (types_[t]->need_gs) split_pair((int)t, *other, st);
      begin_reductions(st);
      other->begin_reductions(st);
      for (size_t r = 0; r < rels_.size(); ++r) {
        if (rels_[r]->theta) continue;
        product_pair(*other, (int)r, st);
        after_product(*rels_[r], (int)r, st);
        other->after_product(*other->rels_[r], (int)r, st);
      }
      phase_products_end(FZ_DFMF, st);
      other->phase_products_end(FZ_DFMF, st);
      phase_update(FZ_DFMF, st);
      other->phase_update(FZ_DFMF, st);
      ++n_paired_;
      ++other->n_paired_;
    }
  }
  void relation_device_ptr(int rel, void** ptr, int64_t* ld, int* dtype) override {
    RelRec& r = relation(rel);
    if (ptr) *ptr = r.data;
    if (ld) *ld = r.ld;
    if (dtype) *dtype = r.x3 ? (int)FZ_BF16X3 : r.storage;      // x3: the pointer is the fp32 master
  }
  int64_t n_paired_ = 0;
  // operand forms of both runs side by side: [n_pad][128] = [bf16(G0 - c0) | bf16(G1 - c1)]
  void split_pair(int t, Engine<T>& other, cudaStream_t st) {
    TypeRec& T0 = *types_[t];
    TypeRec& T1 = *other.types_[t];
    if (!T0.pairGs.p) {
      T0.pairGs.alloc((size_t)T0.n_pad * 2 * kKp * 2);
      std::string e;
      if (!make_tmap_bf16_2d(&T0.tmPair64, T0.pairGs.p, (uint64_t)T0.n_pad, 2 * kKp, 2 * kKp, 64, 64, &e) ||
          !make_tmap_bf16_2d(&T0.tmPair128, T0.pairGs.p, (uint64_t)T0.n_pad, 2 * kKp, 2 * kKp, 64, 128, &e))
        FZ_THROW(FZ_ERR_CUDA, "%s", e.c_str());
    }
    for (int run = 0; run < 2; ++run) {
      Engine<T>& E = run == 0 ? *this : other;
      TypeRec& Tt = run == 0 ? T0 : T1;
      col_sum_partial<T><<<Tt.centre_chunks, 256, 0, st>>>(E.cur(Tt), Tt.k, Tt.n, Tt.k, Tt.centre_rows_per_chunk, Tt.centre_part.template as<double>());
      finish_centre<<<1, 1024, 0, st>>>(Tt.centre_part.template as<double>(), Tt.centre_chunks, Tt.k, Tt.n, Tt.centre.template as<float>());
      __nv_bfloat16* dst = T0.pairGs.template as<__nv_bfloat16>() + run * kKp;
      split_factor_hi<T><<<nblk(Tt.n_pad * kKp, 256), 256, 0, st>>>(E.cur(Tt), Tt.k, dst, 2 * kKp, Tt.n, Tt.n_pad, Tt.k, kKp, Tt.centre.template as<float>());
      E.launches += 3;
      Tt.hi_ptr = dst;
      Tt.hi_ld = 2 * kKp;
    }
  }
  void product_pair(Engine<T>& other, int rel, cudaStream_t st);

  void comm_small(void** ptr, int64_t* count) override {
    need_final();
    *ptr = small_.p;
    *count = small_count_;
  }
  void comm_bpartial(int rel, void** full, void** local, int64_t* local_count, int* dtype) override {
    need_final();
    RelRec& r = relation(rel);
    if (r.theta) FZ_THROW(FZ_ERR_INVALID, "constraint matrices have no B partial");
    *full = r.B.p;
    *local = (world_ > 1) ? r.Bloc.p : r.B.p;
    *local_count = types_[r.tj]->m_loc * types_[r.ti]->k;
    *dtype = kDT;
  }
  void comm_factor(int t, void** full, int64_t* local_count, int* dtype) override {
    need_final();
    if (t < 0 || t >= (int)types_.size()) FZ_THROW(FZ_ERR_INVALID, "unknown type id %d", t);
    TypeRec& Tt = *types_[t];
    *full = Tt.G[Tt.cur].p;
    *local_count = Tt.m_loc * Tt.k;
    *dtype = kDT;
  }

  // ---------------------------------------------------------------------------------------------
  // transform (_dfmf.py:330-458): frozen G_j / S, loop-invariant relation terms computed once.
  void transform_prepare(int target, cudaStream_t st) override {
    need_final();
    check_factors();
    if (world_ != 1) FZ_THROW(FZ_ERR_UNSUPPORTED, "transform runs on replicas (rows are independent), not shards");
    if (target < 0 || target >= (int)types_.size()) FZ_THROW(FZ_ERR_INVALID, "unknown target type");
    TypeRec& Tt = *types_[target];
    tf_target_ = target;
    tf_Cp_.alloc((size_t)Tt.n * Tt.k * sizeof(T));
    tf_Cn_.alloc((size_t)Tt.n * Tt.k * sizeof(T));
    for (auto& rp : rels_) {
      if (rp->theta) { if (rp->ti != target) FZ_THROW(FZ_ERR_INVALID, "constraint on a non-target type in transform"); continue; }
      if (rp->ti != target && rp->tj != target) FZ_THROW(FZ_ERR_INVALID, "relation must include the target object type");
      if (!rp->has_backbone) FZ_THROW(FZ_ERR_INVALID, "relation without a backbone (fz_set_backbone)");
    }
    grams(st);
    run_chain(/*solve=*/false, /*scrub=*/false, st);   // t2 / t5 / W1 / W4 from the given S
    for (auto& rp : rels_) {
      RelRec& r = *rp;
      if (r.theta) continue;
      TypeRec& Ti = *types_[r.ti];
      TypeRec& Tj = *types_[r.tj];
      const bool row_role = (r.ti == target);
      TypeRec& To = row_role ? Tj : Ti;   // the frozen other type
      // E = G_other * (S^T or S)   (n_other x k_t)      _dfmf.py:394 / :408
      r.E.alloc((size_t)To.n * Tt.k * sizeof(T));
      const T* W = row_role ? r.W1.template as<T>() : r.W4.template as<T>();
      gemm(cur(To), To.k, W, Tt.k, r.E.template as<T>(), Tt.k, (int)To.n, Tt.k, To.k, false, st);
      r.Cx.alloc((size_t)Tt.n * Tt.k * sizeof(T));
      if (r.tc) {
        r.Es.alloc((size_t)To.n * gs_terms_ * Tt.kp * 2);
        split_factor<T><<<nblk(To.n * Tt.kp, 256), 256, 0, st>>>(r.E.template as<T>(), Tt.k, r.Es.template as<__nv_bfloat16>(),
                                                                  To.n, To.n, Tt.k, Tt.kp, gs_terms_);
        ++launches;
        std::string e;
        if (!make_tmap_bf16_2d(&r.tmEs, r.Es.p, (uint64_t)To.n, (uint64_t)gs_terms_ * Tt.kp, (uint64_t)gs_terms_ * Tt.kp, 64, 64, &e))
          FZ_THROW(FZ_ERR_CUDA, "%s", e.c_str());
        umma(r.pl, /*trans=*/!row_role, r.tmEs, 0, r.Cx.template as<T>(), Tt.k, (int)Tt.n, (int)To.n, Tt.k, false, st);
      } else {
        if (row_role) gemm((const T*)r.data, r.ld, r.E.template as<T>(), Tt.k, r.Cx.template as<T>(), Tt.k, (int)Tt.n, Tt.k, (int)To.n, false, st);
        else gemm_t((const T*)r.data, r.ld, r.E.template as<T>(), Tt.k, r.Cx.template as<T>(), Tt.k, (int)Tt.n, Tt.k, (int)To.n, st);
      }
      accum_sign_split<T><<<nblk(Tt.n * Tt.k, 256), 256, 0, st>>>(r.Cx.template as<T>(), tf_Cp_.template as<T>(),
                                                                   tf_Cn_.template as<T>(), Tt.n * Tt.k);
      ++launches;
    }
    CUDA_OK(cudaGetLastError());
  }

  void transform_iterate(int n_iters, cudaStream_t st) override {
    need_final();
    if (tf_target_ < 0) FZ_THROW(FZ_ERR_INVALID, "fz_transform_prepare first");
    TypeRec& Tt = *types_[tf_target_];
    bool tc_theta = false;
    for (int id : Tt.thetas) tc_theta = tc_theta || rels_[id]->tc;
    for (int it = 0; it < n_iters; ++it) {
      if (tc_theta) split(Tt, st, false);      // the target factor moves every iteration: so does its operand form
      theta_products(st);
      update_type(tf_target_, 0, st);
      Tt.cur ^= 1;
    }
    CUDA_OK(cudaGetLastError());
  }

  // ---------------------------------------------------------------------------------------------
  void get_factor(int t, void* dst, int64_t ld, int dd, int mem, cudaStream_t st) override {
    need_final();
    if (t < 0 || t >= (int)types_.size()) FZ_THROW(FZ_ERR_INVALID, "unknown type id %d", t);
    TypeRec& Tt = *types_[t];
    wait_gathers(st);
    copy_out(Tt.G[Tt.cur].p, Tt.k, kDT, dst, ld, dd, mem, Tt.n, Tt.k, st);
  }
  void get_backbone(int rel, void* dst, int64_t ld, int dd, int mem, cudaStream_t st) override {
    need_final();
    RelRec& r = relation(rel);
    if (r.theta) FZ_THROW(FZ_ERR_INVALID, "constraint matrices have no backbone");
    copy_out(r.S.p, types_[r.tj]->k, FZ_F64, dst, ld, dd, mem, types_[r.ti]->k, types_[r.tj]->k, st);
  }

  // Frobenius residuals ||R - G_i S G_j^T||_F per relation (_dfmf.py:306-319): every relation's squared sum is accumulated
  // on the device over the local rows, summed over the ranks by one all-reduce when sharded, and read back ONCE.
  void objective(double* per_rel, double* total, cudaStream_t st) override {
    need_final();
    if (world_ != 1 && !comm_) FZ_THROW(FZ_ERR_UNSUPPORTED, "objective on a sharded handle needs fz_comm_init");
    wait_gathers(st);
    int n_rel = 0;
    for (auto& rp : rels_) n_rel += rp->theta ? 0 : 1;
    if (n_rel == 0) { if (total) *total = 0.0; return; }
    if (objective_trace(per_rel, total, n_rel, st)) return;
    err_acc_.alloc((size_t)n_rel * 8, false);
    CUDA_OK(cudaMemsetAsync(err_acc_.p, 0, (size_t)n_rel * 8, st));
    int idx = 0;
    for (auto& rp : rels_) {
      RelRec& r = *rp;
      if (r.theta) continue;
      TypeRec& Ti = *types_[r.ti];
      TypeRec& Tj = *types_[r.tj];
      double* acc = err_acc_.template as<double>() + idx++;
      if (r.rows_loc <= 0) continue;
      ensure_T1(r);
      gemm(cur(Ti) + Ti.row0 * Ti.k, Ti.k, r.W4.template as<T>(), Tj.k, r.T1.template as<T>(), Tj.k, (int)r.rows_loc, Tj.k, Ti.k, false, st);
      dim3 g(nblk(r.cols, 32), nblk(r.rows_loc, 32));
      if (r.storage == FZ_BF16)
        recon_err<T, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)r.data, r.ld, r.T1.template as<T>(), Tj.k, cur(Tj),
                                                       Tj.k, r.rows_loc, r.cols, Tj.k, acc, nullptr, 0);
      else
        recon_err<T, T><<<g, 256, 0, st>>>((const T*)r.data, r.ld, r.T1.template as<T>(), Tj.k, cur(Tj), Tj.k, r.rows_loc, r.cols,
                                           Tj.k, acc, nullptr, 0);
      ++launches;
    }
    CUDA_OK(cudaGetLastError());
    if (comm_) {
      CUDA_OK(cudaEventRecord(ev_c0_, st));
      CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_c0_, 0));
      NCCL_OK(nccl_api().AllReduce(err_acc_.p, err_acc_.p, (size_t)n_rel, ncclFloat64, ncclSum, comm_, comm_stream_));
      CUDA_OK(cudaEventRecord(ev_c1_, comm_stream_));
      CUDA_OK(cudaStreamWaitEvent(st, ev_c1_, 0));
    }
    std::vector<double> sq((size_t)n_rel);
    CUDA_OK(cudaMemcpyAsync(sq.data(), err_acc_.p, (size_t)n_rel * 8, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    double sum = 0.0;
    for (int i = 0; i < n_rel; ++i) {
      const double e = std::sqrt(sq[i]);
      if (per_rel) per_rel[i] = e;
      sum += e;
    }
    if (total) *total = sum;
  }

  // The objective without the n_i x n_j pass: run the products of the CURRENT factors -- the very work the
  // next iteration starts with, which then skips it -- and evaluate ||R||^2 - 2 tr(S^T M) + tr(S^T Gram_i S Gram_j) per
  // relation in fp64.  dfmf handles only (dfmc rewrites R; transform has no live products).  Returns false -- the caller then
  // takes the exact n^2 form -- when the form does not apply or when a residual is so small against ||R|| (< 10 %) that the
  // difference of large numbers would lose it.
  bool objective_trace(double* per_rel, double* total, int n_rel, cudaStream_t st) {
    if (obj_exact_ || tf_target_ >= 0 || dfmc_started_ || it_count_ == 0) return false;
    for (auto& rp : rels_)
      if (rp->mask != nullptr) return false;
    if (!rnorm2_ready_) {
      rnorm2_.alloc((size_t)n_rel * 8);
      int idx = 0;
      for (auto& rp : rels_) {
        RelRec& r = *rp;
        if (r.theta) continue;
        double* dst = rnorm2_.template as<double>() + idx++;
        if (r.rows_loc <= 0) continue;
        const int chunks = (int)std::min<int64_t>(128, std::max<int64_t>(1, (r.rows_loc + 127) / 128));
        const int64_t rpc = (r.rows_loc + chunks - 1) / chunks;
        DevBuf part;
        part.alloc((size_t)chunks * r.cols * 8, false);
        dim3 g(nblk(r.cols, 256), chunks);
        if (r.storage == FZ_BF16) col_sumsq_partial<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)r.data, r.ld, r.rows_loc, r.cols, rpc, part.template as<double>());
        else col_sumsq_partial<T><<<g, 256, 0, st>>>((const T*)r.data, r.ld, r.rows_loc, r.cols, rpc, part.template as<double>());
        total_sum<<<1, 1024, 0, st>>>(part.template as<double>(), (long long)chunks * r.cols, dst);
        launches += 2;
        CUDA_OK(cudaStreamSynchronize(st));
      }
      if (comm_) {
        CUDA_OK(cudaEventRecord(ev_c0_, st));
        CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_c0_, 0));
        NCCL_OK(nccl_api().AllReduce(rnorm2_.p, rnorm2_.p, (size_t)n_rel, ncclFloat64, ncclSum, comm_, comm_stream_));
        CUDA_OK(cudaEventRecord(ev_c1_, comm_stream_));
        CUDA_OK(cudaStreamWaitEvent(st, ev_c1_, 0));
      }
      std::vector<TraceJob> jobs;
      trace_out_.alloc((size_t)n_rel * 8);
      idx = 0;
      for (auto& rp : rels_) {
        RelRec& r = *rp;
        if (r.theta) continue;
        TraceJob j;
        j.M = r.M_raw;
        j.gram_i = types_[r.ti]->gram_raw;
        j.gram_j = types_[r.tj]->gram_raw;
        j.S = r.S.template as<double>();
        j.rnorm2 = rnorm2_.template as<double>() + idx;
        j.work = r.work.template as<double>();
        j.out = trace_out_.template as<double>() + idx;
        j.ki = types_[r.ti]->k;
        j.kj = types_[r.tj]->k;
        jobs.push_back(j);
        ++idx;
      }
      trace_jobs_.alloc(jobs.size() * sizeof(TraceJob), false);
      CUDA_OK(cudaMemcpy(trace_jobs_.p, jobs.data(), jobs.size() * sizeof(TraceJob), cudaMemcpyHostToDevice));
      rnorm2_ready_ = true;
    }
    if (!products_valid_) {
      if (world_ != 1) sharded_products(st);
      else phase_products(FZ_DFMF, st);
      products_valid_ = true;
    }
    trace_objective<<<(unsigned)n_rel, kChainThreads, kChainSmemBytes, st>>>(trace_jobs_.template as<TraceJob>());
    ++launches;
    CUDA_OK(cudaGetLastError());
    std::vector<double> sq((size_t)n_rel), rn((size_t)n_rel);
    CUDA_OK(cudaMemcpyAsync(sq.data(), trace_out_.p, (size_t)n_rel * 8, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(rn.data(), rnorm2_.p, (size_t)n_rel * 8, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    for (int i = 0; i < n_rel; ++i)
      if (!(sq[i] >= 1e-2 * rn[i])) return false;     // residual below 10 % of ||R|| (or not a number): take the exact form
    double sum = 0.0;
    for (int i = 0; i < n_rel; ++i) {
      const double e = std::sqrt(sq[i]);
      if (per_rel) per_rel[i] = e;
      sum += e;
    }
    if (total) *total = sum;
    return true;
  }

  // completed relation G_i S_ij G_j^T (base.py:119-146)
  void complete(int rel, void* dst, int64_t ld, int dd, int mem, cudaStream_t st) override {
    need_final();
    RelRec& r = relation(rel);
    if (r.theta) FZ_THROW(FZ_ERR_INVALID, "constraint matrices cannot be completed");
    wait_gathers(st);
    product_GSG(r.ti, r.tj, r.S.template as<double>(), dst, ld, dd, mem, st);
  }
  // G_i M G_j^T for a caller-given k_i x k_j matrix M: chained profiles G_i (S_ab S_bc ...) G_j^T
  // (examples/dicty_chaining.py:40-53) and completions with any backbone
  void profile_product(int ti, int tj, const void* S, int64_t lds, int sd, int smem, void* dst, int64_t ld, int dd, int mem,
                       cudaStream_t st) override {
    need_final();
    if (ti < 0 || tj < 0 || ti >= (int)types_.size() || tj >= (int)types_.size()) FZ_THROW(FZ_ERR_INVALID, "unknown type id");
    if (S == nullptr || dst == nullptr) FZ_THROW(FZ_ERR_INVALID, "null matrix");
    check_factors();
    wait_gathers(st);
    const int ki = types_[ti]->k, kj = types_[tj]->k;
    DevBuf Sd;
    Sd.alloc((size_t)ki * kj * 8, false);
    copy_in(S, lds, sd, smem, Sd.p, kj, FZ_F64, ki, kj, st);
    product_GSG(ti, tj, Sd.template as<double>(), dst, ld, dd, mem, st);
    CUDA_OK(cudaStreamSynchronize(st));     // Sd dies here
  }

 private:
  // out (n_i x n_j) = G_i S G_j^T with S a device fp64 k_i x k_j matrix.  fp32 engine, ranks <= 64, unsharded: tensor cores
  // (umma_outer.cuh: two bf16 terms per operand, output-bound).  Otherwise the exact CUDA-core tile kernel.
  void product_GSG(int ti, int tj, const double* S_dev, void* dst, int64_t ld, int dd, int mem, cudaStream_t st);
  void product_GSG_simt(TypeRec& Ti, TypeRec& Tj, const T* W, void* dst, int64_t ld, int dd, int mem, cudaStream_t st) {
    DevBuf t1, out;
    t1.alloc((size_t)std::max<int64_t>(1, Ti.n) * Tj.k * sizeof(T), false);
    gemm(cur(Ti), Ti.k, W, Tj.k, t1.template as<T>(), Tj.k, (int)Ti.n, Tj.k, Ti.k, false, st);
    const bool direct = (mem == FZ_DEVICE && dd == kDT);
    if (!direct) out.alloc((size_t)std::max<int64_t>(1, Ti.n) * Tj.n * sizeof(T), false);
    T* C = direct ? (T*)dst : out.template as<T>();
    const int64_t ldc = direct ? ld : Tj.n;
    dim3 g(nblk(Tj.n, 32), nblk(Ti.n, 32));
    recon_err<T, T><<<g, 256, 0, st>>>(nullptr, 0, t1.template as<T>(), Tj.k, cur(Tj), Tj.k, Ti.n, Tj.n, Tj.k, nullptr, C, ldc);
    ++launches;
    CUDA_OK(cudaGetLastError());
    if (!direct) copy_out(out.p, Tj.n, kDT, dst, ld, dd, mem, Ti.n, Tj.n, st);
    CUDA_OK(cudaStreamSynchronize(st));
  }

 public:
  // ---------------------------------------------------------------------------------------------
  // Factor initialisation on the device (reference _init.py:20-61).  The host draws the
  // column samples with numpy's RandomState (bit-exact RNG consumption); the O(k n^2) part -- the means over the sampled
  // columns -- is the product  view(R) * W  with a 0/1 selection matrix W, run through the same streamed kernels as the
  // iteration.  G_t = value + sum over relations touching t of | view * W / p_c |.
  void init_fill(int t, double value, cudaStream_t st) override {
    need_final();
    if (t < 0 || t >= (int)types_.size()) FZ_THROW(FZ_ERR_INVALID, "unknown type id %d", t);
    TypeRec& Tt = *types_[t];
    fill_value<T><<<nblk(Tt.n * Tt.k, 256), 256, 0, st>>>(cur(Tt), (T)value, Tt.n * Tt.k);
    ++launches;
    CUDA_OK(cudaGetLastError());
  }
  // 2-norms of the columns (axis 0) or rows (axis 1) of a relation, fp64, into a HOST buffer (random_c ranks the
  // columns of the oriented relation by norm, _init.py:32-34)
  // Sharded handles (with their own communicator): column norms are summed over the ranks' row blocks, row norms are
  // all-gathered, so every rank returns the same full vector.
  void relation_norms(int rel, int axis, double* dst_host, cudaStream_t st) override {
    need_final();
    if (world_ != 1 && !comm_) FZ_THROW(FZ_ERR_UNSUPPORTED, "device initialisation on a sharded handle needs fz_comm_init");
    RelRec& r = relation(rel);
    if (r.theta) FZ_THROW(FZ_ERR_INVALID, "constraint matrices do not seed factors");
    if (axis != 0 && axis != 1) FZ_THROW(FZ_ERR_INVALID, "axis must be 0 (column norms) or 1 (row norms)");
    if (dst_host == nullptr) FZ_THROW(FZ_ERR_INVALID, "null destination");
    TypeRec& Ti = *types_[r.ti];
    const int64_t count = axis == 0 ? r.cols : Ti.n;
    DevBuf out;
    out.alloc((size_t)std::max<int64_t>(1, axis == 0 ? r.cols : Ti.n_pad) * 8);
    if (axis == 0) {
      const int chunks = (int)std::min<int64_t>(64, std::max<int64_t>(1, (r.rows_loc + 255) / 256));
      const int64_t rows_per_chunk = (r.rows_loc + chunks - 1) / chunks;
      DevBuf part;
      part.alloc((size_t)chunks * r.cols * 8);
      dim3 g(nblk(r.cols, 256), chunks);
      if (r.rows_loc > 0) {
        if (r.storage == FZ_BF16) col_sumsq_partial<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)r.data, r.ld, r.rows_loc, r.cols, rows_per_chunk, part.template as<double>());
        else col_sumsq_partial<T><<<g, 256, 0, st>>>((const T*)r.data, r.ld, r.rows_loc, r.cols, rows_per_chunk, part.template as<double>());
      }
      sum_chunks<<<nblk(r.cols, 256), 256, 0, st>>>(part.template as<double>(), out.template as<double>(), chunks, r.cols);
      if (comm_) all_reduce_on(out.p, (size_t)r.cols, ncclFloat64, st);
      sqrt_inplace<<<nblk(r.cols, 256), 256, 0, st>>>(out.template as<double>(), r.cols);
      launches += 3;
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaMemcpyAsync(dst_host, out.p, (size_t)count * 8, cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
    } else {
      double* mine = out.template as<double>() + Ti.row0;
      if (r.rows_loc > 0) {
        if (r.storage == FZ_BF16) row_norms<__nv_bfloat16><<<nblk(r.rows_loc, 8), 256, 0, st>>>((const __nv_bfloat16*)r.data, r.ld, r.rows_loc, r.cols, mine);
        else row_norms<T><<<nblk(r.rows_loc, 8), 256, 0, st>>>((const T*)r.data, r.ld, r.rows_loc, r.cols, mine);
        ++launches;
      }
      CUDA_OK(cudaGetLastError());
      if (comm_) all_gather_on(out.p, (size_t)Ti.m_loc, ncclFloat64, st);
      CUDA_OK(cudaMemcpyAsync(dst_host, out.p, (size_t)count * 8, cudaMemcpyDeviceToHost, st));
      CUDA_OK(cudaStreamSynchronize(st));
    }
  }
  // collectives of the set-up steps: on the communicator's stream, fenced on both sides against `st`
  void all_reduce_on(void* buf, size_t count, ncclDataType_t dt, cudaStream_t st) {
    CUDA_OK(cudaEventRecord(ev_c0_, st));
    CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_c0_, 0));
    NCCL_OK(nccl_api().AllReduce(buf, buf, count, dt, ncclSum, comm_, comm_stream_));
    CUDA_OK(cudaEventRecord(ev_c1_, comm_stream_));
    CUDA_OK(cudaStreamWaitEvent(st, ev_c1_, 0));
  }
  // in place: this rank's `count_per_rank` elements sit at offset rank * count_per_rank of `buf`
  void all_gather_on(void* buf, size_t count_per_rank, ncclDataType_t dt, cudaStream_t st) {
    const size_t es = (dt == ncclFloat64) ? 8 : (dt == ncclChar ? 1 : 4);
    CUDA_OK(cudaEventRecord(ev_c0_, st));
    CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_c0_, 0));
    NCCL_OK(nccl_api().AllGather((const char*)buf + (size_t)rank_ * count_per_rank * es, buf, count_per_rank, dt, comm_, comm_stream_));
    CUDA_OK(cudaEventRecord(ev_c1_, comm_stream_));
    CUDA_OK(cudaStreamWaitEvent(st, ev_c1_, 0));
  }
  // idx_host: [k_t][p_c] column indices of the oriented relation (t on the rows), one row per latent column.
  // Sharded: the means of a row-role relation are computed for the local rows and all-gathered, those of a column-role
  // relation are partial sums over the local rows and all-reduced (before the absolute value: it is not linear); every rank
  // then adds them to its full copy of the factor.
  void init_add_sampled_means(int t, int rel, const int32_t* idx_host, int p_c, cudaStream_t st) override {
    need_final();
    if (world_ != 1 && !comm_) FZ_THROW(FZ_ERR_UNSUPPORTED, "device initialisation on a sharded handle needs fz_comm_init");
    if (t < 0 || t >= (int)types_.size()) FZ_THROW(FZ_ERR_INVALID, "unknown type id %d", t);
    RelRec& r = relation(rel);
    if (r.theta || (r.ti != t && r.tj != t)) FZ_THROW(FZ_ERR_INVALID, "relation %d does not touch type %d", rel, t);
    if (p_c < 0 || (p_c > 0 && idx_host == nullptr)) FZ_THROW(FZ_ERR_INVALID, "bad sample plan");
    TypeRec& Tt = *types_[t];
    const bool row_role = (r.ti == t);
    TypeRec& To = row_role ? *types_[r.tj] : *types_[r.ti];     // the type the sampled columns index
    // W (n_other x k_t), ones at the sampled positions
    r.E.alloc((size_t)To.n_pad * Tt.k * sizeof(T));
    if (p_c > 0) {
      DevBuf idx;
      idx.alloc((size_t)Tt.k * p_c * sizeof(int32_t), false);
      CUDA_OK(cudaMemcpyAsync(idx.p, idx_host, (size_t)Tt.k * p_c * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      scatter_ones<T><<<nblk((long long)Tt.k * p_c, 256), 256, 0, st>>>(r.E.template as<T>(), Tt.k, (const int32_t*)idx.p, Tt.k, p_c, To.n);
      ++launches;
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaStreamSynchronize(st));                       // idx is freed at scope exit
    }
    r.Cx.alloc((size_t)Tt.n_pad * Tt.k * sizeof(T));            // zeroed: rows this rank does not compute stay 0
    // local rows of the relation: row role -> rows [row0, row0 + rows_loc) of the result; column role -> a partial of all rows
    T* C = r.Cx.template as<T>() + (row_role ? Tt.row0 * Tt.k : 0);
    const int M = (int)(row_role ? r.rows_loc : Tt.n);
    const int K = (int)(row_role ? To.n : r.rows_loc);
    if (r.rows_loc > 0) {
      if (r.tc) {
        r.Es.alloc((size_t)To.n_pad * gs_terms_ * Tt.kp * 2);
        split_factor<T><<<nblk(To.n_pad * Tt.kp, 256), 256, 0, st>>>(r.E.template as<T>(), Tt.k, r.Es.template as<__nv_bfloat16>(), To.n, To.n_pad,
                                                                      Tt.k, Tt.kp, gs_terms_);
        ++launches;
        std::string e;
        if (!make_tmap_bf16_2d(&r.tmEs, r.Es.p, (uint64_t)To.n_pad, (uint64_t)gs_terms_ * Tt.kp, (uint64_t)gs_terms_ * Tt.kp, 64, 64, &e))
          FZ_THROW(FZ_ERR_CUDA, "%s", e.c_str());
        umma(r.pl, /*trans=*/!row_role, r.tmEs, row_role ? 0 : To.row0, C, Tt.k, M, K, Tt.k, false, st);
      } else {
        const T* W = r.E.template as<T>() + (row_role ? 0 : To.row0 * Tt.k);
        if (row_role) gemm((const T*)r.data, r.ld, W, Tt.k, C, Tt.k, M, Tt.k, K, false, st);
        else gemm_t((const T*)r.data, r.ld, W, Tt.k, C, Tt.k, M, Tt.k, K, st);
      }
    }
    if (comm_) {
      const ncclDataType_t dt = (kDT == FZ_F32) ? ncclFloat32 : ncclFloat64;
      if (row_role) all_gather_on(r.Cx.p, (size_t)Tt.m_loc * Tt.k, dt, st);
      else all_reduce_on(r.Cx.p, (size_t)Tt.n * Tt.k, dt, st);
    }
    add_abs_mean<T><<<nblk(Tt.n * Tt.k, 256), 256, 0, st>>>(r.Cx.template as<T>(), cur(Tt), Tt.n * Tt.k, (T)p_c);
    ++launches;
    CUDA_OK(cudaGetLastError());
  }
  void operand_stats(int64_t* single_iters, int64_t* two_term_iters, int64_t* paired_iters, double* err_estimate, double* cond_estimate) override {
    if (single_iters) *single_iters = n_single_;
    if (two_term_iters) *two_term_iters = n_two_;
    if (paired_iters) *paired_iters = n_paired_;
    if (err_estimate) *err_estimate = gate_e_;
    if (cond_estimate) *cond_estimate = gate_cond_est_;
  }
  void init_end() override {
    need_final();
    CUDA_OK(cudaDeviceSynchronize());
    for (auto& rp : rels_) { rp->E.release(); rp->Es.release(); rp->Cx.release(); }
    for (auto& tp : types_) tp->has_factor = true;
    presplit_valid_ = false;
    products_valid_ = false;
  }

 private:
  // ---------------------------------------------------------------------------------------------
  void need_final() const {
    if (!finalized_) FZ_THROW(FZ_ERR_INVALID, "call fz_finalize first");
  }
  void check_factors() const {
    for (auto& t : types_)
      if (!t->has_factor) FZ_THROW(FZ_ERR_INVALID, "a factor was never set (fz_set_factor)");
  }
  RelRec& relation(int rel) {
    if (rel < 0 || rel >= (int)rels_.size()) FZ_THROW(FZ_ERR_INVALID, "unknown relation id %d", rel);
    return *rels_[rel];
  }
  T* cur(TypeRec& t) { return t.G[t.cur].template as<T>(); }
  T* nxt(TypeRec& t) { return t.G[t.cur ^ 1].template as<T>(); }
  void ensure_T1(RelRec& r) {
    if (!r.T1.p) r.T1.alloc((size_t)std::max<int64_t>(1, r.rows_loc) * types_[r.tj]->k * sizeof(T), false);
  }

  // C = X * Y (CUDA cores, exact in T)
  void gemm(const T* X, int64_t ldx, const T* Y, int64_t ldy, T* C, int64_t ldc, int M, int N, int K, bool accumulate,
            cudaStream_t st) {
    if (M <= 0 || N <= 0) return;
    dim3 g(nblk(M, kGemmBM), nblk(N, kGemmBN));
    gemm_simt<T, T, false, false><<<g, 256, 0, st>>>(X, ldx, Y, ldy, C, nullptr, ldc, M, N, K, accumulate ? 1 : 0);
    ++launches;
  }
  // C = X^T * Y, X is K x M
  void gemm_t(const T* X, int64_t ldx, const T* Y, int64_t ldy, T* C, int64_t ldc, int M, int N, int K, cudaStream_t st) {
    if (M <= 0 || N <= 0) return;
    dim3 g(nblk(M, kGemmBM), nblk(N, kGemmBN));
    gemm_simt<T, T, true, false><<<g, 256, 0, st>>>(X, ldx, Y, ldy, C, nullptr, ldc, M, N, K, 0);
    ++launches;
  }

  template <int N, bool TR>
  void umma_launch(const CUtensorMap& tx, const CUtensorMap& tg, const SkinnyParams& p, int ksplit, cudaStream_t st) {
    using Cfg = SkinnyCfg<N>;
    dim3 grid((p.M + kSkBM - 1) / kSkBM, ksplit);
    umma_skinny_kernel<N, TR><<<grid, kSkThreads, Cfg::kSmemBytes, st>>>(tx, tg, p);
    ++launches;
  }
  template <int N, bool TR>
  static void umma_attr() {
    cudaFuncSetAttribute(umma_skinny_kernel<N, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkinnyCfg<N>::kSmemBytes);
  }
  void set_umma_attrs() {
    umma_attr<64, false>(); umma_attr<64, true>();
    umma_attr<128, false>(); umma_attr<128, true>();
    umma_attr<192, false>(); umma_attr<192, true>();
    umma_attr<256, false>(); umma_attr<256, true>();
    cudaFuncSetAttribute(umma_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemBytes);
    cudaFuncSetAttribute(umma_fused1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kF1SmemBytes);
    cudaFuncSetAttribute(umma_fused_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFtSmemBytes);
    cudaFuncSetAttribute(pinv_spd, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemBytes);
    cudaFuncSetAttribute(trace_objective, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemBytes);
    cudaFuncSetAttribute(backbone_chain<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemBytes);
  }
  // tensor-core product of a relation (all its bf16 planes) with a split factor:  C (M x k) (+)= op(R) * Gs[g_row0 + ., :]
  // -- one launch per plane and per 64-column block of the factor, accumulating into C
  void umma(const PlaneSet& ps, bool trans, const CUtensorMap& tg, int64_t g_row0, T* C, int64_t ldc, int M, int K, int k,
            bool accumulate, cudaStream_t st);

  // FZ_BF16X3: split the fp32 master of a relation (part 0) or one sign part of a constraint matrix (+1 / -1) into its bf16
  // planes.  A first pass finds out how many planes hold anything (0/1 data, ratings, small integers: one), so only those are
  // allocated and streamed.  Set-up step: synchronises.
  unsigned plane_grid(int64_t rows) const { return (unsigned)std::max<int64_t>(1, std::min<int64_t>(rows, 16ll * sm_count_)); }
  void build_planes(RelRec& r, PlaneSet& ps, int part, bool all_three, bool keep_one) {
    if (kDT != FZ_F32) FZ_THROW(FZ_ERR_UNSUPPORTED, "bf16 planes need the fp32 engine");
    DevBuf need;
    need.alloc(4 * sizeof(unsigned int));
    unsigned int h[4] = {0, 0, 0, 0};
    if (r.rows_loc > 0) {
      planes_needed<<<plane_grid(r.rows_loc), 256>>>((const float*)r.data, r.ld, r.rows_loc, r.cols, part,
                                                             need.template as<unsigned int>());
      ++launches;
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaMemcpy(h, need.p, sizeof(h), cudaMemcpyDeviceToHost));
    }
    if (h[3]) FZ_THROW(FZ_ERR_INVALID, "a relation with non-finite entries cannot be split into bf16 planes (fill the unknown values first)");
    int n = h[2] ? 3 : (h[1] ? 2 : (h[0] ? 1 : 0));
    if (all_three) n = 3;
    if (keep_one && n == 0) n = 1;
    ps.n = n;
    if (n == 0) return;
    const int64_t rows = std::max<int64_t>(1, r.rows_loc);
    ps.ld = ((r.cols + 63) / 64) * 64;              // rows pitched to 128 bytes, like the engine's own bf16 copies
    ps.stride = rows * ps.ld;
    ps.buf.alloc((size_t)n * ps.stride * 2);        // zeroed: pad columns and absent rows read as 0
    if (r.rows_loc > 0) {
      split_planes<<<plane_grid(r.rows_loc), 256>>>((const float*)r.data, r.ld, ps.buf.template as<__nv_bfloat16>(), ps.ld,
                                                            ps.stride, n, r.rows_loc, r.cols, part);
      ++launches;
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaDeviceSynchronize());
    }
    std::string e;
    for (int q = 0; q < n; ++q) {
      const __nv_bfloat16* base = ps.buf.template as<__nv_bfloat16>() + (size_t)q * ps.stride;
      if (!make_tmap_bf16_2d(&ps.tmX[q], base, (uint64_t)rows, (uint64_t)r.cols, (uint64_t)ps.ld, 64, 128, &e) ||
          !make_tmap_bf16_2d(&ps.tmXT[q], base, (uint64_t)rows, (uint64_t)r.cols, (uint64_t)ps.ld, 64, 64, &e))
        FZ_THROW(FZ_ERR_CUDA, "%s", e.c_str());
    }
  }
  // dfmc rewrote the unknown entries of the master (mask_zero / impute_masked): bring the planes up to date there
  void resplit_masked(RelRec& r, cudaStream_t st) {
    if (!r.x3 || r.mask == nullptr || r.rows_loc <= 0) return;
    split_planes_masked<<<plane_grid(r.rows_loc), 256, 0, st>>>((const float*)r.data, r.ld, r.mask, r.mask_ld,
                                                                        r.pl.buf.template as<__nv_bfloat16>(), r.pl.ld, r.pl.stride, r.pl.n,
                                                                        r.rows_loc, r.cols);
    ++launches;
  }

  void split(TypeRec& t, cudaStream_t st, bool centred = false, const T* G = nullptr) {
    if (G == nullptr) G = cur(t);
    const float* centre = nullptr;
    if (centred) {     // G = 1 c^T + D with c the column means; the bf16 terms represent D
      col_sum_partial<T><<<t.centre_chunks, 256, 0, st>>>(G, t.k, t.n, t.k, t.centre_rows_per_chunk, t.centre_part.template as<double>());
      finish_centre<<<1, 1024, 0, st>>>(t.centre_part.template as<double>(), t.centre_chunks, t.k, t.n, t.centre.template as<float>());
      launches += 2;
      centre = t.centre.template as<float>();
    }
    split_factor<T><<<nblk(t.n_pad * t.kp, 256), 256, 0, st>>>(G, t.k, t.Gs.template as<__nv_bfloat16>(), t.n, t.n_pad, t.k, t.kp,
                                                                gs_terms_, centre);
    ++launches;
    t.hi_ptr = t.Gs.template as<__nv_bfloat16>();
    t.hi_ld = (long long)gs_terms_ * t.kp;
    t.gs_centred = centred;
    if (t.GsT.p != nullptr) {
      split_factor_t<T><<<nblk(t.n_pad, 64), 256, 0, st>>>(cur(t), t.k, t.GsT.template as<__nv_bfloat16>(), t.ldt, t.n, t.ldt, t.k);
      ++launches;
    }
  }

  // Gram matrices of the current factors over the local rows (+ bf16 operand form where needed)
  void gram_of(TypeRec& t, cudaStream_t st) {
    const T* Gl = cur(t) + t.row0 * t.k;
    dim3 g(t.gram_chunks, nblk(t.k, 64), nblk(t.k, 64));
    if (dmma_) gram_partial_dmma<T><<<g, 256, 0, st>>>(Gl, t.k, Gl, t.k, t.gram_part.template as<double>(), t.rows_loc, t.k, t.k,
                                                       t.gram_rows_per_chunk);
    else gram_partial<T><<<g, 256, 0, st>>>(Gl, t.k, Gl, t.k, t.gram_part.template as<double>(), t.rows_loc, t.k, t.k,
                                            t.gram_rows_per_chunk, 0);
    reduce_partials<<<nblk((long long)t.k * t.k, 32), kRedThreads, 0, st>>>(t.gram_part.template as<double>(), t.gram_raw, t.gram_chunks,
                                                                     (long long)t.k * t.k);
    launches += 2;
  }
  void grams(cudaStream_t st, bool centred = false) {
    for (auto& tp : types_) {
      if (tp->need_gs) split(*tp, st, centred);
      gram_of(*tp, st);
    }
  }

  void product_A(RelRec& r, cudaStream_t st) {   // A = R G_j   (local rows of type i)
    TypeRec& Tj = *types_[r.tj];
    if (r.rows_loc <= 0) return;
    if (r.tc) umma(r.pl, false, Tj.tmG, 0, r.A.template as<T>(), Tj.k, (int)r.rows_loc, (int)r.cols, Tj.k, false, st);
    else gemm((const T*)r.data, r.ld, cur(Tj), Tj.k, r.A.template as<T>(), Tj.k, (int)r.rows_loc, Tj.k, (int)r.cols, false, st);
  }
  void product_B(RelRec& r, cudaStream_t st) {   // B = R^T G_i[local rows]   (all rows of type j)
    TypeRec& Ti = *types_[r.ti];
    TypeRec& Tj = *types_[r.tj];
    if (r.rows_loc <= 0) {
      CUDA_OK(cudaMemsetAsync(r.B.p, 0, r.B.bytes, st));
      return;
    }
    if (r.tc) umma(r.pl, true, Ti.tmG, Ti.row0, r.B.template as<T>(), Ti.k, (int)r.cols, (int)r.rows_loc, Ti.k, false, st);
    else gemm_t((const T*)r.data, r.ld, cur(Ti) + Ti.row0 * Ti.k, Ti.k, r.B.template as<T>(), Ti.k, (int)r.cols, Ti.k, (int)r.rows_loc, st);
  }
  // A and B from ONE stream of a bf16 relation (umma_fused.cuh).  Returns false when not applicable.
  bool product_AB_fused(RelRec& r, cudaStream_t st);

  void reduce_M(RelRec& r, cudaStream_t st, bool finish = true) {    // M = G_i[local]^T A    (fp64 accumulate)
    TypeRec& Ti = *types_[r.ti];
    TypeRec& Tj = *types_[r.tj];
    dim3 g(r.m_chunks, nblk(Ti.k, 64), nblk(Tj.k, 64));
    if (dmma_) gram_partial_dmma<T><<<g, 256, 0, st>>>(cur(Ti) + Ti.row0 * Ti.k, Ti.k, r.A.template as<T>(), Tj.k,
                                                       r.M_part.template as<double>(), r.rows_loc, Ti.k, Tj.k, r.m_rows_per_chunk);
    else gram_partial<T><<<g, 256, 0, st>>>(cur(Ti) + Ti.row0 * Ti.k, Ti.k, r.A.template as<T>(), Tj.k, r.M_part.template as<double>(),
                                            r.rows_loc, Ti.k, Tj.k, r.m_rows_per_chunk, 0);
    ++launches;
    if (finish) finish_M(r, st, /*local_rows=*/false);
  }
  // second half of reduce_M: the first-order correction of the single-term form, then the fixed-order sum of all partials.
  // local_rows: the correction uses this rank's reduce-scattered rows of B (sharded handles with their own communicator:
  // 1 / world of the work, but it has to wait for the relation's reduce-scatter) instead of the full-height local partial.
  void finish_M(RelRec& r, cudaStream_t st, bool local_rows) {
    TypeRec& Ti = *types_[r.ti];
    TypeRec& Tj = *types_[r.tj];
    int chunks = r.m_chunks;
    if (single_now_ && !no_corr_ && r.tc) { corr_M(r, st, local_rows); chunks += r.corr_chunks; }
    reduce_partials<<<nblk((long long)Ti.k * Tj.k, 32), kRedThreads, 0, st>>>(r.M_part.template as<double>(), r.M_raw, chunks,
                                                                       (long long)Ti.k * Tj.k);
    ++launches;
  }
  bool corr_deferred() const { return comm_ != nullptr && single_now_ && !no_corr_; }
  // Sharded handles: the rank-1 part of R^T G_i in the centred form, colsum(R) c_i^T with the column sums over ALL ranks' rows
  // (all-reduced once per handle), is added to this rank's reduce-scattered rows of B -- 1 / world of initialising every
  // rank's full-height partial with its local share.
  void rank1_add_local(RelRec& r, cudaStream_t st);
  // Single-term operand form: G_i^T R G_j = G_i^T (R Gs_j + rowsum c_j^T) + (R^T G_i)^T lo_j, and R^T G_i is B up to second
  // order in the residuals.  The partials land behind the M partials and are summed with them in fp64.
  void corr_M(RelRec& r, cudaStream_t st, bool local_rows);
  // Which fused kernel runs this iteration's products (centred operand form only)
  bool choose_single() {
    if (terms_ == FZ_TERMS_CENTRED1) return true;
    if (const char* fs = getenv("FZ_FORCE_SINGLE")) return fs[0] == '1';
    return gate_enabled_ && gate_single_;
  }
  // The single-term form drops R lo_j from A and R^T lo_i from B (lo = the residual bf16 term).  Measure exactly that on a
  // slab of the relation -- its first 128 rows for A, its first 128 columns for B -- with the tensor cores themselves:
  // slots += { |R_slab lo_j|^2, |A_slab|^2, |R_slab^T lo_i|^2, |B_slab|^2 }.  The slots sit in the all-reduce buffer, so a
  // sharded run decides from the sums over the ranks and every rank decides alike.
  void gate_measure(RelRec& r, int rel, cudaStream_t st);
  // Predicted factor error of the single-term form = e (8 + 0.2 sqrt(cond)), e the largest measured relative operand-form
  // error, cond the largest Gram condition estimate: calibrated against the float64 oracle on the synthetic graphs of every
  // initialisation and on dicty (scripts/precision_study.py); single-term runs while it stays below
  // half the stated factor tolerance (1e-3).  One small D2H copy + stream sync per check.
  void gate_decide(cudaStream_t st) {
    std::vector<double> slots((size_t)gate_slot_count_), conds(types_.size());
    if (gate_slot_count_ > 0) CUDA_OK(cudaMemcpyAsync(slots.data(), gate_slots_, slots.size() * 8, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaMemcpyAsync(conds.data(), gate_cond_.p, conds.size() * 8, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    double e = 0.0, cond = 1.0;
    for (size_t i = 0; i + 1 < slots.size(); i += 2) {
      const double num = slots[i], den = slots[i + 1];
      if (den > 0.0) e = std::max(e, std::sqrt(num / den));
      else if (num > 0.0) e = 1.0;
    }
    for (double c : conds) cond = std::max(cond, (c == c) ? c : 1e300);
    gate_e_ = e;
    gate_cond_est_ = cond;
    gate_pred_ = e * (8.0 + 0.2 * std::sqrt(cond));
    gate_single_ = gate_pred_ <= 5e-4;
    if (const char* gl = getenv("FZ_GATE_LOG"))
      if (gl[0] == '1')
        fprintf(stderr, "[fz gate] iteration %lld: operand-form error %.3g, cond %.3g, predicted %.3g -> %s\n", (long long)it_count_, e,
                cond, gate_pred_, gate_single_ ? "single-term" : "two-term");
  }
  // row / column sums of a stored bf16 relation (the relation never changes during dfmf): once per handle
  void ensure_sums(RelRec& r, cudaStream_t st) {
    if (r.sums_ready) return;
    if (r.rows_loc > 0) {
      if (r.storage == FZ_BF16) row_sums<__nv_bfloat16><<<nblk(r.rows_loc, 8), 256, 0, st>>>((const __nv_bfloat16*)r.data, r.ld, r.rows_loc, r.cols, r.rowsum.template as<float>());
      else row_sums<T><<<nblk(r.rows_loc, 8), 256, 0, st>>>((const T*)r.data, r.ld, r.rows_loc, r.cols, r.rowsum.template as<float>());     // x3: the fp32 master
      const int chunks = (int)std::min<int64_t>(128, std::max<int64_t>(1, (r.rows_loc + 127) / 128));
      const int64_t rpc = (r.rows_loc + chunks - 1) / chunks;
      DevBuf part;
      part.alloc((size_t)chunks * r.cols * 8, false);
      dim3 g(nblk(r.cols, 256), chunks);
      if (r.storage == FZ_BF16) col_sums_partial<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)r.data, r.ld, r.rows_loc, r.cols, rpc, part.template as<double>());
      else col_sums_partial<T><<<g, 256, 0, st>>>((const T*)r.data, r.ld, r.rows_loc, r.cols, rpc, part.template as<double>());
      finish_col_sums<<<nblk(r.cols, 256), 256, 0, st>>>(part.template as<double>(), r.colsum.template as<float>(), chunks, r.cols);
      launches += 3;
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaStreamSynchronize(st));   // the partial buffer dies here (once per relation and handle)
    } else {
      CUDA_OK(cudaMemsetAsync(r.colsum.p, 0, r.colsum.bytes, st));
    }
    if (comm_) {     // column sums over every rank's rows (rank1_add_local); every rank gets here at the same relation
      r.colsum_loc.alloc(r.colsum.bytes, false);
      CUDA_OK(cudaMemcpyAsync(r.colsum_loc.p, r.colsum.p, r.colsum.bytes, cudaMemcpyDeviceToDevice, st));
      CUDA_OK(cudaEventRecord(ev_c0_, st));
      CUDA_OK(cudaStreamWaitEvent(comm_stream_, ev_c0_, 0));
      NCCL_OK(nccl_api().AllReduce(r.colsum.p, r.colsum.p, (size_t)r.cols, ncclFloat32, ncclSum, comm_, comm_stream_));
      CUDA_OK(cudaEventRecord(ev_c1_, comm_stream_));
      CUDA_OK(cudaStreamWaitEvent(st, ev_c1_, 0));
    }
    r.sums_ready = true;
  }
  void theta_products(cudaStream_t st) {         // Theta+ G -> den, Theta- G -> num   (_dfmf.py:284-292)
    for (auto& tp : types_) {
      TypeRec& t = *tp;
      bool first = true;
      for (int id : t.thetas) {
        RelRec& r = *rels_[id];
        if (r.rows_loc <= 0) continue;
        if (r.tc) {
          theta_tc(t, r, first, st);      // bf16 planes of Theta+ / Theta- on the tensor cores
        } else {
          dim3 g(nblk(r.rows_loc, kGemmBM), nblk(t.k, kGemmBN));
          gemm_simt<T, T, false, true><<<g, 256, 0, st>>>((const T*)r.data, r.ld, cur(t), t.k, t.thP.template as<T>(),
                                                           t.thN.template as<T>(), t.k, (int)r.rows_loc, t.k, (int)r.cols, first ? 0 : 1);
          ++launches;
        }
        first = false;
      }
    }
  }
  // thP (+)= Theta+ G_t, thN (+)= Theta- G_t from the plane sets of an FZ_BF16X3 constraint matrix and the operand form of G_t
  // that the iteration's products use (in the centred form G = 1 c^T + D the rank-1 parts rowsum(Theta+-) c^T are added here)
  void theta_tc(TypeRec& t, RelRec& r, bool first, cudaStream_t st);

  void build_job_tables() {
    std::vector<PinvJob> pj;
    for (auto& tp : types_) {
      TypeRec& t = *tp;
      PinvJob j;
      j.gram_raw = t.gram_raw;
      j.gram = t.gram.template as<double>();
      j.P = t.P.template as<double>();
      j.work = t.pinv_work.template as<double>();
      j.info = t.info.template as<int>();
      j.cond = gate_cond_.template as<double>() + pj.size();
      j.k = t.k;
      pj.push_back(j);
    }
    pinv_jobs_.alloc(pj.size() * sizeof(PinvJob));
    CUDA_OK(cudaMemcpy(pinv_jobs_.p, pj.data(), pj.size() * sizeof(PinvJob), cudaMemcpyHostToDevice));

    std::vector<BackboneJob<T>> bj;
    for (auto& rp : rels_) {
      RelRec& r = *rp;
      if (r.theta) continue;
      TypeRec& Ti = *types_[r.ti];
      TypeRec& Tj = *types_[r.tj];
      BackboneJob<T> j;
      j.M_raw = r.M_raw;
      j.P_i = Ti.P.template as<double>();
      j.P_j = Tj.P.template as<double>();
      j.gram_i = Ti.gram.template as<double>();
      j.gram_j = Tj.gram.template as<double>();
      j.S = r.S.template as<double>();
      j.t2 = r.t2.template as<double>();
      j.t5 = r.t5.template as<double>();
      j.W1 = r.W1.template as<T>();
      j.W4 = r.W4.template as<T>();
      j.work = r.work.template as<double>();
      j.ki = Ti.k;
      j.kj = Tj.k;
      j.solve = 1;
      j.scrub = 1;
      bj.push_back(j);
    }
    bb_host_ = bj;
    bb_jobs_.alloc(std::max<size_t>(1, bj.size()) * sizeof(BackboneJob<T>));

    std::vector<TypeSumJob<T>> sj;
    for (auto& tp : types_) {
      TypeRec& t = *tp;
      std::vector<const double*> ptrs;
      for (int id : t.row_rels) ptrs.push_back(rels_[id]->t2.template as<double>());
      for (int id : t.col_rels) ptrs.push_back(rels_[id]->t5.template as<double>());
      t.sum_ptrs.alloc(std::max<size_t>(1, ptrs.size()) * sizeof(double*));
      if (!ptrs.empty()) CUDA_OK(cudaMemcpy(t.sum_ptrs.p, ptrs.data(), ptrs.size() * sizeof(double*), cudaMemcpyHostToDevice));
      TypeSumJob<T> j;
      j.mats = t.sum_ptrs.template as<const double*>();
      j.Nsum = t.Nsum.template as<T>();
      j.Dsum = t.Dsum.template as<T>();
      j.n_mats = (int)ptrs.size();
      j.k = t.k;
      sj.push_back(j);
      // update terms (device tables; X pointers are fixed for the lifetime of the handle)
      std::vector<UpdTerm<T>> terms;
      for (int id : t.row_rels) {
        RelRec& r = *rels_[id];
        UpdTerm<T> u;
        u.X = r.A.template as<T>();
        u.ldx = types_[r.tj]->k;
        u.W = r.W1.template as<T>();
        u.kx = types_[r.tj]->k;
        u.pad_ = 0;
        terms.push_back(u);
      }
      for (int id : t.col_rels) {
        RelRec& r = *rels_[id];
        UpdTerm<T> u;
        u.X = (world_ > 1) ? r.Bloc.template as<T>() : r.B.template as<T>();
        u.ldx = types_[r.ti]->k;
        u.W = r.W4.template as<T>();
        u.kx = types_[r.ti]->k;
        u.pad_ = 0;
        terms.push_back(u);
      }
      t.n_terms = (int)terms.size();
      t.upd_terms.alloc(std::max<size_t>(1, terms.size()) * sizeof(UpdTerm<T>));
      if (!terms.empty()) CUDA_OK(cudaMemcpy(t.upd_terms.p, terms.data(), terms.size() * sizeof(UpdTerm<T>), cudaMemcpyHostToDevice));
      std::vector<UpdAdd<T>> adds;
      if (!t.thetas.empty()) {
        UpdAdd<T> a;
        a.num = t.thN.template as<T>();
        a.den = t.thP.template as<T>();
        adds.push_back(a);
      }
      t.n_adds = (int)adds.size();
      t.upd_adds.alloc(2 * sizeof(UpdAdd<T>));
      if (!adds.empty()) CUDA_OK(cudaMemcpy(t.upd_adds.p, adds.data(), adds.size() * sizeof(UpdAdd<T>), cudaMemcpyHostToDevice));
    }
    sum_jobs_.alloc(sj.size() * sizeof(TypeSumJob<T>));
    CUDA_OK(cudaMemcpy(sum_jobs_.p, sj.data(), sj.size() * sizeof(TypeSumJob<T>), cudaMemcpyHostToDevice));
    bb_mode_ = -1;
  }

  std::vector<BackboneJob<T>> bb_host_;
  int bb_mode_ = -1;

  void run_chain(bool solve, bool scrub_small, cudaStream_t st) {
    const int mode = (solve ? 2 : 0) | (scrub_small ? 1 : 0);
    if (mode != bb_mode_ && !bb_host_.empty()) {
      for (auto& j : bb_host_) { j.solve = solve ? 1 : 0; j.scrub = scrub_small ? 1 : 0; }
      CUDA_OK(cudaMemcpyAsync(bb_jobs_.p, bb_host_.data(), bb_host_.size() * sizeof(BackboneJob<T>), cudaMemcpyHostToDevice, st));
      CUDA_OK(cudaStreamSynchronize(st));
      bb_mode_ = mode;
    }
    if (!pinv_done_) {
      pinv_spd<<<(unsigned)types_.size(), kChainThreads, kChainSmemBytes, st>>>(pinv_jobs_.template as<PinvJob>());
      ++launches;
    }
    pinv_done_ = false;
    if (!bb_host_.empty()) {
      backbone_chain<T><<<dim3((unsigned)bb_host_.size(), 2), kChainThreads, kChainSmemBytes, st>>>(bb_jobs_.template as<BackboneJob<T>>());
      ++launches;
    }
    type_sums<T><<<(unsigned)types_.size(), 256, 0, st>>>(sum_jobs_.template as<TypeSumJob<T>>());
    ++launches;
  }

  void update_type(int t, int scrub_terms, cudaStream_t st) {
    TypeRec& Tt = *types_[t];
    if (Tt.rows_loc <= 0) return;
    UpdArgs<T> a;
    a.G = cur(Tt) + Tt.row0 * Tt.k;
    a.Gnew = nxt(Tt) + Tt.row0 * Tt.k;
    a.ldg = Tt.k;
    a.rows = Tt.rows_loc;
    a.kt = Tt.k;
    a.scrub_terms = scrub_terms;
    a.Nsum = Tt.Nsum.template as<T>();
    a.Dsum = Tt.Dsum.template as<T>();
    if (tf_target_ >= 0) {
      // transform: no live relation terms; frozen sums + constraints enter additively
      std::vector<UpdAdd<T>> adds;
      UpdAdd<T> c;
      c.num = tf_Cp_.template as<T>();
      c.den = tf_Cn_.template as<T>();
      adds.push_back(c);
      if (!Tt.thetas.empty()) {
        UpdAdd<T> th;
        th.num = Tt.thN.template as<T>();
        th.den = Tt.thP.template as<T>();
        adds.push_back(th);
      }
      if (!tf_adds_uploaded_) {
        CUDA_OK(cudaMemcpyAsync(Tt.upd_adds.p, adds.data(), adds.size() * sizeof(UpdAdd<T>), cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaStreamSynchronize(st));
        tf_adds_uploaded_ = true;
      }
      a.n_terms = 0;
      a.n_adds = (int)adds.size();
    } else {
      a.n_terms = Tt.n_terms;
      a.n_adds = Tt.n_adds;
    }
    a.terms = Tt.upd_terms.template as<UpdTerm<T>>();
    a.adds = Tt.upd_adds.template as<UpdAdd<T>>();
    fused_update<T><<<nblk(Tt.rows_loc, kUpdRows), 256, 0, st>>>(a);
    ++launches;
  }
  bool tf_adds_uploaded_ = false;
};

template <>
void Engine<float>::umma(const PlaneSet& ps, bool trans, const CUtensorMap& tg, int64_t g_row0, float* C, int64_t ldc, int M, int K, int k,
                         bool accumulate, cudaStream_t st) {
  if (M <= 0) return;
  if (ps.n == 0 || K <= 0) {      // an all-zero matrix (e.g. the empty sign part of a constraint): the product is zero
    if (!accumulate) CUDA_OK(cudaMemsetAsync(C, 0, (size_t)M * ldc * sizeof(float), st));
    return;
  }
  SkinnyParams p;
  p.ldc = ldc;
  p.M = M;
  p.K = K;
  p.kp = kKp;
  p.terms = gs_terms_;
  p.g_row0 = (int)g_row0;
  const int kp_g = ((k + kKp - 1) / kKp) * kKp;      // padded rank of the operand: its terms are kp_g columns apart
  // factors of rank > 64 go through in column blocks: 128 columns per launch (accumulator N = 256) with two terms, else 64
  const int block_w = (k > kKp && gs_terms_ == 2) ? 2 * kKp : kKp;
  const int col_blocks = (kp_g + block_w - 1) / block_w;
  p.g_term_stride = kp_g;
  // split the reduction so that the grid covers the machine a few times over
  const int row_blocks = (M + kSkBM - 1) / kSkBM;
  int ksplit = 1;
  const int target_ctas = 148 * 2;
  if (row_blocks < target_ctas) ksplit = std::min((K + 511) / 512, (target_ctas + row_blocks - 1) / row_blocks);
  ksplit = std::max(1, ksplit);
  int kps = ((K + ksplit - 1) / ksplit + 63) / 64 * 64;
  ksplit = (K + kps - 1) / kps;
  p.k_per_split = kps;
  p.atomic = (ksplit > 1 || ps.n > 1 || accumulate) ? 1 : 0;
  if (p.atomic && !accumulate) CUDA_OK(cudaMemsetAsync(C, 0, (size_t)M * ldc * sizeof(float), st));
  prof_begin(st);
  for (int q = 0; q < ps.n; ++q) {
    const CUtensorMap& tx = trans ? ps.tmXT[q] : ps.tmX[q];
    for (int cb = 0; cb < col_blocks; ++cb) {
      const int w = std::min(block_w, kp_g - cb * block_w);      // 64 or 128 operand columns in this launch
      p.C = C + cb * block_w;
      p.k = std::min(w, k - cb * block_w);
      p.kp = w;
      p.chunks_per_term = w / kKp;
      p.g_col0 = cb * block_w;
      const int N = gs_terms_ * w;
      if (N == 64) { if (trans) umma_launch<64, true>(tx, tg, p, ksplit, st); else umma_launch<64, false>(tx, tg, p, ksplit, st); }
      else if (N == 128) { if (trans) umma_launch<128, true>(tx, tg, p, ksplit, st); else umma_launch<128, false>(tx, tg, p, ksplit, st); }
      else if (N == 192) { if (trans) umma_launch<192, true>(tx, tg, p, ksplit, st); else umma_launch<192, false>(tx, tg, p, ksplit, st); }
      else { if (trans) umma_launch<256, true>(tx, tg, p, ksplit, st); else umma_launch<256, false>(tx, tg, p, ksplit, st); }
    }
  }
  prof_end(st, 2.0 * (double)M * (double)K * ps.n * col_blocks);
}
template <>
bool Engine<float>::product_AB_fused(RelRec& r, cudaStream_t st) {
  if (!fused_ || !r.tc || r.rows_loc <= 0) return false;
  TypeRec& Ti = *types_[r.ti];
  TypeRec& Tj = *types_[r.tj];
  FusedParams p;
  p.A = r.A.template as<float>();
  p.B = r.B.template as<float>();
  p.lda = Tj.k;
  p.ldb = Ti.k;
  p.n_rows = (int)r.rows_loc;
  p.n_cols = (int)r.cols;
  p.k_a = Tj.k;
  p.k_b = Ti.k;
  p.gi_row0 = (int)Ti.row0;
  p.probe_skip_flush = 0;
  p.b_terms = 2;
  p.tma_flush = r.has_tmB ? 1 : 0;
  if (centred_) {
    ensure_sums(r, st);
    p.rowsum = r.rowsum.template as<float>();
    p.cj = Tj.centre.template as<float>();
  }
  const int rows_per_cta = single_now_ ? kF1Blocks * kF1Tile : 2 * kFuTile;
  const int pairs = (int)((r.rows_loc + rows_per_cta - 1) / rows_per_cta);
  const int tiles = (int)((r.cols + kFuTile - 1) / kFuTile);
  int splits = fused_csplit_;
  if (splits <= 0) {
    // Pick the column split that fills whole waves of one CTA per SM: score = wave efficiency x (1 - fixed per-CTA
    // cost of ~3 tile-times for pipeline fill, resident-operand load and the final A flush).
    double best = -1.0;
    splits = 1;
    const int max_splits = std::max(1, std::min(tiles, 64));
    for (int cand = 1; cand <= max_splits; ++cand) {
      const int tps = (tiles + cand - 1) / cand;
      const int eff_splits = (tiles + tps - 1) / tps;
      const double ctas = (double)pairs * eff_splits;
      const double waves = ctas / sm_count_;
      const double wave_eff = waves / std::ceil(waves);
      const double score = wave_eff * (1.0 - 3.0 / (tps + 3.0));
      if (score > best + 1e-9) { best = score; splits = eff_splits; }
    }
  }
  splits = std::max(1, std::min(splits, tiles));
  p.tiles_per_split = (tiles + splits - 1) / splits;
  splits = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  const int n_planes = r.pl.n;                           // FZ_BF16X3: one pass per bf16 plane, all reducing into the same A and B
  p.a_atomic = (splits > 1 || single_now_ || n_planes > 1) ? 1 : 0;     // the single-term kernel always reduces into A
  if (centred_ && comm_ == nullptr) {    // B starts from the rank-1 part colsum(R) c_i^T of the centred form
    rank1_init<<<nblk(Tj.n_pad * Ti.k, 256), 256, 0, st>>>(r.B.template as<float>(), Ti.k, Tj.n_pad, r.cols, Ti.k, r.colsum.template as<float>(),
                                                           Ti.centre.template as<float>());
    ++launches;
  } else {   // (sharded with its own communicator: the rank-1 part goes onto the reduce-scattered rows instead)
    zero_partial(r, st);
  }
  if (p.a_atomic) CUDA_OK(cudaMemsetAsync(r.A.p, 0, (size_t)r.rows_loc * Tj.k * sizeof(float), st));
  dim3 grid(pairs, splits);
  prof_begin(st);
  for (int plane = 0; plane < n_planes; ++plane) {
  const CUtensorMap& tmx = r.pl.tmX[plane];
  if (plane > 0) p.rowsum = nullptr;      // the rank-1 part rowsum(R) c_j^T (row sums of the WHOLE relation) enters once
  if (single_now_) {
    Fused1Params q;
    q.A = p.A; q.B = p.B; q.lda = p.lda; q.ldb = p.ldb; q.rowsum = p.rowsum; q.cj = p.cj;
    q.n_rows = p.n_rows; q.n_cols = p.n_cols; q.k_a = p.k_a; q.k_b = p.k_b; q.gi_row0 = p.gi_row0;
    q.tma_flush = (r.has_tmB ? 1 : 0) | (r.has_tmA ? 2 : 0);
    q.probe = 0;
    // persistent grid: one CTA per SM; three quarters of the (row group x column tile) units in equal static shares, the last
    // quarter in chunks handed out on demand (umma_fused1.cuh: F1Segments)
    const long long units = (long long)pairs * tiles;
    const unsigned ctas = (unsigned)std::max<long long>(1, std::min<long long>(std::max(1, sm_count_ - reserve_sms_), units));
    q.dyn_chunk = 0;
    q.work_counter = nullptr;
    if (dyn_sched_ == 1 && units >= 16ll * ctas) {
      q.dyn_chunk = (int)std::max<long long>(1, std::min<long long>(16, (units / 4) / (4ll * ctas)));
      q.work_counter = sched_ctr_.template as<int>();
      CUDA_OK(cudaMemsetAsync(q.work_counter, 0, sizeof(int), st));
    }
    umma_fused1_kernel<<<ctas, kF1Threads, kF1SmemBytes, st>>>(tmx, Tj.tmG128, Ti.tmG128, r.has_tmB ? r.tmB : tmx,
                                                               r.has_tmA ? r.tmA : tmx, q);
  } else if (fused_ver_ == 4 && r.v4_ok && !centred_) {
    FusedTParams q;
    q.A = p.A; q.B = p.B; q.lda = p.lda; q.ldb = p.ldb;
    q.GiT = Ti.GsT.template as<__nv_bfloat16>();
    q.ldt = Ti.ldt;
    q.n_rows = p.n_rows; q.n_cols = p.n_cols; q.k_a = p.k_a; q.k_b = p.k_b; q.gi_row0 = p.gi_row0;
    q.tiles_per_split = p.tiles_per_split; q.a_atomic = p.a_atomic;
    q.tma_flush = r.has_tmB16 ? 2 : 0;
    q.variant = 0;
    umma_fused_t_kernel<<<grid, kFtThreads, kFtSmemBytes, st>>>(r.tmX256, Tj.tmGT, r.has_tmB16 ? r.tmB16 : r.tmX256, q);
  } else {
    umma_fused_kernel<<<grid, kFuThreads, kFuSmemBytes, st>>>(tmx, Tj.tmG, Ti.tmG128, r.has_tmB ? r.tmB : tmx, r.has_tmB ? r.tmB : tmx, p);
  }
  ++launches;
  }
  prof_end(st, 2.0 * (double)r.rows_loc * (double)r.cols * n_planes, /*passes=*/1);
  return true;
}
template <>
bool Engine<double>::product_AB_fused(RelRec&, cudaStream_t) { return false; }

// S (fp64) -> the compute dtype, row-major k_i x k_j
template <class T>
__global__ void cast_small(const double* __restrict__ S, T* __restrict__ W, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) W[i] = (T)S[i];
}
template <>
void Engine<double>::product_GSG(int ti, int tj, const double* S_dev, void* dst, int64_t ld, int dd, int mem, cudaStream_t st) {
  product_GSG_simt(*types_[ti], *types_[tj], S_dev, dst, ld, dd, mem, st);
}
template <>
void Engine<float>::product_GSG(int ti, int tj, const double* S_dev, void* dst, int64_t ld, int dd, int mem, cudaStream_t st) {
  TypeRec& Ti = *types_[ti];
  TypeRec& Tj = *types_[tj];
  DevBuf W;
  W.alloc((size_t)Ti.k * Tj.k * sizeof(float), false);
  cast_small<float><<<nblk((long long)Ti.k * Tj.k, 256), 256, 0, st>>>(S_dev, W.template as<float>(), Ti.k * Tj.k);
  ++launches;
  const bool tensor = world_ == 1 && Ti.k <= kKp && Tj.k <= kKp && Ti.n >= 128 && Tj.n >= 128 && getenv("FZ_NO_OUTER") == nullptr;
  if (!tensor) {
    product_GSG_simt(Ti, Tj, W.template as<float>(), dst, ld, dd, mem, st);
    return;
  }
  // T1 = G_i S (n_i x k_j), then the operand forms of the output-bound product T1 G_j^T
  DevBuf t1, x3, y3, out;
  t1.alloc((size_t)Ti.n * Tj.k * sizeof(float), false);
  gemm(cur(Ti), Ti.k, W.template as<float>(), Tj.k, t1.template as<float>(), Tj.k, (int)Ti.n, Tj.k, Ti.k, false, st);
  x3.alloc((size_t)Ti.n * kOuK * 2, false);
  y3.alloc((size_t)Tj.n * kOuK * 2, false);
  outer_operand<float><<<nblk(Ti.n * 64, 256), 256, 0, st>>>(t1.template as<float>(), Tj.k, x3.template as<__nv_bfloat16>(), Ti.n, Ti.n, Tj.k, 0);
  outer_operand<float><<<nblk(Tj.n * 64, 256), 256, 0, st>>>(cur(Tj), Tj.k, y3.template as<__nv_bfloat16>(), Tj.n, Tj.n, Tj.k, 1);
  launches += 2;
  const bool direct = mem == FZ_DEVICE && dd == FZ_F32 && (ld % 4) == 0 && ((uintptr_t)dst & 15) == 0;
  if (!direct) out.alloc((size_t)Ti.n * (size_t)(((Tj.n + 3) / 4) * 4) * sizeof(float), false);
  float* C = direct ? (float*)dst : out.template as<float>();
  const int64_t ldc = direct ? ld : ((Tj.n + 3) / 4) * 4;
  CUtensorMap tx, ty, tc;
  std::string e;
  if (!make_tmap_bf16_2d(&tx, x3.p, (uint64_t)Ti.n, kOuK, kOuK, 64, 128, &e) || !make_tmap_bf16_2d(&ty, y3.p, (uint64_t)Tj.n, kOuK, kOuK, 64, 128, &e) ||
      !make_tmap_f32_2d(&tc, C, (uint64_t)Ti.n, (uint64_t)Tj.n, (uint64_t)ldc, 32, 32, &e))
    FZ_THROW(FZ_ERR_CUDA, "%s", e.c_str());
  OuterParams q;
  q.M = (int)Ti.n;
  q.N = (int)Tj.n;
  const long long units = (long long)((Ti.n + kOuTile - 1) / kOuTile) * ((Tj.n + kOuTile - 1) / kOuTile);
  const unsigned ctas = (unsigned)std::max<long long>(1, std::min<long long>(sm_count_, units));
  cudaFuncSetAttribute(umma_outer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kOuSmemBytes);
  umma_outer_kernel<<<ctas, kOuThreads, kOuSmemBytes, st>>>(tx, ty, tc, q);
  ++launches;
  CUDA_OK(cudaGetLastError());
  if (!direct) copy_out(out.p, ldc, FZ_F32, dst, ld, dd, mem, Ti.n, Tj.n, st);
  CUDA_OK(cudaStreamSynchronize(st));     // the operand buffers die here
}
template <>
void Engine<float>::product_pair(Engine<float>& other, int rel, cudaStream_t st) {
  RelRec& r0 = *rels_[rel];
  RelRec& r1 = *other.rels_[rel];
  TypeRec& Ti = *types_[r0.ti];
  TypeRec& Tj = *types_[r0.tj];
  TypeRec& Ti1 = *other.types_[r0.ti];
  TypeRec& Tj1 = *other.types_[r0.tj];
  ensure_sums(r0, st);
  FusedParams p;
  p.A = r0.A.template as<float>();
  p.A2 = r1.A.template as<float>();
  p.B = r0.B.template as<float>();
  p.B2 = r1.B.template as<float>();
  p.lda = Tj.k;
  p.ldb = Ti.k;
  p.n_rows = (int)r0.rows_loc;
  p.n_cols = (int)r0.cols;
  p.k_a = Tj.k;
  p.k_b = Ti.k;
  p.gi_row0 = 0;
  p.probe_skip_flush = 0;
  p.b_terms = 2;
  p.pair = 1;
  p.tma_flush = (r0.has_tmB && r1.has_tmB) ? 1 : 0;
  p.rowsum = r0.rowsum.template as<float>();
  p.cj = Tj.centre.template as<float>();
  p.cj2 = Tj1.centre.template as<float>();
  const int pairs = (int)((r0.rows_loc + 2 * kFuTile - 1) / (2 * kFuTile));
  const int tiles = (int)((r0.cols + kFuTile - 1) / kFuTile);
  int splits = 1;
  double best = -1.0;
  for (int cand = 1; cand <= std::max(1, std::min(tiles, 64)); ++cand) {      // whole waves of one CTA per SM (as product_AB_fused)
    const int tps = (tiles + cand - 1) / cand;
    const int eff = (tiles + tps - 1) / tps;
    const double waves = (double)pairs * eff / sm_count_;
    const double score = waves / std::ceil(waves) * (1.0 - 3.0 / (tps + 3.0));
    if (score > best + 1e-9) { best = score; splits = eff; }
  }
  p.tiles_per_split = (tiles + splits - 1) / splits;
  splits = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.a_atomic = splits > 1 ? 1 : 0;
  // each run's B starts from the rank-1 part of its own centred form; the column sums of the relation are shared
  rank1_init<<<nblk(Tj.n_pad * Ti.k, 256), 256, 0, st>>>(p.B, Ti.k, Tj.n_pad, r0.cols, Ti.k, r0.colsum.template as<float>(), Ti.centre.template as<float>());
  rank1_init<<<nblk(Tj.n_pad * Ti.k, 256), 256, 0, st>>>(p.B2, Ti.k, Tj.n_pad, r0.cols, Ti.k, r0.colsum.template as<float>(), Ti1.centre.template as<float>());
  if (p.a_atomic) {
    CUDA_OK(cudaMemsetAsync(p.A, 0, (size_t)r0.rows_loc * Tj.k * sizeof(float), st));
    CUDA_OK(cudaMemsetAsync(p.A2, 0, (size_t)r0.rows_loc * Tj.k * sizeof(float), st));
  }
  dim3 grid(pairs, splits);
  prof_begin(st);
  umma_fused_kernel<<<grid, kFuThreads, kFuSmemBytes, st>>>(r0.tmX, Tj.tmPair64, Ti.tmPair128, p.tma_flush ? r0.tmB : r0.tmX,
                                                            p.tma_flush ? r1.tmB : r0.tmX, p);
  launches += 3;
  // one pass yields both products of BOTH runs: the algorithmic bytes of the two fits it serves are twice what it streams
  prof_end(st, 2.0 * (double)r0.rows_loc * (double)r0.cols, /*passes=*/1);
  prof_alg_bytes += 2.0 * (double)r0.rows_loc * (double)r0.cols * (profile ? 1.0 : 0.0);
  CUDA_OK(cudaGetLastError());
}
template <>
void Engine<double>::product_pair(Engine<double>&, int, cudaStream_t) {
  FZ_THROW(FZ_ERR_UNSUPPORTED, "batched restarts need the fp32 engine");
}
template <>
void Engine<float>::corr_M(RelRec& r, cudaStream_t st, bool local_rows) {
  TypeRec& Ti = *types_[r.ti];
  TypeRec& Tj = *types_[r.tj];
  dim3 g(r.corr_chunks, nblk(Ti.k, 64), nblk(Tj.k, 64));
  const long long ldgs = Tj.hi_ld;
  const float* B = r.B.template as<float>();
  const float* G = cur(Tj);
  const __nv_bfloat16* Gs = Tj.hi_ptr;
  long long rows = Tj.n;
  int rows_per_chunk = r.corr_rows_per_chunk;
  if (local_rows) {       // this rank's rows of type j: B after the reduce-scatter, the matching rows of the factor
    B = r.Bloc.template as<float>();
    G += Tj.row0 * Tj.k;
    Gs += Tj.row0 * ldgs;
    rows = Tj.rows_loc;
    rows_per_chunk = (int)((((Tj.m_loc + r.corr_chunks - 1) / r.corr_chunks) + 15) / 16 * 16);
  }
  corr_partial<<<g, 256, 0, st>>>(B, Ti.k, G, Tj.k, Gs, ldgs, Tj.centre.template as<float>(),
                                  r.M_part.template as<double>() + (size_t)r.m_chunks * Ti.k * Tj.k, rows, Ti.k, Tj.k, rows_per_chunk);
  ++launches;
}
template <>
void Engine<double>::corr_M(RelRec&, cudaStream_t, bool) {}
template <>
void Engine<float>::rank1_add_local(RelRec& r, cudaStream_t st) {
  TypeRec& Ti = *types_[r.ti];
  TypeRec& Tj = *types_[r.tj];
  if (Tj.rows_loc <= 0) return;
  rank1_add<<<nblk(Tj.rows_loc * Ti.k, 256), 256, 0, st>>>(r.Bloc.template as<float>(), Ti.k, Tj.rows_loc, Ti.k,
                                                           r.colsum.template as<float>() + Tj.row0, Ti.centre.template as<float>());
  ++launches;
}
template <>
void Engine<double>::rank1_add_local(RelRec&, cudaStream_t) {}
template <>
void Engine<float>::theta_tc(TypeRec& t, RelRec& r, bool first, cudaStream_t st) {
  float* outs[2] = {t.thP.template as<float>(), t.thN.template as<float>()};      // den <- Theta+ G, num <- Theta- G
  const PlaneSet* sets[2] = {&r.pl, &r.pl_neg};
  for (int side = 0; side < 2; ++side) {
    umma(*sets[side], false, t.tmG, 0, outs[side], t.k, (int)r.rows_loc, (int)r.cols, t.k, /*accumulate=*/!first, st);
    if (t.gs_centred && sets[side]->n > 0) {
      if (sets[side]->rowsum.p == nullptr) FZ_THROW(FZ_ERR_INVALID, "centred operand form without the constraint's row sums");
      rank1_add<<<nblk(r.rows_loc * t.k, 256), 256, 0, st>>>(outs[side], t.k, r.rows_loc, t.k, sets[side]->rowsum.template as<float>(),
                                                             t.centre.template as<float>());
      ++launches;
    }
  }
}
template <>
void Engine<double>::theta_tc(TypeRec&, RelRec&, bool, cudaStream_t) {
  FZ_THROW(FZ_ERR_UNSUPPORTED, "tensor-core path needs the fp32 engine");
}
template <>
void Engine<float>::gate_measure(RelRec& r, int rel, cudaStream_t st) {
  if (r.theta || !r.tc || r.rows_loc <= 0) return;
  TypeRec& Ti = *types_[r.ti];
  TypeRec& Tj = *types_[r.tj];
  int slot = 0;
  for (int q = 0; q < rel; ++q)
    if (!rels_[q]->theta && rels_[q]->tc) slot += 4;
  float* probe_a = gate_probe_.template as<float>();
  float* probe_b = probe_a + 128 * 64;
  CUDA_OK(cudaMemsetAsync(probe_a, 0, (size_t)2 * 128 * 64 * sizeof(float), st));
  for (int side = 0; side < 2; ++side) {
    const bool trans = side == 1;
    SkinnyParams p;
    p.C = trans ? probe_b : probe_a;
    p.ldc = 64;
    p.M = (int)std::min<int64_t>(128, trans ? r.cols : r.rows_loc);
    p.K = (int)(trans ? r.rows_loc : r.cols);
    p.k = trans ? Ti.k : Tj.k;
    p.kp = kKp;
    p.terms = 1;
    p.g_row0 = trans ? (int)Ti.row0 : 0;
    p.g_col0 = kKp;                                   // the residual term
    int ksplit = std::max(1, std::min(sm_count_, (p.K + 1023) / 1024));
    const int kps = ((p.K + ksplit - 1) / ksplit + 63) / 64 * 64;
    ksplit = (p.K + kps - 1) / kps;
    p.k_per_split = kps;
    p.atomic = 1;
    dim3 grid(1, ksplit);
    if (trans) umma_skinny_kernel<64, true><<<grid, kSkThreads, SkinnyCfg<64>::kSmemBytes, st>>>(r.tmXT, Ti.tmG, p);
    else umma_skinny_kernel<64, false><<<grid, kSkThreads, SkinnyCfg<64>::kSmemBytes, st>>>(r.tmX, Tj.tmG, p);
    // (sharded with a communicator: the local B partial is kept without its rank-1 part, which the slab norm must include)
    const bool add_rank1 = trans && comm_ != nullptr && r.colsum_loc.p != nullptr;
    slab_sumsq<<<1, 256, 0, st>>>(p.C, 64, trans ? r.B.template as<float>() : r.A.template as<float>(), trans ? Ti.k : Tj.k, p.M, p.k,
                                  gate_slots_ + slot + 2 * side, add_rank1 ? r.colsum_loc.template as<float>() : nullptr,
                                  add_rank1 ? Ti.centre.template as<float>() : nullptr);
    launches += 2;
  }
}
template <>
void Engine<double>::gate_measure(RelRec&, int, cudaStream_t) {}

template <>
void Engine<double>::umma(const PlaneSet&, bool, const CUtensorMap&, int64_t, double*, int64_t, int, int, int, bool, cudaStream_t) {
  FZ_THROW(FZ_ERR_UNSUPPORTED, "tensor-core path needs the fp32 engine");
}

}  // namespace fz

// ================================================================================================
// C ABI
// ================================================================================================
struct fz_engine {
  fz::EngineBase* impl;
};

// Every entry point runs with the handle's device current and restores the caller's device afterwards: a host that
// drives several GPUs from one process (or torch code that moves the current device) cannot misdirect a launch.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

#define FZ_GUARD(e, body)                                        \
  if (!(e) || !(e)->impl) return FZ_ERR_INVALID;                 \
  DeviceGuard guard_((e)->impl->device);                         \
  try {                                                          \
    body;                                                        \
  } catch (const fz::FzError& err_) {                            \
    (e)->impl->err = err_.msg;                                   \
    return err_.status;                                          \
  } catch (const std::exception& ex_) {                          \
    (e)->impl->err = ex_.what();                                 \
    return FZ_ERR_INVALID;                                       \
  } catch (...) {                                                \
    (e)->impl->err = "unknown C++ exception";                    \
    return FZ_ERR_INVALID;                                       \
  }                                                              \
  return FZ_OK;

// ---- unknown-value replacement on a device-resident matrix (fusion_graph.py:464-510), in place
namespace {
template <class XT>
int fill_unknown_impl(XT* X, int64_t ld, int64_t rows, int64_t cols, int mode, double value, cudaStream_t st) {
  using namespace fz;
  if (mode == 3) {
    replace_unknown<XT><<<nblk(rows * cols, 256), 256, 0, st>>>(X, ld, rows, cols, 3, nullptr, nullptr, value);
    return cudaGetLastError() == cudaSuccess ? FZ_OK : FZ_ERR_CUDA;
  }
  const bool by_col = (mode == 2);
  const int64_t n_axis = by_col ? cols : rows;
  double *sum = nullptr, *cnt = nullptr, *total = nullptr, *psum = nullptr, *pcnt = nullptr;
  if (cudaMalloc(&sum, (size_t)n_axis * 8) != cudaSuccess || cudaMalloc(&cnt, (size_t)n_axis * 8) != cudaSuccess ||
      cudaMalloc(&total, 16) != cudaSuccess) {
    cudaFree(sum); cudaFree(cnt); cudaFree(total);
    return FZ_ERR_NOMEM;
  }
  int rc = FZ_OK;
  if (!by_col) {
    row_nan_stats<XT><<<nblk(rows, 8), 256, 0, st>>>(X, ld, rows, cols, sum, cnt);
  } else {
    const int chunks = (int)std::min<int64_t>(64, std::max<int64_t>(1, (rows + 255) / 256));
    const int64_t rpc = (rows + chunks - 1) / chunks;
    if (cudaMalloc(&psum, (size_t)chunks * cols * 8) != cudaSuccess || cudaMalloc(&pcnt, (size_t)chunks * cols * 8) != cudaSuccess) {
      rc = FZ_ERR_NOMEM;
    } else {
      dim3 g(nblk(cols, 256), chunks);
      col_nan_stats_partial<XT><<<g, 256, 0, st>>>(X, ld, rows, cols, rpc, psum, pcnt);
      sum_chunks<<<nblk(cols, 256), 256, 0, st>>>(psum, sum, chunks, cols);
      sum_chunks<<<nblk(cols, 256), 256, 0, st>>>(pcnt, cnt, chunks, cols);
    }
  }
  if (rc == FZ_OK) {
    finish_means<<<1, 1024, 0, st>>>(sum, cnt, n_axis, total);
    replace_unknown<XT><<<nblk(rows * cols, 256), 256, 0, st>>>(X, ld, rows, cols, mode, total, sum, 0.0);
    if (cudaGetLastError() != cudaSuccess) rc = FZ_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = FZ_ERR_CUDA;   // temporaries are freed below
  }
  cudaFree(sum); cudaFree(cnt); cudaFree(total); cudaFree(psum); cudaFree(pcnt);
  return rc;
}
}  // namespace


// ---- one process driving all the GPUs of the box: one host thread per handle for the duration of the call
namespace {
template <class F>
int for_each_engine(fz_engine** engines, int n, F&& fn) {
  if (!engines || n <= 0) return FZ_ERR_INVALID;
  for (int i = 0; i < n; ++i)
    if (!engines[i] || !engines[i]->impl) return FZ_ERR_INVALID;
  std::vector<int> rc((size_t)n, FZ_OK);
  std::vector<std::thread> threads;
  for (int i = 1; i < n; ++i) threads.emplace_back([&, i] { rc[(size_t)i] = fn(engines[i]); });
  rc[0] = fn(engines[0]);
  for (auto& t : threads) t.join();
  for (int i = 0; i < n; ++i)
    if (rc[(size_t)i] != FZ_OK) return rc[(size_t)i];
  return FZ_OK;
}
}  // namespace

extern "C" {

int fz_version(void) { return 100; }

int fz_create(fz_engine** out, int device, int compute) {
  if (!out) return FZ_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count <= 0) {
    fz::g_create_error = std::string("no CUDA device available: ") + cudaGetErrorString(ce) +
                         " (this engine has no CPU fallback)";
    return FZ_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    fz::g_create_error = "device index out of range";
    return FZ_ERR_INVALID;
  }
  DeviceGuard guard(device);     // the caller's current device is left as it was
  ce = cudaFree(nullptr);        // context of `device`
  if (ce != cudaSuccess) {
    fz::g_create_error = std::string("cannot use device ") + std::to_string(device) + ": " + cudaGetErrorString(ce);
    return FZ_ERR_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    fz::g_create_error = "this library is built for sm_100a (Blackwell B200) only; found compute capability " +
                         std::to_string(prop.major) + "." + std::to_string(prop.minor);
    return FZ_ERR_UNSUPPORTED;
  }
  fz::EngineBase* impl = nullptr;
  if (compute == FZ_F32) impl = new fz::Engine<float>(device);
  else if (compute == FZ_F64) impl = new fz::Engine<double>(device);
  else {
    fz::g_create_error = "compute dtype must be FZ_F32 or FZ_F64";
    return FZ_ERR_INVALID;
  }
  *out = new fz_engine{impl};
  return FZ_OK;
}

int fz_destroy(fz_engine* e) {
  if (!e) return FZ_OK;
  if (e->impl) {
    DeviceGuard guard(e->impl->device);
    delete e->impl;
    delete e;
    return FZ_OK;
  }
  delete e->impl;
  delete e;
  return FZ_OK;
}

const char* fz_last_error(const fz_engine* e) {
  if (!e || !e->impl) return fz::g_create_error.c_str();
  return e->impl->err.c_str();
}

int64_t fz_launch_count(const fz_engine* e) { return (e && e->impl) ? e->impl->launches : 0; }

int fz_set_shard(fz_engine* e, int world, int rank) { FZ_GUARD(e, e->impl->set_shard(world, rank)) }

int fz_comm_unique_id(void* out128) {
  if (!out128) return FZ_ERR_INVALID;
  const fz::NcclApi& nc = fz::nccl_api();
  if (!nc.ok) {
    fz::g_create_error = "NCCL unavailable: " + nc.error;
    return FZ_ERR_UNSUPPORTED;
  }
  ncclUniqueId id;
  if (nc.GetUniqueId(&id) != ncclSuccess) {
    fz::g_create_error = "ncclGetUniqueId failed";
    return FZ_ERR_CUDA;
  }
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, sizeof(id));
  return FZ_OK;
}
int fz_comm_init(fz_engine* e, const void* unique_id128) { FZ_GUARD(e, e->impl->comm_init(unique_id128)) }

int fz_group_comm_init(fz_engine** engines, int n) {
  unsigned char id[128];
  const int rc = fz_comm_unique_id(id);
  if (rc != FZ_OK) return rc;
  return for_each_engine(engines, n, [&](fz_engine* e) { return fz_comm_init(e, id); });
}
int fz_group_iterate(fz_engine** engines, int n, int algo, int n_iters) {
  return for_each_engine(engines, n, [&](fz_engine* e) {
    const int rc = fz_iterate(e, algo, n_iters, nullptr);
    if (rc != FZ_OK) return rc;
    DeviceGuard guard(e->impl->device);
    return cudaStreamSynchronize(nullptr) == cudaSuccess ? (int)FZ_OK : (int)FZ_ERR_CUDA;
  });
}
int fz_group_init_fill(fz_engine** engines, int n, int t, double value) {
  return for_each_engine(engines, n, [&](fz_engine* e) { return fz_init_fill(e, t, value, nullptr); });
}
int fz_group_relation_norms(fz_engine** engines, int n, int rel, int axis, double* dst_host, int64_t count) {
  // every rank computes (and takes part in the collective); rank 0 writes the caller's buffer, the others scratch copies
  if (count < 0) return FZ_ERR_INVALID;
  std::vector<std::vector<double>> scratch((size_t)std::max(0, n));
  for (int i = 1; i < n; ++i) scratch[(size_t)i].resize((size_t)count + 1);
  return for_each_engine(engines, n, [&](fz_engine* e) {
    int i = 0;
    while (i < n && engines[i] != e) ++i;
    return fz_relation_norms(e, rel, axis, i == 0 ? dst_host : scratch[(size_t)i].data(), nullptr);
  });
}
int fz_group_init_add_sampled_means(fz_engine** engines, int n, int t, int rel, const int32_t* idx_host, int p_c) {
  return for_each_engine(engines, n, [&](fz_engine* e) { return fz_init_add_sampled_means(e, t, rel, idx_host, p_c, nullptr); });
}
int fz_group_init_end(fz_engine** engines, int n) {
  return for_each_engine(engines, n, [&](fz_engine* e) { return fz_init_end(e); });
}
int fz_group_objective(fz_engine** engines, int n, double* per_relation, double* total) {
  // every rank computes its share and takes part in the all-reduce; rank 0's (identical) result is returned
  return for_each_engine(engines, n, [&](fz_engine* e) {
    const bool first = (e == engines[0]);
    return fz_objective(e, first ? per_relation : nullptr, first ? total : nullptr, nullptr);
  });
}

int fz_add_type(fz_engine* e, int64_t n, int k) {
  if (!e || !e->impl) return FZ_ERR_INVALID;
  DeviceGuard guard(e->impl->device);
  try {
    return e->impl->add_type(n, k);
  } catch (const fz::FzError& err_) {
    e->impl->err = err_.msg;
    return err_.status;
  } catch (const std::exception& ex_) {
    e->impl->err = ex_.what();
    return FZ_ERR_INVALID;
  } catch (...) {
    e->impl->err = "unknown C++ exception";
    return FZ_ERR_INVALID;
  }
}

int fz_add_relation(fz_engine* e, int ti, int tj, const void* data, int64_t ld, int src, int mem, int storage, int borrow,
                    const uint8_t* mask, int64_t mask_ld, int mask_mem) {
  if (!e || !e->impl) return FZ_ERR_INVALID;
  DeviceGuard guard(e->impl->device);
  try {
    return e->impl->add_relation(ti, tj, data, ld, src, mem, storage, borrow, mask, mask_ld, mask_mem);
  } catch (const fz::FzError& err_) {
    e->impl->err = err_.msg;
    return err_.status;
  } catch (const std::exception& ex_) {
    e->impl->err = ex_.what();
    return FZ_ERR_INVALID;
  } catch (...) {
    e->impl->err = "unknown C++ exception";
    return FZ_ERR_INVALID;
  }
}

int fz_set_factor(fz_engine* e, int t, const void* G0, int64_t ld, int src, int mem) { FZ_GUARD(e, e->impl->set_factor(t, G0, ld, src, mem)) }
int fz_set_backbone(fz_engine* e, int rel, const void* S, int64_t ld, int src, int mem) { FZ_GUARD(e, e->impl->set_backbone(rel, S, ld, src, mem)) }
int fz_set_split_terms(fz_engine* e, int terms) { FZ_GUARD(e, e->impl->set_split_terms(terms)) }
int fz_finalize(fz_engine* e) { FZ_GUARD(e, e->impl->finalize()) }
int fz_iterate(fz_engine* e, int algo, int n_iters, void* stream) { FZ_GUARD(e, e->impl->iterate(algo, n_iters, (cudaStream_t)stream)) }
int fz_phase_products(fz_engine* e, int algo, void* stream) { FZ_GUARD(e, e->impl->phase_products(algo, (cudaStream_t)stream)) }
int fz_phase_update(fz_engine* e, int algo, void* stream) { FZ_GUARD(e, e->impl->phase_update(algo, (cudaStream_t)stream)) }
int fz_phase_products_begin(fz_engine* e, int algo, void* stream) { FZ_GUARD(e, e->impl->phase_products_begin(algo, (cudaStream_t)stream)) }
int fz_phase_product_relation(fz_engine* e, int algo, int rel, void* stream) {
  FZ_GUARD(e, e->impl->phase_product_relation(algo, rel, (cudaStream_t)stream))
}
int fz_phase_products_end(fz_engine* e, int algo, void* stream) { FZ_GUARD(e, e->impl->phase_products_end(algo, (cudaStream_t)stream)) }
int fz_comm_small(fz_engine* e, void** ptr, int64_t* count) { FZ_GUARD(e, e->impl->comm_small(ptr, count)) }
int fz_comm_bpartial(fz_engine* e, int rel, void** full_ptr, void** local_ptr, int64_t* local_count, int* dtype) {
  FZ_GUARD(e, e->impl->comm_bpartial(rel, full_ptr, local_ptr, local_count, dtype))
}
int fz_comm_factor(fz_engine* e, int t, void** full_ptr, int64_t* local_count, int* dtype) {
  FZ_GUARD(e, e->impl->comm_factor(t, full_ptr, local_count, dtype))
}
int fz_transform_prepare(fz_engine* e, int target, void* stream) { FZ_GUARD(e, e->impl->transform_prepare(target, (cudaStream_t)stream)) }
int fz_transform_iterate(fz_engine* e, int n_iters, void* stream) { FZ_GUARD(e, e->impl->transform_iterate(n_iters, (cudaStream_t)stream)) }
int fz_get_factor(fz_engine* e, int t, void* dst, int64_t ld, int dst_dtype, int mem, void* stream) {
  FZ_GUARD(e, e->impl->get_factor(t, dst, ld, dst_dtype, mem, (cudaStream_t)stream))
}
int fz_get_backbone(fz_engine* e, int rel, void* dst, int64_t ld, int dst_dtype, int mem, void* stream) {
  FZ_GUARD(e, e->impl->get_backbone(rel, dst, ld, dst_dtype, mem, (cudaStream_t)stream))
}
int fz_objective(fz_engine* e, double* per_relation, double* total, void* stream) {
  FZ_GUARD(e, e->impl->objective(per_relation, total, (cudaStream_t)stream))
}
int fz_complete(fz_engine* e, int rel, void* dst, int64_t ld, int dst_dtype, int mem, void* stream) {
  FZ_GUARD(e, e->impl->complete(rel, dst, ld, dst_dtype, mem, (cudaStream_t)stream))
}
int fz_profile_product(fz_engine* e, int ti, int tj, const void* S, int64_t lds, int s_dtype, int s_mem, void* dst, int64_t ld,
                       int dst_dtype, int mem, void* stream) {
  FZ_GUARD(e, e->impl->profile_product(ti, tj, S, lds, s_dtype, s_mem, dst, ld, dst_dtype, mem, (cudaStream_t)stream))
}

int fz_init_fill(fz_engine* e, int t, double value, void* stream) {
  FZ_GUARD(e, e->impl->init_fill(t, value, (cudaStream_t)stream))
}
int fz_relation_norms(fz_engine* e, int rel, int axis, double* dst_host, void* stream) {
  FZ_GUARD(e, e->impl->relation_norms(rel, axis, dst_host, (cudaStream_t)stream))
}
int fz_init_add_sampled_means(fz_engine* e, int t, int rel, const int32_t* idx_host, int p_c, void* stream) {
  FZ_GUARD(e, e->impl->init_add_sampled_means(t, rel, idx_host, p_c, (cudaStream_t)stream))
}
int fz_init_end(fz_engine* e) {
  FZ_GUARD(e, e->impl->init_end())
}
int fz_pair_iterate(fz_engine* e0, fz_engine* e1, int n_iters, void* stream) {
  if (!e1 || !e1->impl) return FZ_ERR_INVALID;
  FZ_GUARD(e0, e0->impl->pair_iterate(e1->impl, n_iters, (cudaStream_t)stream))
}
int fz_relation_device_ptr(fz_engine* e, int rel, void** ptr, int64_t* ld, int* dtype) {
  FZ_GUARD(e, e->impl->relation_device_ptr(rel, ptr, ld, dtype))
}
int fz_operand_stats(fz_engine* e, int64_t* single_iters, int64_t* two_term_iters, int64_t* paired_iters, double* err_estimate,
                     double* cond_estimate) {
  FZ_GUARD(e, e->impl->operand_stats(single_iters, two_term_iters, paired_iters, err_estimate, cond_estimate))
}

int fz_profile(fz_engine* e, int enable) {
  if (!e || !e->impl) return FZ_ERR_INVALID;
  e->impl->profile = enable != 0;
  e->impl->prof_used = 0;
  e->impl->prof_bytes = 0.0;
  e->impl->prof_alg_bytes = 0.0;
  return FZ_OK;
}

int fz_profile_read(fz_engine* e, int64_t* launches, double* total_ms, double* streamed_bytes, double* algorithmic_bytes) {
  if (!e || !e->impl) return FZ_ERR_INVALID;
  double ms = 0.0;
  for (size_t i = 0; i < e->impl->prof_used; ++i) {
    if (cudaEventSynchronize(e->impl->prof_events[i].second) != cudaSuccess) return FZ_ERR_CUDA;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, e->impl->prof_events[i].first, e->impl->prof_events[i].second) != cudaSuccess) return FZ_ERR_CUDA;
    ms += t;
  }
  if (launches) *launches = (int64_t)e->impl->prof_used;
  if (total_ms) *total_ms = ms;
  if (streamed_bytes) *streamed_bytes = e->impl->prof_bytes;
  if (algorithmic_bytes) *algorithmic_bytes = e->impl->prof_alg_bytes;
  return FZ_OK;
}

int fz_fill_uniform(void* dst, int dtype, int64_t ld, int64_t rows, int64_t cols, int64_t row0, uint64_t seed, void* stream) {
  if (!dst || rows < 0 || cols < 0 || ld < cols) return FZ_ERR_INVALID;
  if (rows * cols == 0) return FZ_OK;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, dst) != cudaSuccess || attr.type != cudaMemoryTypeDevice) return FZ_ERR_INVALID;
  DeviceGuard guard(attr.device);                 // launch where the buffer lives, whatever the caller's current device
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = fz::nblk(rows * cols, 256);
  if (dtype == FZ_BF16) fz::fill_hashed_uniform<__nv_bfloat16><<<g, 256, 0, st>>>((__nv_bfloat16*)dst, ld, rows, cols, row0, seed);
  else if (dtype == FZ_F32) fz::fill_hashed_uniform<float><<<g, 256, 0, st>>>((float*)dst, ld, rows, cols, row0, seed);
  else if (dtype == FZ_F64) fz::fill_hashed_uniform<double><<<g, 256, 0, st>>>((double*)dst, ld, rows, cols, row0, seed);
  else return FZ_ERR_INVALID;
  return cudaGetLastError() == cudaSuccess ? FZ_OK : FZ_ERR_CUDA;
}

int fz_unknown_mask(const void* data, int dtype, int64_t ld, int64_t rows, int64_t cols, uint8_t* mask, int64_t mask_ld, void* stream) {
  if (!data || !mask || rows < 0 || cols < 0 || ld < cols || mask_ld < cols) return FZ_ERR_INVALID;
  if (rows * cols == 0) return FZ_OK;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, data) != cudaSuccess || attr.type != cudaMemoryTypeDevice) return FZ_ERR_INVALID;
  DeviceGuard guard(attr.device);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = fz::nblk(rows * cols, 256);
  if (dtype == FZ_BF16) fz::unknown_mask<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)data, ld, rows, cols, mask, mask_ld);
  else if (dtype == FZ_F32) fz::unknown_mask<float><<<g, 256, 0, st>>>((const float*)data, ld, rows, cols, mask, mask_ld);
  else if (dtype == FZ_F64) fz::unknown_mask<double><<<g, 256, 0, st>>>((const double*)data, ld, rows, cols, mask, mask_ld);
  else return FZ_ERR_INVALID;
  return cudaGetLastError() == cudaSuccess ? FZ_OK : FZ_ERR_CUDA;
}

int fz_fill_unknown(void* data, int dtype, int64_t ld, int64_t rows, int64_t cols, int mode, double value, void* stream) {
  if (!data || rows < 0 || cols < 0 || ld < cols || mode < 0 || mode > 3) return FZ_ERR_INVALID;
  if (rows * cols == 0) return FZ_OK;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, data) != cudaSuccess || attr.type != cudaMemoryTypeDevice) return FZ_ERR_INVALID;
  DeviceGuard guard(attr.device);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FZ_BF16) return fill_unknown_impl<__nv_bfloat16>((__nv_bfloat16*)data, ld, rows, cols, mode, value, st);
  if (dtype == FZ_F32) return fill_unknown_impl<float>((float*)data, ld, rows, cols, mode, value, st);
  if (dtype == FZ_F64) return fill_unknown_impl<double>((double*)data, ld, rows, cols, mode, value, st);
  return FZ_ERR_INVALID;
}

}  // extern "C"
