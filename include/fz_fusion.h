/* fz_fusion.h -- C ABI of the B200-native collective matrix tri-factorization engine.
 *
 * The reference (mims-harvard/scikit-fusion @ 88dd02c) has no FFI; its seam is three Python free
 * functions called by keyword from the estimator classes:
 *     dfmf(R, Theta, obj_types, obj_type2rank, ...)            skfusion/fusion/decomposition/_dfmf.py:127
 *     dfmc(R, M, Theta, obj_types, obj_type2rank, ...)         skfusion/fusion/decomposition/_dfmc.py:181
 *     transform(R_ij, Theta_i, target, ranks, G, S, ...)       skfusion/fusion/decomposition/_dfmf.py:330
 * This header is what a binding for that seam binds (ctypes stub in INTEGRATION.md).  One engine
 * handle == one call of one of those functions: describe the block structure (object types,
 * relation / constraint matrices, masks), hand over the initial factors, run N iterations of the
 * multiplicative-update loop on the GPU, read the factors G_t and backbones S_ij back.
 *
 * Conventions
 *   - every function returns 0 on success, a negative fz_status otherwise; fz_last_error() gives text.
 *     No exception crosses the ABI.  A handle is not thread-safe; distinct handles are independent.
 *   - matrices are row-major; `ld` is the leading dimension in ELEMENTS.
 *   - `mem` says where a caller buffer lives (FZ_HOST / FZ_DEVICE).  Inputs are copied (and converted
 *     to the storage dtype) unless `borrow` is set, in which case the device pointer must stay valid
 *     and unmodified for the lifetime of the handle (dfmc never borrows: it rewrites masked entries,
 *     _dfmc.py:268).  Outputs are written into caller-owned buffers.
 *   - all device work is enqueued on the `stream` argument (a cudaStream_t passed as void*; NULL =
 *     the legacy default stream).  Calls that return data to the HOST synchronise that stream.
 *   - every call makes the handle's device current for its duration and restores the caller's device.
 *   - row sharding (one process per GPU): fz_set_shard(world, rank) before any fz_add_type.  Type t
 *     then owns rows [rank*m_t, (rank+1)*m_t) with m_t = ceil(n_t/world); relation (i,j) is given as
 *     the row block of type i's local rows; the iteration is split into fz_phase_* calls and the
 *     caller runs the collectives between them on the buffers fz_comm_* exposes.
 */
#ifndef FZ_FUSION_H
#define FZ_FUSION_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fz_engine fz_engine;

typedef enum {
  FZ_F64 = 0, FZ_F32 = 1, FZ_BF16 = 2, FZ_U8 = 3,
  FZ_BF16X3 = 4   /* storage only: fp32 master + up to three bf16 planes whose sum is the fp32 value exactly */
} fz_dtype;
typedef enum { FZ_HOST = 0, FZ_DEVICE = 1 } fz_mem;
typedef enum { FZ_DFMF = 0, FZ_DFMC = 1 } fz_algo;

typedef enum {
  FZ_OK = 0,
  FZ_ERR_INVALID = -1,     /* bad argument / call order            */
  FZ_ERR_CUDA = -2,        /* CUDA runtime or driver error         */
  FZ_ERR_UNSUPPORTED = -3, /* combination not implemented          */
  FZ_ERR_NOMEM = -4
} fz_status;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* compute = FZ_F32 (fp32 factors + products, fp64 k x k chain) or FZ_F64 (all fp64, parity mode). */
int fz_create(fz_engine** out, int device, int compute);
int fz_destroy(fz_engine* e);
const char* fz_last_error(const fz_engine* e); /* e may be NULL: error of the last failed fz_create */
int fz_version(void);
/* number of engine kernels launched so far on this handle (bench.py's gpu_launches) */
int64_t fz_launch_count(const fz_engine* e);

/* ---- problem description (reference: the R / Theta / M dicts, dfmf.py:69-85, dfmc.py:69-94) ---- */
int fz_set_shard(fz_engine* e, int world, int rank);
/* Collectives inside the library (NCCL over NVLink, resolved at run time from libnccl.so.2).  Every rank of the shard group
 * calls fz_comm_init with the same 128-byte id from fz_comm_unique_id -- each from its own process (one process per GPU:
 * rank 0 makes the id, the host broadcasts it) or its own thread -- any time after fz_set_shard.  fz_iterate and
 * fz_objective on such a handle then run the whole sharded iteration, collectives included, and fz_phase_* / fz_comm_small
 * ... are not needed.  Replaces the reference's only parallelism, the joblib fan-out (_dfmf.py:69-73, dfmf.py:87-95). */
int fz_comm_unique_id(void* out128);
int fz_comm_init(fz_engine* e, const void* unique_id128);
/* One process driving several GPUs: the n handles form one shard group (rank i on handle i, each created on its own device).
 * The calls below run the per-handle call on one host thread per handle and return when all have finished. */
int fz_group_comm_init(fz_engine** engines, int n);
int fz_group_iterate(fz_engine** engines, int n, int algo, int n_iters);      /* fz_iterate on the NULL stream + synchronise */
int fz_group_objective(fz_engine** engines, int n, double* per_relation, double* total);
/* device-side initialisation (below) on a shard group driven from one process: the calls hold collectives, so each runs on
 * one host thread per handle.  `count` = entries of dst_host (the full vector: columns for axis 0, rows of the row type for 1). */
int fz_group_init_fill(fz_engine** engines, int n, int t, double value);
int fz_group_relation_norms(fz_engine** engines, int n, int rel, int axis, double* dst_host, int64_t count);
int fz_group_init_add_sampled_means(fz_engine** engines, int n, int t, int rel, const int32_t* idx_host, int p_c);
int fz_group_init_end(fz_engine** engines, int n);
/* returns the type id (>= 0).  n = number of objects (global), k = factorization rank. */
int fz_add_type(fz_engine* e, int64_t n, int k);
/* Relation between row type ti and column type tj; ti == tj declares a constraint matrix Theta_t.
 *   data        rows_local x n_tj matrix of `src` dtype (rows_local = n_ti unless sharded)
 *   storage     dtype kept on the device: FZ_F64 / FZ_F32 (exact CUDA-core path), FZ_BF16 (tensor-core path, the relation
 *               rounded to bf16; rank <= 64 for the fused kernels; needs ld % 8 == 0 when borrowed, the engine pads its own
 *               copies; constraint matrices and masked relations asked for in bf16 are kept exact in the compute dtype), or
 *               FZ_BF16X3 (fp32 engine): the relation is kept in fp32 AND split into the bf16 planes P0 + P1 + P2 = R
 *               (exactly; all-zero planes are dropped, so 0/1 data, ratings and small integers cost one plane).  The
 *               streamed products run on the tensor cores once per plane, accumulating into the same outputs: no bit of R
 *               is lost, the factor operand carries 16 bits (two bf16 terms).  Valid for masked relations (dfmc re-splits
 *               the imputed entries every iteration), constraint matrices (Theta+ and Theta- as separate plane sets) and
 *               any rank (ranks above 64 take the two-pass kernels over 64-column blocks of the factor)
 *   borrow      1: use the caller's buffer in place -- device memory, or (mem = FZ_HOST) PINNED host memory, which is then read
 *               over PCIe every iteration: the out-of-core mode for relations that do not fit HBM
 *   mask        optional rows_local x n_tj uint8 (non-zero = unknown entry, dfmc), NULL otherwise
 * returns the relation id (>= 0), in insertion order. */
int fz_add_relation(fz_engine* e, int ti, int tj, const void* data, int64_t ld, int src, int mem, int storage,
                    int borrow, const uint8_t* mask, int64_t mask_ld, int mask_mem);
/* Initial factor G_t (n_t x k_t, global rows).  Reference: initialize(), _init.py:6-61 (host side). */
int fz_set_factor(fz_engine* e, int t, const void* G0, int64_t ld, int src, int mem);
/* Frozen backbone S_ij for transform (k_ti x k_tj).  Reference: dfmf.py:112-114. */
int fz_set_backbone(fz_engine* e, int rel, const void* S, int64_t ld, int src, int mem);
/* Operand form of the factors on the tensor-core path (bf16-stored relations):
 *   1..3              plain form: G = G(0) + G(1) (+ G(2)), each term the bf16 rounding of the running residual (default 2)
 *   FZ_TERMS_AUTO     mean-centred form G = 1 c^T + D (c = column means, exact rank-1 algebra outside the MMAs); per
 *                     iteration the engine runs either the two-term kernel on D or the single-term kernel (half the tensor
 *                     work and half the B flush per relation byte) with the first-order effect of the dropped residual
 *                     restored in the fp64 backbone solve -- chosen from a measured estimate of the single-term error
 *                     (DESIGN.md section 4); unsharded and sharded dfmf
 *   FZ_TERMS_CENTRED1 always the single-term kernel (no accuracy gate: for studies) */
enum { FZ_TERMS_AUTO = 0, FZ_TERMS_CENTRED1 = -1 };
int fz_set_split_terms(fz_engine* e, int terms);
/* dfmf iterations run so far with the single-term / the two-term fused kernel, how many of them were batched with a partner
 * handle (fz_pair_iterate), and the last measured operand-form error and Gram condition estimate of the FZ_TERMS_AUTO gate
 * (-1 before the first check).  Any pointer may be NULL. */
int fz_operand_stats(fz_engine* e, int64_t* single_iters, int64_t* two_term_iters, int64_t* paired_iters, double* err_estimate,
                     double* cond_estimate);
/* allocate workspaces, build TMA descriptors; must precede the calls below */
int fz_finalize(fz_engine* e);

/* ---- the hot loop --------------------------------------------------------------------------- */
/* n_iters iterations of the multiplicative-update loop (_dfmf.py:212-296 / _dfmc.py:270-366).
 * Sharded handles need fz_comm_init (the engine then runs the collectives) or the phase calls below. */
int fz_iterate(fz_engine* e, int algo, int n_iters, void* stream);
/* Two restarts batched into one pass over the relations (reference: the n_run fan-out of Dfmf.fuse, dfmf.py:87-95).  e0 and
 * e1 describe the same graph on the same device with different initial factors; e1's bf16 relations must be BORROWED from
 * e0's device copies (fz_relation_device_ptr), both handles use a centred operand form (FZ_TERMS_AUTO / FZ_TERMS_CENTRED1).
 * Per iteration one launch per relation multiplies each relation tile with both runs' single-term operands; iterations in
 * which an accuracy gate measures or refuses the single-term form run one handle after the other, like two fz_iterate calls. */
int fz_pair_iterate(fz_engine* e0, fz_engine* e1, int n_iters, void* stream);
/* device pointer, leading dimension (elements) and dtype of the engine's copy of a relation (valid until fz_destroy) */
int fz_relation_device_ptr(fz_engine* e, int rel, void** ptr, int64_t* ld, int* dtype);
/* Sharded iteration, in order:  products -> [all-reduce small, reduce-scatter B] -> update ->
 * [all-gather factors].  fz_iterate == products + update when world == 1. */
int fz_phase_products(fz_engine* e, int algo, void* stream);
int fz_phase_update(fz_engine* e, int algo, void* stream);
/* fz_phase_products in pieces (dfmf): begin (Gram partials, operand forms), one call per relation (its A, B partial
 * and G_i^T A), end (constraint products).  Lets the caller overlap the reduce-scatter of relation r with the
 * streamed products of relation r+1. */
int fz_phase_products_begin(fz_engine* e, int algo, void* stream);
int fz_phase_product_relation(fz_engine* e, int algo, int rel, void* stream);
int fz_phase_products_end(fz_engine* e, int algo, void* stream);
/* communication buffers (device pointers, valid after fz_finalize) */
int fz_comm_small(fz_engine* e, void** ptr, int64_t* count_f64);               /* all-reduce, fp64 */
int fz_comm_bpartial(fz_engine* e, int rel, void** full_ptr, void** local_ptr, /* reduce-scatter  */
                     int64_t* local_count, int* dtype);
int fz_comm_factor(fz_engine* e, int t, void** full_ptr, int64_t* local_count, int* dtype); /* all-gather */

/* online projection (_dfmf.py:330-458): only `target` moves; other factors and all backbones frozen.
 * fz_transform_prepare computes the loop-invariant terms once; fz_transform_iterate runs the loop. */
int fz_transform_prepare(fz_engine* e, int target, void* stream);
int fz_transform_iterate(fz_engine* e, int n_iters, void* stream);

/* ---- results --------------------------------------------------------------------------------- */
int fz_get_factor(fz_engine* e, int t, void* dst, int64_t ld, int dst_dtype, int mem, void* stream);
int fz_get_backbone(fz_engine* e, int rel, void* dst, int64_t ld, int dst_dtype, int mem, void* stream);
/* Frobenius residuals ||R - G_i S G_j^T||_F per relation (un-squared, _dfmf.py:306-319) and their sum,
 * with the current factors and the backbones of the last iteration.  per_relation may be NULL. */
int fz_objective(fz_engine* e, double* per_relation, double* total, void* stream);
/* completed relation G_i S_ij G_j^T (skfusion/fusion/base/base.py:119-146) into a caller buffer (n_i x n_j) */
int fz_complete(fz_engine* e, int rel, void* dst, int64_t ld, int dst_dtype, int mem, void* stream);
/* G_ti M G_tj^T for a caller-given k_ti x k_tj matrix M (host or device): the chained profiles of the reference's examples,
 * M = S_ab S_bc ... along a path of the fusion graph (base.py:69-96, examples/dicty_chaining.py:40-53), or a completion with
 * any backbone.  fp32 engine with ranks <= 64: an output-bound tcgen05 product written by TMA stores (a device fp32
 * destination with ld % 4 == 0 is written in place); otherwise an exact CUDA-core kernel.  Needs the two factors only: a
 * handle with types, factors and no relations is enough. */
int fz_profile_product(fz_engine* e, int ti, int tj, const void* M, int64_t ldm, int m_dtype, int m_mem, void* dst, int64_t ld,
                       int dst_dtype, int mem, void* stream);

/* ---- factor initialisation on the device ------------------------------------------------------
 * Reference: initialize() / _random_c / _random_vcol, skfusion/fusion/decomposition/_init.py:6-61.  The host keeps the
 * RandomState (the draws must be consumed bit-exactly) and sends, per object type and per relation touching it, the k_t
 * lists of p_c = int(0.2 * cols) sampled column indices of the relation oriented with the type on its rows; the engine
 * computes the column means (a product with a 0/1 selection matrix through the streamed kernels) and accumulates
 *     G_t = value + sum_relations | mean of the sampled columns |                      (_init.py:36-39, 57-60).
 * Call order: fz_finalize, then per type fz_init_fill + fz_init_add_sampled_means per relation, then fz_init_end (which
 * marks every factor as set).  p_c == 0 (fewer than 5 columns) yields NaN factors, as upstream.  Sharded handles need their
 * communicator (fz_comm_init): every rank makes the same calls with the same plans; column norms are summed and the sampled
 * means all-gathered / all-reduced, so every rank ends with the same full factors.
 * fz_relation_norms returns the 2-norms of the columns (axis 0) or rows (axis 1) of a relation in a HOST buffer: random_c
 * samples from the int(0.5 * cols) columns of largest norm (_init.py:30-34). */
int fz_init_fill(fz_engine* e, int t, double value, void* stream);
int fz_relation_norms(fz_engine* e, int rel, int axis, double* dst_host, void* stream);
int fz_init_add_sampled_means(fz_engine* e, int t, int rel, const int32_t* idx_host, int p_c, void* stream);
int fz_init_end(fz_engine* e);

/* ---- measurement ----------------------------------------------------------------------------- */
/* fz_profile(e, 1) brackets every streamed tensor-core product with CUDA events on its launch stream;
 * fz_profile_read returns how many launches were timed, the sum of their durations, the relation bytes
 * they streamed (rows x cols x 2 per launch) and their ALGORITHMIC bytes: one pass over a relation
 * yields both of its products, so a single-product launch is credited with half of what it streams and
 * a fused launch with all of it.  Used by bench.py's roofline leg. */
int fz_profile(fz_engine* e, int enable);
int fz_profile_read(fz_engine* e, int64_t* launches, double* total_ms, double* streamed_bytes, double* algorithmic_bytes);

/* ---- synthetic workloads ----------------------------------------------------- */
/* Fill a rows x cols device matrix (leading dimension ld) with the counter-based uniform [0,1) values
 * value(r, c) = splitmix64(seed, (row0 + r) * cols + c) >> 40 / 2^24, rounded to `dtype`.  The same
 * numbers come out of oracle/fusion_oracle.py:hashed_uniform, on any row sharding. */
int fz_fill_uniform(void* dst, int dtype, int64_t ld, int64_t rows, int64_t cols, int64_t row0, uint64_t seed, void* stream);

/* ---- relation preprocessing on the device -----------------------------------
 * Replace the unknown (non-finite) entries of a DEVICE-resident rows x cols matrix in place, as Relation.filled() does on
 * the host (fill_mean / fill_row / fill_col / fill_const, skfusion/fusion/base/fusion_graph.py:464-510):
 *   mode 0  <- mean of the matrix          1  <- mean of the entry's row        2  <- mean of its column       3  <- value
 * Means follow numpy.nanmean (NaN skipped, +-inf included), summed in fp64; a row / column without any known entry takes
 * the matrix mean.  Masked arrays stay a host (numpy) feature. */
int fz_fill_unknown(void* data, int dtype, int64_t ld, int64_t rows, int64_t cols, int mode, double value, void* stream);
/* mask[r][c] = 1 where the DEVICE-resident matrix holds a non-finite entry: the completion mask Dfmc takes from numpy masked
 * arrays on the host (skfusion/fusion/decomposition/dfmc.py:69-94), extracted without leaving the GPU (before the fill). */
int fz_unknown_mask(const void* data, int dtype, int64_t ld, int64_t rows, int64_t cols, uint8_t* mask, int64_t mask_ld, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FZ_FUSION_H */
