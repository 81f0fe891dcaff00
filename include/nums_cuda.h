/*
 * nums_cuda.h -- C ABI of libnumscuda.so, the sm_100a (B200) kernel library behind
 * `nums_b200.cuda_compute.ComputeCls`.
 *
 * The reference (merrymercy/nums) is pure Python: its per-block kernel surface is the
 * 28-method `ComputeInterface` (nums/core/systems/interfaces.py:73-167) implemented by
 * `numpy_compute.ComputeCls` (nums/core/systems/numpy_compute.py:84-286) on top of
 * NumPy/SciPy.  Every entry point below replaces the NumPy/SciPy call made at the cited
 * line; the Python class that binds them (ctypes) keeps the reference's method names and
 * parameters.  No torch types appear here: plain device pointers, sizes, a CUDA stream.
 *
 * Conventions
 *   - all array pointers are DEVICE pointers on the current CUDA device, unless the
 *     parameter name ends in `_host`;
 *   - arrays are described by `nums_array_t` (dtype, ndim <= 8, shape, strides in ELEMENTS;
 *     stride 0 = broadcast; negative strides are not used);
 *   - every function returns 0 on success and a negative `nums_status_t` otherwise;
 *     `nums_last_error()` returns a thread-local message for the last failure;
 *   - every function only enqueues work on `stream` (a `cudaStream_t` passed as void*) and
 *     never synchronises, unless documented ("host result");
 *   - `ws` / `ws_bytes` is caller-owned scratch in device memory; if it is too small the call
 *     fails with NUMS_ERR_WORKSPACE and `nums_last_workspace_request()` tells how much is
 *     needed.
 */
#ifndef NUMS_CUDA_H_
#define NUMS_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NUMS_MAX_DIMS 8
#define NUMS_ABI_VERSION 1

typedef enum {
  NUMS_OK = 0,
  NUMS_ERR_INVALID = -1,     /* bad argument (shape / dtype / op mismatch)             */
  NUMS_ERR_UNSUPPORTED = -2, /* valid NumPy semantics this library does not implement   */
  NUMS_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed                     */
  NUMS_ERR_WORKSPACE = -4,   /* workspace too small, see nums_last_workspace_request()  */
  NUMS_ERR_NUMERIC = -5      /* singular / not positive definite (host-result calls)    */
} nums_status_t;

/* Storage dtypes (NumPy names: bool_, int32, int64, float32, float64). */
typedef enum {
  NUMS_BOOL = 0,
  NUMS_I32 = 1,
  NUMS_I64 = 2,
  NUMS_F32 = 3,
  NUMS_F64 = 4
} nums_dtype_t;

typedef struct {
  void* data;
  int32_t dtype; /* nums_dtype_t */
  int32_t ndim;
  int64_t shape[NUMS_MAX_DIMS];
  int64_t stride[NUMS_MAX_DIMS]; /* in elements */
} nums_array_t;

/* Binary ufuncs reachable through ComputeCls.bop (numpy_compute.py:233-238, names after
 * nums/core/settings.py:48-61) and ComputeCls.xlogy (numpy_compute.py:203-204). */
typedef enum {
  NUMS_BOP_ADD = 0, NUMS_BOP_SUBTRACT, NUMS_BOP_MULTIPLY, NUMS_BOP_TRUE_DIVIDE,
  NUMS_BOP_FLOOR_DIVIDE, NUMS_BOP_REMAINDER, NUMS_BOP_FMOD, NUMS_BOP_POWER,
  NUMS_BOP_FLOAT_POWER, NUMS_BOP_MAXIMUM, NUMS_BOP_MINIMUM, NUMS_BOP_FMAX, NUMS_BOP_FMIN,
  NUMS_BOP_ARCTAN2, NUMS_BOP_HYPOT, NUMS_BOP_COPYSIGN, NUMS_BOP_NEXTAFTER,
  NUMS_BOP_HEAVISIDE, NUMS_BOP_LOGADDEXP, NUMS_BOP_LOGADDEXP2, NUMS_BOP_LDEXP,
  NUMS_BOP_XLOGY,
  NUMS_BOP_LESS, NUMS_BOP_LESS_EQUAL, NUMS_BOP_GREATER, NUMS_BOP_GREATER_EQUAL,
  NUMS_BOP_EQUAL, NUMS_BOP_NOT_EQUAL,
  NUMS_BOP_LOGICAL_AND, NUMS_BOP_LOGICAL_OR, NUMS_BOP_LOGICAL_XOR,
  NUMS_BOP_BITWISE_AND, NUMS_BOP_BITWISE_OR, NUMS_BOP_BITWISE_XOR,
  NUMS_BOP_LEFT_SHIFT, NUMS_BOP_RIGHT_SHIFT, NUMS_BOP_GCD, NUMS_BOP_LCM,
  NUMS_BOP_COUNT_
} nums_bop_t;

/* Unary ufuncs reachable through ComputeCls.map_uop (numpy_compute.py:184-186). */
typedef enum {
  NUMS_UOP_COPY = 0, /* identity with dtype conversion: astype / strided copies */
  NUMS_UOP_ABS, NUMS_UOP_NEGATIVE, NUMS_UOP_POSITIVE, NUMS_UOP_SIGN, NUMS_UOP_SQRT,
  NUMS_UOP_CBRT, NUMS_UOP_SQUARE, NUMS_UOP_RECIPROCAL, NUMS_UOP_EXP, NUMS_UOP_EXP2,
  NUMS_UOP_EXPM1, NUMS_UOP_LOG, NUMS_UOP_LOG2, NUMS_UOP_LOG10, NUMS_UOP_LOG1P,
  NUMS_UOP_SIN, NUMS_UOP_COS, NUMS_UOP_TAN, NUMS_UOP_ARCSIN, NUMS_UOP_ARCCOS,
  NUMS_UOP_ARCTAN, NUMS_UOP_SINH, NUMS_UOP_COSH, NUMS_UOP_TANH, NUMS_UOP_ARCSINH,
  NUMS_UOP_ARCCOSH, NUMS_UOP_ARCTANH, NUMS_UOP_FLOOR, NUMS_UOP_CEIL, NUMS_UOP_TRUNC,
  NUMS_UOP_RINT, NUMS_UOP_DEG2RAD, NUMS_UOP_RAD2DEG, NUMS_UOP_SPACING,
  NUMS_UOP_ISNAN, NUMS_UOP_ISINF, NUMS_UOP_ISFINITE, NUMS_UOP_SIGNBIT,
  NUMS_UOP_LOGICAL_NOT, NUMS_UOP_INVERT,
  NUMS_UOP_COUNT_
} nums_uop_t;

/* Block reductions reachable through ComputeCls.reduce_axis (numpy_compute.py:177-181). */
typedef enum {
  NUMS_RED_SUM = 0, NUMS_RED_PROD, NUMS_RED_MIN, NUMS_RED_MAX, NUMS_RED_ANY, NUMS_RED_ALL,
  NUMS_RED_COUNT_
} nums_reduce_t;

/* ---- library / diagnostics --------------------------------------------------------- */
int nums_abi_version(void);
const char* nums_last_error(void);
size_t nums_last_workspace_request(void);
/* Number of SMs of the current device (148 on B200); <0 on error. */
int nums_sm_count(void);
/* Total number of CUDA kernels this library has launched in the process so far. */
uint64_t nums_launch_count(void);

/* ---- elementwise ---------------------------------------------------------------------
 * out = ufunc(a, b) with NumPy broadcasting of a and b against out->shape (right aligned,
 * size-1 axes broadcast).  `loop_dtype` is the dtype NumPy's type resolution runs the inner
 * loop in (np.<ufunc>.resolve_dtypes); operands stored in another dtype are converted on
 * load.  out->dtype must be the loop's output dtype.  Replaces np.<ufunc>(a1, a2)
 * (numpy_compute.py:233-238) and scipy.special.xlogy (:203-204). */
int nums_bop(int op, int loop_dtype, const nums_array_t* a, const nums_array_t* b,
             const nums_array_t* out, void* stream);

/* nums_bop for operands that are dense over the output's iteration space or single values (a_n / b_n:
 * n = dense, 1 = broadcast scalar): raw pointers instead of array descriptors, for the per-block
 * dispatch path of BlockArray's elementwise operators (base.py:167-246 -> numpy_compute.py:233-238). */
int nums_bop_flat(int op, int loop_dtype, const void* a, int a_dtype, int64_t a_n, const void* b,
                  int b_dtype, int64_t b_n, void* out, int out_dtype, int64_t n, void* stream);

/* out = ufunc(a) (numpy_compute.py:184-186).  NUMS_UOP_COPY is a dtype-converting strided
 * copy: astype (:206-208), the slice copies of create_block/update_block (:119-169), and
 * materialisation of transposed views (:213-214, :222-229). */
int nums_uop(int op, int loop_dtype, const nums_array_t* a, const nums_array_t* out,
             void* stream);

/* Batched element / slice scatter: for every pair p, dst[o][dst_index[p]][i] = src[o][src_index[p]][i]
 * over dense (outer, dst_len, inner) / (outer, src_len, inner) views of elem_size-byte elements
 * (1, 4 or 8).  One launch for the per-pair loops of update_block_by_index (flattened arrays,
 * outer = inner = 1) and update_block_along_axis (numpy_compute.py:154-169).  Indices are device
 * int64, already normalised and free of duplicate destinations (the caller keeps the last pair,
 * which is what the reference's sequential loop leaves behind). */
int nums_scatter_axis(int elem_size, int64_t outer, int64_t dst_len, int64_t src_len, int64_t inner,
                      int64_t npairs, const int64_t* dst_index, const int64_t* src_index, void* dst,
                      const void* src, void* stream);

/* out = sum of n same-shape contiguous arrays, left to right (np.add.reduce(arrs),
 * numpy_compute.py:210-211).  `arrs_host` is a HOST array of n device pointers. */
int nums_sum_reduce(int n, const void* const* arrs_host, int dtype, int64_t numel, void* out,
                    void* stream);

/* Fill / arange / eye: np.zeros/ones (:96-104), np.arange (:174-175), np.eye (:100-102). */
int nums_fill(const nums_array_t* out, double value, void* stream);
int nums_arange(const nums_array_t* out, double start, double step, void* stream);
int nums_eye(const nums_array_t* out, void* stream);

/* ---- reductions ------------------------------------------------------------------------
 * out = np.<op>(a, axis) for a C-contiguous `a` viewed as (outer, reduce, inner);
 * axis == -1 reduces everything (axis=None).  out is contiguous with outer*inner elements.
 * Sums of bool/int accumulate in int64, float32 sums accumulate in float64.  min/max
 * propagate NaN like np.min/np.max. (numpy_compute.py:177-181) */
int nums_reduce(int op, const void* a, int a_dtype, int64_t outer, int64_t reduce,
                int64_t inner, void* out, int out_dtype, void* ws, size_t ws_bytes,
                void* stream);

/* 1-D argmin (is_max=0) / argmax (is_max=1) with first-occurrence ties, merged with an
 * optional carried optimum that wins only if STRICTLY better (numpy_compute.py:269-283).
 * out_index (int64, device) = index_offset + local index, out_value (a_dtype, device).
 * carried_index / carried_value may be NULL. */
int nums_arg_op(int is_max, const void* a, int a_dtype, int64_t n, int64_t index_offset,
                const int64_t* carried_index, const void* carried_value, int64_t* out_index,
                void* out_value, void* ws, size_t ws_bytes, void* stream);

/* np.allclose(a, b, rtol, atol) on contiguous arrays of the same dtype; writes 0/1 into the
 * device byte *out_flag (numpy_compute.py:261-262). */
int nums_allclose(const void* a, const void* b, int dtype, int64_t numel, double rtol,
                  double atol, uint8_t* out_flag, void* ws, size_t ws_bytes, void* stream);

/* np.where(arr) for a C-contiguous array (numpy_compute.py:188-194).
 * Step 1: count the non-zeros (device int64 *count).  Step 2, after the caller has read the
 * count and allocated `ndim` int64 outputs of that length: write the coordinates, in C order,
 * each axis offset by offsets_host[axis]. */
int nums_nonzero_count(const void* a, int dtype, int64_t numel, int64_t* count, void* ws,
                       size_t ws_bytes, void* stream);
int nums_nonzero_fill(const void* a, int dtype, int ndim, const int64_t* shape_host,
                      const int64_t* offsets_host, int64_t* const* outs_host, void* ws,
                      size_t ws_bytes, void* stream);

/* ---- dense contractions ----------------------------------------------------------------
 * C[m,n] (+)= op(A)[m,k] . op(B)[k,n], all row-major; lda/ldb/ldc are row pitches in
 * elements of the STORED matrices (A stored k x m when trans_a).  accumulate != 0 adds into
 * C.  dtype in {F64, F32, I64, I32}.  F64 runs on the FP64 tensor pipe (DMMA).
 * Replaces np.tensordot(a1, a2, axes) (numpy_compute.py:231-232). */
int nums_gemm(int dtype, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k,
              const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
              int accumulate, void* ws, size_t ws_bytes, void* stream);

/* Grouped / chained contraction: for every problem p,
 *     C_p[m,n] = (Cin_p ? Cin_p : 0) + sum over its terms t of op(A_t)[m,k_t] . op(B_t)[k_t,n]
 * in ONE launch whose CTAs walk the tiles of all problems.  This is the k-chain of
 * BlockArray._tensordot (blockarray.py:460-472: `dot`, then `result_block += dot` per k)
 * accumulated in registers instead of through separate `add` kernels, for all result blocks at
 * once.  float64 only; every operand must be 16-byte aligned with an even pitch.  The tables are
 * HOST arrays (copied to the workspace by the call).  trans_a / trans_b apply to all terms. */
typedef struct {
  const void* A;
  const void* B;
  int64_t lda, ldb, k;
} nums_gemm_term_t;
typedef struct {
  void* C;
  const void* Cin;
  int64_t ldc, ldcin, m, n;
  int32_t term_begin, term_count;
} nums_gemm_problem_t;
int nums_gemm_grouped(int dtype, int trans_a, int trans_b, int nproblems,
                      const nums_gemm_problem_t* problems_host, int nterms,
                      const nums_gemm_term_t* terms_host, void* ws, size_t ws_bytes, void* stream);

/* ---- factorisations ----------------------------------------------------------------------
 * QR of a row-major m x n matrix (np.linalg.qr, numpy_compute.py:240-246).
 * R is k x n (k = min(m,n)), row-major, written with zeros below the diagonal.
 * Q (m x k, row-major) may be NULL for mode "r". */
int nums_qr(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, void* Q, int64_t ldq,
            void* R, int64_t ldr, void* ws, size_t ws_bytes, void* stream);

/* General inverse by LU with partial pivoting (np.linalg.inv, numpy_compute.py:256-257);
 * *info (device int32): 0 ok, j>0 = exact zero pivot at step j. */
int nums_inv(int dtype, int64_t n, const void* A, int64_t lda, void* Ainv, int64_t ldi,
             int32_t* info, void* ws, size_t ws_bytes, void* stream);

/* Lower Cholesky factor (np.linalg.cholesky, numpy_compute.py:248-249); *info as above
 * (j>0 = leading minor j not positive definite). */
int nums_cholesky(int dtype, int64_t n, const void* A, int64_t lda, void* L, int64_t ldl,
                  int32_t* info, void* ws, size_t ws_bytes, void* stream);

/* Newton update of glms.newton (glms.py:362-372: beta <- beta - inv(H) g, stop on max |g| <= tol) in one
 * launch: gh = d entries of g followed by d*d entries of H (the layout nums_lr_grad_hess writes, after
 * the cross-block / cross-rank sum); beta_out = beta - H^-1 g by Gauss-Jordan with partial pivoting;
 * status (device, f64[2]) = {max |g|, info} with info > 0 = zero pivot at that step (np.linalg.inv
 * would raise LinAlgError).  d <= 128. */
int nums_newton_step(int64_t d, const double* gh, const double* beta, double* beta_out, double* status,
                     void* stream);

/* ---- delimited text ingest ------------------------------------------------
 * Device-side replacement of read_csv_block (nums/core/systems/filesystem.py:157-212): the bytes
 * [first, stop) of `text` (device memory, 16-byte aligned, readable up to `stop` rounded up to a
 * multiple of 32) are the whole lines of one chunk, chosen by the caller with the reference's
 * rule (:196-211).  nums_csv_index counts them; nums_csv_parse converts every field with the
 * dtype's converter (:160-190: float(x), np.int64(x), int(float(x)), bool(int(x))) into the
 * row-major rows x cols block `out`.
 * summary (device int64[4]) = {lines, fields, status, byte offset of the first offending field};
 * status is NUMS_CSV_OK, _INVALID (the reference raises ValueError), _UNSUPPORTED (valid for the
 * reference, not handled: hex float, lone '\r', undecidable literal of > 19 digits) or _RAGGED
 * (rows of different lengths; np.array raises ValueError).  nums_csv_index leaves per-tile offsets
 * in `ws`, which nums_csv_parse reads: same workspace, same stream, no launch in between. */
enum { NUMS_CSV_OK = 0, NUMS_CSV_INVALID = 1, NUMS_CSV_UNSUPPORTED = 2, NUMS_CSV_RAGGED = 3 };
int nums_csv_index(const void* text, int64_t first, int64_t stop, int delimiter, int64_t* summary,
                   void* ws, size_t ws_bytes, void* stream);
int nums_csv_parse(const void* text, int64_t first, int64_t stop, int delimiter, int dtype, int64_t rows,
                   int64_t cols, void* out, int64_t* summary, const void* ws, void* stream);

/* Small-matrix stage of the Gram path of qr (np.linalg.qr, numpy_compute.py:240-246, for tall
 * float64 blocks): from G = A^T A (n <= 128, lower triangle read) one launch writes L = chol(G),
 * R = L^T (the qr mode='r' result), L^-1 (any of the three may be NULL) and
 * stats[5] = {info, |L|_1, |L|_inf, |L^-1|_1, |L^-1|_inf} (device doubles; info > 0: G not
 * numerically positive definite at that step, nothing else written).  The caller accepts R when
 * sqrt(|L|_1 |L|_inf |L^-1|_1 |L^-1|_inf) >= cond_2(A) is small and otherwise refines or falls
 * back to nums_qr. */
int nums_gram_factor(int64_t n, const void* G, int64_t ldg, void* L, int64_t ldl, void* R, int64_t ldr,
                     void* Linv, int64_t ldi, double* stats, void* stream);

/* Full SVD of a square n x n matrix by one-sided Jacobi (np.linalg.svd, :251-254):
 * A = U diag(S) Vt, S descending. */
int nums_svd(int dtype, int64_t n, const void* A, int64_t lda, void* U, void* S, void* Vt,
             void* ws, size_t ws_bytes, void* stream);

/* ---- fused logistic-regression step -------------------------------
 * One pass over a row block X (n x d, row-major, f64), y (n): mu = sigmoid(X beta),
 * g = X^T (mu - y), H = X^T diag(mu (1 - mu)) X.  Fuses the call chain of
 * nums/models/glms.py:140-143,213-240 (forward / gradient / hessian).
 * out (device, f64): d entries of g followed by d*d entries of H (row-major). */
int nums_lr_grad_hess(int64_t n, int64_t d, const double* X, int64_t ldx, const double* y,
                      const double* beta, double* out, void* ws, size_t ws_bytes, void* stream);

/* Same, for up to 16 dense row blocks (ldx == d) of one row-sharded matrix in a single launch: the G
 * row blocks of X / y that glms.newton walks (glms.py:362-372).  X_host / y_host / rows_host are
 * HOST arrays of device pointers / row counts.  Needs d == 4 or 12 (mod 16), d <= 48 (the shapes
 * served by the bulk-copy kernel, e.g. HIGGS d = 28); otherwise NUMS_ERR_UNSUPPORTED. */
int nums_lr_grad_hess_blocks(int nblocks, const double* const* X_host, const double* const* y_host,
                             const int64_t* rows_host, int64_t d, const double* beta, double* out,
                             void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NUMS_CUDA_H_ */
