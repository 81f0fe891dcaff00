// Dense block contractions -- replaces np.tensordot(a1, a2, axes) at
// nums/core/systems/numpy_compute.py:231-232 (OpenBLAS dgemm/dgemv in the reference).
//
//   * f64 GEMM: FP64 tensor pipe.  On sm_100a every mma.sync f64 shape lowers to DMMA.8x8x4
//     (checked with cuobjdump), so the kernel is written directly against m8n8k4 fragments:
//     128x128x16 CTA tiles, 8 warps of 32x64, a 4-stage cp.async shared-memory ring padded so
//     that every fragment read is bank-conflict free, optional split-K for small outputs with
//     long contractions (X^T X, the LR Hessian).  Operands may be stored transposed
//     (BlockArray.T is lazy, base.py:72-85), handled by the shared-memory layout, not by a
//     copy.
//   * matrix-vector / vector-vector forms (BlockArray._matvec / _vecdot, blockarray.py:475-580):
//     HBM-bound streaming kernels.
//   * everything else (f32, exact integer tensordot from tests/core/array/test_bop.py:38-42,
//     unaligned f64): a plain shared-memory tiled kernel.
#include "common.cuh"

namespace nums {
namespace {

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ======================================================================================
// FP64 DMMA GEMM
// ======================================================================================
constexpr int BM = 128, BN = 128, BK = 16;
constexpr int kStages = 4;
constexpr int kGemmThreads = 256;
constexpr int kPad = 4;  // row pitch == 4 (mod 16) doubles => conflict-free 8x4 / 4x8 fragment reads

// Shared-memory pitches (in doubles) per operand layout.
//   A stored (m,k) ["N"]: tile BM x BK, pitch BK + 4      A stored (k,m) ["T"]: tile BK x BM, pitch BM + 4
//   B stored (k,n) ["N"]: tile BK x BN, pitch BN + 4      B stored (n,k) ["T"]: tile BN x BK, pitch BK + 4
template <bool TA> struct ATile {
  static constexpr int rows = TA ? BK : BM, cols = TA ? BM : BK, pitch = cols + kPad;
  static constexpr int doubles = rows * pitch;
};
template <bool TB> struct BTile {
  static constexpr int rows = TB ? BN : BK, cols = TB ? BK : BN, pitch = cols + kPad;
  static constexpr int doubles = rows * pitch;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// Copy a (rows x cols) tile of a row-major matrix (pitch ld, extent R x C) starting at
// (r0, c0) into shared memory, 16 bytes (2 doubles) per cp.async, zero-filling out of range.
// Requires ld even and a 16-byte aligned base so every in-range chunk is aligned.
template <int ROWS, int COLS, int PITCH>
__device__ __forceinline__ void load_tile(double* smem, const double* __restrict__ g, int64_t ld,
                                          int64_t R, int64_t C, int64_t r0, int64_t c0) {
  constexpr int kChunksPerRow = COLS / 2;
  constexpr int kChunks = ROWS * kChunksPerRow;
  static_assert(kChunks % kGemmThreads == 0, "tile must divide evenly over the CTA");
#pragma unroll
  for (int it = 0; it < kChunks / kGemmThreads; ++it) {
    const int chunk = it * kGemmThreads + threadIdx.x;
    const int r = chunk / kChunksPerRow, c = (chunk % kChunksPerRow) * 2;
    const int64_t gr = r0 + r, gc = c0 + c;
    int bytes = 0;
    if (gr < R && gc < C) bytes = (gc + 1 < C) ? 16 : 8;
    // keep the address in range even when nothing is read
    const double* src = bytes ? g + gr * ld + gc : g;
    cp_async16(smem + r * PITCH + c, src, bytes);
  }
}

struct GemmParams {
  const double* A;
  const double* B;
  const double* Cin;  // optional addend (may alias C)
  double* C;
  int64_t M, N, K;
  int64_t lda, ldb, ldc, ldcin;
  int64_t k_per_split;   // multiple of BK
  int64_t split_stride;  // elements between split-K partial outputs (M*N) or 0
  int tiles_m, tiles_n;
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(kGemmThreads, 1)
dgemm_dmma_kernel(GemmParams p) {
  extern __shared__ __align__(16) double smem[];
  using AT = ATile<TA>;
  using BT = BTile<TB>;
  double* sA = smem;
  double* sB = smem + kStages * AT::doubles;

  // Tile order: groups of 8 tile-rows sweep the columns, so concurrently running CTAs share
  // A row panels and B column panels in L2.
  int tile = blockIdx.x;
  constexpr int kGroup = 8;
  const int tiles_per_group = kGroup * p.tiles_n;
  const int group = tile / tiles_per_group;
  const int first_m = group * kGroup;
  const int group_rows = min(p.tiles_m - first_m, kGroup);
  const int tm = first_m + (tile % tiles_per_group) % group_rows;
  const int tn = (tile % tiles_per_group) / group_rows;
  const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;

  const int64_t k_begin = (int64_t)blockIdx.z * p.k_per_split;
  int64_t k_end = k_begin + p.k_per_split;
  if (k_end > p.K) k_end = p.K;
  const int KT = (int)((k_end - k_begin + BK - 1) / BK);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp & 3) * 32;   // 4 warps along M
  const int wn = (warp >> 2) * 64;  // 2 warps along N
  const int g = lane >> 2, t = lane & 3;

  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load_stage = [&](int slot, int kt) {
    const int64_t k0 = k_begin + (int64_t)kt * BK;
    double* a = sA + slot * AT::doubles;
    double* b = sB + slot * BT::doubles;
    if (TA) load_tile<BK, BM, AT::pitch>(a, p.A, p.lda, k_end, p.M, k0, m0);
    else load_tile<BM, BK, AT::pitch>(a, p.A, p.lda, p.M, k_end, m0, k0);
    if (TB) load_tile<BN, BK, BT::pitch>(b, p.B, p.ldb, p.N, k_end, n0, k0);
    else load_tile<BK, BN, BT::pitch>(b, p.B, p.ldb, k_end, p.N, k0, n0);
  };

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    {
      const int next = kt + kStages - 1;
      if (next < KT) load_stage(next % kStages, next);
      cp_async_commit();
    }
    const double* a = sA + (kt % kStages) * AT::doubles;
    const double* b = sB + (kt % kStages) * BT::doubles;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[4], bf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = wm + i * 8 + g, k = kk + t;
        af[i] = TA ? a[k * AT::pitch + m] : a[m * AT::pitch + k];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = wn + j * 8 + g, k = kk + t;
        bf[j] = TB ? b[n * BT::pitch + k] : b[k * BT::pitch + n];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(acc[i][j], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // Epilogue: lane (g, t) owns C[m = 8i + g][n = 8j + 2t, 2t + 1].
  double* out = p.C + (int64_t)blockIdx.z * p.split_stride;
  const bool vec_ok = (p.ldc % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + wm + i * 8 + g;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t n = n0 + wn + j * 8 + 2 * t;
      if (n >= p.N) continue;
      double v0 = acc[i][j][0], v1 = acc[i][j][1];
      if (p.Cin != nullptr) {
        v0 += p.Cin[m * p.ldcin + n];
        if (n + 1 < p.N) v1 += p.Cin[m * p.ldcin + n + 1];
      }
      double* dst = out + m * p.ldc + n;
      if (n + 1 < p.N) {
        if (vec_ok) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
        else {
          dst[0] = v0;
          dst[1] = v1;
        }
      } else {
        dst[0] = v0;
      }
    }
  }
}

// out[i] = (cin ? cin[i] : 0) + sum_s partial[s][i]   (split-K fold, sequential in s)
__global__ void __launch_bounds__(256)
splitk_fold_kernel(const double* __restrict__ partial, int splits, int64_t M, int64_t N,
                   const double* __restrict__ cin, int64_t ldcin, double* __restrict__ out, int64_t ldc) {
  const int64_t total = M * N, step = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += step) {
    const int64_t m = i / N, n = i - m * N;
    double acc = partial[i];
    for (int s = 1; s < splits; ++s) acc += partial[(int64_t)s * total + i];
    if (cin != nullptr) acc += cin[m * ldcin + n];
    out[m * ldc + n] = acc;
  }
}

// ======================================================================================
// Generic tiled GEMM (any arithmetic dtype, any alignment) -- correctness path
// ======================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(int ta, int tb, int64_t M, int64_t N, int64_t K, const T* __restrict__ A, int64_t lda,
                 const T* __restrict__ B, int64_t ldb, const T* __restrict__ Cin, int64_t ldcin,
                 T* __restrict__ C, int64_t ldc) {
  constexpr int TS = 64, TK = 16;
  __shared__ T sA[TK][TS + 1];
  __shared__ T sB[TK][TS + 1];
  const int64_t m0 = (int64_t)blockIdx.y * TS, n0 = (int64_t)blockIdx.x * TS;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (int64_t k0 = 0; k0 < K; k0 += TK) {
    for (int e = threadIdx.x; e < TS * TK; e += 256) {
      int kk, mm;
      if (ta) { mm = e % TS; kk = e / TS; } else { kk = e % TK; mm = e / TK; }
      const int64_t gm = m0 + mm, gk = k0 + kk;
      T v = T(0);
      if (gm < M && gk < K) v = ta ? A[gk * lda + gm] : A[gm * lda + gk];
      sA[kk][mm] = v;
    }
    for (int e = threadIdx.x; e < TS * TK; e += 256) {
      int kk, nn;
      if (tb) { kk = e % TK; nn = e / TK; } else { nn = e % TS; kk = e / TS; }
      const int64_t gn = n0 + nn, gk = k0 + kk;
      T v = T(0);
      if (gn < N && gk < K) v = tb ? B[gn * ldb + gk] : B[gk * ldb + gn];
      sB[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= N) continue;
      T v = acc[i][j];
      if (Cin != nullptr) v += Cin[m * ldcin + n];
      C[m * ldc + n] = v;
    }
  }
}

// ======================================================================================
// Matrix-vector and dot kernels (HBM-bound)
// ======================================================================================
// y[r] (+)= sum_c A[r, c] x[c]  for a row-major A (rows x cols): one warp per row.
template <typename T>
__global__ void __launch_bounds__(256)
gemv_rows_kernel(const T* __restrict__ A, int64_t lda, int64_t rows, int64_t cols, const T* __restrict__ x,
                 int64_t incx, const T* __restrict__ yin, T* __restrict__ y, int64_t incy) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * 8;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += warps) {
    const T* row = A + r * lda;
    T acc0 = T(0), acc1 = T(0);
    int64_t c = lane;
    for (; c + 32 < cols; c += 64) {
      acc0 += row[c] * x[c * incx];
      acc1 += row[c + 32] * x[(c + 32) * incx];
    }
    if (c < cols) acc0 += row[c] * x[c * incx];
    T acc = acc0 + acc1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[r * incy] = yin ? acc + yin[r * incy] : acc;
  }
}

// Narrow rows (cols <= 64): a block stages a dense run of rows in shared memory with
// coalesced loads, then each thread reduces one row (pitch cols|1 keeps the reads
// conflict-free).  The LR forward pass X @ beta has cols = 28.
template <typename T>
__global__ void __launch_bounds__(256)
gemv_narrow_rows_kernel(const T* __restrict__ A, int64_t rows, int cols, const T* __restrict__ x,
                        int64_t incx, const T* __restrict__ yin, T* __restrict__ y, int64_t incy) {
  extern __shared__ __align__(16) unsigned char gemv_smem[];
  T* tile = reinterpret_cast<T*>(gemv_smem);
  const int pitch = cols | 1;
  T* xs = tile + 256 * pitch;
  for (int c = threadIdx.x; c < cols; c += 256) xs[c] = x[(int64_t)c * incx];
  const int64_t row_step = (int64_t)gridDim.x * 256;
  for (int64_t r0 = (int64_t)blockIdx.x * 256; r0 < rows; r0 += row_step) {
    const int64_t nrows = (rows - r0 < 256) ? rows - r0 : 256;
    const int64_t count = nrows * cols;
    const T* src = A + r0 * cols;
    __syncthreads();
    for (int e = threadIdx.x; e < (int)count; e += 256) {
      const int rr = e / cols, cc = e - rr * cols;
      tile[rr * pitch + cc] = src[e];
    }
    __syncthreads();
    if ((int64_t)threadIdx.x < nrows) {
      const T* row = tile + threadIdx.x * pitch;
      T acc = T(0);
      for (int c = 0; c < cols; ++c) acc += row[c] * xs[c];
      const int64_t r = r0 + threadIdx.x;
      y[r * incy] = yin ? acc + yin[r * incy] : acc;
    }
  }
}

// Partial y[c] = sum_{r in segment} A[r, c] w[r] for a dense row-major A with few columns
// (cols <= 256): flat coalesced sweep, thread t always sees column t % cols.
// grid.x = segments; partial layout (segment, cols).
template <typename T>
__global__ void __launch_bounds__(256)
gemv_t_narrow_kernel(const T* __restrict__ A, int64_t rows, int cols, const T* __restrict__ w, int64_t incw,
                     int64_t rows_per_seg, T* __restrict__ partial) {
  __shared__ T smem[256];
  const int groups = 256 / cols, active = groups * cols;
  const int64_t r_lo = (int64_t)blockIdx.x * rows_per_seg;
  int64_t r_hi = r_lo + rows_per_seg;
  if (r_hi > rows) r_hi = rows;
  T acc = T(0);
  if ((int)threadIdx.x < active) {
    const int c = threadIdx.x % cols;
    for (int64_t r = r_lo + threadIdx.x / cols; r < r_hi; r += groups) acc += A[r * cols + c] * w[r * incw];
  }
  smem[threadIdx.x] = acc;
  __syncthreads();
  if ((int)threadIdx.x < cols) {
    T total = smem[threadIdx.x];
    for (int gidx = 1; gidx < groups; ++gidx) total += smem[threadIdx.x + gidx * cols];
    partial[(int64_t)blockIdx.x * cols + threadIdx.x] = total;
  }
}

// Wide transposed GEMV: one thread per column, rows split over blockIdx.y.
template <typename T>
__global__ void __launch_bounds__(256)
gemv_t_wide_kernel(const T* __restrict__ A, int64_t lda, int64_t rows, int64_t cols, const T* __restrict__ w,
                   int64_t incw, int64_t rows_per_seg, T* __restrict__ partial) {
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const int64_t r_lo = (int64_t)blockIdx.y * rows_per_seg;
  int64_t r_hi = r_lo + rows_per_seg;
  if (r_hi > rows) r_hi = rows;
  T acc0 = T(0), acc1 = T(0);
  int64_t r = r_lo;
  for (; r + 1 < r_hi; r += 2) {
    acc0 += A[r * lda + c] * w[r * incw];
    acc1 += A[(r + 1) * lda + c] * w[(r + 1) * incw];
  }
  if (r < r_hi) acc0 += A[r * lda + c] * w[r * incw];
  partial[(int64_t)blockIdx.y * cols + c] = acc0 + acc1;
}

// y[c] = (yin ? yin[c] : 0) + sum_s partial[s][c]
template <typename T>
__global__ void __launch_bounds__(256)
fold_vector_kernel(const T* __restrict__ partial, int64_t segs, int64_t cols, const T* __restrict__ yin,
                   T* __restrict__ y, int64_t incy) {
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  T acc = T(0);
  for (int64_t s = 0; s < segs; ++s) acc += partial[s * cols + c];
  y[c * incy] = yin ? acc + yin[c * incy] : acc;
}

// Dot product partials: partial[b] = sum over the block's segment of x[i] * y[i].
template <typename T>
__global__ void __launch_bounds__(256)
dot_partial_kernel(const T* __restrict__ x, int64_t incx, const T* __restrict__ yv, int64_t incy, int64_t n,
                   int64_t seg_len, T* __restrict__ partial) {
  __shared__ T smem[8];
  const int64_t lo = (int64_t)blockIdx.x * seg_len;
  int64_t hi = lo + seg_len;
  if (hi > n) hi = n;
  T acc0 = T(0), acc1 = T(0);
  int64_t i = lo + threadIdx.x;
  for (; i + 256 < hi; i += 512) {
    acc0 += x[i * incx] * yv[i * incy];
    acc1 += x[(i + 256) * incx] * yv[(i + 256) * incy];
  }
  if (i < hi) acc0 += x[i * incx] * yv[i * incy];
  T acc = acc0 + acc1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    T total = smem[0];
    for (int wv = 1; wv < 8; ++wv) total += smem[wv];
    partial[blockIdx.x] = total;
  }
}

// ---- host-side dispatch --------------------------------------------------------------------------------------
template <typename T>
int run_gemv_rows(const T* A, int64_t lda, int64_t rows, int64_t cols, const T* x, int64_t incx,
                  const T* yin, T* y, int64_t incy, cudaStream_t s) {
  if (cols <= 64 && lda == cols) {
    const int pitch = (int)cols | 1;
    const size_t smem = (size_t)(256 * pitch + cols) * sizeof(T);
    const unsigned grid = blocks_for(rows, 256, (int64_t)sm_count() * 8);
    if (smem > 48 * 1024)
      NUMS_CUDA_OK(cudaFuncSetAttribute(gemv_narrow_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    gemv_narrow_rows_kernel<T><<<grid, 256, smem, s>>>(A, rows, (int)cols, x, incx, yin, y, incy);
  } else {
    const unsigned grid = blocks_for(rows, 8, (int64_t)sm_count() * 16);
    gemv_rows_kernel<T><<<grid, 256, 0, s>>>(A, lda, rows, cols, x, incx, yin, y, incy);
  }
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

// y[c] = sum_r A[r, c] w[r], A row-major rows x cols.
template <typename T>
int run_gemv_t(const T* A, int64_t lda, int64_t rows, int64_t cols, const T* w, int64_t incw, const T* yin,
               T* y, int64_t incy, void* ws, size_t ws_bytes, cudaStream_t s) {
  const int64_t want = (int64_t)sm_count() * 8;
  if (cols <= 256 && lda == cols) {
    const int groups = 256 / (int)cols;
    int64_t segs = want;
    const int64_t max_segs = ceil_div(rows, (int64_t)groups * 8);
    if (segs > max_segs) segs = max_segs;
    if (segs < 1) segs = 1;
    const int64_t rows_per_seg = ceil_div(rows, segs);
    segs = ceil_div(rows, rows_per_seg);
    NUMS_NEED_WS((size_t)(segs * cols) * sizeof(T), ws_bytes);
    T* partial = static_cast<T*>(ws);
    gemv_t_narrow_kernel<T><<<(unsigned)segs, 256, 0, s>>>(A, rows, (int)cols, w, incw, rows_per_seg, partial);
    NUMS_LAUNCH_OK();
    fold_vector_kernel<T><<<(unsigned)ceil_div(cols, 256), 256, 0, s>>>(partial, segs, cols, yin, y, incy);
    NUMS_LAUNCH_OK();
    return NUMS_OK;
  }
  const int64_t col_tiles = ceil_div(cols, 256);
  int64_t segs = ceil_div(want, col_tiles);
  const int64_t max_segs = ceil_div(rows, 64);
  if (segs > max_segs) segs = max_segs;
  if (segs > 65535) segs = 65535;
  if (segs < 1) segs = 1;
  const int64_t rows_per_seg = ceil_div(rows, segs);
  segs = ceil_div(rows, rows_per_seg);
  NUMS_NEED_WS((size_t)(segs * cols) * sizeof(T), ws_bytes);
  T* partial = static_cast<T*>(ws);
  dim3 grid((unsigned)col_tiles, (unsigned)segs);
  gemv_t_wide_kernel<T><<<grid, 256, 0, s>>>(A, lda, rows, cols, w, incw, rows_per_seg, partial);
  NUMS_LAUNCH_OK();
  fold_vector_kernel<T><<<(unsigned)col_tiles, 256, 0, s>>>(partial, segs, cols, yin, y, incy);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

template <typename T>
int run_dot(const T* x, int64_t incx, const T* y, int64_t incy, int64_t n, const T* cin, T* out, void* ws,
            size_t ws_bytes, cudaStream_t s) {
  int64_t segs = (int64_t)sm_count() * 4;
  int64_t seg_len = ceil_div(n, segs);
  if (seg_len < 2048) seg_len = 2048;
  segs = ceil_div(n, seg_len);
  NUMS_NEED_WS((size_t)segs * sizeof(T), ws_bytes);
  T* partial = static_cast<T*>(ws);
  dot_partial_kernel<T><<<(unsigned)segs, 256, 0, s>>>(x, incx, y, incy, n, seg_len, partial);
  NUMS_LAUNCH_OK();
  fold_vector_kernel<T><<<1, 256, 0, s>>>(partial, segs, 1, cin, out, 1);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

template <bool TA, bool TB>
int launch_dmma(const GemmParams& p, int splits, cudaStream_t s) {
  const size_t smem = (size_t)kStages * (ATile<TA>::doubles + BTile<TB>::doubles) * sizeof(double);
  NUMS_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma_kernel<TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  dim3 grid((unsigned)(p.tiles_m * p.tiles_n), 1, (unsigned)splits);
  dgemm_dmma_kernel<TA, TB><<<grid, kGemmThreads, smem, s>>>(p);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

int run_dgemm(int ta, int tb, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* B,
              int64_t ldb, const double* Cin, int64_t ldcin, double* C, int64_t ldc, void* ws,
              size_t ws_bytes, cudaStream_t s) {
  GemmParams p;
  p.A = A; p.B = B; p.Cin = Cin; p.C = C;
  p.M = M; p.N = N; p.K = K;
  p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.ldcin = ldcin;
  p.tiles_m = (int)ceil_div(M, BM);
  p.tiles_n = (int)ceil_div(N, BN);
  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
  const int sms = sm_count();
  // Split K when the output has too few tiles to fill the machine and K is long.
  int splits = 1;
  if (tiles * 2 <= sms && K >= 8 * BK) {
    splits = (int)((sms * 2) / tiles);
    const int64_t max_splits = ceil_div(K, 4 * BK);
    if (splits > max_splits) splits = (int)max_splits;
    if (splits > 1024) splits = 1024;
    if (splits < 1) splits = 1;
  }
  int64_t k_per_split = ceil_div(ceil_div(K, splits), BK) * BK;
  splits = (int)ceil_div(K, k_per_split);
  p.k_per_split = k_per_split;
  p.split_stride = 0;
  if (splits > 1) {
    const size_t need = (size_t)splits * M * N * sizeof(double);
    NUMS_NEED_WS(need, ws_bytes);
    p.C = static_cast<double*>(ws);
    p.ldc = N;
    p.Cin = nullptr;
    p.split_stride = M * N;
  }
  int rc;
  if (ta && tb) rc = launch_dmma<true, true>(p, splits, s);
  else if (ta) rc = launch_dmma<true, false>(p, splits, s);
  else if (tb) rc = launch_dmma<false, true>(p, splits, s);
  else rc = launch_dmma<false, false>(p, splits, s);
  if (rc) return rc;
  if (splits > 1) {
    splitk_fold_kernel<<<blocks_for(M * N, 256, (int64_t)sms * 8), 256, 0, s>>>(
        static_cast<const double*>(ws), splits, M, N, Cin, ldcin, C, ldc);
    NUMS_LAUNCH_OK();
  }
  return NUMS_OK;
}

template <typename T>
int run_gemm_typed(int ta, int tb, int64_t M, int64_t N, int64_t K, const T* A, int64_t lda, const T* B,
                   int64_t ldb, T* C, int64_t ldc, int accumulate, void* ws, size_t ws_bytes, cudaStream_t s) {
  const T* Cin = accumulate ? C : nullptr;
  // vector forms ---------------------------------------------------------------------------------
  if (M == 1 && N == 1) {
    const int64_t incx = ta ? lda : 1, incy = tb ? 1 : ldb;
    return run_dot<T>(A, incx, B, incy, K, Cin, C, ws, ws_bytes, s);
  }
  if (N == 1) {  // C[m] = op(A)[m,:] . b
    const int64_t incb = tb ? 1 : ldb;
    if (!ta) return run_gemv_rows<T>(A, lda, M, K, B, incb, Cin, C, ldc, s);
    return run_gemv_t<T>(A, lda, K, M, B, incb, Cin, C, ldc, ws, ws_bytes, s);
  }
  if (M == 1) {  // C[n] = a . op(B)[:, n]
    const int64_t inca = ta ? lda : 1;
    if (tb) return run_gemv_rows<T>(B, ldb, N, K, A, inca, Cin, C, 1, s);
    return run_gemv_t<T>(B, ldb, K, N, A, inca, Cin, C, 1, ws, ws_bytes, s);
  }
  if constexpr (std::is_same<T, double>::value) {
    const bool aligned = (lda % 2 == 0) && (ldb % 2 == 0) &&
                         ((reinterpret_cast<uintptr_t>(A) & 15u) == 0) &&
                         ((reinterpret_cast<uintptr_t>(B) & 15u) == 0);
    const bool worthwhile = M * N >= 32 * 32 || K >= 4096;
    if (aligned && worthwhile)
      return run_dgemm(ta, tb, M, N, K, A, lda, B, ldb, Cin, ldc, C, ldc, ws, ws_bytes, s);
  }
  dim3 grid((unsigned)ceil_div(N, 64), (unsigned)ceil_div(M, 64));
  NUMS_REQUIRE(grid.y <= 65535u, "gemm: M = %lld too large for the generic kernel", (long long)M);
  gemm_simt_kernel<T><<<grid, 256, 0, s>>>(ta, tb, M, N, K, A, lda, B, ldb, Cin, ldc, C, ldc);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

}  // namespace
}  // namespace nums

extern "C" int nums_gemm(int dtype, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k,
                         const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                         int accumulate, void* ws, size_t ws_bytes, void* stream) {
  using namespace nums;
  NUMS_REQUIRE(m >= 0 && n >= 0 && k >= 0, "gemm: negative extent");
  if (m == 0 || n == 0) return NUMS_OK;
  NUMS_REQUIRE(C != nullptr, "gemm: null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (k == 0) {
    if (accumulate) return NUMS_OK;
    nums_array_t out;
    out.data = C; out.dtype = dtype; out.ndim = 2;
    out.shape[0] = m; out.shape[1] = n; out.stride[0] = ldc; out.stride[1] = 1;
    return nums_fill(&out, 0.0, stream);
  }
  NUMS_REQUIRE(A != nullptr && B != nullptr, "gemm: null operand");
  NUMS_REQUIRE(lda >= (trans_a ? m : k) && ldb >= (trans_b ? k : n) && ldc >= n, "gemm: pitch smaller than row");
  switch (dtype) {
    case NUMS_F64:
      return run_gemm_typed<double>(trans_a, trans_b, m, n, k, static_cast<const double*>(A), lda,
                                    static_cast<const double*>(B), ldb, static_cast<double*>(C), ldc,
                                    accumulate, ws, ws_bytes, s);
    case NUMS_F32:
      return run_gemm_typed<float>(trans_a, trans_b, m, n, k, static_cast<const float*>(A), lda,
                                   static_cast<const float*>(B), ldb, static_cast<float*>(C), ldc,
                                   accumulate, ws, ws_bytes, s);
    case NUMS_I64:
      return run_gemm_typed<int64_t>(trans_a, trans_b, m, n, k, static_cast<const int64_t*>(A), lda,
                                     static_cast<const int64_t*>(B), ldb, static_cast<int64_t*>(C), ldc,
                                     accumulate, ws, ws_bytes, s);
    case NUMS_I32:
      return run_gemm_typed<int32_t>(trans_a, trans_b, m, n, k, static_cast<const int32_t*>(A), lda,
                                     static_cast<const int32_t*>(B), ldb, static_cast<int32_t*>(C), ldc,
                                     accumulate, ws, ws_bytes, s);
  }
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "gemm: dtype %s", dtype_name(dtype));
}
