// Block reductions: reduce_axis (np.sum/min/max..., numpy_compute.py:177-181), arg_op
// (np.argmin/argmax with carried optimum, :269-283), allclose (:261-262) and the one-argument
// np.where (:188-194).
//
// All of these are HBM-bound streaming passes.  They are deterministic: no floating-point
// atomics; large reductions are split into a fixed number of segments whose partials are
// folded by a second launch, so results do not change from run to run.
#include <cfloat>
#include <climits>
#include "common.cuh"

namespace nums {
namespace {

constexpr int kThreads = 256;

// ---- combiners -------------------------------------------------------------------------------------
template <int OPK, typename A> struct Comb;
template <typename A> struct Comb<NUMS_RED_SUM, A> {
  static __device__ __forceinline__ A init() { return A(0); }
  static __device__ __forceinline__ A apply(A a, A b) { return a + b; }
  template <typename T> static __device__ __forceinline__ A lift(T v) { return (A)v; }
};
template <typename A> struct Comb<NUMS_RED_PROD, A> {
  static __device__ __forceinline__ A init() { return A(1); }
  static __device__ __forceinline__ A apply(A a, A b) { return a * b; }
  template <typename T> static __device__ __forceinline__ A lift(T v) { return (A)v; }
};
template <typename A> __device__ __forceinline__ A type_max();
template <> __device__ __forceinline__ double type_max<double>() { return INFINITY; }
template <> __device__ __forceinline__ float type_max<float>() { return INFINITY; }
template <> __device__ __forceinline__ int64_t type_max<int64_t>() { return LLONG_MAX; }
template <> __device__ __forceinline__ int32_t type_max<int32_t>() { return INT_MAX; }
template <typename A> __device__ __forceinline__ A type_min();
template <> __device__ __forceinline__ double type_min<double>() { return -INFINITY; }
template <> __device__ __forceinline__ float type_min<float>() { return -INFINITY; }
template <> __device__ __forceinline__ int64_t type_min<int64_t>() { return LLONG_MIN; }
template <> __device__ __forceinline__ int32_t type_min<int32_t>() { return INT_MIN; }

template <typename A> struct Comb<NUMS_RED_MIN, A> {
  static __device__ __forceinline__ A init() { return type_max<A>(); }
  static __device__ __forceinline__ A apply(A a, A b) {
    if (a != a) return a;  // NaN propagates (np.min), first NaN wins
    if (b != b) return b;
    return b < a ? b : a;
  }
  template <typename T> static __device__ __forceinline__ A lift(T v) { return (A)v; }
};
template <typename A> struct Comb<NUMS_RED_MAX, A> {
  static __device__ __forceinline__ A init() { return type_min<A>(); }
  static __device__ __forceinline__ A apply(A a, A b) {
    if (a != a) return a;
    if (b != b) return b;
    return b > a ? b : a;
  }
  template <typename T> static __device__ __forceinline__ A lift(T v) { return (A)v; }
};
template <typename A> struct Comb<NUMS_RED_ANY, A> {
  static __device__ __forceinline__ A init() { return A(0); }
  static __device__ __forceinline__ A apply(A a, A b) { return (a != 0 || b != 0) ? A(1) : A(0); }
  template <typename T> static __device__ __forceinline__ A lift(T v) { return v != T(0) ? A(1) : A(0); }
};
template <typename A> struct Comb<NUMS_RED_ALL, A> {
  static __device__ __forceinline__ A init() { return A(1); }
  static __device__ __forceinline__ A apply(A a, A b) { return (a != 0 && b != 0) ? A(1) : A(0); }
  template <typename T> static __device__ __forceinline__ A lift(T v) { return v != T(0) ? A(1) : A(0); }
};

template <typename A>
__device__ __forceinline__ void store_as(void* out, int dt, int64_t idx, A v) {
  switch (dt) {
    case NUMS_F64: static_cast<double*>(out)[idx] = (double)v; break;
    case NUMS_F32: static_cast<float*>(out)[idx] = (float)v; break;
    case NUMS_I64: static_cast<int64_t*>(out)[idx] = (int64_t)v; break;
    case NUMS_I32: static_cast<int32_t*>(out)[idx] = (int32_t)v; break;
    default: static_cast<uint8_t*>(out)[idx] = v != A(0) ? 1 : 0; break;
  }
}

template <class C, typename A>
__device__ __forceinline__ A block_fold(A v, A* smem) {
  v = warp_reduce(v, [](A x, A y) { return C::apply(x, y); });
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // smem may still be in use by a previous fold
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    A w = lane < (kThreads / 32) ? smem[lane] : C::init();
    // ordered fold so that the first NaN / first optimum wins deterministically
    w = warp_reduce(w, [](A x, A y) { return C::apply(x, y); });
    v = w;
  }
  return v;  // valid in warp 0
}

// Row reduction (inner == 1): grid = (segments, rows).  Each block folds one segment of one
// row; 4 independent accumulators per thread keep 4 loads in flight.
template <int OPK, typename T, typename A>
__global__ void __launch_bounds__(kThreads)
reduce_rows_kernel(const T* __restrict__ in, int64_t R, int64_t seg_len, void* out, int out_dtype) {
  using C = Comb<OPK, A>;
  __shared__ A smem[kThreads / 32];
  const int64_t row = blockIdx.y;
  const int64_t lo = (int64_t)blockIdx.x * seg_len;
  int64_t hi = lo + seg_len;
  if (hi > R) hi = R;
  const T* p = in + row * R;
  A acc0 = C::init(), acc1 = C::init(), acc2 = C::init(), acc3 = C::init();
  int64_t i = lo + threadIdx.x;
  for (; i + 3 * kThreads < hi; i += 4 * kThreads) {
    T v0 = p[i], v1 = p[i + kThreads], v2 = p[i + 2 * kThreads], v3 = p[i + 3 * kThreads];
    acc0 = C::apply(acc0, C::lift(v0));
    acc1 = C::apply(acc1, C::lift(v1));
    acc2 = C::apply(acc2, C::lift(v2));
    acc3 = C::apply(acc3, C::lift(v3));
  }
  for (; i < hi; i += kThreads) acc0 = C::apply(acc0, C::lift(p[i]));
  A acc = C::apply(C::apply(acc0, acc1), C::apply(acc2, acc3));
  acc = block_fold<C, A>(acc, smem);
  if (threadIdx.x == 0) store_as<A>(out, out_dtype, row * gridDim.x + blockIdx.x, acc);
}

// Short rows (R <= 64): one warp per row, 8 rows per block.
template <int OPK, typename T, typename A>
__global__ void __launch_bounds__(kThreads)
reduce_short_rows_kernel(const T* __restrict__ in, int64_t rows, int64_t R, void* out, int out_dtype) {
  using C = Comb<OPK, A>;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* p = in + row * R;
  A acc = C::init();
  for (int64_t i = lane; i < R; i += 32) acc = C::apply(acc, C::lift(p[i]));
  acc = warp_reduce(acc, [](A x, A y) { return C::apply(x, y); });
  if (lane == 0) store_as<A>(out, out_dtype, row, acc);
}

// Column reduction, narrow inner (inner <= kThreads): the block walks a contiguous range of
// rows of one `outer` slab as a flat, fully coalesced stream; thread t always sees column
// t % inner, and the per-thread accumulators of a column are folded through shared memory.
// grid = (segments, outer); output index = (seg * outer + o) * inner + c.
template <int OPK, typename T, typename A>
__global__ void __launch_bounds__(kThreads)
reduce_cols_narrow_kernel(const T* __restrict__ in, int64_t R, int64_t inner, int64_t rows_per_seg,
                          void* out, int out_dtype) {
  using C = Comb<OPK, A>;
  __shared__ A smem[kThreads];
  const int groups = kThreads / (int)inner;       // rows covered per sweep
  const int active = groups * (int)inner;
  const int64_t o = blockIdx.y, seg = blockIdx.x;
  int64_t r_lo = seg * rows_per_seg, r_hi = r_lo + rows_per_seg;
  if (r_hi > R) r_hi = R;
  const T* p = in + (o * R + r_lo) * inner;
  const int64_t count = (r_hi - r_lo) * inner;
  A acc0 = C::init(), acc1 = C::init();
  if ((int)threadIdx.x < active) {
    int64_t i = threadIdx.x;
    for (; i + active < count; i += 2 * active) {
      T v0 = p[i], v1 = p[i + active];
      acc0 = C::apply(acc0, C::lift(v0));
      acc1 = C::apply(acc1, C::lift(v1));
    }
    if (i < count) acc0 = C::apply(acc0, C::lift(p[i]));
  }
  smem[threadIdx.x] = C::apply(acc0, acc1);
  __syncthreads();
  if ((int)threadIdx.x < (int)inner) {
    A acc = smem[threadIdx.x];
    for (int g = 1; g < groups; ++g) acc = C::apply(acc, smem[threadIdx.x + g * (int)inner]);
    store_as<A>(out, out_dtype, (seg * gridDim.y + o) * inner + threadIdx.x, acc);
  }
}

// Column reduction, wide inner: one thread per column, coalesced across threads.
// grid = (column tiles, outer, segments).
template <int OPK, typename T, typename A>
__global__ void __launch_bounds__(kThreads)
reduce_cols_wide_kernel(const T* __restrict__ in, int64_t R, int64_t inner, int64_t rows_per_seg,
                        void* out, int out_dtype) {
  using C = Comb<OPK, A>;
  const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (c >= inner) return;
  const int64_t o = blockIdx.y, seg = blockIdx.z;
  int64_t r_lo = seg * rows_per_seg, r_hi = r_lo + rows_per_seg;
  if (r_hi > R) r_hi = R;
  const T* p = in + o * R * inner + c;
  A acc0 = C::init(), acc1 = C::init(), acc2 = C::init(), acc3 = C::init();
  int64_t r = r_lo;
  for (; r + 3 < r_hi; r += 4) {
    T v0 = p[r * inner], v1 = p[(r + 1) * inner], v2 = p[(r + 2) * inner], v3 = p[(r + 3) * inner];
    acc0 = C::apply(acc0, C::lift(v0));
    acc1 = C::apply(acc1, C::lift(v1));
    acc2 = C::apply(acc2, C::lift(v2));
    acc3 = C::apply(acc3, C::lift(v3));
  }
  for (; r < r_hi; ++r) acc0 = C::apply(acc0, C::lift(p[r * inner]));
  A acc = C::apply(C::apply(acc0, acc1), C::apply(acc2, acc3));
  store_as<A>(out, out_dtype, (seg * gridDim.y + o) * inner + c, acc);
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <int OPK, typename T, typename A>
int run_reduce(const void* in, int64_t outer, int64_t R, int64_t inner, void* out, int out_dtype,
               void* ws, size_t ws_bytes, cudaStream_t s, int depth = 0);

// Fold `S` stacked partial results of type A, laid out (S, outer*inner), into `out`.
template <int OPK, typename A>
int fold_partials(const void* partials, int64_t S, int64_t width, void* out, int out_dtype,
                  void* ws, size_t ws_bytes, cudaStream_t s, int depth) {
  // ANY/ALL/MIN/MAX/SUM/PROD are all closed under A, so the fold is the same reduction on A.
  return run_reduce<OPK, A, A>(partials, 1, S, width, out, out_dtype, ws, ws_bytes, s, depth + 1);
}

template <int OPK, typename T, typename A>
int run_reduce(const void* in_v, int64_t outer, int64_t R, int64_t inner, void* out, int out_dtype,
               void* ws, size_t ws_bytes, cudaStream_t s, int depth) {
  const T* in = static_cast<const T*>(in_v);
  const int sms = sm_count();
  const int64_t want_blocks = (int64_t)sms * 8;
  NUMS_REQUIRE(depth < 4, "reduce: partial folding did not converge");
  if (inner == 1) {
    if (R <= 64 && outer >= 64) {
      reduce_short_rows_kernel<OPK, T, A><<<(unsigned)ceil_div(outer, kThreads / 32), kThreads, 0, s>>>(
          in, outer, R, out, out_dtype);
      NUMS_LAUNCH_OK();
      return NUMS_OK;
    }
    NUMS_REQUIRE(outer <= 65535, "reduce: %lld rows of length %lld is not supported yet",
                 (long long)outer, (long long)R);
    int64_t S = 1;
    if (outer < want_blocks) {
      S = ceil_div(want_blocks, outer);
      const int64_t max_s = ceil_div(R, (int64_t)kThreads * 16);  // >= 4096 elements per segment
      if (S > max_s) S = max_s;
      if (S < 1) S = 1;
    }
    int64_t seg_len = ceil_div(R, S);
    seg_len = ceil_div(seg_len, kThreads) * kThreads;
    S = ceil_div(R, seg_len);
    dim3 grid((unsigned)S, (unsigned)outer);
    if (S == 1) {
      reduce_rows_kernel<OPK, T, A><<<grid, kThreads, 0, s>>>(in, R, seg_len, out, out_dtype);
      NUMS_LAUNCH_OK();
      return NUMS_OK;
    }
    const size_t need = (size_t)(outer * S) * sizeof(A);
    NUMS_NEED_WS(need, ws_bytes);
    reduce_rows_kernel<OPK, T, A><<<grid, kThreads, 0, s>>>(in, R, seg_len, ws, dtype_of<A>::value);
    NUMS_LAUNCH_OK();
    // partials are (outer, S): rows of length S
    char* rest = static_cast<char*>(ws) + ((need + 255) & ~(size_t)255);
    size_t rest_bytes = ws_bytes > ((need + 255) & ~(size_t)255) ? ws_bytes - ((need + 255) & ~(size_t)255) : 0;
    return run_reduce<OPK, A, A>(ws, outer, S, 1, out, out_dtype, rest, rest_bytes, s, depth + 1);
  }
  // inner > 1
  const int64_t width = outer * inner;
  if (inner <= kThreads) {
    NUMS_REQUIRE(outer <= 65535, "reduce: outer extent %lld too large for the narrow column kernel",
                 (long long)outer);
    const int groups = kThreads / (int)inner;
    int64_t S = ceil_div(want_blocks, outer);
    const int64_t max_s = ceil_div(R, (int64_t)groups * 16);
    if (S > max_s) S = max_s;
    if (S < 1) S = 1;
    int64_t rows_per_seg = ceil_div(R, S);
    S = ceil_div(R, rows_per_seg);
    dim3 grid((unsigned)S, (unsigned)outer);
    if (S == 1) {
      reduce_cols_narrow_kernel<OPK, T, A><<<grid, kThreads, 0, s>>>(in, R, inner, rows_per_seg, out, out_dtype);
      NUMS_LAUNCH_OK();
      return NUMS_OK;
    }
    const size_t need = (size_t)(S * width) * sizeof(A);
    NUMS_NEED_WS(need, ws_bytes);
    reduce_cols_narrow_kernel<OPK, T, A><<<grid, kThreads, 0, s>>>(in, R, inner, rows_per_seg, ws,
                                                                   dtype_of<A>::value);
    NUMS_LAUNCH_OK();
    char* rest = static_cast<char*>(ws) + ((need + 255) & ~(size_t)255);
    size_t used = (need + 255) & ~(size_t)255;
    return fold_partials<OPK, A>(ws, S, width, out, out_dtype, rest, ws_bytes > used ? ws_bytes - used : 0, s, depth);
  }
  {
    NUMS_REQUIRE(outer <= 65535, "reduce: outer extent %lld too large for the wide column kernel",
                 (long long)outer);
    const int64_t col_tiles = ceil_div(inner, kThreads);
    int64_t S = ceil_div(want_blocks, col_tiles * outer);
    const int64_t max_s = ceil_div(R, 32);
    if (S > max_s) S = max_s;
    if (S > 65535) S = 65535;
    if (S < 1) S = 1;
    int64_t rows_per_seg = ceil_div(R, S);
    S = ceil_div(R, rows_per_seg);
    dim3 grid((unsigned)col_tiles, (unsigned)outer, (unsigned)S);
    if (S == 1) {
      reduce_cols_wide_kernel<OPK, T, A><<<grid, kThreads, 0, s>>>(in, R, inner, rows_per_seg, out, out_dtype);
      NUMS_LAUNCH_OK();
      return NUMS_OK;
    }
    const size_t need = (size_t)(S * width) * sizeof(A);
    NUMS_NEED_WS(need, ws_bytes);
    reduce_cols_wide_kernel<OPK, T, A><<<grid, kThreads, 0, s>>>(in, R, inner, rows_per_seg, ws,
                                                                 dtype_of<A>::value);
    NUMS_LAUNCH_OK();
    size_t used = (need + 255) & ~(size_t)255;
    char* rest = static_cast<char*>(ws) + used;
    return fold_partials<OPK, A>(ws, S, width, out, out_dtype, rest, ws_bytes > used ? ws_bytes - used : 0, s, depth);
  }
}

template <int OPK>
int reduce_by_dtype(const void* a, int a_dtype, int64_t outer, int64_t R, int64_t inner, void* out,
                    int out_dtype, void* ws, size_t ws_bytes, cudaStream_t s) {
  constexpr bool arith = OPK == NUMS_RED_SUM || OPK == NUMS_RED_PROD;
  constexpr bool logic = OPK == NUMS_RED_ANY || OPK == NUMS_RED_ALL;
  switch (a_dtype) {
    case NUMS_F64:
      return run_reduce<OPK, double, typename std::conditional<logic, int32_t, double>::type>(
          a, outer, R, inner, out, out_dtype, ws, ws_bytes, s);
    case NUMS_F32:
      return run_reduce<OPK, float,
                        typename std::conditional<logic, int32_t,
                                                  typename std::conditional<arith, double, float>::type>::type>(
          a, outer, R, inner, out, out_dtype, ws, ws_bytes, s);
    case NUMS_I64:
      return run_reduce<OPK, int64_t, typename std::conditional<logic, int32_t, int64_t>::type>(
          a, outer, R, inner, out, out_dtype, ws, ws_bytes, s);
    case NUMS_I32:
      return run_reduce<OPK, int32_t,
                        typename std::conditional<logic, int32_t,
                                                  typename std::conditional<arith, int64_t, int32_t>::type>::type>(
          a, outer, R, inner, out, out_dtype, ws, ws_bytes, s);
    case NUMS_BOOL:
      return run_reduce<OPK, uint8_t,
                        typename std::conditional<arith, int64_t, int32_t>::type>(
          a, outer, R, inner, out, out_dtype, ws, ws_bytes, s);
  }
  NUMS_FAIL(NUMS_ERR_INVALID, "reduce: unknown dtype %d", a_dtype);
}

// ---- arg_op -----------------------------------------------------------------------------------------------
template <typename T> struct ArgPair {
  T v;
  int64_t i;
};
// `b` replaces `a` when it is better; NaN counts as the optimum (NumPy returns the first NaN).
template <bool IS_MAX, typename T>
__device__ __forceinline__ ArgPair<T> arg_better(ArgPair<T> a, ArgPair<T> b) {
  if (b.i < 0) return a;
  if (a.i < 0) return b;
  const bool a_nan = a.v != a.v, b_nan = b.v != b.v;
  if (a_nan || b_nan) {
    if (a_nan && b_nan) return b.i < a.i ? b : a;
    return a_nan ? a : b;
  }
  const bool strictly = IS_MAX ? (b.v > a.v) : (b.v < a.v);
  if (strictly || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
template <bool IS_MAX, typename T>
__device__ __forceinline__ ArgPair<T> arg_block_fold(ArgPair<T> p, ArgPair<T>* smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgPair<T> q;
    q.v = __shfl_xor_sync(0xffffffffu, p.v, o);
    q.i = __shfl_xor_sync(0xffffffffu, p.i, o);
    p = arg_better<IS_MAX, T>(p, q);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = p;
  __syncthreads();
  if (warp == 0) {
    ArgPair<T> q = lane < kThreads / 32 ? smem[lane] : ArgPair<T>{T(0), -1};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ArgPair<T> r;
      r.v = __shfl_xor_sync(0xffffffffu, q.v, o);
      r.i = __shfl_xor_sync(0xffffffffu, q.i, o);
      q = arg_better<IS_MAX, T>(q, r);
    }
    p = q;
  }
  return p;
}

template <bool IS_MAX, typename T>
__global__ void __launch_bounds__(kThreads)
arg_stage1_kernel(const T* __restrict__ a, int64_t n, int64_t seg_len, T* __restrict__ pv, int64_t* __restrict__ pi) {
  __shared__ ArgPair<T> smem[kThreads / 32];
  const int64_t lo = (int64_t)blockIdx.x * seg_len;
  int64_t hi = lo + seg_len;
  if (hi > n) hi = n;
  ArgPair<T> best{T(0), -1};
  for (int64_t i = lo + threadIdx.x; i < hi; i += kThreads) best = arg_better<IS_MAX, T>(best, ArgPair<T>{a[i], i});
  best = arg_block_fold<IS_MAX, T>(best, smem);
  if (threadIdx.x == 0) {
    pv[blockIdx.x] = best.v;
    pi[blockIdx.x] = best.i;
  }
}
template <bool IS_MAX, typename T>
__global__ void __launch_bounds__(kThreads)
arg_stage2_kernel(const T* __restrict__ pv, const int64_t* __restrict__ pi, int parts, int64_t offset,
                  const int64_t* carried_index, const T* carried_value, int64_t* out_index, T* out_value) {
  __shared__ ArgPair<T> smem[kThreads / 32];
  ArgPair<T> best{T(0), -1};
  for (int i = threadIdx.x; i < parts; i += kThreads) best = arg_better<IS_MAX, T>(best, ArgPair<T>{pv[i], pi[i]});
  best = arg_block_fold<IS_MAX, T>(best, smem);
  if (threadIdx.x == 0) {
    int64_t idx = best.i + offset;
    T val = best.v;
    if (carried_index != nullptr && carried_value != nullptr) {
      const T cv = *carried_value;
      const bool carried_wins = IS_MAX ? (cv > val) : (cv < val);  // strict, false on NaN
      if (carried_wins) {
        idx = *carried_index;
        val = cv;
      }
    }
    *out_index = idx;
    *out_value = val;
  }
}

template <typename T>
int run_arg_op(int is_max, const void* a, int64_t n, int64_t offset, const int64_t* ci, const void* cv,
               int64_t* out_index, void* out_value, void* ws, size_t ws_bytes, cudaStream_t s) {
  int64_t parts = (int64_t)sm_count() * 4;
  int64_t seg_len = ceil_div(n, parts);
  if (seg_len < 1024) seg_len = 1024;
  parts = ceil_div(n, seg_len);
  const size_t need = (size_t)parts * (sizeof(T) + sizeof(int64_t)) + 256;
  NUMS_NEED_WS(need, ws_bytes);
  int64_t* pi = static_cast<int64_t*>(ws);
  T* pv = reinterpret_cast<T*>(static_cast<char*>(ws) + (((size_t)parts * sizeof(int64_t) + 255) & ~(size_t)255));
  if (is_max) {
    arg_stage1_kernel<true, T><<<(unsigned)parts, kThreads, 0, s>>>(static_cast<const T*>(a), n, seg_len, pv, pi);
    NUMS_LAUNCH_OK();
    arg_stage2_kernel<true, T><<<1, kThreads, 0, s>>>(pv, pi, (int)parts, offset, ci, static_cast<const T*>(cv),
                                                     out_index, static_cast<T*>(out_value));
  } else {
    arg_stage1_kernel<false, T><<<(unsigned)parts, kThreads, 0, s>>>(static_cast<const T*>(a), n, seg_len, pv, pi);
    NUMS_LAUNCH_OK();
    arg_stage2_kernel<false, T><<<1, kThreads, 0, s>>>(pv, pi, (int)parts, offset, ci, static_cast<const T*>(cv),
                                                      out_index, static_cast<T*>(out_value));
  }
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

// ---- allclose -------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
allclose_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t n, double rtol, double atol,
                uint8_t* flag) {
  const int64_t step = (int64_t)gridDim.x * kThreads;
  bool ok = true;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += step) {
    const double x = (double)a[i], y = (double)b[i];
    const bool close = (x == y) || (isfinite(x) && isfinite(y) && fabs(x - y) <= atol + rtol * fabs(y));
    ok = ok && close;
  }
  if (!ok) *flag = 0;  // benign race: every writer stores the same value
}

// ---- nonzero -----------------------------------------------------------------------------------------------------
constexpr int kNzItems = 8;
constexpr int kNzTile = kThreads * kNzItems;

template <typename T>
__global__ void __launch_bounds__(kThreads)
nonzero_count_kernel(const T* __restrict__ a, int64_t n, int64_t* __restrict__ block_counts) {
  __shared__ int smem[kThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kNzTile + (int64_t)threadIdx.x * kNzItems;
  int c = 0;
#pragma unroll
  for (int j = 0; j < kNzItems; ++j)
    if (base + j < n && a[base + j] != T(0)) ++c;
  c = warp_reduce(c, [](int x, int y) { return x + y; });
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += smem[w];
    block_counts[blockIdx.x] = t;
  }
}

// In-place exclusive scan of block_counts[0..nblocks) by one block; total -> *count.
__global__ void __launch_bounds__(kThreads)
nonzero_scan_kernel(int64_t* __restrict__ block_counts, int64_t nblocks, int64_t* __restrict__ count) {
  __shared__ int64_t warp_tot[kThreads / 32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < nblocks; base += kThreads) {
    const int64_t i = base + threadIdx.x;
    int64_t v = i < nblocks ? block_counts[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int64_t before = carry;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    if (i < nblocks) block_counts[i] = before + incl - v;
    __syncthreads();
    if (threadIdx.x == kThreads - 1) carry = before + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}

struct NzMeta {
  int ndim;
  int64_t shape[NUMS_MAX_DIMS];
  int64_t offset[NUMS_MAX_DIMS];
  int64_t* out[NUMS_MAX_DIMS];
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
nonzero_fill_kernel(const T* __restrict__ a, int64_t n, const int64_t* __restrict__ block_offsets, NzMeta m) {
  __shared__ int warp_tot[kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * kNzTile + (int64_t)threadIdx.x * kNzItems;
  bool flag[kNzItems];
  int c = 0;
#pragma unroll
  for (int j = 0; j < kNzItems; ++j) {
    flag[j] = base + j < n && a[base + j] != T(0);
    c += flag[j] ? 1 : 0;
  }
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += warp_tot[w];
  int64_t pos = block_offsets[blockIdx.x] + before + incl - c;
#pragma unroll
  for (int j = 0; j < kNzItems; ++j) {
    if (!flag[j]) continue;
    int64_t lin = base + j;
    for (int d = m.ndim - 1; d >= 0; --d) {
      const int64_t q = lin / m.shape[d];
      m.out[d][pos] = (lin - q * m.shape[d]) + m.offset[d];
      lin = q;
    }
    ++pos;
  }
}

template <typename F>
int by_storage(int dtype, F&& f) {
  switch (dtype) {
    case NUMS_F64: return f(double());
    case NUMS_F32: return f(float());
    case NUMS_I64: return f(int64_t());
    case NUMS_I32: return f(int32_t());
    case NUMS_BOOL: return f(uint8_t());
  }
  NUMS_FAIL(NUMS_ERR_INVALID, "unknown dtype %d", dtype);
}

}  // namespace
}  // namespace nums

extern "C" int nums_reduce(int op, const void* a, int a_dtype, int64_t outer, int64_t reduce,
                           int64_t inner, void* out, int out_dtype, void* ws, size_t ws_bytes,
                           void* stream) {
  using namespace nums;
  NUMS_REQUIRE(outer >= 0 && reduce >= 0 && inner >= 0, "reduce: negative extent");
  NUMS_REQUIRE(dtype_size(a_dtype) > 0 && dtype_size(out_dtype) > 0, "reduce: unknown dtype");
  if (outer * inner == 0) return NUMS_OK;
  NUMS_REQUIRE(reduce > 0, "reduce: zero-size reduction axis must be handled by the caller");
  NUMS_REQUIRE(a != nullptr && out != nullptr, "reduce: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (op) {
    case NUMS_RED_SUM: return reduce_by_dtype<NUMS_RED_SUM>(a, a_dtype, outer, reduce, inner, out, out_dtype, ws, ws_bytes, s);
    case NUMS_RED_PROD: return reduce_by_dtype<NUMS_RED_PROD>(a, a_dtype, outer, reduce, inner, out, out_dtype, ws, ws_bytes, s);
    case NUMS_RED_MIN: return reduce_by_dtype<NUMS_RED_MIN>(a, a_dtype, outer, reduce, inner, out, out_dtype, ws, ws_bytes, s);
    case NUMS_RED_MAX: return reduce_by_dtype<NUMS_RED_MAX>(a, a_dtype, outer, reduce, inner, out, out_dtype, ws, ws_bytes, s);
    case NUMS_RED_ANY: return reduce_by_dtype<NUMS_RED_ANY>(a, a_dtype, outer, reduce, inner, out, out_dtype, ws, ws_bytes, s);
    case NUMS_RED_ALL: return reduce_by_dtype<NUMS_RED_ALL>(a, a_dtype, outer, reduce, inner, out, out_dtype, ws, ws_bytes, s);
  }
  NUMS_FAIL(NUMS_ERR_INVALID, "reduce: unknown op id %d", op);
}

extern "C" int nums_arg_op(int is_max, const void* a, int a_dtype, int64_t n, int64_t index_offset,
                           const int64_t* carried_index, const void* carried_value,
                           int64_t* out_index, void* out_value, void* ws, size_t ws_bytes,
                           void* stream) {
  using namespace nums;
  NUMS_REQUIRE(n > 0, "arg_op: attempt to get argmin/argmax of an empty sequence");
  NUMS_REQUIRE(a && out_index && out_value, "arg_op: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (a_dtype) {
    case NUMS_F64: return run_arg_op<double>(is_max, a, n, index_offset, carried_index, carried_value, out_index, out_value, ws, ws_bytes, s);
    case NUMS_F32: return run_arg_op<float>(is_max, a, n, index_offset, carried_index, carried_value, out_index, out_value, ws, ws_bytes, s);
    case NUMS_I64: return run_arg_op<int64_t>(is_max, a, n, index_offset, carried_index, carried_value, out_index, out_value, ws, ws_bytes, s);
    case NUMS_I32: return run_arg_op<int32_t>(is_max, a, n, index_offset, carried_index, carried_value, out_index, out_value, ws, ws_bytes, s);
    case NUMS_BOOL: return run_arg_op<uint8_t>(is_max, a, n, index_offset, carried_index, carried_value, out_index, out_value, ws, ws_bytes, s);
  }
  NUMS_FAIL(NUMS_ERR_INVALID, "arg_op: unknown dtype %d", a_dtype);
}

extern "C" int nums_allclose(const void* a, const void* b, int dtype, int64_t numel, double rtol,
                             double atol, uint8_t* out_flag, void* ws, size_t ws_bytes, void* stream) {
  using namespace nums;
  (void)ws;
  (void)ws_bytes;
  NUMS_REQUIRE(out_flag != nullptr, "allclose: null flag pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  NUMS_CUDA_OK(cudaMemsetAsync(out_flag, 1, 1, s));
  if (numel == 0) return NUMS_OK;
  const unsigned grid = blocks_for(numel, kThreads, (int64_t)sm_count() * 16);
  return by_storage(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    allclose_kernel<T><<<grid, kThreads, 0, s>>>(static_cast<const T*>(a), static_cast<const T*>(b), numel, rtol,
                                                 atol, out_flag);
    NUMS_LAUNCH_OK();
    return NUMS_OK;
  });
}

extern "C" int nums_nonzero_count(const void* a, int dtype, int64_t numel, int64_t* count, void* ws,
                                  size_t ws_bytes, void* stream) {
  using namespace nums;
  NUMS_REQUIRE(count != nullptr, "nonzero: null count pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (numel == 0) {
    NUMS_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int64_t), s));
    return NUMS_OK;
  }
  const int64_t nblocks = ceil_div(numel, kNzTile);
  NUMS_NEED_WS((size_t)nblocks * sizeof(int64_t), ws_bytes);
  int64_t* counts = static_cast<int64_t*>(ws);
  int rc = by_storage(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    nonzero_count_kernel<T><<<(unsigned)nblocks, kThreads, 0, s>>>(static_cast<const T*>(a), numel, counts);
    NUMS_LAUNCH_OK();
    return NUMS_OK;
  });
  if (rc) return rc;
  nonzero_scan_kernel<<<1, kThreads, 0, s>>>(counts, nblocks, count);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

extern "C" int nums_nonzero_fill(const void* a, int dtype, int ndim, const int64_t* shape_host,
                                 const int64_t* offsets_host, int64_t* const* outs_host, void* ws,
                                 size_t ws_bytes, void* stream) {
  using namespace nums;
  NUMS_REQUIRE(ndim >= 1 && ndim <= NUMS_MAX_DIMS, "nonzero: ndim %d out of range", ndim);
  NzMeta m;
  m.ndim = ndim;
  int64_t numel = 1;
  for (int d = 0; d < ndim; ++d) {
    m.shape[d] = shape_host[d];
    m.offset[d] = offsets_host ? offsets_host[d] : 0;
    m.out[d] = outs_host[d];
    numel *= shape_host[d];
  }
  if (numel == 0) return NUMS_OK;
  const int64_t nblocks = ceil_div(numel, kNzTile);
  NUMS_NEED_WS((size_t)nblocks * sizeof(int64_t), ws_bytes);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t* offsets = static_cast<const int64_t*>(ws);
  return by_storage(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    nonzero_fill_kernel<T><<<(unsigned)nblocks, kThreads, 0, s>>>(static_cast<const T*>(a), numel, offsets, m);
    NUMS_LAUNCH_OK();
    return NUMS_OK;
  });
}
