// Binary elementwise ufuncs with NumPy broadcasting -- replaces np.<ufunc>(a1, a2) at
// nums/core/systems/numpy_compute.py:233-238 (and scipy.special.xlogy at :203-204).
//
// HBM-bound: the dense same-dtype case moves 128 bits per thread per access with four
// independent accesses in flight per operand; everything else goes through a strided
// kernel that converts dtypes on load.  Compiled with -fmad=false so IEEE add/sub/mul/div
// stay correctly rounded (bit-exact against NumPy).
#include "ops.cuh"

namespace nums {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

enum { kBothArrays = 0, kScalarA = 1, kScalarB = 2 };

// Streaming (evict-first) 128-bit accessors: every operand byte is touched exactly once.
template <typename S, int N>
__device__ __forceinline__ Vec<S, N> load_vec(const S* base, int64_t vec_index) {
  using V = Vec<S, N>;
  static_assert(sizeof(V) == 16 || sizeof(V) == 8 || sizeof(V) == 4 || sizeof(V) == 2 || sizeof(V) == 1, "");
  V out;
  if constexpr (sizeof(V) == 16) {
    int4 raw = __ldcs(reinterpret_cast<const int4*>(base) + vec_index);
    memcpy(&out, &raw, 16);
  } else {
    out = reinterpret_cast<const V*>(base)[vec_index];
  }
  return out;
}
template <typename S, int N>
__device__ __forceinline__ void store_vec(S* base, int64_t vec_index, const Vec<S, N>& v) {
  using V = Vec<S, N>;
  if constexpr (sizeof(V) == 16) {
    int4 raw;
    memcpy(&raw, &v, 16);
    __stcs(reinterpret_cast<int4*>(base) + vec_index, raw);
  } else if constexpr (sizeof(V) == 8) {
    int2 raw;
    memcpy(&raw, &v, 8);
    __stcs(reinterpret_cast<int2*>(base) + vec_index, raw);
  } else {
    reinterpret_cast<V*>(base)[vec_index] = v;
  }
}

template <typename T> __device__ __forceinline__ T from_storage(typename storage_of<T>::type s) {
  if constexpr (std::is_same<T, bool>::value) return s != 0;
  else return s;
}
template <typename T> __device__ __forceinline__ typename storage_of<T>::type to_storage(T v) {
  if constexpr (std::is_same<T, bool>::value) return v ? 1 : 0;
  else return v;
}

// Dense kernel: out[i] = OP(a[i], b[i]) with optional scalar operand.
template <class OP, typename T, int MODE>
__global__ void __launch_bounds__(kThreads)
bop_dense_kernel(const typename storage_of<T>::type* __restrict__ a,
                 const typename storage_of<T>::type* __restrict__ b,
                 const void* __restrict__ scalar, int scalar_dtype,
                 typename storage_of<typename OP::Out>::type* __restrict__ out, int64_t n) {
  using S = typename storage_of<T>::type;
  using O = typename OP::Out;
  using OS = typename storage_of<O>::type;
  constexpr int VEC = 16 / sizeof(S);
  constexpr int64_t kTile = (int64_t)kThreads * kUnroll * VEC;
  T sc = T(0);
  if constexpr (MODE != kBothArrays) sc = load_as<T>(scalar, scalar_dtype, 0);
  const int64_t base = (int64_t)blockIdx.x * kTile;
  if (base + kTile <= n) {
    const int64_t v0 = base / VEC + threadIdx.x;
    Vec<S, VEC> va[kUnroll], vb[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if constexpr (MODE != kScalarA) va[u] = load_vec<S, VEC>(a, v0 + (int64_t)u * kThreads);
      if constexpr (MODE != kScalarB) vb[u] = load_vec<S, VEC>(b, v0 + (int64_t)u * kThreads);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      Vec<OS, VEC> vo;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        T x = (MODE == kScalarA) ? sc : from_storage<T>(va[u].v[j]);
        T y = (MODE == kScalarB) ? sc : from_storage<T>(vb[u].v[j]);
        vo.v[j] = to_storage<O>(OP::apply(x, y));
      }
      store_vec<OS, VEC>(out, v0 + (int64_t)u * kThreads, vo);
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < n; i += kThreads) {
      T x = (MODE == kScalarA) ? sc : from_storage<T>(a[i]);
      T y = (MODE == kScalarB) ? sc : from_storage<T>(b[i]);
      out[i] = to_storage<O>(OP::apply(x, y));
    }
  }
}

// Row / column broadcast kernel: out(r, c) = OP(X(r, c), v) with X and out dense (rows x cols) and the
// other operand either one value per row (`s * X` with s of shape (n, 1), glms.py:236) or one value per
// column (`X - mean` with mean of shape (1, d), application.py:510-512).  Same 128-bit streaming accesses
// as the dense kernel; the small operand goes through the read-only cache (each value is reused `cols` /
// `rows` times).  cols is a multiple of the vector width, so a vector never straddles two rows.
enum { kPerRow = 0, kPerCol = 1 };
template <class OP, typename T, int KIND, bool SMALL_IS_A>
__global__ void __launch_bounds__(kThreads)
bop_rowcol_kernel(const typename storage_of<T>::type* __restrict__ x, const void* __restrict__ small, int small_dtype,
                  int64_t small_stride, typename storage_of<typename OP::Out>::type* __restrict__ out,
                  uint32_t n, uint32_t cols) {
  using S = typename storage_of<T>::type;
  using O = typename OP::Out;
  using OS = typename storage_of<O>::type;
  constexpr int VEC = 16 / sizeof(S);
  constexpr uint32_t kTile = kThreads * kUnroll * VEC;
  const uint32_t base = blockIdx.x * kTile;
  if (base + kTile <= n) {
    const uint32_t v0 = base / VEC + threadIdx.x;
    Vec<S, VEC> vx[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) vx[u] = load_vec<S, VEC>(x, (int64_t)v0 + (int64_t)u * kThreads);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t e = (v0 + u * kThreads) * VEC;
      const uint32_t row = e / cols, col = e - row * cols;
      T per_row = T(0);
      if constexpr (KIND == kPerRow) per_row = load_as<T>(small, small_dtype, (int64_t)row * small_stride);
      Vec<OS, VEC> vo;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        T sv = per_row;
        if constexpr (KIND == kPerCol) sv = load_as<T>(small, small_dtype, (int64_t)(col + j) * small_stride);
        T xv = from_storage<T>(vx[u].v[j]);
        vo.v[j] = to_storage<O>(SMALL_IS_A ? OP::apply(sv, xv) : OP::apply(xv, sv));
      }
      store_vec<OS, VEC>(out, (int64_t)v0 + (int64_t)u * kThreads, vo);
    }
  } else {
    for (uint32_t i = base + threadIdx.x; i < n; i += kThreads) {
      const uint32_t row = i / cols, col = i - row * cols;
      T sv = load_as<T>(small, small_dtype, (int64_t)(KIND == kPerRow ? row : col) * small_stride);
      T xv = from_storage<T>(x[i]);
      out[i] = to_storage<O>(SMALL_IS_A ? OP::apply(sv, xv) : OP::apply(xv, sv));
    }
  }
}

// General kernel: any broadcast / stride pattern over <= 8 collapsed axes, operands converted
// from their storage dtype to the loop dtype on load (e.g. f64 array (+) f32 0-d scalar from
// BlockArray.from_scalar, blockarray.py:47-58; `s * X`, glms.py:236; `X - mean`,
// application.py:510-512).
template <class OP, typename T, typename IDX>
__global__ void __launch_bounds__(kThreads)
bop_strided_kernel(DevLayout<3> L, const void* __restrict__ a, int a_dtype,
                   const void* __restrict__ b, int b_dtype,
                   typename storage_of<typename OP::Out>::type* __restrict__ out, IDX n) {
  using O = typename OP::Out;
  const IDX step = (IDX)gridDim.x * kThreads;
  for (IDX i = (IDX)blockIdx.x * kThreads + threadIdx.x; i < n; i += step) {
    int64_t off[3];
    unravel<3, IDX>(L, i, off);
    T x = load_as<T>(a, a_dtype, off[1]);
    T y = load_as<T>(b, b_dtype, off[2]);
    out[off[0]] = to_storage<O>(OP::apply(x, y));
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <template <typename> class OPT, typename T>
int launch_bop(const Layout3& L, const nums_array_t* a, const nums_array_t* b,
               const nums_array_t* out, cudaStream_t stream) {
  using OP = OPT<T>;
  using O = typename OP::Out;
  using S = typename storage_of<T>::type;
  using OS = typename storage_of<O>::type;
  NUMS_REQUIRE(out->dtype == dtype_of<O>::value, "bop: output dtype %s, the %s loop produces %s",
               dtype_name(out->dtype), dtype_name(dtype_of<T>::value), dtype_name(dtype_of<O>::value));
  const int64_t n = L.numel;
  if (n == 0) return NUMS_OK;
  constexpr int VEC = 16 / sizeof(S);
  const bool out_dense = layout_contiguous(L, 0);
  const bool a_same = a->dtype == dtype_of<T>::value, b_same = b->dtype == dtype_of<T>::value;
  const bool a_dense = layout_contiguous(L, 1) && a_same, b_dense = layout_contiguous(L, 2) && b_same;
  const bool a_scalar = layout_scalar(L, 1), b_scalar = layout_scalar(L, 2);
  const bool out_al = (reinterpret_cast<uintptr_t>(out->data) % (VEC * sizeof(OS))) == 0;
  if (out_dense && out_al && n > 1) {
    const unsigned grid = blocks_for(n, kThreads * kUnroll * VEC);
    OS* o = static_cast<OS*>(out->data);
    if (a_dense && b_dense && aligned16(a->data) && aligned16(b->data)) {
      bop_dense_kernel<OP, T, kBothArrays><<<grid, kThreads, 0, stream>>>(
          static_cast<const S*>(a->data), static_cast<const S*>(b->data), nullptr, 0, o, n);
      NUMS_LAUNCH_OK();
      return NUMS_OK;
    }
    if (a_dense && b_scalar && aligned16(a->data)) {
      bop_dense_kernel<OP, T, kScalarB><<<grid, kThreads, 0, stream>>>(
          static_cast<const S*>(a->data), nullptr, b->data, b->dtype, o, n);
      NUMS_LAUNCH_OK();
      return NUMS_OK;
    }
    if (b_dense && a_scalar && aligned16(b->data)) {
      bop_dense_kernel<OP, T, kScalarA><<<grid, kThreads, 0, stream>>>(
          nullptr, static_cast<const S*>(b->data), a->data, a->dtype, o, n);
      NUMS_LAUNCH_OK();
      return NUMS_OK;
    }
  }
  // (a 2-D collapsed layout is dense for an operand when its strides are {cols, 1}: the axes could not be merged
  // only because the OTHER operand is constant along one of them)
  auto dense2d = [&](int o) { return L.stride[o][1] == 1 && L.stride[o][0] == L.shape[1]; };
  if constexpr (std::is_floating_point<T>::value)     // (the float loops are the ones on the hot path; keeps build time down)
  if (out_al && L.ndim == 2 && dense2d(0) && n < (int64_t)0x7fffffff && L.shape[1] % VEC == 0) {
    // (rows, cols) (+) (rows, 1) / (1, cols): one dense operand, the other constant along one axis
    const unsigned grid = blocks_for(n, kThreads * kUnroll * VEC);
    OS* o = static_cast<OS*>(out->data);
    const uint32_t cols = (uint32_t)L.shape[1];
    for (int small_op = 1; small_op <= 2; ++small_op) {
      const int big_op = 3 - small_op;
      const nums_array_t* big = big_op == 1 ? a : b;
      const nums_array_t* small = small_op == 1 ? a : b;
      const bool big_dense = dense2d(big_op) && (big_op == 1 ? a_same : b_same) && aligned16(big->data);
      if (!big_dense) continue;
      const int64_t s0 = L.stride[small_op][0], s1 = L.stride[small_op][1];
      const bool per_row = s1 == 0 && s0 != 0, per_col = s0 == 0 && s1 != 0;
      if (!per_row && !per_col) continue;
      const S* xp = static_cast<const S*>(big->data);
#define NUMS_ROWCOL(KIND, SMALL_A)                                                                      \
      bop_rowcol_kernel<OP, T, KIND, SMALL_A><<<grid, kThreads, 0, stream>>>(                             \
          xp, small->data, small->dtype, (KIND) == kPerRow ? s0 : s1, o, (uint32_t)n, cols)
      if (per_row && small_op == 1) NUMS_ROWCOL(kPerRow, true);
      else if (per_row) NUMS_ROWCOL(kPerRow, false);
      else if (small_op == 1) NUMS_ROWCOL(kPerCol, true);
      else NUMS_ROWCOL(kPerCol, false);
#undef NUMS_ROWCOL
      NUMS_LAUNCH_OK();
      return NUMS_OK;
    }
  }
  DevLayout<3> D;
  D.ndim = L.ndim;
  bool fits32 = n < (int64_t)0x7fffffff;
  for (int d = 0; d < L.ndim; ++d) {
    D.shape[d] = (uint32_t)L.shape[d];
    D.shape64[d] = L.shape[d];
    for (int o = 0; o < 3; ++o) D.stride[o][d] = L.stride[o][d];
  }
  const unsigned grid = blocks_for(n, kThreads, (int64_t)sm_count() * 32);
  if (fits32)
    bop_strided_kernel<OP, T, uint32_t><<<grid, kThreads, 0, stream>>>(
        D, a->data, a->dtype, b->data, b->dtype, static_cast<OS*>(out->data), (uint32_t)n);
  else
    bop_strided_kernel<OP, T, uint64_t><<<grid, kThreads, 0, stream>>>(
        D, a->data, a->dtype, b->data, b->dtype, static_cast<OS*>(out->data), (uint64_t)n);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

enum TypeClass { kAll, kNumeric, kFloat, kInt, kIntBool };

template <template <typename> class OPT>
int dispatch_types(TypeClass cls, int loop_dtype, const Layout3& L, const nums_array_t* a,
                   const nums_array_t* b, const nums_array_t* out, cudaStream_t s, const char* name) {
  switch (loop_dtype) {
    case NUMS_F64: if (cls == kAll || cls == kNumeric || cls == kFloat) return launch_bop<OPT, double>(L, a, b, out, s); break;
    case NUMS_F32: if (cls == kAll || cls == kNumeric || cls == kFloat) return launch_bop<OPT, float>(L, a, b, out, s); break;
    case NUMS_I64: if (cls != kFloat) return launch_bop<OPT, int64_t>(L, a, b, out, s); break;
    case NUMS_I32: if (cls != kFloat) return launch_bop<OPT, int32_t>(L, a, b, out, s); break;
    case NUMS_BOOL: if (cls == kAll || cls == kIntBool) return launch_bop<OPT, bool>(L, a, b, out, s); break;
  }
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "bop %s has no %s loop", name, dtype_name(loop_dtype));
}

// Ops whose integer/bool instantiations would not compile are restricted by a second
// dispatcher that only names floating-point types.
template <template <typename> class OPT>
int dispatch_float(int loop_dtype, const Layout3& L, const nums_array_t* a, const nums_array_t* b,
                   const nums_array_t* out, cudaStream_t s, const char* name) {
  if (loop_dtype == NUMS_F64) return launch_bop<OPT, double>(L, a, b, out, s);
  if (loop_dtype == NUMS_F32) return launch_bop<OPT, float>(L, a, b, out, s);
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "bop %s has no %s loop", name, dtype_name(loop_dtype));
}
template <template <typename> class OPT>
int dispatch_numeric(int loop_dtype, const Layout3& L, const nums_array_t* a, const nums_array_t* b,
                     const nums_array_t* out, cudaStream_t s, const char* name) {
  if (loop_dtype == NUMS_I64) return launch_bop<OPT, int64_t>(L, a, b, out, s);
  if (loop_dtype == NUMS_I32) return launch_bop<OPT, int32_t>(L, a, b, out, s);
  return dispatch_float<OPT>(loop_dtype, L, a, b, out, s, name);
}
template <template <typename> class OPT>
int dispatch_int(int loop_dtype, bool with_bool, const Layout3& L, const nums_array_t* a,
                 const nums_array_t* b, const nums_array_t* out, cudaStream_t s, const char* name) {
  if (loop_dtype == NUMS_I64) return launch_bop<OPT, int64_t>(L, a, b, out, s);
  if (loop_dtype == NUMS_I32) return launch_bop<OPT, int32_t>(L, a, b, out, s);
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "bop %s has no %s loop", name, dtype_name(loop_dtype));
}
template <template <typename> class OPT>
int dispatch_intbool(int loop_dtype, const Layout3& L, const nums_array_t* a, const nums_array_t* b,
                     const nums_array_t* out, cudaStream_t s, const char* name) {
  if (loop_dtype == NUMS_BOOL) return launch_bop<OPT, bool>(L, a, b, out, s);
  return dispatch_int<OPT>(loop_dtype, false, L, a, b, out, s, name);
}

}  // namespace
}  // namespace nums

extern "C" int nums_bop(int op, int loop_dtype, const nums_array_t* a, const nums_array_t* b,
                        const nums_array_t* out, void* stream) {
  using namespace nums;
  if (int rc = check_array(a, "bop a")) return rc;
  if (int rc = check_array(b, "bop b")) return rc;
  if (int rc = check_array(out, "bop out")) return rc;
  const nums_array_t* arrs[3] = {out, a, b};
  Layout3 L;
  if (int rc = build_layout(arrs, 3, &L)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define ALL(OPT, NAME) return dispatch_types<op::OPT>(kAll, loop_dtype, L, a, b, out, s, NAME)
#define NUM(OPT, NAME) return dispatch_numeric<op::OPT>(loop_dtype, L, a, b, out, s, NAME)
#define FLT(OPT, NAME) return dispatch_float<op::OPT>(loop_dtype, L, a, b, out, s, NAME)
#define INT(OPT, NAME) return dispatch_int<op::OPT>(loop_dtype, false, L, a, b, out, s, NAME)
#define INB(OPT, NAME) return dispatch_intbool<op::OPT>(loop_dtype, L, a, b, out, s, NAME)
  switch (op) {
    case NUMS_BOP_ADD: ALL(Add, "add");
    case NUMS_BOP_SUBTRACT: NUM(Subtract, "subtract");
    case NUMS_BOP_MULTIPLY: ALL(Multiply, "multiply");
    case NUMS_BOP_TRUE_DIVIDE: FLT(TrueDivide, "true_divide");
    case NUMS_BOP_FLOOR_DIVIDE: NUM(FloorDivide, "floor_divide");
    case NUMS_BOP_REMAINDER: NUM(Remainder, "remainder");
    case NUMS_BOP_FMOD: NUM(Fmod, "fmod");
    case NUMS_BOP_POWER: NUM(Power, "power");
    case NUMS_BOP_FLOAT_POWER: FLT(Power, "float_power");
    case NUMS_BOP_MAXIMUM: ALL(Maximum, "maximum");
    case NUMS_BOP_MINIMUM: ALL(Minimum, "minimum");
    case NUMS_BOP_FMAX: ALL(Fmax, "fmax");
    case NUMS_BOP_FMIN: ALL(Fmin, "fmin");
    case NUMS_BOP_ARCTAN2: FLT(Arctan2, "arctan2");
    case NUMS_BOP_HYPOT: FLT(Hypot, "hypot");
    case NUMS_BOP_COPYSIGN: FLT(Copysign, "copysign");
    case NUMS_BOP_NEXTAFTER: FLT(Nextafter, "nextafter");
    case NUMS_BOP_HEAVISIDE: FLT(Heaviside, "heaviside");
    case NUMS_BOP_LOGADDEXP: FLT(Logaddexp, "logaddexp");
    case NUMS_BOP_LOGADDEXP2: FLT(Logaddexp2, "logaddexp2");
    case NUMS_BOP_LDEXP: FLT(Ldexp, "ldexp");
    case NUMS_BOP_XLOGY: FLT(Xlogy, "xlogy");
    case NUMS_BOP_LESS: ALL(Less, "less");
    case NUMS_BOP_LESS_EQUAL: ALL(LessEqual, "less_equal");
    case NUMS_BOP_GREATER: ALL(Greater, "greater");
    case NUMS_BOP_GREATER_EQUAL: ALL(GreaterEqual, "greater_equal");
    case NUMS_BOP_EQUAL: ALL(Equal, "equal");
    case NUMS_BOP_NOT_EQUAL: ALL(NotEqual, "not_equal");
    case NUMS_BOP_LOGICAL_AND: ALL(LogicalAnd, "logical_and");
    case NUMS_BOP_LOGICAL_OR: ALL(LogicalOr, "logical_or");
    case NUMS_BOP_LOGICAL_XOR: ALL(LogicalXor, "logical_xor");
    case NUMS_BOP_BITWISE_AND: INB(BitAnd, "bitwise_and");
    case NUMS_BOP_BITWISE_OR: INB(BitOr, "bitwise_or");
    case NUMS_BOP_BITWISE_XOR: INB(BitXor, "bitwise_xor");
    case NUMS_BOP_LEFT_SHIFT: INT(LeftShift, "left_shift");
    case NUMS_BOP_RIGHT_SHIFT: INT(RightShift, "right_shift");
    case NUMS_BOP_GCD: INT(Gcd, "gcd");
    case NUMS_BOP_LCM: INT(Lcm, "lcm");
  }
#undef ALL
#undef NUM
#undef FLT
#undef INT
#undef INB
  NUMS_FAIL(NUMS_ERR_INVALID, "bop: unknown op id %d", op);
}

// Same as nums_bop for the two shapes that dominate the per-block dispatch path -- both operands dense over
// the output's iteration space, or one of them a single value -- without the three array descriptors: the
// host hands over raw pointers and element counts (a_n / b_n: n = dense, 1 = broadcast scalar).
extern "C" int nums_bop_flat(int op, int loop_dtype, const void* a, int a_dtype, int64_t a_n, const void* b,
                             int b_dtype, int64_t b_n, void* out, int out_dtype, int64_t n, void* stream) {
  using namespace nums;
  NUMS_REQUIRE((a_n == n || a_n == 1) && (b_n == n || b_n == 1) && n >= 0,
               "bop_flat: operands must be dense over the output or single values");
  nums_array_t arr[3];
  const void* ptr[3] = {a, b, out};
  const int dt[3] = {a_dtype, b_dtype, out_dtype};
  const int64_t cnt[3] = {a_n, b_n, n};
  for (int i = 0; i < 3; ++i) {
    arr[i].data = const_cast<void*>(ptr[i]);
    arr[i].dtype = dt[i];
    arr[i].ndim = 1;
    arr[i].shape[0] = cnt[i];
    arr[i].stride[0] = (cnt[i] == n && n != 1) ? 1 : 0;
  }
  if (n == 1) arr[2].stride[0] = 1;
  return nums_bop(op, loop_dtype, &arr[0], &arr[1], &arr[2], stream);
}
