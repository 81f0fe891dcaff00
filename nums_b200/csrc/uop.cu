// Unary elementwise ufuncs, dtype-converting strided copies, n-ary block sum and constant
// fills -- replaces np.<ufunc>(arr) (numpy_compute.py:184-186), arr.astype (:206-208), the
// slice assignments of create_block / update_block (:119-169), np.add.reduce(arrs) (:210-211)
// and np.zeros / ones / eye / arange (:96-104, :174-175).
#include "ops.cuh"

namespace nums {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

template <typename T> __device__ __forceinline__ T from_storage(typename storage_of<T>::type s) {
  if constexpr (std::is_same<T, bool>::value) return s != 0;
  else return s;
}
template <typename T> __device__ __forceinline__ typename storage_of<T>::type to_storage(T v) {
  if constexpr (std::is_same<T, bool>::value) return v ? 1 : 0;
  else return v;
}

template <class OP, typename T>
__global__ void __launch_bounds__(kThreads)
uop_dense_kernel(const typename storage_of<T>::type* __restrict__ a,
                 typename storage_of<typename OP::Out>::type* __restrict__ out, int64_t n) {
  using S = typename storage_of<T>::type;
  using O = typename OP::Out;
  using OS = typename storage_of<O>::type;
  constexpr int VEC = 16 / sizeof(S);
  constexpr int64_t kTile = (int64_t)kThreads * kUnroll * VEC;
  const int64_t base = (int64_t)blockIdx.x * kTile;
  if (base + kTile <= n) {
    const int64_t v0 = base / VEC + threadIdx.x;
    Vec<S, VEC> va[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      int4 raw = __ldcs(reinterpret_cast<const int4*>(a) + v0 + (int64_t)u * kThreads);
      memcpy(&va[u], &raw, 16);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      Vec<OS, VEC> vo;
#pragma unroll
      for (int j = 0; j < VEC; ++j) vo.v[j] = to_storage<O>(OP::apply(from_storage<T>(va[u].v[j])));
      reinterpret_cast<Vec<OS, VEC>*>(out)[v0 + (int64_t)u * kThreads] = vo;
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < n; i += kThreads)
      out[i] = to_storage<O>(OP::apply(from_storage<T>(a[i])));
  }
}

template <class OP, typename T, typename IDX>
__global__ void __launch_bounds__(kThreads)
uop_strided_kernel(DevLayout<2> L, const void* __restrict__ a, int a_dtype,
                   typename storage_of<typename OP::Out>::type* __restrict__ out, IDX n) {
  using O = typename OP::Out;
  const IDX step = (IDX)gridDim.x * kThreads;
  for (IDX i = (IDX)blockIdx.x * kThreads + threadIdx.x; i < n; i += step) {
    int64_t off[2];
    unravel<2, IDX>(L, i, off);
    out[off[0]] = to_storage<O>(OP::apply(load_as<T>(a, a_dtype, off[1])));
  }
}

// 2-D transposing copy through shared memory (both sides coalesced): out[r, c] = a[c, r]
// where `a` is a dense (cols, rows) matrix.  Used when a lazily transposed block
// (Block.transpose, base.py:72-85) has to be materialised.
template <typename S>
__global__ void __launch_bounds__(256)
transpose2d_kernel(const S* __restrict__ a, S* __restrict__ out, int64_t rows, int64_t cols) {
  __shared__ S tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {  // read a[c0 + j, r0 + tx]
    int64_t c = c0 + j, r = r0 + tx;
    if (c < cols && r < rows) tile[j][tx] = a[c * rows + r];
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {  // write out[r0 + j, c0 + tx]
    int64_t r = r0 + j, c = c0 + tx;
    if (r < rows && c < cols) out[r * cols + c] = tile[tx][j];
  }
}

template <template <typename> class OPT, typename T>
int launch_uop(const Layout3& L, const nums_array_t* a, const nums_array_t* out, cudaStream_t stream,
               bool is_copy) {
  using OP = OPT<T>;
  using O = typename OP::Out;
  using S = typename storage_of<T>::type;
  using OS = typename storage_of<O>::type;
  NUMS_REQUIRE(out->dtype == dtype_of<O>::value, "uop: output dtype %s, the %s loop produces %s",
               dtype_name(out->dtype), dtype_name(dtype_of<T>::value), dtype_name(dtype_of<O>::value));
  const int64_t n = L.numel;
  if (n == 0) return NUMS_OK;
  constexpr int VEC = 16 / sizeof(S);
  const bool same = a->dtype == dtype_of<T>::value;
  if (same && layout_contiguous(L, 0) && layout_contiguous(L, 1) &&
      (reinterpret_cast<uintptr_t>(a->data) & 15u) == 0 &&
      (reinterpret_cast<uintptr_t>(out->data) % (VEC * sizeof(OS))) == 0) {
    uop_dense_kernel<OP, T><<<blocks_for(n, kThreads * kUnroll * VEC), kThreads, 0, stream>>>(
        static_cast<const S*>(a->data), static_cast<OS*>(out->data), n);
    NUMS_LAUNCH_OK();
    return NUMS_OK;
  }
  // Dense 2-D transpose: out (rows, cols) dense, a walks (1, rows) over the same space.
  if (is_copy && same && L.ndim == 2 && L.stride[0][1] == 1 && L.stride[0][0] == L.shape[1] &&
      L.stride[1][0] == 1 && L.stride[1][1] == L.shape[0] && L.shape[0] >= 16 && L.shape[1] >= 16) {
    dim3 grid((unsigned)((L.shape[1] + 31) / 32), (unsigned)((L.shape[0] + 31) / 32));
    if (grid.y <= 65535u) {
      transpose2d_kernel<S><<<grid, 256, 0, stream>>>(static_cast<const S*>(a->data),
                                                      static_cast<S*>(out->data), L.shape[0], L.shape[1]);
      NUMS_LAUNCH_OK();
      return NUMS_OK;
    }
  }
  DevLayout<2> D;
  D.ndim = L.ndim;
  for (int d = 0; d < L.ndim; ++d) {
    D.shape[d] = (uint32_t)L.shape[d];
    D.shape64[d] = L.shape[d];
    D.stride[0][d] = L.stride[0][d];
    D.stride[1][d] = L.stride[1][d];
  }
  const unsigned grid = blocks_for(n, kThreads, (int64_t)sm_count() * 32);
  if (n < (int64_t)0x7fffffff)
    uop_strided_kernel<OP, T, uint32_t><<<grid, kThreads, 0, stream>>>(D, a->data, a->dtype,
                                                                      static_cast<OS*>(out->data), (uint32_t)n);
  else
    uop_strided_kernel<OP, T, uint64_t><<<grid, kThreads, 0, stream>>>(D, a->data, a->dtype,
                                                                      static_cast<OS*>(out->data), (uint64_t)n);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

#define NUMS_UOP_FAIL(name) \
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "uop %s has no %s loop", name, dtype_name(loop_dtype))

template <template <typename> class OPT>
int uop_all(int loop_dtype, const Layout3& L, const nums_array_t* a, const nums_array_t* out,
            cudaStream_t s, const char* name, bool is_copy = false) {
  switch (loop_dtype) {
    case NUMS_F64: return launch_uop<OPT, double>(L, a, out, s, is_copy);
    case NUMS_F32: return launch_uop<OPT, float>(L, a, out, s, is_copy);
    case NUMS_I64: return launch_uop<OPT, int64_t>(L, a, out, s, is_copy);
    case NUMS_I32: return launch_uop<OPT, int32_t>(L, a, out, s, is_copy);
    case NUMS_BOOL: return launch_uop<OPT, bool>(L, a, out, s, is_copy);
  }
  NUMS_UOP_FAIL(name);
}
template <template <typename> class OPT>
int uop_float(int loop_dtype, const Layout3& L, const nums_array_t* a, const nums_array_t* out,
              cudaStream_t s, const char* name) {
  if (loop_dtype == NUMS_F64) return launch_uop<OPT, double>(L, a, out, s, false);
  if (loop_dtype == NUMS_F32) return launch_uop<OPT, float>(L, a, out, s, false);
  NUMS_UOP_FAIL(name);
}
template <template <typename> class OPT>
int uop_numeric(int loop_dtype, const Layout3& L, const nums_array_t* a, const nums_array_t* out,
                cudaStream_t s, const char* name) {
  if (loop_dtype == NUMS_I64) return launch_uop<OPT, int64_t>(L, a, out, s, false);
  if (loop_dtype == NUMS_I32) return launch_uop<OPT, int32_t>(L, a, out, s, false);
  return uop_float<OPT>(loop_dtype, L, a, out, s, name);
}
template <template <typename> class OPT>
int uop_intbool(int loop_dtype, const Layout3& L, const nums_array_t* a, const nums_array_t* out,
                cudaStream_t s, const char* name) {
  if (loop_dtype == NUMS_I64) return launch_uop<OPT, int64_t>(L, a, out, s, false);
  if (loop_dtype == NUMS_I32) return launch_uop<OPT, int32_t>(L, a, out, s, false);
  if (loop_dtype == NUMS_BOOL) return launch_uop<OPT, bool>(L, a, out, s, false);
  NUMS_UOP_FAIL(name);
}

// ---- n-ary sum ------------------------------------------------------------------------------------------
constexpr int kMaxSumInputs = 32;
struct SumInputs {
  const void* p[kMaxSumInputs];
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
sum_reduce_kernel(SumInputs in, int n_in, T* __restrict__ out, int64_t n, int accumulate) {
  const int64_t step = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += step) {
    T acc = accumulate ? out[i] : T(0);
    int first = 0;
    if (!accumulate) {
      acc = static_cast<const T*>(in.p[0])[i];
      first = 1;
    }
    for (int j = first; j < n_in; ++j) acc = acc + static_cast<const T*>(in.p[j])[i];
    out[i] = acc;
  }
}

template <typename T>
int launch_sum_reduce(int n_in, const void* const* arrs, int64_t numel, void* out, cudaStream_t s) {
  const unsigned grid = blocks_for(numel, kThreads, (int64_t)sm_count() * 16);
  for (int done = 0; done < n_in; done += kMaxSumInputs) {
    SumInputs in;
    int chunk = n_in - done < kMaxSumInputs ? n_in - done : kMaxSumInputs;
    for (int j = 0; j < chunk; ++j) in.p[j] = arrs[done + j];
    sum_reduce_kernel<T><<<grid, kThreads, 0, s>>>(in, chunk, static_cast<T*>(out), numel, done > 0);
    NUMS_LAUNCH_OK();
  }
  return NUMS_OK;
}

// ---- fills ------------------------------------------------------------------------------------------------
template <typename OS>
__global__ void __launch_bounds__(kThreads)
fill_kernel(DevLayout<1> L, OS* __restrict__ out, OS value, int64_t n) {
  const int64_t step = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += step) {
    int64_t off[1];
    unravel<1, uint64_t>(L, (uint64_t)i, off);
    out[off[0]] = value;
  }
}
template <typename OS>
__global__ void __launch_bounds__(kThreads)
arange_kernel(OS* __restrict__ out, double start, double step_v, int64_t n) {
  const int64_t step = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += step) {
    if constexpr (std::is_floating_point<OS>::value) out[i] = (OS)(start + (double)i * step_v);
    else out[i] = (OS)((int64_t)start + i * (int64_t)step_v);
  }
}
template <typename OS>
__global__ void __launch_bounds__(kThreads)
eye_kernel(OS* __restrict__ out, int64_t rows, int64_t cols) {
  const int64_t n = rows * cols, step = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += step)
    out[i] = (i / cols == i % cols) ? OS(1) : OS(0);
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

// dst[o][dst_index[p]][i] = src[o][src_index[p]][i] for every pair p: the per-pair loops of
// update_block_by_index / update_block_along_axis (numpy_compute.py:154-169) as ONE launch.  Both
// arrays are dense (outer, length, inner) views of the same element size; the caller removed
// duplicate destinations (last pair wins, as in the reference's sequential loop), so pairs are
// independent.  W = element width in bytes (1, 4, 8): the copy is bitwise.
template <typename W>
__global__ void __launch_bounds__(256)
scatter_axis_kernel(W* __restrict__ dst, const W* __restrict__ src, const int64_t* __restrict__ dst_index,
                    const int64_t* __restrict__ src_index, int64_t outer, int64_t dst_len, int64_t src_len,
                    int64_t inner, int64_t npairs) {
  const int64_t total = outer * npairs * inner;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e % inner;
    const int64_t p = (e / inner) % npairs;
    const int64_t o = e / (inner * npairs);
    dst[(o * dst_len + dst_index[p]) * inner + i] = src[(o * src_len + src_index[p]) * inner + i];
  }
}

}  // namespace
}  // namespace nums

extern "C" int nums_uop(int op, int loop_dtype, const nums_array_t* a, const nums_array_t* out,
                        void* stream) {
  using namespace nums;
  if (int rc = check_array(a, "uop a")) return rc;
  if (int rc = check_array(out, "uop out")) return rc;
  const nums_array_t* arrs[2] = {out, a};
  Layout3 L;
  if (int rc = build_layout(arrs, 2, &L)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define ALL(OPT, NAME) return uop_all<op::OPT>(loop_dtype, L, a, out, s, NAME)
#define NUM(OPT, NAME) return uop_numeric<op::OPT>(loop_dtype, L, a, out, s, NAME)
#define FLT(OPT, NAME) return uop_float<op::OPT>(loop_dtype, L, a, out, s, NAME)
#define INB(OPT, NAME) return uop_intbool<op::OPT>(loop_dtype, L, a, out, s, NAME)
  switch (op) {
    case NUMS_UOP_COPY: return uop_all<op::Copy>(loop_dtype, L, a, out, s, "copy", true);
    case NUMS_UOP_ABS: ALL(Abs, "absolute");
    case NUMS_UOP_NEGATIVE: NUM(Negative, "negative");
    case NUMS_UOP_POSITIVE: NUM(Positive, "positive");
    case NUMS_UOP_SIGN: NUM(Sign, "sign");
    case NUMS_UOP_SQRT: FLT(Sqrt, "sqrt");
    case NUMS_UOP_CBRT: FLT(Cbrt, "cbrt");
    case NUMS_UOP_SQUARE: ALL(Square, "square");
    case NUMS_UOP_RECIPROCAL: NUM(Reciprocal, "reciprocal");
    case NUMS_UOP_EXP: FLT(Exp, "exp");
    case NUMS_UOP_EXP2: FLT(Exp2, "exp2");
    case NUMS_UOP_EXPM1: FLT(Expm1, "expm1");
    case NUMS_UOP_LOG: FLT(Log, "log");
    case NUMS_UOP_LOG2: FLT(Log2, "log2");
    case NUMS_UOP_LOG10: FLT(Log10, "log10");
    case NUMS_UOP_LOG1P: FLT(Log1p, "log1p");
    case NUMS_UOP_SIN: FLT(Sin, "sin");
    case NUMS_UOP_COS: FLT(Cos, "cos");
    case NUMS_UOP_TAN: FLT(Tan, "tan");
    case NUMS_UOP_ARCSIN: FLT(Arcsin, "arcsin");
    case NUMS_UOP_ARCCOS: FLT(Arccos, "arccos");
    case NUMS_UOP_ARCTAN: FLT(Arctan, "arctan");
    case NUMS_UOP_SINH: FLT(Sinh, "sinh");
    case NUMS_UOP_COSH: FLT(Cosh, "cosh");
    case NUMS_UOP_TANH: FLT(Tanh, "tanh");
    case NUMS_UOP_ARCSINH: FLT(Arcsinh, "arcsinh");
    case NUMS_UOP_ARCCOSH: FLT(Arccosh, "arccosh");
    case NUMS_UOP_ARCTANH: FLT(Arctanh, "arctanh");
    case NUMS_UOP_FLOOR: ALL(Floor, "floor");
    case NUMS_UOP_CEIL: ALL(Ceil, "ceil");
    case NUMS_UOP_TRUNC: ALL(Trunc, "trunc");
    case NUMS_UOP_RINT: FLT(Rint, "rint");
    case NUMS_UOP_DEG2RAD: FLT(Deg2rad, "deg2rad");
    case NUMS_UOP_RAD2DEG: FLT(Rad2deg, "rad2deg");
    case NUMS_UOP_SPACING: FLT(Spacing, "spacing");
    case NUMS_UOP_ISNAN: ALL(Isnan, "isnan");
    case NUMS_UOP_ISINF: ALL(Isinf, "isinf");
    case NUMS_UOP_ISFINITE: ALL(Isfinite, "isfinite");
    case NUMS_UOP_SIGNBIT: NUM(Signbit, "signbit");
    case NUMS_UOP_LOGICAL_NOT: ALL(LogicalNot, "logical_not");
    case NUMS_UOP_INVERT: INB(Invert, "invert");
  }
#undef ALL
#undef NUM
#undef FLT
#undef INB
  NUMS_FAIL(NUMS_ERR_INVALID, "uop: unknown op id %d", op);
}

extern "C" int nums_sum_reduce(int n, const void* const* arrs_host, int dtype, int64_t numel,
                               void* out, void* stream) {
  using namespace nums;
  NUMS_REQUIRE(n >= 1 && arrs_host != nullptr, "sum_reduce: needs at least one input");
  NUMS_REQUIRE(numel >= 0, "sum_reduce: negative numel");
  if (numel == 0) return NUMS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case NUMS_F64: return launch_sum_reduce<double>(n, arrs_host, numel, out, s);
    case NUMS_F32: return launch_sum_reduce<float>(n, arrs_host, numel, out, s);
    case NUMS_I64: return launch_sum_reduce<int64_t>(n, arrs_host, numel, out, s);
    case NUMS_I32: return launch_sum_reduce<int32_t>(n, arrs_host, numel, out, s);
  }
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "sum_reduce: dtype %s", dtype_name(dtype));
}

extern "C" int nums_fill(const nums_array_t* out, double value, void* stream) {
  using namespace nums;
  if (int rc = check_array(out, "fill out")) return rc;
  const nums_array_t* arrs[1] = {out};
  Layout3 L;
  if (int rc = build_layout(arrs, 1, &L)) return rc;
  if (L.numel == 0) return NUMS_OK;
  DevLayout<1> D;
  D.ndim = L.ndim;
  for (int d = 0; d < L.ndim; ++d) {
    D.shape[d] = (uint32_t)L.shape[d];
    D.shape64[d] = L.shape[d];
    D.stride[0][d] = L.stride[0][d];
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = blocks_for(L.numel, kThreads, (int64_t)sm_count() * 16);
  return by_storage(out->dtype, [&](auto tag) -> int {
    using OS = decltype(tag);
    OS v;
    if constexpr (std::is_same<OS, uint8_t>::value) v = value != 0.0 ? 1 : 0;
    else v = (OS)value;
    fill_kernel<OS><<<grid, kThreads, 0, s>>>(D, static_cast<OS*>(out->data), v, L.numel);
    NUMS_LAUNCH_OK();
    return NUMS_OK;
  });
}

extern "C" int nums_arange(const nums_array_t* out, double start, double step, void* stream) {
  using namespace nums;
  if (int rc = check_array(out, "arange out")) return rc;
  NUMS_REQUIRE(out->ndim == 1 && (out->shape[0] <= 1 || out->stride[0] == 1), "arange: dense 1-D output required");
  NUMS_REQUIRE(out->dtype != NUMS_BOOL, "arange: bool output unsupported");
  const int64_t n = out->shape[0];
  if (n == 0) return NUMS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = blocks_for(n, kThreads, (int64_t)sm_count() * 16);
  return by_storage(out->dtype, [&](auto tag) -> int {
    using OS = decltype(tag);
    arange_kernel<OS><<<grid, kThreads, 0, s>>>(static_cast<OS*>(out->data), start, step, n);
    NUMS_LAUNCH_OK();
    return NUMS_OK;
  });
}

extern "C" int nums_eye(const nums_array_t* out, void* stream) {
  using namespace nums;
  if (int rc = check_array(out, "eye out")) return rc;
  NUMS_REQUIRE(out->ndim == 2, "eye: 2-D output required");
  NUMS_REQUIRE((out->shape[1] <= 1 || out->stride[1] == 1) && (out->shape[0] <= 1 || out->stride[0] == out->shape[1]),
               "eye: dense output required");
  const int64_t n = out->shape[0] * out->shape[1];
  if (n == 0) return NUMS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = blocks_for(n, kThreads, (int64_t)sm_count() * 16);
  return by_storage(out->dtype, [&](auto tag) -> int {
    using OS = decltype(tag);
    eye_kernel<OS><<<grid, kThreads, 0, s>>>(static_cast<OS*>(out->data), out->shape[0], out->shape[1]);
    NUMS_LAUNCH_OK();
    return NUMS_OK;
  });
}

extern "C" int nums_scatter_axis(int elem_size, int64_t outer, int64_t dst_len, int64_t src_len, int64_t inner,
                                 int64_t npairs, const int64_t* dst_index, const int64_t* src_index, void* dst,
                                 const void* src, void* stream) {
  using namespace nums;
  NUMS_REQUIRE(outer >= 0 && dst_len >= 0 && src_len >= 0 && inner >= 0 && npairs >= 0, "scatter_axis: negative extent");
  const int64_t total = outer * npairs * inner;
  if (total == 0) return NUMS_OK;
  NUMS_REQUIRE(dst && src && dst_index && src_index, "scatter_axis: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = blocks_for(total, 256, (int64_t)sm_count() * 16);
  switch (elem_size) {
    case 1:
      scatter_axis_kernel<uint8_t><<<grid, 256, 0, s>>>(static_cast<uint8_t*>(dst), static_cast<const uint8_t*>(src), dst_index,
                                                        src_index, outer, dst_len, src_len, inner, npairs);
      break;
    case 4:
      scatter_axis_kernel<uint32_t><<<grid, 256, 0, s>>>(static_cast<uint32_t*>(dst), static_cast<const uint32_t*>(src),
                                                         dst_index, src_index, outer, dst_len, src_len, inner, npairs);
      break;
    case 8:
      scatter_axis_kernel<uint64_t><<<grid, 256, 0, s>>>(static_cast<uint64_t*>(dst), static_cast<const uint64_t*>(src),
                                                         dst_index, src_index, outer, dst_len, src_len, inner, npairs);
      break;
    default:
      NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "scatter_axis: element size %d", elem_size);
  }
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}
