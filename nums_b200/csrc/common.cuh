// Shared helpers for libnumscuda (sm_100a).  Internal header, not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "../../include/nums_cuda.h"

namespace nums {

// ---- error plumbing --------------------------------------------------------------------
void set_error(const char* fmt, ...);
void set_workspace_request(size_t bytes);
void count_launch();
int sm_count();

#define NUMS_FAIL(code, ...)      \
  do {                            \
    ::nums::set_error(__VA_ARGS__); \
    return (code);                \
  } while (0)

#define NUMS_REQUIRE(cond, ...)                           \
  do {                                                    \
    if (!(cond)) NUMS_FAIL(NUMS_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define NUMS_CUDA_OK(expr)                                                        \
  do {                                                                            \
    cudaError_t e__ = (expr);                                                     \
    if (e__ != cudaSuccess)                                                       \
      NUMS_FAIL(NUMS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                __FILE__, __LINE__);                                              \
  } while (0)

// Every kernel launch of the library goes through this macro: it bumps the process-wide launch
// counter (nums_launch_count) and surfaces launch-configuration errors.
#define NUMS_LAUNCH_OK()                  \
  do {                                    \
    ::nums::count_launch();               \
    NUMS_CUDA_OK(cudaGetLastError());     \
  } while (0)

#define NUMS_NEED_WS(need, have)                                                   \
  do {                                                                             \
    if ((size_t)(need) > (size_t)(have)) {                                         \
      ::nums::set_workspace_request((size_t)(need));                               \
      NUMS_FAIL(NUMS_ERR_WORKSPACE, "workspace too small: need %zu bytes, have %zu", \
                (size_t)(need), (size_t)(have));                                   \
    }                                                                              \
  } while (0)

// ---- dtypes --------------------------------------------------------------------------------
__host__ __device__ inline int dtype_size(int dt) {
  switch (dt) {
    case NUMS_BOOL: return 1;
    case NUMS_I32: return 4;
    case NUMS_I64: return 8;
    case NUMS_F32: return 4;
    case NUMS_F64: return 8;
  }
  return 0;
}
inline const char* dtype_name(int dt) {
  switch (dt) {
    case NUMS_BOOL: return "bool";
    case NUMS_I32: return "int32";
    case NUMS_I64: return "int64";
    case NUMS_F32: return "float32";
    case NUMS_F64: return "float64";
  }
  return "?";
}
template <typename T> struct dtype_of;
template <> struct dtype_of<bool> { static constexpr int value = NUMS_BOOL; };
template <> struct dtype_of<int32_t> { static constexpr int value = NUMS_I32; };
template <> struct dtype_of<int64_t> { static constexpr int value = NUMS_I64; };
template <> struct dtype_of<float> { static constexpr int value = NUMS_F32; };
template <> struct dtype_of<double> { static constexpr int value = NUMS_F64; };

// C-style conversion with NumPy's "x != 0" rule for casts to bool.
template <typename To, typename From>
__host__ __device__ __forceinline__ To convert(From v) {
  if constexpr (std::is_same<To, bool>::value) return v != From(0);
  else return static_cast<To>(v);
}

// Load element `idx` of an array stored as `dt`, converted to T.
template <typename T>
__device__ __forceinline__ T load_as(const void* p, int dt, int64_t idx) {
  switch (dt) {
    case NUMS_F64: return convert<T>(static_cast<const double*>(p)[idx]);
    case NUMS_F32: return convert<T>(static_cast<const float*>(p)[idx]);
    case NUMS_I64: return convert<T>(static_cast<const int64_t*>(p)[idx]);
    case NUMS_I32: return convert<T>(static_cast<const int32_t*>(p)[idx]);
    default: return convert<T>(static_cast<const uint8_t*>(p)[idx] != 0);
  }
}

// Storage type for bool results is one byte holding 0/1.
template <typename T> struct storage_of { using type = T; };
template <> struct storage_of<bool> { using type = uint8_t; };

template <typename T, int N>
struct alignas(sizeof(T) * N) Vec {
  T v[N];
};

// ---- broadcast / stride bookkeeping (host) -------------------------------------------------------
struct Layout3 {   // up to 3 operands sharing an iteration space (out, a, b)
  int ndim;
  int64_t shape[NUMS_MAX_DIMS];
  int64_t stride[3][NUMS_MAX_DIMS];
  int64_t numel;
};

// Broadcast `arrs[1..n)` against arrs[0] (the output) and collapse mergeable axes.
// Returns 0 or a negative status.
int build_layout(const nums_array_t* const* arrs, int n, Layout3* L);
bool layout_contiguous(const Layout3& L, int operand);  // dense row-major over the collapsed space
bool layout_scalar(const Layout3& L, int operand);      // every stride is 0

int check_array(const nums_array_t* a, const char* what);
int64_t array_numel(const nums_array_t* a);

// ---- kernel-side layout ---------------------------------------------------------------------------
template <int NOPS>
struct DevLayout {
  int ndim;
  uint32_t shape[NUMS_MAX_DIMS];      // used by the 32-bit index path
  int64_t shape64[NUMS_MAX_DIMS];
  int64_t stride[NOPS][NUMS_MAX_DIMS];
};

template <int NOPS, typename IDX>
__device__ __forceinline__ void unravel(const DevLayout<NOPS>& L, IDX linear, int64_t (&off)[NOPS]) {
#pragma unroll
  for (int o = 0; o < NOPS; ++o) off[o] = 0;
#pragma unroll 1
  for (int d = L.ndim - 1; d >= 0; --d) {
    IDX dim = (sizeof(IDX) == 4) ? (IDX)L.shape[d] : (IDX)L.shape64[d];
    IDX q = linear / dim;
    IDX r = linear - q * dim;
    linear = q;
#pragma unroll
    for (int o = 0; o < NOPS; ++o) off[o] += (int64_t)r * L.stride[o][d];
  }
}

inline unsigned blocks_for(int64_t work_items, int per_block, int64_t cap = 0x7fffffffLL) {
  int64_t b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (unsigned)b;
}

// ---- warp / block reductions -------------------------------------------------------------------------
template <typename T, typename F>
__device__ __forceinline__ T warp_reduce(T v, F f) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = f(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace nums
