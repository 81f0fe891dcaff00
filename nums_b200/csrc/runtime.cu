// Library plumbing: thread-local error text, workspace requests, array/layout validation.
#include <atomic>
#include "common.cuh"

namespace nums {

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local char g_error[512] = "";
static thread_local size_t g_ws_request = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void set_workspace_request(size_t bytes) { g_ws_request = bytes; }

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}

int check_array(const nums_array_t* a, const char* what) {
  NUMS_REQUIRE(a != nullptr, "%s: null array descriptor", what);
  NUMS_REQUIRE(a->ndim >= 0 && a->ndim <= NUMS_MAX_DIMS, "%s: ndim %d out of range", what, a->ndim);
  NUMS_REQUIRE(dtype_size(a->dtype) > 0, "%s: unknown dtype %d", what, a->dtype);
  int64_t n = 1;
  for (int i = 0; i < a->ndim; ++i) {
    NUMS_REQUIRE(a->shape[i] >= 0, "%s: negative extent on axis %d", what, i);
    NUMS_REQUIRE(a->stride[i] >= 0, "%s: negative stride on axis %d", what, i);
    n *= a->shape[i];
  }
  NUMS_REQUIRE(n == 0 || a->data != nullptr, "%s: null data pointer", what);
  return NUMS_OK;
}

int64_t array_numel(const nums_array_t* a) {
  int64_t n = 1;
  for (int i = 0; i < a->ndim; ++i) n *= a->shape[i];
  return n;
}

int build_layout(const nums_array_t* const* arrs, int n, Layout3* L) {
  const nums_array_t* out = arrs[0];
  int nd = out->ndim;
  int64_t shape[NUMS_MAX_DIMS];
  int64_t stride[3][NUMS_MAX_DIMS];
  for (int d = 0; d < nd; ++d) {
    shape[d] = out->shape[d];
    stride[0][d] = out->stride[d];
  }
  for (int o = 1; o < n; ++o) {
    const nums_array_t* a = arrs[o];
    NUMS_REQUIRE(a->ndim <= nd, "operand %d has more axes (%d) than the output (%d)", o, a->ndim, nd);
    int shift = nd - a->ndim;
    for (int d = 0; d < nd; ++d) {
      if (d < shift) {
        stride[o][d] = 0;
        continue;
      }
      int64_t ext = a->shape[d - shift];
      if (ext == shape[d] && ext != 1) stride[o][d] = a->stride[d - shift];
      else if (ext == 1) stride[o][d] = 0;
      else NUMS_FAIL(NUMS_ERR_INVALID, "operand %d: extent %lld on axis %d does not broadcast to %lld",
                     o, (long long)ext, d, (long long)shape[d]);
    }
  }
  // Drop size-1 axes, then merge neighbours that are jointly dense for every operand.
  int m = 0;
  int64_t numel = 1;
  for (int d = 0; d < nd; ++d) {
    numel *= shape[d];
    if (shape[d] == 1) continue;
    if (m > 0) {
      bool mergeable = true;
      for (int o = 0; o < n; ++o)
        if (L->stride[o][m - 1] != stride[o][d] * shape[d]) mergeable = false;
      if (mergeable) {
        L->shape[m - 1] *= shape[d];
        for (int o = 0; o < n; ++o) L->stride[o][m - 1] = stride[o][d];
        continue;
      }
    }
    L->shape[m] = shape[d];
    for (int o = 0; o < n; ++o) L->stride[o][m] = stride[o][d];
    ++m;
  }
  if (m == 0) {  // scalar (or all-ones) iteration space
    L->shape[0] = 1;
    for (int o = 0; o < n; ++o) L->stride[o][0] = 0;
    m = 1;
  }
  L->ndim = m;
  L->numel = numel;
  return NUMS_OK;
}

bool layout_contiguous(const Layout3& L, int o) {
  if (L.numel <= 1) return true;
  return L.ndim == 1 && L.stride[o][0] == 1;
}
bool layout_scalar(const Layout3& L, int o) {
  for (int d = 0; d < L.ndim; ++d)
    if (L.stride[o][d] != 0) return false;
  return true;
}

}  // namespace nums

extern "C" {
int nums_abi_version(void) { return NUMS_ABI_VERSION; }
const char* nums_last_error(void) { return nums::g_error; }
size_t nums_last_workspace_request(void) { return nums::g_ws_request; }
int nums_sm_count(void) { return nums::sm_count(); }
uint64_t nums_launch_count(void) { return nums::g_launches.load(std::memory_order_relaxed); }
}
