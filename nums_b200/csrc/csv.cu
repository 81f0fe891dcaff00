// Delimited text -> dense block, on the device.  Replaces the per-chunk parser of the reference's CSV
// ingest, read_csv_block (nums/core/systems/filesystem.py:157-212): lines are split at the
// delimiter and every field goes through the dtype's converter (:160-190).  The host side
// (cuda_compute.read_csv_block) decides which whole lines belong to the chunk exactly as the
// reference does (:196-211) and hands the byte range [first, stop) of those lines to two kernels:
//
//   nums_csv_index : one pass over the bytes; per 8 KiB tile the number of field separators
//                    (delimiter or '\n') and of line ends, then an exclusive scan over the tiles.
//                    The totals come back to the host, which sizes the output (the reference returns
//                    the shape next to the block for the same reason, :212).
//   nums_csv_parse : the same tiles again; the field starts of a tile are compacted into shared
//                    memory with a block-wide scan, which also gives every field its global index
//                    (row = index / columns, column = index % columns); consecutive threads then
//                    convert consecutive fields (csv_parse.cuh: correctly rounded decimal -> double).
//
// Both passes are byte-bound integer work: 16-byte loads, no tensor cores.  Ragged rows, malformed
// literals and unsupported ones (hex floats) are reported through a status word with the byte
// offset of the first offender, so the host can raise the reference's ValueError.
#include "common.cuh"
#include "csv_parse.cuh"

namespace nums {
namespace {

constexpr int kTileBytes = 8192;
constexpr int kCsvThreads = 256;
constexpr int kBytesPerThread = kTileBytes / kCsvThreads;   // 32
constexpr int kMaxFieldsPerTile = kTileBytes;               // every byte a separator, worst case

struct CsvRange {
  const uint8_t* text;   // 16-byte aligned; readable up to `stop` rounded up to a multiple of 32
  int64_t first, stop;   // bytes [first, stop) hold whole lines; the last one may lack its '\n'
  int64_t origin;        // first rounded down to a multiple of 16: tiles start here, so loads are aligned
  uint8_t delimiter;
};

__device__ __forceinline__ bool is_separator(uint8_t c, uint8_t delimiter) { return c == delimiter || c == '\n'; }

// The 32 bytes a thread owns in a tile, as two 16-byte loads.
struct Chunk {
  uint8_t b[kBytesPerThread];
  __device__ __forceinline__ void load(const uint8_t* p) {
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    *reinterpret_cast<uint4*>(b) = lo;
    *reinterpret_cast<uint4*>(b + 16) = hi;
  }
};

// Per tile: {separators, newlines}; also flags a '\r' that is not followed by '\n' (the reference's
// text-mode reader would split the line there).
__global__ void __launch_bounds__(kCsvThreads)
csv_count_kernel(CsvRange r, int64_t tiles, int64_t* __restrict__ tile_seps, int64_t* __restrict__ tile_lines,
                 int64_t* __restrict__ summary) {
  __shared__ int warp_seps[kCsvThreads / 32], warp_lines[kCsvThreads / 32];
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t base = r.origin + tile * kTileBytes + (int64_t)threadIdx.x * kBytesPerThread;
    int seps = 0, lines = 0;
    bool lone_cr = false;
    if (base < r.stop && base + kBytesPerThread > r.first) {
      alignas(16) Chunk ch;
      ch.load(r.text + base);
#pragma unroll
      for (int i = 0; i < kBytesPerThread; ++i) {
        const int64_t pos = base + i;
        if (pos < r.first || pos >= r.stop) continue;
        const uint8_t c = ch.b[i];
        seps += is_separator(c, r.delimiter);
        lines += c == '\n';
        if (c == '\r' && pos + 1 < r.stop) {
          const uint8_t next = i + 1 < kBytesPerThread ? ch.b[i + 1] : r.text[pos + 1];
          if (next != '\n') lone_cr = true;
        }
      }
    }
    if (lone_cr) atomicMax(reinterpret_cast<unsigned long long*>(summary + 2), (unsigned long long)NUMS_CSV_UNSUPPORTED);
    for (int o = 16; o > 0; o >>= 1) {
      seps += __shfl_xor_sync(0xffffffffu, seps, o);
      lines += __shfl_xor_sync(0xffffffffu, lines, o);
    }
    if ((threadIdx.x & 31) == 0) {
      warp_seps[threadIdx.x >> 5] = seps;
      warp_lines[threadIdx.x >> 5] = lines;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0, l = 0;
      for (int w = 0; w < kCsvThreads / 32; ++w) {
        s += warp_seps[w];
        l += warp_lines[w];
      }
      tile_seps[tile] = s;
      tile_lines[tile] = l;
    }
    __syncthreads();
  }
}

// Exclusive scan of the per-tile separator counts (in place) and the totals:
// summary[0] = lines (a last line without '\n' included), summary[1] = fields.
__global__ void __launch_bounds__(1024)
csv_scan_kernel(CsvRange r, int64_t tiles, int64_t* __restrict__ tile_seps, const int64_t* __restrict__ tile_lines,
                int64_t* __restrict__ summary) {
  __shared__ int64_t partial[1024];
  __shared__ int64_t line_partial[1024];
  const int t = threadIdx.x;
  const int64_t per = (tiles + 1023) / 1024;
  const int64_t lo = t * per, hi = min(tiles, lo + per);
  int64_t s = 0, l = 0;
  for (int64_t i = lo; i < hi; ++i) {
    s += tile_seps[i];
    l += tile_lines[i];
  }
  partial[t] = s;
  line_partial[t] = l;
  __syncthreads();
  if (t == 0) {
    int64_t run = 0, lines = 0;
    for (int i = 0; i < 1024; ++i) {
      const int64_t v = partial[i];
      partial[i] = run;
      run += v;
      lines += line_partial[i];
    }
    const bool open_last = r.stop > r.first && r.text[r.stop - 1] != '\n';   // final line without terminator
    summary[0] = lines + (open_last ? 1 : 0);
    summary[1] = run + (open_last ? 1 : 0);
    if (open_last && r.text[r.stop - 1] == r.delimiter) {   // "a,b," at end of file: the empty last field is a ValueError
      atomicMax(reinterpret_cast<unsigned long long*>(summary + 2), (unsigned long long)NUMS_CSV_INVALID);
      atomicMin(reinterpret_cast<unsigned long long*>(summary + 3), (unsigned long long)r.stop);
    }
  }
  __syncthreads();
  int64_t run = partial[t];
  for (int64_t i = lo; i < hi; ++i) {
    const int64_t v = tile_seps[i];
    tile_seps[i] = run;
    run += v;
  }
}

__device__ __forceinline__ void report(int64_t* summary, int code, int64_t offset) {
  atomicMax(reinterpret_cast<unsigned long long*>(summary + 2), (unsigned long long)code);
  atomicMin(reinterpret_cast<unsigned long long*>(summary + 3), (unsigned long long)offset);
}

template <typename T> struct Convert;
template <> struct Convert<double> {      // floatconv, filesystem.py:163-167
  static __device__ int run(const uint8_t* p, int n, double* out) { return csv::parse_float(p, n, out); }
};
template <> struct Convert<float> {       // float(x), then np.array(..., dtype=float32) rounds once more
  static __device__ int run(const uint8_t* p, int n, float* out) {
    double d;
    const int st = csv::parse_float(p, n, &d);
    *out = (float)d;
    return st;
  }
};
template <> struct Convert<int64_t> {     // np.int64(x), :173-174
  static __device__ int run(const uint8_t* p, int n, int64_t* out) { return csv::parse_int64(p, n, out); }
};
template <> struct Convert<int32_t> {     // int(float(x)), :175-176; out-of-range values overflow in np.array
  static __device__ int run(const uint8_t* p, int n, int32_t* out) {
    double d;
    const int st = csv::parse_float(p, n, &d);
    if (st != csv::FIELD_OK) return st;
    if (!(d > -2147483649.0 && d < 2147483648.0)) return csv::FIELD_INVALID;   // also NaN / inf
    *out = (int32_t)d;                    // truncation toward zero, like int()
    return csv::FIELD_OK;
  }
};
template <> struct Convert<bool> {        // bool(int(x)), :169-170
  static __device__ int run(const uint8_t* p, int n, bool* out) {
    int64_t v;
    const int st = csv::parse_int64(p, n, &v);
    *out = v != 0;
    return st;
  }
};

template <typename T>
__global__ void __launch_bounds__(kCsvThreads)
csv_parse_kernel(CsvRange r, int64_t tiles, const int64_t* __restrict__ tile_offsets, int64_t rows, int64_t cols,
                 T* __restrict__ out, int64_t* __restrict__ summary) {
  // starts[k] = byte offset (relative to the tile) of the k-th field that STARTS in this tile
  __shared__ uint16_t starts[kMaxFieldsPerTile];
  __shared__ int warp_total[kCsvThreads / 32];
  __shared__ int tile_fields;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t tile_base = r.origin + tile * kTileBytes;
    const int64_t base = tile_base + (int64_t)threadIdx.x * kBytesPerThread;
    // A field starts at `first` and right after every separator (unless that separator is the last byte).
    uint32_t mask = 0;   // bit i: a field starts at base + i
    if (base < r.stop && base + kBytesPerThread > r.first) {
      alignas(16) Chunk ch;
      ch.load(r.text + base);
      uint8_t prev = base > r.first ? r.text[base - 1] : 0;
#pragma unroll
      for (int i = 0; i < kBytesPerThread; ++i) {
        const int64_t pos = base + i;
        const bool start = pos >= r.first && pos < r.stop && (pos == r.first || is_separator(prev, r.delimiter));
        mask |= (uint32_t)start << i;
        prev = ch.b[i];
      }
    }
    const int mine = __popc(mask);
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_total[warp] = incl;
    __syncthreads();
    int before = incl - mine;
    for (int w = 0; w < warp; ++w) before += warp_total[w];
    if (threadIdx.x == kCsvThreads - 1) tile_fields = before + mine;
    for (uint32_t m = mask; m; m &= m - 1) starts[before++] = (uint16_t)(threadIdx.x * kBytesPerThread + __ffs(m) - 1);
    __syncthreads();
    int64_t first_index = 0;
    if (tile > 0) {
      // fields that start before this tile: the one at `first` plus one per separator before
      // tile_base - 1 (a separator AT tile_base - 1 announces a field of this tile)
      first_index = tile_offsets[tile] + 1 - (is_separator(r.text[tile_base - 1], r.delimiter) ? 1 : 0);
    }
    const int count = tile_fields;
    for (int k = threadIdx.x; k < count; k += kCsvThreads) {
      const int64_t begin = tile_base + starts[k];
      int64_t end = begin;
      while (end < r.stop && !is_separator(r.text[end], r.delimiter)) ++end;
      const bool line_end = end >= r.stop || r.text[end] == '\n';
      int64_t trimmed = end;
      if (line_end && trimmed > begin && r.text[trimmed - 1] == '\r') --trimmed;   // line.strip("\r\n")
      const int64_t index = first_index + k;
      const int64_t row = index / cols, col = index - row * cols;
      if (row >= rows || line_end != (col == cols - 1)) {
        report(summary, NUMS_CSV_RAGGED, begin);
        continue;
      }
      if (trimmed - begin > (1 << 20)) {
        report(summary, NUMS_CSV_INVALID, begin);
        continue;
      }
      T value;
      const int st = Convert<T>::run(r.text + begin, (int)(trimmed - begin), &value);
      if (st == csv::FIELD_OK) out[row * cols + col] = value;
      else report(summary, st == csv::FIELD_INVALID ? NUMS_CSV_INVALID : NUMS_CSV_UNSUPPORTED, begin);
    }
    __syncthreads();
  }
}

inline int64_t tile_count(int64_t first, int64_t stop) {
  const int64_t origin = first & ~(int64_t)15;
  return stop > first ? (stop - origin + kTileBytes - 1) / kTileBytes : 0;
}

}  // namespace
}  // namespace nums

extern "C" int nums_csv_index(const void* text, int64_t first, int64_t stop, int delimiter, int64_t* summary,
                              void* ws, size_t ws_bytes, void* stream) {
  using namespace nums;
  NUMS_REQUIRE(text && summary && first >= 0 && stop >= first, "csv_index: bad range");
  NUMS_REQUIRE((reinterpret_cast<uintptr_t>(text) & 15u) == 0, "csv_index: text must be 16-byte aligned");
  NUMS_REQUIRE(delimiter > 0 && delimiter < 256 && delimiter != '\n' && delimiter != '\r', "csv_index: bad delimiter");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t tiles = tile_count(first, stop);
  NUMS_NEED_WS((size_t)(2 * tiles + 2) * sizeof(int64_t), ws_bytes);
  // summary = {lines, fields, status, offset of the first offending field}
  const int64_t init[4] = {0, 0, NUMS_CSV_OK, INT64_MAX};
  NUMS_CUDA_OK(cudaMemcpyAsync(summary, init, sizeof(init), cudaMemcpyHostToDevice, s));
  if (tiles == 0) return NUMS_OK;
  int64_t* tile_seps = static_cast<int64_t*>(ws);
  int64_t* tile_lines = tile_seps + tiles;
  CsvRange r{static_cast<const uint8_t*>(text), first, stop, first & ~(int64_t)15, (uint8_t)delimiter};
  const unsigned grid = (unsigned)(tiles < (int64_t)sm_count() * 8 ? tiles : (int64_t)sm_count() * 8);
  csv_count_kernel<<<grid, kCsvThreads, 0, s>>>(r, tiles, tile_seps, tile_lines, summary);
  NUMS_LAUNCH_OK();
  csv_scan_kernel<<<1, 1024, 0, s>>>(r, tiles, tile_seps, tile_lines, summary);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

extern "C" int nums_csv_parse(const void* text, int64_t first, int64_t stop, int delimiter, int dtype, int64_t rows,
                              int64_t cols, void* out, int64_t* summary, const void* ws, void* stream) {
  using namespace nums;
  NUMS_REQUIRE(text && out && summary && ws && first >= 0 && stop >= first, "csv_parse: bad range");
  NUMS_REQUIRE(rows >= 0 && cols >= 1, "csv_parse: bad shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t tiles = tile_count(first, stop);
  if (tiles == 0 || rows == 0) return NUMS_OK;
  const int64_t* offsets = static_cast<const int64_t*>(ws);   // left there by nums_csv_index
  CsvRange r{static_cast<const uint8_t*>(text), first, stop, first & ~(int64_t)15, (uint8_t)delimiter};
  const unsigned grid = (unsigned)(tiles < (int64_t)sm_count() * 8 ? tiles : (int64_t)sm_count() * 8);
  switch (dtype) {
    case NUMS_F64:
      csv_parse_kernel<double><<<grid, kCsvThreads, 0, s>>>(r, tiles, offsets, rows, cols, static_cast<double*>(out), summary);
      break;
    case NUMS_F32:
      csv_parse_kernel<float><<<grid, kCsvThreads, 0, s>>>(r, tiles, offsets, rows, cols, static_cast<float*>(out), summary);
      break;
    case NUMS_I64:
      csv_parse_kernel<int64_t><<<grid, kCsvThreads, 0, s>>>(r, tiles, offsets, rows, cols, static_cast<int64_t*>(out), summary);
      break;
    case NUMS_I32:
      csv_parse_kernel<int32_t><<<grid, kCsvThreads, 0, s>>>(r, tiles, offsets, rows, cols, static_cast<int32_t*>(out), summary);
      break;
    case NUMS_BOOL:
      csv_parse_kernel<bool><<<grid, kCsvThreads, 0, s>>>(r, tiles, offsets, rows, cols, static_cast<bool*>(out), summary);
      break;
    default:
      NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "csv_parse: dtype %s", dtype_name(dtype));
  }
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}
