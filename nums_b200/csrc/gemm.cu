// Dense block contractions -- replaces np.tensordot(a1, a2, axes) at
// nums/core/systems/numpy_compute.py:231-232 (OpenBLAS dgemm/dgemv in the reference).
//
//   * f64 GEMM: FP64 tensor pipe.  On sm_100a every mma.sync f64 shape lowers to DMMA.8x8x4
//     (checked with cuobjdump), so the kernels are written directly against m8n8k4 fragments:
//     128x128x32 CTA tiles, 8 MMA warps of 32x64, a 3-stage shared-memory ring padded so that
//     every fragment read is bank-conflict free, optional split-K for small outputs with long
//     contractions (X^T X, the LR Hessian), grouped launches over many output blocks whose
//     tiles accumulate a whole k-chain of terms in registers.  Operands may be stored
//     transposed (BlockArray.T is lazy, base.py:72-85), handled by the shared-memory layout,
//     not by a copy.  Two feeds for the same tile code: dgemm_dmma_tma_kernel (default: a
//     producer warp issues two tensor-map TMA copies per k-tile, mbarrier hand-off, 36.3 TFLOP/s)
//     and dgemm_dmma_kernel (cp.async issued by the MMA warps, 33.4 TFLOP/s; the fallback when
//     cuTensorMapEncodeTiled is unavailable, and NUMS_GEMM_FEED=cpasync for comparisons).
//   * matrix-vector / vector-vector forms (BlockArray._matvec / _vecdot, blockarray.py:475-580):
//     HBM-bound streaming kernels.
//   * everything else (f32, exact integer tensordot from tests/core/array/test_bop.py:38-42,
//     unaligned f64): a plain shared-memory tiled kernel.
#include <cuda.h>   // CUtensorMap types; the encoder itself is fetched through cudaGetDriverEntryPoint
#include <cstdlib>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace nums {
namespace {

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ======================================================================================
// FP64 DMMA GEMM
// ======================================================================================
constexpr int BM = 128, BN = 128, BK = 32;
constexpr int kStages = 3;
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
  // not volatile: the scheduler may interleave the k-loop's cp.async address arithmetic with the MMAs
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// Streams (ROWS x COLS) tiles of a row-major matrix into shared memory, 16 bytes per cp.async,
// zero-filling everything out of range.  One of the two tile axes is the contraction axis and
// advances by BK per k-tile (K_IS_ROW: tile rows run along k); the other is fixed for the CTA.
// All per-thread addressing is precomputed once per term, so a k-tile costs one pointer bump,
// one byte-count and ITERS cp.async per thread.  Requires an even pitch and a 16-byte aligned
// base so that every in-range chunk is aligned.
template <int ROWS, int COLS, int PITCH, bool K_IS_ROW>
struct TileLoader {
  static constexpr int kChunksPerRow = COLS / 2;
  static constexpr int kRowStep = kGemmThreads / kChunksPerRow;  // tile rows covered per sweep
  static constexpr int ITERS = ROWS / kRowStep;
  static_assert(kGemmThreads % kChunksPerRow == 0 && ROWS % kRowStep == 0, "tile must divide evenly over the CTA");

  const double* ptr;   // element (r_base, c) of the current tile
  int64_t sweep;       // kRowStep * ld
  int64_t k_step;      // pointer bump per k-tile
  int smem_off;
  int r_base, c;
  uint32_t fixed_ok;   // K_IS_ROW: 0/8/16 bytes valid along the fixed (column) axis; else row-valid bit mask

  __device__ __forceinline__ void init(const double* src, int64_t ld, int64_t fixed0, int64_t fixed_max, int64_t k0) {
    r_base = threadIdx.x / kChunksPerRow;
    c = (threadIdx.x % kChunksPerRow) * 2;
    smem_off = r_base * PITCH + c;
    sweep = (int64_t)kRowStep * ld;
    if (K_IS_ROW) {
      const int64_t gc = fixed0 + c;
      fixed_ok = gc + 1 < fixed_max ? 16u : (gc < fixed_max ? 8u : 0u);
      ptr = src + (k0 + r_base) * ld + (fixed_ok ? gc : 0);
      k_step = (int64_t)BK * ld;
    } else {
      fixed_ok = 0;
#pragma unroll
      for (int it = 0; it < ITERS; ++it)
        if (fixed0 + r_base + it * kRowStep < fixed_max) fixed_ok |= 1u << it;
      ptr = src + (fixed0 + r_base) * ld + k0 + c;
      k_step = BK;
    }
  }
  // One of the ITERS cp.async of a tile (k_left = k_hi - k_cur: valid contraction indices from the
  // start of this tile).  The pieces are issued between the MMA steps of the previous tile so
  // that their issue cost hides behind the other warp's DMMA stream.
  __device__ __forceinline__ void load_piece(double* smem, int it, int64_t k_left, const double* safe) {
    double* dst = smem + smem_off + it * kRowStep * PITCH;
    int bytes;
    if (K_IS_ROW) {
      bytes = (r_base + it * kRowStep < k_left) ? (int)fixed_ok : 0;
    } else {
      const int64_t left = k_left - c;
      const int col_bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);
      bytes = ((fixed_ok >> it) & 1u) ? col_bytes : 0;
    }
    cp_async16(dst, bytes ? ptr + it * sweep : safe, bytes);
  }
  __device__ __forceinline__ void advance() { ptr += k_step; }
  __device__ __forceinline__ void load(double* smem, int64_t k_left, const double* safe) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) load_piece(smem, it, k_left, safe);
    advance();
  }
};

// One contraction term: C += op(A)[M x K] . op(B)[K x N].
struct GemmTerm {
  const double* A;
  const double* B;
  int64_t lda, ldb, K;
};
// One output block: C = Cin + sum over `term_count` terms (the k-chain of BlockArray._tensordot,
// blockarray.py:460-472, accumulated in registers instead of through `add` kernels).
struct GemmProblem {
  double* C;
  const double* Cin;  // optional addend (may alias C)
  int64_t ldc, ldcin, M, N;
  int32_t tiles_m, tiles_n;
  int32_t term_begin, term_count;
  int32_t tile_begin;  // first CTA index of this problem
  int32_t pad_;
};

struct GemmParams {
  GemmProblem single;           // used when nproblems == 1 (no table upload)
  GemmTerm single_term;
  const GemmProblem* problems;  // device tables when nproblems > 1 or term_count > 1
  const GemmTerm* terms;
  int32_t nproblems;
  int32_t use_table;
  int64_t k_per_split;   // split-K (single problem, single term only); multiple of BK
  int64_t split_stride;  // elements between split-K partial outputs (M*N) or 0
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(kGemmThreads, 1)
dgemm_dmma_kernel(GemmParams p) {
  extern __shared__ __align__(16) double smem[];
  using AT = ATile<TA>;
  using BT = BTile<TB>;
  double* sA = smem;
  double* sB = smem + kStages * AT::doubles;

  // ---- which problem / tile does this CTA own? ---------------------------------------------------
  GemmProblem prob;
  if (p.use_table) {
    int lo = 0, hi = p.nproblems - 1;
    while (lo < hi) {  // last problem whose tile_begin <= blockIdx.x
      const int mid = (lo + hi + 1) >> 1;
      if (p.problems[mid].tile_begin <= (int)blockIdx.x) lo = mid;
      else hi = mid - 1;
    }
    prob = p.problems[lo];
  } else {
    prob = p.single;
  }
  // Tile order inside a problem: groups of 8 tile-rows sweep the columns, so concurrently running
  // CTAs share A row panels and B column panels in L2.
  const int tile = (int)blockIdx.x - prob.tile_begin;
  constexpr int kGroup = 8;
  const int tiles_per_group = kGroup * prob.tiles_n;
  const int group = tile / tiles_per_group;
  const int first_m = group * kGroup;
  const int group_rows = min(prob.tiles_m - first_m, kGroup);
  const int tm = first_m + (tile % tiles_per_group) % group_rows;
  const int tn = (tile % tiles_per_group) / group_rows;
  const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;
  const int64_t M = prob.M, N = prob.N;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp & 3) * 32;   // 4 warps along M
  const int wn = (warp >> 2) * 64;  // 2 warps along N
  const int g = lane >> 2, t = lane & 3;

  // Accumulators start from the addend Cin (or zero): the loads go straight into the accumulator
  // registers, all in flight at once, and their latency hides behind the pipeline prologue.  (Adding
  // Cin in the epilogue instead needs 128 temporaries the kernel does not have, so the compiler
  // serialises the loads: measured +12 % on a 16-block SUMMA step.)
  double acc[4][8][2];
  if (prob.Cin != nullptr && blockIdx.z == 0) {
    const bool cin_vec = (prob.ldcin % 2 == 0) && ((reinterpret_cast<uintptr_t>(prob.Cin) & 15u) == 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + wm + i * 8 + g;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t n = n0 + wn + j * 8 + 2 * t;
        double v0 = 0.0, v1 = 0.0;
        if (m < M && n < N) {
          const double* src = prob.Cin + m * prob.ldcin + n;
          if (n + 1 < N && cin_vec) {
            const double2 v = *reinterpret_cast<const double2*>(src);
            v0 = v.x;
            v1 = v.y;
          } else {
            v0 = src[0];
            if (n + 1 < N) v1 = src[1];
          }
        }
        acc[i][j][0] = v0;
        acc[i][j][1] = v1;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  }

  // ---- producer cursor over (term, k-tile) -----------------------------------------------------------
  int term_idx = 0;
  GemmTerm term = p.use_table ? p.terms[prob.term_begin] : p.single_term;
  int64_t k_lo = 0, k_hi = term.K;
  if (p.k_per_split > 0) {  // split-K slice of a single-term problem
    k_lo = (int64_t)blockIdx.z * p.k_per_split;
    k_hi = k_lo + p.k_per_split;
    if (k_hi > term.K) k_hi = term.K;
  }
  int64_t k_cur = k_lo;
  // total number of k-tiles this CTA will consume
  int KT = 0;
  if (p.use_table) {
    for (int s = 0; s < prob.term_count; ++s)
      KT += (int)((p.terms[prob.term_begin + s].K + BK - 1) / BK);
  } else {
    KT = (int)((k_hi - k_lo + BK - 1) / BK);
  }

  // A tile: "N" storage (m, k) -> rows fixed, k along columns; "T" storage (k, m) -> k along rows.
  TileLoader<AT::rows, AT::cols, AT::pitch, TA> la;
  // B tile: "N" storage (k, n) -> k along rows; "T" storage (n, k) -> k along columns.
  TileLoader<BT::rows, BT::cols, BT::pitch, !TB> lb;
  la.init(term.A, term.lda, m0, M, k_cur);
  lb.init(term.B, term.ldb, n0, N, k_cur);

  auto finish_tile = [&]() {   // after all pieces of one k-tile have been issued
    la.advance();
    lb.advance();
    k_cur += BK;
    if (k_cur >= k_hi && term_idx + 1 < prob.term_count) {  // advance to the next term of the chain
      ++term_idx;
      term = p.terms[prob.term_begin + term_idx];
      k_cur = 0;
      k_hi = term.K;
      la.init(term.A, term.lda, m0, M, 0);
      lb.init(term.B, term.ldb, n0, N, 0);
    }
  };
  using LA = decltype(la);
  using LB = decltype(lb);
  static_assert(LA::ITERS == BK / 4 && LB::ITERS == BK / 4, "one A piece and one B piece per MMA k-step");

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < KT) {
      la.load(sA + s * AT::doubles, k_hi - k_cur, term.A);
      lb.load(sB + s * BT::doubles, k_hi - k_cur, term.B);
      // load() already advanced the pointers
      k_cur += BK;
      if (k_cur >= k_hi && term_idx + 1 < prob.term_count) {
        ++term_idx;
        term = p.terms[prob.term_begin + term_idx];
        k_cur = 0;
        k_hi = term.K;
        la.init(term.A, term.lda, m0, M, 0);
        lb.init(term.B, term.ldb, n0, N, 0);
      }
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    const int next = kt + kStages - 1;
    const bool prefetch = next < KT;
    double* na = sA + (next % kStages) * AT::doubles;
    double* nb = sB + (next % kStages) * BT::doubles;
    const int64_t k_left = k_hi - k_cur;
    const double* a = sA + (kt % kStages) * AT::doubles;
    const double* b = sB + (kt % kStages) * BT::doubles;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      if (prefetch) {   // slot `next` was consumed in iteration kt - 1: free after the barrier above
        la.load_piece(na, kk / 4, k_left, term.A);
        lb.load_piece(nb, kk / 4, k_left, term.B);
      }
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
    if (prefetch) finish_tile();
    cp_async_commit();
  }
  cp_async_wait<0>();

  // Epilogue: lane (g, t) owns C[m = 8i + g][n = 8j + 2t, 2t + 1].
  double* out = prob.C + (int64_t)blockIdx.z * p.split_stride;
  const bool vec_ok = (prob.ldc % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + wm + i * 8 + g;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t n = n0 + wn + j * 8 + 2 * t;
      if (n >= N) continue;
      double* dst = out + m * prob.ldc + n;
      if (n + 1 < N) {
        if (vec_ok) *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
        else {
          dst[0] = acc[i][j][0];
          dst[1] = acc[i][j][1];
        }
      } else {
        dst[0] = acc[i][j][0];
      }
    }
  }
}

// ======================================================================================
// TMA-fed, warp-specialised variant of the same tile kernel
// ======================================================================================
// Same tiles, fragments and shared-memory layout as dgemm_dmma_kernel, but the ring is filled by
// the TMA engine from 2-D tensor maps: a ninth (producer) warp issues TWO
// cp.async.bulk.tensor.2d copies per k-tile (UTMALDG), one per operand.  The box of each map is
// (tile rows) x (tile columns + 4): the four extra columns are fetched and never read, which makes
// the dense box land exactly at the padded pitch (== 4 mod 16 doubles) the conflict-free fragment
// reads need -- no swizzle, no per-row copies (a first version issued one bulk copy per tile row and
// was limited by the ~60-cycle issue rate of the copy engine whenever rows were 256 bytes).  The
// engine zero-fills everything outside the tensor, so ragged M / N / K edges need no special code.
// `full` mbarriers (transaction bytes) hand stages to the eight MMA warps, `empty` mbarriers hand
// them back; there is no CTA-wide barrier in the main loop and no copy instruction in the MMA warps.
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_addr_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE;\n"
      "bra MBAR_WAIT;\n"
      "MBAR_DONE:\n"
      "}\n" ::"r"(smem_addr_u32(bar)), "r"(parity) : "memory");
}
// One box of a 2-D tensor map -> shared memory; c0 = column (innermost) index, c1 = row index.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
          smem_addr_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_addr_u32(bar))
      : "memory");
}

constexpr int kTmaThreads = kGemmThreads + 32;

// Tensor maps of one contraction term (operand A, operand B) and its K.
struct alignas(64) TmaTerm {
  CUtensorMap a;
  CUtensorMap b;
  int64_t K;
  int64_t pad_[7];
};
static_assert(sizeof(TmaTerm) % 64 == 0, "tensor maps must stay 64-byte aligned in the table");

// The term table is written by table_copy_kernel (generic proxy) in an earlier launch on the same
// stream; the copy engine reads descriptors through the tensormap proxy, so the reader acquires
// each pair of maps before its first use (the writer releases, see table_copy_kernel).
__device__ __forceinline__ void tensormap_acquire(const TmaTerm* term) {
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;\n" ::"l"(reinterpret_cast<uint64_t>(&term->a)) : "memory");
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;\n" ::"l"(reinterpret_cast<uint64_t>(&term->b)) : "memory");
}

// Descriptor tables: page-locked host staging (mapped into the device address space) -> workspace.
// A kernel rather than cudaMemcpyAsync, because a host-to-device copy queues on the copy engine
// BEHIND every block upload already submitted on other streams (measured: the first launch group of
// a pipelined matmul started only after all 4 GiB of operands had landed), while a kernel only
// waits for its own stream.
__global__ void __launch_bounds__(256)
table_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t count) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
  asm volatile("fence.proxy.tensormap::generic.release.gpu;\n" ::: "memory");
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(kTmaThreads, 1)
dgemm_dmma_tma_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ TmaTerm single_tma,
                      const TmaTerm* __restrict__ tma_terms) {
  extern __shared__ __align__(128) double smem[];
  using AT = ATile<TA>;
  using BT = BTile<TB>;
  double* sA = smem;
  double* sB = smem + kStages * AT::doubles;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + kStages * BT::doubles);
  uint64_t* empty = full + kStages;
  constexpr uint32_t kStageBytes = (uint32_t)((AT::doubles + BT::doubles) * sizeof(double));

  GemmProblem prob;
  if (p.use_table) {
    int lo = 0, hi = p.nproblems - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (p.problems[mid].tile_begin <= (int)blockIdx.x) lo = mid;
      else hi = mid - 1;
    }
    prob = p.problems[lo];
  } else {
    prob = p.single;
  }
  const int tile = (int)blockIdx.x - prob.tile_begin;
  constexpr int kGroup = 8;
  const int tiles_per_group = kGroup * prob.tiles_n;
  const int group = tile / tiles_per_group;
  const int first_m = group * kGroup;
  const int group_rows = min(prob.tiles_m - first_m, kGroup);
  const int tm = first_m + (tile % tiles_per_group) % group_rows;
  const int tn = (tile % tiles_per_group) / group_rows;
  const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;
  const int64_t M = prob.M, N = prob.N;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kGemmThreads / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  const TmaTerm* first_term = p.use_table ? tma_terms + prob.term_begin : &single_tma;
  int64_t k_lo = 0, k_hi = first_term->K;
  if (p.k_per_split > 0) {   // split-K slice of a single-term problem (k_per_split is a multiple of BK)
    k_lo = (int64_t)blockIdx.z * p.k_per_split;
    k_hi = k_lo + p.k_per_split;
    if (k_hi > first_term->K) k_hi = first_term->K;
  }
  int KT = 0;
  if (p.use_table) {
    for (int s = 0; s < prob.term_count; ++s) KT += (int)((tma_terms[prob.term_begin + s].K + BK - 1) / BK);
  } else {
    KT = (int)((k_hi - k_lo + BK - 1) / BK);
  }

  if (warp == kGemmThreads / 32) {
    // ===== producer warp: one elected lane drives the copy engine =====
    if (lane == 0) {
      const TmaTerm* term = first_term;
      int term_idx = 0;
      if (p.use_table) tensormap_acquire(term);
      int k_cur = (int)k_lo;
      int64_t k_end = k_hi;
      for (int kt = 0; kt < KT; ++kt) {
        const int slot = kt % kStages;
        const int round = kt / kStages;
        if (round > 0) {
          mbar_wait(&empty[slot], (uint32_t)((round - 1) & 1));
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads before async writes
        }
        mbar_expect_tx(&full[slot], kStageBytes);
        // A: "N" storage (m, k) -> box origin (col k, row m).  "T" storage (k, m) -> (col m, row k).
        tma_load_2d(sA + slot * AT::doubles, &term->a, TA ? (int)m0 : k_cur, TA ? k_cur : (int)m0, &full[slot]);
        // B: "N" storage (k, n) -> (col n, row k).  "T" storage (n, k) -> (col k, row n).
        tma_load_2d(sB + slot * BT::doubles, &term->b, TB ? k_cur : (int)n0, TB ? (int)n0 : k_cur, &full[slot]);
        k_cur += BK;
        if (k_cur >= k_end && term_idx + 1 < prob.term_count) {
          ++term_idx;
          ++term;
          tensormap_acquire(term);
          k_cur = 0;
          k_end = term->K;
        }
      }
    }
    return;
  }

  // ===== MMA warps =====
  const int wm = (warp & 3) * 32;
  const int wn = (warp >> 2) * 64;
  const int g = lane >> 2, t = lane & 3;
  double acc[4][8][2];
  if (prob.Cin != nullptr && blockIdx.z == 0) {
    const bool cin_vec = (prob.ldcin % 2 == 0) && ((reinterpret_cast<uintptr_t>(prob.Cin) & 15u) == 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + wm + i * 8 + g;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t n = n0 + wn + j * 8 + 2 * t;
        double v0 = 0.0, v1 = 0.0;
        if (m < M && n < N) {
          const double* src = prob.Cin + m * prob.ldcin + n;
          if (n + 1 < N && cin_vec) {
            const double2 v = *reinterpret_cast<const double2*>(src);
            v0 = v.x;
            v1 = v.y;
          } else {
            v0 = src[0];
            if (n + 1 < N) v1 = src[1];
          }
        }
        acc[i][j][0] = v0;
        acc[i][j][1] = v1;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  }

  for (int kt = 0; kt < KT; ++kt) {
    const int slot = kt % kStages;
    mbar_wait(&full[slot], (uint32_t)((kt / kStages) & 1));
    const double* a = sA + slot * AT::doubles;
    const double* b = sB + slot * BT::doubles;
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
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);
  }

  double* out = prob.C + (int64_t)blockIdx.z * p.split_stride;
  const bool vec_ok = (prob.ldc % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + wm + i * 8 + g;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t n = n0 + wn + j * 8 + 2 * t;
      if (n >= N) continue;
      double* dst = out + m * prob.ldc + n;
      if (n + 1 < N) {
        if (vec_ok) *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
        else {
          dst[0] = acc[i][j][0];
          dst[1] = acc[i][j][1];
        }
      } else {
        dst[0] = acc[i][j][0];
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

// The same fold for MANY partials of a SMALL result (per-CTA partials of the streaming kernels: up to
// one per SM for a 28 x 28 output): one warp per output element, lane l adds partials l, l + 32, ...
// (four independent loads in flight), then a fixed-order butterfly -- deterministic, and ~3 us where
// the one-thread-per-element loop above spends `splits` dependent L2 round trips (20 us at 148).
__global__ void __launch_bounds__(256)
fold_partials_warp_kernel(const double* __restrict__ partial, int splits, int64_t M, int64_t N,
                          const double* __restrict__ cin, int64_t ldcin, double* __restrict__ out, int64_t ldc) {
  const int lane = threadIdx.x & 31;
  const int64_t total = M * N;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= total) return;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int s = lane;
  for (; s + 96 < splits; s += 128) {
    a0 += partial[(int64_t)s * total + i];
    a1 += partial[(int64_t)(s + 32) * total + i];
    a2 += partial[(int64_t)(s + 64) * total + i];
    a3 += partial[(int64_t)(s + 96) * total + i];
  }
  for (; s < splits; s += 32) a0 += partial[(int64_t)s * total + i];
  double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const int64_t m = i / N, n = i - m * N;
    if (cin != nullptr) acc += cin[m * ldcin + n];
    out[m * ldc + n] = acc;
  }
}

// Larger results (the 128 x 128 Gram matrix: 16384 elements x ~148 partials): a CTA folds 32 consecutive elements,
// eight threads per element each adding every eighth partial (coalesced 256-byte rows), then a fixed-order sum of the
// eight through shared memory -- deterministic.
__global__ void __launch_bounds__(256)
fold_partials_tiled_kernel(const double* __restrict__ partial, int splits, int64_t M, int64_t N,
                           const double* __restrict__ cin, int64_t ldcin, double* __restrict__ out, int64_t ldc) {
  __shared__ double part[8][32];
  const int e = threadIdx.x & 31, sg = threadIdx.x >> 5;
  const int64_t total = M * N;
  const int64_t i = (int64_t)blockIdx.x * 32 + e;
  double a0 = 0.0, a1 = 0.0;
  if (i < total) {
    int s = sg;
    for (; s + 8 < splits; s += 16) {
      a0 += partial[(int64_t)s * total + i];
      a1 += partial[(int64_t)(s + 8) * total + i];
    }
    if (s < splits) a0 += partial[(int64_t)s * total + i];
  }
  part[sg][e] = a0 + a1;
  __syncthreads();
  if (sg == 0 && i < total) {
    double acc = part[0][e];
#pragma unroll
    for (int q = 1; q < 8; ++q) acc += part[q][e];
    const int64_t m = i / N, n = i - m * N;
    if (cin != nullptr) acc += cin[m * ldcin + n];
    out[m * ldc + n] = acc;
  }
}

inline int launch_fold_partials(const double* partial, int splits, int64_t M, int64_t N, const double* cin,
                                int64_t ldcin, double* out, int64_t ldc, cudaStream_t s) {
  if (splits >= 16 && M * N > 4096 && M * N <= (1 << 20)) {
    fold_partials_tiled_kernel<<<(unsigned)ceil_div(M * N, 32), 256, 0, s>>>(partial, splits, M, N, cin, ldcin, out, ldc);
  } else if (splits >= 16 && M * N <= 4096) {
    fold_partials_warp_kernel<<<(unsigned)ceil_div(M * N, 8), 256, 0, s>>>(partial, splits, M, N, cin, ldcin, out, ldc);
  } else {
    splitk_fold_kernel<<<blocks_for(M * N, 256, (int64_t)sm_count() * 8), 256, 0, s>>>(partial, splits, M, N, cin, ldcin,
                                                                                       out, ldc);
  }
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

// ======================================================================================
// Skinny A^T B: C[M,N] = A^T B with M, N <= 32 and a very long contraction (K rows)
// ======================================================================================
// The shapes of the LR Hessian X^T (s X) and gradient-like products through the unfused interface
// path (glms.py:232-238: (28, n_b) . (n_b, 28)) and of small Gram matrices.  A 128x128 output tile
// would waste 95 % of the MMA work there; instead every CTA streams 256-row chunks of A (K x M) and
// B (K x N) through a cp.async ring and each warp turns its 32 rows into eight rank-4 DMMA updates
// of the whole (<= 4 x 4 blocks of 8 x 8) output, i.e. the kernel is HBM bound: 8 (M + N) bytes per
// row.  Per-CTA partials are folded by splitk_fold_kernel (deterministic order).
constexpr int kSkinnyStages = 3;

__host__ __device__ constexpr int skinny_pitch(int cols) {   // smallest pitch >= cols with pitch % 16 in {4, 12}
  int p = cols;
  while (!((p % 16 == 4) || (p % 16 == 12))) ++p;
  return p;
}

template <int MB, int NB, int kSkinnyRows>   // kSkinnyRows = 256 or 128 rows per chunk (shared-memory budget)
__global__ void __launch_bounds__(256, 1)
dgemm_tn_skinny_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                       int M, int N, int64_t K, int same, double* __restrict__ partial) {
  constexpr int PA = skinny_pitch(MB * 8), PB = skinny_pitch(NB * 8);
  extern __shared__ __align__(16) double sk_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  double* ringA = sk_smem;
  double* ringB = sk_smem + (size_t)kSkinnyStages * kSkinnyRows * PA;
  const int total_doubles = kSkinnyStages * kSkinnyRows * (PA + (same ? 0 : PB));
  for (int i = threadIdx.x; i < total_doubles; i += 256) sk_smem[i] = 0.0;   // padding columns stay zero
  __syncthreads();

  double acc[MB][NB][2];
#pragma unroll
  for (int i = 0; i < MB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int64_t nchunks = (K + kSkinnyRows - 1) / kSkinnyRows;
  auto load_rows = [&](double* dst, int pitch, const double* src, int64_t ld, int cols, int64_t r0) {
    // each warp copies rows warp, warp + 8, ...; lanes cover the 16-byte pieces of a row
    const int pieces = cols >> 1;
    for (int r = warp; r < kSkinnyRows; r += 8) {
      const int64_t gr = r0 + r;
      const bool in = gr < K;
      for (int c = lane; c < pieces; c += 32)
        cp_async16(dst + r * pitch + 2 * c, in ? src + gr * ld + 2 * c : src, in ? 16 : 0);
    }
  };
  auto load_chunk = [&](int slot, int64_t chunk) {
    const int64_t r0 = chunk * kSkinnyRows;
    load_rows(ringA + (size_t)slot * kSkinnyRows * PA, PA, A, lda, M, r0);
    if (!same) load_rows(ringB + (size_t)slot * kSkinnyRows * PB, PB, B, ldb, N, r0);
  };

  int64_t chunk = blockIdx.x;
  for (int s = 0; s < kSkinnyStages - 1; ++s) {
    const int64_t c = chunk + (int64_t)s * gridDim.x;
    if (c < nchunks) load_chunk(s, c);
    cp_async_commit();
  }
  int slot = 0;
  for (; chunk < nchunks; chunk += gridDim.x) {
    cp_async_wait<kSkinnyStages - 2>();
    __syncthreads();
    {
      const int64_t nxt = chunk + (int64_t)(kSkinnyStages - 1) * gridDim.x;
      int nslot = slot + kSkinnyStages - 1;
      if (nslot >= kSkinnyStages) nslot -= kSkinnyStages;
      if (nxt < nchunks) load_chunk(nslot, nxt);
      cp_async_commit();
    }
    constexpr int kWarpRows = kSkinnyRows / 8;
    const double* a = ringA + (size_t)slot * kSkinnyRows * PA + (size_t)warp * kWarpRows * PA;
    const double* b = same ? a : ringB + (size_t)slot * kSkinnyRows * PB + (size_t)warp * kWarpRows * PB;
    const int pb = same ? PA : PB;
#pragma unroll
    for (int q = 0; q < kWarpRows / 4; ++q) {
      double af[MB], bf[NB];
#pragma unroll
      for (int i = 0; i < MB; ++i) af[i] = a[(4 * q + t) * PA + 8 * i + g];
#pragma unroll
      for (int j = 0; j < NB; ++j) bf[j] = b[(4 * q + t) * pb + 8 * j + g];
#pragma unroll
      for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) dmma884(acc[i][j], af[i], bf[j]);
    }
    ++slot;
    if (slot == kSkinnyStages) slot = 0;
  }
  cp_async_wait<0>();
  __syncthreads();
  // CTA fold: 8 warps -> one (MB*8) x (NB*8) tile in shared memory, then the M x N corner goes out
  constexpr int TM = MB * 8, TN = NB * 8;
  double* red = sk_smem;   // 8 * TM * TN doubles <= 64 KB, the ring is larger
  double* mine = red + warp * TM * TN;
#pragma unroll
  for (int i = 0; i < MB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      mine[(8 * i + g) * TN + 8 * j + 2 * t] = acc[i][j][0];
      mine[(8 * i + g) * TN + 8 * j + 2 * t + 1] = acc[i][j][1];
    }
  __syncthreads();
  double* out = partial + (size_t)blockIdx.x * M * N;
  for (int e = threadIdx.x; e < M * N; e += 256) {
    const int r = e / N, c = e - r * N;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w * TM * TN + r * TN + c];
    out[e] = v;
  }
}

template <int MB, int NB, int kSkinnyRows>
int launch_skinny_rows(const double* A, int64_t lda, const double* B, int64_t ldb, int M, int N, int64_t K,
                       const double* Cin, int64_t ldcin, double* C, int64_t ldc, void* ws, size_t ws_bytes,
                       cudaStream_t s) {
  const int same = (A == B && lda == ldb && M == N) ? 1 : 0;
  constexpr int PA = skinny_pitch(MB * 8), PB = skinny_pitch(NB * 8);
  const size_t smem = (size_t)kSkinnyStages * kSkinnyRows * (PA + (same ? 0 : PB)) * sizeof(double);
  const int64_t nchunks = (K + kSkinnyRows - 1) / kSkinnyRows;
  int grid = sm_count();
  if (grid > nchunks) grid = (int)nchunks;
  NUMS_NEED_WS((size_t)grid * M * N * sizeof(double), ws_bytes);
  NUMS_CUDA_OK(cudaFuncSetAttribute(dgemm_tn_skinny_kernel<MB, NB, kSkinnyRows>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dgemm_tn_skinny_kernel<MB, NB, kSkinnyRows><<<grid, 256, smem, s>>>(A, lda, B, ldb, M, N, K, same,
                                                                     static_cast<double*>(ws));
  NUMS_LAUNCH_OK();
  return launch_fold_partials(static_cast<const double*>(ws), grid, M, N, Cin, ldcin, C, ldc, s);
}

// --------------------------------------------------------------------------------------
// Dense operands (lda == M, ldb == N): a chunk of 128 rows is ONE contiguous run of memory per operand,
// so a producer warp fetches it with one bulk asynchronous copy each (cp.async.bulk -> UBLKCP, tracked by
// the stage's `full` mbarrier) and the eight MMA warps free-run over the ring, handing stages back through
// `empty` mbarriers -- no per-thread copy instructions, no CTA barrier per chunk (the cp.async version
// above spends 32 copy instructions with 14 of 32 lanes active per warp and chunk, and measured 1.55 TB/s
// on the 28 x n_b by n_b x 28 Hessian product of glms.hessian, glms.py:232-238).  Tiles keep the operand's
// own pitch (M resp. N doubles): conflict free when the pitch is 4 or 12 (mod 16), e.g. 28; columns past
// M / N in the last 8-wide block and rows past K in the last chunk are masked in registers.
// --------------------------------------------------------------------------------------
inline bool skinny_dense_enabled() {   // NUMS_SKINNY_DENSE=0 selects the cp.async ring (A/B measurements)
  static const bool on = []() { const char* v = getenv("NUMS_SKINNY_DENSE"); return !(v && v[0] == '0'); }();
  return on;
}
constexpr int kSkDenseRows = 128;
constexpr int kSkDenseStages = 3;
constexpr int kSkDenseThreads = 256 + 32;

__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_addr_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr_u32(bar))
               : "memory");
}

template <int MB, int NB>
__global__ void __launch_bounds__(kSkDenseThreads, 1)
dgemm_tn_skinny_dense_kernel(const double* __restrict__ A, const double* __restrict__ B, int M, int N, int64_t K,
                             int same, double* __restrict__ partial) {
  extern __shared__ __align__(128) double skd_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int stage_doubles = kSkDenseRows * (M + (same ? 0 : N));
  double* ring = skd_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)kSkDenseStages * stage_doubles + 32);   // 32 doubles of slack:
  uint64_t* empty = full + kSkDenseStages;                                    // masked fragment reads run past the last row
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSkDenseStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  const int64_t nchunks = (K + kSkDenseRows - 1) / kSkDenseRows;
  const int64_t my_chunks = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  double acc[MB][NB][2];
#pragma unroll
  for (int i = 0; i < MB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  if (warp == 8) {
    if (lane == 0) {
      for (int64_t i = 0; i < my_chunks; ++i) {
        const int slot = (int)(i % kSkDenseStages);
        const int64_t round = i / kSkDenseStages;
        if (round > 0) {
          mbar_wait(&empty[slot], (uint32_t)((round - 1) & 1));
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads before async writes
        }
        const int64_t r0 = ((int64_t)blockIdx.x + i * gridDim.x) * kSkDenseRows;
        const int64_t rows = (K - r0 < kSkDenseRows) ? (K - r0) : kSkDenseRows;
        const uint32_t bytes_a = (uint32_t)(rows * M * sizeof(double));
        const uint32_t bytes_b = same ? 0u : (uint32_t)(rows * N * sizeof(double));
        double* dst = ring + (size_t)slot * stage_doubles;
        mbar_expect_tx(&full[slot], bytes_a + bytes_b);
        bulk_copy_g2s(dst, A + r0 * M, bytes_a, &full[slot]);
        if (!same) bulk_copy_g2s(dst + kSkDenseRows * M, B + r0 * N, bytes_b, &full[slot]);
      }
    }
  } else {
    constexpr int kWarpRows = kSkDenseRows / 8;
    const bool a_edge = g < M - 8 * (MB - 1);     // lane's column exists in the last 8-wide block of A / B
    const bool b_edge = g < N - 8 * (NB - 1);
    for (int64_t i = 0; i < my_chunks; ++i) {
      const int slot = (int)(i % kSkDenseStages);
      mbar_wait(&full[slot], (uint32_t)((i / kSkDenseStages) & 1));
      const int64_t r0 = ((int64_t)blockIdx.x + i * gridDim.x) * kSkDenseRows;
      const int rows = (int)((K - r0 < kSkDenseRows) ? (K - r0) : kSkDenseRows);
      const double* a = ring + (size_t)slot * stage_doubles + (size_t)warp * kWarpRows * M;
      const double* b = same ? a : ring + (size_t)slot * stage_doubles + kSkDenseRows * M + (size_t)warp * kWarpRows * N;
      const int pb = same ? M : N;
#pragma unroll
      for (int q = 0; q < kWarpRows / 4; ++q) {
        const bool row_ok = warp * kWarpRows + 4 * q + t < rows;
        double af[MB], bf[NB];
#pragma unroll
        for (int ii = 0; ii < MB; ++ii) {
          const double v = a[(4 * q + t) * M + 8 * ii + g];
          af[ii] = (row_ok && (ii < MB - 1 || a_edge)) ? v : 0.0;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const double v = b[(4 * q + t) * pb + 8 * j + g];
          bf[j] = (row_ok && (j < NB - 1 || b_edge)) ? v : 0.0;
        }
#pragma unroll
        for (int ii = 0; ii < MB; ++ii)
#pragma unroll
          for (int j = 0; j < NB; ++j) dmma884(acc[ii][j], af[ii], bf[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
  }
  __syncthreads();
  // CTA fold: 8 warps -> one (MB*8) x (NB*8) tile in shared memory, then the M x N corner goes out
  constexpr int TM = MB * 8, TN = NB * 8;
  double* red = skd_smem;   // 8 * TM * TN doubles <= 64 KB; launch_skinny_dense checks that the ring is larger
  if (warp < 8) {
    double* mine = red + warp * TM * TN;
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        mine[(8 * i + g) * TN + 8 * j + 2 * t] = acc[i][j][0];
        mine[(8 * i + g) * TN + 8 * j + 2 * t + 1] = acc[i][j][1];
      }
  }
  __syncthreads();
  double* out = partial + (size_t)blockIdx.x * M * N;
  for (int e = threadIdx.x; e < M * N; e += kSkDenseThreads) {
    const int r = e / N, c = e - r * N;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w * TM * TN + r * TN + c];
    out[e] = v;
  }
}

template <int MB, int NB>
int launch_skinny_dense(const double* A, const double* B, int M, int N, int64_t K, const double* Cin, int64_t ldcin,
                        double* C, int64_t ldc, void* ws, size_t ws_bytes, cudaStream_t s, bool* handled) {
  const int same = (A == B && M == N) ? 1 : 0;
  const size_t ring_bytes = (size_t)kSkDenseStages * kSkDenseRows * (M + (same ? 0 : N)) * sizeof(double);
  const size_t smem = ring_bytes + 32 * sizeof(double) + 2 * kSkDenseStages * sizeof(uint64_t);
  *handled = false;
  if (smem > 227 * 1024 || ring_bytes < (size_t)8 * (MB * 8) * (NB * 8) * sizeof(double)) return NUMS_OK;
  *handled = true;
  const int64_t nchunks = (K + kSkDenseRows - 1) / kSkDenseRows;
  int grid = sm_count();
  if (grid > nchunks) grid = (int)nchunks;
  NUMS_NEED_WS((size_t)grid * M * N * sizeof(double), ws_bytes);
  NUMS_CUDA_OK(cudaFuncSetAttribute(dgemm_tn_skinny_dense_kernel<MB, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  dgemm_tn_skinny_dense_kernel<MB, NB><<<grid, kSkDenseThreads, smem, s>>>(A, B, M, N, K, same, static_cast<double*>(ws));
  NUMS_LAUNCH_OK();
  return launch_fold_partials(static_cast<const double*>(ws), grid, M, N, Cin, ldcin, C, ldc, s);
}

// --------------------------------------------------------------------------------------
// Gram matrix of a tall dense block with 128 columns: C = A^T A  (the TSQR leaf on the Gram path, config 3:
// cuda_compute._gram_of; application.py:784-814 asks for qr(block, mode="r") of 2 097 152 x 128 blocks).
// The general GEMM computes all 256 8x8 blocks of the 128 x 128 result; by symmetry only the 136 on or above the
// block diagonal are needed, and because the m8n8k4 A and B fragments of A^T A are the SAME registers (lane (g, t)
// holds X[row t][8 c + g] for block column c either way) a warp that owns block row r needs no operand beyond the
// fragments of the block columns it multiplies with.  Block rows are paired so that every one of the eight MMA
// warps owns 17 blocks: warp w takes block row w (16 - w blocks) and block row 15 - w (w + 1 blocks) and loads
// 16 - w fragments per rank-4 update.  All warps consume the same 64-row chunk; a producer warp feeds a 3-deep
// ring with one bulk copy per ROW (1 KB) into rows padded to 132 doubles (pitch = 4 mod 16: conflict-free
// fragment reads; a dense 128-double pitch would put rows t = 0..3 of a fragment on the same banks).  Tensor-pipe
// time is 136 / 256 of the GEMM's: 2 m n^2 "flops" in the time of 1.06 m n^2.
// --------------------------------------------------------------------------------------
bool encode_matrix_map(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                       int box_cols);
constexpr int kSyrkN = 128;
constexpr int kSyrkRows = 64;
constexpr int kSyrkStages = 3;
constexpr int kSyrkPitch = kSyrkN + 4;
constexpr int kSyrkThreads = 256 + 32;
constexpr int kSyrkStageDoubles = kSyrkRows * kSyrkPitch;

template <int W>
__device__ __forceinline__ void syrk_consumer(double* ring, uint64_t* full, uint64_t* empty, int64_t K, int64_t my_chunks,
                                              int lane, bool zero_filled) {
  constexpr int NB = kSyrkN / 8;             // 16 block columns
  constexpr int R1 = NB - 1 - W;             // the second block row of this warp
  const int g = lane >> 2, t = lane & 3;
  double acc0[NB - W][2], acc1[W + 1][2];
#pragma unroll
  for (int c = 0; c < NB - W; ++c) acc0[c][0] = acc0[c][1] = 0.0;
#pragma unroll
  for (int c = 0; c < W + 1; ++c) acc1[c][0] = acc1[c][1] = 0.0;
  for (int64_t i = 0; i < my_chunks; ++i) {
    const int slot = (int)(i % kSyrkStages);
    mbar_wait(&full[slot], (uint32_t)((i / kSyrkStages) & 1));
    const int64_t r0 = ((int64_t)blockIdx.x + i * gridDim.x) * kSyrkRows;
    const int rows = (int)((K - r0 < kSyrkRows) ? (K - r0) : kSyrkRows);
    const double* tile = ring + (size_t)slot * kSyrkStageDoubles;
    if (rows == kSyrkRows || zero_filled) {       // (uniform) every chunk but possibly the last
#pragma unroll 4
      for (int q = 0; q < kSyrkRows / 4; ++q) {
        const double* xr = tile + (4 * q + t) * kSyrkPitch + g;
        double f[NB];
#pragma unroll
        for (int c = W; c < NB; ++c) f[c] = xr[8 * c];
#pragma unroll
        for (int c = W; c < NB; ++c) dmma884(acc0[c - W], f[W], f[c]);
#pragma unroll
        for (int c = R1; c < NB; ++c) dmma884(acc1[c - R1], f[R1], f[c]);
      }
    } else {
#pragma unroll 2
      for (int q = 0; q < kSyrkRows / 4; ++q) {
        const bool row_ok = 4 * q + t < rows;     // rows past the end of the last chunk hold stale data
        const double* xr = tile + (4 * q + t) * kSyrkPitch + g;
        double f[NB];
#pragma unroll
        for (int c = W; c < NB; ++c) {
          const double v = xr[8 * c];
          f[c] = row_ok ? v : 0.0;
        }
#pragma unroll
        for (int c = W; c < NB; ++c) dmma884(acc0[c - W], f[W], f[c]);
#pragma unroll
        for (int c = R1; c < NB; ++c) dmma884(acc1[c - R1], f[R1], f[c]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);
  }
  // every warp waits until all chunks are consumed by everybody, then the ring becomes the 128 x 128 result image
  asm volatile("bar.sync 1, 256;\n" ::: "memory");
  double* img = ring;
#pragma unroll
  for (int c = W; c < NB; ++c) {
    img[(8 * W + g) * kSyrkN + 8 * c + 2 * t] = acc0[c - W][0];
    img[(8 * W + g) * kSyrkN + 8 * c + 2 * t + 1] = acc0[c - W][1];
  }
#pragma unroll
  for (int c = R1; c < NB; ++c) {
    img[(8 * R1 + g) * kSyrkN + 8 * c + 2 * t] = acc1[c - R1][0];
    img[(8 * R1 + g) * kSyrkN + 8 * c + 2 * t + 1] = acc1[c - R1][1];
  }
}

__global__ void __launch_bounds__(kSyrkThreads, 1)
dsyrk128_stream_kernel(const double* __restrict__ A, int64_t K, double* __restrict__ partial,
                       const __grid_constant__ CUtensorMap map, int use_map) {
  extern __shared__ __align__(128) double syrk_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* ring = syrk_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)kSyrkStages * kSyrkStageDoubles);
  uint64_t* empty = full + kSyrkStages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSyrkStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const int64_t nchunks = (K + kSyrkRows - 1) / kSyrkRows;
  const int64_t my_chunks = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (warp == 8) {
    for (int64_t i = 0; i < my_chunks; ++i) {
      const int slot = (int)(i % kSyrkStages);
      const int64_t round = i / kSyrkStages;
      const int64_t r0 = ((int64_t)blockIdx.x + i * gridDim.x) * kSyrkRows;
      const int rows = (int)((K - r0 < kSyrkRows) ? (K - r0) : kSyrkRows);
      double* dst = ring + (size_t)slot * kSyrkStageDoubles;
      if (lane == 0) {
        if (round > 0) {
          mbar_wait(&empty[slot], (uint32_t)((round - 1) & 1));
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads before async writes
        }
        if (use_map) {
          // ONE tensor-map copy per chunk: a box of 64 rows x 132 columns of the K x 128 matrix -- the four columns
          // past the matrix and the rows past K are zero-filled by the copy engine, so the box lands at the padded
          // pitch and the last chunk needs no masking (the same trick as the GEMM's tiles)
          mbar_expect_tx(&full[slot], (uint32_t)(kSyrkStageDoubles * sizeof(double)));
          tma_load_2d(dst, &map, 0, (int)r0, &full[slot]);
        } else {
          mbar_expect_tx(&full[slot], (uint32_t)(rows * kSyrkN * sizeof(double)));
        }
      }
      __syncwarp();
      if (!use_map) {      // no encoder: one bulk copy per row into the padded rows
        for (int r = lane; r < rows; r += 32)
          bulk_copy_g2s(dst + r * kSyrkPitch, A + (r0 + r) * kSyrkN, (uint32_t)(kSyrkN * sizeof(double)), &full[slot]);
      }
    }
  } else {
    switch (warp) {
      case 0: syrk_consumer<0>(ring, full, empty, K, my_chunks, lane, use_map != 0); break;
      case 1: syrk_consumer<1>(ring, full, empty, K, my_chunks, lane, use_map != 0); break;
      case 2: syrk_consumer<2>(ring, full, empty, K, my_chunks, lane, use_map != 0); break;
      case 3: syrk_consumer<3>(ring, full, empty, K, my_chunks, lane, use_map != 0); break;
      case 4: syrk_consumer<4>(ring, full, empty, K, my_chunks, lane, use_map != 0); break;
      case 5: syrk_consumer<5>(ring, full, empty, K, my_chunks, lane, use_map != 0); break;
      case 6: syrk_consumer<6>(ring, full, empty, K, my_chunks, lane, use_map != 0); break;
      default: syrk_consumer<7>(ring, full, empty, K, my_chunks, lane, use_map != 0); break;
    }
  }
  __syncthreads();
  // the blocks on or above the block diagonal were computed; everything below is their mirror image
  double* out = partial + (size_t)blockIdx.x * kSyrkN * kSyrkN;
  for (int e = threadIdx.x; e < kSyrkN * kSyrkN; e += kSyrkThreads) {
    const int r = e / kSyrkN, c = e - r * kSyrkN;
    out[e] = (r / 8 <= c / 8) ? ring[r * kSyrkN + c] : ring[c * kSyrkN + r];
  }
}

inline bool syrk_stream_enabled() {   // NUMS_SYRK_STREAM=0 sends A^T A through the general GEMM (A/B measurements)
  static const bool on = []() { const char* v = getenv("NUMS_SYRK_STREAM"); return !(v && v[0] == '0'); }();
  return on;
}

int launch_syrk128(const double* A, int64_t K, const double* Cin, int64_t ldcin, double* C, int64_t ldc, void* ws,
                   size_t ws_bytes, cudaStream_t s) {
  const size_t smem = (size_t)kSyrkStages * kSyrkStageDoubles * sizeof(double) + 2 * kSyrkStages * sizeof(uint64_t);
  const int64_t nchunks = (K + kSyrkRows - 1) / kSyrkRows;
  int grid = sm_count();
  if (grid > nchunks) grid = (int)nchunks;
  NUMS_NEED_WS((size_t)grid * kSyrkN * kSyrkN * sizeof(double), ws_bytes);
  NUMS_CUDA_OK(cudaFuncSetAttribute(dsyrk128_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const int use_map = encode_matrix_map(&map, A, K, kSyrkN, kSyrkN, kSyrkRows, kSyrkPitch) ? 1 : 0;
  dsyrk128_stream_kernel<<<grid, kSyrkThreads, smem, s>>>(A, K, static_cast<double*>(ws), map, use_map);
  NUMS_LAUNCH_OK();
  return launch_fold_partials(static_cast<const double*>(ws), grid, kSyrkN, kSyrkN, Cin, ldcin, C, ldc, s);
}

// --------------------------------------------------------------------------------------
// Tall block times small square: C (M x 128) = A (M x 128, dense) . B (128 x 128, dense) -- the explicit Q of
// TSQR, Q_i = X_i R^-1 (application.py:833-845; config 3: eight 2 097 152 x 128 blocks).  On the tiled GEMM every
// 128 x 128 output tile has only four k-tiles to amortise its pipeline fill and its epilogue over (27 TFLOP/s).
// Here B never moves: warp w keeps the 128 x 16 slab of B it multiplies with as 64 fragment registers per lane
// for the whole kernel, a producer warp streams 64-row chunks of A through the same padded tensor-map ring as
// the SYRK kernel above, every warp reads the whole chunk (conflict-free fragment reads at pitch 132) and writes
// its 16 output columns straight from the accumulators.  Two LDS per four DMMAs, no CTA barrier, no tile
// prologue: the kernel runs at the tensor pipe's rate.
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSyrkThreads, 1)
dgemm_tall128_stream_kernel(const double* __restrict__ B, double* __restrict__ C, int64_t M,
                            const __grid_constant__ CUtensorMap map) {
  extern __shared__ __align__(128) double tall_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* ring = tall_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)kSyrkStages * kSyrkStageDoubles);
  uint64_t* empty = full + kSyrkStages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSyrkStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const int64_t nchunks = (M + kSyrkRows - 1) / kSyrkRows;
  const int64_t my_chunks = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (warp == 8) {
    if (lane == 0) {
      for (int64_t i = 0; i < my_chunks; ++i) {
        const int slot = (int)(i % kSyrkStages);
        const int64_t round = i / kSyrkStages;
        if (round > 0) {
          mbar_wait(&empty[slot], (uint32_t)((round - 1) & 1));
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads before async writes
        }
        const int64_t r0 = ((int64_t)blockIdx.x + i * gridDim.x) * kSyrkRows;
        mbar_expect_tx(&full[slot], (uint32_t)(kSyrkStageDoubles * sizeof(double)));
        tma_load_2d(ring + (size_t)slot * kSyrkStageDoubles, &map, 0, (int)r0, &full[slot]);
      }
    }
    return;
  }
  const int g = lane >> 2, t = lane & 3;
  // this warp's slab of B as fragments: bf[ks][j] = B[4 ks + t][16 warp + 8 j + g]
  double bf[kSyrkN / 4][2];
#pragma unroll
  for (int ks = 0; ks < kSyrkN / 4; ++ks)
#pragma unroll
    for (int j = 0; j < 2; ++j) bf[ks][j] = B[(size_t)(4 * ks + t) * kSyrkN + 16 * warp + 8 * j + g];
  for (int64_t i = 0; i < my_chunks; ++i) {
    const int slot = (int)(i % kSyrkStages);
    mbar_wait(&full[slot], (uint32_t)((i / kSyrkStages) & 1));
    const int64_t r0 = ((int64_t)blockIdx.x + i * gridDim.x) * kSyrkRows;
    const double* tile = ring + (size_t)slot * kSyrkStageDoubles;
#pragma unroll 1
    for (int rb = 0; rb < kSyrkRows / 8; rb += 2) {        // two 8-row blocks at a time: four independent accumulators
      double acc[2][2][2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[u][j][0] = acc[u][j][1] = 0.0;
      const double* a0 = tile + (8 * rb + g) * kSyrkPitch + t;
      const double* a1 = a0 + 8 * kSyrkPitch;
#pragma unroll
      for (int ks = 0; ks < kSyrkN / 4; ++ks) {
        const double x0 = a0[4 * ks], x1 = a1[4 * ks];
        dmma884(acc[0][0], x0, bf[ks][0]);
        dmma884(acc[0][1], x0, bf[ks][1]);
        dmma884(acc[1][0], x1, bf[ks][0]);
        dmma884(acc[1][1], x1, bf[ks][1]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t row = r0 + 8 * (rb + u) + g;
        if (row < M) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            __stcs(reinterpret_cast<double2*>(C + row * kSyrkN + 16 * warp + 8 * j + 2 * t),
                   make_double2(acc[u][j][0], acc[u][j][1]));
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);
  }
}

inline bool tall128_stream_enabled() {   // NUMS_TALL_STREAM=0 sends X . R^-1 through the tiled GEMM (A/B measurements)
  static const bool on = []() { const char* v = getenv("NUMS_TALL_STREAM"); return !(v && v[0] == '0'); }();
  return on;
}

// Returns NUMS_OK with *handled = false when the tensor-map encoder is unavailable (the caller falls back).
int launch_tall128(const double* A, const double* B, double* C, int64_t M, cudaStream_t s, bool* handled) {
  *handled = false;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (!encode_matrix_map(&map, A, M, kSyrkN, kSyrkN, kSyrkRows, kSyrkPitch)) return NUMS_OK;
  *handled = true;
  const size_t smem = (size_t)kSyrkStages * kSyrkStageDoubles * sizeof(double) + 2 * kSyrkStages * sizeof(uint64_t);
  const int64_t nchunks = (M + kSyrkRows - 1) / kSyrkRows;
  int grid = sm_count();
  if (grid > nchunks) grid = (int)nchunks;
  NUMS_CUDA_OK(cudaFuncSetAttribute(dgemm_tall128_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dgemm_tall128_stream_kernel<<<grid, kSyrkThreads, smem, s>>>(B, C, M, map);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

template <int MB, int NB>
int launch_skinny(const double* A, int64_t lda, const double* B, int64_t ldb, int M, int N, int64_t K,
                  const double* Cin, int64_t ldcin, double* C, int64_t ldc, void* ws, size_t ws_bytes,
                  cudaStream_t s) {
  const bool same = (A == B && lda == ldb && M == N);
  if (lda == M && ldb == N && skinny_dense_enabled()) {
    bool handled = false;
    const int rc = launch_skinny_dense<MB, NB>(A, B, M, N, K, Cin, ldcin, C, ldc, ws, ws_bytes, s, &handled);
    if (rc != NUMS_OK || handled) return rc;
  }
  constexpr int PA = skinny_pitch(MB * 8), PB = skinny_pitch(NB * 8);
  const size_t big = (size_t)kSkinnyStages * 256 * (PA + (same ? 0 : PB)) * sizeof(double);
  if (big <= 224 * 1024)
    return launch_skinny_rows<MB, NB, 256>(A, lda, B, ldb, M, N, K, Cin, ldcin, C, ldc, ws, ws_bytes, s);
  return launch_skinny_rows<MB, NB, 128>(A, lda, B, ldb, M, N, K, Cin, ldcin, C, ldc, ws, ws_bytes, s);
}

int run_skinny(const double* A, int64_t lda, const double* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
               const double* Cin, int64_t ldcin, double* C, int64_t ldc, void* ws, size_t ws_bytes, cudaStream_t s) {
  const int mb = (int)((M + 7) / 8), nb = (int)((N + 7) / 8);
#define NUMS_SKINNY(MBv, NBv) \
  if (mb == MBv && nb == NBv) return launch_skinny<MBv, NBv>(A, lda, B, ldb, (int)M, (int)N, K, Cin, ldcin, C, ldc, ws, ws_bytes, s)
  NUMS_SKINNY(1, 1); NUMS_SKINNY(1, 2); NUMS_SKINNY(1, 3); NUMS_SKINNY(1, 4);
  NUMS_SKINNY(2, 1); NUMS_SKINNY(2, 2); NUMS_SKINNY(2, 3); NUMS_SKINNY(2, 4);
  NUMS_SKINNY(3, 1); NUMS_SKINNY(3, 2); NUMS_SKINNY(3, 3); NUMS_SKINNY(3, 4);
  NUMS_SKINNY(4, 1); NUMS_SKINNY(4, 2); NUMS_SKINNY(4, 3); NUMS_SKINNY(4, 4);
#undef NUMS_SKINNY
  NUMS_FAIL(NUMS_ERR_INVALID, "skinny gemm: bad block counts %d x %d", mb, nb);
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

// float64, even cols <= 32, 16-byte aligned dense A: a warp owns 32 consecutive rows = one contiguous run of
// 16 * cols 16-byte pieces; lane l loads pieces l, l + 32, ... (all HALF = cols / 2 loads independent and
// coalesced), multiplies each by the matching pair of x (held in registers: a 32-row run starts on a row
// boundary, so the pair a lane needs for its k-th piece never changes), parks the products in a per-warp
// scratch and lane l adds the HALF products of row l in order.  No CTA barrier, no integer division in the
// loop; the staged version above measured 2.8 TB/s on the LR forward pass X beta (cols = 28).
template <int HALF>
__global__ void __launch_bounds__(256, 2)
gemv_rows_f64_narrow_kernel(const double* __restrict__ A, int64_t rows, const double* __restrict__ x, int64_t incx,
                            const double* __restrict__ yin, double* __restrict__ y, int64_t incy) {
  constexpr int COLS = 2 * HALF, PITCH = HALF + 1;
  __shared__ double scratch[8][32 * PITCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sc = scratch[warp];
  double2 xv[HALF];
#pragma unroll
  for (int k = 0; k < HALF; ++k) {
    const int p = lane + 32 * k, c = p % HALF;
    xv[k] = make_double2(x[(int64_t)(2 * c) * incx], x[(int64_t)(2 * c + 1) * incx]);
  }
  const int64_t chunks = (rows + 31) / 32;
  const int64_t step = (int64_t)gridDim.x * 8;
  for (int64_t ch = (int64_t)blockIdx.x * 8 + warp; ch < chunks; ch += step) {
    const int64_t r0 = ch * 32;
    const int left = (int)((rows - r0 < 32) ? rows - r0 : 32);
    const double2* src = reinterpret_cast<const double2*>(A + r0 * COLS);
    double2 v[HALF];
    if (left == 32) {
#pragma unroll
      for (int k = 0; k < HALF; ++k) v[k] = __ldcs(src + lane + 32 * k);
    } else {
#pragma unroll
      for (int k = 0; k < HALF; ++k) v[k] = (lane + 32 * k < left * HALF) ? __ldcs(src + lane + 32 * k) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < HALF; ++k) {
      const int p = lane + 32 * k;          // piece p = column pair p % HALF of row p / HALF (constant divisor)
      sc[p + p / HALF] = fma(v[k].y, xv[k].y, v[k].x * xv[k].x);      // == (p / HALF) * PITCH + p % HALF
    }
    __syncwarp();
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < HALF; ++c) acc += sc[lane * PITCH + c];
    __syncwarp();
    if (lane < left) {
      const int64_t r = r0 + lane;
      y[r * incy] = yin ? acc + yin[r * incy] : acc;
    }
  }
}

template <int HALF>
int launch_gemv_rows_f64_narrow(const double* A, int64_t rows, const double* x, int64_t incx, const double* yin,
                                double* y, int64_t incy, cudaStream_t s) {
  const unsigned grid = blocks_for((rows + 31) / 32, 8, (int64_t)sm_count() * 6);
  gemv_rows_f64_narrow_kernel<HALF><<<grid, 256, 0, s>>>(A, rows, x, incx, yin, y, incy);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

inline int run_gemv_rows_f64_narrow(const double* A, int64_t rows, int cols, const double* x, int64_t incx,
                                    const double* yin, double* y, int64_t incy, cudaStream_t s) {
  switch (cols / 2) {
#define NUMS_GEMV_NARROW(H) case H: return launch_gemv_rows_f64_narrow<H>(A, rows, x, incx, yin, y, incy, s)
    NUMS_GEMV_NARROW(1); NUMS_GEMV_NARROW(2); NUMS_GEMV_NARROW(3); NUMS_GEMV_NARROW(4);
    NUMS_GEMV_NARROW(5); NUMS_GEMV_NARROW(6); NUMS_GEMV_NARROW(7); NUMS_GEMV_NARROW(8);
    NUMS_GEMV_NARROW(9); NUMS_GEMV_NARROW(10); NUMS_GEMV_NARROW(11); NUMS_GEMV_NARROW(12);
    NUMS_GEMV_NARROW(13); NUMS_GEMV_NARROW(14); NUMS_GEMV_NARROW(15); NUMS_GEMV_NARROW(16);
#undef NUMS_GEMV_NARROW
  }
  NUMS_FAIL(NUMS_ERR_INVALID, "gemv: narrow kernel does not serve %d columns", cols);
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

// y[c] = (yin ? yin[c] : 0) + sum_s partial[s][c]: one warp per column, lane l adds segments l, l + 32, ...
// (four independent loads in flight) and a fixed-order butterfly finishes -- deterministic.  (One thread per
// column walking all segments was a chain of `segs` dependent L2 round trips: 54 us for the 1184 partials
// of the LR gradient X^T (mu - y), more than the 57 us sweep over X that produced them.)
template <typename T>
__global__ void __launch_bounds__(256)
fold_vector_kernel(const T* __restrict__ partial, int64_t segs, int64_t cols, const T* __restrict__ yin,
                   T* __restrict__ y, int64_t incy) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= cols) return;
  T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
  int64_t s = lane;
  for (; s + 96 < segs; s += 128) {
    a0 += partial[s * cols + c];
    a1 += partial[(s + 32) * cols + c];
    a2 += partial[(s + 64) * cols + c];
    a3 += partial[(s + 96) * cols + c];
  }
  for (; s < segs; s += 32) a0 += partial[s * cols + c];
  T acc = (a0 + a1) + (a2 + a3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[c * incy] = yin ? acc + yin[c * incy] : acc;
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
  if constexpr (std::is_same<T, double>::value) {
    if (cols >= 2 && cols <= 32 && cols % 2 == 0 && lda == cols && rows >= 4096
        && (reinterpret_cast<uintptr_t>(A) & 15u) == 0)
      return run_gemv_rows_f64_narrow(A, rows, (int)cols, x, incx, yin, y, incy, s);
  }
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
    fold_vector_kernel<T><<<(unsigned)ceil_div(cols, 8), 256, 0, s>>>(partial, segs, cols, yin, y, incy);
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
  fold_vector_kernel<T><<<(unsigned)ceil_div(cols, 8), 256, 0, s>>>(partial, segs, cols, yin, y, incy);
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

// ---- tensor maps -------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (the library does not link
// libcuda).  NUMS_GEMM_FEED=cpasync forces the LDGSTS ring for A/B comparisons.
EncodeTiledFn tensor_map_encoder() {
  static const EncodeTiledFn fn = []() -> EncodeTiledFn {
    const char* v = getenv("NUMS_GEMM_FEED");
    if (v && strcmp(v, "cpasync") == 0) return nullptr;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// Map of a row-major (rows x cols) float64 matrix with pitch ld, fetched in (box_rows x box_cols) boxes.
bool encode_matrix_map(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                       int box_cols) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc || rows < 1 || cols < 1 || rows >= (1LL << 31) || cols >= (1LL << 31)) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t elem[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, elem,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Maps of one term: op(A) is M x K, op(B) is K x N; boxes are the padded shared-memory tiles.
bool encode_term(TmaTerm* out, int ta, int tb, int64_t M, int64_t N, const GemmTerm& t) {
  memset(out, 0, sizeof(*out));
  out->K = t.K;
  const bool a_ok = ta ? encode_matrix_map(&out->a, t.A, t.K, M, t.lda, ATile<true>::rows, ATile<true>::pitch)
                       : encode_matrix_map(&out->a, t.A, M, t.K, t.lda, ATile<false>::rows, ATile<false>::pitch);
  if (!a_ok) return false;
  return tb ? encode_matrix_map(&out->b, t.B, N, t.K, t.ldb, BTile<true>::rows, BTile<true>::pitch)
            : encode_matrix_map(&out->b, t.B, t.K, N, t.ldb, BTile<false>::rows, BTile<false>::pitch);
}

template <bool TA, bool TB>
int launch_dmma_tma(const GemmParams& p, const TmaTerm& single, const TmaTerm* table, unsigned tiles, int splits,
                    cudaStream_t s) {
  const size_t smem = (size_t)kStages * (ATile<TA>::doubles + BTile<TB>::doubles) * sizeof(double) +
                      2 * kStages * sizeof(uint64_t);
  NUMS_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma_tma_kernel<TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  dim3 grid(tiles, 1, (unsigned)splits);
  dgemm_dmma_tma_kernel<TA, TB><<<grid, kTmaThreads, smem, s>>>(p, single, table);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

template <bool TA, bool TB>
int launch_dmma(const GemmParams& p, unsigned tiles, int splits, cudaStream_t s) {
  const size_t smem = (size_t)kStages * (ATile<TA>::doubles + BTile<TB>::doubles) * sizeof(double);
  NUMS_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma_kernel<TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  dim3 grid(tiles, 1, (unsigned)splits);
  dgemm_dmma_kernel<TA, TB><<<grid, kGemmThreads, smem, s>>>(p);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

// `single` (one term, maps passed as kernel parameters) or `table` (device array, one entry per term)
// select the TMA-fed kernel; with neither, the cp.async ring runs.
int dispatch_dmma(int ta, int tb, const GemmParams& p, unsigned tiles, int splits, cudaStream_t s,
                  const TmaTerm* single = nullptr, const TmaTerm* table = nullptr) {
  if (single || table) {
    static const TmaTerm none = {};
    const TmaTerm& one = single ? *single : none;
    if (ta && tb) return launch_dmma_tma<true, true>(p, one, table, tiles, splits, s);
    if (ta) return launch_dmma_tma<true, false>(p, one, table, tiles, splits, s);
    if (tb) return launch_dmma_tma<false, true>(p, one, table, tiles, splits, s);
    return launch_dmma_tma<false, false>(p, one, table, tiles, splits, s);
  }
  if (ta && tb) return launch_dmma<true, true>(p, tiles, splits, s);
  if (ta) return launch_dmma<true, false>(p, tiles, splits, s);
  if (tb) return launch_dmma<false, true>(p, tiles, splits, s);
  return launch_dmma<false, false>(p, tiles, splits, s);
}

// Page-locked staging for the descriptor tables of a grouped launch.  cudaMemcpyAsync from pageable
// memory synchronises the stream before it copies, which would block the host behind every queued
// GEMM (and with it the issue of the next SUMMA broadcasts); from page-locked memory it is an
// ordinary asynchronous copy.  A small ring of slots, each guarded by an event recorded after its
// copy, keeps the host from overwriting tables that have not been transferred yet.
class TableStaging {
 public:
  static constexpr int kSlots = 8;
  // Locks the ring; returns a pinned buffer of at least `bytes` (nullptr: use pageable memory).
  char* acquire(size_t bytes) {
    mu_.lock();
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    Slot& sl = slots_[dev][next_[dev]];
    current_ = &sl;
    next_[dev] = (next_[dev] + 1) % kSlots;
    if (sl.event && cudaEventSynchronize(sl.event) != cudaSuccess) { (void)cudaGetLastError(); current_ = nullptr; return nullptr; }
    if (sl.cap < bytes) {
      if (sl.ptr) cudaFreeHost(sl.ptr);
      sl.ptr = nullptr;
      sl.cap = 0;
      const size_t want = bytes < (size_t)(64 << 10) ? (size_t)(64 << 10) : bytes + bytes / 2;
      if (cudaHostAlloc(reinterpret_cast<void**>(&sl.ptr), want, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
        (void)cudaGetLastError();
        sl.ptr = nullptr;
        current_ = nullptr;
        return nullptr;
      }
      sl.cap = want;
    }
    return sl.ptr;
  }
  // Marks the slot busy until everything queued on `s` so far has run, and unlocks the ring.
  void release(cudaStream_t s) {
    if (current_) {
      if (!current_->event && cudaEventCreateWithFlags(&current_->event, cudaEventDisableTiming) != cudaSuccess) {
        (void)cudaGetLastError();
        current_->event = nullptr;
        cudaStreamSynchronize(s);   // no event: fall back to a hard wait so the slot is free again
      } else {
        cudaEventRecord(current_->event, s);
      }
    }
    current_ = nullptr;
    mu_.unlock();
  }

 private:
  static constexpr int kMaxDevices = 16;
  struct Slot { char* ptr = nullptr; size_t cap = 0; cudaEvent_t event = nullptr; };
  Slot slots_[kMaxDevices][kSlots];
  int next_[kMaxDevices] = {};
  Slot* current_ = nullptr;
  std::mutex mu_;
};
TableStaging g_table_staging;

int run_dgemm(int ta, int tb, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* B,
              int64_t ldb, const double* Cin, int64_t ldcin, double* C, int64_t ldc, void* ws,
              size_t ws_bytes, cudaStream_t s) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.single.C = C; p.single.Cin = Cin; p.single.ldc = ldc; p.single.ldcin = ldcin;
  p.single.M = M; p.single.N = N;
  p.single.tiles_m = (int)ceil_div(M, BM);
  p.single.tiles_n = (int)ceil_div(N, BN);
  p.single.term_begin = 0; p.single.term_count = 1; p.single.tile_begin = 0;
  p.single_term.A = A; p.single_term.B = B; p.single_term.lda = lda; p.single_term.ldb = ldb; p.single_term.K = K;
  p.nproblems = 1;
  p.use_table = 0;
  const int64_t tiles = (int64_t)p.single.tiles_m * p.single.tiles_n;
  NUMS_REQUIRE(tiles < 0x7fffffffLL, "gemm: too many output tiles");
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
  if (splits > 1) {
    const int64_t k_per_split = ceil_div(ceil_div(K, splits), BK) * BK;
    splits = (int)ceil_div(K, k_per_split);
    p.k_per_split = k_per_split;
  }
  if (splits > 1) {
    const size_t need = (size_t)splits * M * N * sizeof(double);
    NUMS_NEED_WS(need, ws_bytes);
    p.single.C = static_cast<double*>(ws);
    p.single.ldc = N;
    p.single.Cin = nullptr;
    p.split_stride = M * N;
  } else {
    p.k_per_split = 0;
  }
  TmaTerm maps;
  const bool tma_ok = encode_term(&maps, ta, tb, M, N, p.single_term);
  if (int rc = dispatch_dmma(ta, tb, p, (unsigned)tiles, splits, s, tma_ok ? &maps : nullptr)) return rc;
  if (splits > 1) {
    splitk_fold_kernel<<<blocks_for(M * N, 256, (int64_t)sms * 8), 256, 0, s>>>(
        static_cast<const double*>(ws), splits, M, N, Cin, ldcin, C, ldc);
    NUMS_LAUNCH_OK();
  }
  return NUMS_OK;
}

bool dmma_operand_ok(const void* ptr, int64_t ld) {
  return (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0);
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
    const bool aligned = dmma_operand_ok(A, lda) && dmma_operand_ok(B, ldb);
    if (aligned && ta && !tb && M <= 32 && N <= 32 && M % 2 == 0 && N % 2 == 0 && K >= 8192)
      return run_skinny(A, lda, B, ldb, M, N, K, Cin, ldc, C, ldc, ws, ws_bytes, s);
    if (aligned && ta && !tb && A == B && M == kSyrkN && N == kSyrkN && lda == kSyrkN && ldb == kSyrkN && K >= 16384
        && syrk_stream_enabled())
      return launch_syrk128(A, K, Cin, ldc, C, ldc, ws, ws_bytes, s);
    if (aligned && !ta && !tb && N == kSyrkN && K == kSyrkN && lda == kSyrkN && ldb == kSyrkN && ldc == kSyrkN
        && M >= 16384 && Cin == nullptr && (reinterpret_cast<uintptr_t>(C) & 15u) == 0 && tall128_stream_enabled()) {
      bool handled = false;
      const int rc = launch_tall128(A, B, C, M, s, &handled);
      if (rc != NUMS_OK || handled) return rc;
    }
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

extern "C" int nums_gemm_grouped(int dtype, int trans_a, int trans_b, int nproblems,
                                 const nums_gemm_problem_t* problems_host, int nterms,
                                 const nums_gemm_term_t* terms_host, void* ws, size_t ws_bytes,
                                 void* stream) {
  using namespace nums;
  NUMS_REQUIRE(dtype == NUMS_F64, "gemm_grouped: float64 only (dtype %s)", dtype_name(dtype));
  NUMS_REQUIRE(nproblems >= 1 && nterms >= nproblems && problems_host && terms_host, "gemm_grouped: empty group");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t prob_bytes = ((size_t)nproblems * sizeof(GemmProblem) + 255) & ~(size_t)255;
  const size_t term_bytes = ((size_t)nterms * sizeof(GemmTerm) + 255) & ~(size_t)255;
  std::vector<GemmProblem> probs((size_t)nproblems);
  std::vector<GemmTerm> terms((size_t)nterms);
  int64_t tile_cursor = 0;
  for (int i = 0; i < nproblems; ++i) {
    const nums_gemm_problem_t& src = problems_host[i];
    NUMS_REQUIRE(src.m >= 1 && src.n >= 1 && src.C != nullptr, "gemm_grouped: problem %d is empty", i);
    NUMS_REQUIRE(src.term_count >= 1 && src.term_begin >= 0 && src.term_begin + src.term_count <= nterms,
                 "gemm_grouped: problem %d has a bad term range", i);
    GemmProblem& d = probs[(size_t)i];
    d.C = static_cast<double*>(src.C);
    d.Cin = static_cast<const double*>(src.Cin);
    d.ldc = src.ldc; d.ldcin = src.ldcin; d.M = src.m; d.N = src.n;
    d.tiles_m = (int)ceil_div(src.m, BM);
    d.tiles_n = (int)ceil_div(src.n, BN);
    d.term_begin = src.term_begin; d.term_count = src.term_count;
    d.tile_begin = (int)tile_cursor;
    d.pad_ = 0;
    tile_cursor += (int64_t)d.tiles_m * d.tiles_n;
    NUMS_REQUIRE(tile_cursor < 0x7fffffffLL, "gemm_grouped: too many tiles");
  }
  for (int i = 0; i < nterms; ++i) {
    const nums_gemm_term_t& src = terms_host[i];
    NUMS_REQUIRE(src.k >= 1 && src.A && src.B, "gemm_grouped: term %d is empty", i);
    NUMS_REQUIRE(dmma_operand_ok(src.A, src.lda) && dmma_operand_ok(src.B, src.ldb),
                 "gemm_grouped: term %d is not 16-byte aligned with even pitches", i);
    GemmTerm& d = terms[(size_t)i];
    d.A = static_cast<const double*>(src.A);
    d.B = static_cast<const double*>(src.B);
    d.lda = src.lda; d.ldb = src.ldb; d.K = src.k;
  }
  // tensor maps, one pair per term (the TMA-fed kernel); any failure selects the cp.async ring
  std::vector<TmaTerm> maps;
  bool tma_ok = tensor_map_encoder() != nullptr;
  if (tma_ok) {
    maps.resize((size_t)nterms);
    for (int i = 0; i < nproblems && tma_ok; ++i) {
      const GemmProblem& d = probs[(size_t)i];
      for (int t = d.term_begin; t < d.term_begin + d.term_count && tma_ok; ++t)
        tma_ok = encode_term(&maps[(size_t)t], trans_a, trans_b, d.M, d.N, terms[(size_t)t]);
    }
  }
  const size_t map_bytes = tma_ok ? (size_t)nterms * sizeof(TmaTerm) : 0;
  const size_t total_bytes = prob_bytes + term_bytes + map_bytes;
  NUMS_NEED_WS(total_bytes, ws_bytes);
  char* base = static_cast<char*>(ws);
  NUMS_REQUIRE((reinterpret_cast<uintptr_t>(base) & 63u) == 0, "gemm_grouped: workspace must be 64-byte aligned");
  // one table image [problems | terms | maps] -> one copy, from page-locked staging when available
  char* stage = g_table_staging.acquire(total_bytes);
  std::vector<char> pageable;
  char* image = stage;
  if (!image) {
    pageable.resize(total_bytes);
    image = pageable.data();
  }
  memcpy(image, probs.data(), (size_t)nproblems * sizeof(GemmProblem));
  memcpy(image + prob_bytes, terms.data(), (size_t)nterms * sizeof(GemmTerm));
  if (tma_ok) memcpy(image + prob_bytes + term_bytes, maps.data(), map_bytes);
  cudaError_t copy_err = cudaSuccess;
  void* mapped = nullptr;
  if (stage && cudaHostGetDevicePointer(&mapped, stage, 0) == cudaSuccess && mapped) {
    const size_t count = total_bytes / sizeof(uint4);   // every section is a multiple of 16 bytes
    table_copy_kernel<<<(unsigned)((count + 255) / 256 < 64 ? (count + 255) / 256 : 64), 256, 0, s>>>(
        static_cast<const uint4*>(mapped), reinterpret_cast<uint4*>(base), count);
    copy_err = cudaGetLastError();
    count_launch();
  } else {
    (void)cudaGetLastError();
    copy_err = cudaMemcpyAsync(base, image, total_bytes, cudaMemcpyHostToDevice, s);
  }
  g_table_staging.release(s);
  NUMS_CUDA_OK(copy_err);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.problems = reinterpret_cast<const GemmProblem*>(base);
  p.terms = reinterpret_cast<const GemmTerm*>(base + prob_bytes);
  p.nproblems = nproblems;
  p.use_table = 1;
  return dispatch_dmma(trans_a, trans_b, p, (unsigned)tile_cursor, 1, s, nullptr,
                       tma_ok ? reinterpret_cast<const TmaTerm*>(base + prob_bytes + term_bytes) : nullptr);
}
