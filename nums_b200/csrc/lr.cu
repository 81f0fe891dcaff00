// Fused logistic-regression gradient + Hessian over one row block.
//
// The reference composes this from ~15 per-block kernel calls per Newton iteration
// (nums/models/glms.py:140-143 forward, :213-214 link_inv, :222-227 gradient, :232-238
// hessian; call trace in SURVEY.md 3.5): X is streamed about six times and an X-sized
// temporary `s * X` is materialised.  Here one pass over X produces
//     g = X^T (mu - y),   H = X^T diag(mu (1 - mu)) X,   mu = 1 / (1 + exp(-X beta)).
//
// Layout / roofline: X (n x d, row-major f64) is read exactly once, 8 n (d + 1) bytes.  With
// d = 28 the Hessian needs d(d+1)/2 = 406 FMA per 232 bytes, about the FP64 ridge of a B200,
// so the rank-k update runs on the FP64 tensor pipe (DMMA m8n8k4, only the upper-triangular
// 8x8 blocks), the dot products X beta and the gradient ride along in the same fragment
// layout, and exp() is evaluated once per row (not once per fragment lane).
//
// Each CTA is persistent (grid = #SMs), stages 256-row tiles through a cp.async ring and
// writes one partial (g | H) to the workspace; a second tiny kernel folds the partials in a
// fixed order (deterministic, no atomics) and mirrors H.
#include <cstdlib>
#include "common.cuh"

namespace nums {
namespace {

constexpr int kLrThreads = 256;
constexpr int kTileRows = 256;  // 8 warps x 32 rows

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// pitch == 4 (mod 16) doubles => the (row = t, feature = g) fragment reads are conflict free
__host__ __device__ constexpr int lr_pitch(int nb) { return ((nb * 8 + 11) / 16) * 16 + 4; }

template <int NB>  // NB = ceil(d / 8) feature blocks
__global__ void __launch_bounds__(kLrThreads, 1)
lr_grad_hess_kernel(const double* __restrict__ X, int64_t ldx, const double* __restrict__ y,
                    const double* __restrict__ beta, int64_t n, int d, int stages,
                    double* __restrict__ partial) {
  constexpr int PITCH = lr_pitch(NB);
  constexpr int NTRI = NB * (NB + 1) / 2;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int chunks_per_row = d / 2;  // d is even (checked on the host)

  // zero the whole ring once: padding columns are never written by cp.async afterwards
  for (int i = threadIdx.x; i < stages * kTileRows * PITCH; i += kLrThreads) smem[i] = 0.0;
  __syncthreads();

  double bfrag[NB];  // beta[8 bi + g]
#pragma unroll
  for (int bi = 0; bi < NB; ++bi) bfrag[bi] = (8 * bi + g < d) ? beta[8 * bi + g] : 0.0;

  double hacc[NTRI][2];
  double gacc[NB];
#pragma unroll
  for (int i = 0; i < NTRI; ++i) hacc[i][0] = hacc[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < NB; ++i) gacc[i] = 0.0;

  const int64_t ntiles = (n + kTileRows - 1) / kTileRows;
  auto load_tile = [&](int slot, int64_t tile) {
    double* dst = smem + (size_t)slot * kTileRows * PITCH;
    const int64_t r0 = tile * kTileRows;
    const int total = kTileRows * chunks_per_row;
    for (int c = threadIdx.x; c < total; c += kLrThreads) {
      const int r = c / chunks_per_row, cc = (c - r * chunks_per_row) * 2;
      const int64_t gr = r0 + r;
      const bool in = gr < n;
      cp_async16(dst + r * PITCH + cc, in ? X + gr * ldx + cc : X, in ? 16 : 0);
    }
  };

  int64_t tile = blockIdx.x;
  // prologue: prefetch stages-1 tiles
  for (int s = 0; s < stages - 1; ++s) {
    const int64_t tl = tile + (int64_t)s * gridDim.x;
    if (tl < ntiles) load_tile(s, tl);
    cp_async_commit();
  }
  int slot = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    if (stages == 3) cp_async_wait<1>();
    else cp_async_wait<0>();
    __syncthreads();
    {
      const int64_t nxt = tile + (int64_t)(stages - 1) * gridDim.x;
      int nslot = slot + stages - 1;
      if (nslot >= stages) nslot -= stages;
      if (nxt < ntiles) load_tile(nslot, nxt);
      cp_async_commit();
    }
    const double* xs = smem + (size_t)slot * kTileRows * PITCH + (size_t)warp * 32 * PITCH;
    const int64_t row0 = tile * kTileRows + warp * 32;

    // pass 1: z for the warp's 32 rows.  After the reduction every lane with the same t holds
    // z[4 q + t] for group q; lane (g, t) keeps the one with q == g, i.e. row 4 g + t.
    double zmine = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double* xr = xs + (4 * q + t) * PITCH + g;
      double z = 0.0;
#pragma unroll
      for (int bi = 0; bi < NB; ++bi) z = fma(xr[8 * bi], bfrag[bi], z);
      z += __shfl_xor_sync(0xffffffffu, z, 4);
      z += __shfl_xor_sync(0xffffffffu, z, 8);
      z += __shfl_xor_sync(0xffffffffu, z, 16);
      if (q == g) zmine = z;
    }
    const int64_t myrow = row0 + 4 * g + t;
    const double yv = myrow < n ? y[myrow] : 0.0;
    const double mu = 1.0 / (1.0 + exp(-zmine));
    const double e_mine = mu - yv;
    const double s_mine = mu * (1.0 - mu);

    // pass 2: rank-4 updates per group.
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double s = __shfl_sync(0xffffffffu, s_mine, 4 * q + t);
      const double e = __shfl_sync(0xffffffffu, e_mine, 4 * q + t);
      const double* xr = xs + (4 * q + t) * PITCH + g;
      double xf[NB], af[NB];
#pragma unroll
      for (int bi = 0; bi < NB; ++bi) {
        xf[bi] = xr[8 * bi];
        af[bi] = s * xf[bi];
        gacc[bi] = fma(e, xf[bi], gacc[bi]);
      }
      int idx = 0;
#pragma unroll
      for (int bi = 0; bi < NB; ++bi)
#pragma unroll
        for (int bj = bi; bj < NB; ++bj) {
          dmma884(hacc[idx], af[bi], xf[bj]);
          ++idx;
        }
    }
    ++slot;
    if (slot == stages) slot = 0;
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- CTA fold through shared memory (reuses the ring) ----------------------------------------
  // layout per warp: [NB*8 gradient][ (NB*8)^2 hessian ]
  constexpr int D8 = NB * 8;
  constexpr int PER_WARP = D8 + D8 * D8;
  double* red = smem;  // 8 * PER_WARP doubles <= ring size (checked on the host)
  for (int i = threadIdx.x; i < 8 * PER_WARP; i += kLrThreads) red[i] = 0.0;
  __syncthreads();
  double* mine = red + warp * PER_WARP;
#pragma unroll
  for (int bi = 0; bi < NB; ++bi) {
    double v = gacc[bi];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (t == 0) mine[8 * bi + g] = v;
  }
  {
    int idx = 0;
#pragma unroll
    for (int bi = 0; bi < NB; ++bi)
#pragma unroll
      for (int bj = bi; bj < NB; ++bj) {
        double* h = mine + D8 + (8 * bi + g) * D8 + 8 * bj + 2 * t;
        h[0] = hacc[idx][0];
        h[1] = hacc[idx][1];
        ++idx;
      }
  }
  __syncthreads();
  double* out = partial + (size_t)blockIdx.x * (d + d * d);
  for (int i = threadIdx.x; i < d + d * d; i += kLrThreads) {
    int src;
    if (i < d) src = i;
    else {
      const int r = (i - d) / d, c = (i - d) - r * d;
      // only the upper triangle is read (blocks below the block diagonal were never computed, and
      // mirroring inside diagonal blocks makes H exactly symmetric)
      src = (r <= c) ? D8 + r * D8 + c : D8 + c * D8 + r;
    }
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) acc += red[w * PER_WARP + src];
    out[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// Fast path: dense row blocks (ldx == d) with d == 4 or 12 (mod 16), e.g. the HIGGS shape d = 28.
//
// A 256-row tile of X is then one contiguous run of 256*d*8 bytes, so the whole tile is fetched by
// ONE bulk asynchronous copy (cp.async.bulk.shared::cluster.global -> UBLKCP, the TMA engine)
// issued by a single thread and tracked by an mbarrier -- no per-thread address arithmetic, no
// LDGSTS issue slots.  With pitch d == 4 or 12 (mod 16) doubles the (row = t, feature = g)
// fragment reads of the dense tile are bank-conflict free without padding.
// A ninth warp is the producer (one lane: wait for the slot's `empty` mbarrier, issue the copy);
// the eight consumer warps free-run through the tiles, each signalling `empty` on its own, so warps
// drift apart and one warp's scalar pass 1 overlaps the others' tensor-pipe pass 2.
//   pass 1: lane l owns row l of the warp's 32 rows: z = x . beta with 128-bit shared loads,
//           mu = sigmoid(z); s = mu (1 - mu) and e = mu - y go to a per-warp shared scratch
//           (y is prefetched one tile ahead into a register);
//   pass 2: 8 rank-4 updates; fragments straight from the tile, s / e broadcast-read from the
//           scratch (no shuffles); H on the DMMA pipe (upper-triangular 8x8 blocks), g by DFMA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kDenseStages = 3;
constexpr int kDenseThreads = kLrThreads + 32;   // 8 consumer warps + 1 producer warp
constexpr int kMaxLrBlocks = 16;

// Several row blocks of one row-sharded matrix handled by a single launch (the G row blocks of
// X / y in glms.newton): tile t belongs to block b with tile_begin[b] <= t < tile_begin[b + 1].
struct LrBlocks {
  const double* X[kMaxLrBlocks];
  const double* y[kMaxLrBlocks];
  int64_t rows[kMaxLrBlocks];
  int64_t tile_begin[kMaxLrBlocks + 1];
  int count;
};

__device__ __forceinline__ int lr_find_block(const LrBlocks& blk, int64_t tile) {
  int b = 0;
  while (b + 1 < blk.count && tile >= blk.tile_begin[b + 1]) ++b;
  return b;
}

template <int D>   // number of features, compile time: pass 1 becomes straight-line code
__global__ void __launch_bounds__(kDenseThreads, 1)
lr_grad_hess_dense_kernel(const __grid_constant__ LrBlocks blk, const double* __restrict__ beta,
                          double* __restrict__ partial) {
  constexpr int d = D;
  constexpr int NB = (D + 7) / 8;
  constexpr int NTRI = NB * (NB + 1) / 2;
  extern __shared__ __align__(128) unsigned char lr_smem[];
  const int tile_doubles = kTileRows * d;
  double* ring = reinterpret_cast<double*>(lr_smem);                       // kDenseStages tiles
  double* scratch = ring + (size_t)kDenseStages * tile_doubles;            // 8 warps x 2 buffers x (32 s + 32 e)
  double* bsm = scratch + 8 * 128;                                          // beta, zero padded to NB*8
  uint64_t* full = reinterpret_cast<uint64_t*>(bsm + NB * 8);              // data landed (producer -> consumers)
  uint64_t* empty = full + kDenseStages;                                   // slot released (8 warps -> producer)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  for (int i = threadIdx.x; i < kDenseStages * tile_doubles; i += kDenseThreads) ring[i] = 0.0;
  for (int i = threadIdx.x; i < NB * 8; i += kDenseThreads) bsm[i] = i < d ? beta[i] : 0.0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kDenseStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  fence_proxy_async();   // the zero fill (generic proxy) happens-before the bulk copies (async proxy)
  __syncthreads();

  const int64_t ntiles = blk.tile_begin[blk.count];
  const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  // Two accumulator sets used by alternating row groups: consecutive DMMAs into the same
  // accumulator are then 2 x NTRI instructions apart, which covers the DMMA result latency.
  double hacc[2][NTRI][2];
#pragma unroll
  for (int i = 0; i < NTRI; ++i) hacc[0][i][0] = hacc[0][i][1] = hacc[1][i][0] = hacc[1][i][1] = 0.0;

  if (warp == 8) {
    // ===== producer warp: one lane feeds the ring with bulk (TMA) copies =====
    if (lane == 0) {
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int slot = (int)(i % kDenseStages);
        const int64_t round = i / kDenseStages;
        if (round > 0) mbar_wait(&empty[slot], (uint32_t)((round - 1) & 1));
        fence_proxy_async();   // consumers' generic-proxy reads of this slot precede the async write
        const int64_t tile = (int64_t)blockIdx.x + i * gridDim.x;
        const int b = lr_find_block(blk, tile);
        const int64_t r0 = (tile - blk.tile_begin[b]) * kTileRows;
        const int64_t rows = (blk.rows[b] - r0 < kTileRows) ? (blk.rows[b] - r0) : kTileRows;
        const uint32_t bytes = (uint32_t)(rows * d * sizeof(double));
        mbar_expect_tx(&full[slot], bytes);
        bulk_copy_g2s(ring + (size_t)slot * tile_doubles, blk.X[b] + r0 * d, bytes, &full[slot]);
      }
    }
  } else {
    // ===== consumer warps: 32 rows each per tile, free-running (no CTA barrier per tile) =====
    // Software pipelined: while the tensor pipe works through the rank-4 updates of tile i
    // (pass 2), the same warp already evaluates z / sigmoid for tile i + 1 (pass 1); the two are
    // independent instruction streams in one basic block, so the long FP64 latency chains of
    // pass 1 (dot product, exp, divide) hide behind the 16-cycle DMMA cadence.
    double* scr = scratch + warp * 128;          // [buffer][s: 32 | e: 32]
    const int gcol_block = d >> 3, gcol_lane = d & 7;   // padding column that carries e (gradient)
    auto fetch_y = [&](int64_t tile, bool& valid) -> double {
      const int b = lr_find_block(blk, tile);
      const int64_t row = (tile - blk.tile_begin[b]) * kTileRows + warp * 32 + lane;
      valid = row < blk.rows[b];
      return valid ? blk.y[b][row] : 0.0;
    };
    // pass 1 is cut into slices so that it can be issued between the MMA groups of pass 2:
    // slice q accumulates features [4q, 4q + 4) of this lane's row into four partial sums.
    auto dot_slice = [&](const double* xr, int q, double (&z)[4]) {
      const int j = 4 * q;
      if (j + 4 <= D) {
        const double2 v = *reinterpret_cast<const double2*>(xr + j);
        const double2 w = *reinterpret_cast<const double2*>(xr + j + 2);
        z[0] = fma(v.x, bsm[j], z[0]);
        z[1] = fma(v.y, bsm[j + 1], z[1]);
        z[2] = fma(w.x, bsm[j + 2], z[2]);
        z[3] = fma(w.y, bsm[j + 3], z[3]);
      } else if (j + 2 <= D) {
        const double2 v = *reinterpret_cast<const double2*>(xr + j);
        z[0] = fma(v.x, bsm[j], z[0]);
        z[1] = fma(v.y, bsm[j + 1], z[1]);
      }
    };
    constexpr int kSlices = (D + 3) / 4;
    auto finish = [&](const double (&z)[4], double yv, bool valid, double* out_s, double* out_e) {
      const double mu = 1.0 / (1.0 + exp(-((z[0] + z[1]) + (z[2] + z[3]))));
      out_s[lane] = valid ? mu * (1.0 - mu) : 0.0;   // rows past the end may hold stale data: weight 0
      out_e[lane] = valid ? mu - yv : 0.0;
    };

    if (my_tiles > 0) {
      bool valid0 = false;
      const double y0 = fetch_y(blockIdx.x, valid0);
      mbar_wait(&full[0], 0u);
      double z[4] = {0.0, 0.0, 0.0, 0.0};
      const double* xr = ring + (size_t)warp * 32 * d + lane * d;
#pragma unroll
      for (int q = 0; q < kSlices; ++q) dot_slice(xr, q, z);
      finish(z, y0, valid0, scr, scr + 32);
      __syncwarp();
    }
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int slot = (int)(i % kDenseStages);
      const bool has_next = i + 1 < my_tiles;
      const int nslot = has_next ? (int)((i + 1) % kDenseStages) : slot;
      bool valid_next = false;
      const double y_next = has_next ? fetch_y((int64_t)blockIdx.x + (i + 1) * gridDim.x, valid_next) : 0.0;
      if (has_next) mbar_wait(&full[nslot], (uint32_t)(((i + 1) / kDenseStages) & 1));
      const double* xs = ring + (size_t)slot * tile_doubles + (size_t)warp * 32 * d;
      const double* xr_next = ring + (size_t)nslot * tile_doubles + (size_t)warp * 32 * d + lane * d;
      const double* cur_s = scr + (i & 1) * 64;
      const double* cur_e = cur_s + 32;
      double* nxt_s = scr + ((i + 1) & 1) * 64;
      double z[4] = {0.0, 0.0, 0.0, 0.0};

      // pass 2 of tile i: H += x (s x)^T with DMMA; the padding column d of the B operand carries e,
      // so the gradient X^T e accumulates in column d of the same tensor-pipe product.  Between the
      // groups, slices of pass 1 of tile i + 1 (same tile again when there is none: result unused).
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const double sv = cur_s[4 * q + t];
        const double ev = cur_e[4 * q + t];
        const double* xr = xs + (4 * q + t) * d + g;
        double xf[NB], bf[NB];
#pragma unroll
        for (int bi = 0; bi < NB; ++bi) {
          xf[bi] = (8 * bi + g < d) ? xr[8 * bi] : 0.0;
          bf[bi] = sv * xf[bi];
        }
        if (g == gcol_lane) bf[NB - 1] = ev;   // gcol_block == NB - 1 for every d this kernel serves
        if (q < kSlices) dot_slice(xr_next, q, z);
        int idx = 0;
#pragma unroll
        for (int bi = 0; bi < NB; ++bi)
#pragma unroll
          for (int bj = bi; bj < NB; ++bj) {
            dmma884(hacc[q & 1][idx], xf[bi], bf[bj]);
            ++idx;
          }
      }
#pragma unroll
      for (int q = 8; q < kSlices; ++q) dot_slice(xr_next, q, z);
      finish(z, y_next, valid_next && has_next, nxt_s, nxt_s + 32);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);   // this warp is done with the slot
      (void)gcol_block;
    }
  }
  __syncthreads();

  // ---- CTA fold through shared memory (reuses the ring) ----------------------------------------
  constexpr int D8 = NB * 8;
  constexpr int PER_WARP = D8 + D8 * D8;
  double* red = ring;
  for (int i = threadIdx.x; i < 8 * PER_WARP; i += kDenseThreads) red[i] = 0.0;
  __syncthreads();
  if (warp < 8) {
    double* mine = red + warp * PER_WARP;
    int idx = 0;
#pragma unroll
    for (int bi = 0; bi < NB; ++bi)
#pragma unroll
      for (int bj = bi; bj < NB; ++bj) {
        double* h = mine + D8 + (8 * bi + g) * D8 + 8 * bj + 2 * t;
        h[0] = hacc[0][idx][0] + hacc[1][idx][0];
        h[1] = hacc[0][idx][1] + hacc[1][idx][1];
        ++idx;
      }
  }
  __syncthreads();
  double* out = partial + (size_t)blockIdx.x * (d + d * d);
  for (int i = threadIdx.x; i < d + d * d; i += kDenseThreads) {
    int src;
    if (i < d) src = D8 + i * D8 + d;     // gradient: column d (the padding column that carried e)
    else {
      const int r = (i - d) / d, c = (i - d) - r * d;
      src = (r <= c) ? D8 + r * D8 + c : D8 + c * D8 + r;
    }
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) acc += red[w * PER_WARP + src];
    out[i] = acc;
  }
}

__global__ void __launch_bounds__(256)
lr_fold_kernel(const double* __restrict__ partial, int parts, int len, double* __restrict__ out) {
  // one warp per output element: lane l adds partials l, l + 32, ... then a fixed-order butterfly (deterministic);
  // a single thread walking all ~148 partials was a chain of dependent L2 round trips
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= len) return;
  double acc = 0.0;
  for (int p = lane; p < parts; p += 32) acc += partial[(size_t)p * len + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[i] = acc;
}

// ---------------------------------------------------------------------------------------------
// d = 28 (the HIGGS width): the Hessian from SEVEN 8x8x4 tensor-pipe products per rank-4 update instead of
// the ten upper-triangular 8x8 blocks above.  The 28 features are seven groups of four; an m8n8k4 product
// with row set (G_a | G_b) and column set (G_c | G_d) yields the four 4x4 blocks (a,c) (a,d) (b,c) (b,d).
// The 28 unordered group pairs (21 off-diagonal + 7 diagonal) are covered exactly once by the seven lines
// {k, k+1, k+3} (mod 7) of the Fano plane -- every two points lie on exactly one line -- taking for line k
// rows (G_k | G_{k+1}) and columns (G_k | G_{k+3}): blocks (k,k), (k,k+3), (k+1,k), (k+1,k+3).  That is the
// minimum (28 blocks / 4 per product) and 30 % less tensor-pipe time than the triangular cover (measured with the
// round-1 structure: 0.613 -> 0.587 ms on 11M x 28; DMMA shares the FP64 pipe with DFMA / DMUL).  Operands: lane (g, t) of a warp holds row t, feature gi = g & 3 of
// seven groups, loaded ROTATED by one group on the upper half of the warp (g >= 4), so that the A operand of
// product k is register zv[k] on every lane; the B operand is s * zv[k] on the lower half and s * zv[(k+2) % 7]
// on the upper half (seven selects).  No padding column is left for the gradient: g = X^T e is accumulated
// next to it with seven FMAs per update (the plain FP64 pipe is otherwise idle).
// ---------------------------------------------------------------------------------------------
// Structure: pass 1 of a tile (x . beta, sigmoid, one row per lane) simply precedes its pass 2 inside a warp, and the
// latency chains of pass 1 are hidden by OTHER warps' tensor-pipe work -- thread-level parallelism instead of the
// software pipelining of the round-1 kernel (pass 1 of tile i + 1 interleaved with pass 2 of tile i), which measured
// slower (0.605 vs 0.577 ms with the same eight warps).  CW consumer warps of 32 rows each; every scheduler has its
// own FP64 pipe, so CW must be a multiple of four or the tile time is set by the schedulers that got one warp more
// (nine or ten consumer warps measured SLOWER than eight).  Twelve consumer warps -- three per scheduler, the most that
// leaves 168 registers per thread -- leave no room for a producer warp (a 13th warp would cap every thread at 128
// registers), so with INWARP lane 0 of warp 0 feeds the ring between its tiles: before tile i it waits until all
// warps have handed back the slot of tile i - 1 and refills it with tile i - 1 + STAGES.  Tiles of CW x 32 rows,
// STAGES-deep ring; a consumer needs only its current tile.  Measured on 11 M x 28: <12, 2, in-warp> 0.536 ms,
// <11, 2, producer warp> 0.558, <8, 3, producer warp> 0.577 (profiles/r2_lr_kernel_experiments.md).
template <int CW, int STAGES, bool INWARP>   // INWARP: no producer warp, lane 0 of warp 0 feeds the ring between its tiles
__global__ void __launch_bounds__(32 * (CW + (INWARP ? 0 : 1)), 1)
lr_grad_hess_fano28_tlp_kernel(const __grid_constant__ LrBlocks blk, const double* __restrict__ beta,
                               double* __restrict__ partial) {
  constexpr int d = 28;
  constexpr int NG = 7;
  constexpr int ROWS = 32 * CW;
  constexpr int THREADS = 32 * (CW + (INWARP ? 0 : 1));
  extern __shared__ __align__(128) unsigned char lr_smem[];
  constexpr int tile_doubles = ROWS * d;
  double* ring = reinterpret_cast<double*>(lr_smem);
  double* scratch = ring + (size_t)STAGES * tile_doubles;                   // CW warps x (32 s + 32 e)
  double* bsm = scratch + CW * 64;
  uint64_t* full = reinterpret_cast<uint64_t*>(bsm + 32);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int half = g >> 2, gi = g & 3;

  for (int i = threadIdx.x; i < STAGES * tile_doubles; i += THREADS) ring[i] = 0.0;
  for (int i = threadIdx.x; i < 32; i += THREADS) bsm[i] = i < d ? beta[i] : 0.0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  fence_proxy_async();
  __syncthreads();

  const int64_t ntiles = blk.tile_begin[blk.count];
  const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  double hacc[NG][2];
  double gacc[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) hacc[k][0] = hacc[k][1] = gacc[k] = 0.0;

  auto issue_tile = [&](int64_t i) {      // one bulk copy: tile i of this CTA into slot i % STAGES
    const int slot = (int)(i % STAGES);
    const int64_t tile = (int64_t)blockIdx.x + i * gridDim.x;
    const int b = lr_find_block(blk, tile);
    const int64_t r0 = (tile - blk.tile_begin[b]) * ROWS;
    const int64_t rows = (blk.rows[b] - r0 < ROWS) ? (blk.rows[b] - r0) : ROWS;
    const uint32_t bytes = (uint32_t)(rows * d * sizeof(double));
    mbar_expect_tx(&full[slot], bytes);
    bulk_copy_g2s(ring + (size_t)slot * tile_doubles, blk.X[b] + r0 * d, bytes, &full[slot]);
  };
  if (INWARP && threadIdx.x == 0) {
    for (int64_t i = 0; i < STAGES && i < my_tiles; ++i) issue_tile(i);
  }
  if (!INWARP && warp == CW) {
    if (lane == 0) {
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int slot = (int)(i % STAGES);
        const int64_t round = i / STAGES;
        if (round > 0) mbar_wait(&empty[slot], (uint32_t)((round - 1) & 1));
        fence_proxy_async();
        const int64_t tile = (int64_t)blockIdx.x + i * gridDim.x;
        const int b = lr_find_block(blk, tile);
        const int64_t r0 = (tile - blk.tile_begin[b]) * ROWS;
        const int64_t rows = (blk.rows[b] - r0 < ROWS) ? (blk.rows[b] - r0) : ROWS;
        const uint32_t bytes = (uint32_t)(rows * d * sizeof(double));
        mbar_expect_tx(&full[slot], bytes);
        bulk_copy_g2s(ring + (size_t)slot * tile_doubles, blk.X[b] + r0 * d, bytes, &full[slot]);
      }
    }
  } else {
    double* scr_s = scratch + warp * 64;
    double* scr_e = scr_s + 32;
    int zo[NG];
#pragma unroll
    for (int j = 0; j < NG; ++j) zo[j] = 4 * ((j + half) % NG) + gi;
    const int swz = (lane >> 2) & 1;           // chunk order c ^ swz: conflict-free 128-bit reads at a 224-byte row stride
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int slot = (int)(i % STAGES);
      if (INWARP && warp == 0 && i >= 1) {
        // the slot of tile i - 1 is refilled with tile i - 1 + STAGES as soon as all warps have handed it back
        const int64_t nxt = i - 1 + STAGES;
        if (lane == 0 && nxt < my_tiles) {
          mbar_wait(&empty[(i - 1) % STAGES], (uint32_t)(((i - 1) / STAGES) & 1));
          fence_proxy_async();
          issue_tile(nxt);
        }
        __syncwarp();
      }
      const int64_t tile = (int64_t)blockIdx.x + i * gridDim.x;
      const int b = lr_find_block(blk, tile);
      const int64_t row = (tile - blk.tile_begin[b]) * ROWS + warp * 32 + lane;
      const bool valid = row < blk.rows[b];
      const double yv = valid ? blk.y[b][row] : 0.0;
      mbar_wait(&full[slot], (uint32_t)((i / STAGES) & 1));
      const double* xs = ring + (size_t)slot * tile_doubles + (size_t)warp * 32 * d;
      {   // pass 1: lane l owns row l of the warp's 32 rows
        const double* xrow = xs + lane * d;
        double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
#pragma unroll
        for (int c = 0; c < d / 2; c += 2) {
          const double2 x0 = *reinterpret_cast<const double2*>(xrow + 2 * (c ^ swz));
          const double2 x1 = *reinterpret_cast<const double2*>(xrow + 2 * ((c + 1) ^ swz));
          const double2 b0 = *reinterpret_cast<const double2*>(bsm + 2 * (c ^ swz));
          const double2 b1 = *reinterpret_cast<const double2*>(bsm + 2 * ((c + 1) ^ swz));
          z0 = fma(x0.x, b0.x, z0);
          z1 = fma(x0.y, b0.y, z1);
          z2 = fma(x1.x, b1.x, z2);
          z3 = fma(x1.y, b1.y, z3);
        }
        const double mu = 1.0 / (1.0 + exp(-((z0 + z1) + (z2 + z3))));
        scr_s[lane] = valid ? mu * (1.0 - mu) : 0.0;   // rows past the end may hold stale data: weight 0
        scr_e[lane] = valid ? mu - yv : 0.0;
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q) {   // pass 2: H += x (s x)^T, seven DMMAs per rank-4 update; g += e x
        const double sv = scr_s[4 * q + t];
        const double ev = scr_e[4 * q + t];
        const double* xr = xs + (4 * q + t) * d;
        double zv[NG], sz[NG];
#pragma unroll
        for (int j = 0; j < NG; ++j) zv[j] = xr[zo[j]];
#pragma unroll
        for (int j = 0; j < NG; ++j) {
          sz[j] = sv * zv[j];
          gacc[j] = fma(ev, zv[j], gacc[j]);
        }
#pragma unroll
        for (int k = 0; k < NG; ++k) {
          const double bk = half ? sz[(k + 2) % NG] : sz[k];
          dmma884(hacc[k], zv[k], bk);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
  }
  __syncthreads();

  constexpr int PER_WARP = d + d * d;
  double* red = ring;          // CW x 812 doubles <= the ring
  if (warp < CW) {
    double* mine = red + warp * PER_WARP;
#pragma unroll
    for (int j = 0; j < NG; ++j) {
      double v = gacc[j];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (t == 0 && half == 0) mine[4 * j + gi] = v;
    }
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      const int rowf = half ? 4 * ((k + 1) % NG) + gi : 4 * k + gi;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = 2 * t + e;
        const int colf = n < 4 ? 4 * k + n : 4 * ((k + 3) % NG) + n - 4;
        mine[d + rowf * d + colf] = hacc[k][e];
      }
    }
  }
  __syncthreads();
  double* out = partial + (size_t)blockIdx.x * PER_WARP;
  for (int i = threadIdx.x; i < PER_WARP; i += THREADS) {
    int src = i;
    if (i >= d) {
      const int r = (i - d) / d, c = (i - d) - r * d;
      const int diff = (c / 4 - r / 4 + NG) % NG;
      const bool as_is = diff == 0 ? r <= c : (diff == 2 || diff == 3 || diff == 6);
      src = as_is ? d + r * d + c : d + c * d + r;
    }
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < CW; ++w) acc += red[w * PER_WARP + src];
    out[i] = acc;
  }
}

template <int CW, int STAGES, bool INWARP>
int launch_lr_fano28_tlp(LrBlocks blk, const double* beta, double* out, void* ws, size_t ws_bytes, cudaStream_t s) {
  constexpr int d = 28, ROWS = 32 * CW;
  int64_t tiles = 0;      // the block table arrives with tile counts for kTileRows-row tiles: recount
  for (int b = 0; b < blk.count; ++b) {
    blk.tile_begin[b] = tiles;
    tiles += (blk.rows[b] + ROWS - 1) / ROWS;
  }
  blk.tile_begin[blk.count] = tiles;
  const size_t tile_bytes = (size_t)ROWS * d * sizeof(double);
  const size_t smem = STAGES * tile_bytes + (CW * 64 + 32) * sizeof(double) + 2 * STAGES * sizeof(uint64_t);
  const int len = d + d * d;
  int grid = sm_count();
  if (grid > tiles) grid = (int)tiles;
  if (grid < 1) grid = 1;
  NUMS_NEED_WS((size_t)grid * len * sizeof(double), ws_bytes);
  NUMS_CUDA_OK(cudaFuncSetAttribute(lr_grad_hess_fano28_tlp_kernel<CW, STAGES, INWARP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  lr_grad_hess_fano28_tlp_kernel<CW, STAGES, INWARP><<<grid, 32 * (CW + (INWARP ? 0 : 1)), smem, s>>>(blk, beta, static_cast<double*>(ws));
  NUMS_LAUNCH_OK();
  lr_fold_kernel<<<(len + 7) / 8, 256, 0, s>>>(static_cast<const double*>(ws), grid, len, out);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

inline bool lr_fano_enabled() {    // NUMS_LR_FANO=0 selects the triangular cover (A/B measurements)
  static const bool on = []() { const char* v = getenv("NUMS_LR_FANO"); return !(v && v[0] == '0'); }();
  return on;
}

int launch_lr_fano28(const LrBlocks& blk, const double* beta, double* out, void* ws, size_t ws_bytes, cudaStream_t s) {
  static const int variant = []() { const char* v = getenv("NUMS_LR_TILE"); return v ? atoi(v) : 1202; }();
  switch (variant) {      // NUMS_LR_TILE = <consumer warps><stages, 2 digits>: A/B measurements
    case 1102: return launch_lr_fano28_tlp<11, 2, false>(blk, beta, out, ws, ws_bytes, s);
    case 803: return launch_lr_fano28_tlp<8, 3, false>(blk, beta, out, ws, ws_bytes, s);
    default: return launch_lr_fano28_tlp<12, 2, true>(blk, beta, out, ws, ws_bytes, s);
  }
}

template <int D>
int launch_lr_dense(const LrBlocks& blk, const double* beta, double* out, void* ws, size_t ws_bytes,
                    cudaStream_t s) {
  constexpr int d = D;
  constexpr int NB = (D + 7) / 8;
  const size_t tile_bytes = (size_t)kTileRows * d * sizeof(double);
  const size_t smem = kDenseStages * tile_bytes + (8 * 128 + NB * 8) * sizeof(double) + 2 * kDenseStages * sizeof(uint64_t);
  const int len = d + d * d;
  const int64_t ntiles = blk.tile_begin[blk.count];
  int grid = sm_count();
  if (grid > ntiles) grid = (int)ntiles;
  if (grid < 1) grid = 1;
  NUMS_NEED_WS((size_t)grid * len * sizeof(double), ws_bytes);
  NUMS_CUDA_OK(cudaFuncSetAttribute(lr_grad_hess_dense_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lr_grad_hess_dense_kernel<D><<<grid, kDenseThreads, smem, s>>>(blk, beta, static_cast<double*>(ws));
  NUMS_LAUNCH_OK();
  lr_fold_kernel<<<(len + 7) / 8, 256, 0, s>>>(static_cast<const double*>(ws), grid, len, out);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

// d for which the dense (bulk-copy) kernel applies: pitch d conflict free (d == 4 or 12 mod 16, which
// also leaves padding column d free for the gradient), 3 tiles + scratch fit, and the per-CTA fold
// (8 x (D8 + D8^2) doubles) fits in the ring.
bool lr_dense_ok(int64_t d) {
  if (!(d == 4 || d == 12 || d == 20 || d == 28 || d == 36 || d == 44)) return false;
  const int nb = (int)((d + 7) / 8);
  const size_t tile_bytes = (size_t)kTileRows * d * sizeof(double);
  const size_t smem = kDenseStages * tile_bytes + (8 * 128 + nb * 8 + 8) * sizeof(double);
  return smem <= 227 * 1024 && (size_t)8 * (nb * 8 + nb * 8 * nb * 8) * sizeof(double) <= kDenseStages * tile_bytes;
}

int dispatch_lr_dense(const LrBlocks& blk, const double* beta, int d, double* out, void* ws, size_t ws_bytes,
                      cudaStream_t s) {
  switch (d) {
    case 4: return launch_lr_dense<4>(blk, beta, out, ws, ws_bytes, s);
    case 12: return launch_lr_dense<12>(blk, beta, out, ws, ws_bytes, s);
    case 20: return launch_lr_dense<20>(blk, beta, out, ws, ws_bytes, s);
    case 28:    // NUMS_LR_FANO=0 selects the triangular cover (A/B measurements)
      return lr_fano_enabled() ? launch_lr_fano28(blk, beta, out, ws, ws_bytes, s)
                               : launch_lr_dense<28>(blk, beta, out, ws, ws_bytes, s);
    case 36: return launch_lr_dense<36>(blk, beta, out, ws, ws_bytes, s);
    case 44: return launch_lr_dense<44>(blk, beta, out, ws, ws_bytes, s);
  }
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "lr_grad_hess: d = %d", d);
}

template <int NB>
int launch_lr(const double* X, int64_t ldx, const double* y, const double* beta, int64_t n, int d,
              double* out, void* ws, size_t ws_bytes, cudaStream_t s) {
  constexpr int PITCH = lr_pitch(NB);
  if (ldx == d && lr_dense_ok(d)) {
    LrBlocks blk;
    memset(&blk, 0, sizeof(blk));
    blk.count = 1;
    blk.X[0] = X; blk.y[0] = y; blk.rows[0] = n;
    blk.tile_begin[0] = 0;
    blk.tile_begin[1] = (n + kTileRows - 1) / kTileRows;
    return dispatch_lr_dense(blk, beta, d, out, ws, ws_bytes, s);
  }
  const size_t stage_bytes = (size_t)kTileRows * PITCH * sizeof(double);
  int stages = 3;
  if (3 * stage_bytes > 220 * 1024) stages = 2;
  const size_t smem = stages * stage_bytes;
  NUMS_REQUIRE(smem <= 227 * 1024, "lr_grad_hess: d = %d needs too much shared memory", d);
  NUMS_REQUIRE((size_t)8 * (NB * 8 + NB * 8 * NB * 8) * sizeof(double) <= smem,
               "lr_grad_hess: reduction scratch does not fit");
  const int len = d + d * d;
  const int64_t ntiles = (n + kTileRows - 1) / kTileRows;
  int grid = sm_count();
  if (grid > ntiles) grid = (int)ntiles;
  if (grid < 1) grid = 1;
  NUMS_NEED_WS((size_t)grid * len * sizeof(double), ws_bytes);
  NUMS_CUDA_OK(cudaFuncSetAttribute(lr_grad_hess_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lr_grad_hess_kernel<NB><<<grid, kLrThreads, smem, s>>>(X, ldx, y, beta, n, d, stages, static_cast<double*>(ws));
  NUMS_LAUNCH_OK();
  lr_fold_kernel<<<(len + 7) / 8, 256, 0, s>>>(static_cast<const double*>(ws), grid, len, out);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

// Newton update of glms.newton (glms.py:362-372) as one single-CTA kernel: from the summed buffer
// g | H it solves H step = g (Gauss-Jordan on the augmented matrix, partial pivoting like LAPACK's
// LU behind np.linalg.inv), writes beta_out = beta - step and status = {max |g|, info}.  Replaces the
// inv / tensordot / sub / abs / max launches and the 4-byte pivot read-back of the interface path with
// one launch and one 16-byte read-back per iteration.
constexpr int kNewtonThreads = 256;
constexpr int kNewtonMaxD = 128;

// max that propagates NaN, like np.max
__device__ __forceinline__ double max_nan(double a, double b) { return a != a ? a : (b != b ? b : fmax(a, b)); }

__global__ void __launch_bounds__(kNewtonThreads, 1)
newton_step_kernel(int d, const double* __restrict__ gh, const double* __restrict__ beta,
                   double* __restrict__ beta_out, double* __restrict__ status) {
  extern __shared__ __align__(16) double aug[];   // d x (d + 2): [H | g], pitch d + 2
  __shared__ int piv_row;
  __shared__ int failed;
  __shared__ double red[kNewtonThreads / 32];
  const int pitch = d + 2;
  const int tid = threadIdx.x;
  if (tid == 0) failed = 0;
  for (int e = tid; e < d * d; e += kNewtonThreads) {
    const int i = e / d, c = e - i * d;
    aug[i * pitch + c] = gh[d + e];
  }
  double gmax = 0.0;
  for (int i = tid; i < d; i += kNewtonThreads) {
    const double g = gh[i];
    aug[i * pitch + d] = g;
    gmax = max_nan(gmax, fabs(g));
  }
  for (int o = 16; o > 0; o >>= 1) gmax = max_nan(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  if ((tid & 31) == 0) red[tid >> 5] = gmax;
  __syncthreads();
  for (int k = 0; k < d; ++k) {
    if (tid < 32) {     // pivot: first maximum of |aug[i][k]|, i >= k
      double best = -1.0;
      int bi = k;
      for (int i = k + tid; i < d; i += 32) {
        const double a = fabs(aug[i * pitch + k]);
        if (a > best) {
          best = a;
          bi = i;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      if (tid == 0) {
        piv_row = bi;
        if (!(best > 0.0) && failed == 0) failed = k + 1;
      }
    }
    __syncthreads();
    const int p = piv_row;
    if (p != k) {
      for (int c = k + tid; c <= d; c += kNewtonThreads) {
        const double a = aug[k * pitch + c];
        aug[k * pitch + c] = aug[p * pitch + c];
        aug[p * pitch + c] = a;
      }
      __syncthreads();
    }
    const double inv_pivot = 1.0 / aug[k * pitch + k];
    __syncthreads();
    for (int c = k + tid; c <= d; c += kNewtonThreads) aug[k * pitch + c] *= inv_pivot;
    __syncthreads();
    // eliminate column k from every other row (columns k+1 .. d; column k itself is not read again)
    const int width = d - k;
    for (int e = tid; e < d * width; e += kNewtonThreads) {
      const int i = e / width, c = k + 1 + (e - i * width);
      if (i != k) aug[i * pitch + c] = fma(-aug[i * pitch + k], aug[k * pitch + c], aug[i * pitch + c]);
    }
    __syncthreads();
  }
  for (int i = tid; i < d; i += kNewtonThreads) beta_out[i] = beta[i] - aug[i * pitch + d];
  if (tid == 0) {
    double top = red[0];
    for (int w = 1; w < kNewtonThreads / 32; ++w) top = max_nan(top, red[w]);
    status[0] = top;
    status[1] = (double)failed;
  }
}

// The same update for d <= 32 in ONE warp: lane i owns row i of the augmented matrix [H | g] (shared memory, odd
// pitch: conflict free), the pivot search is a shuffle butterfly and the warp never meets a CTA barrier (the kernel
// above spends about four per pivot: ~45 us at d = 28, more than half of what the fused gradient / Hessian kernel
// needs for one GPU's share of config 4 on eight GPUs).  Same operations in the same order as above (first maximum
// wins ties, fma elimination): bit-identical results.
__global__ void __launch_bounds__(32, 1)
newton_step_warp_kernel(int d, const double* __restrict__ gh, const double* __restrict__ beta,
                        double* __restrict__ beta_out, double* __restrict__ status) {
  constexpr int P = 35;                 // pitch: columns 0 .. d - 1 of H, the right-hand side in column d (d <= 32)
  __shared__ double aug[32 * P];
  const int lane = threadIdx.x;
  const bool live = lane < d;
  for (int e = lane; e < d * d; e += 32) {
    const int i = e / d, c = e - i * d;
    aug[i * P + c] = gh[d + e];
  }
  const double g = live ? gh[lane] : 0.0;
  if (live) aug[lane * P + d] = g;
  double gmax = live ? fabs(g) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gmax = max_nan(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  __syncwarp();
  double* mine = aug + lane * P;
  int failed = 0;
  for (int k = 0; k < d; ++k) {
    // first maximum of |a[i][k]|, i >= k; a NaN never wins a comparison (as in the kernel above: `a > best`)
    const double cand = (live && lane >= k) ? fabs(mine[k]) : -1.0;
    double best = cand > -1.0 ? cand : -1.0;
    int bi = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    best = __shfl_sync(0xffffffffu, best, 0);     // (every lane holds the same pair already; make it explicit)
    bi = __shfl_sync(0xffffffffu, bi, 0);
    if (!(best > 0.0)) {
      if (failed == 0) failed = k + 1;
      if (best < 0.0) bi = k;           // nothing but NaNs below the diagonal: keep the diagonal like the kernel above
    }
    const int p = bi;
    if (p != k) {                       // uniform: lanes swap the two rows column-wise
      for (int c = k + lane; c <= d; c += 32) {
        const double a = aug[k * P + c];
        aug[k * P + c] = aug[p * P + c];
        aug[p * P + c] = a;
      }
      __syncwarp();
    }
    const double inv_pivot = 1.0 / aug[k * P + k];
    __syncwarp();
    for (int c = k + lane; c <= d; c += 32) aug[k * P + c] *= inv_pivot;
    __syncwarp();
    if (live && lane != k) {
      const double f = mine[k];
      const double* piv = aug + k * P;
      for (int c = k + 1; c <= d; ++c) mine[c] = fma(-f, piv[c], mine[c]);
    }
    __syncwarp();
  }
  if (live) beta_out[lane] = beta[lane] - mine[d];
  if (lane == 0) {
    status[0] = gmax;
    status[1] = (double)failed;
  }
}

}  // namespace
}  // namespace nums

extern "C" int nums_newton_step(int64_t d, const double* gh, const double* beta, double* beta_out, double* status,
                                void* stream) {
  using namespace nums;
  NUMS_REQUIRE(d >= 1 && d <= kNewtonMaxD, "newton_step: d = %lld outside the supported range [1, %d]", (long long)d,
               kNewtonMaxD);
  NUMS_REQUIRE(gh && beta && beta_out && status, "newton_step: null pointer");
  if (d <= 32) {
    newton_step_warp_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>((int)d, gh, beta, beta_out, status);
    NUMS_LAUNCH_OK();
    return NUMS_OK;
  }
  const size_t smem = (size_t)d * (d + 2) * sizeof(double);
  NUMS_CUDA_OK(cudaFuncSetAttribute(newton_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  newton_step_kernel<<<1, kNewtonThreads, smem, static_cast<cudaStream_t>(stream)>>>((int)d, gh, beta, beta_out, status);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

extern "C" int nums_lr_grad_hess(int64_t n, int64_t d, const double* X, int64_t ldx, const double* y,
                                 const double* beta, double* out, void* ws, size_t ws_bytes,
                                 void* stream) {
  using namespace nums;
  NUMS_REQUIRE(n >= 1 && d >= 1, "lr_grad_hess: empty block");
  NUMS_REQUIRE(X && y && beta && out, "lr_grad_hess: null pointer");
  if (d > 48 || d % 2 != 0 || ldx % 2 != 0 || (reinterpret_cast<uintptr_t>(X) & 15u) != 0)
    NUMS_FAIL(NUMS_ERR_UNSUPPORTED,
              "lr_grad_hess: needs even d <= 48, even row pitch and a 16-byte aligned X (d=%lld, ldx=%lld)",
              (long long)d, (long long)ldx);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch ((d + 7) / 8) {
    case 1: return launch_lr<1>(X, ldx, y, beta, n, (int)d, out, ws, ws_bytes, s);
    case 2: return launch_lr<2>(X, ldx, y, beta, n, (int)d, out, ws, ws_bytes, s);
    case 3: return launch_lr<3>(X, ldx, y, beta, n, (int)d, out, ws, ws_bytes, s);
    case 4: return launch_lr<4>(X, ldx, y, beta, n, (int)d, out, ws, ws_bytes, s);
    case 5: return launch_lr<5>(X, ldx, y, beta, n, (int)d, out, ws, ws_bytes, s);
    case 6: return launch_lr<6>(X, ldx, y, beta, n, (int)d, out, ws, ws_bytes, s);
  }
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "lr_grad_hess: d = %lld", (long long)d);
}

extern "C" int nums_lr_grad_hess_blocks(int nblocks, const double* const* X_host, const double* const* y_host,
                                        const int64_t* rows_host, int64_t d, const double* beta, double* out,
                                        void* ws, size_t ws_bytes, void* stream) {
  using namespace nums;
  NUMS_REQUIRE(nblocks >= 1 && nblocks <= kMaxLrBlocks, "lr_grad_hess_blocks: 1..%d blocks per call", kMaxLrBlocks);
  NUMS_REQUIRE(X_host && y_host && rows_host && beta && out, "lr_grad_hess_blocks: null pointer");
  if (!lr_dense_ok(d))
    NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "lr_grad_hess_blocks: d = %lld is not served by the dense kernel", (long long)d);
  LrBlocks blk;
  memset(&blk, 0, sizeof(blk));
  blk.count = nblocks;
  int64_t tiles = 0;
  for (int b = 0; b < nblocks; ++b) {
    NUMS_REQUIRE(rows_host[b] >= 1 && X_host[b] && y_host[b], "lr_grad_hess_blocks: block %d is empty", b);
    NUMS_REQUIRE((reinterpret_cast<uintptr_t>(X_host[b]) & 15u) == 0, "lr_grad_hess_blocks: X block %d is not 16-byte aligned", b);
    blk.X[b] = X_host[b]; blk.y[b] = y_host[b]; blk.rows[b] = rows_host[b];
    blk.tile_begin[b] = tiles;
    tiles += (rows_host[b] + kTileRows - 1) / kTileRows;
  }
  blk.tile_begin[nblocks] = tiles;
  return dispatch_lr_dense(blk, beta, (int)d, out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
