// Dense factorisations of one block -- replaces np.linalg.qr / inv / cholesky / svd at
// nums/core/systems/numpy_compute.py:240-257 (LAPACK geqrf, getrf+getri, potrf, gesdd in the
// reference).
//
// QR (the TSQR leaf, application.py:784-814): a streaming Householder factorisation.  Every CTA
// keeps the current n x n triangle R in shared memory and annihilates one chunk of rows after
// another against it; annihilating column j of [R; chunk] only couples row j of R with the
// chunk (R is upper triangular), so a step costs 2 c (n - j) FMA for a c-row chunk and the
// whole leaf costs the textbook 2 m n^2.  The per-CTA triangles are then stacked and reduced by
// the same kernel (a tree whose fan-in is 4) until one R is left.  This is the exact
// Householder R of the block (up to row signs), backward stable for any conditioning.
//
// inv / cholesky: single-CTA in-place elimination in shared memory (global scratch beyond
// ~160 x 160): the matrices on the path are the 28 x 28 LR Hessian and the 128 x 128 TSQR R,
// so this is latency, not throughput.
#include <cstdlib>
#include "common.cuh"

namespace nums {
namespace {

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// =====================================================================================
// streaming Householder R
// =====================================================================================
constexpr int kQrThreads = 256;
constexpr int kQrWarps = kQrThreads / 32;

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// CR = chunk rows per lane (chunk = 32 * CR rows).
// A: row-major (m x n, pitch lda).  CTA b handles rows [b * rows_per_cta, (b+1) * rows_per_cta).
// Rout: per-CTA dense n x n triangles (zeros below the diagonal), CTA b at Rout + b * n * n.
template <typename T, int CR>
__global__ void __launch_bounds__(kQrThreads, 1)
tsqr_stream_kernel(const T* __restrict__ A, int64_t lda, int64_t m, int n, int64_t rows_per_cta,
                   T* __restrict__ Rout, int64_t r_rows, int64_t ldr) {
  extern __shared__ __align__(16) unsigned char qr_smem[];
  T* R = reinterpret_cast<T*>(qr_smem);
  const int pitch = n | 1;              // odd pitch: column walks are bank-conflict free
  T* C = R + (size_t)n * pitch;         // chunk, 32*CR rows
  constexpr int CH = 32 * CR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < n * pitch; i += kQrThreads) R[i] = T(0);

  const int64_t row_lo = (int64_t)blockIdx.x * rows_per_cta;
  int64_t row_hi = row_lo + rows_per_cta;
  if (row_hi > m) row_hi = m;

  for (int64_t base = row_lo; base < row_hi; base += CH) {
    __syncthreads();
    // coalesced load of the chunk, zero rows past the end
    for (int e = threadIdx.x; e < CH * n; e += kQrThreads) {
      const int i = e / n, c = e - i * n;
      const int64_t gr = base + i;
      C[i * pitch + c] = gr < row_hi ? A[gr * lda + c] : T(0);
    }
    __syncthreads();
    for (int j = 0; j < n; ++j) {
      // Householder vector for column j of [R(j,j); chunk(:, j)] (LAPACK dlarfg), computed
      // redundantly by every warp so no broadcast is needed.
      T v[CR];
      T sigma = T(0);
#pragma unroll
      for (int r = 0; r < CR; ++r) {
        v[r] = C[(lane + 32 * r) * pitch + j];
        sigma += v[r] * v[r];
      }
      sigma = warp_sum(sigma);
      const T alpha = R[j * pitch + j];
      T tau = T(0), beta = alpha;
      if (sigma != T(0)) {
        const T nrm = sqrt(alpha * alpha + sigma);
        beta = alpha >= T(0) ? -nrm : nrm;
        tau = (beta - alpha) / beta;
        const T scale = T(1) / (alpha - beta);
#pragma unroll
        for (int r = 0; r < CR; ++r) v[r] *= scale;
      }
      if (tau != T(0)) {
        // apply H = I - tau [1; v][1; v]^T to the trailing columns, 2 columns per warp in flight
        int c = j + 1 + warp;
        for (; c + kQrWarps < n; c += 2 * kQrWarps) {
          const int c2 = c + kQrWarps;
          T w0 = T(0), w1 = T(0);
          T x0[CR], x1[CR];
#pragma unroll
          for (int r = 0; r < CR; ++r) {
            x0[r] = C[(lane + 32 * r) * pitch + c];
            x1[r] = C[(lane + 32 * r) * pitch + c2];
            w0 += v[r] * x0[r];
            w1 += v[r] * x1[r];
          }
          w0 = warp_sum(w0);
          w1 = warp_sum(w1);
          const T r0 = R[j * pitch + c], r1 = R[j * pitch + c2];
          const T t0 = tau * (w0 + r0), t1 = tau * (w1 + r1);
#pragma unroll
          for (int r = 0; r < CR; ++r) {
            C[(lane + 32 * r) * pitch + c] = x0[r] - t0 * v[r];
            C[(lane + 32 * r) * pitch + c2] = x1[r] - t1 * v[r];
          }
          if (lane == 0) {
            R[j * pitch + c] = r0 - t0;
            R[j * pitch + c2] = r1 - t1;
          }
        }
        if (c < n) {
          T w0 = T(0);
          T x0[CR];
#pragma unroll
          for (int r = 0; r < CR; ++r) {
            x0[r] = C[(lane + 32 * r) * pitch + c];
            w0 += v[r] * x0[r];
          }
          w0 = warp_sum(w0);
          const T r0 = R[j * pitch + c];
          const T t0 = tau * (w0 + r0);
#pragma unroll
          for (int r = 0; r < CR; ++r) C[(lane + 32 * r) * pitch + c] = x0[r] - t0 * v[r];
          if (lane == 0) R[j * pitch + c] = r0 - t0;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) R[j * pitch + j] = beta;
    }
  }
  __syncthreads();
  T* out = Rout + (size_t)blockIdx.x * r_rows * ldr;
  for (int e = threadIdx.x; e < (int)r_rows * n; e += kQrThreads) {
    const int i = e / n, c = e - i * n;
    out[(size_t)i * ldr + c] = (c >= i) ? R[i * pitch + c] : T(0);
  }
}

template <typename T, int CR>
int launch_tsqr(const T* A, int64_t lda, int64_t m, int n, int64_t rows_per_cta, int ctas, T* Rout,
                int64_t r_rows, int64_t ldr, cudaStream_t s) {
  const size_t smem = (size_t)(n + 32 * CR) * (n | 1) * sizeof(T);
  NUMS_CUDA_OK(cudaFuncSetAttribute(tsqr_stream_kernel<T, CR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tsqr_stream_kernel<T, CR><<<ctas, kQrThreads, smem, s>>>(A, lda, m, n, rows_per_cta, Rout, r_rows, ldr);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

template <typename T>
int launch_tsqr_auto(const T* A, int64_t lda, int64_t m, int n, int64_t rows_per_cta, int ctas, T* Rout,
                     int64_t r_rows, int64_t ldr, cudaStream_t s) {
  const size_t budget = 200 * 1024;
  const size_t row_bytes = (size_t)(n | 1) * sizeof(T);
  if ((size_t)(n + 128) * row_bytes <= budget) return launch_tsqr<T, 4>(A, lda, m, n, rows_per_cta, ctas, Rout, r_rows, ldr, s);
  if ((size_t)(n + 64) * row_bytes <= budget) return launch_tsqr<T, 2>(A, lda, m, n, rows_per_cta, ctas, Rout, r_rows, ldr, s);
  if ((size_t)(n + 32) * row_bytes <= 227 * 1024) return launch_tsqr<T, 1>(A, lda, m, n, rows_per_cta, ctas, Rout, r_rows, ldr, s);
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "qr: %d columns do not fit the shared-memory TSQR leaf", n);
}

// =====================================================================================
// streaming Householder R, blocked: WY panels + FP64 tensor-pipe trailing updates (n <= 128)
// =====================================================================================
// Same factorization as tsqr_stream_kernel -- the running triangle R stays in shared memory and every
// 128-row chunk C is annihilated against it; reflector j couples only row j of R with the chunk -- but the
// reflectors are generated 8 at a time (one warp, the 128 x 8 panel held in registers, no CTA barrier inside a
// panel) and applied to everything right of the panel as one compact-WY block reflector
//     Q^T = H_7 ... H_0 = I - V T^T V^T,   V = [I_8 ; Vc]   (rows j0..j0+7 of R; the chunk rows)
//     W = R[J, K] + Vc^T C[:, K];   Y = -T^T W;   R[J, K] += Y;   C[:, K] += Vc Y
// whose two products run on the FP64 tensor pipe (mma.sync m8n8k4 -> DMMA) straight out of shared memory.
// T follows LAPACK dlarft (forward, columnwise) with V^T V = I + Vc^T Vc.  The unblocked kernel spends its time
// in 128 CTA-wide barriers and ~16 dependent warp reductions per column; here a chunk costs 16 panel
// factorizations (8 short steps each, one warp) plus 2 m n^2 flops of DMMA work.  Input may be float32 or
// float64; the arithmetic is float64 either way.  Columns are padded to a multiple of 8 with zeros (a zero
// column yields tau = 0, i.e. H = I).
constexpr int kWyThreads = 256;
constexpr int kWyChunk = 128;
constexpr int kWyVPitch = 12;   // 12 = 12 (mod 16): 4 x 4 fragment reads of Vc hit 16 distinct 8-byte banks

__device__ __forceinline__ void qr_dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

__host__ __device__ inline int wy_roff(int i, int np) { return i * np - (i * (i - 1)) / 2; }   // packed upper triangle
inline size_t wy_smem_bytes(int np) {
  const size_t tri = ((size_t)np * (np + 1) / 2 + 1) & ~(size_t)1;
  const size_t pc = np + 4;
  return (tri + (size_t)kWyChunk * pc + (size_t)kWyChunk * kWyVPitch + 2 * 8 * pc + 64) * sizeof(double);
}

template <typename TIN>
__global__ void __launch_bounds__(kWyThreads, 1)
tsqr_wy_kernel(const TIN* __restrict__ A, int64_t lda, int64_t m, int n, int np, int64_t rows_per_cta,
               TIN* __restrict__ Rout, int64_t r_rows, int64_t ldr) {
  extern __shared__ __align__(16) unsigned char wy_smem[];
  const int pc = np + 4;                                   // = 4 or 12 (mod 16): conflict-free 4 x 4 fragment reads
  const int tri = (np * (np + 1) / 2 + 1) & ~1;
  double* Rp = reinterpret_cast<double*>(wy_smem);         // packed upper triangle of the np x np factor
  double* C = Rp + tri;                                    // chunk, kWyChunk x np
  double* V = C + kWyChunk * pc;                           // Vc of the current panel, kWyChunk x 8
  double* W = V + kWyChunk * kWyVPitch;                    // 8 x np
  double* Y = W + 8 * pc;                                  // 8 x np
  double* Tm = Y + 8 * pc;                                 // 8 x 8
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < tri; i += kWyThreads) Rp[i] = 0.0;

  const int64_t row_lo = (int64_t)blockIdx.x * rows_per_cta;
  int64_t row_hi = row_lo + rows_per_cta;
  if (row_hi > m) row_hi = m;

  for (int64_t base = row_lo; base < row_hi; base += kWyChunk) {
    __syncthreads();
    // chunk load: 8 independent global loads in flight per thread (consecutive threads read consecutive
    // elements of a row), zero fill past the last row / column
    {
      const int total = kWyChunk * np;
      for (int e0 = threadIdx.x; e0 < total; e0 += kWyThreads * 8) {
        double val[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = e0 + u * kWyThreads;
          const int i = e / np, c = e - i * np;
          const int64_t gr = base + i;
          val[u] = (e < total && gr < row_hi && c < n) ? (double)A[gr * lda + c] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = e0 + u * kWyThreads;
          if (e < total) {
            const int i = e / np, c = e - i * np;
            C[i * pc + c] = val[u];
          }
        }
      }
    }
    __syncthreads();
    for (int j0 = 0; j0 < np; j0 += 8) {
      if (warp == 0) {
        // ---- panel factorization: rows lane, lane+32, lane+64, lane+96 of the 128 x 8 panel in registers.
        // Everything a step needs lives in registers (the panel, row j of R, this lane's row of T), so the 8
        // steps run without a single barrier or shared-memory round trip; the reductions of a step (7 - t
        // projections onto the other panel columns, t inner products with the earlier reflectors) are
        // independent butterflies that the scheduler interleaves.
        double x[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int k = 0; k < 8; ++k) x[r][k] = C[(lane + 32 * r) * pc + j0 + k];
        double trow[8];                                       // lane s < 8 keeps row s of T
#pragma unroll
        for (int q = 0; q < 8; ++q) trow[q] = 0.0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          double* Rrow = Rp + wy_roff(j0 + t, np);           // Rrow[k - t] = R[j0 + t][j0 + k]
          double rr[8];
#pragma unroll
          for (int k = t; k < 8; ++k) rr[k] = Rrow[k - t];
          // One batch of independent butterflies: p[q] = x_t . (column q).  p[t] is the squared norm that
          // fixes the reflector; the others are the projections, needed only after scaling (v = scale x_t), so
          // their reduction latency overlaps the sqrt / reciprocal chain instead of following it.
          double p[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            p[q] = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r) p[q] += x[r][t] * x[r][q];
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) p[q] = warp_sum(p[q]);
          const double sigma = p[t];
          const double alpha = rr[t];
          double tau = 0.0, beta = alpha, scale = 0.0;
          if (sigma != 0.0) {                                 // LAPACK dlarfg
            const double nrm = sqrt(alpha * alpha + sigma);
            beta = alpha >= 0.0 ? -nrm : nrm;
            const double inv_beta = 1.0 / beta;
            scale = 1.0 / (alpha - beta);
            tau = (beta - alpha) * inv_beta;
          }
          double v[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) v[r] = x[r][t] * scale;
          double d[8];                                        // v . (column q): q < t -> Vc^T Vc, q > t -> projections
#pragma unroll
          for (int q = 0; q < 8; ++q) d[q] = scale * p[q];
#pragma unroll
          for (int k = t + 1; k < 8; ++k) {                   // H_t on the rest of the panel
            const double tw = tau * (d[k] + rr[k]);
#pragma unroll
            for (int r = 0; r < 4; ++r) x[r][k] -= tw * v[r];
            rr[k] -= tw;
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) x[r][t] = v[r];
          // T(0:t-1, t) = -tau T(0:t-1, 0:t-1) (Vc^T Vc)(0:t-1, t) ; T(t, t) = tau     (dlarft, forward columnwise)
          double acc = 0.0;
#pragma unroll
          for (int u = 0; u < t; ++u)
            if (u >= lane) acc += trow[u] * d[u];
          trow[t] = lane < t ? -tau * acc : (lane == t ? tau : 0.0);
          if (lane == 0) {
            Rrow[0] = beta;
#pragma unroll
            for (int k = t + 1; k < 8; ++k) Rrow[k - t] = rr[k];
          }
        }
        if (lane < 8) {
#pragma unroll
          for (int q = 0; q < 8; ++q) Tm[lane * 8 + q] = trow[q];
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int k = 0; k < 8; ++k) V[(lane + 32 * r) * kWyVPitch + k] = x[r][k];
      }
      __syncthreads();
      const int k_first = j0 + 8;
      const int ntiles = (np - k_first) >> 3;
      if (ntiles > 0) {
        const int fr = lane & 3, fc = lane >> 2;              // fragment coordinates of this lane
        // ---- W = R[J, K] + Vc^T C[:, K]: one 8-column tile per warp at a time, 128-deep contraction
        for (int tile = warp; tile < ntiles; tile += kWyThreads / 32) {
          const int k0 = k_first + 8 * tile;
          double acc0[2] = {0.0, 0.0}, acc1[2] = {0.0, 0.0};
#pragma unroll 4
          for (int r0 = 0; r0 < kWyChunk; r0 += 8) {
            qr_dmma884(acc0, V[(r0 + fr) * kWyVPitch + fc], C[(r0 + fr) * pc + k0 + fc]);
            qr_dmma884(acc1, V[(r0 + 4 + fr) * kWyVPitch + fc], C[(r0 + 4 + fr) * pc + k0 + fc]);
          }
          const int t = fc, kk = k0 + 2 * fr;                 // accumulator holds W[t][kk], W[t][kk + 1]
          const double* Rrow = Rp + wy_roff(j0 + t, np) - (j0 + t);
          W[t * pc + kk] = acc0[0] + acc1[0] + Rrow[kk];
          W[t * pc + kk + 1] = acc0[1] + acc1[1] + Rrow[kk + 1];
        }
        __syncthreads();
        // ---- Y = -T^T W ; R[J, K] += Y
        const int ncols = np - k_first;
        for (int e = threadIdx.x; e < 8 * ncols; e += kWyThreads) {
          const int t = e / ncols, k = k_first + (e - t * ncols);
          double y = 0.0;
          for (int q = 0; q <= t; ++q) y += Tm[q * 8 + t] * W[q * pc + k];
          Y[t * pc + k] = -y;
          Rp[wy_roff(j0 + t, np) + k - (j0 + t)] -= y;
        }
        __syncthreads();
        // ---- C[:, K] += Vc Y: 16 row tiles, two per warp; the two A fragments of a row tile are loaded once
        for (int mt = warp; mt < kWyChunk / 8; mt += kWyThreads / 32) {
          const int m0 = 8 * mt;
          const double a0 = V[(m0 + fc) * kWyVPitch + fr], a1 = V[(m0 + fc) * kWyVPitch + 4 + fr];
          for (int tile = 0; tile < ntiles; ++tile) {
            const int k0 = k_first + 8 * tile;
            double2* cp = reinterpret_cast<double2*>(&C[(m0 + fc) * pc + k0 + 2 * fr]);
            double2 cv = *cp;
            double acc[2] = {cv.x, cv.y};
            qr_dmma884(acc, a0, Y[fr * pc + k0 + fc]);
            qr_dmma884(acc, a1, Y[(4 + fr) * pc + k0 + fc]);
            cv.x = acc[0];
            cv.y = acc[1];
            *cp = cv;
          }
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  TIN* out = Rout + (size_t)blockIdx.x * r_rows * ldr;
  for (int i = warp; i < (int)r_rows; i += kWyThreads / 32) {
    const double* Rrow = Rp + wy_roff(i, np) - i;
    for (int c = lane; c < n; c += 32) out[(size_t)i * ldr + c] = (c >= i) ? (TIN)Rrow[c] : TIN(0);
  }
}

template <typename T>
int launch_tsqr_wy(const T* A, int64_t lda, int64_t m, int n, int64_t rows_per_cta, int ctas, T* Rout, int64_t r_rows,
                   int64_t ldr, cudaStream_t s) {
  const int np = (n + 7) & ~7;
  const size_t smem = wy_smem_bytes(np);
  NUMS_CUDA_OK(cudaFuncSetAttribute(tsqr_wy_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tsqr_wy_kernel<T><<<ctas, kWyThreads, smem, s>>>(A, lda, m, n, np, rows_per_cta, Rout, r_rows, ldr);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

bool qr_use_wy(int n) {
  static const bool disabled = [] { const char* e = getenv("NUMS_QR_UNBLOCKED"); return e && e[0] == '1'; }();
  return !disabled && n <= 128;
}

template <typename T>
int launch_leaf(const T* A, int64_t lda, int64_t m, int n, int64_t rows_per_cta, int ctas, T* Rout, int64_t r_rows,
                int64_t ldr, cudaStream_t s) {
  if (qr_use_wy(n)) return launch_tsqr_wy<T>(A, lda, m, n, rows_per_cta, ctas, Rout, r_rows, ldr, s);
  return launch_tsqr_auto<T>(A, lda, m, n, rows_per_cta, ctas, Rout, r_rows, ldr, s);
}

template <typename T>
int run_qr_r(int64_t m, int64_t n64, const T* A, int64_t lda, T* R, int64_t ldr, void* ws, size_t ws_bytes,
             cudaStream_t s) {
  const int n = (int)n64;
  const int64_t k = m < n ? m : n;
  // Leaf: at least 4n rows per CTA (so the leaf dominates the merges), at most one CTA per SM.
  int64_t min_rows = 4 * (int64_t)n;
  if (min_rows < 256) min_rows = 256;
  int64_t ctas = ceil_div(m, min_rows);
  if (ctas > sm_count()) ctas = sm_count();
  if (ctas < 1) ctas = 1;
  int64_t rows_per_cta = ceil_div(m, ctas);
  ctas = ceil_div(m, rows_per_cta);
  if (ctas == 1) return launch_leaf<T>(A, lda, m, n, rows_per_cta, 1, R, k, ldr, s);
  // ping-pong buffers of stacked n x n triangles
  const size_t level0 = (size_t)ctas * n * n * sizeof(T);
  const size_t level1 = (size_t)ceil_div(ctas, 4) * n * n * sizeof(T);
  const size_t off1 = (level0 + 255) & ~(size_t)255;
  NUMS_NEED_WS(off1 + level1, ws_bytes);
  T* buf0 = static_cast<T*>(ws);
  T* buf1 = reinterpret_cast<T*>(static_cast<char*>(ws) + off1);
  if (int rc = launch_leaf<T>(A, lda, m, n, rows_per_cta, (int)ctas, buf0, n, n, s)) return rc;
  T* src = buf0;
  T* dst = buf1;
  int64_t count = ctas;
  while (count > 1) {
    const int64_t next = ceil_div(count, 4);
    const int64_t rows = count * n;
    if (next == 1) return launch_leaf<T>(src, n, rows, n, rows, 1, R, k, ldr, s);
    if (int rc = launch_leaf<T>(src, n, rows, n, 4 * (int64_t)n, (int)next, dst, n, n, s)) return rc;
    T* tmp = src;
    src = dst;
    dst = tmp;
    count = next;
  }
  return NUMS_OK;
}

// =====================================================================================
// inverse (Gauss-Jordan, partial pivoting) and Cholesky, one CTA
// =====================================================================================
constexpr int kLaThreads = 1024;

template <typename T>
__global__ void __launch_bounds__(kLaThreads, 1)
inv_kernel(const T* __restrict__ A, int64_t lda, int n, T* __restrict__ out, int64_t ldo, T* scratch,
           int use_smem, int32_t* info) {
  extern __shared__ __align__(16) unsigned char la_smem[];
  const int pitch = use_smem ? (n | 1) : n;
  T* M = use_smem ? reinterpret_cast<T*>(la_smem) : scratch;
  __shared__ int piv_row;
  __shared__ int perm[1024];
  __shared__ T colk[1024];
  __shared__ int failed;
  const int tid = threadIdx.x;
  if (tid == 0) failed = 0;
  for (int e = tid; e < n * n; e += kLaThreads) {
    const int i = e / n, c = e - i * n;
    M[i * pitch + c] = A[(int64_t)i * lda + c];
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    // pivot search by warp 0: first maximum of |M[i][k]|, i >= k
    if (tid < 32) {
      T best = T(-1);
      int bi = k;
      for (int i = k + tid; i < n; i += 32) {
        const T a = fabs(M[i * pitch + k]);
        if (a > best) {
          best = a;
          bi = i;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const T ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      if (tid == 0) {
        piv_row = bi;
        perm[k] = bi;
        if (!(best > T(0)) && failed == 0) failed = k + 1;
      }
    }
    __syncthreads();
    const int p = piv_row;
    if (p != k) {
      for (int c = tid; c < n; c += kLaThreads) {
        const T a = M[k * pitch + c];
        M[k * pitch + c] = M[p * pitch + c];
        M[p * pitch + c] = a;
      }
    }
    __syncthreads();
    const T pivot = M[k * pitch + k];
    const T rp = T(1) / pivot;
    // save column k, then overwrite it with the unit column (in-place inversion)
    for (int i = tid; i < n; i += kLaThreads) colk[i] = M[i * pitch + k];
    __syncthreads();
    for (int i = tid; i < n; i += kLaThreads) M[i * pitch + k] = (i == k) ? T(1) : T(0);
    __syncthreads();
    for (int c = tid; c < n; c += kLaThreads) M[k * pitch + c] *= rp;
    __syncthreads();
    for (int e = tid; e < n * n; e += kLaThreads) {
      const int i = e / n, c = e - i * n;
      if (i != k) M[i * pitch + c] -= colk[i] * M[k * pitch + c];
    }
    __syncthreads();
  }
  // undo the row interchanges as column interchanges, in reverse order
  for (int k = n - 1; k >= 0; --k) {
    const int p = perm[k];
    if (p != k) {
      for (int i = tid; i < n; i += kLaThreads) {
        const T a = M[i * pitch + k];
        M[i * pitch + k] = M[i * pitch + p];
        M[i * pitch + p] = a;
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < n * n; e += kLaThreads) {
    const int i = e / n, c = e - i * n;
    out[(int64_t)i * ldo + c] = M[i * pitch + c];
  }
  if (tid == 0 && info != nullptr) *info = failed;
}

template <typename T>
__global__ void __launch_bounds__(kLaThreads, 1)
cholesky_kernel(const T* __restrict__ A, int64_t lda, int n, T* __restrict__ out, int64_t ldo, T* scratch,
                int use_smem, int32_t* info) {
  extern __shared__ __align__(16) unsigned char la_smem[];
  const int pitch = use_smem ? (n | 1) : n;
  T* M = use_smem ? reinterpret_cast<T*>(la_smem) : scratch;
  __shared__ int failed;
  const int tid = threadIdx.x;
  if (tid == 0) failed = 0;
  for (int e = tid; e < n * n; e += kLaThreads) {
    const int i = e / n, c = e - i * n;
    M[i * pitch + c] = A[(int64_t)i * lda + c];
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const T dkk = M[k * pitch + k];
    if (!(dkk > T(0))) {
      if (tid == 0 && failed == 0) failed = k + 1;
    }
    const T l = sqrt(dkk);
    __syncthreads();
    for (int i = k + tid; i < n; i += kLaThreads) M[i * pitch + k] = (i == k) ? l : M[i * pitch + k] / l;
    __syncthreads();
    const int rem = n - k - 1;
    for (int e = tid; e < rem * rem; e += kLaThreads) {
      const int i = k + 1 + e / rem, c = k + 1 + e % rem;
      if (c <= i) M[i * pitch + c] -= M[i * pitch + k] * M[c * pitch + k];
    }
    __syncthreads();
  }
  for (int e = tid; e < n * n; e += kLaThreads) {
    const int i = e / n, c = e - i * n;
    out[(int64_t)i * ldo + c] = (c <= i) ? M[i * pitch + c] : T(0);
  }
  if (tid == 0 && info != nullptr) *info = failed;
}


// =====================================================================================
// Gram-matrix QR, small-matrix stage: Cholesky + triangular inverse + condition bound, one launch
// =====================================================================================
// For the Gram path of qr (cuda_compute.qr_r_ex): given G = A^T A (n <= 128) produce L = chol(G),
// R = L^T, L^-1 and stats = {info, |L|_1, |L|_inf, |L^-1|_1, |L^-1|_inf}, the four norms of the
// rigorous bound cond_2(A) <= sqrt(|L|_1 |L|_inf |L^-1|_1 |L^-1|_inf).  Everything lives in one
// n x (n+1) shared-memory array: L in the lower triangle, (L^-1)^T in the strictly upper part
// shifted right by one column.  Four lanes share each row / column dot product, so the 2n
// dependent steps cost a few hundred cycles each: ~60 us instead of the ~1.1 ms the separate
// cholesky + general inverse + norm-reduction launches took.
constexpr int kGfMaxN = 128;
constexpr int kGfThreads = 4 * kGfMaxN;

__global__ void __launch_bounds__(kGfThreads, 1)
gram_factor_kernel(const double* __restrict__ G, int64_t ldg, int n, double* __restrict__ L, int64_t ldl,
                   double* __restrict__ R, int64_t ldr, double* __restrict__ Linv, int64_t ldi,
                   double* __restrict__ stats) {
  extern __shared__ __align__(16) double gf[];
  __shared__ double pivot;
  __shared__ double sums[4][kGfMaxN];
  const int pitch = n + 1;
  const int tid = threadIdx.x, row = tid >> 2, q = tid & 3;
  for (int e = tid; e < n * n; e += kGfThreads) {
    const int i = e / n, c = e - i * n;
    if (c <= i) gf[i * pitch + c] = G[(int64_t)i * ldg + c];
  }
  __syncthreads();

  // ---- left-looking Cholesky: column k of L from rows k.. of (G - L[:, :k] L[k, :k]^T) ------------------
  int failed = 0;
  for (int k = 0; k < n; ++k) {
    const bool active = row >= k && row < n;
    double acc = 0.0;
    if (active) {
      const double* li = gf + row * pitch;
      const double* lk = gf + k * pitch;
      for (int j = q; j < k; j += 4) acc = fma(li[j], lk[j], acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    const double v = active ? gf[row * pitch + k] - acc : 0.0;
    if (row == k && q == 0) pivot = v;
    __syncthreads();
    const double pv = pivot;
    if (!(pv > 0.0)) {      // same value in every thread: uniform exit
      failed = k + 1;
      break;
    }
    const double d = sqrt(pv);
    if (active && q == 0) gf[row * pitch + k] = (row == k) ? d : v / d;
    __syncthreads();
  }
  if (failed) {
    if (tid == 0) {
      stats[0] = (double)failed;
      stats[1] = stats[2] = stats[3] = stats[4] = 0.0;
    }
    return;
  }

  // ---- X = L^-1 by forward substitution, one column per 4-lane group; X[i][j] -> gf[j][i + 1] -----------
  {
    const int col = row;                       // column of X owned by this group
    const int first = (tid >> 5) * 8;          // smallest column in this warp: uniform loop bounds
    for (int i = first; i < n; ++i) {
      double acc = 0.0;
      const bool mine = col < n && i > col;
      if (mine) {
        const double* li = gf + i * pitch;
        const double* xj = gf + col * pitch + 1;
        for (int k = col + q; k < i; k += 4) acc = fma(li[k], xj[k], acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (col < n && i >= col && q == 0) {
        const double lii = gf[i * pitch + i];
        gf[col * pitch + i + 1] = (i == col) ? 1.0 / lii : -acc / lii;
      }
      __syncwarp();
    }
  }
  __syncthreads();

  // ---- norms: row and column abs-sums of L and X ---------------------------------------------------------
  {
    const int what = tid / n, t = tid - what * n;   // 0: L rows, 1: L columns, 2: X rows, 3: X columns
    if (what < 4) {
      double acc = 0.0;
      if (what == 0) for (int j = 0; j <= t; ++j) acc += fabs(gf[t * pitch + j]);
      else if (what == 1) for (int i = t; i < n; ++i) acc += fabs(gf[i * pitch + t]);
      else if (what == 2) for (int j = 0; j <= t; ++j) acc += fabs(gf[j * pitch + t + 1]);
      else for (int i = t; i < n; ++i) acc += fabs(gf[t * pitch + i + 1]);
      sums[what][t] = acc;
    }
  }
  __syncthreads();
  if (tid < 4) {
    double top = 0.0;
    for (int t = 0; t < n; ++t) top = fmax(top, sums[tid][t]);
    // |M|_1 = max column sum, |M|_inf = max row sum
    const int slot = tid == 0 ? 2 : tid == 1 ? 1 : tid == 2 ? 4 : 3;
    stats[slot] = top;
    if (tid == 0) stats[0] = 0.0;
  }
  for (int e = tid; e < n * n; e += kGfThreads) {
    const int i = e / n, c = e - i * n;
    const double l = (c <= i) ? gf[i * pitch + c] : 0.0;           // L[i][c]
    const double lt = (c >= i) ? gf[c * pitch + i] : 0.0;          // R[i][c] = L[c][i]
    const double x = (c <= i) ? gf[c * pitch + i + 1] : 0.0;       // X[i][c]
    if (L) L[(int64_t)i * ldl + c] = l;
    if (R) R[(int64_t)i * ldr + c] = lt;
    if (Linv) Linv[(int64_t)i * ldi + c] = x;
  }
}

// =====================================================================================
// SVD of a square matrix: one-sided Jacobi (Hestenes), one CTA
// =====================================================================================
// Works on At (row j = column j of A) and Vt (row j = column j of V) in global scratch (L2
// resident for the n <= 1024 this path sees: svd is only ever applied to the n x n R factor,
// application.py:946).  Each round of the round-robin ordering rotates n/2 disjoint column
// pairs, one warp per pair; sweeps repeat until no pair needed a rotation.
template <typename T>
__global__ void __launch_bounds__(kLaThreads, 1)
svd_jacobi_kernel(const T* __restrict__ A, int64_t lda, int n, T* __restrict__ U, T* __restrict__ S,
                  T* __restrict__ Vt_out, T* __restrict__ At, T* __restrict__ Vt, T* __restrict__ sig) {
  __shared__ int rotated;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kLaThreads / 32;
  for (int e = tid; e < n * n; e += kLaThreads) {
    const int j = e / n, i = e - j * n;   // At[j][i] = A[i][j]
    At[e] = A[(int64_t)i * lda + j];
    Vt[e] = (i == j) ? T(1) : T(0);
  }
  __syncthreads();
  const int np = n + (n & 1);             // pad to even with a dummy column
  const T eps = std::is_same<T, double>::value ? T(1.1e-16) : T(6e-8);
  for (int sweep = 0; sweep < 60; ++sweep) {
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int round = 0; round < np - 1; ++round) {
      for (int k = warp; k < np / 2; k += kWarps) {
        int p, q;
        if (k == 0) { p = np - 1; q = round; }
        else { p = (round + k) % (np - 1); q = (round - k + np - 1) % (np - 1); }
        if (p >= n || q >= n) continue;   // dummy column
        if (p > q) { const int tmp = p; p = q; q = tmp; }
        T* ap = At + (size_t)p * n;
        T* aq = At + (size_t)q * n;
        T alpha = T(0), beta = T(0), gamma = T(0);
        for (int i = lane; i < n; i += 32) {
          const T x = ap[i], y = aq[i];
          alpha += x * x; beta += y * y; gamma += x * y;
        }
        alpha = warp_sum(alpha); beta = warp_sum(beta); gamma = warp_sum(gamma);
        if (fabs(gamma) > eps * sqrt(alpha * beta) && alpha > T(0) && beta > T(0)) {
          const T zeta = (beta - alpha) / (T(2) * gamma);
          const T tt = (zeta >= T(0) ? T(1) : T(-1)) / (fabs(zeta) + sqrt(T(1) + zeta * zeta));
          const T c = T(1) / sqrt(T(1) + tt * tt), sn = c * tt;
          T* vp = Vt + (size_t)p * n;
          T* vq = Vt + (size_t)q * n;
          for (int i = lane; i < n; i += 32) {
            const T x = ap[i], y = aq[i];
            ap[i] = c * x - sn * y;
            aq[i] = sn * x + c * y;
            const T vx = vp[i], vy = vq[i];
            vp[i] = c * vx - sn * vy;
            vq[i] = sn * vx + c * vy;
          }
          if (lane == 0) rotated = 1;
        }
      }
      __syncthreads();
    }
    const int again = rotated;
    __syncthreads();
    if (!again) break;
  }
  // singular values, descending order
  for (int j = warp; j < n; j += kWarps) {
    T acc = T(0);
    for (int i = lane; i < n; i += 32) { const T x = At[(size_t)j * n + i]; acc += x * x; }
    acc = warp_sum(acc);
    if (lane == 0) sig[j] = sqrt(acc);
  }
  __syncthreads();
  for (int j = tid; j < n; j += kLaThreads) {
    const T sj = sig[j];
    int rank = 0;
    for (int i = 0; i < n; ++i) {
      const T si = sig[i];
      if (si > sj || (si == sj && i < j)) ++rank;
    }
    S[rank] = sj;
    // stash the destination in the (now unused) diagonal-free slot: reuse sig as int via a second pass
    reinterpret_cast<int*>(sig + n)[j] = rank;
  }
  __syncthreads();
  const int* rank_of = reinterpret_cast<const int*>(sig + n);
  for (int e = tid; e < n * n; e += kLaThreads) {
    const int j = e / n, i = e - j * n;
    const int r = rank_of[j];
    const T sj = sig[j];
    U[(size_t)i * n + r] = sj > T(0) ? At[e] / sj : T(0);
    Vt_out[(size_t)r * n + i] = Vt[e];
  }
}

template <typename T, typename K>
int run_single_cta(K kernel, int64_t n, const T* A, int64_t lda, T* out, int64_t ldo, int32_t* info, void* ws,
                   size_t ws_bytes, cudaStream_t s, const char* what) {
  NUMS_REQUIRE(n >= 1 && n <= 1024, "%s: n = %lld outside the supported range [1, 1024]", what, (long long)n);
  const size_t smem = (size_t)n * (n | 1) * sizeof(T);
  const bool use_smem = smem <= 200 * 1024;
  if (!use_smem) NUMS_NEED_WS((size_t)n * n * sizeof(T), ws_bytes);
  if (use_smem) NUMS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<1, kLaThreads, use_smem ? smem : 0, s>>>(A, lda, (int)n, out, ldo, static_cast<T*>(ws), use_smem ? 1 : 0, info);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

template <typename T>
int run_svd(int64_t n, const T* A, int64_t lda, T* U, T* S, T* Vt, void* ws, size_t ws_bytes, cudaStream_t s) {
  NUMS_REQUIRE(n >= 1 && n <= 1024, "svd: n = %lld outside the supported range [1, 1024]", (long long)n);
  const size_t need = ((size_t)2 * n * n + 3 * n + 8) * sizeof(T);
  NUMS_NEED_WS(need, ws_bytes);
  T* At = static_cast<T*>(ws);
  T* Vw = At + (size_t)n * n;
  T* sig = Vw + (size_t)n * n;   // n singular values followed by n ints (rank table)
  svd_jacobi_kernel<T><<<1, kLaThreads, 0, s>>>(A, lda, (int)n, U, S, Vt, At, Vw, sig);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

}  // namespace
}  // namespace nums

extern "C" int nums_qr(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, void* Q, int64_t ldq,
                       void* R, int64_t ldr, void* ws, size_t ws_bytes, void* stream) {
  using namespace nums;
  (void)ldq;
  NUMS_REQUIRE(m >= 1 && n >= 1, "qr: empty matrix");
  NUMS_REQUIRE(A && R, "qr: null pointer");
  if (Q != nullptr)
    NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "qr: explicit Q is formed by the caller as A R^-1 (+ one refinement)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == NUMS_F64) return run_qr_r<double>(m, n, static_cast<const double*>(A), lda, static_cast<double*>(R), ldr, ws, ws_bytes, s);
  if (dtype == NUMS_F32) return run_qr_r<float>(m, n, static_cast<const float*>(A), lda, static_cast<float*>(R), ldr, ws, ws_bytes, s);
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "qr: dtype %s", dtype_name(dtype));
}

extern "C" int nums_inv(int dtype, int64_t n, const void* A, int64_t lda, void* Ainv, int64_t ldi,
                        int32_t* info, void* ws, size_t ws_bytes, void* stream) {
  using namespace nums;
  NUMS_REQUIRE(A && Ainv, "inv: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == NUMS_F64)
    return run_single_cta<double>(inv_kernel<double>, n, static_cast<const double*>(A), lda, static_cast<double*>(Ainv), ldi, info, ws, ws_bytes, s, "inv");
  if (dtype == NUMS_F32)
    return run_single_cta<float>(inv_kernel<float>, n, static_cast<const float*>(A), lda, static_cast<float*>(Ainv), ldi, info, ws, ws_bytes, s, "inv");
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "inv: dtype %s", dtype_name(dtype));
}

extern "C" int nums_cholesky(int dtype, int64_t n, const void* A, int64_t lda, void* L, int64_t ldl,
                             int32_t* info, void* ws, size_t ws_bytes, void* stream) {
  using namespace nums;
  NUMS_REQUIRE(A && L, "cholesky: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == NUMS_F64)
    return run_single_cta<double>(cholesky_kernel<double>, n, static_cast<const double*>(A), lda, static_cast<double*>(L), ldl, info, ws, ws_bytes, s, "cholesky");
  if (dtype == NUMS_F32)
    return run_single_cta<float>(cholesky_kernel<float>, n, static_cast<const float*>(A), lda, static_cast<float*>(L), ldl, info, ws, ws_bytes, s, "cholesky");
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "cholesky: dtype %s", dtype_name(dtype));
}

extern "C" int nums_gram_factor(int64_t n, const void* G, int64_t ldg, void* L, int64_t ldl, void* R, int64_t ldr,
                                void* Linv, int64_t ldi, double* stats, void* stream) {
  using namespace nums;
  NUMS_REQUIRE(n >= 1 && n <= kGfMaxN, "gram_factor: n = %lld outside the supported range [1, %d]", (long long)n, kGfMaxN);
  NUMS_REQUIRE(G && stats, "gram_factor: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)n * (n + 1) * sizeof(double);
  NUMS_CUDA_OK(cudaFuncSetAttribute(gram_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_factor_kernel<<<1, kGfThreads, smem, s>>>(static_cast<const double*>(G), ldg, (int)n, static_cast<double*>(L), ldl,
                                                 static_cast<double*>(R), ldr, static_cast<double*>(Linv), ldi, stats);
  NUMS_LAUNCH_OK();
  return NUMS_OK;
}

extern "C" int nums_svd(int dtype, int64_t n, const void* A, int64_t lda, void* U, void* S, void* Vt,
                        void* ws, size_t ws_bytes, void* stream) {
  using namespace nums;
  NUMS_REQUIRE(A && U && S && Vt, "svd: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == NUMS_F64)
    return run_svd<double>(n, static_cast<const double*>(A), lda, static_cast<double*>(U), static_cast<double*>(S),
                           static_cast<double*>(Vt), ws, ws_bytes, s);
  if (dtype == NUMS_F32)
    return run_svd<float>(n, static_cast<const float*>(A), lda, static_cast<float*>(U), static_cast<float*>(S),
                          static_cast<float*>(Vt), ws, ws_bytes, s);
  NUMS_FAIL(NUMS_ERR_UNSUPPORTED, "svd: dtype %s", dtype_name(dtype));
}
