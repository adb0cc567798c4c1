// svd_block.cuh -- building blocks of the truncating SVD of the op sweep (reference: the TruncBond/TruncThresh SVD inside
// TensorTrains.compress!, call site src/recursive_bp_factor.jl:127) for matrices too large for a one-CTA Jacobi:
// un-squared block subspace iteration
//     Z = orth(M^T Q),  Y = M Z,  Y = Q R  (Householder),  Ritz values = singular values of the b x b factor R
// * the two tall products are DMMA (mma.sync m8n8k4 f64) GEMMs whose A operand streams from L2 in either orientation
//   (no transposed copy of M) and whose B operand is the b-column block in shared memory;
// * the blocks are orthonormalised by an in-place Householder QR with explicit thin Q (unconditionally orthonormal, also
//   for the numerically rank-deficient blocks that graded spectra produce), one __syncthreads per column;
// * only the b x b triangular factor goes through the one-sided Jacobi (Ritz values every iteration, Ritz vectors at the
//   end), instead of the p x b block itself.
#pragma once
#include "common.cuh"
#include "jacobi.cuh"
#include "qr_ft.cuh"

namespace mpbp {

constexpr int SUB_BMAX = 64;   // storage bound of the block of the subspace iterations
constexpr int SUB_BLOCK = 48;  // block width used (>= 2d+8 at d = 20): convergence ~ (sigma_{b+1}/sigma_k)^2
constexpr int SUB_MAXIT = 60;

// leading dimension == 4 (mod 8): the DMMA B-fragment loads (lane (g,q4) reads B[q4 + ld*g]) then hit 16 distinct
// 8-byte bank pairs per half-warp (2 wavefronts for 256 bytes = the minimum)
__host__ __device__ inline int blk_ld(int rows) { return rows + ((12 - (rows & 7)) & 7); }

// C (m x b, column-major, ldc) = A (m x k) * B (k x b, column-major, ldb).
// A(i, kk) = A[i*sr + kk*sc] in global memory (either orientation of a column-major matrix); B in shared (or global)
// memory; b a multiple of 8, b <= 64.  Every thread of the CTA calls; the caller synchronises afterwards.
// Each warp owns pairs of 8-row tiles: one B fragment feeds two DMMAs; A fragments are prefetched one chunk ahead.
__device__ inline void blk_gemm_dmma(const double* __restrict__ A, const long long sr, const long long sc, const int m, const int k,
                                     const double* B, const int ldb, const int b, double* C, const int ldc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q4 = lane & 3;
  const int nt = b >> 3;
  constexpr int KU = 4;  // k-steps (of 4) per prefetched chunk
  const int npair = (m + 15) >> 4;
  for (int mp = warp; mp < npair; mp += NW) {
    const int i0 = mp * 16 + g, i1 = i0 + 8;
    const bool ok0 = i0 < m, ok1 = i1 < m;
    double acc0[8][2], acc1[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc0[t][0] = acc0[t][1] = acc1[t][0] = acc1[t][1] = 0.0;
    const double* a0 = A + (long long)i0 * sr + (long long)q4 * sc;
    const double* a1 = A + (long long)i1 * sr + (long long)q4 * sc;
    double f0[KU], f1[KU], n0[KU], n1[KU];
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int kk = 4 * u + q4;
      f0[u] = (ok0 && kk < k) ? a0[(long long)(4 * u) * sc] : 0.0;
      f1[u] = (ok1 && kk < k) ? a1[(long long)(4 * u) * sc] : 0.0;
    }
    for (int k0 = 0; k0 < k; k0 += 4 * KU) {
      const int kn = k0 + 4 * KU;
#pragma unroll
      for (int u = 0; u < KU; ++u) {
        const int kk = kn + 4 * u + q4;
        n0[u] = (ok0 && kk < k) ? a0[(long long)(kn + 4 * u) * sc] : 0.0;
        n1[u] = (ok1 && kk < k) ? a1[(long long)(kn + 4 * u) * sc] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < KU; ++u) {
        const int kb = k0 + 4 * u + q4;
        const bool kok = kb < k;
        const double* bp = B + kb + (size_t)ldb * g;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          if (t < nt) {
            const double bf = kok ? bp[(size_t)ldb * 8 * t] : 0.0;
            dmma884(acc0[t][0], acc0[t][1], f0[u], bf);
            dmma884(acc1[t][0], acc1[t][1], f1[u], bf);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < KU; ++u) {
        f0[u] = n0[u];
        f1[u] = n1[u];
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (t < nt) {
        const int c = 8 * t + 2 * q4;
        if (ok0) {
          C[i0 + (size_t)ldc * c] = acc0[t][0];
          C[i0 + (size_t)ldc * (c + 1)] = acc0[t][1];
        }
        if (ok1) {
          C[i1 + (size_t)ldc * c] = acc1[t][0];
          C[i1 + (size_t)ldc * (c + 1)] = acc1[t][1];
        }
      }
    }
  }
}

// 8 simultaneous warp-wide sums: v[j] per lane -> every lane holds all 8 totals (transpose-reduce: 9 + 8 shuffles
// instead of 40)
__device__ __forceinline__ void warp_reduce8(double (&v)[8]) {
  const int lane = threadIdx.x & 31;
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  double w4[4], w2[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double send = b4 ? v[i] : v[i + 4], keep = b4 ? v[i + 4] : v[i];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double send = b3 ? w4[i] : w4[i + 2], keep = b3 ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  double t;
  {
    const double send = b2 ? w2[0] : w2[1], keep = b2 ? w2[1] : w2[0];
    t = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  t += __shfl_xor_sync(0xffffffffu, t, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 1);
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = __shfl_sync(0xffffffffu, t, (((c >> 2) & 1) << 4) | (((c >> 1) & 1) << 3) | ((c & 1) << 2));
}

// one Householder step on the warp's own trailing columns c = c0, c0+NW, ... (< b), all at once: the 8 dot products
// share every load of the reflector column and their reductions overlap (the step is latency-bound otherwise).
// mode 0 (forward): w = tau (cc[k] + sc x.cc),  cc[k] -= w ;   mode 1 (backward): w = tau sc x.cc,  cc[k] = -w.
__device__ __forceinline__ void hh_apply_cols(double* W, const int rows, const int b, const int ld, const int k, const double tau,
                                              const double sc, const int mode) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* col = W + (size_t)k * ld;
  const int c0 = k + 1 + warp;
  if (c0 >= b) return;
  const int nc = min(8, (b - c0 + NW - 1) / NW);
  double* cp[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cp[j] = W + (size_t)(c0 + NW * min(j, nc - 1)) * ld;
  double acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0;
  for (int i = k + 1 + lane; i < rows; i += 32) {
    const double x = col[i];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nc) acc[j] += x * cp[j][i];
  }
  warp_reduce8(acc);
  double ws[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double w = mode == 0 ? tau * (cp[j][k] + sc * acc[j]) : tau * sc * acc[j];
    ws[j] = w * sc;
    acc[j] = w;
  }
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nc) cp[j][k] = mode == 0 ? cp[j][k] - acc[j] : -acc[j];
  }
  for (int i = k + 1 + lane; i < rows; i += 32) {
    const double x = col[i];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nc) cp[j][i] -= ws[j] * x;
  }
}

// In-place Householder QR with explicit thin Q of the rows x b block W (column-major, ld), rows >= b, b <= 64:
// on return the columns of W are orthonormal (to machine precision, whatever the conditioning of the input) and, if
// Rout != nullptr, Rout (b x b column-major, ld = b) holds the triangular factor (zeros below the diagonal).
// scal: shared scratch of 3*b doubles.  Every warp recomputes the column-k reflector (no cross-warp reduction) and
// updates its own trailing columns c = k+1+warp (mod NW); one __syncthreads per column, forward and backward.
// W may live in shared or global memory.  All threads call; synchronised on return.
__device__ inline void hh_orth(double* W, const int rows, const int b, const int ld, double* Rout, double* scal) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* s_tau = scal;
  double* s_sc = scal + b;
  double* s_beta = scal + 2 * b;
  __syncthreads();
  for (int k = 0; k < b; ++k) {
    const double* col = W + (size_t)k * ld;
    double s0 = 0.0, s1 = 0.0;
    int i = k + 1 + lane;
    for (; i + 32 < rows; i += 64) {
      const double x = col[i], y = col[i + 32];
      s0 += x * x;
      s1 += y * y;
    }
    if (i < rows) s0 += col[i] * col[i];
    const double sig2 = warp_sum(s0 + s1);
    const double alpha = col[k];
    double tau = 0.0, sc = 0.0, beta = alpha;
    if (sig2 > 0.0) {
      const double n2 = alpha * alpha + sig2;
      const double rs = rsqrt(n2);
      const double nrm = n2 * rs;
      beta = alpha >= 0.0 ? -nrm : nrm;
      const double u = alpha - beta;
      tau = alpha >= 0.0 ? u * rs : -u * rs;
      sc = 1.0 / u;
    }
    hh_apply_cols(W, rows, b, ld, k, tau, sc, 0);
    if (warp == (k & (NW - 1)) && lane == 0) {
      s_tau[k] = tau;
      s_sc[k] = sc;
      s_beta[k] = beta;
    }
    __syncthreads();
  }
  if (Rout) {
    for (int idx = threadIdx.x; idx < b * b; idx += NT) {
      const int j = idx % b, c = idx / b;
      Rout[idx] = j < c ? W[j + (size_t)c * ld] : (j == c ? s_beta[c] : 0.0);
    }
    __syncthreads();
  }
  // Q = H_0 ... H_{b-1} [I; 0], accumulated backwards in place (column k holds x with v = [1; sc*x] until its turn;
  // rows <= k of the columns c > k are zero at that point: their stale R entries are never read)
  for (int k = b - 1; k >= 0; --k) {
    const double tau = s_tau[k], sc = s_sc[k];
    hh_apply_cols(W, rows, b, ld, k, tau, sc, 1);
    __syncthreads();
    double* ck = W + (size_t)k * ld;
    const double ts = tau * sc;
    for (int i = threadIdx.x; i < rows; i += NT) ck[i] = i < k ? 0.0 : (i == k ? 1.0 - tau : -ts * ck[i]);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Cholesky-QR2 orthonormalisation of the rows x b block W (column-major, ld) in shared memory, with the Householder
// routine above as the unconditional fallback:
//   1. scale every column to unit norm (D);  2. G = W^T W (DMMA, both operands from the block);  3. G = R^T R (Cholesky,
//   one warp, b x b in shared memory);  4. W <- W R^-1 (one thread per row, forward substitution);  repeat 2-4 once.
// Q is orthonormal to machine precision when cond(W D^-1) < ~1e6: true for the blocks of a subspace iteration once it has
// started to converge (columns ~ sigma_j v_j: orthogonal up to their scale).  A pivot below 1e-13 (or a zero column) means
// the block is too ill-conditioned for this route: the caller falls back to hh_orth on the ORIGINAL block, which is why
// the input is preserved in `backup` (global scratch, rows x b with ld) until the first factorisation succeeded.
// On success returns true; Rout (b x b column-major, optional) = R2 R1 D, the triangular factor of the input block.
// G1, G2: shared scratch of b*b doubles each; dsc: b doubles.
// ---------------------------------------------------------------------------------------------------------------------
__device__ inline void blk_gram_dmma(const double* W, const int rows, const int b, const int ld, double* G) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q4 = lane & 3;
  const int nt = b >> 3;
  const int ntile = nt * (nt + 1) / 2;  // upper triangle of tiles (i <= j)
  for (int tix = warp; tix < ntile; tix += NW) {
    int ti = 0, rem = tix;
    while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
    const int tj = ti + rem;
    const double* ap = W + (size_t)(8 * ti + g) * ld + q4;  // A(i, k) = W[k + ld*i]
    const double* bp = W + (size_t)(8 * tj + g) * ld + q4;  // B(k, j) = W[k + ld*j]
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
    int k = 0;
    for (; k + 8 <= rows; k += 8) {
      dmma884(c0, c1, ap[k], bp[k]);
      dmma884(e0, e1, ap[k + 4], bp[k + 4]);
    }
    for (; k < rows; k += 4) {
      const bool ok = k + q4 < rows;
      dmma884(c0, c1, ok ? ap[k] : 0.0, ok ? bp[k] : 0.0);
    }
    c0 += e0;
    c1 += e1;
    const int i = 8 * ti + g, j = 8 * tj + 2 * q4;
    G[i + (size_t)b * j] = c0;
    G[i + (size_t)b * (j + 1)] = c1;
    G[j + (size_t)b * i] = c0;  // mirror (the Cholesky below reads the upper triangle only; kept symmetric for clarity)
    G[j + 1 + (size_t)b * i] = c1;
  }
}
// in-place upper Cholesky G = R^T R by the whole CTA (column-major b x b, upper triangle in/out), right-looking, two
// __syncthreads per column; *ok_flag = 0 on a pivot <= tol (all threads then return together).  All threads call.
__device__ inline void blk_chol_cta(double* G, const int b, const double tol, int* ok_flag) {
  for (int k = 0; k < b; ++k) {
    const double piv = G[k + (size_t)b * k];
    if (!(piv > tol)) {  // uniform: every thread reads the same value
      if (threadIdx.x == 0) *ok_flag = 0;
      __syncthreads();
      return;
    }
    const double inv = rsqrt(piv);
    __syncthreads();  // everyone has read the pivot
    for (int j = k + threadIdx.x; j < b; j += NT) G[k + (size_t)b * j] = j == k ? piv * inv : G[k + (size_t)b * j] * inv;  // row k of R
    __syncthreads();
    // trailing update G[i,j] -= R[k,i] R[k,j] for k < i <= j, one entry per thread
    const int nr = b - k - 1;
    for (int e = threadIdx.x; e < nr * nr; e += NT) {
      const int i = k + 1 + e % nr, j = k + 1 + e / nr;
      if (i <= j) G[i + (size_t)b * j] -= G[k + (size_t)b * i] * G[k + (size_t)b * j];
    }
    __syncthreads();
  }
}
// R <- R^-1 IN PLACE (upper triangular b x b, column-major, ld b, shared memory): rows from the bottom up; at row i thread j
// (j >= i) forms X[i,j] = -(sum_{k=i+1..j} R[i,k] X[k,j]) / R[i,i] from row i of R (still intact) and the finished rows k > i of
// its own column; the row is written after a barrier.  All threads call (b <= NT).
__device__ inline void blk_tri_inverse_inplace(double* R, const int b) {
  const int j = threadIdx.x;
  for (int i = b - 1; i >= 0; --i) {
    double v = 0.0;
    if (j < b && j >= i) {
      if (j == i) v = 1.0 / R[i + (size_t)b * i];
      else {
        double a0 = 0.0, a1 = 0.0;
        int k = i + 1;
        for (; k + 1 <= j; k += 2) {
          a0 += R[i + (size_t)b * k] * R[k + (size_t)b * j];
          a1 += R[i + (size_t)b * (k + 1)] * R[k + 1 + (size_t)b * j];
        }
        if (k <= j) a0 += R[i + (size_t)b * k] * R[k + (size_t)b * j];
        v = -(a0 + a1) / R[i + (size_t)b * i];
      }
    }
    __syncthreads();
    if (j < b && j >= i) R[i + (size_t)b * j] = v;
    __syncthreads();
  }
}
// W <- W X in place (X upper triangular b x b): each warp owns 8-row tiles, reads the whole tile row (b/4 A fragments) into
// registers before it writes; k-steps above the diagonal of X are skipped.  b <= 64, multiple of 8.
__device__ inline void blk_trmm_dmma(double* W, const int rows, const int b, const int ld, const double* X) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q4 = lane & 3;
  const int nt = b >> 3, nks = b >> 2;
  for (int mt = warp; mt < (rows + 7) / 8; mt += NW) {
    const int i = 8 * mt + g;
    const bool ok = i < rows;
    double af[SUB_BMAX / 4];
#pragma unroll
    for (int ks = 0; ks < SUB_BMAX / 4; ++ks) af[ks] = (ok && ks < nks) ? W[i + (size_t)ld * (4 * ks + q4)] : 0.0;
    __syncwarp();
#pragma unroll
    for (int t = 0; t < SUB_BMAX / 8; ++t) {
      if (t < nt) {
        double c0 = 0.0, c1 = 0.0;
        const double* xp = X + q4 + (size_t)b * (8 * t + g);  // B(k, n) = X[k + b*n]
#pragma unroll
        for (int ks = 0; ks < SUB_BMAX / 4; ++ks)
          if (ks <= 2 * t + 1) dmma884(c0, c1, af[ks], xp[4 * ks]);  // X[k, n] = 0 for k > n: k-steps beyond the tile's columns vanish
        if (ok) {
          W[i + (size_t)ld * (8 * t + 2 * q4)] = c0;
          W[i + (size_t)ld * (8 * t + 2 * q4 + 1)] = c1;
        }
      }
    }
  }
}
// G: shared b*b (Gram -> Cholesky factor -> its inverse, in place); Xg: b*b global temp (product of the two triangular
// factors); Racc: shared b*b (only touched when want_R); dsc: b doubles; backup: global, rows x b with the block's ld
__device__ inline bool cholqr2_orth(double* W, const int rows, const int b, const int ld, const bool want_R, double* Racc, double* G,
                                    double* Xg, double* dsc, double* backup, int* s_flag) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  // backup + column norms
  for (int j = warp; j < b; j += NW) {
    double s = 0.0;
    for (int i = lane; i < rows; i += 32) {
      const double x = W[i + (size_t)ld * j];
      backup[i + (size_t)ld * j] = x;
      s += x * x;
    }
    s = warp_sum(s);
    if (lane == 0) dsc[j] = sqrt(s);
  }
  if (threadIdx.x == 0) *s_flag = 1;
  __syncthreads();
  double dmax = 0.0;
  for (int j = 0; j < b; ++j) dmax = fmax(dmax, dsc[j]);
  bool good = dmax > 0.0;
  for (int j = 0; j < b; ++j) good = good && (dsc[j] > 1e-150 * dmax);
  if (!good) return false;  // (numerically) zero column: the Householder route handles it; the block is untouched so far
  for (int j = warp; j < b; j += NW) {
    const double f = 1.0 / dsc[j];
    for (int i = lane; i < rows; i += 32) W[i + (size_t)ld * j] *= f;
  }
  __syncthreads();
  for (int pass = 0; pass < 2; ++pass) {
    blk_gram_dmma(W, rows, b, ld, G);
    __syncthreads();
    blk_chol_cta(G, b, pass == 0 ? 1e-13 : 0.5, s_flag);
    __syncthreads();
    if (*s_flag == 0) {
      // too ill-conditioned for the Gram route: restore the input, the caller runs the Householder route
      for (int j = warp; j < b; j += NW)
        for (int i = lane; i < rows; i += 32) W[i + (size_t)ld * j] = backup[i + (size_t)ld * j];
      __syncthreads();
      return false;
    }
    if (want_R) {
      if (pass == 0) {
        for (int idx = threadIdx.x; idx < b * b; idx += NT) {
          const int i = idx % b, j = idx / b;
          Racc[idx] = i <= j ? G[idx] * dsc[j] : 0.0;  // R1 D
        }
      } else {
        for (int idx = threadIdx.x; idx < b * b; idx += NT) {
          const int i = idx % b, j = idx / b;
          double acc = 0.0;
          if (i <= j)
            for (int k = i; k <= j; ++k) acc += G[i + (size_t)b * k] * Racc[k + (size_t)b * j];
          Xg[idx] = acc;  // R2 (R1 D)
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < b * b; idx += NT) Racc[idx] = Xg[idx];
      }
      __syncthreads();
    }
    for (int idx = threadIdx.x; idx < b * b; idx += NT)
      if (idx % b > idx / b) G[idx] = 0.0;  // strictly lower part: the in-place inverse and the DMMA product read full tiles
    __syncthreads();
    blk_tri_inverse_inplace(G, b);
    blk_trmm_dmma(W, rows, b, ld, G);
    __syncthreads();
  }
  return true;
}

}  // namespace mpbp
