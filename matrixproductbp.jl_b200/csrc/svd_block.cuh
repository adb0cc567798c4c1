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

}  // namespace mpbp
