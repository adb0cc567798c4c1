// jacobi.cuh -- one-sided (Hestenes) Jacobi SVD on the columns of a small column-major matrix, one CTA.
//
// Role in the hot path: every truncating SVD of the reference (TensorTrains truncators inside compress!,
// call sites src/recursive_bp_factor.jl:127,156) only needs the leading singular vectors on ONE side.
// After convergence the columns of A are mutually orthogonal: their norms are the singular values and the
// normalised columns the left singular vectors of the input.  High relative accuracy, warp-shuffle
// reductions, round-robin (tournament) pair ordering so that c/2 rotations run concurrently.
#pragma once
#include <cstdio>
#include "common.cuh"

namespace mpbp {

constexpr int JACOBI_MAX_SWEEPS = 60;
constexpr double JACOBI_EPS = 1.1102230246251565e-16;  // tolerance = 2*sqrt(p)*eps (LAPACK dgesvj: sqrt(m)*eps)

// A: p x c column-major (lda), in shared or global memory.  flag: one int in shared memory.
// Returns the number of sweeps used (JACOBI_MAX_SWEEPS+1 if not converged).  All threads must call.
// Columns whose norm is below JACOBI_ZERO x (largest column norm) are numerically zero: they are not rotated
// (a column inside the span of the others would otherwise shrink by eps per sweep forever) and callers zero
// them in U (jacobi_inv_sigma).  LAPACK resolves such directions only to eps*sigma_max as well.
constexpr double JACOBI_ZERO = 1e-15;
__device__ __forceinline__ double jacobi_inv_sigma(double s, double smax) {
  return (s > 2.0 * JACOBI_ZERO * smax && s > 0.0) ? 1.0 / s : 0.0;
}
// G lanes cooperate on one column pair (G = 8, 16 or 32): 32/G pairs run concurrently per warp, which is what the
// latency-bound pair step needs (a b = 64 block has 32 pairs per round = one pass over 8 warps at G = 8).
template <int G>
__device__ inline int jacobi_cols_g(double* A, const int p, const int c, const int lda, int* flag, const double thr2,
                                    const int max_sweeps) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int GPW = 32 / G;                 // groups per warp
  const int grp = w * GPW + lane / G, lg = lane % G;
  const int ngrp = NW * GPW;
  const int ce = c + (c & 1);
  const double tol = 2.0 * JACOBI_EPS * sqrt((double)max(p, 64));
  const double tol2 = tol * tol;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    __syncthreads();
    if (threadIdx.x == 0) *flag = 0;
    __syncthreads();
    for (int round = 0; round < ce - 1; ++round) {
      for (int pr0 = 0; pr0 < ce / 2; pr0 += ngrp) {
        const int pr = pr0 + grp;
        int i = 0, j = 0;
        bool act = pr < ce / 2;
        if (act) {
          if (pr == 0) {
            i = ce - 1;
            j = round;
          } else {
            i = (round + pr) % (ce - 1);
            j = (round - pr + (ce - 1)) % (ce - 1);
          }
          act = (i < c) && (j < c);
          if (i > j) {
            const int t = i;
            i = j;
            j = t;
          }
        }
        double* ai = A + (size_t)i * lda;
        double* aj = A + (size_t)j * lda;
        double a = 0.0, b = 0.0, g = 0.0;
        if (act) {
          // two independent accumulator sets break the loop-carried DFMA chains
          double a2 = 0.0, b2 = 0.0, g2 = 0.0;
          int k = lg;
#pragma unroll 2
          for (; k + G < p; k += 2 * G) {
            const double x = ai[k], y = aj[k], x2 = ai[k + G], y2 = aj[k + G];
            a += x * x;
            b += y * y;
            g += x * y;
            a2 += x2 * x2;
            b2 += y2 * y2;
            g2 += x2 * y2;
          }
          if (k < p) {
            const double x = ai[k], y = aj[k];
            a += x * x;
            b += y * y;
            g += x * y;
          }
          a += a2;
          b += b2;
          g += g2;
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
          g += __shfl_xor_sync(0xffffffffu, g, o);
        }
        // rotate when |g| > tol*sqrt(a*b) (sqrt-free test) and neither column is numerically zero
        if (act && g * g > tol2 * a * b && a > thr2 && b > thr2 && a * b > 0.0) {
          // t = sign(zeta)/(|zeta| + sqrt(1+zeta^2)), zeta = (b-a)/(2g)  ==  2g*sign(d)/(|d| + sqrt(d^2 + 4g^2)), d = b-a
          const double d = b - a;
          const double t = (d >= 0.0 ? 2.0 * g : -2.0 * g) / (fabs(d) + sqrt(d * d + 4.0 * g * g));
          const double cs = rsqrt(1.0 + t * t), sn = cs * t;
#pragma unroll 4
          for (int k = lg; k < p; k += G) {
            const double x = ai[k], y = aj[k];
            ai[k] = cs * x - sn * y;
            aj[k] = sn * x + cs * y;
          }
          if (lg == 0) *flag = 1;
        }
      }
      __syncthreads();
    }
    if (*flag == 0) break;
  }
  __syncthreads();
  return sweep;
}

// max_sweeps < JACOBI_MAX_SWEEPS: approximate orthogonalisation (enough inside a subspace iteration)
__device__ inline int jacobi_cols(double* A, const int p, const int c, const int lda, int* flag,
                                  const int max_sweeps = JACOBI_MAX_SWEEPS) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (c < 2) return 0;
  __shared__ double s_scale[NW];
  {
    double mx = 0.0;
    for (int j = w; j < c; j += NW) {
      double s = 0.0;
      for (int k = lane; k < p; k += 32) { const double x = A[k + (size_t)j * lda]; s += x * x; }
      mx = fmax(mx, warp_sum(s));
    }
    __syncthreads();
    if (lane == 0) s_scale[w] = mx;
    __syncthreads();
  }
  double scale2 = 0.0;
#pragma unroll
  for (int ww = 0; ww < NW; ++ww) scale2 = fmax(scale2, s_scale[ww]);
  const double thr2 = JACOBI_ZERO * JACOBI_ZERO * scale2;
  const int npair = (c + 1) / 2;
  int sweep;
  if (npair > 2 * NW && p >= 16) sweep = jacobi_cols_g<8>(A, p, c, lda, flag, thr2, max_sweeps);
  else if (npair > NW && p >= 32) sweep = jacobi_cols_g<16>(A, p, c, lda, flag, thr2, max_sweeps);
  else sweep = jacobi_cols_g<32>(A, p, c, lda, flag, thr2, max_sweeps);
  if (sweep >= JACOBI_MAX_SWEEPS && max_sweeps >= JACOBI_MAX_SWEEPS && threadIdx.x == 0) {
    printf("[mpbp] jacobi not converged: p=%d c=%d\n", p, c);
  }
  return sweep;
}

// After jacobi_cols: column norms -> sig[c]; order[r] = index of the r-th largest column (ties by index).
// sig, order in shared memory (c entries each).
__device__ inline void jacobi_sort(const double* A, const int p, const int c, const int lda, double* sig,
                                   int* order) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int j = w; j < c; j += NW) {
    const double* aj = A + (size_t)j * lda;
    double s = 0.0;
    for (int k = lane; k < p; k += 32) s += aj[k] * aj[k];
    s = warp_sum(s);
    if (lane == 0) sig[j] = sqrt(s);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < c; j += NT) {
    const double sj = sig[j];
    int rank = 0;
    for (int i = 0; i < c; ++i) {
      const double si = sig[i];
      rank += (si > sj) || (si == sj && i < j);
    }
    order[rank] = j;
  }
  __syncthreads();
}

// Deflation + pre-sorting: column norms of Ag (p x c, column-major, lda = p), columns at or below
// JACOBI_ZERO x (largest norm) are dropped (they carry < 1e-15 of the matrix), the others are copied to `dst`
// (p x c_eff, lda = p) in order of decreasing norm (a norm-sorted start converges in fewer sweeps).
// sig / order: scratch of c entries.  Returns c_eff (>= 1 unless the matrix is exactly zero).
__device__ inline int jacobi_compact(const double* Ag, const int p, const int c, double* dst, double* sig, int* order,
                                     int* s_int) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int j = w; j < c; j += NW) {
    const double* aj = Ag + (size_t)j * p;
    double s = 0.0;
    for (int k = lane; k < p; k += 32) s += aj[k] * aj[k];
    s = warp_sum(s);
    if (lane == 0) sig[j] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) *s_int = 0;
  __syncthreads();
  double mx = 0.0;
  for (int j = 0; j < c; ++j) mx = fmax(mx, sig[j]);
  const double thr2 = JACOBI_ZERO * JACOBI_ZERO * mx;
  for (int j = threadIdx.x; j < c; j += NT) {
    const double sj = sig[j];
    if (sj > thr2) {
      int rank = 0;
      for (int i = 0; i < c; ++i) {
        const double si = sig[i];
        rank += (si > sj) || (si == sj && i < j);
      }
      order[rank] = j;  // every kept column outranks every dropped one
      atomicAdd(s_int, 1);
    }
  }
  __syncthreads();
  const int ceff = *s_int;
  for (int idx = threadIdx.x; idx < ceff * p; idx += NT) {
    const int kk = idx / p, k = idx % p;
    dst[idx] = Ag[k + (size_t)order[kk] * p];
  }
  __syncthreads();
  return ceff;
}

}  // namespace mpbp
