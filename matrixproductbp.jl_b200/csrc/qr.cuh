// qr.cuh -- Q-less Householder QR of a tall row-major matrix, one CTA per matrix (FP64).
//
// Role in the hot path (DESIGN.md "sweep 1"): TensorTrains.compress! first right-orthogonalises the
// bond-D Kronecker train with un-truncated SVDs (reference call site src/recursive_bp_factor.jl:127).
// Only the triangular factor of each site is needed to reproduce the truncations of the second sweep
// (gauge invariance), so the device computes R of A = QR and never forms Q.
//
// Algorithm: right-looking blocked Householder, panel width QB = 8.
//   panel   : thread-owns-rows, panel held in registers, one CTA-wide reduction per column
//             (Gram-row trick: the reflector's dot products with the remaining panel columns and its own
//             norm come out of the same reduction);
//   T       : compact-WY factor from V^T V (one more reduction);
//   trailing: thread-owns-column, two passes over the trailing matrix (W = V^T A ; A -= V T^T W) with V
//             staged in shared memory in row chunks.
#pragma once
#include "common.cuh"

namespace mpbp {

constexpr int QB = 8;
constexpr int QR_MAX_RPT = 8;                  // rows per thread in the panel -> m <= 2048
constexpr int QR_MAX_M = QR_MAX_RPT * NT;
constexpr int QR_MAX_CPT = 4;                  // trailing columns per thread  -> n <= 1024 + 8
constexpr int QR_NPAIR = QB * (QB - 1) / 2;    // 28

struct QRShared {
  double* V;   // vrows * QB
  int vrows;
  double* red;  // NW * QR_NPAIR
  double* g;    // 2 * QR_NPAIR   (double-buffered broadcast)
  double* row;  // 2 * QB
  double* tau;  // QB
  double* T;    // QB * QB
};
__host__ __device__ inline size_t qr_shared_doubles(int vrows) {
  return (size_t)vrows * QB + NW * QR_NPAIR + 2 * QR_NPAIR + 2 * QB + QB + QB * QB;
}
__device__ inline QRShared qr_carve(double* smem, int vrows) {
  QRShared s;
  s.V = smem;
  s.vrows = vrows;
  s.red = s.V + (size_t)vrows * QB;
  s.g = s.red + NW * QR_NPAIR;
  s.row = s.g + 2 * QR_NPAIR;
  s.tau = s.row + 2 * QB;
  s.T = s.tau + QB;
  return s;
}

template <int RPT>
__device__ void qr_panels(double* __restrict__ A, const int m, const int n, const int lda, const QRShared sm) {
  const int tid = threadIdx.x;
  const int kmax = min(m - 1, n);
  for (int j0 = 0; j0 < kmax; j0 += QB) {
    const int pw = min(QB, n - j0);
    const int bw = min(pw, kmax - j0);
    double P[RPT][QB];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int i = j0 + tid + NT * r;
#pragma unroll
      for (int c = 0; c < QB; ++c) P[r][c] = (i < m && c < pw) ? A[(size_t)i * lda + j0 + c] : 0.0;
    }
    // ---- panel factorisation (k loop fully unrolled so that P[][] stays in registers) ----
#pragma unroll
    for (int k = 0; k < QB; ++k) {
      if (k >= bw) break;
      double g[QB];
#pragma unroll
      for (int c = 0; c < QB; ++c) g[c] = 0.0;
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const bool below = (r > 0) || (tid > k);
        if (below) {
          const double vk = P[r][k];
#pragma unroll
          for (int c = k; c < QB; ++c) g[c] += vk * P[r][c];
        }
      }
      double* rowb = sm.row + (k & 1) * QB;
      double* gb = sm.g + (k & 1) * QR_NPAIR;
      if (tid == k) {
#pragma unroll
        for (int c = 0; c < QB; ++c) rowb[c] = P[0][c];
      }
      block_sum<QB>(g, sm.red, gb);
      const double alpha = rowb[k], sig2 = gb[k];
      double tau = 0.0, sc = 0.0, beta = alpha;
      if (sig2 > 0.0) {
        const double nrm = sqrt(alpha * alpha + sig2);
        beta = alpha >= 0.0 ? -nrm : nrm;
        tau = (beta - alpha) / beta;
        sc = 1.0 / (alpha - beta);
      }
      double s[QB];
#pragma unroll
      for (int c = 0; c < QB; ++c) s[c] = (c > k && c < pw) ? tau * (rowb[c] + sc * gb[c]) : 0.0;
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const bool below = (r > 0) || (tid > k);
        if (below) {
          const double v = P[r][k] * sc;
          P[r][k] = v;
#pragma unroll
          for (int c = k + 1; c < QB; ++c) P[r][c] -= s[c] * v;
        }
      }
      if (tid == k) {
        P[0][k] = beta;
#pragma unroll
        for (int c = k + 1; c < QB; ++c) P[0][c] -= s[c];
      }
      if (tid == 0) sm.tau[k] = tau;
    }
    // ---- V^T V (strict upper) for the compact-WY factor ----
    {
      double S[QR_NPAIR];
#pragma unroll
      for (int p = 0; p < QR_NPAIR; ++p) S[p] = 0.0;
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        double v[QB];
        const bool diagrow = (r == 0) && (tid < bw);
#pragma unroll
        for (int c = 0; c < QB; ++c) {
          double x = (c < bw) ? P[r][c] : 0.0;
          if (diagrow) x = (c < tid) ? x : (c == tid ? 1.0 : 0.0);
          v[c] = x;
        }
        int p = 0;
#pragma unroll
        for (int a = 0; a < QB; ++a)
#pragma unroll
          for (int b = a + 1; b < QB; ++b) S[p++] += v[a] * v[b];
      }
      block_sum<QR_NPAIR>(S, sm.red, sm.g);
      if (tid == 0) {
        // dlarft, forward / columnwise: T(0:k,k) = -tau_k * T(0:k,0:k) * V(:,0:k)^T v_k
        auto SS = [&](int a, int b) {  // a < b
          return sm.g[a * QB - a * (a + 1) / 2 + (b - a - 1)];
        };
        for (int k = 0; k < QB; ++k)
          for (int a = 0; a < QB; ++a) sm.T[a * QB + k] = 0.0;
        for (int k = 0; k < bw; ++k) {
          const double tk = sm.tau[k];
          sm.T[k * QB + k] = tk;
          for (int a = 0; a < k; ++a) {
            double acc = 0.0;
            for (int l = a; l < k; ++l) acc += sm.T[a * QB + l] * SS(l, k);
            sm.T[a * QB + k] = -tk * acc;
          }
        }
      }
    }
    // ---- write the panel back (R on/above the diagonal, V below) ----
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int i = j0 + tid + NT * r;
      if (i < m) {
#pragma unroll
        for (int c = 0; c < QB; ++c)
          if (c < pw) A[(size_t)i * lda + j0 + c] = P[r][c];
      }
    }
    __syncthreads();
    // ---- trailing update ----
    const int c0 = j0 + pw;
    if (c0 >= n) continue;
    const int nrows = m - j0;
    double w[QR_MAX_CPT][QB];
#pragma unroll
    for (int cg = 0; cg < QR_MAX_CPT; ++cg)
#pragma unroll
      for (int k = 0; k < QB; ++k) w[cg][k] = 0.0;
    const int nchunks = (nrows + sm.vrows - 1) / sm.vrows;
    for (int ch = 0; ch < nchunks; ++ch) {
      const int base = ch * sm.vrows;
      const int cr = min(sm.vrows, nrows - base);
      for (int idx = tid; idx < cr * QB; idx += NT) {
        const int ii = idx / QB, k = idx % QB;
        const int a = base + ii;  // row offset inside the panel rows
        double v = 0.0;
        if (k < bw) {
          if (a > k) v = A[(size_t)(j0 + a) * lda + j0 + k];
          else if (a == k) v = 1.0;
        }
        sm.V[idx] = v;
      }
      __syncthreads();
#pragma unroll
      for (int cg = 0; cg < QR_MAX_CPT; ++cg) {
        const int c = c0 + tid + NT * cg;
        if (c < n) {
          const double* Ap = A + (size_t)(j0 + base) * lda + c;
          for (int ii = 0; ii < cr; ++ii) {
            const double a = Ap[(size_t)ii * lda];
            const double2* vv = reinterpret_cast<const double2*>(sm.V + ii * QB);
            const double2 v0 = vv[0], v1 = vv[1], v2 = vv[2], v3 = vv[3];
            w[cg][0] += v0.x * a; w[cg][1] += v0.y * a; w[cg][2] += v1.x * a; w[cg][3] += v1.y * a;
            w[cg][4] += v2.x * a; w[cg][5] += v2.y * a; w[cg][6] += v3.x * a; w[cg][7] += v3.y * a;
          }
        }
      }
      if (nchunks > 1) __syncthreads();
    }
    // w <- T^T w
#pragma unroll
    for (int cg = 0; cg < QR_MAX_CPT; ++cg) {
      double t[QB];
#pragma unroll
      for (int k = 0; k < QB; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < QB; ++l)
          if (l <= k) acc += sm.T[l * QB + k] * w[cg][l];
        t[k] = acc;
      }
#pragma unroll
      for (int k = 0; k < QB; ++k) w[cg][k] = t[k];
    }
    for (int ch = 0; ch < nchunks; ++ch) {
      const int base = ch * sm.vrows;
      const int cr = min(sm.vrows, nrows - base);
      if (nchunks > 1) {
        for (int idx = tid; idx < cr * QB; idx += NT) {
          const int ii = idx / QB, k = idx % QB;
          const int a = base + ii;
          double v = 0.0;
          if (k < bw) {
            if (a > k) v = A[(size_t)(j0 + a) * lda + j0 + k];
            else if (a == k) v = 1.0;
          }
          sm.V[idx] = v;
        }
        __syncthreads();
      }
#pragma unroll
      for (int cg = 0; cg < QR_MAX_CPT; ++cg) {
        const int c = c0 + tid + NT * cg;
        if (c < n) {
          double* Ap = A + (size_t)(j0 + base) * lda + c;
          for (int ii = 0; ii < cr; ++ii) {
            const double2* vv = reinterpret_cast<const double2*>(sm.V + ii * QB);
            const double2 v0 = vv[0], v1 = vv[1], v2 = vv[2], v3 = vv[3];
            double acc = v0.x * w[cg][0];
            acc += v0.y * w[cg][1]; acc += v1.x * w[cg][2]; acc += v1.y * w[cg][3];
            acc += v2.x * w[cg][4]; acc += v2.y * w[cg][5]; acc += v3.x * w[cg][6]; acc += v3.y * w[cg][7];
            Ap[(size_t)ii * lda] -= acc;
          }
        }
      }
      __syncthreads();
    }
  }
}

// CTA-wide: factor A (m x n row-major, lda; destroyed) and write R (min(m,n) x n row-major, ldr) with
// explicit zeros below the diagonal.  If `normalize`, R is divided by its max-abs (the scale of the
// sweep-1 factors is irrelevant).  Requires m <= QR_MAX_M, n <= NT*QR_MAX_CPT + QB.
__device__ inline void qr_r_cta(double* A, int m, int n, int lda, double* R, int ldr, bool normalize,
                                double* smem, int vrows) {
  QRShared sm = qr_carve(smem, vrows);
  if (m <= NT) qr_panels<1>(A, m, n, lda, sm);
  else if (m <= 2 * NT) qr_panels<2>(A, m, n, lda, sm);
  else if (m <= 4 * NT) qr_panels<4>(A, m, n, lda, sm);
  else qr_panels<8>(A, m, n, lda, sm);
  __syncthreads();
  const int k = min(m, n);
  double mx = 0.0;
  if (normalize) {
    for (int idx = threadIdx.x; idx < k * n; idx += NT) {
      const int i = idx / n, c = idx % n;
      if (c >= i) mx = fmax(mx, fabs(A[(size_t)i * lda + c]));
    }
    mx = block_max(mx, sm.red);
  }
  const double f = (normalize && mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
  for (int idx = threadIdx.x; idx < k * n; idx += NT) {
    const int i = idx / n, c = idx % n;
    R[(size_t)i * ldr + c] = (c >= i) ? A[(size_t)i * lda + c] * f : 0.0;
  }
  __syncthreads();
}

}  // namespace mpbp
