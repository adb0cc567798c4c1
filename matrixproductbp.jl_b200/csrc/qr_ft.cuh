// qr_ft.cuh -- Q-less QR of a tall row-major matrix by a flat-tree (row-block sequential) Householder sweep,
// one CTA per matrix, FP64 tensor-pipe (DMMA mma.sync.m8n8k4.f64) trailing updates.  The dominant kernel of
// the MPBP hot path (sweep 1 of TensorTrains.compress!, reference call site src/recursive_bp_factor.jl:127).
//
//   R (n x n, upper triangular) lives in global memory and stays L2-resident (n = 400: 1.3 MB);
//   the matrix is consumed in row blocks of H rows staged ONCE in shared memory (HBM traffic = one read of A):
//     for each row block A_i (H x n):   [R; A_i] = Q [R'; 0]
//       for each panel of 8 columns:
//         warp 0  : Householder factorisation of [R_jj (8x8 upper); A_i[:, panel] (H x 8)]  -> V (H x 8), T (8x8)
//         8 warps : for each 8-column slab c of the trailing columns, all warp-local:
//                     W  = R_jc + V^T A_ic        (DMMA, K = H)
//                     W' = T^T W                  (DMMA)
//                     R_jc -= W' ;  A_ic -= V W'  (DMMA, K = 8)
//   The per-column chain of the panel (one warp-wide reduction per column) is sequential; two CTAs per SM
//   (two independent matrices) overlap the chain of one with the DMMA updates of the other.
//
// Reflectors have the structure [e_k ; v] (the R part is a unit vector), so V^T V = I + V_A^T V_A and the
// compact-WY factor T follows the LAPACK dlarft recurrence.
#pragma once
#include "common.cuh"

namespace mpbp {

constexpr int FT_B = 8;

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__host__ __device__ inline int ft_ld(int n) {
  const int n8 = (n + 7) & ~7;
  return n8 + 4 + ((n8 & 8) ? 8 : 0);  // ld % 16 == 4 : conflict-free B-fragment loads
}
template <int H>
__host__ __device__ inline size_t ft_smem_doubles(int n) {
  return (size_t)H * ft_ld(n) + FT_B * (H + 4) + FT_B * FT_B + NW * 72 + 16;
}

// A: m x n row-major (lda), read-only.  R: n x n row-major (ldr) in global memory, fully overwritten; on return
// rows 0..min(m,n)-1 hold the R factor (rows >= m are rounding noise when m < n and must be ignored).
// If normalize: rows 0..min(m,n)-1 are divided by their max-abs.
template <int H>
__device__ void qr_ft_cta(const double* __restrict__ A, const int m, const int n, const int lda, double* __restrict__ R,
                          const int ldr, const bool normalize, double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n8 = (n + 7) & ~7;
  const int ld = ft_ld(n);
  constexpr int LDV = H + 4;
  double* Ablk = smem;                        // H x ld
  double* Vt = Ablk + (size_t)H * ld;         // 8 x LDV   (V^T, A-part of the reflectors)
  double* Tm = Vt + FT_B * LDV;               // 8 x 8
  double* Ws = Tm + FT_B * FT_B + warp * 72;  // per-warp 8x8 scratch (ld 9)
  // zero R
  for (int idx = tid; idx < n * n; idx += NT) R[(size_t)(idx / n) * ldr + (idx % n)] = 0.0;
  const int g = lane >> 2, q4 = lane & 3;  // DMMA fragment coordinates
  for (int row0 = 0; row0 < m; row0 += H) {
    __syncthreads();
    // ---- stage the row block (zero padded) ----
    for (int idx = tid; idx < H * n8; idx += NT) {
      const int i = idx / n8, c = idx % n8;
      const int gi = row0 + i;
      Ablk[(size_t)i * ld + c] = (gi < m && c < n) ? A[(size_t)gi * lda + c] : 0.0;
    }
    __syncthreads();
    for (int j0 = 0; j0 < n; j0 += FT_B) {
      // trailing-slab geometry; the first R fragment of every warp is prefetched across the panel phase
      const int nslab = (n8 - j0 - FT_B) / FT_B;
      const int rr = j0 + g;
      double nr0 = 0.0, nr1 = 0.0;
      {
        const int cc = j0 + FT_B + warp * FT_B + 2 * q4;
        if (warp < nslab && rr < n) {
          if (cc < n) nr0 = R[(size_t)rr * ldr + cc];
          if (cc + 1 < n) nr1 = R[(size_t)rr * ldr + cc + 1];
        }
      }
      // ================= panel factorisation (warp 0) =================
      if (warp == 0) {
        double a[FT_B];
        const bool rowok = lane < H;
#pragma unroll
        for (int c = 0; c < FT_B; ++c) a[c] = rowok ? Ablk[(size_t)lane * ld + j0 + c] : 0.0;
        // R_jj (8x8 upper) preloaded once: lane k holds row j0+k; rows are broadcast with shuffles
        double rrow[FT_B];
#pragma unroll
        for (int c = 0; c < FT_B; ++c)
          rrow[c] = (lane < FT_B && j0 + lane < n && j0 + c < n && c >= lane) ? R[(size_t)(j0 + lane) * ldr + j0 + c] : 0.0;
        double T[FT_B][FT_B];
#pragma unroll
        for (int x = 0; x < FT_B; ++x)
#pragma unroll
          for (int y = 0; y < FT_B; ++y) T[x][y] = 0.0;
#pragma unroll
        for (int k = 0; k < FT_B; ++k) {
          // one batched reduction: g[c] = a_k . a_c (c >= k)  and  z[l] = v_l . a_k (l < k)
          double red[FT_B];
#pragma unroll
          for (int c = 0; c < FT_B; ++c) red[c] = a[k] * a[c];
#pragma unroll
          for (int c = 0; c < FT_B; ++c) red[c] = warp_sum(red[c]);
          double rk[FT_B];
#pragma unroll
          for (int c = 0; c < FT_B; ++c) rk[c] = __shfl_sync(0xffffffffu, rrow[c], k);
          const double alpha = rk[k], sig2 = red[k];
          double tau = 0.0, sc = 0.0, beta = alpha;
          if (sig2 > 0.0) {
            const double nrm = sqrt(alpha * alpha + sig2);
            beta = alpha >= 0.0 ? -nrm : nrm;
            tau = (beta - alpha) / beta;
            sc = 1.0 / (alpha - beta);
          }
          const double v = a[k] * sc;
          a[k] = v;
#pragma unroll
          for (int c = k + 1; c < FT_B; ++c) {
            const double s = tau * (rk[c] + sc * red[c]);
            a[c] -= s * v;
            rk[c] -= s;
          }
          if (lane == k) {
            rrow[k] = beta;
#pragma unroll
            for (int c = k + 1; c < FT_B; ++c) rrow[c] = rk[c];
          }
          // T(0:k,k) = -tau * T(0:k,0:k) * (V_A(:,0:k)^T v_k);   z[l] = sc * (v_l . a_k) = sc * red[l]
          T[k][k] = tau;
#pragma unroll
          for (int x = 0; x < k; ++x) {
            double acc = 0.0;
#pragma unroll
            for (int l = x; l < k; ++l) acc += T[x][l] * (sc * red[l]);
            T[x][k] = -tau * acc;
          }
        }
        if (lane < FT_B && j0 + lane < n) {
#pragma unroll
          for (int c = 0; c < FT_B; ++c)
            if (c >= lane && j0 + c < n) R[(size_t)(j0 + lane) * ldr + j0 + c] = rrow[c];
        }
        if (rowok) {
#pragma unroll
          for (int k = 0; k < FT_B; ++k) Vt[k * LDV + lane] = a[k];
        }
        if (lane == 0) {
#pragma unroll
          for (int x = 0; x < FT_B; ++x)
#pragma unroll
            for (int y = 0; y < FT_B; ++y) Tm[x * FT_B + y] = T[x][y];
        }
        __threadfence_block();
      }
      __syncthreads();
      // ================= trailing update (all warps, one 8-column slab at a time) =================
      if (nslab > 0) {
        double va1[H / 4], va2[H / 8][2], at[2];
#pragma unroll
        for (int s = 0; s < H / 4; ++s) va1[s] = Vt[g * LDV + 4 * s + q4];
#pragma unroll
        for (int r = 0; r < H / 8; ++r)
#pragma unroll
          for (int s = 0; s < 2; ++s) va2[r][s] = Vt[(4 * s + q4) * LDV + 8 * r + g];
#pragma unroll
        for (int s = 0; s < 2; ++s) at[s] = Tm[(4 * s + q4) * FT_B + g];  // T^T[g][4s+q4]
        for (int sl = warp; sl < nslab; sl += NW) {
          const int c0 = j0 + FT_B + sl * FT_B;
          // R_jc fragment (C layout): rows j0+g, cols c0 + 2*q4 + {0,1}; the next slab's fragment is prefetched
          const int cc = c0 + 2 * q4;
          const bool ok0 = (rr < n) && (cc < n), ok1 = (rr < n) && (cc + 1 < n);
          double* rp = R + (size_t)rr * ldr + cc;
          const double r0 = nr0, r1 = nr1;
          {
            const int ccn = cc + NW * FT_B;
            nr0 = (sl + NW < nslab && rr < n && ccn < n) ? rp[NW * FT_B] : 0.0;
            nr1 = (sl + NW < nslab && rr < n && ccn + 1 < n) ? rp[NW * FT_B + 1] : 0.0;
          }
          double w0 = 0.0, w1 = 0.0;
          const double* bp = Ablk + (size_t)q4 * ld + c0 + g;
#pragma unroll
          for (int s = 0; s < H / 4; ++s) dmma884(w0, w1, va1[s], bp[(size_t)(4 * s) * ld]);
          w0 += r0;
          w1 += r1;
          __syncwarp();
          Ws[g * 9 + 2 * q4] = w0;
          Ws[g * 9 + 2 * q4 + 1] = w1;
          __syncwarp();
          double p0 = 0.0, p1 = 0.0;
#pragma unroll
          for (int s = 0; s < 2; ++s) dmma884(p0, p1, at[s], Ws[(4 * s + q4) * 9 + g]);
          if (ok0) rp[0] = r0 - p0;
          if (ok1) rp[1] = r1 - p1;
          __syncwarp();
          Ws[g * 9 + 2 * q4] = -p0;
          Ws[g * 9 + 2 * q4 + 1] = -p1;
          __syncwarp();
          const double b0 = Ws[q4 * 9 + g], b1 = Ws[(4 + q4) * 9 + g];
#pragma unroll
          for (int r = 0; r < H / 8; ++r) {
            double2* cp = reinterpret_cast<double2*>(Ablk + (size_t)(8 * r + g) * ld + cc);
            double2 cv = *cp;
            dmma884(cv.x, cv.y, va2[r][0], b0);
            dmma884(cv.x, cv.y, va2[r][1], b1);
            *cp = cv;
          }
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (normalize) {
    __shared__ double redn[NW + 1];
    const int k = min(m, n);
    double mx = 0.0;
    for (int idx = tid; idx < k * n; idx += NT) {
      const int i = idx / n, c = idx % n;
      if (c >= i) mx = fmax(mx, fabs(R[(size_t)i * ldr + c]));
    }
    mx = block_max(mx, redn);
    if (mx > 0.0 && isfinite(mx)) {
      const double f = 1.0 / mx;
      for (int idx = tid; idx < k * n; idx += NT) {
        const int i = idx / n, c = idx % n;
        if (c >= i) R[(size_t)i * ldr + c] *= f;
      }
    }
  }
  __syncthreads();
}

}  // namespace mpbp
