// qr_ft.cuh -- Q-less QR of a tall row-major matrix by a flat-tree (row-block sequential) Householder sweep,
// one CTA per matrix, FP64 tensor-pipe (DMMA mma.sync.m8n8k4.f64) trailing updates.  The dominant kernel of
// the MPBP hot path (sweep 1 of TensorTrains.compress!, reference call site src/recursive_bp_factor.jl:127).
//
//   R (n x n, upper triangular) lives in global memory and stays L2-resident (n = 400: 1.3 MB);
//   the matrix is consumed in row blocks of H rows staged ONCE in shared memory (HBM traffic = one read of A):
//     for each row block A_i (H x n):   [R; A_i] = Q [R'; 0]
//       for each panel j of 8 columns:
//         warp 0    : (look-ahead) applies the update of panel j to the 8 columns of panel j+1, then factors
//                     [R_jj (8x8 upper); A_i[:, panel j+1] (H x 8)]  -> V (H x 8), T (8x8)   (double-buffered)
//         warps 1-7 : update of panel j on the remaining 8-column slabs c, all warp-local:
//                       W  = R_jc + V^T A_ic        (DMMA, K = H)
//                       W' = T^T W                  (DMMA)
//                       R_jc -= W' ;  A_ic -= V W'  (DMMA, K = 8)
//       one __syncthreads per panel.
//   The per-column chain of the panel (one warp-wide batched reduction per column) therefore runs concurrently
//   with the DMMA updates; two CTAs per SM (two independent matrices) fill the remaining bubbles.
//
// Reflectors have the structure [e_k ; v] (the R part is a unit vector), so V^T V = I + V_A^T V_A and the
// compact-WY factor T follows the LAPACK dlarft recurrence.
#pragma once
#include "common.cuh"
#include "bulk_copy.cuh"

namespace mpbp {

constexpr int FT_B = 8;

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__host__ __device__ inline int ft_ld(int n) {
  const int n8 = (n + 7) & ~7;
  return n8 + 8 + ((n8 & 8) ? 8 : 0);  // ld % 16 == 8; with the XOR-4 column swizzle of rows (i>>1)&1 both the
                                        // DMMA B-fragment loads and the 128-bit C-tile accesses are conflict-free
}
__device__ __forceinline__ int ft_sw(int row) { return ((row >> 1) & 1) << 2; }
template <int H>
__host__ __device__ inline size_t ft_smem_doubles(int n) {
  return (size_t)H * ft_ld(n) + 2 * (FT_B * (H + 4) + FT_B * FT_B) + NW * 72 + 16 + (H == 64 ? 2 * 128 : 0) + 288;
}

// H = 64: the otherwise idle partner warp of the panel warp stages, one panel ahead, the two 8x8 blocks of R the panel
// warp needs first in its serial chain -- R_{j,j+1} (look-ahead slab update) and R_{j+1,j+1} (panel factorisation) --
// from global/L2 into shared memory, so that their load latency leaves the critical path.
// dst[0..63] = R[8 jpt + r, 8 (jpt+1) + c], dst[64..127] = R[8 (jpt+1) + r, 8 (jpt+1) + c], zero outside the matrix.
__device__ __forceinline__ void ft_prefetch_R(const double* __restrict__ R, const int ldr, const int n, const int jpt, double* dst) {
  const int lane = threadIdx.x & 31;
  const int j0 = jpt * FT_B, c0 = j0 + FT_B;
  // four independent loads in flight (clamped address + select: no branch between them), then the stores
  double v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = lane + 32 * i;
    const int r = (e >> 3) & 7, c = e & 7;
    const int row = ((e >> 6) ? c0 : j0) + r, col = c0 + c;
    const bool ok = (row < n && col < n);
    const double x = R[ok ? (size_t)row * ldr + col : (size_t)0];
    v[i] = ok ? x : 0.0;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) dst[lane + 32 * i] = v[i];
}

// Householder factorisation of [R_jj ; Ablk[:, j0..j0+8)] by ONE warp.  Writes V^T (A part) to Vt (8 x LDV), the
// compact-WY T (8x8) to Tm, the new R_jj block to global R.
template <int H>
__device__ __forceinline__ void ft_panel(const double* Ablk, const int ld, const int j0, const int n, double* __restrict__ R,
                                         const int ldr, double* Vt, double* Tm, const double* rjj_s = nullptr) {
  constexpr int LDV = H + 4;
  constexpr int RPL = (H + 31) / 32;  // rows of the block per lane
  const int lane = threadIdx.x & 31;
  double a[RPL][FT_B];
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    const int row = lane + 32 * r;
#pragma unroll
    for (int c = 0; c < FT_B; ++c) a[r][c] = (row < H) ? Ablk[(size_t)row * ld + ((j0 + c) ^ ft_sw(row))] : 0.0;
  }
  // R_jj (8x8 upper) preloaded once: lane k holds row j0+k; rows are broadcast with shuffles
  double rrow[FT_B];
#pragma unroll
  for (int c = 0; c < FT_B; ++c) {
    if (rjj_s) rrow[c] = (lane < FT_B && c >= lane) ? rjj_s[lane * FT_B + c] : 0.0;  // staged by ft_prefetch_R (zero padded)
    else rrow[c] = (lane < FT_B && j0 + lane < n && j0 + c < n && c >= lane) ? R[(size_t)(j0 + lane) * ldr + j0 + c] : 0.0;
  }
  double T[FT_B][FT_B];
#pragma unroll
  for (int x = 0; x < FT_B; ++x)
#pragma unroll
    for (int y = 0; y < FT_B; ++y) T[x][y] = 0.0;
#pragma unroll
  for (int k = 0; k < FT_B; ++k) {
    // one batched reduction: red[c] = a_k . a_c (c >= k)  and  red[l] = v_l . a_k (l < k)
    double red[FT_B];
    {
      // transpose-reduce: 4+2+1+1+1 shuffles leave every lane with one complete sum, 8 more broadcast them
      double w4[4], w2[2];
      const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double lo = 0.0, hi = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
          lo += a[r][k] * a[r][i];
          hi += a[r][k] * a[r][i + 4];
        }
        const double send = b4 ? lo : hi, keep = b4 ? hi : lo;
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double send = b3 ? w4[i] : w4[i + 2], keep = b3 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      double tsum;
      {
        const double send = b2 ? w2[0] : w2[1], keep = b2 ? w2[1] : w2[0];
        tsum = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      tsum += __shfl_xor_sync(0xffffffffu, tsum, 2);
      tsum += __shfl_xor_sync(0xffffffffu, tsum, 1);
#pragma unroll
      for (int c = 0; c < FT_B; ++c)
        red[c] = __shfl_sync(0xffffffffu, tsum, (((c >> 2) & 1) << 4) | (((c >> 1) & 1) << 3) | ((c & 1) << 2));
    }
    double rk[FT_B];
#pragma unroll
    for (int c = 0; c < FT_B; ++c) rk[c] = (c >= k) ? __shfl_sync(0xffffffffu, rrow[c], k) : 0.0;
    const double alpha = rk[k], sig2 = red[k];
    double tau = 0.0, sc = 0.0, beta = alpha;
    if (sig2 > 0.0) {
      const double n2 = alpha * alpha + sig2;
      const double rs = rsqrt(n2);          // 1/||x||
      const double nrm = n2 * rs;
      beta = alpha >= 0.0 ? -nrm : nrm;
      const double u = alpha - beta;         // |u| = |alpha| + nrm, no cancellation
      tau = alpha >= 0.0 ? u * rs : -u * rs; // (beta - alpha)/beta = -u/beta
      sc = 1.0 / u;
    }
    double v[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
      v[r] = a[r][k] * sc;
      a[r][k] = v[r];
    }
#pragma unroll
    for (int c = k + 1; c < FT_B; ++c) {
      const double s = tau * (rk[c] + sc * red[c]);
#pragma unroll
      for (int r = 0; r < RPL; ++r) a[r][c] -= s * v[r];
      rk[c] -= s;
    }
    if (lane == k) {
      rrow[k] = beta;
#pragma unroll
      for (int c = k + 1; c < FT_B; ++c) rrow[c] = rk[c];
    }
    // T(0:k,k) = -tau * T(0:k,0:k) * (V_A(:,0:k)^T v_k);   V_l^T v_k = sc * red[l]
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
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    const int row = lane + 32 * r;
    if (row < H) {
#pragma unroll
      for (int k = 0; k < FT_B; ++k) Vt[k * LDV + row] = a[r][k];
    }
  }
  // T is replicated in every lane: lane x writes row x
#pragma unroll
  for (int x = 0; x < FT_B; ++x)
    if (lane == x) {
#pragma unroll
      for (int y = 0; y < FT_B; ++y) Tm[x * FT_B + y] = T[x][y];
    }
}

// ft_panel with fewer instructions on the (issue-bound) panel warp: the eight reduced dots go through a 16-double warp
// scratch (one store + four broadcast loads instead of 16 shuffles), and row k of R_jj is read from the staged copy when
// there is one (broadcast loads instead of shuffles).  scr: >= 16 doubles of warp-private shared memory.
template <int H>
__device__ __forceinline__ void ft_panel_l(const double* Ablk, const int ld, const int j0, const int n, double* __restrict__ R,
                                         const int ldr, double* Vt, double* Tm, const double* __restrict__ rjj_s, double* __restrict__ scr) {
  constexpr int LDV = H + 4;
  constexpr int RPL = (H + 31) / 32;  // rows of the block per lane
  const int lane = threadIdx.x & 31;
  double a[RPL][FT_B];
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    const int row = lane + 32 * r;
#pragma unroll
    for (int c = 0; c < FT_B; ++c) a[r][c] = (row < H) ? Ablk[(size_t)row * ld + ((j0 + c) ^ ft_sw(row))] : 0.0;
  }
  // R_jj (8x8 upper) preloaded once: lane k holds row j0+k; rows are broadcast with shuffles
  double rrow[FT_B];
#pragma unroll
  for (int c = 0; c < FT_B; ++c) {
    if (rjj_s) rrow[c] = (lane < FT_B && c >= lane) ? rjj_s[lane * FT_B + c] : 0.0;  // staged by ft_prefetch_R (zero padded)
    else rrow[c] = (lane < FT_B && j0 + lane < n && j0 + c < n && c >= lane) ? R[(size_t)(j0 + lane) * ldr + j0 + c] : 0.0;
  }
  // compact-WY T, ONE ROW PER LANE (lane x < 8 holds row x; T is upper triangular, so the zeros left of the diagonal make
  // the recurrence index-free): 2k+3 instead of k(k+3)/2 + k FP64 instructions in column step k
  double Trow[FT_B];
#pragma unroll
  for (int y = 0; y < FT_B; ++y) Trow[y] = 0.0;
#pragma unroll
  for (int k = 0; k < FT_B; ++k) {
    // one batched reduction: red[c] = a_k . a_c (c >= k)  and  red[l] = v_l . a_k (l < k)
    double red[FT_B];
    {
      // transpose-reduce: 4+2+1+1+1 shuffles leave every lane with one complete sum, 8 more broadcast them
      double w4[4], w2[2];
      const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double lo = 0.0, hi = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
          lo += a[r][k] * a[r][i];
          hi += a[r][k] * a[r][i + 4];
        }
        const double send = b4 ? lo : hi, keep = b4 ? hi : lo;
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double send = b3 ? w4[i] : w4[i + 2], keep = b3 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      double tsum;
      {
        const double send = b2 ? w2[0] : w2[1], keep = b2 ? w2[1] : w2[0];
        tsum = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      tsum += __shfl_xor_sync(0xffffffffu, tsum, 2);
      tsum += __shfl_xor_sync(0xffffffffu, tsum, 1);
      // lane bits 4,3,2 = column bits 2,1,0: every lane stores its sum, four broadcast loads hand all eight to every lane
      double* rb = scr + ((k & 1) << 3);
      rb[(((lane >> 4) & 1) << 2) | (((lane >> 3) & 1) << 1) | ((lane >> 2) & 1)] = tsum;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < FT_B; c += 2) {
        const double2 t2 = *reinterpret_cast<const double2*>(rb + c);
        red[c] = t2.x;
        red[c + 1] = t2.y;
      }
    }
    double rk[FT_B];
#pragma unroll
    for (int c = 0; c < FT_B; ++c) rk[c] = (c >= k) ? (rjj_s ? rjj_s[k * FT_B + c] : __shfl_sync(0xffffffffu, rrow[c], k)) : 0.0;
    const double alpha = rk[k], sig2 = red[k];
    double tau = 0.0, sc = 0.0, beta = alpha;
    if (sig2 > 0.0) {
      const double n2 = alpha * alpha + sig2;
      const double rs = rsqrt(n2);          // 1/||x||
      const double nrm = n2 * rs;
      beta = alpha >= 0.0 ? -nrm : nrm;
      const double u = alpha - beta;         // |u| = |alpha| + nrm, no cancellation
      tau = alpha >= 0.0 ? u * rs : -u * rs; // (beta - alpha)/beta = -u/beta
      sc = 1.0 / u;
    }
    double v[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
      v[r] = a[r][k] * sc;
      a[r][k] = v[r];
    }
#pragma unroll
    for (int c = k + 1; c < FT_B; ++c) {
      const double s = tau * (rk[c] + sc * red[c]);
#pragma unroll
      for (int r = 0; r < RPL; ++r) a[r][c] -= s * v[r];
      rk[c] -= s;
    }
    if (lane == k) {
      rrow[k] = beta;
#pragma unroll
      for (int c = k + 1; c < FT_B; ++c) rrow[c] = rk[c];
    }
    // T(0:k,k) = -tau * T(0:k,0:k) * (V_A(:,0:k)^T v_k);   V_l^T v_k = sc * red[l]
    {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < k; ++l) acc += Trow[l] * (sc * red[l]);
      Trow[k] = (lane == k) ? tau : ((lane < k) ? -tau * acc : 0.0);
    }
  }
  if (lane < FT_B && j0 + lane < n) {
#pragma unroll
    for (int c = 0; c < FT_B; ++c)
      if (c >= lane && j0 + c < n) R[(size_t)(j0 + lane) * ldr + j0 + c] = rrow[c];
  }
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    const int row = lane + 32 * r;
    if (row < H) {
#pragma unroll
      for (int k = 0; k < FT_B; ++k) Vt[k * LDV + row] = a[r][k];
    }
  }
  if (lane < FT_B) {
#pragma unroll
    for (int y = 0; y < FT_B; ++y) Tm[lane * FT_B + y] = Trow[y];
  }
}

// Branch-free reciprocal square root / reciprocal of a NORMAL double (the caller excludes zero, subnormal and non-finite
// arguments): the hardware seed (20 mantissa bits) followed by the same Newton steps libdevice uses (one cubic step for
// rsqrt, two quadratic steps for rcp; result within 1-2 ulp), without libdevice's special-case branches -- on the serial
// panel chain the reconvergence bookkeeping of those branches costs ~60 cycles per column step.
__device__ __forceinline__ double ft_rsqrt_fast(const double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y * y, 1.0);
  const double t = fma(e, 0.375, 0.5);
  return fma(t, y * e, y);
}
__device__ __forceinline__ double ft_rcp_fast(const double u) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(u));
  double e = fma(-u, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y);
  e = fma(-u, y, 1.0);
  return fma(y, e, y);
}

// ft_panel_l with the cross-lane reduction of the eight dots done through shared memory: every lane stores its eight
// partial dots (column-major 8 x 36 scratch, conflict-free), lane (c, qt) = (lane >> 2, lane & 3) adds the eight partials
// i*4+qt of column c, two shuffle levels finish the sum -- 36 instead of 61 instructions per column step on the
// issue-bound panel warp.  STAGED: row k of R_jj is read from / written back to the 64-double staged block rjj (shared
// memory) and the block is copied to global memory once at the end.  pred: 288 doubles, scr: 16 doubles (warp-private).
template <int H, bool STAGED>
__device__ __forceinline__ void ft_panel_m(const double* Ablk, const int ld, const int j0, const int n, double* __restrict__ R,
                                           const int ldr, double* Vt, double* Tm, double* rjj, double* scr, double* pred) {
  constexpr int LDV = H + 4;
  constexpr int RPL = (H + 31) / 32;
  constexpr int LDP = 36;
  const int lane = threadIdx.x & 31;
  double a[RPL][FT_B];
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    const int row = lane + 32 * r;
#pragma unroll
    for (int c = 0; c < FT_B; ++c) a[r][c] = (row < H) ? Ablk[(size_t)row * ld + ((j0 + c) ^ ft_sw(row))] : 0.0;
  }
  double rrow[FT_B];
  if (!STAGED) {
#pragma unroll
    for (int c = 0; c < FT_B; ++c)
      rrow[c] = (lane < FT_B && j0 + lane < n && j0 + c < n && c >= lane) ? R[(size_t)(j0 + lane) * ldr + j0 + c] : 0.0;
  }
  double Trow[FT_B];
#pragma unroll
  for (int y = 0; y < FT_B; ++y) Trow[y] = 0.0;
  const int pc = lane >> 2, pq = lane & 3;
  const double* prd = pred + pc * LDP + pq;
#pragma unroll
  for (int k = 0; k < FT_B; ++k) {
    double rk[FT_B];
#pragma unroll
    for (int c = 0; c < FT_B; ++c) {
      if (STAGED) rk[c] = (c >= k) ? rjj[k * FT_B + c] : 0.0;
      else rk[c] = (c >= k) ? __shfl_sync(0xffffffffu, rrow[c], k) : 0.0;
    }
    // red[c] = a_k . a_c (c >= k)  and  red[l] = v_l . a_k (l < k)
    double red[FT_B];
    {
#pragma unroll
      for (int i = 0; i < FT_B; ++i) {
        double pp = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; ++r) pp += a[r][k] * a[r][i];
        pred[i * LDP + lane] = pp;
      }
      __syncwarp();
      double x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = prd[4 * i];
      double tsum = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
      tsum += __shfl_xor_sync(0xffffffffu, tsum, 1);
      tsum += __shfl_xor_sync(0xffffffffu, tsum, 2);
      double* rb = scr + ((k & 1) << 3);
      rb[pc] = tsum;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < FT_B; c += 2) {
        const double2 t2 = *reinterpret_cast<const double2*>(rb + c);
        red[c] = t2.x;
        red[c + 1] = t2.y;
      }
    }
    const double alpha = rk[k], sig2 = red[k];
    double tau, sc, beta;
    {
      // branch-free: a column whose A part is exactly zero (or whose [alpha; a] norm^2 is below 1e-280) gets tau = 0
      const double n2 = fma(alpha, alpha, sig2);
      const bool ok = (sig2 > 0.0) && (n2 > 1e-280);  // (inf / NaN flow through and are flagged downstream)
      const double n2s = ok ? n2 : 1.0;
      const double rs = ft_rsqrt_fast(n2s);
      const double nrm = n2s * rs;
      const double bt = alpha >= 0.0 ? -nrm : nrm;
      const double u = alpha - bt;  // |u| = |alpha| + nrm >= nrm > 0
      const double tt = alpha >= 0.0 ? u * rs : -u * rs;
      const double rc = ft_rcp_fast(u);
      tau = ok ? tt : 0.0;
      sc = ok ? rc : 0.0;
      beta = ok ? bt : alpha;
    }
    double v[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
      v[r] = a[r][k] * sc;
      a[r][k] = v[r];
    }
#pragma unroll
    for (int c = k + 1; c < FT_B; ++c) {
      const double s = tau * (rk[c] + sc * red[c]);
#pragma unroll
      for (int r = 0; r < RPL; ++r) a[r][c] -= s * v[r];
      rk[c] -= s;
    }
    rk[k] = beta;
    if (STAGED) {
      if (lane == 0) {
#pragma unroll
        for (int c = k; c < FT_B; ++c) rjj[k * FT_B + c] = rk[c];
      }
    } else {
      if (lane == k) {
#pragma unroll
        for (int c = k; c < FT_B; ++c) rrow[c] = rk[c];
      }
    }
    {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < k; ++l) acc += Trow[l] * (sc * red[l]);
      Trow[k] = (lane == k) ? tau : ((lane < k) ? -tau * acc : 0.0);
    }
  }
  if (STAGED) {
    __syncwarp();
#pragma unroll
    for (int e = lane; e < FT_B * FT_B; e += 32) {
      const int r = e >> 3, c = e & 7;
      if (c >= r && j0 + r < n && j0 + c < n) R[(size_t)(j0 + r) * ldr + j0 + c] = rjj[e];
    }
  } else if (lane < FT_B && j0 + lane < n) {
#pragma unroll
    for (int c = 0; c < FT_B; ++c)
      if (c >= lane && j0 + c < n) R[(size_t)(j0 + lane) * ldr + j0 + c] = rrow[c];
  }
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    const int row = lane + 32 * r;
    if (row < H) {
#pragma unroll
      for (int k = 0; k < FT_B; ++k) Vt[k * LDV + row] = a[r][k];
    }
  }
  if (lane < FT_B) {
#pragma unroll
    for (int y = 0; y < FT_B; ++y) Tm[lane * FT_B + y] = Trow[y];
  }
}

#ifndef FT_PANEL_LEAN
#define FT_PANEL_LEAN 2
#endif
#ifndef FT_PANEL_DMMA
#define FT_PANEL_DMMA 0  // measured slower on B200 (10.5 vs 11.5 TF/s at 1600x400): the panel warp is issue-bound and this variant issues ~245 instead of ~200 instructions per column
#endif

// Same factorisation, column-per-lane-group layout: lane (g, q4) holds column j0+g at the rows 4s+q4 -- exactly the DMMA
// A- AND B-fragment of the panel, so the dots of a column step come from the tensor pipe (G = P^T P, H/4 DMMAs whose
// k-reduction replaces the five dependent shuffle levels of ft_panel) and the reflector update touches one column per
// lane.  Row k of G sits in the four lanes g = k; eight broadcasts hand it to every lane, the scalar path is ft_panel's.
template <int H>
__device__ __forceinline__ void ft_panel_g(const double* Ablk, const int ld, const int j0, const int n, double* __restrict__ R,
                                           const int ldr, double* Vt, double* Tm, const double* rjj_s = nullptr) {
  constexpr int LDV = H + 4;
  constexpr int NS = H / 4;  // rows per lane
  const int lane = threadIdx.x & 31, g = lane >> 2, q4 = lane & 3;
  double a[NS];
  {
    const double* bp = Ablk + (size_t)q4 * ld + ((j0 + g) ^ (((q4 >> 1) & 1) << 2));
#pragma unroll
    for (int s = 0; s < NS; ++s) a[s] = bp[(size_t)(4 * s) * ld];
  }
  double rrow[FT_B];
#pragma unroll
  for (int c = 0; c < FT_B; ++c) {
    if (rjj_s) rrow[c] = (lane < FT_B && c >= lane) ? rjj_s[lane * FT_B + c] : 0.0;
    else rrow[c] = (lane < FT_B && j0 + lane < n && j0 + c < n && c >= lane) ? R[(size_t)(j0 + lane) * ldr + j0 + c] : 0.0;
  }
  double T[FT_B][FT_B];
#pragma unroll
  for (int x = 0; x < FT_B; ++x)
#pragma unroll
    for (int y = 0; y < FT_B; ++y) T[x][y] = 0.0;
#pragma unroll
  for (int k = 0; k < FT_B; ++k) {
    // column k at this lane's rows (same q4 group, lane g = k); independent of the dots below
    double ak[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) ak[s] = __shfl_sync(0xffffffffu, a[s], 4 * k + q4);
    // G = P^T P on the tensor pipe, two accumulator chains
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
    for (int s = 0; s < NS; s += 2) {
      dmma884(c0, c1, a[s], a[s]);
      dmma884(e0, e1, a[s + 1], a[s + 1]);
    }
    c0 += e0;
    c1 += e1;
    // red[c] = G[k][c]: a_k . a_c (c >= k),  v_c . a_k (c < k)
    double red[FT_B];
#pragma unroll
    for (int c = 0; c < FT_B; ++c) red[c] = __shfl_sync(0xffffffffu, (c & 1) ? c1 : c0, 4 * k + (c >> 1));
    double rk[FT_B];
#pragma unroll
    for (int c = 0; c < FT_B; ++c) rk[c] = (c >= k) ? __shfl_sync(0xffffffffu, rrow[c], k) : 0.0;
    const double alpha = rk[k], sig2 = red[k];
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
    // a_c -= s_c v (c > k), a_k <- v = a_k sc: this lane's column only
    double coef = (g == k) ? sc : 0.0;
#pragma unroll
    for (int c = k + 1; c < FT_B; ++c) {
      const double s = tau * (rk[c] + sc * red[c]);
      rk[c] -= s;
      if (g == c) coef = -(s * sc);
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) a[s] = fma(coef, ak[s], (g == k) ? 0.0 : a[s]);
    if (lane == k) {
      rrow[k] = beta;
#pragma unroll
      for (int c = k + 1; c < FT_B; ++c) rrow[c] = rk[c];
    }
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
#pragma unroll
  for (int s = 0; s < NS; ++s) Vt[g * LDV + 4 * s + q4] = a[s];
#pragma unroll
  for (int x = 0; x < FT_B; ++x)
    if (lane == x) {
#pragma unroll
      for (int y = 0; y < FT_B; ++y) Tm[x * FT_B + y] = T[x][y];
    }
}

template <int H>
__device__ __forceinline__ void ft_panel_sel(const double* Ablk, const int ld, const int j0, const int n, double* __restrict__ R,
                                             const int ldr, double* Vt, double* Tm, double* scr, double* pred, double* rjj_s = nullptr) {
#if FT_PANEL_DMMA
  ft_panel_g<H>(Ablk, ld, j0, n, R, ldr, Vt, Tm, rjj_s);
#elif FT_PANEL_LEAN == 2
  if (rjj_s) ft_panel_m<H, true>(Ablk, ld, j0, n, R, ldr, Vt, Tm, rjj_s, scr, pred);
  else ft_panel_m<H, false>(Ablk, ld, j0, n, R, ldr, Vt, Tm, nullptr, scr, pred);
#elif FT_PANEL_LEAN
  ft_panel_l<H>(Ablk, ld, j0, n, R, ldr, Vt, Tm, rjj_s, scr);
#else
  ft_panel<H>(Ablk, ld, j0, n, R, ldr, Vt, Tm, rjj_s);
#endif
}

// fragments of V^T / T of the current panel held in registers by an updating warp
template <int H>
struct FtFrags {
  double va1[H / 4], va2[H / 8][2], at[2];
  __device__ __forceinline__ void load(const double* Vt, const double* Tm) {
    constexpr int LDV = H + 4;
    const int lane = threadIdx.x & 31, g = lane >> 2, q4 = lane & 3;
#pragma unroll
    for (int s = 0; s < H / 4; ++s) va1[s] = Vt[g * LDV + 4 * s + q4];
#pragma unroll
    for (int r = 0; r < H / 8; ++r)
#pragma unroll
      for (int s = 0; s < 2; ++s) va2[r][s] = Vt[(4 * s + q4) * LDV + 8 * r + g];
#pragma unroll
    for (int s = 0; s < 2; ++s) at[s] = Tm[(4 * s + q4) * FT_B + g];  // T^T[g][4s+q4]
  }
};

// update of one 8-column slab at column c0 by panel rows j0..j0+7 (warp-local).  (r0, r1) = R_jc fragment.
template <int H>
__device__ __forceinline__ void ft_update_slab(double* Ablk, const int ld, const int c0, const FtFrags<H>& f, double* Ws,
                                               double* rp, const bool ok0, const bool ok1, const double r0, const double r1) {
  const int lane = threadIdx.x & 31, g = lane >> 2, q4 = lane & 3;
  const int cc = c0 + 2 * q4;
  double w0 = 0.0, w1 = 0.0, x0 = 0.0, x1 = 0.0;
  // B fragments of GEMM1: rows 4s+q4 (swizzle bit = q4 bit 1), column c0+g
  const double* bp = Ablk + (size_t)q4 * ld + ((c0 + g) ^ (((q4 >> 1) & 1) << 2));
  double bf[H / 4];
#pragma unroll
  for (int s = 0; s < H / 4; ++s) bf[s] = bp[(size_t)(4 * s) * ld];
  // C tiles of GEMM2 (rows 8r+g, columns cc, cc+1), loaded early so that their latency hides behind GEMM1
  double2* cp[H / 8];
  double2 cv[H / 8];
#pragma unroll
  for (int r = 0; r < H / 8; ++r) {
    cp[r] = reinterpret_cast<double2*>(Ablk + (size_t)(8 * r + g) * ld + (cc ^ ft_sw(g)));
    cv[r] = *cp[r];
  }
#pragma unroll
  for (int s = 0; s < H / 4; s += 2) {
    dmma884(w0, w1, f.va1[s], bf[s]);
    dmma884(x0, x1, f.va1[s + 1], bf[s + 1]);
  }
  w0 += x0 + r0;
  w1 += x1 + r1;
  // per-warp 8x8 scratch, ld 8, column swizzle 4*((row>>1)&1): C-layout 128-bit stores and B-layout loads conflict-free
  double2* wst = reinterpret_cast<double2*>(Ws + g * 8 + ((2 * q4) ^ ft_sw(g)));
  const double* wld0 = Ws + q4 * 8 + (g ^ ft_sw(q4));
  const double* wld1 = Ws + (4 + q4) * 8 + (g ^ ft_sw(4 + q4));
  __syncwarp();
  *wst = make_double2(w0, w1);
  __syncwarp();
  double p0 = 0.0, p1 = 0.0;
  dmma884(p0, p1, f.at[0], *wld0);
  dmma884(p0, p1, f.at[1], *wld1);
  if (ok0) rp[0] = r0 - p0;
  if (ok1) rp[1] = r1 - p1;
  __syncwarp();
  *wst = make_double2(-p0, -p1);
  __syncwarp();
  const double b0 = *wld0, b1 = *wld1;
#pragma unroll
  for (int r = 0; r < H / 8; ++r) dmma884(cv[r].x, cv[r].y, f.va2[r][0], b0);
#pragma unroll
  for (int r = 0; r < H / 8; ++r) dmma884(cv[r].x, cv[r].y, f.va2[r][1], b1);
#pragma unroll
  for (int r = 0; r < H / 8; ++r) *cp[r] = cv[r];
}

// A: m x n row-major (lda), read-only.  R: n x n row-major (ldr) in global memory, fully overwritten with the
// upper-triangular factor (R^T R = A^T A).  For m < n use the matrix itself as the factor instead (callers do).
// If normalize: R is divided by its max-abs.
// tri_n > 0: A is a stack of m / tri_n upper-triangular tri_n x n blocks (the chunk factors of a TSQR split): the first
// triangle is copied into R instead of being factored, and a row block that starts at local row r of its triangle skips
// the panels left of column r (they hold only zeros there) -- the merge then costs ~0.45 of a dense stack.
template <int H>
__device__ void qr_ft_cta(const double* __restrict__ A, const int m, const int n, const int lda, double* __restrict__ R,
                          const int ldr, const bool normalize, double* smem, const int tri_n = 0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n8 = (n + 7) & ~7;
  const int ld = ft_ld(n);
  constexpr int LDV = H + 4;
  constexpr int VT_SZ = FT_B * LDV + FT_B * FT_B;
  double* Ablk = smem;                       // H x ld
  double* VT0 = Ablk + (size_t)H * ld;       // 2 x (V^T 8 x LDV, T 8x8)
  double* Ws = VT0 + 2 * VT_SZ + warp * 72;  // per-warp 8x8 scratch (ld 9)
  constexpr bool PREF = (H == 64);             // warp 4 stages the panel warp's R blocks one panel ahead
  double* Rst = VT0 + 2 * VT_SZ + NW * 72 + 16;  // 2 x 128 doubles (H = 64 only)
  double* Pred = Rst + (H == 64 ? 2 * 128 : 0);   // 8 x 36 partial-dot scratch of the panel warp
  if (tri_n > 0) {
    for (int idx = tid; idx < n * n; idx += NT) {
      const int i = idx / n, c = idx % n;
      R[(size_t)i * ldr + c] = (c >= i && i < tri_n) ? A[(size_t)i * lda + c] : 0.0;
    }
  } else {
    for (int idx = tid; idx < n * n; idx += NT) R[(size_t)(idx / n) * ldr + (idx % n)] = 0.0;
  }
  const int g = lane >> 2, q4 = lane & 3;  // DMMA fragment coordinates
  const int npanel = n8 / FT_B;
  // TMA bulk staging needs 16-byte aligned, 16-byte-multiple rows (even n and lda); odd bonds use the plain loop
  __shared__ uint64_t ft_mbar;
  const bool bulk = ((n & 1) == 0) && ((lda & 1) == 0) && aligned16(A) && aligned16(smem);
  uint32_t phase = 0;
  if (bulk && tid == 0) mbar_init(&ft_mbar, 1);
  // row blocks: dense -> [row0, row0+H) ; triangle stack -> blocks never straddle two triangles
  const int seg_rows = tri_n > 0 ? tri_n : m;
  const int nseg = tri_n > 0 ? m / tri_n : 1;
  const int blocks_per_seg = (seg_rows + H - 1) / H;
  for (int blk = (tri_n > 0 ? blocks_per_seg : 0); blk < nseg * blocks_per_seg; ++blk) {
    const int seg = blk / blocks_per_seg, lrow0 = (blk % blocks_per_seg) * H;
    const int row0 = seg * seg_rows + lrow0;
    const int row_end = seg * seg_rows + seg_rows;      // rows of this block stay below row_end
    const int jp0 = tri_n > 0 ? min(lrow0 / FT_B, npanel - 1) : 0;  // first panel with a non-zero column in this block
    __syncthreads();
    // ---- stage the row block (zero padded) ----
    if (bulk) {
      // one TMA bulk copy per row (n contiguous doubles), completion on the mbarrier; the other threads zero the padding
      const int vrows = min(H, min(row_end, m) - row0);
      if (warp == 0) {
        if (lane == 0) {
          fence_proxy_async();
          mbar_expect_tx(&ft_mbar, (uint32_t)(vrows * n * 8));
        }
        __syncwarp();
        const uint64_t pol = l2_policy_evict_first();
        for (int i = lane; i < vrows; i += 32) bulk_g2s_hint(Ablk + (size_t)i * ld, A + (size_t)(row0 + i) * lda, (uint32_t)(n * 8), &ft_mbar, pol);
      }
      const int npad = n8 - n;
      for (int idx = tid; idx < vrows * npad; idx += NT) Ablk[(size_t)(idx / npad) * ld + n + idx % npad] = 0.0;
      for (int idx = tid; idx < (H - vrows) * n8; idx += NT) Ablk[(size_t)(vrows + idx / n8) * ld + idx % n8] = 0.0;
      mbar_wait(&ft_mbar, phase);
      phase ^= 1;
      __syncthreads();
      // swizzle in place: rows with bit 1 set swap the two halves of every 8-column group (c -> c ^ 4)
      const int ng = n8 >> 3;
      for (int idx = tid; idx < (H / 2) * ng * 4; idx += NT) {
        const int e = idx & 3, gq = (idx >> 2) % ng, ih = (idx >> 2) / ng;
        const int i = ((ih >> 1) << 2) + 2 + (ih & 1);
        double* p = Ablk + (size_t)i * ld + 8 * gq + e;
        const double t0 = p[0];
        p[0] = p[4];
        p[4] = t0;
      }
    } else {
      for (int idx = tid; idx < H * n8; idx += NT) {
        const int i = idx / n8, c = idx % n8;
        const int gi = row0 + i;
        Ablk[(size_t)i * ld + (c ^ ft_sw(i))] = (gi < row_end && gi < m && c < n) ? A[(size_t)gi * lda + c] : 0.0;
      }
    }
    __syncthreads();
    if (warp == 0) ft_panel_sel<H>(Ablk, ld, jp0 * FT_B, n, R, ldr, VT0 + (jp0 & 1) * VT_SZ, VT0 + (jp0 & 1) * VT_SZ + FT_B * LDV, Ws, Pred);
    if (PREF && warp == 4 && jp0 + 1 < npanel) ft_prefetch_R(R, ldr, n, jp0, Rst + (jp0 & 1) * 128);
    for (int jp = jp0; jp < npanel; ++jp) {
      const int j0 = jp * FT_B;
      double* Vt = VT0 + (jp & 1) * VT_SZ;
      double* Tm = Vt + FT_B * LDV;
      double* Vtn = VT0 + ((jp + 1) & 1) * VT_SZ;
      __syncthreads();  // V/T of panel jp ready; update of panel jp-1 complete
      const int nslab = npanel - jp - 1;
      if (nslab <= 0) continue;
      const int rr = j0 + g;
      FtFrags<H> f;
      f.load(Vt, Tm);
      if (warp == 0) {
        // look-ahead: bring the next panel's columns up to date, then factor it while warps 1-7 update the rest
        const int c0 = j0 + FT_B, cc = c0 + 2 * q4;
        const bool ok0 = (rr < n) && (cc < n), ok1 = (rr < n) && (cc + 1 < n);
        double* rp = R + (size_t)rr * ldr + cc;
        double* rs = Rst + (jp & 1) * 128;
        const double r0 = PREF ? rs[g * FT_B + 2 * q4] : (ok0 ? rp[0] : 0.0);
        const double r1 = PREF ? rs[g * FT_B + 2 * q4 + 1] : (ok1 ? rp[1] : 0.0);
        ft_update_slab<H>(Ablk, ld, c0, f, Ws, rp, ok0, ok1, r0, r1);
        __syncwarp();
        ft_panel_sel<H>(Ablk, ld, c0, n, R, ldr, Vtn, Vtn + FT_B * LDV, Ws, Pred, PREF ? rs + 64 : nullptr);
      } else {
        // slabs 1 .. nslab-1 over the updating warps.  With H = 64 (one CTA per SM) warp 4, which shares the SM
        // sub-partition and its FP64 pipe with the panel warp, sits out: the step is bound by the panel's dependent
        // FP64 chain, which must not queue behind DMMAs (measured +7 % on isolated 1600..4000 x 400 batches).
        constexpr bool IDLE = (H == 64);
        constexpr int NUPD = IDLE ? NW - 2 : NW - 1;
        if (IDLE && warp == 4) {
          if (PREF && jp + 2 < npanel) ft_prefetch_R(R, ldr, n, jp + 1, Rst + ((jp + 1) & 1) * 128);
          continue;
        }
        int sl = (!IDLE || warp < 4) ? warp : warp - 1;
        double nr0 = 0.0, nr1 = 0.0;
        if (sl < nslab && rr < n) {
          const int cc = j0 + FT_B + sl * FT_B + 2 * q4;
          if (cc < n) nr0 = R[(size_t)rr * ldr + cc];
          if (cc + 1 < n) nr1 = R[(size_t)rr * ldr + cc + 1];
        }
        for (; sl < nslab; sl += NUPD) {
          const int c0 = j0 + FT_B + sl * FT_B, cc = c0 + 2 * q4;
          const bool ok0 = (rr < n) && (cc < n), ok1 = (rr < n) && (cc + 1 < n);
          double* rp = R + (size_t)rr * ldr + cc;
          const double r0 = nr0, r1 = nr1;
          {
            const int ccn = cc + NUPD * FT_B;
            const bool more = (sl + NUPD < nslab) && (rr < n);
            nr0 = (more && ccn < n) ? rp[NUPD * FT_B] : 0.0;
            nr1 = (more && ccn + 1 < n) ? rp[NUPD * FT_B + 1] : 0.0;
          }
          ft_update_slab<H>(Ablk, ld, c0, f, Ws, rp, ok0, ok1, r0, r1);
        }
      }
    }
  }
  __syncthreads();
  if (bulk && tid == 0) mbar_inval(&ft_mbar);
  if (normalize) {
    __shared__ double redn[NW + 1];
    double mx = 0.0;
    for (int idx = tid; idx < n * n; idx += NT) {
      const int i = idx / n, c = idx % n;
      if (c >= i) mx = fmax(mx, fabs(R[(size_t)i * ldr + c]));
    }
    mx = block_max(mx, redn);
    if (mx > 0.0 && isfinite(mx)) {
      const double fs = 1.0 / mx;
      for (int idx = tid; idx < n * n; idx += NT) {
        const int i = idx / n, c = idx % n;
        if (c >= i) R[(size_t)i * ldr + c] *= fs;
      }
    }
  }
  __syncthreads();
}

}  // namespace mpbp
