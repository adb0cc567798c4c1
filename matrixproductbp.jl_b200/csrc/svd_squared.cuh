// svd_squared.cuh -- blocked subspace iteration with ONE orthonormalisation per application of M M^T, the block kept in
// shared memory and orthonormalised / Rayleigh-Ritz'ed by one-sided Jacobi (round-1 algorithm).
// Used when the p x b and n x b blocks fit shared memory; the un-squared Householder variant of svd_block.cuh covers the rest
// (D = 900).  Reference code replaced: the truncating SVD inside TensorTrains.compress!, src/recursive_bp_factor.jl:127.
#pragma once
#include "common.cuh"
#include "jacobi.cuh"

namespace mpbp {

__device__ inline void normalize_cols_sq(double* W, int rows, int b, const double* sig, const bool squared = false) {
  double smax = 0.0;
  for (int j = 0; j < b; ++j) smax = fmax(smax, sig[j]);
  if (squared) smax *= smax;
  for (int j = threadIdx.x >> 5; j < b; j += NW) {
    const double f = jacobi_inv_sigma(squared ? sig[j] * sig[j] : sig[j], smax);
    for (int k = threadIdx.x & 31; k < rows; k += 32) W[k + (size_t)j * rows] *= f;
  }
  __syncthreads();
}
// OUT[r + rows_out*j] = sum_k MT(k, r) * W[k + kdim*j]   with M addressed as M[a + p*rr]
// OUT (rows_out x b) = Mx W  where Mx (rows_out x kdim) is column-major with leading dimension rows_out
// (thread per output row: consecutive lanes read consecutive addresses of Mx), W (kdim x b) in shared memory.
__device__ inline void sub_gemm(const double* __restrict__ Mx, int rows_out, int kdim, const double* W, int b, double* OUT) {
  for (int r = threadIdx.x; r < rows_out; r += NT) {
    for (int j0 = 0; j0 < b; j0 += 16) {
      double acc[16];
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) acc[jj] = 0.0;
      const double* w = W + (size_t)j0 * kdim;
      const int nj = min(16, b - j0);
      // the Mx loads come from L2: keep 8 of them in flight per thread
      int k = 0;
      for (; k + 8 <= kdim; k += 8) {
        double m8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) m8[u] = Mx[r + (size_t)rows_out * (k + u)];
        if (nj == 16 && (kdim & 1) == 0) {
          // 128-bit shared-memory loads: two consecutive k per load (kdim even -> 16-byte aligned)
#pragma unroll
          for (int u = 0; u < 8; u += 2)
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const double2 ww = *reinterpret_cast<const double2*>(w + k + u + (size_t)jj * kdim);
              acc[jj] += m8[u] * ww.x;
              acc[jj] += m8[u + 1] * ww.y;
            }
        } else if (nj == 16) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) acc[jj] += m8[u] * w[k + u + (size_t)jj * kdim];
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) acc[jj] += m8[u] * w[k + u + (size_t)jj * kdim];
        }
      }
      for (; k < kdim; ++k) {
        const double m = Mx[r + (size_t)rows_out * k];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj)
          if (jj < nj) acc[jj] += m * w[k + (size_t)jj * kdim];
      }
#pragma unroll
      for (int jj = 0; jj < 16; ++jj)
        if (jj < nj) OUT[r + (size_t)rows_out * (j0 + jj)] = acc[jj];
    }
  }
}

// On return the p x b block W (shared memory, lda = p) holds orthonormal left singular vectors, sig / order (b entries) their
// singular values and descending order; returns ||M||_F^2.  Qg >= p*64, Zg >= n*64, Mt >= p*n doubles of global scratch.
__device__ inline double svd_subspace_squared(const double* Mcm, double* Qg, double* Zg, double* Mt, const int p, const int rn,
                                              const Trunc tr, const int dcap, const int jac_doubles, double* sig, int* order,
                                              double* sprev, double* W, int* flagp, int* s_donep, double* red, int* err,
                                              double* stats, int* b_out) {
  int& flag = *flagp;
  int& s_done = *s_donep;
  const int c = min(p, rn);
  double nrm2_all = -1.0;
  {
    const int n = rn;
    const double* M = Mcm;  // column-major p x n
    long long tph = clock64();
    auto phase = [&](int which) {
      __syncthreads();
      const long long now = clock64();
      if (stats && threadIdx.x == 0) atomicAdd(stats + 8 + which, (double)(now - tph));
      tph = now;
    };
    for (int idx = threadIdx.x; idx < p * n; idx += NT) Mt[(idx / p) + (size_t)n * (idx % p)] = M[idx];
    phase(0);
    int b = min(min(max(SUB_BLOCK, min(SUB_BMAX, 2 * (tr.kind == 1 ? dcap : tr.d) + 8)), c), jac_doubles / max(p, n));
    b &= ~7;
    if (b < 8 || b < min(c, (tr.kind == 1 ? dcap : tr.d))) {
      if (threadIdx.x == 0) atomicOr(err, ERR_BOND_OVERFLOW);  // shared memory cannot hold a block wide enough
      b = max(b, 8);
    }
    // ---- start block: the b largest-norm columns of M ----
    double* nrm = W;
    int* sel = reinterpret_cast<int*>(W + n);
    for (int rr = threadIdx.x >> 5; rr < n; rr += NW) {
      double s = 0.0;
      for (int a = threadIdx.x & 31; a < p; a += 32) { const double x = M[a + (size_t)p * rr]; s += x * x; }
      s = warp_sum(s);
      if ((threadIdx.x & 31) == 0) nrm[rr] = s;
    }
    __syncthreads();
    double fro = 0.0;
    for (int rr = threadIdx.x; rr < n; rr += NT) fro += nrm[rr];
    nrm2_all = block_sum1(fro, red);
    for (int rr = threadIdx.x; rr < n; rr += NT) {
      const double sj = nrm[rr];
      int rank = 0;
      for (int i = 0; i < n; ++i) rank += (nrm[i] > sj) || (nrm[i] == sj && i < rr);
      if (rank < b) sel[rank] = rr;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < b; j += NT) order[j] = sel[j];
    __syncthreads();
    for (int idx = threadIdx.x; idx < p * b; idx += NT) W[idx] = M[(idx % p) + (size_t)p * order[idx / p]];
    for (int j = threadIdx.x; j < SUB_BMAX; j += NT) sprev[j] = 0.0;
    __syncthreads();
    phase(1);
    // start block: two sweeps are enough (only a well-conditioned basis is needed here)
    int sw = 0;
    jacobi_cols(W, p, b, p, &flag, 2);
    jacobi_sort(W, p, b, p, sig, order);
    normalize_cols_sq(W, p, b, sig);
    phase(2);
    int extra = -1;
    const int kchk = min(b, tr.kind == 1 ? dcap : tr.d);
    int nit = 0;
    for (int it = 0; it < SUB_MAXIT; ++it) {
      // one application of M M^T per iteration, ONE orthonormalisation: the kept singular values span only a few
      // orders of magnitude, so the squared spectrum of the block stays far inside FP64 range
      sub_gemm(Mt, n, p, W, b, Zg);  // Z = M^T Q   (n x b, not orthonormalised)
      __syncthreads();
      for (int idx = threadIdx.x; idx < n * b; idx += NT) W[idx] = Zg[idx];
      __syncthreads();
      sub_gemm(M, p, n, W, b, Qg);  // Y = M Z
      __syncthreads();
      for (int idx = threadIdx.x; idx < p * b; idx += NT) W[idx] = Qg[idx];
      phase(3);
      jacobi_cols(W, p, b, p, &flag);
      jacobi_sort(W, p, b, p, sig, order);
      for (int j = threadIdx.x; j < b; j += NT) sig[j] = sqrt(sig[j]);  // singular values of M (Y ~ U Sigma^2)
      __syncthreads();
      ++nit;
      if (threadIdx.x == 0) {
        // Ritz values of the squared iteration carry a noise floor eps*sigma_1^2/sigma_i: tolerate it here, the
        // final un-squared refinement below restores the small directions
        const double s1 = sig[order[0]];
        bool conv = true;
        for (int i = 0; i < kchk; ++i) {
          const double s = sig[order[i]];
          const double tol_i = 1e-13 * s1 + 8e-16 * s1 * s1 / fmax(s, 1e-300);
          conv = conv && (fabs(s - sprev[i]) <= tol_i);
          sprev[i] = s;
        }
        if (extra < 0 && conv) extra = 1;
        else if (extra > 0) extra--;
        s_done = (extra == 0);
      }
      __syncthreads();
      normalize_cols_sq(W, p, b, sig, true);
      phase(4);
      if (s_done) break;
    }
    const bool converged = s_done;
    if (!converged) {
      // The block iteration hit its cap (clustered singular values across the block edge).  Exact fallback, rare
      // and slow on purpose: one-sided Jacobi on ALL n columns of M in global memory (work copy in Mt, whose
      // transposed copy is no longer needed), then the b leading columns become the block.  Counted in stats[4].
      if (threadIdx.x == 0)
        printf("[mpbp] subspace SVD not converged after %d iterations (p=%d n=%d b=%d, s1=%.3e s_k=%.3e): exact Jacobi fallback\n",
               nit, p, n, b, sig[order[0]], sig[order[kchk - 1]]);
      __syncthreads();
      for (int idx = threadIdx.x; idx < p * n; idx += NT) Mt[idx] = M[idx];
      __syncthreads();
      sw = max(sw, jacobi_cols(Mt, p, n, p, &flag));
      double* sall = W;                                  // n doubles
      int* oall = reinterpret_cast<int*>(W + n);          // n ints
      jacobi_sort(Mt, p, n, p, sall, oall);
      for (int idx = threadIdx.x; idx < p * b; idx += NT) Qg[idx] = Mt[(idx % p) + (size_t)p * oall[idx / p]];
      for (int j = threadIdx.x; j < b; j += NT) sprev[j] = sall[oall[j]];
      __syncthreads();
      for (int idx = threadIdx.x; idx < p * b; idx += NT) W[idx] = Qg[idx];
      for (int j = threadIdx.x; j < b; j += NT) {
        sig[j] = sprev[j];
        order[j] = j;
      }
      __syncthreads();
      normalize_cols_sq(W, p, b, sig);
      phase(5);
    } else {
      // final un-squared Rayleigh-Ritz refinement: Z = orth(M^T Q), Y = M Z, SVD(Y) -> U, sigma
      sub_gemm(Mt, n, p, W, b, Zg);
      __syncthreads();
      for (int idx = threadIdx.x; idx < n * b; idx += NT) W[idx] = Zg[idx];
      __syncthreads();
      sw = max(sw, jacobi_cols(W, n, b, n, &flag));
      jacobi_sort(W, n, b, n, sig, order);
      normalize_cols_sq(W, n, b, sig);
      sub_gemm(M, p, n, W, b, Qg);
      __syncthreads();
      for (int idx = threadIdx.x; idx < p * b; idx += NT) W[idx] = Qg[idx];
      __syncthreads();
      sw = max(sw, jacobi_cols(W, p, b, p, &flag));
      jacobi_sort(W, p, b, p, sig, order);
      normalize_cols_sq(W, p, b, sig);
      phase(5);
    }
    if (stats && threadIdx.x == 0) {
      atomicAdd(stats + 0, 1.0);
      atomicAdd(stats + 1, (double)nit);
      atomicAdd(stats + 2, (double)b);
      atomicAdd(stats + 3, (double)sw);
      if (!converged) atomicAdd(stats + 4, 1.0);  // hit SUB_MAXIT: the exact Jacobi fallback above produced the result
    }
    if (sw >= JACOBI_MAX_SWEEPS && threadIdx.x == 0) atomicOr(err, ERR_JACOBI_NOCONV);
    *b_out = b;
  }
  return nrm2_all;
}

}  // namespace mpbp
