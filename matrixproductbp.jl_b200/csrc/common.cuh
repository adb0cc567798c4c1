// common.cuh -- shared device-side definitions of the MPBP engine (sm_100a, FP64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace mpbp {

constexpr int NT = 256;  // threads per CTA for every engine kernel
constexpr int NW = NT / 32;

// Reference to one tensor train held on the device (arena TT or message slot).
// site t lives at data + t*stride, stored column-major [m, n, p] with the ACTUAL dims
// (bonds[t], bonds[t+1], P).  value(x) = exp(*ls) * prod_t A_t[:, :, x_t].
struct TTRef {
  double* data;
  int* bonds;  // L+1
  double* ls;  // log-scale
  int stride;  // doubles per site slot (capacity)
  int P;       // physical size of a site
};

// truncation policy (TensorTrains.TruncBond / TruncThresh / TruncBondThresh)
struct Trunc {
  int kind;  // 0 bond, 1 thresh, 2 bond+thresh
  int d;
  double eps;
};

// error flags raised by kernels (bit mask in a device int)
enum { ERR_BOND_OVERFLOW = 1, ERR_NAN = 2, ERR_JACOBI_NOCONV = 4 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// CTA-wide sum of K values per thread; result broadcast to all threads through `out` (smem, >= K doubles).
// scratch: smem, >= NW*K doubles.  Two __syncthreads.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* scratch, double* out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) scratch[w * K + k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0;
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) s += scratch[ww * K + threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

// CTA-wide max of one value; broadcast.  scratch >= NW+1 doubles.
__device__ __forceinline__ double block_max(double v, double* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double r = scratch[0];
#pragma unroll
  for (int ww = 1; ww < NW; ++ww) r = fmax(r, scratch[ww]);
  return r;
}
__device__ __forceinline__ double block_sum1(double v, double* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double r = 0;
#pragma unroll
  for (int ww = 0; ww < NW; ++ww) r += scratch[ww];
  return r;
}

// number of singular values kept by the policy; s[] sorted descending, n > 0.
__device__ __forceinline__ int trunc_keep(const Trunc& tr, const double* s, int n) {
  int k = n;
  if (tr.kind == 1 || tr.kind == 2) {
    double nrm = 0;
    for (int i = 0; i < n; ++i) nrm += s[i] * s[i];
    nrm = sqrt(nrm);
    int last = 0;
    for (int i = 0; i < n; ++i)
      if (s[i] > tr.eps * nrm) last = i + 1;
    k = last > 0 ? last : 1;
  }
  if (tr.kind == 0 || tr.kind == 2) k = min(k, tr.d);
  return k;
}

}  // namespace mpbp
