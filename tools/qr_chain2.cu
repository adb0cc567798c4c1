// qr_chain2.cu -- PROTOTYPE (not part of the product library; measured and rejected in round 2): flat-tree DMMA Q-less QR
// with TWO alternating panel-chain warps and look-ahead depth 2.  Written at the end of round 1 from the measured
// per-panel budget of the product kernel (profiles/r01_qr_ft_ncu.md): on the panel warp, barrier 800 + look-ahead
// slab update 2240 + panel load 435 + eight Householder columns 4970 cycles, i.e. the serial chain binds on every
// panel.  Here the slab update and the panel load leave the critical path:
//
//   step s (one __syncthreads per step, as in the product kernel):
//     producer  C(s)   = warp 0 (s even) / warp 4 (s odd): factors panel s from REGISTERS, publishing every reflector
//                        (v_k, tau_k) to shared memory as soon as it exists (named barrier 1+k, bar.arrive);
//     consumer  C(s+1) = the other chain warp: (i) applies the finished panel s-1 to slab s+1 with the usual DMMA slab
//                        update, (ii) loads slab s+1 into registers in panel layout, (iii) applies the reflectors of
//                        panel s one by one as they are published (bar.sync 1+k), so that at the end of the step it
//                        holds slab s+1 fully up to date and becomes the producer of step s+1;
//     warps 1,2,3,5,6,7: apply panel s-1 to the slabs >= s+2 (DMMA), exactly as the product kernel does.
//   Expected critical path per panel: ~5000 (columns) + ~350 (last streamed reflector) instead of ~8900 cycles.
//   MEASURED (round 2, profiles/r02_qr_chain2_result.txt): correct on the first run (error 4e-15) but SLOWER than the product
//   kernel, 8.4 vs 10.0 TF/s at 1600x400x148: the per-column publish lengthens the producer chain by more than it saves.
//
// Build / run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/qr_chain2 tools/qr_chain2.cu
//   ./tools/qr_chain2 [m n batch]      (prints max |R^T R - A^T A| / max|A^T A| of matrix 0 and TF/s for both kernels)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../matrixproductbp.jl_b200/csrc/qr_ft.cuh"

namespace mpbp {

constexpr int C2_H = 64;
constexpr int C2_LDV = C2_H + 4;
constexpr int C2_VT = FT_B * C2_LDV + FT_B * FT_B + FT_B;  // V^T, T, tau
constexpr int C2_RPL = C2_H / 32;

__host__ __device__ inline size_t c2_smem_doubles(int n) { return (size_t)C2_H * ft_ld(n) + 2 * C2_VT + NW * 72 + 16; }

__device__ __forceinline__ void c2_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void c2_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// 8 simultaneous warp-wide sums: in[c] (c = 0..7) per lane -> every lane gets all 8 totals (transpose-reduce)
__device__ __forceinline__ void c2_reduce8(const double (&in)[FT_B], double (&out)[FT_B]) {
  const int lane = threadIdx.x & 31;
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  double w4[4], w2[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double send = b4 ? in[i] : in[i + 4], keep = b4 ? in[i + 4] : in[i];
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
  for (int c = 0; c < FT_B; ++c)
    out[c] = __shfl_sync(0xffffffffu, t, (((c >> 2) & 1) << 4) | (((c >> 1) & 1) << 3) | ((c & 1) << 2));
}

__device__ __forceinline__ void c2_load_slab(const double* Ablk, int ld, int c0, double (&a)[C2_RPL][FT_B]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < C2_RPL; ++r) {
    const int row = lane + 32 * r;
#pragma unroll
    for (int c = 0; c < FT_B; ++c) a[r][c] = Ablk[(size_t)row * ld + ((c0 + c) ^ ft_sw(row))];
  }
}

// lane k < 8 holds row (r0 + k) of the 8x8 block of R at (r0, c0); `upper`: only the upper triangle (diagonal block)
__device__ __forceinline__ void c2_load_rblock(const double* R, int ldr, int r0, int c0, int n, bool upper, double (&rrow)[FT_B]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < FT_B; ++c)
    rrow[c] = (lane < FT_B && r0 + lane < n && c0 + c < n && (!upper || c >= lane)) ? R[(size_t)(r0 + lane) * ldr + c0 + c] : 0.0;
}
__device__ __forceinline__ void c2_store_rblock(double* R, int ldr, int r0, int c0, int n, bool upper, const double (&rrow)[FT_B]) {
  const int lane = threadIdx.x & 31;
  if (lane < FT_B && r0 + lane < n) {
#pragma unroll
    for (int c = 0; c < FT_B; ++c)
      if (c0 + c < n && (!upper || c >= lane)) R[(size_t)(r0 + lane) * ldr + c0 + c] = rrow[c];
  }
}

// producer: Householder factorisation of [R_jj ; a] with a already in registers; publishes (v_k, tau_k) per column
__device__ __forceinline__ void c2_factor(double (&a)[C2_RPL][FT_B], const int j0, const int n, double* __restrict__ R, const int ldr,
                                          double* Vt, double* Tm, double* taus, const bool publish) {
  const int lane = threadIdx.x & 31;
  double rpre[FT_B], rrow[FT_B];
  c2_load_rblock(R, ldr, j0, j0, n, true, rpre);
#pragma unroll
  for (int c = 0; c < FT_B; ++c) rrow[c] = 0.0;
  double T[FT_B][FT_B];
#pragma unroll
  for (int x = 0; x < FT_B; ++x)
#pragma unroll
    for (int y = 0; y < FT_B; ++y) T[x][y] = 0.0;
#pragma unroll
  for (int k = 0; k < FT_B; ++k) {
    double part[FT_B], red[FT_B], rk[FT_B];
#pragma unroll
    for (int c = 0; c < FT_B; ++c) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < C2_RPL; ++r) s += a[r][k] * a[r][c];
      part[c] = s;
    }
#pragma unroll
    for (int c = 0; c < FT_B; ++c) rk[c] = (c >= k) ? __shfl_sync(0xffffffffu, rpre[c], k) : 0.0;
    c2_reduce8(part, red);  // red[c] = a_k . a_c (c >= k), = v_c . a_k (c < k)
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
    double v[C2_RPL];
#pragma unroll
    for (int r = 0; r < C2_RPL; ++r) {
      v[r] = a[r][k] * sc;
      a[r][k] = v[r];
      Vt[k * C2_LDV + lane + 32 * r] = v[r];
    }
    if (lane == 0) taus[k] = tau;
    if (publish) {
      __threadfence_block();
      c2_arrive(1 + k);
    }
#pragma unroll
    for (int c = k + 1; c < FT_B; ++c) {
      const double s = tau * (rk[c] + sc * red[c]);
#pragma unroll
      for (int r = 0; r < C2_RPL; ++r) a[r][c] -= s * v[r];
      rk[c] -= s;
    }
    if (lane == k) {
      rrow[k] = beta;
#pragma unroll
      for (int c = k + 1; c < FT_B; ++c) rrow[c] = rk[c];
    }
    // reflector k only changes row k of the R part (its R component is e_k): the rows k' > k broadcast later from
    // rpre are still the original ones
    T[k][k] = tau;
#pragma unroll
    for (int x = 0; x < k; ++x) {
      double acc = 0.0;
#pragma unroll
      for (int l = x; l < k; ++l) acc += T[x][l] * (sc * red[l]);
      T[x][k] = -tau * acc;
    }
  }
  c2_store_rblock(R, ldr, j0, j0, n, true, rrow);
#pragma unroll
  for (int x = 0; x < FT_B; ++x)
    if (lane == x) {
#pragma unroll
      for (int y = 0; y < FT_B; ++y) Tm[x * FT_B + y] = T[x][y];
    }
}

// consumer: apply the reflectors of the panel at rows j0.. (being factored by the other chain warp) to the slab at
// columns c0.. held in registers, one reflector at a time as they are published; updates the R block (j0, c0)
__device__ __forceinline__ void c2_stream(double (&a)[C2_RPL][FT_B], const int j0, const int c0, const int n, double* __restrict__ R,
                                          const int ldr, const double* Vt, const double* taus) {
  const int lane = threadIdx.x & 31;
  double rblk[FT_B], rnew[FT_B];
  c2_load_rblock(R, ldr, j0, c0, n, false, rblk);
#pragma unroll
  for (int c = 0; c < FT_B; ++c) rnew[c] = rblk[c];
#pragma unroll
  for (int k = 0; k < FT_B; ++k) {
    double rk[FT_B];
#pragma unroll
    for (int c = 0; c < FT_B; ++c) rk[c] = __shfl_sync(0xffffffffu, rblk[c], k);
    c2_wait(1 + k);
    double v[C2_RPL];
#pragma unroll
    for (int r = 0; r < C2_RPL; ++r) v[r] = Vt[k * C2_LDV + lane + 32 * r];
    const double tau = taus[k];
    double part[FT_B], red[FT_B];
#pragma unroll
    for (int c = 0; c < FT_B; ++c) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < C2_RPL; ++r) s += v[r] * a[r][c];
      part[c] = s;
    }
    c2_reduce8(part, red);
#pragma unroll
    for (int c = 0; c < FT_B; ++c) {
      const double w = tau * (rk[c] + red[c]);
#pragma unroll
      for (int r = 0; r < C2_RPL; ++r) a[r][c] -= v[r] * w;
      if (lane == k) rnew[c] = rk[c] - w;
    }
  }
  c2_store_rblock(R, ldr, j0, c0, n, false, rnew);
}

// same contract as qr_ft_cta<64> (drop-in once validated): A m x n row-major read-only, R n x n fully overwritten with
// the upper-triangular factor, optionally divided by its max-abs
__device__ void qr_chain2_cta(const double* __restrict__ A, const int m, const int n, const int lda, double* __restrict__ R, const int ldr,
                              const bool normalize, double* smem) {
  constexpr int H = C2_H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n8 = (n + 7) & ~7;
  const int ld = ft_ld(n);
  double* Ablk = smem;
  double* VT0 = Ablk + (size_t)H * ld;
  double* Ws = VT0 + 2 * C2_VT + warp * 72;
  for (int idx = tid; idx < n * n; idx += NT) R[(size_t)(idx / n) * ldr + (idx % n)] = 0.0;
  const int g = lane >> 2, q4 = lane & 3;
  const int npanel = n8 / FT_B;
  const bool chain = (warp == 0 || warp == 4);
  const int urank = warp < 4 ? warp - 1 : warp - 2;  // 0..5 for the updating warps 1,2,3,5,6,7
  double a[C2_RPL][FT_B];  // chain warps: the slab they stream into / factor
  for (int row0 = 0; row0 < m; row0 += H) {
    __syncthreads();
    for (int idx = tid; idx < H * n8; idx += NT) {
      const int i = idx / n8, c = idx % n8;
      const int gi = row0 + i;
      Ablk[(size_t)i * ld + (c ^ ft_sw(i))] = (gi < m && c < n) ? A[(size_t)gi * lda + c] : 0.0;
    }
    __syncthreads();
    if (warp == 0) c2_load_slab(Ablk, ld, 0, a);
    for (int s = 0; s < npanel; ++s) {
      const int producer = (s & 1) ? 4 : 0;
      double* Vs = VT0 + (s & 1) * C2_VT;            // panel s (being produced)
      double* Vp = VT0 + ((s + 1) & 1) * C2_VT;      // panel s-1 (complete)
      const int j0 = s * FT_B;
      if (warp == producer) {
        c2_factor(a, j0, n, R, ldr, Vs, Vs + FT_B * C2_LDV, Vs + FT_B * C2_LDV + FT_B * FT_B, s + 1 < npanel);
      } else if (chain) {
        if (s + 1 < npanel) {
          const int c0 = j0 + FT_B;
          if (s >= 1) {
            // (i) the finished panel s-1 on slab s+1 (rows of panel s-1 in R)
            FtFrags<H> f;
            f.load(Vp, Vp + FT_B * C2_LDV);
            const int rr = j0 - FT_B + g, cc = c0 + 2 * q4;
            const bool ok0 = (rr < n) && (cc < n), ok1 = (rr < n) && (cc + 1 < n);
            double* rp = R + (size_t)rr * ldr + cc;
            const double r0 = ok0 ? rp[0] : 0.0, r1 = ok1 ? rp[1] : 0.0;
            ft_update_slab<H>(Ablk, ld, c0, f, Ws, rp, ok0, ok1, r0, r1);
            __syncwarp();
          }
          c2_load_slab(Ablk, ld, c0, a);                                                       // (ii)
          c2_stream(a, j0, c0, n, R, ldr, Vs, Vs + FT_B * C2_LDV + FT_B * FT_B);                // (iii)
        }
      } else if (s >= 1) {
        // panel s-1 on the slabs p >= s+2
        FtFrags<H> f;
        f.load(Vp, Vp + FT_B * C2_LDV);
        const int rr = j0 - FT_B + g;
        for (int p = s + 2 + urank; p < npanel; p += NW - 2) {
          const int c0 = p * FT_B, cc = c0 + 2 * q4;
          const bool ok0 = (rr < n) && (cc < n), ok1 = (rr < n) && (cc + 1 < n);
          double* rp = R + (size_t)rr * ldr + cc;
          const double r0 = ok0 ? rp[0] : 0.0, r1 = ok1 ? rp[1] : 0.0;
          ft_update_slab<H>(Ablk, ld, c0, f, Ws, rp, ok0, ok1, r0, r1);
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
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

__global__ void __launch_bounds__(NT, 1) k_chain2(const double* A, int m, int n, double* R) {
  extern __shared__ double smem[];
  qr_chain2_cta(A + (size_t)blockIdx.x * m * n, m, n, n, R + (size_t)blockIdx.x * n * n, n, false, smem);
}
__global__ void __launch_bounds__(NT, 1) k_base(const double* A, int m, int n, double* R) {
  extern __shared__ double smem[];
  qr_ft_cta<64>(A + (size_t)blockIdx.x * m * n, m, n, n, R + (size_t)blockIdx.x * n * n, n, false, smem);
}

}  // namespace mpbp

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__);      \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

int main(int argc, char** argv) {
  using namespace mpbp;
  const int m = argc > 1 ? atoi(argv[1]) : 1600, n = argc > 2 ? atoi(argv[2]) : 400, batch = argc > 3 ? atoi(argv[3]) : 148;
  std::vector<double> A((size_t)m * n);
  unsigned long long st = 88172645463325252ull;
  for (auto& x : A) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    x = (double)(st >> 11) / 9007199254740992.0 - 0.5;
  }
  double *dA, *dR;
  CK(cudaMalloc(&dA, sizeof(double) * (size_t)batch * m * n));
  CK(cudaMalloc(&dR, sizeof(double) * (size_t)batch * n * n));
  for (int b = 0; b < batch; ++b) CK(cudaMemcpy(dA + (size_t)b * m * n, A.data(), sizeof(double) * m * n, cudaMemcpyHostToDevice));
  std::vector<double> G((size_t)n * n, 0.0);
  double gmax = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      double s = 0.0;
      for (int r = 0; r < m; ++r) s += A[(size_t)r * n + i] * A[(size_t)r * n + j];
      G[(size_t)i * n + j] = G[(size_t)j * n + i] = s;
      gmax = fmax(gmax, fabs(s));
    }
  const size_t sm2 = c2_smem_doubles(n) * 8, sm1 = ft_smem_doubles<64>(n) * 8;
  CK(cudaFuncSetAttribute(k_chain2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  CK(cudaFuncSetAttribute(k_base, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const double fl = batch * (2.0 * m * n * n - 2.0 / 3.0 * n * n * n);
  for (int which = 0; which < 2; ++which) {
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_base<<<batch, NT, sm1>>>(dA, m, n, dR);
      else k_chain2<<<batch, NT, sm2>>>(dA, m, n, dR);
      cudaEventRecord(e1);
      CK(cudaGetLastError());
      CK(cudaEventSynchronize(e1));
      float t;
      cudaEventElapsedTime(&t, e0, e1);
      best = t < best ? t : best;
    }
    std::vector<double> Rh((size_t)n * n);
    CK(cudaMemcpy(Rh.data(), dR + (size_t)(batch - 1) * n * n, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    double err = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = i; j < n; ++j) {
        double s = 0.0;
        for (int r = 0; r <= i; ++r) s += Rh[(size_t)r * n + i] * Rh[(size_t)r * n + j];
        err = fmax(err, fabs(s - G[(size_t)i * n + j]));
      }
    printf("%s %dx%d b%d: %.2f ms %.2f TF/s  err=%.1e\n", which ? "chain2" : "base  ", m, n, batch, best, fl / best / 1e9, err / gmax);
  }
  return 0;
}
