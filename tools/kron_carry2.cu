// kron_carry2.cu -- PROTOTYPE (not part of the product library, not yet run on a GPU): register-tiled stage 2 of
// k_kron_carry.  The product kernel feeds every FMA of stage 2 with one shared-memory load (4 FMAs per 1 LDG + 4 LDS);
// here a thread owns MB = 4 consecutive m1 values x RB = 4 right-bond columns (16 FMAs per 4 LDG + 4 LDS).
// The harness builds one heavy op of the bench shape (bonds 20, Glauber prefix product ny1 = 5, ny2 = 2 -> nyo = 6),
// replicates it `batch` times, checks the new kernel against k_kron_carry<4> element by element and times both.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/kron_carry2 tools/kron_carry2.cu
//   ./tools/kron_carry2 [batch]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../matrixproductbp.jl_b200/csrc/kernels.cuh"

namespace mpbp {

template <int RB, int MB>
__global__ void __launch_bounds__(NT) k_kron_carry_rt(const OpDesc* ops, int t, int L) {
  extern __shared__ double smem[];
  const OpDesc& op = ops[blockIdx.x];
  const int x = blockIdx.y;
  if (x >= op.q) return;
  const int bl1 = op.a.bonds[t], br1 = op.a.bonds[t + 1], bl2 = op.b.bonds[t], br2 = op.b.bonds[t + 1];
  const int Dl = bl1 * bl2, Dr = br1 * br2;
  const int rn = op.r[t + 1];
  const int rr0 = blockIdx.z * KC_RC;
  if (rr0 >= rn) return;
  const double* A1 = op.a.data + (size_t)t * op.a.stride;
  const double* A2 = op.b.data + (size_t)t * op.b.stride;
  const double* Lm = (t + 1 < L) ? op.Lbuf + (size_t)(t + 1) * op.Lstride : nullptr;
  const double* pyy = op.pyy + (size_t)t * op.pyy_tstride;
  const int ny1 = op.ny1, ny2 = op.ny2, nyo = op.nyo;
  const int nz = bl2 * br1 * ny2;
  double* Lcol = smem;         // RB x Dr
  double* Z = smem + RB * Dr;  // RB x nz
  const int rr1 = min(rr0 + KC_RC, rn);
  for (int rb = rr0; rb < rr1; rb += RB) {
    const int nb = min(RB, rr1 - rb);
    for (int i = threadIdx.x; i < RB * Dr; i += NT) {
      const int u = i / Dr, e = i % Dr;
      Lcol[i] = (u < nb) ? (Lm ? Lm[e + (size_t)Dr * (rb + u)] : 1.0) : 0.0;
    }
    __syncthreads();
    // stage 1 (unchanged): Z[u][m2, n1, y2] = sum_n2 A2[m2,n2,y2,x] L[(n1,n2), rb+u]
    for (int idx = threadIdx.x; idx < nz; idx += NT) {
      const int m2 = idx % bl2, n1 = (idx / bl2) % br1, y2 = idx / (bl2 * br1);
      const double* a2 = A2 + m2 + (size_t)bl2 * br2 * (y2 + ny2 * x);
      double acc[RB];
#pragma unroll
      for (int u = 0; u < RB; ++u) acc[u] = 0.0;
      for (int n2 = 0; n2 < br2; ++n2) {
        const double a = a2[bl2 * n2];
        const double* lc = Lcol + n1 + br1 * n2;
#pragma unroll
        for (int u = 0; u < RB; ++u) acc[u] += a * lc[u * Dr];
      }
#pragma unroll
      for (int u = 0; u < RB; ++u) Z[u * nz + idx] = acc[u];
    }
    __syncthreads();
    // stage 2, register-tiled: a thread owns m1 = MB*mg .. MB*mg+MB-1 for one (m2, y) and the RB columns
    const int nmg = (bl1 + MB - 1) / MB;
    const int no = nmg * bl2 * nyo;
    for (int idx = threadIdx.x; idx < no; idx += NT) {
      const int mg = idx % nmg, m2 = (idx / nmg) % bl2, y = idx / (nmg * bl2);
      const int m1 = mg * MB;
      double acc[MB][RB];
#pragma unroll
      for (int j = 0; j < MB; ++j)
#pragma unroll
        for (int u = 0; u < RB; ++u) acc[j][u] = 0.0;
      for (int y2 = 0; y2 < ny2; ++y2)
        for (int y1 = 0; y1 < ny1; ++y1) {
          const double pv = pyy[y + nyo * (y1 + ny1 * (y2 + ny2 * x))];
          if (pv != 0.0) {
            const double* a1 = A1 + m1 + (size_t)bl1 * br1 * (y1 + ny1 * x);
            const double* z = Z + m2 + bl2 * br1 * y2;
            double s[MB][RB];
#pragma unroll
            for (int j = 0; j < MB; ++j)
#pragma unroll
              for (int u = 0; u < RB; ++u) s[j][u] = 0.0;
            for (int n1 = 0; n1 < br1; ++n1) {
              double av[MB], zv[RB];
#pragma unroll
              for (int j = 0; j < MB; ++j) av[j] = (m1 + j < bl1) ? a1[bl1 * n1 + j] : 0.0;
#pragma unroll
              for (int u = 0; u < RB; ++u) zv[u] = z[bl2 * n1 + u * nz];
#pragma unroll
              for (int j = 0; j < MB; ++j)
#pragma unroll
                for (int u = 0; u < RB; ++u) s[j][u] += av[j] * zv[u];
            }
#pragma unroll
            for (int j = 0; j < MB; ++j)
#pragma unroll
              for (int u = 0; u < RB; ++u) acc[j][u] += pv * s[j][u];
          }
        }
#pragma unroll
      for (int j = 0; j < MB; ++j)
#pragma unroll
        for (int u = 0; u < RB; ++u)
          if (m1 + j < bl1 && u < nb) op.M[(m1 + j + bl1 * m2) + (size_t)Dl * (rb + u + (size_t)rn * (y + nyo * x))] = acc[j][u];
    }
    __syncthreads();
  }
}

}  // namespace mpbp

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) {                                                   \
      printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__);  \
      return 1;                                                                \
    }                                                                          \
  } while (0)

template <class T>
static T* dev_copy(const std::vector<T>& v) {
  T* p = nullptr;
  cudaMalloc((void**)&p, sizeof(T) * v.size());
  cudaMemcpy(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
  return p;
}

int main(int argc, char** argv) {
  using namespace mpbp;
  const int batch = argc > 1 ? atoi(argv[1]) : 32;
  const int L = 3, t = 1, d = 20, q = 2, ny1 = 5, ny2 = 2, nyo = 6, rn = 400;
  const int D = d * d, X = nyo * q;
  unsigned long long st = 88172645463325252ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0 - 0.5; };
  std::vector<double> a((size_t)L * d * d * ny1 * q), b((size_t)L * d * d * ny2 * q), Lb((size_t)L * D * D), pyy((size_t)L * nyo * ny1 * ny2 * q, 0.0);
  for (auto& v : a) v = rnd();
  for (auto& v : b) v = rnd();
  for (auto& v : Lb) v = rnd();
  for (int tt = 0; tt < L; ++tt)
    for (int x = 0; x < q; ++x)
      for (int y2 = 0; y2 < ny2; ++y2)
        for (int y1 = 0; y1 < ny1; ++y1) pyy[(size_t)tt * nyo * ny1 * ny2 * q + (y1 + y2) + nyo * (y1 + ny1 * (y2 + ny2 * x))] = 1.0;
  std::vector<int> bonds = {1, d, d, 1}, r = {1, rn, rn, 1};
  double *da = dev_copy(a), *db = dev_copy(b), *dL = dev_copy(Lb), *dp = dev_copy(pyy);
  int *dbonds = dev_copy(bonds), *dr = dev_copy(r);
  const size_t msz = (size_t)rn * X * D;
  double *dM0, *dM1;
  CK(cudaMalloc(&dM0, sizeof(double) * msz * batch));
  CK(cudaMalloc(&dM1, sizeof(double) * msz * batch));
  CK(cudaMemset(dM0, 0, sizeof(double) * msz * batch));
  CK(cudaMemset(dM1, 0, sizeof(double) * msz * batch));
  std::vector<OpDesc> ops0(batch), ops1(batch);
  for (int k = 0; k < batch; ++k) {
    OpDesc op;
    memset(&op, 0, sizeof op);
    op.a = TTRef{da, dbonds, nullptr, d * d * ny1 * q, ny1 * q};
    op.b = TTRef{db, dbonds, nullptr, d * d * ny2 * q, ny2 * q};
    op.ny1 = ny1; op.ny2 = ny2; op.nyo = nyo; op.q = q;
    op.pyy = dp; op.pyy_tstride = nyo * ny1 * ny2 * q;
    op.r = dr; op.Lbuf = dL; op.Lstride = (long long)D * D;
    op.M = dM0 + msz * k;
    ops0[k] = op;
    op.M = dM1 + msz * k;
    ops1[k] = op;
  }
  OpDesc *d0 = dev_copy(ops0), *d1 = dev_copy(ops1);
  const size_t smem = 4 * ((size_t)D + (size_t)d * d * ny2) * 8;
  CK(cudaFuncSetAttribute(k_kron_carry<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_kron_carry_rt<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(batch, q, (rn + KC_RC - 1) / KC_RC);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const double flops = (double)batch * q * rn * (2.0 * d * d * ny2 * d + 2.0 * D * nyo * 2 /*pairs per y (<=)*/ * d);
  for (int which = 0; which < 2; ++which) {
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_kron_carry<4><<<grid, NT, smem>>>(d0, t, L);
      else k_kron_carry_rt<4, 4><<<grid, NT, smem>>>(d1, t, L);
      cudaEventRecord(e1);
      CK(cudaGetLastError());
      CK(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    printf("%s: %.3f ms  (~%.2f TF/s)\n", which ? "register-tiled" : "product       ", best, flops / best / 1e9);
  }
  std::vector<double> h0(msz), h1(msz);
  CK(cudaMemcpy(h0.data(), dM0 + msz * (batch - 1), sizeof(double) * msz, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h1.data(), dM1 + msz * (batch - 1), sizeof(double) * msz, cudaMemcpyDeviceToHost));
  double err = 0.0, mx = 0.0;
  for (size_t i = 0; i < msz; ++i) {
    err = fmax(err, fabs(h0[i] - h1[i]));
    mx = fmax(mx, fabs(h0[i]));
  }
  printf("max |diff| = %.3e (max |M| = %.3e)\n", err, mx);
  return 0;
}
