// kernels.cuh -- batched FP64 kernels of the MPBP node update (sm_100a).  See DESIGN.md for the math.
//
// Reference code each kernel replaces:
//   k_btilde          compute_prob_ys, first map      src/recursive_bp_factor.jl:108-115
//   k_kron_carry      op(): Kronecker build + right-orthogonalisation carry   :118-127 (+ TensorTrains.compress! sweep 1)
//   k_qr_stage        un-truncated R->L sweep of compress! (triangular factors only)
//   k_kron_proj, k_gemm_m2t, k_qr_small, k_jacobi_project   truncating L->R sweep of compress! (:127)
//   k_finalize        _f_bp_partial :73-87, mpem2 src/mpems.jl:67-94, compress!(..., is_orthogonal=:left),
//                     normalize_eachmatrix!, normalize! / set_msg!   src/recursive_bp_factor.jl:154-158,168-179
//   k_belief          f_bp_partial_i, mpem2, marginalize, normalize!, marginals  :160-163, src/mpbp.jl:237
//   k_pair_belief     pair_belief_as_mpem / pair_belief  src/bp_core.jl:95-109, src/mpbp.jl:202-235
#pragma once
#include "common.cuh"
#include "jacobi.cuh"
#include "qr.cuh"
#include "qr_ft.cuh"
#include "svd_block.cuh"
#include "svd_squared.cuh"

namespace mpbp {

// ------------------------------------------------------------------------------------------------
// B~_k = sum_{x_k} Pxy[y,x_k,x_i] psi[x_i,x_k] mu_{k->i}[m,n,x_k,x_i]
// ------------------------------------------------------------------------------------------------
struct BtJob {
  TTRef msg;  // [m,n,xk,xi]
  TTRef out;  // [m,n,y,xi]
  const double* psi;  // [t][xi + qi*xk]
  const double* pxy;  // [y + ny1*(xk + qk*xi)] per t
  int pxy_tstride;
  int qk, qi, ny1;
};

__global__ void __launch_bounds__(NT) k_btilde(const BtJob* jobs, int L) {
  const BtJob jb = jobs[blockIdx.x];
  const int t = blockIdx.y;
  const int bl = jb.msg.bonds[t], br = jb.msg.bonds[t + 1];
  const double* A = jb.msg.data + (size_t)t * jb.msg.stride;
  double* O = jb.out.data + (size_t)t * jb.out.stride;
  const double* psi = jb.psi + (size_t)t * jb.qi * jb.qk;
  const double* pxy = jb.pxy + (size_t)t * jb.pxy_tstride;
  const int mn = bl * br;
  const int tot = mn * jb.ny1 * jb.qi;
  for (int idx = threadIdx.x; idx < tot; idx += NT) {
    const int e = idx % mn, y = (idx / mn) % jb.ny1, x = idx / (mn * jb.ny1);
    double acc = 0.0;
    for (int xk = 0; xk < jb.qk; ++xk)
      acc += pxy[y + jb.ny1 * (xk + jb.qk * x)] * psi[x + jb.qi * xk] * A[e + mn * (xk + jb.qk * x)];
    O[idx] = acc;
  }
  if (t == 0) {
    for (int i = threadIdx.x; i <= L; i += NT) jb.out.bonds[i] = jb.msg.bonds[i];
    if (threadIdx.x == 0) *jb.out.ls = *jb.msg.ls;
  }
}

// ------------------------------------------------------------------------------------------------
// generic BPFactor (exhaustive trace, src/bp_core.jl:18-57 / :60-93): Kronecker product of the incoming messages of
// all neighbours but `skip` (skip = -1: all of them, dummy-neighbour path), reweighted by psi:
//   out[(m_k)_k, (n_k)_k, y = (x_k)_k, x] = prod_k mu_k[m_k, n_k, x_k, x] * psi_k[x, x_k]      (first k fastest)
// The dense factor table is applied afterwards by k_finalize / k_belief exactly like prob_y_partial / prob_y.
// ------------------------------------------------------------------------------------------------
constexpr int GEN_MAXZ = 8;
struct GenJob {
  TTRef msg[GEN_MAXZ];
  const double* psi[GEN_MAXZ];  // [t][x + q*xk]
  int qk[GEN_MAXZ];
  int nk, skip, q;
  TTRef out;
};
__global__ void __launch_bounds__(NT) k_generic_kron(const GenJob* jobs, int L, int dcap, int* err) {
  const GenJob& jb = jobs[blockIdx.x];
  const int t = blockIdx.y;
  int bl[GEN_MAXZ], br[GEN_MAXZ];
  long long Ml = 1, Mr = 1, Y = 1;
  for (int k = 0; k < jb.nk; ++k) {
    if (k == jb.skip) continue;
    bl[k] = jb.msg[k].bonds[t];
    br[k] = jb.msg[k].bonds[t + 1];
    Ml *= bl[k];
    Mr *= br[k];
    Y *= jb.qk[k];
  }
  if (Ml > dcap || Mr > dcap) {
    // loud error; leave a VALID (bond-1) site behind so that the kernels that follow stay inside their buffers
    if (threadIdx.x == 0) {
      atomicOr(err, ERR_BOND_OVERFLOW);
      jb.out.bonds[t] = 1;
      if (t == L - 1) jb.out.bonds[L] = 1;
      if (t == 0) *jb.out.ls = 0.0;
    }
    double* Oz = jb.out.data + (size_t)t * jb.out.stride;
    for (int idx = threadIdx.x; idx < jb.out.stride; idx += NT) Oz[idx] = 0.0;
    return;
  }
  const int q = jb.q;
  double* O = jb.out.data + (size_t)t * jb.out.stride;
  const long long tot = Ml * Mr * Y * q;
  for (long long idx = threadIdx.x; idx < tot; idx += NT) {
    long long r = idx;
    long long m = r % Ml; r /= Ml;
    long long n = r % Mr; r /= Mr;
    long long y = r % Y;
    const int x = (int)(r / Y);
    double v = 1.0;
    for (int k = 0; k < jb.nk; ++k) {
      if (k == jb.skip) continue;
      const int mk = (int)(m % bl[k]); m /= bl[k];
      const int nk_ = (int)(n % br[k]); n /= br[k];
      const int xk = (int)(y % jb.qk[k]); y /= jb.qk[k];
      const double* A = jb.msg[k].data + (size_t)t * jb.msg[k].stride;
      v *= A[mk + (size_t)bl[k] * (nk_ + (size_t)br[k] * (xk + jb.qk[k] * x))] * jb.psi[k][(size_t)t * q * jb.qk[k] + x + q * xk];
    }
    O[idx] = v;
  }
  if (threadIdx.x == 0) {
    jb.out.bonds[t] = (int)Ml;
    if (t == L - 1) jb.out.bonds[L] = (int)Mr;
    if (t == 0) {
      double ls = 0.0;
      for (int k = 0; k < jb.nk; ++k)
        if (k != jb.skip) ls += *jb.msg[k].ls;
      *jb.out.ls = ls;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// heavy op descriptor
// ------------------------------------------------------------------------------------------------
struct OpDesc {
  TTRef a, b, o;
  int ny1, ny2, nyo, q;
  const double* pyy;  // [y + nyo*(y1 + ny1*(y2 + ny2*x))] per t
  int pyy_tstride;
  int* r;             // [L+1]  r[t] = #cols of the sweep-1 factor L_t ; r[L] = 1
  double* Lbuf;       // L_t at Lbuf + t*Lstride, column-major D_l(t) x r[t]
  long long Lstride;
  double* M;          // tall scratch, row-major (r*X) x D_l
  double* Ms;         // TSQR stack scratch
  double* G;          // [mt, n, y, x]
  double* M2T;        // row-major r x (dX)  == column-major (dX) x r
  double* R2;         // (dX) x (dX)
  double* Pc[2];      // carry [mt, (m1,m2)]
};

__global__ void k_op_setup(const OpDesc* ops, int nops, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nops) return;
  const OpDesc& op = ops[i];
  op.r[L] = 1;
  op.o.bonds[0] = 1;
  *op.o.ls = *op.a.ls + *op.b.ls;
}

constexpr int KC_RC = 8;  // right-bond columns per CTA in k_kron_carry

// M_t[(m1,m2), rr, y, x] = sum_{y1,y2} Pyy sum_{n1,n2} B1[m1,n1,y1,x] B2[m2,n2,y2,x] L_{t+1}[(n1,n2), rr]
// grid (nops, q, ceil(rcap/KC_RC)).  RB right-bond columns are processed together so that every operand load
// (B1/B2 element) feeds RB independent FMAs.  dyn smem: RB * (Dcap + d*d*nymax) doubles.
template <int RB>
__global__ void __launch_bounds__(NT) k_kron_carry(const OpDesc* ops, int t, int L) {
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
  double* Lcol = smem;            // RB x Dr
  double* Z = smem + RB * Dr;     // RB x nz
  const int rr1 = min(rr0 + KC_RC, rn);
  for (int rb = rr0; rb < rr1; rb += RB) {
    const int nb = min(RB, rr1 - rb);
    for (int i = threadIdx.x; i < RB * Dr; i += NT) {
      const int u = i / Dr, e = i % Dr;
      Lcol[i] = (u < nb) ? (Lm ? Lm[e + (size_t)Dr * (rb + u)] : 1.0) : 0.0;
    }
    __syncthreads();
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
    const int no = Dl * nyo;
    for (int idx = threadIdx.x; idx < no; idx += NT) {
      const int c = idx % Dl, y = idx / Dl;
      const int m1 = c % bl1, m2 = c / bl1;
      double acc[RB];
#pragma unroll
      for (int u = 0; u < RB; ++u) acc[u] = 0.0;
      for (int y2 = 0; y2 < ny2; ++y2)
        for (int y1 = 0; y1 < ny1; ++y1) {
          const double pv = pyy[y + nyo * (y1 + ny1 * (y2 + ny2 * x))];
          if (pv != 0.0) {
            const double* a1 = A1 + m1 + (size_t)bl1 * br1 * (y1 + ny1 * x);
            const double* z = Z + m2 + bl2 * br1 * y2;
            double s[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) s[u] = 0.0;
            for (int n1 = 0; n1 < br1; ++n1) {
              const double a = a1[bl1 * n1];
#pragma unroll
              for (int u = 0; u < RB; ++u) s[u] += a * z[bl2 * n1 + u * nz];
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) acc[u] += pv * s[u];
          }
        }
#pragma unroll
      for (int u = 0; u < RB; ++u)
        if (u < nb) op.M[c + (size_t)Dl * (rb + u + (size_t)rn * (y + nyo * x))] = acc[u];
    }
    __syncthreads();
  }
}

// DMMA version of k_kron_carry (same contract, same output).  The two contractions are GEMMs:
//   stage 1 (per y_S)        Z_yS[m_S, (n_F,u)] = sum_{n_S} S[m_S,n_S,y_S,x] * Lc[(n_F,n_S), u]          M = bl_S, K = br_S, N = br_F*RB
//   stage 2 (per output y)   out_y[m_F, (m_S,u)] = sum_{(y_F,y_S)} Pyy * sum_{n_F} F[m_F,n_F,y_F,x] * Z_yS[m_S,n_F,u]   M = bl_F, K = br_F, N = bl_S*RB
// where u indexes the RB right-bond columns of L_{t+1} handled by the CTA and (F, S) = (B1, B2) or (B2, B1): the operand
// with FEWER auxiliary states is contracted first (Z holds ny_S copies in shared memory).  M is padded to a multiple of 8
// and K to a multiple of 4 with zeros (bonds are 1..30); one B fragment feeds every m-tile.  A fragments come straight
// from L1/L2 (a site of an operand train is a few KB).  grid (nops, q, ceil(rcap/RB)).
// dyn smem: RB*DrP + nySmax*RB*brFmax*ZP doubles (see kc_mma_smem_doubles).
__host__ __device__ inline int kc_pad4(int n) { return n + ((12 - (n & 7)) & 7); }  // == 4 (mod 8): conflict-free B fragments
__host__ __device__ inline size_t kc_mma_smem_doubles(int RB, int Dcap, int dcap, int nyS, int nymax) {
  return (size_t)RB * (Dcap + 8) + (size_t)nyS * RB * dcap * kc_pad4(dcap) + (size_t)nymax * nymax * nyS +
         (size_t)(nymax + nyS) * dcap * dcap + 8;  // + the two operand sites (all auxiliary states of this x)
}
template <int RB>
__global__ void __launch_bounds__(NT) k_kron_carry_mma(const OpDesc* ops, int t, int L, int nyS_cap, int ny_cap, double* flops) {
  extern __shared__ double smem[];
  const OpDesc& op = ops[blockIdx.x];
  const int x = blockIdx.y;
  if (x >= op.q) return;
  const int bl1 = op.a.bonds[t], br1 = op.a.bonds[t + 1], bl2 = op.b.bonds[t], br2 = op.b.bonds[t + 1];
  const int Dl = bl1 * bl2, Dr = br1 * br2;
  const int rn = op.r[t + 1];
  const int rr0 = blockIdx.z * RB;
  if (rr0 >= rn) return;
  const int nb = min(RB, rn - rr0);
  const double* A1 = op.a.data + (size_t)t * op.a.stride;
  const double* A2 = op.b.data + (size_t)t * op.b.stride;
  const double* Lm = (t + 1 < L) ? op.Lbuf + (size_t)(t + 1) * op.Lstride : nullptr;
  const double* pyy = op.pyy + (size_t)t * op.pyy_tstride;
  const int ny1 = op.ny1, ny2 = op.ny2, nyo = op.nyo;
  // roles: S is contracted first
  const bool swap = ny1 < ny2;  // true: S = B1, F = B2
  const double* Sd = swap ? A1 : A2;
  const double* Fd = swap ? A2 : A1;
  const int blS = swap ? bl1 : bl2, brS = swap ? br1 : br2, nyS = swap ? ny1 : ny2;
  const int blF = swap ? bl2 : bl1, brF = swap ? br2 : br1, nyF = swap ? ny2 : ny1;
  const int sLF = swap ? br1 : 1, sLS = swap ? 1 : br1;        // L row index = nF*sLF + nS*sLS
  const int sOF = swap ? bl1 : 1, sOS = swap ? 1 : bl1;        // out row index = mF*sOF + mS*sOS
  const int pF = swap ? nyo * ny1 : nyo, pS = swap ? nyo : nyo * ny1;  // pyy index = y + yF*pF + yS*pS + nyo*ny1*ny2*x
  if (nyS > nyS_cap || nyo > ny_cap || nyF > ny_cap) return;  // host guarantees this does not happen (launch sized from the group maxima)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q4 = lane & 3;
  const int DrP = Dr + 8;
  const int ZP = kc_pad4(blS);
  double* Lc = smem;                       // [u][DrP]
  double* Z = smem + (size_t)RB * DrP;     // [(yS*RB + u)*brF + nF][ZP]  (mS fastest)
  double* Ps = Z + (size_t)nyS_cap * RB * brF * ZP;  // Pyy[y + nyo*(yF + nyF*yS)] of this x: the (y_S, y_F) loops test it per pair,
                                                      // from global memory every test would be an exposed L2 round trip
  for (int i = threadIdx.x; i < nyo * nyF * nyS; i += NT) {
    const int y = i % nyo, yF = (i / nyo) % nyF, yS = i / (nyo * nyF);
    Ps[i] = pyy[y + yF * pF + yS * pS + (size_t)nyo * ny1 * ny2 * x];
  }
  // the two operand sites of this x (contiguous: [bl, br, ny] blocks) also live in shared memory: the A fragments of both
  // stages are re-read per n-tile group / per (y_S, y_F) pair, and from L2 every re-read is an exposed round trip
  double* Fs = Ps + (((size_t)nyo * nyF * nyS + 1) & ~(size_t)1);
  double* Ss = Fs + (((size_t)blF * brF * nyF + 1) & ~(size_t)1);
  const double* Fg = Fd + (size_t)blF * brF * nyF * x;
  const double* Sg = Sd + (size_t)blS * brS * nyS * x;
  const int nFs = blF * brF * nyF, nSs = blS * brS * nyS;
  if (flops && blockIdx.z == 0 && threadIdx.x == 0) {
    int npairs = 0;
    for (int yS = 0; yS < nyS; ++yS)
      for (int yF = 0; yF < nyF; ++yF) {
        bool any = false;
        for (int y = 0; y < nyo; ++y) any |= pyy[y + yF * pF + yS * pS + (size_t)nyo * ny1 * ny2 * x] != 0.0;
        npairs += any;
      }
    atomicAdd(flops, 2.0 * rn * ((double)blS * brS * brF * nyS + (double)blF * brF * blS * npairs));
  }
  // stage the RB columns of L_{t+1} (one TMA bulk copy per column, Dr contiguous doubles) and the two operand sites (one
  // bulk copy each) when 16-byte aligned; plain loads otherwise
  __shared__ uint64_t kc_mbar;
  const bool bulkL = Lm != nullptr && (Dr & 1) == 0 && aligned16(Lm) && aligned16(Lc);
  const bool bulkF = (nFs & 1) == 0 && aligned16(Fg) && aligned16(Fs);
  const bool bulkS = (nSs & 1) == 0 && aligned16(Sg) && aligned16(Ss);
  if (bulkL || bulkF || bulkS) {
    if (threadIdx.x == 0) mbar_init(&kc_mbar, 1);
    __syncthreads();
    if (warp == 0) {
      if (lane == 0) mbar_expect_tx(&kc_mbar, (uint32_t)((bulkL ? nb * Dr : 0) + (bulkF ? nFs : 0) + (bulkS ? nSs : 0)) * 8u);
      __syncwarp();
      if (bulkL)
        for (int u = lane; u < nb; u += 32) bulk_g2s(Lc + (size_t)u * DrP, Lm + (size_t)Dr * (rr0 + u), (uint32_t)(Dr * 8), &kc_mbar);
      if (bulkF && lane == 0) bulk_g2s(Fs, Fg, (uint32_t)(nFs * 8), &kc_mbar);
      if (bulkS && lane == 1) bulk_g2s(Ss, Sg, (uint32_t)(nSs * 8), &kc_mbar);
    }
  }
  if (bulkL) {
    for (int i = threadIdx.x; i < nb * (DrP - Dr); i += NT) Lc[(size_t)(i / (DrP - Dr)) * DrP + Dr + i % (DrP - Dr)] = 0.0;
    for (int i = threadIdx.x; i < (RB - nb) * DrP; i += NT) Lc[(size_t)nb * DrP + i] = 0.0;
  } else {
    for (int i = threadIdx.x; i < RB * DrP; i += NT) {
      const int u = i / DrP, e = i % DrP;
      Lc[i] = (u < nb && e < Dr) ? (Lm ? Lm[e + (size_t)Dr * (rr0 + u)] : 1.0) : 0.0;
    }
  }
  if (!bulkF)
    for (int i = threadIdx.x; i < nFs; i += NT) Fs[i] = Fg[i];
  if (!bulkS)
    for (int i = threadIdx.x; i < nSs; i += NT) Ss[i] = Sg[i];
  if (bulkL || bulkF || bulkS) mbar_wait(&kc_mbar, 0);
  __syncthreads();
  // ---------------- stage 1 ----------------
  // A fragments (the operand site, a few KB in L1/L2) are hoisted into registers per auxiliary state: the inner loops
  // then issue one shared-memory B-fragment load per m-tile group of DMMAs.
  const int mtS = (blS + 7) >> 3, ksS = (brS + 3) >> 2;
  const int ncol1 = brF * RB, nt1 = (ncol1 + 7) >> 3;
  for (int yS = 0; yS < nyS; ++yS) {
    const double* Sy = Ss + (size_t)blS * brS * yS;
    double af[4][8];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int mr = 8 * mt + g, kB = 4 * ks + q4;
        af[mt][ks] = (mr < blS && kB < brS) ? Sy[mr + blS * kB] : 0.0;
      }
    for (int nt = warp; nt < nt1; nt += NW) {
      const int cB = 8 * nt + g;                       // B-fragment column of this lane
      const bool cBok = cB < ncol1;
      const int boff = cBok ? (cB / brF) * DrP + (cB % brF) * sLF : 0;  // column (nF, u), nF fastest
      double acc[4][2];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) acc[mt][0] = acc[mt][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        if (ks < ksS) {
          const int kB = 4 * ks + q4;
          const double bf = (cBok && kB < brS) ? Lc[boff + kB * sLS] : 0.0;
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
            if (mt < mtS) dmma884(acc[mt][0], acc[mt][1], af[mt][ks], bf);
        }
      }
      // C tile: rows mS = 8mt+g, columns c0, c0+1
      const int c0 = 8 * nt + 2 * q4;
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        if (mt < mtS) {
          const int mr = 8 * mt + g;
          if (mr < ZP) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = c0 + e;
              if (c < ncol1) {
                const int u = c / brF, nF = c % brF;
                Z[((size_t)(yS * RB + u) * brF + nF) * ZP + mr] = mr < blS ? acc[mt][e] : 0.0;
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // ---------------- stage 2 ----------------
  const int mtF = (blF + 7) >> 3, ksF = (brF + 3) >> 2;
  const int ncol2 = blS * RB, nt2 = (ncol2 + 7) >> 3;
  constexpr int TG = 4;  // n-tiles a warp carries at once (their accumulators persist across the (y_S, y_F) pairs)
  for (int y = 0; y < nyo; ++y) {
    for (int ntb = warp; ntb < nt2; ntb += NW * TG) {
      int zoff[TG];
      bool cok[TG];
#pragma unroll
      for (int j = 0; j < TG; ++j) {
        const int cB = 8 * (ntb + NW * j) + g;
        cok[j] = (ntb + NW * j < nt2) && cB < ncol2;
        zoff[j] = cok[j] ? ((cB / blS) * brF) * ZP + (cB % blS) : 0;  // column (mS, u), mS fastest; + (yS*RB*brF + kB)*ZP
      }
      double acc[TG][4][2];
#pragma unroll
      for (int j = 0; j < TG; ++j)
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) acc[j][mt][0] = acc[j][mt][1] = 0.0;
      for (int yS = 0; yS < nyS; ++yS) {
        const double* zy = Z + (size_t)yS * RB * brF * ZP;
        for (int yF = 0; yF < nyF; ++yF) {
          const double pv = Ps[y + nyo * (yF + nyF * yS)];
          if (pv == 0.0) continue;
          const double* Fy = Fs + (size_t)blF * brF * yF;
          double af[4][8];
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const int mr = 8 * mt + g, kB = 4 * ks + q4;
              af[mt][ks] = (mr < blF && kB < brF) ? pv * Fy[mr + blF * kB] : 0.0;
            }
#pragma unroll
          for (int j = 0; j < TG; ++j) {
            if (ntb + NW * j < nt2) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                if (ks < ksF) {
                  const int kB = 4 * ks + q4;
                  const double bf = (cok[j] && kB < brF) ? zy[zoff[j] + (size_t)kB * ZP] : 0.0;
#pragma unroll
                  for (int mt = 0; mt < 4; ++mt)
                    if (mt < mtF) dmma884(acc[j][mt][0], acc[j][mt][1], af[mt][ks], bf);
                }
              }
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < TG; ++j) {
        const int nt = ntb + NW * j;
        if (nt < nt2) {
          const int c0 = 8 * nt + 2 * q4;
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            if (mt < mtF) {
              const int mF = 8 * mt + g;
              if (mF < blF) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int c = c0 + e;
                  if (c < ncol2) {
                    const int u = c / blS, mS = c % blS;
                    if (u < nb) op.M[(size_t)(mF * sOF + mS * sOS) + (size_t)Dl * (rr0 + u + (size_t)rn * (y + nyo * x))] = acc[j][mt][e];
                  }
                }
              }
            }
          }
        }
      }
    }
  }
}

// Sweep-1 factor of site t with the flat-tree DMMA QR: L_t = R^T of M_t ((rn*X) x Dl), r[t] = min(rows, Dl).
// When a launch holds few matrices (nsplit > 1) tall matrices are split TSQR-style over several CTAs: row chunks
// are factored independently into op.Ms (n x n each) and k_qr_ft_merge factors the stack.
// split_min: matrices with fewer than split_min * n rows stay whole (in a full launch only the tall outliers are split: they
// are the long pole of the launch, the split trades ~15 % more flops on them for a 2x shorter critical path)
__device__ __forceinline__ void ft_split(int m, int n, int nsplit, int split_min, int& nch, int& ch_rows) {
  const int n8 = (n + 7) & ~7;
  if (m < split_min * n8) { nch = 1; ch_rows = m; return; }
  // latency of chunk stage + merge stage ~ m/k + k*n rows  ->  k ~ sqrt(m/n)  (the merge factors a k*n-row stack)
  nch = min(nsplit, (int)(sqrt((double)m / (double)n8) + 0.5));
  if (nch < 2) { nch = 1; ch_rows = m; return; }
  ch_rows = (((m + nch - 1) / nch) + 31) & ~31;
  nch = (m + ch_rows - 1) / ch_rows;
}
__device__ __forceinline__ double qr_flops(double mm, double nn) {
  return mm >= nn ? 2.0 * mm * nn * nn - (2.0 / 3.0) * nn * nn * nn : 2.0 * nn * mm * mm - (2.0 / 3.0) * mm * mm * mm;
}
template <int H>
__global__ void __launch_bounds__(NT, (H == 32 ? 2 : 1)) k_qr_ft(const OpDesc* ops, int t, int nsplit, double* flops, int split_min) {
  extern __shared__ double smem[];
  const OpDesc& op = ops[blockIdx.x];
  const int Dl = op.a.bonds[t] * op.b.bonds[t];
  const int m = op.r[t + 1] * op.nyo * op.q;
  double* Lt = op.Lbuf + (size_t)t * op.Lstride;
  if (m <= Dl) {
    if (blockIdx.y > 0) return;
    // wide case: M_t^T (m x Dl) itself is a valid factor (L = A^T, L L^T = A^T A); no factorisation needed
    __shared__ double redc[NW + 1];
    double mx = 0.0;
    for (int idx = threadIdx.x; idx < m * Dl; idx += NT) mx = fmax(mx, fabs(op.M[idx]));
    mx = block_max(mx, redc);
    const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
    for (int idx = threadIdx.x; idx < m * Dl; idx += NT) Lt[idx] = op.M[idx] * f;
    if (threadIdx.x == 0) op.r[t] = m;
    return;
  }
  int nch, ch_rows;
  ft_split(m, Dl, nsplit, split_min, nch, ch_rows);
  const int ch = blockIdx.y;
  if (ch >= nch) return;
  const int row0 = ch * ch_rows, rows = min(ch_rows, m - row0);
  // roofline numerator = ALGORITHMIC flops of the unsplit matrix (2mn^2 - 2/3 n^3); what the TSQR split executes on
  // top of that (per-chunk triangles + merge) is accumulated separately in flops[16]
  if (flops && threadIdx.x == 0) {
    if (ch == 0) atomicAdd(flops, qr_flops(m, Dl));
    if (nch > 1) atomicAdd(flops + 16, qr_flops(rows, Dl) - (ch == 0 ? qr_flops(m, Dl) : 0.0));
  }
  if (nch == 1) {
    qr_ft_cta<H>(op.M, m, Dl, Dl, Lt, Dl, true, smem);
    if (threadIdx.x == 0) op.r[t] = Dl;
  } else {
    qr_ft_cta<H>(op.M + (size_t)row0 * Dl, rows, Dl, Dl, op.Ms + (size_t)ch * Dl * Dl, Dl, false, smem);
  }
}
template <int H>
__global__ void __launch_bounds__(NT, (H == 32 ? 2 : 1)) k_qr_ft_merge(const OpDesc* ops, int t, int nsplit, double* flops, int tri, int split_min) {
  extern __shared__ double smem[];
  const OpDesc& op = ops[blockIdx.x];
  const int Dl = op.a.bonds[t] * op.b.bonds[t];
  const int m = op.r[t + 1] * op.nyo * op.q;
  if (m <= Dl) return;
  int nch, ch_rows;
  ft_split(m, Dl, nsplit, split_min, nch, ch_rows);
  if (nch == 1) return;
  if (flops && threadIdx.x == 0) atomicAdd(flops + 16, qr_flops((double)nch * Dl, Dl));
  qr_ft_cta<H>(op.Ms, nch * Dl, Dl, Dl, op.Lbuf + (size_t)t * op.Lstride, Dl, true, smem, /*tri_n=*/tri ? Dl : 0);
  if (threadIdx.x == 0) op.r[t] = Dl;
}

// G_t[mt, (n1,n2), y, x] = sum_{y1,y2} Pyy sum_{m1,m2} Pc[mt,(m1,m2)] B1[m1,n1,y1,x] B2[m2,n2,y2,x]
// grid (nops, q, nyo_max); dyn smem dcap*dcap*dcap doubles
__global__ void __launch_bounds__(NT) k_kron_proj(const OpDesc* ops, int t) {
  extern __shared__ double smem[];
  const OpDesc& op = ops[blockIdx.x];
  const int x = blockIdx.y, y = blockIdx.z;
  if (x >= op.q || y >= op.nyo) return;
  const int bl1 = op.a.bonds[t], br1 = op.a.bonds[t + 1], bl2 = op.b.bonds[t], br2 = op.b.bonds[t + 1];
  const int Dr = br1 * br2;
  const int dt = op.o.bonds[t];
  const double* Pc = (t == 0) ? nullptr : op.Pc[(t - 1) & 1];
  const double* A1 = op.a.data + (size_t)t * op.a.stride;
  const double* A2 = op.b.data + (size_t)t * op.b.stride;
  const double* pyy = op.pyy + (size_t)t * op.pyy_tstride;
  const int ny1 = op.ny1, ny2 = op.ny2, nyo = op.nyo;
  double* Y = smem;
  double* Gs = op.G + (size_t)dt * Dr * (y + nyo * x);
  bool first = true;
  for (int y1 = 0; y1 < ny1; ++y1) {
    bool any = false;
    for (int y2 = 0; y2 < ny2; ++y2) any |= (pyy[y + nyo * (y1 + ny1 * (y2 + ny2 * x))] != 0.0);
    if (!any) continue;
    const int nyy = dt * bl2 * br1;
    for (int idx = threadIdx.x; idx < nyy; idx += NT) {
      const int mt = idx % dt, m2 = (idx / dt) % bl2, n1 = idx / (dt * bl2);
      const double* a1 = A1 + (size_t)bl1 * (n1 + br1 * (y1 + ny1 * x));
      double acc = 0.0;
      if (Pc) {
        const double* pc = Pc + mt + (size_t)dt * bl1 * m2;
        for (int m1 = 0; m1 < bl1; ++m1) acc += pc[dt * m1] * a1[m1];
      } else {
        acc = a1[0];  // t == 0: bl1 = bl2 = dt = 1
      }
      Y[idx] = acc;
    }
    __syncthreads();
    for (int y2 = 0; y2 < ny2; ++y2) {
      const double pv = pyy[y + nyo * (y1 + ny1 * (y2 + ny2 * x))];
      if (pv == 0.0) continue;
      const double* a2 = A2 + (size_t)bl2 * br2 * (y2 + ny2 * x);
      const int ng = dt * Dr;
      for (int idx = threadIdx.x; idx < ng; idx += NT) {
        const int mt = idx % dt, n = idx / dt, n1 = n % br1, n2 = n / br1;
        const double* yy = Y + mt + (size_t)dt * bl2 * n1;
        const double* aa = a2 + (size_t)bl2 * n2;
        double s = 0.0;
        for (int m2 = 0; m2 < bl2; ++m2) s += yy[dt * m2] * aa[m2];
        Gs[idx] = first ? pv * s : Gs[idx] + pv * s;
      }
      first = false;
    }
    __syncthreads();
  }
  if (first) {
    const int ng = dt * Dr;
    for (int idx = threadIdx.x; idx < ng; idx += NT) Gs[idx] = 0.0;
  }
}

// M2T[a + dX*rr] = sum_n G[mt + dt*(n + Dr*yx)] * L_{t+1}[n + Dr*rr],  a = mt + dt*yx
// grid (nops, ceil(dXcap/32), ceil(rcap/32)), 32x32 tiles, 4 outputs per thread
__global__ void __launch_bounds__(NT) k_gemm_m2t(const OpDesc* ops, int t) {
  __shared__ double Gs[32][33];
  __shared__ double Ls[32][33];
  const OpDesc& op = ops[blockIdx.x];
  const int br1 = op.a.bonds[t + 1], br2 = op.b.bonds[t + 1];
  const int Dr = br1 * br2;
  const int dt = op.o.bonds[t];
  const int X = op.nyo * op.q;
  const int dX = dt * X;
  const int rn = op.r[t + 1];
  const int a0 = blockIdx.y * 32, r0 = blockIdx.z * 32;
  if (a0 >= dX || r0 >= rn) return;
  const double* Lm = op.Lbuf + (size_t)(t + 1) * op.Lstride;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty 0..7
  double acc[4] = {0, 0, 0, 0};
  for (int n0 = 0; n0 < Dr; n0 += 32) {
    // Gs[aa][nn], Ls[nn][rr]
    for (int k = 0; k < 4; ++k) {
      const int nn = ty + 8 * k;
      const int a = a0 + tx, n = n0 + nn;
      Gs[tx][nn] = (a < dX && n < Dr) ? op.G[(a % dt) + (size_t)dt * (n + (size_t)Dr * (a / dt))] : 0.0;
      const int nl = n0 + tx, rr = r0 + nn;
      Ls[tx][nn] = (nl < Dr && rr < rn) ? Lm[nl + (size_t)Dr * rr] : 0.0;  // Ls[n][rr]
    }
    __syncthreads();
#pragma unroll 8
    for (int nn = 0; nn < 32; ++nn) {
      const double g = Gs[tx][nn];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += g * Ls[nn][ty + 8 * k];
    }
    __syncthreads();
  }
  for (int k = 0; k < 4; ++k) {
    const int a = a0 + tx, rr = r0 + ty + 8 * k;
    if (a < dX && rr < rn) op.M2T[a + (size_t)dX * rr] = acc[k];
  }
}

// if r_{t+1} > dX: R2 = R-factor of M2T (r x dX); Jacobi then runs on R2^T (dX x dX)
__host__ __device__ inline bool svd_direct(int p, int n, int jac_doubles);
template <int H>
__global__ void __launch_bounds__(NT, (H == 32 ? 2 : 1)) k_qr_small(const OpDesc* ops, int t, int jac_doubles) {
  extern __shared__ double smem[];
  const OpDesc& op = ops[blockIdx.x];
  const int dX = op.o.bonds[t] * op.nyo * op.q;
  const int rn = op.r[t + 1];
  if (rn <= dX || !svd_direct(dX, rn, jac_doubles)) return;
  qr_ft_cta<H>(op.M2T, rn, dX, dX, op.R2, dX, true, smem);
}

// ---- truncated SVD of M2 (p x n, p = d~X, n = r_{t+1}) : only the leading left singular vectors are needed ----
// small matrices : one-sided Jacobi on the (QR-reduced) matrix in shared memory ("direct");
// large matrices : un-squared block subspace iteration (svd_block.cuh)
//                    Z = orth(M^T Q),  Y = M Z = Q R,  Ritz values = singular values of the b x b factor R
//                  with b <= 64 columns: DMMA GEMMs, Householder orthonormalisation of the blocks, Jacobi only on R.
//                  Converges like (sigma_{b+1}/sigma_k)^2 per iteration; validated against exact SVDs in the tests.
__host__ __device__ inline bool svd_direct(int p, int n, int jac_doubles) {
  const int c = p < n ? p : n;
  return c <= SUB_BMAX && (long long)p * c <= jac_doubles;
}
// doubles of global scratch svd_left_cta needs for a p x n matrix (subspace path + exact fallback)
__host__ __device__ inline size_t svd_scratch_doubles(int p, int n) {
  const size_t mx = (size_t)(p > n ? p : n) + 8;
  return 3 * SUB_BMAX * mx + (size_t)p * n + SUB_BMAX * SUB_BMAX;
}
__device__ inline void normalize_cols(double* W, int rows, int b, int ld, const double* sig) {
  double smax = 0.0;
  for (int j = 0; j < b; ++j) smax = fmax(smax, sig[j]);
  for (int j = threadIdx.x >> 5; j < b; j += NW) {
    const double f = jacobi_inv_sigma(sig[j], smax);
    for (int k = threadIdx.x & 31; k < rows; k += 32) W[k + (size_t)j * ld] *= f;
  }
  __syncthreads();
}

// Leading left singular vectors of M (p x n column-major at Mcm; when n > p and the direct path applies, R2 must
// hold the Q-less QR factor of M^T as produced by k_qr_small).  On return (all threads): columns of A (p x ceff,
// lda = p, in shared memory) are orthonormal left singular vectors, sig[col] their singular values, order[] sorts
// them descending; nrm2_all = ||M||_F^2 when known (subspace path) else -1.
// smem: [sig 64][order 64 ints][sprev 64][W : jac_doubles].  scratch (global): svd_scratch_doubles(p, n) doubles.
struct SvdLeft {
  double* A;
  int ceff;
  double nrm2_all;
};
__device__ inline SvdLeft svd_left_cta(const double* Mcm, const double* R2, double* scratch, const int p, const int rn, const Trunc tr,
                                       const int dcap, const int jac_doubles, double* smem, int* flagp, int* s_donep, double* red,
                                       int* err, double* stats, const int mode = 0) {
  int& flag = *flagp;
  int& s_done = *s_donep;
  const int c = min(p, rn);
  double* sig = smem;                                   // 64
  int* order = reinterpret_cast<int*>(smem + SUB_BMAX);  // 64 ints
  double* sprev = smem + SUB_BMAX + SUB_BMAX / 2;        // 64
  double* W = sprev + SUB_BMAX;                          // jac_doubles
  double* A;
  int ceff;
  double nrm2_all = -1.0;
  if (svd_direct(p, rn, jac_doubles)) {
    const double* Ag = (rn > p) ? R2 : Mcm;  // column-major p x c, lda = p
    for (int i = threadIdx.x; i < p * c; i += NT) W[i] = Ag[i];
    __syncthreads();
    A = W;
    ceff = c;
    const int sweeps = jacobi_cols(A, p, ceff, p, &flag);
    if (sweeps >= JACOBI_MAX_SWEEPS && threadIdx.x == 0) atomicOr(err, ERR_JACOBI_NOCONV);
    jacobi_sort(A, p, ceff, p, sig, order);
    normalize_cols(A, p, ceff, p, sig);
  } else if (mode == 0 && (long long)max(p, rn) * (min(min(max(SUB_BLOCK, min(SUB_BMAX, 2 * (tr.kind == 1 ? dcap : tr.d) + 8)), c), SUB_BMAX) & ~7) <= jac_doubles) {
    // the full-width block fits shared memory: squared iteration with Jacobi orthonormalisation (svd_squared.cuh)
    const size_t mx8 = (size_t)max(p, rn) + 8;
    int b = 0;
    nrm2_all = svd_subspace_squared(Mcm, scratch + SUB_BMAX * mx8, scratch, scratch + 3 * SUB_BMAX * mx8, p, rn, tr, dcap, jac_doubles, sig,
                                    order, sprev, W, flagp, s_donep, red, err, stats, &b);
    A = W;
    ceff = b;
  } else {
    const int n = rn;
    const double* M = Mcm;  // column-major p x n
    long long tph = clock64();
    auto phase = [&](int which) {
      __syncthreads();
      const long long now = clock64();
      if (stats && threadIdx.x == 0) atomicAdd(stats + 8 + which, (double)(now - tph));
      tph = now;
    };
    const int kwant = tr.kind == 1 ? dcap : tr.d;
    int b = min(min(max(SUB_BLOCK, min(SUB_BMAX, 2 * kwant + 8)), c), SUB_BMAX) & ~7;
    if (b < 8 || b < min(c, kwant)) {
      if (threadIdx.x == 0) atomicOr(err, ERR_BOND_OVERFLOW);  // the block cannot hold the kept bond
      b = max(b, 8);
    }
    const int ldq = blk_ld(p), ldz = blk_ld(n);
    const size_t mx8 = (size_t)max(p, n) + 8;
    double* Zg = scratch;                         // GEMM output n x b (ld = ldz)
    double* Qg = scratch + SUB_BMAX * mx8;        // GEMM output p x b (ld = ldq)
    double* Gblk = scratch + 2 * SUB_BMAX * mx8;  // the block itself when shared memory cannot hold it
    double* Mwork = scratch + 3 * SUB_BMAX * mx8;  // p x n work copy of the exact fallback
    double* Xg = Mwork + (size_t)p * n;            // b x b (inverse triangular factor of the Cholesky-QR route)
    // shared-memory carve-up of W: [block : cap][S : 64*64][scal : 3*64][dsc : 64][Gs : b*b (mode 2)]
    const int fixed = SUB_BMAX * SUB_BMAX + 4 * SUB_BMAX + (mode == 2 ? b * b : 0);
    const int cap = jac_doubles - fixed;
    double* S = W + cap;
    double* scal = S + SUB_BMAX * SUB_BMAX;
    double* dsc = scal + 3 * SUB_BMAX;
    double* Gs = dsc + SUB_BMAX;
    double* blk = ((long long)max(ldq, ldz) * b <= cap) ? W : Gblk;
    // orthonormalise the block: Cholesky-QR2 (mode 2, block in shared memory) with the Householder route as fallback
    auto orth = [&](const int rows, const int ld, const bool want_R) {
      bool done = false;
      if (mode == 2 && blk == W) done = cholqr2_orth(blk, rows, b, ld, want_R, S, Gs, Xg, dsc, Gblk, &flag);
      if (!done) hh_orth(blk, rows, b, ld, want_R ? S : nullptr, scal);
      if (stats && threadIdx.x == 0 && !done) atomicAdd(stats + 5, 1.0);
    };
    // ---- start block: the b largest-norm columns of M ----
    double* nrm = (n + n / 2 + 2 <= SUB_BMAX * SUB_BMAX) ? S : Zg;  // n column norms + n ints
    int* sel = reinterpret_cast<int*>(nrm + n);
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
    for (int idx = threadIdx.x; idx < p * b; idx += NT) blk[(idx % p) + (size_t)ldq * (idx / p)] = M[(idx % p) + (size_t)p * order[idx / p]];
    for (int j = threadIdx.x; j < SUB_BMAX; j += NT) sprev[j] = 0.0;
    __syncthreads();
    phase(1);
    orth(p, ldq, false);
    phase(2);
    int sw = 0;
    int extra = -1;
    __shared__ int s_extra;
    double dprev = 0.0;
    const int kchk = min(b, kwant);
    int nit = 0;
    for (int it = 0; it < SUB_MAXIT; ++it) {
      blk_gemm_dmma(M, (long long)p, 1LL, n, p, blk, ldq, b, Zg, ldz);  // Z = M^T Q   (n x b)
      __syncthreads();
      for (int idx = threadIdx.x; idx < n * b; idx += NT) { const size_t o = (idx % n) + (size_t)ldz * (idx / n); blk[o] = Zg[o]; }
      phase(3);
      orth(n, ldz, false);
      phase(4);
      blk_gemm_dmma(M, 1LL, (long long)p, p, n, blk, ldz, b, Qg, ldq);  // Y = M Z   (p x b)
      __syncthreads();
      for (int idx = threadIdx.x; idx < p * b; idx += NT) { const size_t o = (idx % p) + (size_t)ldq * (idx / p); blk[o] = Qg[o]; }
      phase(3);
      orth(p, ldq, true);  // Y = Q R, R -> S (b x b)
      phase(4);
      ++nit;
      if (extra > 1) {
        // blind iterations after convergence was detected (their count came from the contraction rate): no Ritz step
        if (threadIdx.x == 0) s_done = 0;
        --extra;
        __syncthreads();
        continue;
      }
      sw = max(sw, jacobi_cols(S, b, b, b, &flag));  // R V = U_r Sigma: the columns of S become U_r Sigma
      jacobi_sort(S, b, b, b, sig, order);
      if (threadIdx.x == 0) {
        // Ritz VALUES converge like angle^2: once they are stationary to 1e-13 sigma_1 the subspace angle is still up to
        // ~sqrt(1e-13).  The contraction of the last step (value error ratio r = rho_angle^2) says how many more
        // iterations bring the angle to 1e-11 (1 for decaying spectra, many for flat ones, where it matters).
        const double s1 = sig[order[0]];
        double dm = 0.0;
        for (int i = 0; i < kchk; ++i) {
          const double s = sig[order[i]];
          dm = fmax(dm, fabs(s - sprev[i]));
          sprev[i] = s;
        }
        dm = s1 > 0.0 ? dm / s1 : 0.0;
        int ex = extra;
        if (ex < 0 && dm <= 1e-13) {
          const double r = fmin(0.9, fmax(1e-8, dprev > 0.0 ? dm / dprev : 1e-8));
          const double ang = sqrt(fmax(dm, 1e-18));
          ex = (int)fmin(30.0, fmax(1.0, ceil(log(1e-11 / ang) / (0.5 * log(r)))));
        } else if (ex > 0) ex--;
        dprev = dm;
        s_extra = ex;
        s_done = (ex == 0);
      }
      __syncthreads();
      extra = s_extra;  // every thread follows the same schedule
      phase(5);
      if (s_done) break;
    }
    const bool converged = s_done;
    if (converged) {
      // Ritz vectors: U = Q U_r  (p x b), sorted by decreasing singular value
      for (int idx = threadIdx.x; idx < b * b; idx += NT) {
        const int l = idx % b, j = idx / b;
        const int cj = order[j];
        Zg[idx] = S[l + (size_t)b * cj] * jacobi_inv_sigma(sig[cj], sig[order[0]]);
      }
      for (int j = threadIdx.x; j < b; j += NT) sprev[j] = sig[order[j]];
      __syncthreads();
      blk_gemm_dmma(blk, 1LL, (long long)ldq, p, b, Zg, b, b, Qg, p);  // U = Q U_r  (p x b, compact ld = p)
      __syncthreads();
    } else {
      // The block iteration hit its cap (clustered singular values across the block edge).  Exact fallback, rare
      // and slow on purpose: one-sided Jacobi on ALL n columns of a work copy of M in global memory, then the b leading
      // columns become the block.  Counted in stats[4].
      if (threadIdx.x == 0)
        printf("[mpbp] subspace SVD not converged after %d iterations (p=%d n=%d b=%d, s1=%.3e s_k=%.3e): exact Jacobi fallback\n",
               nit, p, n, b, sig[order[0]], sig[order[kchk - 1]]);
      __syncthreads();
      for (int idx = threadIdx.x; idx < p * n; idx += NT) Mwork[idx] = M[idx];
      __syncthreads();
      sw = max(sw, jacobi_cols(Mwork, p, n, p, &flag));
      double* sall = Zg;                              // n doubles
      int* oall = reinterpret_cast<int*>(Zg + n);      // n ints
      jacobi_sort(Mwork, p, n, p, sall, oall);
      for (int j = threadIdx.x; j < b; j += NT) sprev[j] = sall[oall[j]];
      __syncthreads();
      const double smax = sprev[0];
      for (int idx = threadIdx.x; idx < p * b; idx += NT)
        Qg[idx] = Mwork[(idx % p) + (size_t)p * oall[idx / p]] * jacobi_inv_sigma(sprev[idx / p], smax);
      __syncthreads();
    }
    // compact copy into shared memory when it fits (lda = p), singular values sorted: order = identity
    A = Qg;
    if ((long long)p * b <= jac_doubles) {
      for (int idx = threadIdx.x; idx < p * b; idx += NT) W[idx] = Qg[idx];
      A = W;
    }
    for (int j = threadIdx.x; j < b; j += NT) {
      sig[j] = sprev[j];
      order[j] = j;
    }
    __syncthreads();
    phase(0);
    if (stats && threadIdx.x == 0) {
      atomicAdd(stats + 0, 1.0);
      atomicAdd(stats + 1, (double)nit);
      atomicAdd(stats + 2, (double)b);
      atomicAdd(stats + 3, (double)sw);
      if (!converged) atomicAdd(stats + 4, 1.0);  // hit SUB_MAXIT: the exact Jacobi fallback above produced the result
      // executed flops of the iteration: two p x n x b GEMMs and two Householder orthonormalisations (4 rows b^2 each)
      atomicAdd(stats + 6, (double)nit * (4.0 * p * n * b + 4.0 * (p + n) * b * b) + 2.0 * p * b * b);
    }
    if (sw >= JACOBI_MAX_SWEEPS && threadIdx.x == 0) atomicOr(err, ERR_JACOBI_NOCONV);
    ceff = b;
  }
  SvdLeft out;
  out.A = A;
  out.ceff = ceff;
  out.nrm2_all = nrm2_all;
  return out;
}

// truncated SVD (left vectors) of M2 (dX x r), output site, new carry Pc_t = U^T G_t (rescaled).
// dyn smem: [sig 64][order 64 ints][sprev 64][W : jac_doubles]
__global__ void __launch_bounds__(NT) k_jacobi_project(const OpDesc* ops, int t, Trunc tr, int dcap, int jac_doubles,
                                                       int* err, double* stats, int svd_mode) {
  extern __shared__ double smem[];
  __shared__ int flag;
  __shared__ int s_keep;
  __shared__ int s_done;
  __shared__ double red[NW + 1];
  const OpDesc& op = ops[blockIdx.x];
  const int br1 = op.a.bonds[t + 1], br2 = op.b.bonds[t + 1];
  const int Dr = br1 * br2;
  const int dt = op.o.bonds[t];
  const int X = op.nyo * op.q;
  const int p = dt * X;
  const int rn = op.r[t + 1];
  const int c = min(p, rn);
  double* sig = smem;
  int* order = reinterpret_cast<int*>(smem + SUB_BMAX);
  // global scratch of the subspace path: the tall sweep-1 matrix op.M is dead during sweep 2 (D^2 X doubles >= what is needed
  // whenever the subspace path can occur, i.e. min(p, n) > 64)
  const SvdLeft sv = svd_left_cta(op.M2T, op.R2, op.M, p, rn, tr, dcap, jac_doubles, smem, &flag, &s_done, red, err, stats, svd_mode);
  double* A = sv.A;
  const int ceff = sv.ceff;
  const double nrm2_all = sv.nrm2_all;
  if (threadIdx.x == 0) {
    // truncation policy on the sorted singular values (c = number of singular values of the reference's SVD)
    double nrm2 = 0.0;
    for (int i = 0; i < ceff; ++i) nrm2 += sig[i] * sig[i];
    if (nrm2_all >= 0.0) nrm2 = nrm2_all;  // subspace path: Frobenius norm of the whole matrix
    int k = c;
    if (tr.kind == 1 || tr.kind == 2) {
      // values below 2e-15 sigma_1 are rounding noise (zero columns of U): TruncThresh(0.0) keeps the numerical rank
      const double lim = fmax(tr.eps * sqrt(nrm2), 2.0 * JACOBI_ZERO * sig[order[0]]);
      int last = 0;
      for (int i = 0; i < ceff; ++i)
        if (sig[order[i]] > lim) last = i + 1;
      if (last == ceff && ceff < c && tr.kind == 1) atomicOr(err, ERR_BOND_OVERFLOW);  // more values above eps than the block holds
      k = last > 0 ? last : 1;
    }
    if (tr.kind == 0 || tr.kind == 2) k = min(k, tr.d);
    if (k > dcap) {
      atomicOr(err, ERR_BOND_OVERFLOW);
      k = dcap;
    }
    if (!(nrm2 == nrm2)) atomicOr(err, ERR_NAN);
    s_keep = k;
    op.o.bonds[t + 1] = k;
  }
  __syncthreads();
  const int keep = s_keep;
  const int kuse = min(keep, ceff);  // columns kuse..keep-1 of U are zero
  // output site A_t[mt, kk, yx] = U[mt + dt*yx, kk]
  double* O = op.o.data + (size_t)t * op.o.stride;
  for (int idx = threadIdx.x; idx < dt * keep * X; idx += NT) {
    const int mt = idx % dt, kk = (idx / dt) % keep, yx = idx / (dt * keep);
    O[idx] = kk < kuse ? A[(mt + dt * yx) + (size_t)order[kk] * p] : 0.0;
  }
  // carry Pc_t[kk + keep*n] = sum_a U[a,kk] G[a; n]   (a = mt + dt*yx): a (Dr x keep) = G^T (Dr x p) U (p x keep) GEMM on the
  // tensor pipe.  K runs over (yx, mt in steps of 4) so that no index division is needed; rows of U beyond kuse are zero.
  double* Pn = op.Pc[t & 1];
  double mx = 0.0;
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q4 = lane & 3;
    const int ntk = (keep + 7) >> 3;  // <= 4 (dcap <= 30)
    for (int mtile = warp; mtile < (Dr + 7) / 8; mtile += NW) {
      const int n = 8 * mtile + g;
      const bool nok = n < Dr;
      double acc[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = 0.0;
      const double* ub[4];
      bool uok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = 8 * j + g;
        uok[j] = j < ntk && kk < kuse;
        ub[j] = A + (size_t)order[uok[j] ? kk : 0] * p;
      }
      for (int yx = 0; yx < X; ++yx) {
        const double* gp = op.G + (size_t)dt * (n + (size_t)Dr * yx);
        for (int m0 = 0; m0 < dt; m0 += 4) {
          const int mt = m0 + q4;
          const bool kok = mt < dt;
          const double af = (nok && kok) ? gp[mt] : 0.0;
          const int a = mt + dt * yx;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < ntk) dmma884(acc[j][0], acc[j][1], af, (uok[j] && kok) ? ub[j][a] : 0.0);
        }
      }
      if (nok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < ntk) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int kk = 8 * j + 2 * q4 + e;
              if (kk < keep) {
                Pn[kk + (size_t)keep * n] = acc[j][e];
                mx = fmax(mx, fabs(acc[j][e]));
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();
  mx = block_max(mx, red);
  if (mx > 0.0 && isfinite(mx)) {
    const double f = 1.0 / mx;
    for (int idx = threadIdx.x; idx < keep * Dr; idx += NT) Pn[idx] *= f;
    if (threadIdx.x == 0) *op.o.ls += log(mx);
  }
}

// last site: A_L[mt, 0, yx] = G_L (Dr = 1), max-abs rescaled
__global__ void __launch_bounds__(NT) k_op_last(const OpDesc* ops, int t) {
  __shared__ double red[NW + 1];
  const OpDesc& op = ops[blockIdx.x];
  const int dt = op.o.bonds[t];
  const int X = op.nyo * op.q;
  double* O = op.o.data + (size_t)t * op.o.stride;
  double mx = 0.0;
  for (int idx = threadIdx.x; idx < dt * X; idx += NT) mx = fmax(mx, fabs(op.G[idx]));
  mx = block_max(mx, red);
  const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
  for (int idx = threadIdx.x; idx < dt * X; idx += NT) O[idx] = op.G[idx] * f;
  if (threadIdx.x == 0) {
    op.o.bonds[t + 1] = 1;
    if (f != 1.0) *op.o.ls += log(mx);
  }
}

// ------------------------------------------------------------------------------------------------
// outgoing message: MPEM3 build + MPEM3->MPEM2 + truncating R->L sweep + normalisation, one CTA per edge
// ------------------------------------------------------------------------------------------------
struct FinJob {
  TTRef c;    // C_j [m,n,y,x]
  TTRef out;  // message slot [m,n,x,xj]
  int nyc, q, qj;
  const double* W;  // [x' + q*(x + q*(xj + qj*y))] per t
  int w_tstride;
  const double* phi;   // [t][x]
  double* logz_out;
  // scratch (global)
  double* Rbuf;  // R index t at Rbuf + t*rstride : row-major k_t x (bl_t*q)
  int rstride;
  int* kdim;     // [L+1]
  double* Bt;    // [m,n,x,xj,x'] current site
  double* S;     // sweep-A matrix / sweep-B N^T
  double* H;     // [m,x,xj,k']
  double* R2;    // (d*q*qj)^2 : QR pre-reduction of a wide sweep-B matrix
  double* Pr[2];
};

__device__ inline void fin_build_B(const FinJob& jb, int t, int L, double* Bt) {
  const int bl = jb.c.bonds[t], br = jb.c.bonds[t + 1];
  const int q = jb.q, qj = jb.qj, ny = jb.nyc;
  const double* C = jb.c.data + (size_t)t * jb.c.stride;
  const double* phi = jb.phi + (size_t)t * q;
  const int mn = bl * br;
  const int tot = mn * q * qj * q;
  if (t < L - 1) {
    const double* W = jb.W + (size_t)t * jb.w_tstride;
    for (int idx = threadIdx.x; idx < tot; idx += NT) {
      const int e = idx % mn, x = (idx / mn) % q, xj = (idx / (mn * q)) % qj, xn = idx / (mn * q * qj);
      double acc = 0.0;
      for (int y = 0; y < ny; ++y) acc += W[xn + q * (x + q * (xj + qj * y))] * C[e + mn * (y + ny * x)];
      Bt[idx] = acc * phi[x];
    }
  } else {
    for (int idx = threadIdx.x; idx < tot; idx += NT) {
      const int e = idx % mn, x = (idx / mn) % q;
      double acc = 0.0;
      for (int y = 0; y < ny; ++y) acc += C[e + mn * (y + ny * x)];
      Bt[idx] = acc * phi[x];
    }
  }
  __syncthreads();
}

// dyn smem: qr scratch (vrows) followed by jacobi cache (jac_doubles)
__global__ void __launch_bounds__(NT) k_finalize(const FinJob* jobs, int L, Trunc tr, int dcap, int vrows,
                                                 int jac_doubles, int* err) {
  extern __shared__ double smem[];
  __shared__ int flag;
  __shared__ int s_keep;
  __shared__ double red[NW + 1];
  __shared__ double s_ls;
  const FinJob& jb = jobs[blockIdx.x];
  const int q = jb.q, qj = jb.qj;
  double* qrs = smem;
  double* jsm = smem + qr_shared_doubles(vrows);
  if (threadIdx.x == 0) {
    jb.kdim[0] = 1;
    s_ls = *jb.c.ls;
  }
  __syncthreads();
  // ---------------- sweep A (L->R): triangular factors of the left-orthonormalisation ----------------
  for (int t = 0; t < L - 1; ++t) {
    const int bl = jb.c.bonds[t], br = jb.c.bonds[t + 1];
    fin_build_B(jb, t, L, jb.Bt);
    const int kprev = jb.kdim[t];
    const int nrows = kprev * q * qj, ncols = br * q;
    const double* Rp = jb.Rbuf + (size_t)t * jb.rstride;  // kprev x (bl*q) row-major (unused for t==0)
    for (int idx = threadIdx.x; idx < nrows * ncols; idx += NT) {
      const int col = idx % ncols, row = idx / ncols;
      const int n = col % br, xn = col / br;
      const int mt = row % kprev, x = (row / kprev) % q, xj = row / (kprev * q);
      const double* b = jb.Bt + (size_t)bl * (n + br * (x + q * (xj + qj * xn)));
      double acc = 0.0;
      if (t == 0) acc = b[0];
      else {
        const double* rp = Rp + (size_t)mt * (bl * q) + bl * x;
        for (int m = 0; m < bl; ++m) acc += rp[m] * b[m];
      }
      jb.S[idx] = acc;
    }
    __syncthreads();
    double* Rn = jb.Rbuf + (size_t)(t + 1) * jb.rstride;
    qr_r_cta(jb.S, nrows, ncols, ncols, Rn, ncols, true, qrs, vrows);
    if (threadIdx.x == 0) jb.kdim[t + 1] = min(nrows, ncols);
    __syncthreads();
  }
  // ---------------- sweep B (R->L): truncating SVDs ----------------
  if (threadIdx.x == 0) {
    jb.out.bonds[L] = 1;
    jb.out.bonds[0] = 1;
  }
  __syncthreads();
  for (int t = L - 1; t >= 0; --t) {
    const int bl = jb.c.bonds[t], br = jb.c.bonds[t + 1];
    fin_build_B(jb, t, L, jb.Bt);
    const int kn = jb.out.bonds[t + 1];  // k~_{t+1}
    const double* Pr = (t == L - 1) ? nullptr : jb.Pr[(t + 1) & 1];  // [(n + br*x') + br*q*k']
    // H[m + bl*(x + q*(xj + qj*k'))]
    const int nh = bl * q * qj * kn;
    for (int idx = threadIdx.x; idx < nh; idx += NT) {
      const int m = idx % bl, x = (idx / bl) % q, xj = (idx / (bl * q)) % qj, kp = idx / (bl * q * qj);
      double acc = 0.0;
      if (!Pr) acc = jb.Bt[m + (size_t)bl * (0 + br * (x + q * (xj + qj * 0)))];
      else {
        const double* pr = Pr + (size_t)br * q * kp;
        for (int xn = 0; xn < q; ++xn) {
          const double* b = jb.Bt + m + (size_t)bl * br * (x + q * (xj + qj * xn));
          for (int n = 0; n < br; ++n) acc += b[bl * n] * pr[n + br * xn];
        }
      }
      jb.H[idx] = acc;
    }
    __syncthreads();
    double* O = jb.out.data + (size_t)t * jb.out.stride;
    if (t == 0) {
      // first site: [1, k', x, xj] = H[0,x,xj,k'], rescaled
      double mx = 0.0;
      for (int idx = threadIdx.x; idx < nh; idx += NT) mx = fmax(mx, fabs(jb.H[idx]));
      mx = block_max(mx, red);
      const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
      for (int idx = threadIdx.x; idx < nh; idx += NT) {
        const int x = idx % q, xj = (idx / q) % qj, kp = idx / (q * qj);
        O[kp + kn * (x + q * xj)] = jb.H[idx] * f;
      }
      if (threadIdx.x == 0 && f != 1.0) s_ls += log(mx);
      __syncthreads();
      break;
    }
    // N^T column-major: rows rr = k' + kn*(x + q*xj), cols mt
    const int kprev = jb.kdim[t];
    const int p = kn * q * qj;
    int c = kprev;
    const double* Rp = jb.Rbuf + (size_t)t * jb.rstride;
    const bool wide = c > p;  // more columns than rows: reduce with a Q-less QR first (right vectors of N = those of R)
    const int cp = max(c, p);
    double* jcache = jsm + cp + (cp + 1) / 2 + 1;
    double* A = (!wide && (long long)p * c <= jac_doubles) ? jcache : jb.S;
    double* sig = jsm;
    int* order = reinterpret_cast<int*>(jsm + cp);
    for (int idx = threadIdx.x; idx < p * c; idx += NT) {
      const int rr = idx % p, mt = idx / p;
      const int kp = rr % kn, x = (rr / kn) % q, xj = rr / (kn * q);
      const double* rp = Rp + (size_t)mt * (bl * q) + bl * x;
      const double* h = jb.H + (size_t)bl * (x + q * (xj + qj * kp));
      double acc = 0.0;
      for (int m = 0; m < bl; ++m) acc += rp[m] * h[m];
      A[idx] = acc;
    }
    __syncthreads();
    if (wide) {
      // A is N row-major (c x p); R (p x p, row-major) == R^T column-major
      qr_r_cta(A, c, p, p, jb.R2, p, false, qrs, vrows);
      A = jb.R2;
      c = p;
      if ((long long)p * c <= jac_doubles) {
        for (int idx = threadIdx.x; idx < p * c; idx += NT) jcache[idx] = A[idx];
        A = jcache;
      }
      __syncthreads();
    }
    const int sweeps = jacobi_cols(A, p, c, p, &flag);
    if (sweeps >= JACOBI_MAX_SWEEPS && threadIdx.x == 0) atomicOr(err, ERR_JACOBI_NOCONV);
    jacobi_sort(A, p, c, p, sig, order);
    if (threadIdx.x == 0) {
      const int ns = min(p, c);
      double nrm2 = 0.0;
      for (int i = 0; i < ns; ++i) nrm2 += sig[order[i]] * sig[order[i]];
      int k = ns;
      if (tr.kind == 1 || tr.kind == 2) {
        const double lim = fmax(tr.eps * sqrt(nrm2), 2.0 * JACOBI_ZERO * sig[order[0]]);
        int last = 0;
        for (int i = 0; i < ns; ++i)
          if (sig[order[i]] > lim) last = i + 1;
        k = last > 0 ? last : 1;
      }
      if (tr.kind == 0 || tr.kind == 2) k = min(k, tr.d);
      if (k > dcap) {
        atomicOr(err, ERR_BOND_OVERFLOW);
        k = dcap;
      }
      if (!(nrm2 == nrm2)) atomicOr(err, ERR_NAN);
      s_keep = k;
      jb.out.bonds[t] = k;
    }
    __syncthreads();
    const int keep = s_keep;
    for (int kk = threadIdx.x >> 5; kk < keep; kk += NW) {
      const int col = order[kk];
      const double s = sig[col];
      const double f = jacobi_inv_sigma(s, sig[order[0]]);
      for (int k = threadIdx.x & 31; k < p; k += 32) A[k + (size_t)col * p] *= f;
    }
    __syncthreads();
    // message site: O[kk + keep*rr] = V^T[kk, rr]
    for (int idx = threadIdx.x; idx < keep * p; idx += NT) {
      const int kk = idx % keep, rr = idx / keep;
      O[idx] = A[rr + (size_t)order[kk] * p];
    }
    // Pr_t[(m + bl*a) + bl*q*kk] = sum_{xj,k'} H[m,a,xj,k'] V^T[kk,(k',a,xj)]
    double* Pn = jb.Pr[t & 1];
    double mx = 0.0;
    for (int idx = threadIdx.x; idx < bl * q * keep; idx += NT) {
      const int m = idx % bl, a = (idx / bl) % q, kk = idx / (bl * q);
      const double* v = A + (size_t)order[kk] * p;
      double acc = 0.0;
      for (int xj = 0; xj < qj; ++xj)
        for (int kp = 0; kp < kn; ++kp)
          acc += jb.H[m + (size_t)bl * (a + q * (xj + qj * kp))] * v[kp + kn * (a + q * xj)];
      Pn[idx] = acc;
      mx = fmax(mx, fabs(acc));
    }
    mx = block_max(mx, red);
    if (mx > 0.0 && isfinite(mx)) {
      const double f = 1.0 / mx;
      for (int idx = threadIdx.x; idx < bl * q * keep; idx += NT) Pn[idx] *= f;
      if (threadIdx.x == 0) s_ls += log(mx);
    }
    __syncthreads();
  }
  // ---------------- normalisation: Z = sum_x prod_t A_t ----------------
  // l (1 x k) <- l * sum_{x,xj} A_t[:,:,x,xj]; vectors in jsm
  {
    double* l0 = jsm;
    double* l1 = jsm + dcap + 1;
    if (threadIdx.x == 0) l0[0] = 1.0;
    __syncthreads();
    double logZ = 0.0;
    for (int t = 0; t < L; ++t) {
      const int bl = jb.out.bonds[t], br = jb.out.bonds[t + 1];
      const double* O = jb.out.data + (size_t)t * jb.out.stride;
      const int P = q * qj;
      double mx = 0.0;
      for (int n = threadIdx.x; n < br; n += NT) {
        double acc = 0.0;
        for (int pp = 0; pp < P; ++pp)
          for (int m = 0; m < bl; ++m) acc += l0[m] * O[m + (size_t)bl * (n + br * pp)];
        l1[n] = acc;
        mx = fmax(mx, fabs(acc));
      }
      mx = block_max(mx, red);
      const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
      for (int n = threadIdx.x; n < br; n += NT) l1[n] *= f;
      if (f != 1.0) logZ += log(mx);
      __syncthreads();
      double* tmp = l0;
      l0 = l1;
      l1 = tmp;
    }
    if (threadIdx.x == 0) {
      const double z = l0[0];
      logZ += log(fabs(z));
      *jb.logz_out = s_ls + logZ;
      *jb.out.ls = -logZ;
      if (!(logZ == logZ) || !(z > 0.0)) atomicOr(err, ERR_NAN);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// damping (set_msg!, src/recursive_bp_factor.jl:168-179):  mu <- normalize!(compress!(mu_new + c * mu_old; svd_trunc)),
// c = damp/(1-damp), both operands normalised.  The sum is the block-diagonal train of bond a+b (TensorTrains._compose);
// compress! = un-truncated R->L sweep (triangular factors only) + truncating L->R sweep, same scheme as the op.
// One CTA per edge; `b` and `out` may be the same slot (in-place on the old message).
// ------------------------------------------------------------------------------------------------
struct DampJob {
  TTRef a, b, out;
  int q, qj;
  double coef;
  double* Lbuf;  // L_t at Lbuf + t*lstride, column-major bl_t x r_t
  int lstride;
  int* r;        // [L+1]
  double* S;     // (r*P) x bl row-major scratch
  double* G;     // [mt, n, p]
  double* M2;    // column-major (dt*P) x r
  double* R2;    // QR pre-reduction when r > dt*P
  double* Pc[2];
};
constexpr int DAMP_MAXL = 192;  // static shared memory must stay below the 2 KB head-room left for it

__global__ void __launch_bounds__(NT) k_damp(const DampJob* jobs, int L, Trunc tr, int dcap, int vrows, int jac_doubles, int* err) {
  extern __shared__ double smem[];
  __shared__ int flag, s_keep;
  __shared__ double red[NW + 1];
  __shared__ double s_ls;
  __shared__ int ba[DAMP_MAXL + 1], bb[DAMP_MAXL + 1];
  const DampJob& jb = jobs[blockIdx.x];
  const int P = jb.q * jb.qj;
  double* qrs = smem;
  double* jsm = smem + qr_shared_doubles(vrows);
  for (int i = threadIdx.x; i <= L; i += NT) {
    ba[i] = jb.a.bonds[i];
    bb[i] = jb.b.bonds[i];
  }
  __syncthreads();
  const double fa = exp(*jb.a.ls), fb = jb.coef * exp(*jb.b.ls);
  // element (m, n, p) of site t of the sum train
  auto site = [&](int t, int m, int n, int p) -> double {
    const int al = ba[t], ar = ba[t + 1], bl = bb[t], br = bb[t + 1];
    const double* A = jb.a.data + (size_t)t * jb.a.stride;
    const double* B = jb.b.data + (size_t)t * jb.b.stride;
    const bool first = (t == 0), last = (t == L - 1);
    if (first && last) return fa * A[p] + fb * B[p];
    if (first) return n < ar ? fa * A[(size_t)al * (n + ar * p)] : fb * B[(size_t)bl * ((n - ar) + br * p)];
    if (last) return m < al ? A[m + (size_t)al * ar * p] : B[(m - al) + (size_t)bl * br * p];
    if (m < al && n < ar) return A[m + (size_t)al * (n + ar * p)];
    if (m >= al && n >= ar) return B[(m - al) + (size_t)bl * ((n - ar) + br * p)];
    return 0.0;
  };
  auto sbl = [&](int t) { return t == 0 ? 1 : ba[t] + bb[t]; };
  auto sbr = [&](int t) { return t == L - 1 ? 1 : ba[t + 1] + bb[t + 1]; };
  if (threadIdx.x == 0) {
    jb.r[L] = 1;
    s_ls = 0.0;
  }
  __syncthreads();
  // ---------------- sweep 1 (R->L): L_t with  K_t ... K_L = L_t * (isometry) ----------------
  for (int t = L - 1; t >= 1; --t) {
    const int bl = sbl(t), br = sbr(t), rn = jb.r[t + 1];
    const double* Ln = jb.Lbuf + (size_t)(t + 1) * jb.lstride;  // br x rn (unused for t = L-1)
    const int rows = rn * P;
    for (int idx = threadIdx.x; idx < rows * bl; idx += NT) {
      const int m = idx % bl, row = idx / bl, rr = row % rn, p = row / rn;
      double acc = 0.0;
      if (t == L - 1) acc = site(t, m, 0, p);
      else
        for (int n = 0; n < br; ++n) acc += site(t, m, n, p) * Ln[n + (size_t)br * rr];
      jb.S[idx] = acc;
    }
    __syncthreads();
    double* Lt = jb.Lbuf + (size_t)t * jb.lstride;
    qr_r_cta(jb.S, rows, bl, bl, Lt, bl, true, qrs, vrows);
    if (threadIdx.x == 0) jb.r[t] = min(rows, bl);
    __syncthreads();
  }
  // ---------------- sweep 2 (L->R): truncating ----------------
  if (threadIdx.x == 0) jb.out.bonds[0] = 1;
  __syncthreads();
  int dt = 1;
  for (int t = 0; t < L; ++t) {
    const int bl = sbl(t), br = sbr(t);
    const double* Pc = (t == 0) ? nullptr : jb.Pc[(t - 1) & 1];  // dt x bl
    for (int idx = threadIdx.x; idx < dt * br * P; idx += NT) {
      const int mt = idx % dt, n = (idx / dt) % br, p = idx / (dt * br);
      double acc = 0.0;
      if (!Pc) acc = site(t, 0, n, p);
      else
        for (int m = 0; m < bl; ++m) acc += Pc[mt + (size_t)dt * m] * site(t, m, n, p);
      jb.G[idx] = acc;
    }
    __syncthreads();
    double* O = jb.out.data + (size_t)t * jb.out.stride;
    if (t == L - 1) {
      double mx = 0.0;
      for (int idx = threadIdx.x; idx < dt * P; idx += NT) mx = fmax(mx, fabs(jb.G[idx]));
      mx = block_max(mx, red);
      const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
      for (int idx = threadIdx.x; idx < dt * P; idx += NT) O[idx] = jb.G[idx] * f;
      if (threadIdx.x == 0) {
        jb.out.bonds[L] = 1;
        if (f != 1.0) s_ls += log(mx);
      }
      __syncthreads();
      break;
    }
    const int rn = jb.r[t + 1];
    const double* Ln = jb.Lbuf + (size_t)(t + 1) * jb.lstride;  // br x rn
    const int p_rows = dt * P;
    for (int idx = threadIdx.x; idx < p_rows * rn; idx += NT) {
      const int a = idx % p_rows, rr = idx / p_rows, mt = a % dt, p = a / dt;
      double acc = 0.0;
      for (int n = 0; n < br; ++n) acc += jb.G[mt + (size_t)dt * (n + (size_t)br * p)] * Ln[n + (size_t)br * rr];
      jb.M2[idx] = acc;
    }
    __syncthreads();
    int c = rn;
    double* A = jb.M2;
    if (c > p_rows) {
      qr_r_cta(jb.M2, c, p_rows, p_rows, jb.R2, p_rows, false, qrs, vrows);
      A = jb.R2;
      c = p_rows;
    }
    const int cp = max(c, p_rows);
    double* sig = jsm;
    int* order = reinterpret_cast<int*>(jsm + cp);
    double* cache = jsm + cp + (cp + 1) / 2 + 1;
    if ((long long)p_rows * c <= jac_doubles) {
      for (int idx = threadIdx.x; idx < p_rows * c; idx += NT) cache[idx] = A[idx];
      A = cache;
      __syncthreads();
    }
    const int sweeps = jacobi_cols(A, p_rows, c, p_rows, &flag);
    if (sweeps >= JACOBI_MAX_SWEEPS && threadIdx.x == 0) atomicOr(err, ERR_JACOBI_NOCONV);
    jacobi_sort(A, p_rows, c, p_rows, sig, order);
    if (threadIdx.x == 0) {
      double nrm2 = 0.0;
      for (int i = 0; i < c; ++i) nrm2 += sig[i] * sig[i];
      int kk = c;
      if (tr.kind == 1 || tr.kind == 2) {
        const double lim = fmax(tr.eps * sqrt(nrm2), 2.0 * JACOBI_ZERO * sig[order[0]]);
        int last = 0;
        for (int i = 0; i < c; ++i)
          if (sig[order[i]] > lim) last = i + 1;
        kk = last > 0 ? last : 1;
      }
      if (tr.kind == 0 || tr.kind == 2) kk = min(kk, tr.d);
      if (kk > dcap) {
        atomicOr(err, ERR_BOND_OVERFLOW);
        kk = dcap;
      }
      if (!(nrm2 == nrm2)) atomicOr(err, ERR_NAN);
      s_keep = kk;
      jb.out.bonds[t + 1] = kk;
    }
    __syncthreads();
    const int keep = s_keep;
    normalize_cols(A, p_rows, c, p_rows, sig);
    for (int idx = threadIdx.x; idx < dt * keep * P; idx += NT) {
      const int mt = idx % dt, kk = (idx / dt) % keep, p = idx / (dt * keep);
      O[idx] = A[(mt + dt * p) + (size_t)order[kk] * p_rows];
    }
    double* Pn = jb.Pc[t & 1];
    double mx = 0.0;
    for (int idx = threadIdx.x; idx < keep * br; idx += NT) {
      const int kk = idx % keep, n = idx / keep;
      const double* u = A + (size_t)order[kk] * p_rows;
      double acc = 0.0;
      for (int p = 0; p < P; ++p)
        for (int mt = 0; mt < dt; ++mt) acc += u[mt + dt * p] * jb.G[mt + (size_t)dt * (n + (size_t)br * p)];
      Pn[idx] = acc;
      mx = fmax(mx, fabs(acc));
    }
    mx = block_max(mx, red);
    if (mx > 0.0 && isfinite(mx)) {
      const double f = 1.0 / mx;
      for (int idx = threadIdx.x; idx < keep * br; idx += NT) Pn[idx] *= f;
      if (threadIdx.x == 0) s_ls += log(mx);
    }
    __syncthreads();
    dt = keep;
  }
  // ---------------- normalize! ----------------
  {
    double* l0 = jsm;
    double* l1 = jsm + dcap + 1;
    if (threadIdx.x == 0) l0[0] = 1.0;
    __syncthreads();
    double logZ = 0.0;
    for (int t = 0; t < L; ++t) {
      const int bl = jb.out.bonds[t], br = jb.out.bonds[t + 1];
      const double* O = jb.out.data + (size_t)t * jb.out.stride;
      double mx = 0.0;
      for (int n = threadIdx.x; n < br; n += NT) {
        double acc = 0.0;
        for (int pp = 0; pp < P; ++pp)
          for (int m = 0; m < bl; ++m) acc += l0[m] * O[m + (size_t)bl * (n + br * pp)];
        l1[n] = acc;
        mx = fmax(mx, fabs(acc));
      }
      mx = block_max(mx, red);
      const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
      for (int n = threadIdx.x; n < br; n += NT) l1[n] *= f;
      if (f != 1.0) logZ += log(mx);
      __syncthreads();
      double* tmp = l0;
      l0 = l1;
      l1 = tmp;
    }
    if (threadIdx.x == 0) {
      const double z = l0[0];
      logZ += log(fabs(z));
      *jb.out.ls = -logZ;
      if (!(logZ == logZ) || !(z > 0.0)) atomicOr(err, ERR_NAN);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// belief of node i from `full`: marginals + log z_i, one CTA per node
// ------------------------------------------------------------------------------------------------
struct BelJob {
  TTRef full;  // [m,n,y,x]
  int ny, q;
  const double* Wd;  // [x' + q*(x + q*y)] per t
  int w_tstride;
  const double* phi;
  double* marg;   // [t][x] output (stride q)
  double* logz;   // output scalar
  double* bw;     // scratch [L][dcap*q]
  double* Bt;     // scratch [m,n,x,x']
  // two-time marginals (option "twovar"; nullptr = off)
  double* tv;     // output [t][u][x_t + q*x_u], (t,u) stride q2cap, zero unless t < u <= t + maxdist
  double* Btall;  // scratch [L-1][dcap*dcap*q*q]
  double* fwall;  // scratch [L][dcap*q]
};

__global__ void __launch_bounds__(NT) k_belief(const BelJob* jobs, int L, int dcap, int* err) {
  __shared__ double red[NW + 1];
  extern __shared__ double smem[];  // fw0, fw1 : dcap*q each
  const BelJob& jb = jobs[blockIdx.x];
  const int q = jb.q, ny = jb.ny;
  const int bwst = dcap * q;
  double logZ = 0.0;
  // backward pass
  for (int t = L - 1; t >= 0; --t) {
    const int bl = jb.full.bonds[t], br = jb.full.bonds[t + 1];
    const double* C = jb.full.data + (size_t)t * jb.full.stride;
    const double* phi = jb.phi + (size_t)t * q;
    double* bw = jb.bw + (size_t)t * bwst;
    const int mn = bl * br;
    double mx = 0.0;
    if (t == L - 1) {
      for (int idx = threadIdx.x; idx < bl * q; idx += NT) {
        const int m = idx % bl, x = idx / bl;
        double acc = 0.0;
        for (int y = 0; y < ny; ++y) acc += C[m + mn * (y + ny * x)];
        acc *= phi[x];
        bw[idx] = acc;
        mx = fmax(mx, fabs(acc));
      }
    } else {
      const double* W = jb.Wd + (size_t)t * jb.w_tstride;
      const double* bn = jb.bw + (size_t)(t + 1) * bwst;  // [n + br*x']
      // Bt[m,n,x,x'] then contract
      for (int idx = threadIdx.x; idx < mn * q * q; idx += NT) {
        const int e = idx % mn, x = (idx / mn) % q, xn = idx / (mn * q);
        double acc = 0.0;
        for (int y = 0; y < ny; ++y) acc += W[xn + q * (x + q * y)] * C[e + mn * (y + ny * x)];
        jb.Bt[idx] = acc * phi[x];
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < bl * q; idx += NT) {
        const int m = idx % bl, x = idx / bl;
        double acc = 0.0;
        for (int xn = 0; xn < q; ++xn)
          for (int n = 0; n < br; ++n) acc += jb.Bt[m + bl * (n + br * (x + q * xn))] * bn[n + br * xn];
        bw[idx] = acc;
        mx = fmax(mx, fabs(acc));
      }
    }
    mx = block_max(mx, red);
    const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
    for (int idx = threadIdx.x; idx < bl * q; idx += NT) bw[idx] *= f;
    if (f != 1.0) logZ += log(mx);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double z = 0.0;
    for (int x = 0; x < q; ++x) z += jb.bw[x];  // bl = 1 at t = 0
    const double lz = *jb.full.ls + logZ + log(z);
    *jb.logz = lz;
    if (!(lz == lz)) atomicOr(err, ERR_NAN);
  }
  // forward pass + marginals
  double* fw0 = smem;
  double* fw1 = smem + bwst;
  for (int t = 0; t < L; ++t) {
    const int bl = jb.full.bonds[t], br = jb.full.bonds[t + 1];
    const double* bw = jb.bw + (size_t)t * bwst;
    if (threadIdx.x < q) {
      const int x = threadIdx.x;
      double acc = 0.0;
      if (t == 0) acc = bw[x];
      else
        for (int m = 0; m < bl; ++m) acc += fw0[m + bl * x] * bw[m + bl * x];
      red[x] = acc;  // q <= NW+1 assumed (q <= 8)
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int x = 0; x < q; ++x) s += red[x];
      for (int x = 0; x < q; ++x) jb.marg[(size_t)t * q + x] = red[x] / s;
    }
    __syncthreads();
    if (t == L - 1) break;
    // fw_t[n,x'] = sum_{m,x} fw_{t-1}[m,x] B_t[m,n,x,x']
    const double* C = jb.full.data + (size_t)t * jb.full.stride;
    const double* phi = jb.phi + (size_t)t * q;
    const double* W = jb.Wd + (size_t)t * jb.w_tstride;
    const int mn = bl * br;
    for (int idx = threadIdx.x; idx < mn * q * q; idx += NT) {
      const int e = idx % mn, x = (idx / mn) % q, xn = idx / (mn * q);
      double acc = 0.0;
      for (int y = 0; y < ny; ++y) acc += W[xn + q * (x + q * y)] * C[e + mn * (y + ny * x)];
      jb.Bt[idx] = acc * phi[x];
    }
    __syncthreads();
    double mx = 0.0;
    for (int idx = threadIdx.x; idx < br * q; idx += NT) {
      const int n = idx % br, xn = idx / br;
      double acc = 0.0;
      for (int x = 0; x < q; ++x)
        for (int m = 0; m < bl; ++m)
          acc += (t == 0 ? 1.0 : fw0[m + bl * x]) * jb.Bt[m + bl * (n + br * (x + q * xn))];
      fw1[idx] = acc;
      mx = fmax(mx, fabs(acc));
    }
    mx = block_max(mx, red);
    const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
    for (int idx = threadIdx.x; idx < br * q; idx += NT) fw1[idx] *= f;
    __syncthreads();
    double* tmp = fw0;
    fw0 = fw1;
    fw1 = tmp;
  }
}

// ------------------------------------------------------------------------------------------------
// marginals of a (compressed) MPEM2 summed over its second variable: bp.b[i] = marginalize(mu) on the generic path, where
// the reference compresses the dummy-neighbour message before marginalising (src/mpbp.jl:145-154, src/mpems.jl:27-29).
// p_t[x] ~ l_{t-1} (sum_xj A_t[:,:,x,xj]) r_{t+1}.  One CTA per job; rv: global scratch (L+1)*dcap doubles.
// ------------------------------------------------------------------------------------------------
struct MargJob {
  TTRef tt;      // [m,n,x,xj]
  int q, qj;
  double* marg;  // [t][x]
  double* rv;
};
__global__ void __launch_bounds__(NT) k_msg_marginals(const MargJob* jobs, int L, int dcap) {
  __shared__ double red[NW + 1];
  __shared__ double pm[8];
  extern __shared__ double smem[];  // l0, l1 : dcap each
  const MargJob& jb = jobs[blockIdx.x];
  const int q = jb.q, P = jb.q * jb.qj;
  const int* bonds = jb.tt.bonds;
  // right vectors r_t (length bonds[t]) = (sum_p A_t) r_{t+1}, rescaled
  if (threadIdx.x == 0) jb.rv[(size_t)L * dcap] = 1.0;
  __syncthreads();
  for (int t = L - 1; t >= 0; --t) {
    const int bl = bonds[t], br = bonds[t + 1];
    const double* A = jb.tt.data + (size_t)t * jb.tt.stride;
    const double* rn = jb.rv + (size_t)(t + 1) * dcap;
    double* rt = jb.rv + (size_t)t * dcap;
    double mx = 0.0;
    for (int m = threadIdx.x; m < bl; m += NT) {
      double acc = 0.0;
      for (int pp = 0; pp < P; ++pp)
        for (int n = 0; n < br; ++n) acc += A[m + (size_t)bl * (n + br * pp)] * rn[n];
      rt[m] = acc;
      mx = fmax(mx, fabs(acc));
    }
    mx = block_max(mx, red);
    const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
    for (int m = threadIdx.x; m < bl; m += NT) rt[m] *= f;
    __syncthreads();
  }
  double* l0 = smem;
  double* l1 = smem + dcap;
  if (threadIdx.x == 0) l0[0] = 1.0;
  __syncthreads();
  for (int t = 0; t < L; ++t) {
    const int bl = bonds[t], br = bonds[t + 1];
    const double* A = jb.tt.data + (size_t)t * jb.tt.stride;
    const double* rn = jb.rv + (size_t)(t + 1) * dcap;
    if (threadIdx.x < q) {
      const int x = threadIdx.x;
      double acc = 0.0;
      for (int xj = 0; xj < jb.qj; ++xj)
        for (int n = 0; n < br; ++n) {
          double s = 0.0;
          for (int m = 0; m < bl; ++m) s += l0[m] * A[m + (size_t)bl * (n + br * (x + q * xj))];
          acc += s * rn[n];
        }
      pm[x] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int x = 0; x < q; ++x) s += pm[x];
      for (int x = 0; x < q; ++x) jb.marg[(size_t)t * q + x] = pm[x] / s;
    }
    double mx = 0.0;
    for (int n = threadIdx.x; n < br; n += NT) {
      double acc = 0.0;
      for (int pp = 0; pp < P; ++pp)
        for (int m = 0; m < bl; ++m) acc += l0[m] * A[m + (size_t)bl * (n + br * pp)];
      l1[n] = acc;
      mx = fmax(mx, fabs(acc));
    }
    mx = block_max(mx, red);
    const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
    for (int n = threadIdx.x; n < br; n += NT) l1[n] *= f;
    __syncthreads();
    double* tmp = l0;
    l0 = l1;
    l1 = tmp;
  }
}

// ------------------------------------------------------------------------------------------------
// two-time marginals of the belief, b_i(x^t, x^u) for t < u <= t + maxdist: what TensorTrains.twovar_marginals
// returns for bp.b[i] (beliefs_tu / autocorrelations / autocovariances, src/mpbp.jl:239-255,289-296).  Same transfer
// formulation as k_belief (the belief MPEM is never built): with B_s[m,n,x,x'] the site tensors of
// f_bp_partial_i(full), F_t[m,x] the forward and G_u[n,y] the backward vectors (jb.bw, left there by k_belief),
//   p_tu[x0,y] ~ sum_n ( [F_t delta_{x0,x}] B_t ... B_{u-1} )[x0; n,y] G_u[n,y].
// One CTA per node, launched right after k_belief on the same jobs.  dyn smem: 2 * q*dcap*q doubles.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_twovar(const BelJob* jobs, int L, int dcap, int maxdist, int q2cap) {
  __shared__ double red[NW + 1];
  __shared__ double pbuf[64];
  extern __shared__ double smem[];
  const BelJob& jb = jobs[blockIdx.x];
  if (!jb.tv) return;
  const int q = jb.q, ny = jb.ny;
  const int bst = dcap * dcap * q * q, fst = dcap * q;
  const int* bonds = jb.full.bonds;
  for (int idx = threadIdx.x; idx < L * L * q2cap; idx += NT) jb.tv[idx] = 0.0;
  // site tensors
  for (int t = 0; t < L - 1; ++t) {
    const int bl = bonds[t], br = bonds[t + 1], mn = bl * br;
    const double* C = jb.full.data + (size_t)t * jb.full.stride;
    const double* phi = jb.phi + (size_t)t * q;
    const double* W = jb.Wd + (size_t)t * jb.w_tstride;
    double* B = jb.Btall + (size_t)t * bst;
    for (int idx = threadIdx.x; idx < mn * q * q; idx += NT) {
      const int e = idx % mn, x = (idx / mn) % q, xn = idx / (mn * q);
      double acc = 0.0;
      for (int y = 0; y < ny; ++y) acc += W[xn + q * (x + q * y)] * C[e + mn * (y + ny * x)];
      B[idx] = acc * phi[x];
    }
  }
  for (int idx = threadIdx.x; idx < q; idx += NT) jb.fwall[idx] = 1.0;  // bonds[0] = 1
  __syncthreads();
  // forward vectors F_{t+1}[n,x'] = sum_{m,x} F_t[m,x] B_t[m,n,x,x'], each rescaled by its max-abs
  for (int t = 0; t < L - 1; ++t) {
    const int bl = bonds[t], br = bonds[t + 1];
    const double* F = jb.fwall + (size_t)t * fst;
    double* Fn = jb.fwall + (size_t)(t + 1) * fst;
    const double* B = jb.Btall + (size_t)t * bst;
    double mx = 0.0;
    for (int idx = threadIdx.x; idx < br * q; idx += NT) {
      const int n = idx % br, xn = idx / br;
      double acc = 0.0;
      for (int x = 0; x < q; ++x)
        for (int m = 0; m < bl; ++m) acc += F[m + bl * x] * B[m + bl * (n + br * (x + q * xn))];
      Fn[idx] = acc;
      mx = fmax(mx, fabs(acc));
    }
    mx = block_max(mx, red);
    const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
    for (int idx = threadIdx.x; idx < br * q; idx += NT) Fn[idx] *= f;
    __syncthreads();
  }
  double* M0 = smem;
  double* M1 = smem + (size_t)q * fst;
  for (int t = 0; t < L - 1; ++t) {
    const int bt = bonds[t];
    const double* F = jb.fwall + (size_t)t * fst;
    for (int idx = threadIdx.x; idx < q * bt * q; idx += NT) {
      const int x0 = idx % q, r = idx / q, x = r / bt;
      M0[idx] = (x0 == x) ? F[r] : 0.0;
    }
    __syncthreads();
    const int umax = min(L - 1, t + maxdist);
    for (int u = t + 1; u <= umax; ++u) {
      const int bl = bonds[u - 1], br = bonds[u];
      const double* B = jb.Btall + (size_t)(u - 1) * bst;
      const double* G = jb.bw + (size_t)u * fst;  // [n + br*y], includes site u
      double mx = 0.0;
      for (int idx = threadIdx.x; idx < q * br * q; idx += NT) {
        const int x0 = idx % q, r = idx / q, n = r % br, y = r / br;
        double acc = 0.0;
        for (int x = 0; x < q; ++x)
          for (int m = 0; m < bl; ++m) acc += M0[x0 + q * (m + bl * x)] * B[m + bl * (n + br * (x + q * y))];
        M1[idx] = acc;
        mx = fmax(mx, fabs(acc));
      }
      mx = block_max(mx, red);
      if (threadIdx.x < q * q) {
        const int x0 = threadIdx.x % q, y = threadIdx.x / q;
        double acc = 0.0;
        for (int n = 0; n < br; ++n) acc += M1[x0 + q * (n + br * y)] * G[n + br * y];
        pbuf[threadIdx.x] = acc;
      }
      __syncthreads();
      if (threadIdx.x < q * q) {
        double sum = 0.0;
        for (int k = 0; k < q * q; ++k) sum += pbuf[k];
        jb.tv[((size_t)t * L + u) * q2cap + threadIdx.x] = pbuf[threadIdx.x] / sum;
      }
      const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
      for (int idx = threadIdx.x; idx < q * br * q; idx += NT) M1[idx] *= f;
      __syncthreads();
      double* tmp = M0;
      M0 = M1;
      M1 = tmp;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// pair beliefs: one CTA per directed edge.  env matrices [a + da*b]
// ------------------------------------------------------------------------------------------------
struct PairJob {
  TTRef A;  // mu_e   [a,a',xs,xd]
  TTRef B;  // mu_rev [b,b',xd,xs]
  const double* psi;  // [t][xs + qs*xd]
  int qs, qd;
  double* out;   // [t][xs + qs*xd]
  double* logz;  // scalar
  double* Renv;  // scratch [L+1][dcap*dcap]
};

__global__ void __launch_bounds__(NT) k_pair_belief(const PairJob* jobs, int L, int dcap) {
  extern __shared__ double smem[];  // tmp [dcap*dcap], Lenv0, Lenv1
  __shared__ double red[NW + 1];
  __shared__ double pm[64];
  const PairJob& jb = jobs[blockIdx.x];
  const int qs = jb.qs, qd = jb.qd;
  const int est = dcap * dcap;
  double* tmp = smem;
  double* L0 = smem + est;
  double* L1 = smem + 2 * est;
  double logZ = 0.0;
  if (threadIdx.x == 0) jb.Renv[(size_t)L * est] = 1.0;
  __syncthreads();
  for (int t = L - 1; t >= 0; --t) {
    const int al = jb.A.bonds[t], ar = jb.A.bonds[t + 1], bl = jb.B.bonds[t], br = jb.B.bonds[t + 1];
    const double* At = jb.A.data + (size_t)t * jb.A.stride;
    const double* Bt = jb.B.data + (size_t)t * jb.B.stride;
    const double* psi = jb.psi + (size_t)t * qs * qd;
    const double* Rn = jb.Renv + (size_t)(t + 1) * est;  // [a' + ar*b']
    double* Rt = jb.Renv + (size_t)t * est;              // [a + al*b]
    for (int idx = threadIdx.x; idx < al * bl; idx += NT) Rt[idx] = 0.0;
    __syncthreads();
    for (int xs = 0; xs < qs; ++xs)
      for (int xd = 0; xd < qd; ++xd) {
        const double ps = psi[xs + qs * xd];
        if (ps == 0.0) continue;
        const double* Ax = At + (size_t)al * ar * (xs + qs * xd);
        const double* Bx = Bt + (size_t)bl * br * (xd + qd * xs);
        // tmp[a + al*b'] = sum_a' Ax[a,a'] Rn[a',b']
        for (int idx = threadIdx.x; idx < al * br; idx += NT) {
          const int a = idx % al, b2 = idx / al;
          double acc = 0.0;
          for (int a2 = 0; a2 < ar; ++a2) acc += Ax[a + al * a2] * Rn[a2 + ar * b2];
          tmp[idx] = acc;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < al * bl; idx += NT) {
          const int a = idx % al, b = idx / al;
          double acc = 0.0;
          for (int b2 = 0; b2 < br; ++b2) acc += tmp[a + al * b2] * Bx[b + bl * b2];
          Rt[idx] += ps * acc;
        }
        __syncthreads();
      }
    double mx = 0.0;
    for (int idx = threadIdx.x; idx < al * bl; idx += NT) mx = fmax(mx, fabs(Rt[idx]));
    mx = block_max(mx, red);
    const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
    for (int idx = threadIdx.x; idx < al * bl; idx += NT) Rt[idx] *= f;
    if (f != 1.0) logZ += log(mx);
    __syncthreads();
  }
  if (threadIdx.x == 0) *jb.logz = *jb.A.ls + *jb.B.ls + logZ + log(jb.Renv[0]);
  // forward: Lenv [a + al*b], marginals
  if (threadIdx.x == 0) L0[0] = 1.0;
  __syncthreads();
  for (int t = 0; t < L; ++t) {
    const int al = jb.A.bonds[t], ar = jb.A.bonds[t + 1], bl = jb.B.bonds[t], br = jb.B.bonds[t + 1];
    const double* At = jb.A.data + (size_t)t * jb.A.stride;
    const double* Bt = jb.B.data + (size_t)t * jb.B.stride;
    const double* psi = jb.psi + (size_t)t * qs * qd;
    const double* Rn = jb.Renv + (size_t)(t + 1) * est;
    for (int idx = threadIdx.x; idx < ar * br; idx += NT) L1[idx] = 0.0;
    __syncthreads();
    for (int xs = 0; xs < qs; ++xs)
      for (int xd = 0; xd < qd; ++xd) {
        const double ps = psi[xs + qs * xd];
        const double* Ax = At + (size_t)al * ar * (xs + qs * xd);
        const double* Bx = Bt + (size_t)bl * br * (xd + qd * xs);
        // tmp[a' + ar*b] = sum_a Ax[a,a'] L0[a,b]
        for (int idx = threadIdx.x; idx < ar * bl; idx += NT) {
          const int a2 = idx % ar, b = idx / ar;
          double acc = 0.0;
          for (int a = 0; a < al; ++a) acc += Ax[a + al * a2] * L0[a + al * b];
          tmp[idx] = acc;
        }
        __syncthreads();
        // new[a',b'] = sum_b tmp[a',b] Bx[b,b'] ; marginal weight = sum new .* Rn
        double part = 0.0;
        for (int idx = threadIdx.x; idx < ar * br; idx += NT) {
          const int a2 = idx % ar, b2 = idx / ar;
          double acc = 0.0;
          for (int b = 0; b < bl; ++b) acc += tmp[a2 + ar * b] * Bx[b + bl * b2];
          acc *= ps;
          L1[idx] += acc;
          part += acc * Rn[idx];
        }
        part = block_sum1(part, red);
        if (threadIdx.x == 0) pm[xs + qs * xd] = part;
        __syncthreads();
      }
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int i = 0; i < qs * qd; ++i) s += pm[i];
      for (int i = 0; i < qs * qd; ++i) jb.out[(size_t)t * qs * qd + i] = pm[i] / s;
    }
    double mx = 0.0;
    for (int idx = threadIdx.x; idx < ar * br; idx += NT) mx = fmax(mx, fabs(L1[idx]));
    mx = block_max(mx, red);
    const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
    for (int idx = threadIdx.x; idx < ar * br; idx += NT) L1[idx] *= f;
    __syncthreads();
    double* sw = L0;
    L0 = L1;
    L1 = sw;
  }
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
// alternate marginals p(x_i^t, x_j^{t+1}) per directed edge i->j, t = 0..T-1 (src/mpbp.jl:270-280: the (t, t+1)
// two-time marginal of the pair-belief MPEM summed over x_j^t and x_i^{t+1}).  One CTA per directed edge, d x d
// environments as in k_pair_belief; with Lenv_t, Renv_{t+2} the environments around sites t, t+1:
//   P_xs[a',b'] = sum_xd psi_t[xs,xd]      (A_t[xs,xd]^T  Lenv_t      B_t[xd,xs]   )[a',b']
//   Q_y [a',b'] = sum_xs psi_{t+1}[xs,y]   (A_{t+1}[xs,y] Renv_{t+2}  B_{t+1}[y,xs]^T)[a',b']
//   p[xs,y] ~ <P_xs, Q_y>.
// dyn smem: (4 + qs) * dcap*dcap doubles.  out: [t][xs + qs*y], edge stride as for the pair beliefs (slot t = T unused).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_alt_marginal(const PairJob* jobs, int L, int dcap) {
  extern __shared__ double smem[];
  __shared__ double red[NW + 1];
  __shared__ double pm[64];
  const PairJob& jb = jobs[blockIdx.x];
  const int qs = jb.qs, qd = jb.qd;
  const int est = dcap * dcap;
  double* tmp = smem;
  double* L0 = smem + est;
  double* L1 = smem + 2 * est;
  double* Q = smem + 3 * est;
  double* P = smem + 4 * est;  // [xs][a' + ar*b']
  if (threadIdx.x == 0) jb.Renv[(size_t)L * est] = 1.0;
  __syncthreads();
  // backward environments, identical to k_pair_belief
  for (int t = L - 1; t >= 0; --t) {
    const int al = jb.A.bonds[t], ar = jb.A.bonds[t + 1], bl = jb.B.bonds[t], br = jb.B.bonds[t + 1];
    const double* At = jb.A.data + (size_t)t * jb.A.stride;
    const double* Bt = jb.B.data + (size_t)t * jb.B.stride;
    const double* psi = jb.psi + (size_t)t * qs * qd;
    const double* Rn = jb.Renv + (size_t)(t + 1) * est;
    double* Rt = jb.Renv + (size_t)t * est;
    for (int idx = threadIdx.x; idx < al * bl; idx += NT) Rt[idx] = 0.0;
    __syncthreads();
    for (int xs = 0; xs < qs; ++xs)
      for (int xd = 0; xd < qd; ++xd) {
        const double ps = psi[xs + qs * xd];
        if (ps == 0.0) continue;
        const double* Ax = At + (size_t)al * ar * (xs + qs * xd);
        const double* Bx = Bt + (size_t)bl * br * (xd + qd * xs);
        for (int idx = threadIdx.x; idx < al * br; idx += NT) {
          const int a = idx % al, b2 = idx / al;
          double acc = 0.0;
          for (int a2 = 0; a2 < ar; ++a2) acc += Ax[a + al * a2] * Rn[a2 + ar * b2];
          tmp[idx] = acc;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < al * bl; idx += NT) {
          const int a = idx % al, b = idx / al;
          double acc = 0.0;
          for (int b2 = 0; b2 < br; ++b2) acc += tmp[a + al * b2] * Bx[b + bl * b2];
          Rt[idx] += ps * acc;
        }
        __syncthreads();
      }
    double mx = 0.0;
    for (int idx = threadIdx.x; idx < al * bl; idx += NT) mx = fmax(mx, fabs(Rt[idx]));
    mx = block_max(mx, red);
    const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
    for (int idx = threadIdx.x; idx < al * bl; idx += NT) Rt[idx] *= f;
    __syncthreads();
  }
  if (threadIdx.x == 0) L0[0] = 1.0;
  __syncthreads();
  for (int t = 0; t < L - 1; ++t) {
    const int al = jb.A.bonds[t], ar = jb.A.bonds[t + 1], bl = jb.B.bonds[t], br = jb.B.bonds[t + 1];
    const int cr = jb.A.bonds[t + 2], dr = jb.B.bonds[t + 2];
    const double* At = jb.A.data + (size_t)t * jb.A.stride;
    const double* Bt = jb.B.data + (size_t)t * jb.B.stride;
    const double* psi = jb.psi + (size_t)t * qs * qd;
    const double* At1 = jb.A.data + (size_t)(t + 1) * jb.A.stride;
    const double* Bt1 = jb.B.data + (size_t)(t + 1) * jb.B.stride;
    const double* psi1 = jb.psi + (size_t)(t + 1) * qs * qd;
    const double* Rn2 = jb.Renv + (size_t)(t + 2) * est;  // [a'' + cr*b'']
    const int nab = ar * br;
    for (int idx = threadIdx.x; idx < qs * est; idx += NT) P[idx] = 0.0;
    __syncthreads();
    for (int xs = 0; xs < qs; ++xs)
      for (int xd = 0; xd < qd; ++xd) {
        const double ps = psi[xs + qs * xd];
        if (ps == 0.0) continue;
        const double* Ax = At + (size_t)al * ar * (xs + qs * xd);
        const double* Bx = Bt + (size_t)bl * br * (xd + qd * xs);
        for (int idx = threadIdx.x; idx < ar * bl; idx += NT) {
          const int a2 = idx % ar, b = idx / ar;
          double acc = 0.0;
          for (int a = 0; a < al; ++a) acc += Ax[a + al * a2] * L0[a + al * b];
          tmp[idx] = acc;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < nab; idx += NT) {
          const int a2 = idx % ar, b2 = idx / ar;
          double acc = 0.0;
          for (int b = 0; b < bl; ++b) acc += tmp[a2 + ar * b] * Bx[b + bl * b2];
          P[(size_t)xs * est + idx] += ps * acc;
        }
        __syncthreads();
      }
    for (int y = 0; y < qd; ++y) {
      for (int idx = threadIdx.x; idx < nab; idx += NT) Q[idx] = 0.0;
      __syncthreads();
      for (int xs = 0; xs < qs; ++xs) {
        const double ps = psi1[xs + qs * y];
        if (ps == 0.0) continue;
        const double* Ax = At1 + (size_t)ar * cr * (xs + qs * y);
        const double* Bx = Bt1 + (size_t)br * dr * (y + qd * xs);
        for (int idx = threadIdx.x; idx < ar * dr; idx += NT) {
          const int a = idx % ar, b2 = idx / ar;
          double acc = 0.0;
          for (int a2 = 0; a2 < cr; ++a2) acc += Ax[a + ar * a2] * Rn2[a2 + cr * b2];
          tmp[idx] = acc;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < nab; idx += NT) {
          const int a = idx % ar, b = idx / ar;
          double acc = 0.0;
          for (int b2 = 0; b2 < dr; ++b2) acc += tmp[a + ar * b2] * Bx[b + br * b2];
          Q[idx] += ps * acc;
        }
        __syncthreads();
      }
      for (int xs = 0; xs < qs; ++xs) {
        double part = 0.0;
        for (int idx = threadIdx.x; idx < nab; idx += NT) part += P[(size_t)xs * est + idx] * Q[idx];
        part = block_sum1(part, red);
        if (threadIdx.x == 0) pm[xs + qs * y] = part;
        __syncthreads();
      }
    }
    if (threadIdx.x == 0) {
      double sum = 0.0;
      for (int i = 0; i < qs * qd; ++i) sum += pm[i];
      for (int i = 0; i < qs * qd; ++i) jb.out[(size_t)t * qs * qd + i] = pm[i] / sum;
    }
    // Lenv_{t+1} = sum_xs P_xs, rescaled
    double mx = 0.0;
    for (int idx = threadIdx.x; idx < nab; idx += NT) {
      double v = 0.0;
      for (int xs = 0; xs < qs; ++xs) v += P[(size_t)xs * est + idx];
      L1[idx] = v;
      mx = fmax(mx, fabs(v));
    }
    mx = block_max(mx, red);
    const double f = (mx > 0.0 && isfinite(mx)) ? 1.0 / mx : 1.0;
    for (int idx = threadIdx.x; idx < nab; idx += NT) L0[idx] = L1[idx] * f;
    __syncthreads();
  }
  if (threadIdx.x < qs * qd) jb.out[(size_t)(L - 1) * qs * qd + threadIdx.x] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// forward sampler of the prior dynamics (reference onesample!, src/sampling.jl:30-59): x_i^0 ~ phi_i^0 / sum,
// x_i^{t+1} ~ w_i^t(. | x_{di}^t, x_i^t).  The transition probability of a RecursiveBPFactor is evaluated from the class
// tables exactly as the factor's functor does (src/recursive_bp_factor.jl:34-46), folding the neighbours in cavity order:
// P_1 = Pxy_1, P_{k+1} = Pyy(k,1)[P_k, Pxy_{k+1}], P_full = Pyy(z,0)[P_z, Minit], p(x') = sum_y Wd[x',x,y] P_full[y];
// a generic BPFactor indexes its dense table with the joint neighbour state.  One thread per node, one launch per time
// step.  Counter-based RNG (splitmix64 of (seed, node, time)) and no FMA contraction, so the oracle's restatement
// reproduces the trajectory bit for bit.
// ------------------------------------------------------------------------------------------------
constexpr int SAMP_MAXZ = 32;
constexpr int SAMP_MAXNY = 64;
struct SampCls {
  int z, q, generic;
  int ny[SAMP_MAXZ + 1];
  int qn[SAMP_MAXZ];
  const double* pxy;
  long long pxy_ts, pxy_off[SAMP_MAXZ];
  const double* pyy;
  long long pyy_off[SAMP_MAXZ + 1], pyy_ts[SAMP_MAXZ + 1];  // [k] : pair (k,1) for 1 <= k < z ; [z] : pair (z,0)
  const double* wd;
  long long wd_ts;
  const double* minit;
  long long minit_ts;
};
__host__ __device__ inline double samp_uniform(unsigned long long seed, long long i, int t) {
  unsigned long long zz = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1) + 0xD1B54A32D192ED03ull * (unsigned long long)(t + 2);
  zz = (zz ^ (zz >> 30)) * 0xBF58476D1CE4E5B9ull;
  zz = (zz ^ (zz >> 27)) * 0x94D049BB133111EBull;
  zz = zz ^ (zz >> 31);
  return (double)(zz >> 11) * (1.0 / 9007199254740992.0);
}
__device__ inline int samp_draw(const double* p, int q, double u) {
  double S = 0.0;
  for (int x = 0; x < q; ++x) S = __dadd_rn(S, p[x]);
  double c = 0.0;
  for (int x = 0; x < q; ++x) {
    c = __dadd_rn(c, __ddiv_rn(p[x], S));
    if (u < c) return x;
  }
  return q - 1;
}
// t < 0: initial condition; else step t -> t+1.  X[i*L + t], states 0-based.
__global__ void k_sample_step(const SampCls* cls, const int* class_of_node, const int64_t* colptr, const int64_t* dst, const double* phi,
                              const int64_t* phi_off, const int* qarr, long long N, int L, int t, unsigned long long seed, int* X, int* err) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int q = qarr[i];
  double p[8];
  if (t < 0) {
    for (int x = 0; x < q; ++x) p[x] = phi[phi_off[i] + x];
    X[i * L] = samp_draw(p, q, samp_uniform(seed, i, -1));
    return;
  }
  const int ci = class_of_node[i];
  if (ci < 0) { atomicOr(err, ERR_NAN); return; }
  const SampCls& c = cls[ci];
  const int x = X[i * L + t];
  const int64_t e0 = colptr[i];
  if (c.generic) {
    long long yall = 0, mul = 1;
    for (int k = 0; k < c.z; ++k) {
      yall += mul * X[dst[e0 + k] * L + t];
      mul *= c.qn[k];
    }
    const double* wd = c.wd + (size_t)t * c.wd_ts;
    for (int xn = 0; xn < q; ++xn) p[xn] = wd[xn + q * (x + (size_t)q * yall)];
  } else {
    double P[SAMP_MAXNY], Pn[SAMP_MAXNY];
    const double* minit = c.minit + (size_t)t * c.minit_ts;
    const int ny0 = c.ny[0], ny1 = c.ny[1];
    int nycur;
    if (c.z == 0) {
      nycur = ny0;
      for (int y = 0; y < ny0; ++y) P[y] = minit[y + ny0 * x];
    } else {
      const double* pxy_t = c.pxy + (size_t)t * c.pxy_ts;
      {
        const int xk = X[dst[e0] * L + t];
        const double* px = pxy_t + c.pxy_off[0];
        for (int y = 0; y < ny1; ++y) P[y] = px[y + ny1 * (xk + c.qn[0] * x)];
        nycur = ny1;
      }
      for (int k = 1; k < c.z; ++k) {
        const int xk = X[dst[e0 + k] * L + t];
        const double* px = pxy_t + c.pxy_off[k];
        const double* pyy = c.pyy + c.pyy_off[k] + (size_t)t * c.pyy_ts[k];
        const int nyn = c.ny[k + 1];
        for (int y = 0; y < nyn; ++y) {
          double acc = 0.0;
          for (int y1 = 0; y1 < nycur; ++y1)
            for (int y2 = 0; y2 < ny1; ++y2)
              acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(pyy[y + nyn * (y1 + nycur * (y2 + ny1 * x))], P[y1]), px[y2 + ny1 * (xk + c.qn[k] * x)]));
          Pn[y] = acc;
        }
        for (int y = 0; y < nyn; ++y) P[y] = Pn[y];
        nycur = nyn;
      }
      // full = op(p[z-1], init): pair (z, 0)
      const double* pyy = c.pyy + c.pyy_off[c.z] + (size_t)t * c.pyy_ts[c.z];
      const int nyz = c.ny[c.z];
      for (int y = 0; y < nyz; ++y) {
        double acc = 0.0;
        for (int y1 = 0; y1 < nycur; ++y1)
          for (int y0 = 0; y0 < ny0; ++y0)
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(pyy[y + nyz * (y1 + nycur * (y0 + ny0 * x))], P[y1]), minit[y0 + ny0 * x]));
        Pn[y] = acc;
      }
      for (int y = 0; y < nyz; ++y) P[y] = Pn[y];
      nycur = nyz;
    }
    const double* wd = c.wd + (size_t)t * c.wd_ts;
    for (int xn = 0; xn < q; ++xn) {
      double acc = 0.0;
      for (int y = 0; y < nycur; ++y) acc = __dadd_rn(acc, __dmul_rn(wd[xn + q * (x + q * y)], P[y]));
      p[xn] = acc;
    }
  }
  X[i * L + t + 1] = samp_draw(p, q, samp_uniform(seed, i, t));
}

// ------------------------------------------------------------------------------------------------
// Minit TT: [1,1,y,x] = prob_y0 per t
struct InitJob {
  TTRef out;
  const double* minit;  // [y + ny0*x] per t
  int tstride, n;       // n = ny0*q
};
__global__ void k_init_tt(const InitJob* jobs, int njobs, int L) {
  const int j = blockIdx.x;
  if (j >= njobs) return;
  const InitJob& jb = jobs[j];
  for (int idx = threadIdx.x; idx < L * jb.n; idx += blockDim.x) {
    const int t = idx / jb.n, e = idx % jb.n;
    jb.out.data[(size_t)t * jb.out.stride + e] = jb.minit[(size_t)t * jb.tstride + e];
  }
  for (int i = threadIdx.x; i <= L; i += blockDim.x) jb.out.bonds[i] = 1;
  if (threadIdx.x == 0) *jb.out.ls = 0.0;
}

// f_i = (z/2 - 1) logz_i - 1/2 sum_j logz_{i->j}      (src/recursive_bp_factor.jl:163)
struct FJob {
  const double* logzi;
  const double* logzij;  // contiguous z entries
  int z;
  double* f;
};
__global__ void k_free_energy(const FJob* jobs, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const FJob jb = jobs[i];
  double s = 0.0;
  for (int j = 0; j < jb.z; ++j) s += jb.logzij[j];
  *jb.f = (0.5 * jb.z - 1.0) * (*jb.logzi) - 0.5 * s;
}

// flat message: every site all-ones with bond 1, normalised: ls = -L*log(qs*qd)
__global__ void k_flat_messages(double* data, int* bonds, double* ls, const int* qprod, long long slot, int L,
                                long long E2, int sstride) {
  const long long e = blockIdx.x;
  if (e >= E2) return;
  const int P = qprod[e];
  for (int idx = threadIdx.x; idx < L * P; idx += blockDim.x) {
    const int t = idx / P, k = idx % P;
    data[e * slot + (size_t)t * sstride + k] = 1.0;
  }
  for (int i = threadIdx.x; i <= L; i += blockDim.x) bonds[e * (L + 1) + i] = 1;
  if (threadIdx.x == 0) ls[e] = -(double)L * log((double)P);
}

// means and delta: mean[i,t] = sum_x obs[i,x] marg[i,t,x]; delta = max |new - old| over the listed nodes
__global__ void k_means_delta(const double* marg, const int64_t* moff, const int* q, const double* obs, int qmax,
                              const int64_t* nodes, long long nn, int L, double* means, double* delta) {
  const long long k = blockIdx.x;
  if (k >= nn) return;
  const long long i = nodes ? (long long)nodes[k] : k;
  const int qi = q[i];
  double mx = 0.0;
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    double m = 0.0;
    for (int x = 0; x < qi; ++x) m += (obs ? obs[i * qmax + x] : (double)(x + 1)) * marg[moff[i] + (size_t)t * qi + x];
    const double old = means[i * L + t];
    mx = fmax(mx, fabs(m - old));
    means[i * L + t] = m;
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0 && mx > 0.0) {
    // atomic max on non-negative doubles via long long compare
    atomicMax(reinterpret_cast<unsigned long long*>(delta), (unsigned long long)__double_as_longlong(mx));
  }
}

}  // namespace mpbp
