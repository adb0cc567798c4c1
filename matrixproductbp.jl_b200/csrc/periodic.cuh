// periodic.cuh -- periodic-in-time MPBP on the device.
//
// Reference code replaced (all paths relative to /root/reference):
//   src/mpems.jl:96-155                    PeriodicMPEM3, trace evaluation, mpem2(::PeriodicMPEM3)
//   src/recursive_bp_factor.jl:89-101      _f_bp_partial for PeriodicMPEM2 (every site carries the factor)
//   src/recursive_bp_factor.jl:104-165     compute_prob_ys / op / onebpiter! on ring tensor trains
//   src/mpbp.jl:399-409                    periodic_mpbp (flat periodic messages)
//   TensorTrains.jl (un-vendored)          compress! / orthogonalize_left! / orthogonalize_right! / normalize! / marginals of a
//                                          PeriodicTensorTrain: full turns around the ring, the closing bond included
//
// A ring tensor train closes its matrix product with a trace, so its sweeps are not isometric-gauge truncations: the
// gauge-invariance argument that lets the open-boundary engine replace SVD sweeps by Q-less QRs
// does not hold.  This path therefore performs the ring sweeps LITERALLY (one SVD per site and sweep, carry multiplied
// into the neighbour, the closing bond visited once per turn).  One CTA updates one node: B~_k, the 3z-2 cavity `op`s
// (Kronecker product + ring compress + normalize_eachmatrix), per out-edge f_bp_partial -> periodic mpem2 ->
// compress(left) -> normalize, and the belief / log z_i from bond x state transfer matrices (the belief train of the
// reference is an exact refactorisation of them, so no SVD is needed for it).  The SVD is a one-sided (Hestenes) Jacobi
// on the shorter side of the matrix, one warp per column pair, round-robin pairing, in global / L2 memory.
//
// Singular values below 1e-14 sigma_1 are dropped under every truncation policy (the reference keeps them as
// functionally irrelevant bond dimensions up to the cap).
//
// The file is written against a tiny execution-model shim (PER_*) so that tests/host_emul compiles THE SAME source with
// g++ as a one-thread CPU emulation and checks it against oracle/periodic.py in the CPU test tier; the product only
// ever runs the CUDA build.
#pragma once
#include <math.h>
#include <stddef.h>

#ifdef PER_HOST
#define PER_FN static inline
#define PER_TID 0
#define PER_NTH 1
#define PER_LANE 0
#define PER_WSZ 1
#define PER_WARP 0
#define PER_NWARPS 1
#define PER_SYNC() ((void)0)
static inline double per_warp_sum(double v) { return v; }
static inline double per_warp_max(double v) { return v; }
#else
#define PER_FN __device__ inline
#define PER_TID ((int)threadIdx.x)
#define PER_NTH ((int)blockDim.x)
#define PER_LANE ((int)(threadIdx.x & 31))
#define PER_WSZ 32
#define PER_WARP ((int)(threadIdx.x >> 5))
#define PER_NWARPS ((int)(blockDim.x >> 5))
#define PER_SYNC() __syncthreads()
__device__ __forceinline__ double per_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double per_warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif
#define PER_FOR(i, n) for (int i = PER_TID; i < (int)(n); i += PER_NTH)

namespace mpbp_per {

constexpr int PER_MAXZ = 10;     // largest node degree of the periodic path
constexpr int PER_MAXW = 64;     // warps per CTA upper bound (reduction scratch)
constexpr int PER_ERR_BOND = 1;  // same bit values as mpbp::ERR_*
constexpr int PER_ERR_NAN = 2;
constexpr int PER_ERR_NOCONV = 4;

// ring tensor train: site t at data + t*stride, column-major [bl[t], bl[(t+1)%L], X]; bonds has L+1 entries with
// bonds[L] == bonds[0] (the closing bond), so that a message slot reads like an open one with non-trivial ends.
// value(x) = exp(*ls) * trace prod_t A_t[:, :, x_t]
struct PTT {
  double* data;
  int* bonds;
  double* ls;
  int stride;
  int X;
};

struct PTrunc {
  int kind;  // 0 bond, 1 thresh, 2 bond + thresh
  int d;
  double eps;
};

struct PerOp {
  int a, b, o;            // TT registers
  int ny1, ny2, nyo;
  const double* pyy;      // [y + nyo*(y1 + ny1*(y2 + ny2*x))] per t
  int pyy_ts;
};

// scratch of one node update (global memory)
struct PerWS {
  PTT K;          // ring workspace, site capacity `cap` doubles
  int cap;        // doubles per workspace site / per matrix buffer below
  double *G, *Lf, *Rf, *Tmp, *Mb;  // `cap` doubles each: Jacobi matrix, the two SVD factors, site copy, unfolding
  double* W;      // wcap*wcap: accumulated rotations
  int wcap;
  double* sig;    // 2*wcap
  int* perm;      // wcap
  double* red;    // PER_MAXW + 8: block reductions, scalar broadcast
  int* ib;        // 8 ints: flags / broadcast
  double* tm;     // transfer matrices of the belief: (3L + 3 + 2q) * wcap^2
};

struct PerNode {
  int z, q, L, dmax;
  PTrunc tr;
  double damp;  // set_msg! damping (src/recursive_bp_factor.jl:168-179); the old message is read from msg_out[j]
  int nreg;
  PTT reg[4 * PER_MAXZ + 4];
  int nops;
  PerOp ops[3 * PER_MAXZ];
  // B~_k
  PTT msg_in[PER_MAXZ];
  const double* psi[PER_MAXZ];  // [t][xi + q*xk]
  const double* pxy[PER_MAXZ];  // [y + ny1*(xk + qk*xi)] per t
  int pxy_ts, ny1;
  int qn[PER_MAXZ];
  int src_reg[PER_MAXZ];
  // init
  const double* minit;  // [t][y + ny0*x]
  int minit_ts, ny0, init_reg;
  // outgoing messages
  int dest_reg[PER_MAXZ];
  PTT msg_out[PER_MAXZ];
  const double* Wp[PER_MAXZ];  // [x' + q*(x + q*(xj + qj*y))] per t
  int w_ts, nyc;
  // belief
  int full_reg;
  const double* Wd;  // [x' + q*(x + q*y)] per t
  int wd_ts, nyz;
  const double* phi;  // [t][x]
  double *marg, *logzi, *logzij, *f;
  double* tv;        // two-time marginals b_i(x^t, x^u) (twovar_marginals(bp.b[i]), src/mpbp.jl:239), nullptr = off:
  int tv_maxdist;    //   [t][u][x_t + q*x_u], (t,u) stride tv_q2cap, zero unless t < u <= t + tv_maxdist
  int tv_q2cap;
  int* err;  // the handle's error word (OR of the PER_ERR_* bits of every node)
  PerWS ws;
};

// ------------------------------------------------------------------------------------------------
// small CTA-level helpers (results broadcast through global scratch: no shared memory on this path)
// ------------------------------------------------------------------------------------------------
PER_FN double per_block_max(double v, double* red) {
  v = per_warp_max(v);
  PER_SYNC();
  if (PER_LANE == 0) red[PER_WARP] = v;
  PER_SYNC();
  double r = red[0];
  for (int w = 1; w < PER_NWARPS; ++w) r = fmax(r, red[w]);
  return r;
}

// largest |a_i| over n entries; NaN / Inf entries raise `bad`
PER_FN double per_maxabs(const double* a, int n, double* red, int* err) {
  double m = 0.0;
  int bad = 0;
  PER_FOR(i, n) {
    const double v = fabs(a[i]);
    if (!(v <= 1.79e308)) bad = 1;  // NaN or Inf
    else m = fmax(m, v);
  }
  if (bad) *err |= PER_ERR_NAN;  // (plain read-modify-write of the node's own error word: concurrent writers all set this same bit,
                                 // and any non-zero word aborts the node update at the next per_failed())
  return per_block_max(m, red);
}

PER_FN int per_trunc_keep(const PTrunc& tr, const double* s, int n) {
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
  if (tr.kind == 0 || tr.kind == 2) k = k < tr.d ? k : tr.d;
  return k;
}

// ------------------------------------------------------------------------------------------------
// one-sided Jacobi: orthogonalise the c columns (length len, column-major) of G in place, W (c x c) = product of the
// rotations (G_final = G_initial * W).  sig = column norms sorted descending, perm = the matching column indices.
// G is expected scaled to max |entry| = 1 (columns with squared norm below 1e-40 count as zero).
// ------------------------------------------------------------------------------------------------
PER_FN void per_jacobi(double* G, int len, int c, double* W, double* sig, int* perm, int* ib, int* err) {
  PER_FOR(i, c * c) W[i] = (i % c == i / c) ? 1.0 : 0.0;
  const int cc = c + (c & 1);
  const int npairs = cc / 2;
  const double tol = 4.5e-16 * sqrt((double)len) + 2.3e-16;
  const int maxsweep = 50;
  PER_SYNC();
  for (int sweep = 0; sweep < maxsweep && c > 1; ++sweep) {
    if (PER_TID == 0) { ib[0] = 0; ib[1] = 0; }
    PER_SYNC();
    for (int r = 0; r < cc - 1; ++r) {
      for (int p = PER_WARP; p < npairs; p += PER_NWARPS) {
        int i, j;
        if (p == 0) { i = cc - 1; j = r; }
        else { i = (r + p) % (cc - 1); j = (r + cc - 1 - p) % (cc - 1); }
        if (i >= c || j >= c) continue;  // the padding player of an odd tournament
        if (i > j) { const int s_ = i; i = j; j = s_; }
        double* gi = G + (size_t)i * len;
        double* gj = G + (size_t)j * len;
        double aii = 0, ajj = 0, aij = 0;
        for (int k = PER_LANE; k < len; k += PER_WSZ) {
          const double x = gi[k], y = gj[k];
          aii += x * x;
          ajj += y * y;
          aij += x * y;
        }
        aii = per_warp_sum(aii);
        ajj = per_warp_sum(ajj);
        aij = per_warp_sum(aij);
        if (aii < 1e-40 || ajj < 1e-40) continue;
        const double lim = sqrt(aii) * sqrt(ajj);
        if (fabs(aij) <= tol * lim) continue;
        const double zeta = (ajj - aii) / (2.0 * aij);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
        for (int k = PER_LANE; k < len; k += PER_WSZ) {
          const double x = gi[k], y = gj[k];
          gi[k] = cs * x - sn * y;
          gj[k] = sn * x + cs * y;
        }
        double* wi = W + (size_t)i * c;
        double* wj = W + (size_t)j * c;
        for (int k = PER_LANE; k < c; k += PER_WSZ) {
          const double x = wi[k], y = wj[k];
          wi[k] = cs * x - sn * y;
          wj[k] = sn * x + cs * y;
        }
        if (PER_LANE == 0) {
          ib[0] = 1;
          if (fabs(aij) > 1e-10 * lim) ib[1] = 1;
        }
      }
      PER_SYNC();
    }
    const int any = ib[0], coarse = ib[1];
    PER_SYNC();
    if (!any) break;
    if (sweep == maxsweep - 1 && coarse && PER_TID == 0) *err |= PER_ERR_NOCONV;
  }
  // column norms
  for (int k = PER_WARP; k < c; k += PER_NWARPS) {
    const double* g = G + (size_t)k * len;
    double s = 0;
    for (int i = PER_LANE; i < len; i += PER_WSZ) s += g[i] * g[i];
    s = per_warp_sum(s);
    if (PER_LANE == 0) sig[c + k] = sqrt(s);  // unsorted norms in the upper half of sig (capacity 2*wcap)
  }
  PER_SYNC();
  if (PER_TID == 0) {
    for (int k = 0; k < c; ++k) perm[k] = k;
    for (int a = 1; a < c; ++a) {  // insertion sort, descending, stable
      const int pk = perm[a];
      const double v = sig[c + pk];
      int b = a - 1;
      while (b >= 0 && sig[c + perm[b]] < v) { perm[b + 1] = perm[b]; --b; }
      perm[b + 1] = pk;
    }
    for (int k = 0; k < c; ++k) sig[k] = sig[c + perm[k]];
  }
  PER_SYNC();
}

// ------------------------------------------------------------------------------------------------
// truncated SVD  M (R x C, column-major) = U diag(lam) Vt  as two explicit factors:
//   wl = 0 (orthogonalize_right):  Lf = U lam (R x r),  Rf = Vt     (r x C)
//   wl = 1 (orthogonalize_left) :  Lf = U     (R x r),  Rf = lam Vt (r x C)
// M is divided by its largest |entry| first and the log of it is added to *ls (the reference does the same before
// every SVD of a sweep).  Returns r (the same value in every thread).
// ------------------------------------------------------------------------------------------------
PER_FN int per_svd_factor(const double* M, int R, int C, int wl, const PTrunc& tr, double* ls, PerWS& ws, int* err) {
  const double mx = per_maxabs(M, R * C, ws.red, err);
  double sc = 1.0;
  if (mx > 0.0) {
    sc = 1.0 / mx;
    if (PER_TID == 0) *ls += log(mx);
  }
  const bool tall = R >= C;
  const int len = tall ? R : C, c = tall ? C : R;
  double* G = ws.G;
  if (tall) {
    PER_FOR(i, R * C) G[i] = M[i] * sc;
  } else {
    PER_FOR(i, R * C) {
      const int row = i % R, col = i / R;
      G[col + (size_t)C * row] = M[i] * sc;
    }
  }
  PER_SYNC();
  per_jacobi(G, len, c, ws.W, ws.sig, ws.perm, ws.ib, err);
  if (PER_TID == 0) {
    int nz = 0;
    while (nz < c && ws.sig[nz] > 1e-14 * ws.sig[0] && ws.sig[nz] > 0.0) ++nz;
    if (nz < 1) nz = 1;
    int k = per_trunc_keep(tr, ws.sig, c);
    if (k > nz) k = nz;
    ws.ib[2] = k;
  }
  PER_SYNC();
  const int r = ws.ib[2];
  const double* W = ws.W;
  // F1(i,k) = tall ? G[i + R*pk] (= lam u) : W[i + R*pk] (= u);   F2(k,j) = tall ? W[j + C*pk] (= v) : G[j + C*pk] (= lam v)
  PER_FOR(idx, R * r) {
    const int i = idx % R, k = idx / R, pk = ws.perm[k];
    const double s = ws.sig[k], inv = s > 0.0 ? 1.0 / s : 0.0;
    const double f1 = tall ? G[i + (size_t)R * pk] : W[i + (size_t)R * pk];
    ws.Lf[idx] = tall ? (wl == 0 ? f1 : f1 * inv) : (wl == 0 ? f1 * s : f1);
  }
  PER_FOR(idx, r * C) {
    const int k = idx % r, j = idx / r, pk = ws.perm[k];
    const double s = ws.sig[k], inv = s > 0.0 ? 1.0 / s : 0.0;
    const double f2 = tall ? W[j + (size_t)C * pk] : G[j + (size_t)C * pk];
    ws.Rf[idx] = tall ? (wl == 0 ? f2 : f2 * s) : (wl == 0 ? f2 * inv : f2);
  }
  PER_SYNC();
  return r;
}

PER_FN int per_bl(const PTT& A, int t) { return A.bonds[t]; }                  // left bond of site t (t < L)
PER_FN int per_br(const PTT& A, int t, int L) { return A.bonds[(t + 1) % L]; }  // right bond of site t

PER_FN void per_close(const PTT& A, int L) {  // keep the duplicate of the closing bond in sync
  PER_SYNC();
  if (PER_TID == 0) A.bonds[L] = A.bonds[0];
  PER_SYNC();
}

// one full turn right to left (TensorTrains orthogonalize_right! on a ring): A[t] <- Vt, U lam into the site on its left
PER_FN void per_orth_right(const PTT& A, int L, const PTrunc& tr, PerWS& ws, int cap, int* err) {
  const int X = A.X;
  for (int t = L - 1; t >= 0; --t) {
    const int m = per_bl(A, t), n = per_br(A, t, L);
    double* At = A.data + (size_t)t * A.stride;
    const int r = per_svd_factor(At, m, n * X, 0, tr, A.ls, ws, err);
    PER_FOR(i, r * n * X) At[i] = ws.Rf[i];
    PER_SYNC();
    if (PER_TID == 0) A.bonds[t] = r;
    const int p = (t - 1 + L) % L;
    double* Ap = A.data + (size_t)p * A.stride;
    const int mp = (p == t) ? r : per_bl(A, p);  // L == 1: the site is its own left neighbour
    if ((size_t)mp * r * X > (size_t)cap) { if (PER_TID == 0) *err |= PER_ERR_BOND; PER_SYNC(); return; }
    PER_FOR(i, mp * m * X) ws.Tmp[i] = Ap[i];
    PER_SYNC();
    PER_FOR(idx, mp * r * X) {
      const int a = idx % mp, k = (idx / mp) % r, x = idx / (mp * r);
      double acc = 0;
      for (int b = 0; b < m; ++b) acc += ws.Tmp[a + (size_t)mp * (b + (size_t)m * x)] * ws.Lf[b + (size_t)m * k];
      Ap[idx] = acc;
    }
    PER_SYNC();
  }
  per_close(A, L);
}

// one full turn left to right (orthogonalize_left!): A[t] <- U, lam Vt into the site on its right
PER_FN void per_orth_left(const PTT& A, int L, const PTrunc& tr, PerWS& ws, int cap, int* err) {
  const int X = A.X;
  for (int t = 0; t < L; ++t) {
    const int m = per_bl(A, t), n = per_br(A, t, L);
    double* At = A.data + (size_t)t * A.stride;
    PER_FOR(idx, m * n * X) {  // M[(a, x), b] = A[a, b, x]
      const int a = idx % m, b = (idx / m) % n, x = idx / (m * n);
      ws.Mb[a + (size_t)m * x + (size_t)m * X * b] = At[idx];
    }
    PER_SYNC();
    const int r = per_svd_factor(ws.Mb, m * X, n, 1, tr, A.ls, ws, err);
    PER_FOR(idx, m * r * X) {
      const int a = idx % m, k = (idx / m) % r, x = idx / (m * r);
      At[idx] = ws.Lf[a + (size_t)m * x + (size_t)m * X * k];
    }
    PER_SYNC();
    const int nx = (t + 1) % L;
    if (PER_TID == 0) A.bonds[nx] = r;
    double* An = A.data + (size_t)nx * A.stride;
    // A[nx] is [n, n2, X]; if nx == t (L == 1) the site was just rewritten as [m, r, X] with m == n
    const int n2 = (nx == t) ? r : per_br(A, nx, L);
    PER_FOR(i, n * n2 * X) ws.Tmp[i] = An[i];
    PER_SYNC();
    PER_FOR(idx, r * n2 * X) {
      const int k = idx % r, c2 = (idx / r) % n2, x = idx / (r * n2);
      double acc = 0;
      for (int b = 0; b < n; ++b) acc += ws.Rf[k + (size_t)r * b] * ws.Tmp[b + (size_t)n * (c2 + (size_t)n2 * x)];
      An[idx] = acc;
    }
    PER_SYNC();
  }
  per_close(A, L);
}

PER_FN void per_normalize_eachmatrix(const PTT& A, int L, PerWS& ws, int* err) {
  for (int t = 0; t < L; ++t) {
    double* At = A.data + (size_t)t * A.stride;
    const int n = per_bl(A, t) * per_br(A, t, L) * A.X;
    const double mx = per_maxabs(At, n, ws.red, err);
    if (mx > 0.0) {
      const double sc = 1.0 / mx;
      PER_FOR(i, n) At[i] *= sc;
      if (PER_TID == 0) *A.ls += log(mx);
    }
    PER_SYNC();
  }
}

// C (m x n) = A (m x k) * B (k x n), column-major, then scaled by its largest |entry|; returns log of that scale
PER_FN double per_matmul_scaled(double* Cm, const double* Am, const double* Bm, int m, int k, int n, double* red, int* err) {
  PER_FOR(idx, m * n) {
    const int i = idx % m, j = idx / m;
    double acc = 0;
    for (int l = 0; l < k; ++l) acc += Am[i + (size_t)m * l] * Bm[l + (size_t)k * j];
    Cm[idx] = acc;
  }
  PER_SYNC();
  const double mx = per_maxabs(Cm, m * n, red, err);
  if (mx > 0.0) {
    const double sc = 1.0 / mx;
    PER_FOR(i, m * n) Cm[i] *= sc;
    PER_SYNC();
    return log(mx);
  }
  return 0.0;
}

// log | trace prod_t S_t |,  S_t = sum_x A_t[:, :, x]   (TensorTrains normalization of a ring).  Uses ws.tm (3 matrices).
PER_FN double per_ring_lognorm(const PTT& A, int L, PerWS& ws, int* err) {
  const int X = A.X;
  const int m0 = per_bl(A, 0);
  double* cur = ws.tm;
  double* nxt = ws.tm + (size_t)ws.wcap * ws.wcap;
  double* S = ws.tm + 2 * (size_t)ws.wcap * ws.wcap;
  double acc = 0.0;
  PER_FOR(i, m0 * m0) cur[i] = (i % m0 == i / m0) ? 1.0 : 0.0;
  PER_SYNC();
  for (int t = 0; t < L; ++t) {
    const int m = per_bl(A, t), n = per_br(A, t, L);
    const double* At = A.data + (size_t)t * A.stride;
    PER_FOR(i, m * n) {
      double s = 0;
      for (int x = 0; x < X; ++x) s += At[i + (size_t)m * n * x];
      S[i] = s;
    }
    PER_SYNC();
    acc += per_matmul_scaled(nxt, cur, S, m0, m, n, ws.red, err);
    double* sw = cur; cur = nxt; nxt = sw;
  }
  double tr = 0;
  for (int i = 0; i < m0; ++i) tr += cur[i + (size_t)m0 * i];
  PER_SYNC();
  if (!(tr > 0.0)) { if (PER_TID == 0) *err |= PER_ERR_NAN; return acc; }
  return acc + log(tr);
}

PER_FN void per_copy_tt(const PTT& dst, const PTT& src, int L, int* err) {
  for (int t = 0; t < L; ++t) {
    const int n = src.bonds[t] * src.bonds[(t + 1) % L] * src.X;
    if (n > dst.stride) { if (PER_TID == 0) *err |= PER_ERR_BOND; continue; }
    PER_FOR(i, n) dst.data[(size_t)t * dst.stride + i] = src.data[(size_t)t * src.stride + i];
  }
  PER_FOR(t, L + 1) dst.bonds[t] = src.bonds[t % L];
  if (PER_TID == 0) *dst.ls = *src.ls;
  PER_SYNC();
}

// compress!(A; svd_trunc) of a ring: un-truncated turn right to left, truncating turn left to right
PER_FN void per_compress(const PTT& A, int L, const PTrunc& tr, PerWS& ws, int* err) {
  const PTrunc none{1, 0, 0.0};  // TruncThresh(0.0)
  per_orth_right(A, L, none, ws, ws.cap, err);
  per_orth_left(A, L, tr, ws, ws.cap, err);
}

// ------------------------------------------------------------------------------------------------
// op: Kronecker product of two B~ trains (also squares the closing bond), compress, normalize_eachmatrix
// ------------------------------------------------------------------------------------------------
PER_FN void per_op(const PerNode& nd, const PerOp& op, PerWS& ws, int* err) {
  const int L = nd.L, q = nd.q;
  const PTT& A = nd.reg[op.a];
  const PTT& B = nd.reg[op.b];
  const PTT& O = nd.reg[op.o];
  PTT K = ws.K;
  K.X = op.nyo * q;
  int over = 0;
  for (int t = 0; t < L; ++t) {
    const int m1 = per_bl(A, t), n1 = per_br(A, t, L), m2 = per_bl(B, t), n2 = per_br(B, t, L);
    if ((size_t)m1 * m2 * n1 * n2 * K.X > (size_t)ws.cap) over = 1;
  }
  if (over) { if (PER_TID == 0) *err |= PER_ERR_BOND; return; }
  for (int t = 0; t < L; ++t) {
    const int m1 = per_bl(A, t), n1 = per_br(A, t, L), m2 = per_bl(B, t), n2 = per_br(B, t, L);
    const double* a = A.data + (size_t)t * A.stride;
    const double* b = B.data + (size_t)t * B.stride;
    const double* pyy = op.pyy + (size_t)t * op.pyy_ts;
    double* o = K.data + (size_t)t * K.stride;
    const int M = m1 * m2, N = n1 * n2;
    PER_FOR(idx, M * N * K.X) {
      const int mm = idx % M, nn = (idx / M) % N, yx = idx / (M * N);
      const int y = yx % op.nyo, x = yx / op.nyo;
      const int i1 = mm % m1, i2 = mm / m1, j1 = nn % n1, j2 = nn / n1;
      double acc = 0;
      for (int y2 = 0; y2 < op.ny2; ++y2) {
        const double bv = b[i2 + (size_t)m2 * (j2 + (size_t)n2 * (y2 + (size_t)op.ny2 * x))];
        if (bv == 0.0) continue;
        double s = 0;
        for (int y1 = 0; y1 < op.ny1; ++y1)
          s += pyy[y + op.nyo * (y1 + op.ny1 * (y2 + op.ny2 * x))] * a[i1 + (size_t)m1 * (j1 + (size_t)n1 * (y1 + (size_t)op.ny1 * x))];
        acc += s * bv;
      }
      o[idx] = acc;
    }
    if (PER_TID == 0) K.bonds[t] = M;
  }
  if (PER_TID == 0) *K.ls = *A.ls + *B.ls;
  per_close(K, L);
  per_compress(K, L, nd.tr, ws, err);
  per_normalize_eachmatrix(K, L, ws, err);
  // the result must fit a bond-dmax register
  int big = 0;
  for (int t = 0; t < L; ++t) big = big || K.bonds[t] > nd.dmax;
  if (big) { if (PER_TID == 0) *err |= PER_ERR_BOND; PER_SYNC(); return; }
  per_copy_tt(O, K, L, err);
}

// site t of f_bp_partial (src/recursive_bp_factor.jl:89-101): Bt[m,n,x,xj,x'] = sum_y W[x',x,xj,y] phi_t[x] C_t[m,n,y,x]
PER_FN void per_build_B(const PTT& C, int t, int L, int ny, int q, int qj, const double* W, const double* phi, double* Bt) {
  const int m = per_bl(C, t), n = per_br(C, t, L);
  const double* c = C.data + (size_t)t * C.stride;
  const int mn = m * n;
  PER_FOR(idx, mn * q * qj * q) {
    const int e = idx % mn, x = (idx / mn) % q, xj = (idx / (mn * q)) % qj, xn = idx / (mn * q * qj);
    double acc = 0;
    for (int y = 0; y < ny; ++y) acc += W[xn + q * (x + q * (xj + qj * y))] * c[e + (size_t)mn * (y + (size_t)ny * x)];
    Bt[idx] = acc * phi[(size_t)t * q + x];
  }
  PER_SYNC();
}

// outgoing message j: f_bp_partial_ij -> mpem2(::PeriodicMPEM3) (src/mpems.jl:123-155) -> compress!(left) ->
// normalize_eachmatrix! -> normalize! ; returns log z_{i->j} in every thread
PER_FN double per_finalize(const PerNode& nd, int j, PerWS& ws, int* err) {
  const int L = nd.L, q = nd.q, qj = nd.qn[j];
  const PTT& C = nd.reg[nd.dest_reg[j]];
  const double* Wt = nd.Wp[j];
  PTT K = ws.K;
  K.X = q * qj;
  if (PER_TID == 0) *K.ls = *C.ls;
  PER_SYNC();
  // Bnew lives in ws.Tmp ([m', n, x, xj, x'], m' = carried bond); the raw site of f_bp_partial is built in ws.Mb first
  int mcur = per_bl(C, 0);
  per_build_B(C, 0, L, nd.nyc, q, qj, Wt, nd.phi, ws.Tmp);  // (site 0: time slice 0 of the table)
  for (int t = 0; t < L; ++t) {
    const int n = per_br(C, t, L);
    const int R = q * qj * mcur, Cc = n * q;
    if ((size_t)R * Cc > (size_t)ws.cap) { if (PER_TID == 0) *err |= PER_ERR_BOND; PER_SYNC(); return 0.0; }
    // M[(x, xj, m), (n, x')] = Bnew[m, n, x, xj, x']
    PER_FOR(idx, R * Cc) {
      const int mm = idx % mcur, nn = (idx / mcur) % n, x = (idx / (mcur * n)) % q, xj = (idx / (mcur * n * q)) % qj,
                xn = idx / (mcur * n * q * qj);
      ws.Mb[(x + q * (xj + qj * mm)) + (size_t)R * (nn + n * xn)] = ws.Tmp[idx];
    }
    PER_SYNC();
    const PTrunc none{1, 0, 0.0};
    const int r = per_svd_factor(ws.Mb, R, Cc, 1, none, K.ls, ws, err);  // Lf = U (R x r), Rf = lam Vt (r x n q)
    if ((size_t)mcur * r * q * qj > (size_t)K.stride) { if (PER_TID == 0) *err |= PER_ERR_BOND; PER_SYNC(); return 0.0; }
    double* Kt = K.data + (size_t)t * K.stride;
    // C_t[m, k, x, xj] = U[(x, xj, m), k]
    PER_FOR(idx, mcur * r * q * qj) {
      const int mm = idx % mcur, k = (idx / mcur) % r, x = (idx / (mcur * r)) % q, xj = idx / (mcur * r * q);
      Kt[idx] = ws.Lf[(x + q * (xj + qj * mm)) + (size_t)R * k];
    }
    if (PER_TID == 0) { K.bonds[t] = mcur; }
    PER_SYNC();
    if (t < L - 1) {
      // Bnew_{t+1}[k, n2, x', xj', x''] = sum_l (lam Vt)[k, (l, x')] B_{t+1}[l, n2, x', xj', x'']
      per_build_B(C, t + 1, L, nd.nyc, q, qj, Wt + (size_t)(t + 1) * nd.w_ts, nd.phi, ws.Mb);
      const int n2 = per_br(C, t + 1, L);
      if ((size_t)r * n2 * q * qj * q > (size_t)ws.cap) { if (PER_TID == 0) *err |= PER_ERR_BOND; PER_SYNC(); return 0.0; }
      PER_FOR(idx, r * n2 * q * qj * q) {
        const int k = idx % r, c2 = (idx / r) % n2, x = (idx / (r * n2)) % q, rest = idx / (r * n2 * q);  // rest = (xj', x'')
        double acc = 0;
        for (int l = 0; l < n; ++l)
          acc += ws.Rf[k + (size_t)r * (l + n * x)] * ws.Mb[l + (size_t)n * (c2 + (size_t)n2 * (x + (size_t)q * rest))];
        ws.Tmp[idx] = acc;
      }
      PER_SYNC();
      mcur = r;
    } else {
      // close the ring: C_0[k, n0, x0, xj0] <- sum_l (lam Vt)[k, (l, x0)] C_0[l, n0, x0, xj0]   (l = old left bond of site 0)
      double* K0 = K.data;
      const int l0 = K.bonds[0];                       // == n (right bond of the last site of C == left bond of site 0)
      const int n0 = (L == 1) ? r : K.bonds[1 % L];    // right bond of site 0 (L == 1: just written as r)
      if ((size_t)r * n0 * q * qj > (size_t)K.stride || (size_t)l0 * n0 * q * qj > (size_t)ws.cap) {
        if (PER_TID == 0) *err |= PER_ERR_BOND;
        PER_SYNC();
        return 0.0;
      }
      PER_FOR(i, l0 * n0 * q * qj) ws.Mb[i] = K0[i];
      PER_SYNC();
      PER_FOR(idx, r * n0 * q * qj) {
        const int k = idx % r, c2 = (idx / r) % n0, x = (idx / (r * n0)) % q, xj = idx / (r * n0 * q);
        double acc = 0;
        for (int l = 0; l < l0; ++l)
          acc += ws.Rf[k + (size_t)r * (l + l0 * x)] * ws.Mb[l + (size_t)l0 * (c2 + (size_t)n0 * (x + (size_t)q * xj))];
        K0[idx] = acc;
      }
      PER_SYNC();
      if (PER_TID == 0) K.bonds[0] = r;
    }
    // right bond of site t = r: recorded as the left bond of site t+1 (t < L-1) or of site 0 (above)
    if (t < L - 1) { if (PER_TID == 0) K.bonds[t + 1] = r; PER_SYNC(); }
  }
  per_close(K, L);
  per_orth_right(K, L, nd.tr, ws, K.stride < ws.cap ? K.stride : ws.cap, err);  // compress!(...; is_orthogonal = :left)
  per_normalize_eachmatrix(K, L, ws, err);
  const double lz = *K.ls + per_ring_lognorm(K, L, ws, err);  // normalize!: log z_{i->j}
  PER_SYNC();
  const PTT& out = nd.msg_out[j];
  double lnorm = lz;  // log-normalisation of what K holds
  if (nd.damp > 0.0) {
    // set_msg!: mu <- mu_new + damp/(1-damp) mu_old (block-diagonal ring sum, _compose), compress!, normalize!
    const PTT& old = out;
    const int X = K.X;
    const double fa = exp((*K.ls - lz) / L), fb = exp(*old.ls / L), coef = nd.damp / (1.0 - nd.damp);
    int over = 0;
    for (int t = 0; t < L; ++t)
      over = over || (size_t)(per_bl(K, t) + per_bl(old, t)) * (per_br(K, t, L) + per_br(old, t, L)) * X > (size_t)ws.cap;
    if (over) { if (PER_TID == 0) *err |= PER_ERR_BOND; PER_SYNC(); return lz; }
    for (int t = 0; t < L; ++t) {
      const int ma = per_bl(K, t), na = per_br(K, t, L), mb = per_bl(old, t), nb = per_br(old, t, L);
      const int Mr = ma + mb, Nc = na + nb;
      double* Kt = K.data + (size_t)t * K.stride;
      const double* Ot = old.data + (size_t)t * old.stride;
      const double fo = (t == 0) ? fb * coef : fb;
      PER_FOR(idx, Mr * Nc * X) {
        const int r = idx % Mr, c = (idx / Mr) % Nc, x = idx / (Mr * Nc);
        double v = 0.0;
        if (r < ma && c < na) v = Kt[r + (size_t)ma * (c + (size_t)na * x)] * fa;
        else if (r >= ma && c >= na) v = Ot[(r - ma) + (size_t)mb * ((c - na) + (size_t)nb * x)] * fo;
        ws.Tmp[idx] = v;
      }
      PER_SYNC();
      PER_FOR(i, Mr * Nc * X) Kt[i] = ws.Tmp[i];
      PER_SYNC();
    }
    if (PER_TID == 0) {
      for (int t = 0; t < L; ++t) K.bonds[t] += old.bonds[t];
      *K.ls = 0.0;
    }
    per_close(K, L);
    per_compress(K, L, nd.tr, ws, err);
    lnorm = *K.ls + per_ring_lognorm(K, L, ws, err);
    PER_SYNC();
  }
  int big = 0;
  for (int t = 0; t < L; ++t) big = big || K.bonds[t] > nd.dmax;
  if (big) { if (PER_TID == 0) *err |= PER_ERR_BOND; PER_SYNC(); return lz; }
  per_copy_tt(out, K, L, err);
  if (PER_TID == 0) *out.ls = *K.ls - lnorm;  // stored normalised: exp(ls) trace prod sum_x A = 1
  PER_SYNC();
  return lz;
}

// belief of node i and log z_i from the transfer matrices T_t[(m,x),(n,x')] = sum_y Wd[x',x,y] phi_t[x] full_t[m,n,y,x]
// (bp.b[i] = marginalize(mpem2(f_bp_partial_i(full))), logz = normalize!(bp.b[i]); src/recursive_bp_factor.jl:160-162)
PER_FN double per_belief(const PerNode& nd, PerWS& ws, int* err) {
  const int L = nd.L, q = nd.q, ny = nd.nyz;
  const PTT& Fu = nd.reg[nd.full_reg];
  const size_t msz = (size_t)ws.wcap * ws.wcap;
  double* Tm = ws.tm + 3 * msz;         // L transfer matrices
  double* Pre = ws.tm + (3 + L) * msz;  // L prefix products  Pre_t = T_0 ... T_{t-1}   (Pre_0 = I)
  double* Suf = ws.tm;                  // running suffix (2 buffers) + scratch
  const int d0 = per_bl(Fu, 0) * q;
  for (int t = 0; t < L; ++t) {
    const int m = per_bl(Fu, t), n = per_br(Fu, t, L);
    const double* a = Fu.data + (size_t)t * Fu.stride;
    const double* Wd = nd.Wd + (size_t)t * nd.wd_ts;
    double* T = Tm + t * msz;
    const int mq = m * q;
    PER_FOR(idx, mq * n * q) {
      const int mm = idx % m, x = (idx / m) % q, nn = (idx / mq) % n, xn = idx / (mq * n);
      double acc = 0;
      for (int y = 0; y < ny; ++y) acc += Wd[xn + q * (x + q * y)] * a[mm + (size_t)m * (nn + (size_t)n * (y + (size_t)ny * x))];
      T[(mm + m * x) + (size_t)mq * (nn + n * xn)] = acc * nd.phi[(size_t)t * q + x];
    }
  }
  PER_SYNC();
  // prefixes (scaled; only ratios matter for the marginals, the logs are summed for log z)
  double lacc = 0.0;
  PER_FOR(i, d0 * d0) Pre[i] = (i % d0 == i / d0) ? 1.0 : 0.0;
  PER_SYNC();
  for (int t = 0; t < L - 1; ++t) {
    const int k = per_bl(Fu, t) * q, n = per_br(Fu, t, L) * q;
    lacc += per_matmul_scaled(Pre + (t + 1) * msz, Pre + t * msz, Tm + t * msz, d0, k, n, ws.red, err);
  }
  // suffixes from the right: Suf_t = T_t ... T_{L-1}  (dims (bl_t q) x d0);  marg_t(x) ~ sum_m (Suf_t Pre_t)[(m,x),(m,x)]
  double* cur = Suf;
  double* nxt = Suf + msz;
  double* prod = Suf + 2 * msz;
  double logz = 0.0;
  for (int t = L - 1; t >= 0; --t) {
    const int k = per_bl(Fu, t) * q, n = per_br(Fu, t, L) * q;  // T_t is k x n
    if (t == L - 1) {
      PER_FOR(i, k * n) nxt[i] = Tm[t * msz + i];  // n == d0
      PER_SYNC();
    } else {
      per_matmul_scaled(nxt, Tm + t * msz, cur, k, n, d0, ws.red, err);
    }
    double* sw = cur; cur = nxt; nxt = sw;
    // prod = Suf_t (k x d0) * Pre_t (d0 x k): only its diagonal is needed
    const int m = per_bl(Fu, t);
    PER_FOR(i, k) {
      double acc = 0;
      for (int l = 0; l < d0; ++l) acc += cur[i + (size_t)k * l] * Pre[t * msz + l + (size_t)d0 * i];
      prod[i] = acc;
    }
    PER_SYNC();
    double tot = 0;
    for (int i = 0; i < k; ++i) tot += prod[i];
    if (!(tot > 0.0) && PER_TID == 0) *err |= PER_ERR_NAN;
    PER_FOR(x, q) {
      double s = 0;
      for (int mm = 0; mm < m; ++mm) s += prod[mm + m * x];
      nd.marg[(size_t)t * q + x] = s / tot;
    }
    if (t == L - 1) logz = log(tot) + lacc;  // trace(Pre_{L-1} T_{L-1}) with Pre scaled by exp(-lacc)
    if (nd.tv) {  // keep every suffix for the two-time marginals below
      double* Sa = ws.tm + (3 + 2 * (size_t)L + t) * msz;
      PER_FOR(i, k * d0) Sa[i] = cur[i];
    }
    PER_SYNC();
  }
  if (nd.tv) {
    // b(x_t, x_u) ~ trace(Pre_t P_{x_t} T_t ... T_{u-1} P_{x_u} Suf_u),  P_x = projector on the rows (m, x).
    // V_x = Pre_t P_x T_t T_{t+1} ... is carried for all x_t together (common rescaling keeps their relative weights).
    const int q2c = nd.tv_q2cap;
    double* Sall = ws.tm + (3 + 2 * (size_t)L) * msz;
    double* Va = ws.tm + (3 + 3 * (size_t)L) * msz;
    double* Vb = Va + (size_t)q * msz;
    PER_FOR(i, L * L * q2c) nd.tv[i] = 0.0;
    PER_SYNC();
    for (int t = 0; t < L - 1; ++t) {
      const int mt = per_bl(Fu, t), kt = mt * q, nt = per_br(Fu, t, L) * q;
      // V_x (d0 x nt) = Pre_t[:, (m, x)] * T_t[(m, x), :]
      PER_FOR(idx, q * d0 * nt) {
        const int x = idx / (d0 * nt), a = idx % d0, c = (idx / d0) % nt;
        double acc = 0;
        for (int mm = 0; mm < mt; ++mm) acc += Pre[t * msz + a + (size_t)d0 * (mm + mt * x)] * Tm[t * msz + (mm + mt * x) + (size_t)kt * c];
        Va[x * msz + a + (size_t)d0 * c] = acc;
      }
      PER_SYNC();
      double* Vc = Va;
      double* Vn = Vb;
      for (int u = t + 1; u < L && u - t <= nd.tv_maxdist; ++u) {
        const int mu = per_bl(Fu, u), ku = mu * q, nu = per_br(Fu, u, L) * q;  // V_x is d0 x ku here
        const double* Su = Sall + u * msz;                                     // ku x d0
        double* o = nd.tv + ((size_t)t * L + u) * q2c;
        PER_FOR(xx, q * q) {
          const int xt = xx % q, xu = xx / q;
          double acc = 0;
          for (int mm = 0; mm < mu; ++mm)
            for (int a = 0; a < d0; ++a) acc += Vc[xt * msz + a + (size_t)d0 * (mm + mu * xu)] * Su[(mm + mu * xu) + (size_t)ku * a];
          o[xx] = acc;
        }
        PER_SYNC();
        double tot = 0;
        for (int xx = 0; xx < q * q; ++xx) tot += o[xx];
        PER_SYNC();
        if (!(tot > 0.0) && PER_TID == 0) *err |= PER_ERR_NAN;
        PER_FOR(xx, q * q) o[xx] /= tot;
        if (u + 1 < L && u + 1 - t <= nd.tv_maxdist) {
          // V_x <- V_x T_u for every x, rescaled by ONE common factor
          PER_FOR(idx, q * d0 * nu) {
            const int x = idx / (d0 * nu), a = idx % d0, c = (idx / d0) % nu;
            double acc = 0;
            for (int l = 0; l < ku; ++l) acc += Vc[x * msz + a + (size_t)d0 * l] * Tm[u * msz + l + (size_t)ku * c];
            Vn[x * msz + a + (size_t)d0 * c] = acc;
          }
          PER_SYNC();
          double mx = 0.0;
          for (int x = 0; x < q; ++x) mx = fmax(mx, per_maxabs(Vn + x * msz, d0 * nu, ws.red, err));
          if (mx > 0.0) {
            const double sc = 1.0 / mx;
            PER_FOR(idx, q * d0 * nu) Vn[(idx / (d0 * nu)) * msz + idx % (d0 * nu)] *= sc;
          }
          PER_SYNC();
          double* sw = Vc; Vc = Vn; Vn = sw;
        }
      }
      PER_SYNC();
    }
  }
  return logz + *Fu.ls;
}

// ------------------------------------------------------------------------------------------------
// the node update (onebpiter!, src/recursive_bp_factor.jl:146-165, on ring trains; damp = 0)
// ------------------------------------------------------------------------------------------------
// the node's own error word (ws.ib[4]) is read CTA-uniformly: barrier, read, barrier
PER_FN int per_failed(const int* err) {
  PER_SYNC();
  const int e = *err;
  PER_SYNC();
  return e;
}
PER_FN void per_report(const PerNode& nd, const int* err) {
  if (PER_TID == 0 && *err) {
#ifdef PER_HOST
    *nd.err |= *err;
#else
    atomicOr(nd.err, *err);
#endif
  }
}

PER_FN void per_node_update(const PerNode& nd) {
  PerWS ws = nd.ws;
  const int L = nd.L, q = nd.q, z = nd.z;
  int* err = ws.ib + 4;
  if (PER_TID == 0) *err = 0;
  PER_SYNC();
  // B~_k = sum_{x_k} Pxy[y,x_k,x_i] psi[x_i,x_k] mu_{k->i}[m,n,x_k,x_i]      (:108-115)
  for (int k = 0; k < z; ++k) {
    const PTT& in = nd.msg_in[k];
    const PTT& out = nd.reg[nd.src_reg[k]];
    const int qk = nd.qn[k], ny1 = nd.ny1;
    for (int t = 0; t < L; ++t) {
      const int mn = per_bl(in, t) * per_br(in, t, L);
      const double* A = in.data + (size_t)t * in.stride;
      double* O = out.data + (size_t)t * out.stride;
      const double* psi = nd.psi[k] + (size_t)t * q * qk;
      const double* pxy = nd.pxy[k] + (size_t)t * nd.pxy_ts;
      PER_FOR(idx, mn * ny1 * q) {
        const int e = idx % mn, y = (idx / mn) % ny1, x = idx / (mn * ny1);
        double acc = 0.0;
        for (int xk = 0; xk < qk; ++xk) acc += pxy[y + ny1 * (xk + qk * x)] * psi[x + q * xk] * A[e + (size_t)mn * (xk + qk * x)];
        O[idx] = acc;
      }
    }
    PER_FOR(t, L + 1) out.bonds[t] = in.bonds[t % L];
    if (PER_TID == 0) *out.ls = *in.ls;
  }
  {
    const PTT& out = nd.reg[nd.init_reg];
    const int n = nd.ny0 * q;
    PER_FOR(idx, L * n) out.data[(size_t)(idx / n) * out.stride + idx % n] = nd.minit[(size_t)(idx / n) * nd.minit_ts + idx % n];
    PER_FOR(t, L + 1) out.bonds[t] = 1;
    if (PER_TID == 0) *out.ls = 0.0;
  }
  PER_SYNC();
  for (int o = 0; o < nd.nops; ++o) {
    per_op(nd, nd.ops[o], ws, err);
    if (per_failed(err)) { per_report(nd, err); return; }
  }
  double sumlz = 0.0;
  for (int j = 0; j < z; ++j) {
    const double lz = per_finalize(nd, j, ws, err);
    if (per_failed(err)) { per_report(nd, err); return; }
    if (PER_TID == 0) nd.logzij[j] = lz;
    sumlz += lz;
  }
  const double lzi = per_belief(nd, ws, err);
  if (PER_TID == 0) {
    *nd.logzi = lzi;
    *nd.f = (0.5 * z - 1.0) * lzi - 0.5 * sumlz;
  }
  per_failed(err);
  per_report(nd, err);
}

// ------------------------------------------------------------------------------------------------
// pair belief b_ij^t(x_i, x_j) and log z_ij of one edge from the two ring messages (pair_belief, src/bp_core.jl:95-109 on
// PeriodicMPEM2s): the bond-d^2 product train is never compressed by the reference, so its marginals and normalisation are
// read off the transfer matrices  TM_t[(a,b),(a',b')] = sum_{xi,xj} A_t[a,a',xi,xj] B_t[b,b',xj,xi] psi_t[xi,xj].
// ------------------------------------------------------------------------------------------------
struct PerPair {
  PTT a, b;           // mu_ij [a,a',xi,xj], mu_ji [b,b',xj,xi]
  const double* psi;  // [t][xi + qi*xj]
  int qi, qj, L;
  double* out;        // [t][xi + qi*xj]
  double* logz;
  double* tm;         // (2L + 3) * wcap^2
  int wcap;           // dmax^2
  double* red;        // PER_MAXW + 8
  int* err;
};

PER_FN void per_pair_belief(const PerPair& pj) {
  const int L = pj.L, qi = pj.qi, qj = pj.qj;
  const PTT& A = pj.a;
  const PTT& B = pj.b;
  const size_t msz = (size_t)pj.wcap * pj.wcap;
  double* Tm = pj.tm + 3 * msz;
  double* Pre = pj.tm + (3 + L) * msz;
  double* cur = pj.tm;
  double* nxt = pj.tm + msz;
  double* Q = pj.tm + 2 * msz;
  int lerr = 0;
  const int d0 = per_bl(A, 0) * per_bl(B, 0);
  for (int t = 0; t < L; ++t) {
    const int ma = per_bl(A, t), na = per_br(A, t, L), mb = per_bl(B, t), nb = per_br(B, t, L);
    const double* a = A.data + (size_t)t * A.stride;
    const double* b = B.data + (size_t)t * B.stride;
    const double* psi = pj.psi + (size_t)t * qi * qj;
    double* T = Tm + t * msz;
    const int K = ma * mb, N = na * nb;
    PER_FOR(idx, K * N) {
      const int r = idx % K, c = idx / K;
      const int ia = r % ma, ib = r / ma, ja = c % na, jb = c / na;
      double acc = 0;
      for (int xi = 0; xi < qi; ++xi)
        for (int xj = 0; xj < qj; ++xj)
          acc += a[ia + (size_t)ma * (ja + (size_t)na * (xi + qi * xj))] * b[ib + (size_t)mb * (jb + (size_t)nb * (xj + qj * xi))] *
                 psi[xi + qi * xj];
      T[idx] = acc;
    }
  }
  PER_SYNC();
  double lacc = 0.0;
  PER_FOR(i, d0 * d0) Pre[i] = (i % d0 == i / d0) ? 1.0 : 0.0;
  PER_SYNC();
  for (int t = 0; t < L - 1; ++t) {
    const int k = per_bl(A, t) * per_bl(B, t), n = per_br(A, t, L) * per_br(B, t, L);
    lacc += per_matmul_scaled(Pre + (t + 1) * msz, Pre + t * msz, Tm + t * msz, d0, k, n, pj.red, &lerr);
  }
  // cur = Suf_{t+1} = TM_{t+1} ... TM_{L-1}  (n_t x d0), Suf_L = I
  PER_FOR(i, d0 * d0) cur[i] = (i % d0 == i / d0) ? 1.0 : 0.0;
  PER_SYNC();
  for (int t = L - 1; t >= 0; --t) {
    const int ma = per_bl(A, t), na = per_br(A, t, L), mb = per_bl(B, t), nb = per_br(B, t, L);
    const int K = ma * mb, N = na * nb;
    // Q = Suf_{t+1} (N x d0) * Pre_t (d0 x K)
    per_matmul_scaled(Q, cur, Pre + t * msz, N, d0, K, pj.red, &lerr);
    const double* a = A.data + (size_t)t * A.stride;
    const double* b = B.data + (size_t)t * B.stride;
    const double* psi = pj.psi + (size_t)t * qi * qj;
    double* o = pj.out + (size_t)t * qi * qj;
    PER_FOR(xx, qi * qj) {
      const int xi = xx % qi, xj = xx / qi;
      double acc = 0;
      for (int c = 0; c < N; ++c) {
        const int ja = c % na, jb = c / na;
        for (int r = 0; r < K; ++r) {
          const int ia = r % ma, ib = r / ma;
          acc += a[ia + (size_t)ma * (ja + (size_t)na * (xi + qi * xj))] * b[ib + (size_t)mb * (jb + (size_t)nb * (xj + qj * xi))] *
                 Q[c + (size_t)N * r];
        }
      }
      o[xx] = acc * psi[xx];
    }
    PER_SYNC();
    double tot = 0;
    for (int xx = 0; xx < qi * qj; ++xx) tot += o[xx];
    PER_SYNC();
    if (!(tot > 0.0)) lerr |= PER_ERR_NAN;
    PER_FOR(xx, qi * qj) o[xx] /= tot;
    if (t == L - 1 && PER_TID == 0) {
      // trace(Pre_{L-1} TM_{L-1}) from the diagonal of TM_{L-1} (K x d0) * Pre_{L-1} (d0 x K)
      double tr = 0;
      const double* T = Tm + t * msz;
      for (int r = 0; r < K; ++r)
        for (int l = 0; l < d0; ++l) tr += T[r + (size_t)K * l] * Pre[t * msz + l + (size_t)d0 * r];
      if (!(tr > 0.0)) lerr |= PER_ERR_NAN;
      *pj.logz = log(tr) + lacc + *A.ls + *B.ls;
    }
    if (t > 0) {
      per_matmul_scaled(nxt, Tm + t * msz, cur, K, N, d0, pj.red, &lerr);
      double* sw = cur; cur = nxt; nxt = sw;
    }
  }
  if (lerr && PER_TID == 0) {
#ifdef PER_HOST
    *pj.err |= lerr;
#else
    atomicOr(pj.err, lerr);
#endif
  }
}

}  // namespace mpbp_per
