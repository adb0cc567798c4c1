// TEST INFRASTRUCTURE ONLY.  One-thread CPU emulation of the periodic node-update kernel: compiles
// matrixproductbp.jl_b200/csrc/periodic.cuh UNCHANGED with the host shim (PER_HOST: one thread, barriers are no-ops) so that
// the CPU test tier can check the kernel's arithmetic and indexing against oracle/periodic.py without a GPU.  Never loaded by
// the product (tests/test_periodic_host_emul.py builds it with g++ into tests/host_emul/_build/).
#define PER_HOST 1
#include "../../matrixproductbp.jl_b200/csrc/periodic_plan.h"

#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

using namespace mpbp_per;

extern "C" int per_host_node_update(int z, int q, const int* qn, int T, int dmax, const int* ny, int nt, const double* pxy, int npairs,
                                    const int* pd1, const int* pd2, const double* pyy, const double* w, const double* wd,
                                    const double* minit, const double* phi, const double* psi, int trunc_kind, int trunc_d,
                                    double trunc_eps, double damp, int sstride, const int* in_bonds, const double* in_data, const double* in_ls,
                                    int* out_bonds, double* out_data, double* out_ls, double* marg, double* logzi, double* logzij,
                                    double* f, double* tv, int tv_maxdist, int tv_q2cap, int* err) {
  const int L = T + 1;
  const bool td = nt > 1;
  std::vector<size_t> pxy_off(z), w_off(z);
  size_t tot = 0;
  for (int k = 0; k < z; ++k) { pxy_off[k] = tot; tot += (size_t)ny[1] * qn[k] * q; }
  const size_t pxy_ts = td ? tot : 0;
  std::map<std::pair<int, int>, std::pair<size_t, size_t>> pyym;
  size_t off = 0;
  for (int p = 0; p < npairs; ++p) {
    const size_t sz = (size_t)ny[pd1[p] + pd2[p]] * ny[pd1[p]] * ny[pd2[p]] * q;
    pyym[{pd1[p], pd2[p]}] = {off, td ? sz : 0};
    off += sz * nt;
  }
  tot = 0;
  for (int j = 0; j < z; ++j) { w_off[j] = tot; tot += (size_t)q * q * qn[j] * ny[z > 0 ? z - 1 : 0]; }
  const size_t w_ts = td ? tot : 0;
  const size_t mis = (size_t)ny[0] * q;
  std::vector<double> mi((size_t)L * mis);
  for (int t = 0; t < L; ++t) memcpy(mi.data() + t * mis, minit + (td ? t * mis : 0), mis * 8);
  PerClassView c;
  c.z = z;
  c.q = q;
  c.qn = qn;
  c.ny = ny;
  c.pxy = pxy;
  c.pxy_off = pxy_off.data();
  c.pxy_ts = pxy_ts;
  c.pyy = [&](int d1, int d2, const double** p, size_t* ts) {
    auto it = pyym.find({d1, d2});
    if (it == pyym.end()) return false;
    *p = pyy + it->second.first;
    *ts = it->second.second;
    return true;
  };
  c.w = w;
  c.w_off = w_off.data();
  c.w_ts = w_ts;
  c.wd = wd;
  c.wd_ts = td ? (size_t)q * q * ny[z] : 0;
  c.minit = mi.data();
  c.minit_ts = mis;
  std::vector<void*> blocks;
  auto take = [&](size_t bytes) -> void* {
    void* p = calloc(bytes + 256, 1);
    blocks.push_back(p);
    return p;
  };
  PerNode nd;
  memset((void*)&nd, 0, sizeof nd);
  int rc = 0;
  if (!per_plan_node(c, L, dmax, take, nd)) rc = 2;
  nd.tr = PTrunc{trunc_kind, trunc_d, trunc_eps};
  nd.damp = damp;
  std::vector<std::vector<int>> ib(z, std::vector<int>(L + 1)), ob(z, std::vector<int>(L + 1));
  std::vector<double> ils(z), ols(z);
  size_t psi_off = 0;
  for (int k = 0; k < z; ++k) {
    for (int t = 0; t <= L; ++t) ib[k][t] = in_bonds[k * (L + 1) + t];
    ils[k] = in_ls[k];
    nd.msg_in[k] = PTT{const_cast<double*>(in_data) + (size_t)k * L * sstride, ib[k].data(), &ils[k], sstride, qn[k] * q};
    for (int t = 0; t <= L; ++t) ob[k][t] = out_bonds[k * (L + 1) + t];  // (damping reads the old message from the out slot)
    ols[k] = out_ls[k];
    nd.msg_out[k] = PTT{out_data + (size_t)k * L * sstride, ob[k].data(), &ols[k], sstride, q * qn[k]};
    nd.psi[k] = psi + psi_off;
    psi_off += (size_t)L * q * qn[k];
  }
  nd.phi = phi;
  nd.marg = marg;
  nd.logzi = logzi;
  nd.logzij = logzij;
  nd.f = f;
  nd.tv = tv;
  nd.tv_maxdist = tv_maxdist;
  nd.tv_q2cap = tv_q2cap;
  *err = 0;
  nd.err = err;
  if (rc == 0) per_node_update(nd);
  for (int k = 0; k < z; ++k) {
    for (int t = 0; t <= L; ++t) out_bonds[k * (L + 1) + t] = ob[k][t];
    out_ls[k] = ols[k];
  }
  for (void* p : blocks) free(p);
  return rc;
}

// the truncated-SVD building block alone: M (R x C column-major) -> Lf (R x r), Rf (r x C), returns r
extern "C" int per_host_svd(const double* M, int R, int C, int wl, int trunc_kind, int trunc_d, double trunc_eps, double* Lf, double* Rf,
                            double* ls, int* err) {
  PerWS ws;
  memset((void*)&ws, 0, sizeof ws);
  const size_t cap = (size_t)R * C, wc = (size_t)std::min(R, C);
  std::vector<double> G(cap), lf(cap), rf(cap), W(wc * wc), sig(2 * wc), red(PER_MAXW + 8);
  std::vector<int> perm(wc), ibuf(8);
  ws.cap = (int)cap;
  ws.wcap = (int)wc;
  ws.G = G.data();
  ws.Lf = lf.data();
  ws.Rf = rf.data();
  ws.W = W.data();
  ws.sig = sig.data();
  ws.perm = perm.data();
  ws.red = red.data();
  ws.ib = ibuf.data();
  *err = 0;
  *ls = 0.0;
  const int r = per_svd_factor(M, R, C, wl, PTrunc{trunc_kind, trunc_d, trunc_eps}, ls, ws, err);
  memcpy(Lf, lf.data(), sizeof(double) * R * r);
  memcpy(Rf, rf.data(), sizeof(double) * r * C);
  return r;
}

// pair belief of one edge from two ring messages given in slot format
extern "C" int per_host_pair_belief(int qi, int qj, int T, int dmax, int sstride, const int* bonds_a, const double* data_a, double ls_a,
                                    const int* bonds_b, const double* data_b, double ls_b, const double* psi, double* out,
                                    double* logz, int* err) {
  const int L = T + 1, D2 = dmax * dmax;
  std::vector<int> ba(bonds_a, bonds_a + L + 1), bb(bonds_b, bonds_b + L + 1);
  std::vector<double> tm((size_t)(2 * L + 3) * D2 * D2), red(PER_MAXW + 8);
  PerPair pj;
  memset((void*)&pj, 0, sizeof pj);
  pj.a = PTT{const_cast<double*>(data_a), ba.data(), &ls_a, sstride, qi * qj};
  pj.b = PTT{const_cast<double*>(data_b), bb.data(), &ls_b, sstride, qi * qj};
  pj.psi = psi;
  pj.qi = qi;
  pj.qj = qj;
  pj.L = L;
  pj.out = out;
  pj.logz = logz;
  pj.tm = tm.data();
  pj.wcap = D2;
  pj.red = red.data();
  *err = 0;
  pj.err = err;
  per_pair_belief(pj);
  return 0;
}
