// periodic_plan.h -- host-side planning of ONE periodic node update (TT registers, the 3z-2 cavity ops in dependency
// order, scratch sizing) for the kernel in periodic.cuh.  Plain C++: shared by engine.cu (device pointers, arena
// allocator) and by the CPU emulation harness of the test tier (tests/host_emul, host pointers, malloc).
// Cavity order = CavityTools.cavity as used at src/recursive_bp_factor.jl:140 (prefix products, full, suffix products,
// all-but-one), identical to the open-boundary planner (engine.cu: build_plan).
#pragma once
#include "periodic.cuh"

#include <algorithm>
#include <functional>

namespace mpbp_per {

// device- (or host-) resident tables of one node class, in the layout of mpbp_add_node_class (include/mpbp.h)
struct PerClassView {
  int z = 0, q = 0;
  const int* qn = nullptr;  // z
  const int* ny = nullptr;  // z+1
  const double* pxy = nullptr;
  const size_t* pxy_off = nullptr;  // z
  size_t pxy_ts = 0;
  std::function<bool(int, int, const double**, size_t*)> pyy;  // (d1, d2) -> table, time stride
  const double* w = nullptr;
  const size_t* w_off = nullptr;  // z
  size_t w_ts = 0;
  const double* wd = nullptr;
  size_t wd_ts = 0;
  const double* minit = nullptr;
  size_t minit_ts = 0;
};

// scratch capacities of a node of this class
inline void per_caps(const PerClassView& c, int dmax, int* cap, int* wcap) {
  const int z = c.z, q = c.q;
  long long cp = 1, wc = (long long)dmax * q;
  int qjmax = 1;
  for (int k = 0; k < z; ++k) qjmax = std::max(qjmax, c.qn[k]);
  auto op = [&](int ca, int cb, int nyo) {
    const long long D = (long long)ca * cb;
    cp = std::max(cp, D * D * nyo * q);
    wc = std::max(wc, D);
  };
  if (z == 1) op(dmax, 1, c.ny[1]);
  if (z >= 2) {
    for (int k = 1; k < z; ++k) op(dmax, dmax, c.ny[k + 1]);
    op(dmax, 1, c.ny[z]);
    for (int k = z - 1; k >= 1; --k) op(dmax, k == z - 1 ? 1 : dmax, c.ny[z - k]);
    for (int k = 1; k < z; ++k) op(dmax, k == z - 1 ? 1 : dmax, c.ny[z - 1]);
  }
  cp = std::max(cp, (long long)dmax * dmax * q * q * q * qjmax);           // finalize: unfolding / carried site
  cp = std::max(cp, (long long)dmax * dmax * std::max(c.ny[1], c.ny[z]) * q);  // registers copied through the workspace
  cp = std::max(cp, 4LL * dmax * dmax * q * qjmax);                             // damping: block-diagonal sum of two messages
  wc = std::max(wc, 2LL * dmax);
  *cap = (int)std::min<long long>(cp, 0x7fffffff);
  *wcap = (int)wc;
}

// take(bytes) returns 256-byte aligned storage or nullptr.  Fills everything of `nd` that depends on the class only
// (registers, ops, tables, scratch); the caller adds messages, reweightings, outputs and the truncation.
template <class Take>
bool per_plan_node(const PerClassView& c, int L, int dmax, Take&& take, PerNode& nd) {
  const int z = c.z, q = c.q;
  if (z > PER_MAXZ) return false;
  bool ok = true;
  nd.z = z;
  nd.q = q;
  nd.L = L;
  nd.dmax = dmax;
  nd.nreg = 0;
  nd.nops = 0;
  auto new_reg = [&](int capb, int ny) {
    PTT r;
    r.X = ny * q;
    r.stride = capb * capb * r.X;
    r.data = (double*)take(sizeof(double) * (size_t)L * r.stride);
    r.bonds = (int*)take(sizeof(int) * (L + 1));
    r.ls = (double*)take(sizeof(double));
    ok = ok && r.data && r.bonds && r.ls;
    nd.reg[nd.nreg] = r;
    return nd.nreg++;
  };
  auto add_op = [&](int a, int da, int b, int db) {
    PerOp op;
    op.a = a;
    op.b = b;
    op.ny1 = c.ny[da];
    op.ny2 = c.ny[db];
    op.nyo = c.ny[da + db];
    size_t ts = 0;
    op.pyy = nullptr;
    if (!c.pyy(da, db, &op.pyy, &ts)) ok = false;
    op.pyy_ts = (int)ts;
    op.o = new_reg(dmax, op.nyo);
    nd.ops[nd.nops++] = op;
    return op.o;
  };
  nd.ny1 = z > 0 ? c.ny[1] : 1;
  nd.pxy_ts = (int)c.pxy_ts;
  for (int k = 0; k < z; ++k) {
    nd.qn[k] = c.qn[k];
    nd.pxy[k] = c.pxy + c.pxy_off[k];
    nd.Wp[k] = c.w + c.w_off[k];
    nd.src_reg[k] = new_reg(dmax, c.ny[1]);
  }
  nd.w_ts = (int)c.w_ts;
  nd.nyc = z > 0 ? c.ny[z - 1] : 1;
  nd.minit = c.minit;
  nd.minit_ts = (int)c.minit_ts;
  nd.ny0 = c.ny[0];
  nd.init_reg = new_reg(1, c.ny[0]);
  nd.Wd = c.wd;
  nd.wd_ts = (int)c.wd_ts;
  nd.nyz = c.ny[z];
  if (z == 0) {
    nd.full_reg = nd.init_reg;
  } else if (z == 1) {
    nd.dest_reg[0] = nd.init_reg;
    nd.full_reg = add_op(nd.src_reg[0], 1, nd.init_reg, 0);
  } else {
    int p[PER_MAXZ], s[PER_MAXZ + 1];
    p[0] = nd.src_reg[0];
    for (int k = 1; k < z; ++k) p[k] = add_op(p[k - 1], k, nd.src_reg[k], 1);
    nd.full_reg = add_op(p[z - 1], z, nd.init_reg, 0);
    s[z] = nd.init_reg;
    for (int k = z - 1; k >= 1; --k) s[k] = add_op(nd.src_reg[k], 1, s[k + 1], z - 1 - k);
    for (int k = 1; k < z; ++k) nd.dest_reg[k] = add_op(p[k - 1], k, s[k + 1], z - 1 - k);
    nd.dest_reg[0] = s[1];
  }
  // scratch
  PerWS& ws = nd.ws;
  per_caps(c, dmax, &ws.cap, &ws.wcap);
  const size_t cap = (size_t)ws.cap, wc = (size_t)ws.wcap;
  ws.K.stride = ws.cap;
  ws.K.X = 1;
  ws.K.data = (double*)take(8 * cap * L);
  ws.K.bonds = (int*)take(4 * (L + 1));
  ws.K.ls = (double*)take(8);
  ws.G = (double*)take(8 * cap);
  ws.Lf = (double*)take(8 * cap);
  ws.Rf = (double*)take(8 * cap);
  ws.Tmp = (double*)take(8 * cap);
  ws.Mb = (double*)take(8 * cap);
  ws.W = (double*)take(8 * wc * wc);
  ws.sig = (double*)take(8 * 2 * wc);
  ws.perm = (int*)take(4 * wc);
  ws.red = (double*)take(8 * (PER_MAXW + 8));
  ws.ib = (int*)take(4 * 8);
  ws.tm = (double*)take(8 * (size_t)(3 * L + 3 + 2 * q) * wc * wc);
  ok = ok && ws.K.data && ws.K.bonds && ws.K.ls && ws.G && ws.Lf && ws.Rf && ws.Tmp && ws.Mb && ws.W && ws.sig && ws.perm &&
       ws.red && ws.ib && ws.tm;
  return ok;
}

}  // namespace mpbp_per
