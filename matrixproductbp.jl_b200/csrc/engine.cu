// engine.cu -- host runtime of the MPBP engine: device state, arena, cavity-DAG planner, launch sequencing,
// and the extern "C" boundary declared in include/mpbp.h.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/mpbp.h"
#include "kernels.cuh"
#include "periodic_plan.h"

using namespace mpbp;

static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CUDA_OK(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) return fail("CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

namespace {

constexpr int QR_NSPLIT_MAX = 8;

struct NodeClass {
  int z = 0, q = 0, nt = 1;
  bool generic = false;
  std::vector<int> qn, ny;
  double* d_pxy = nullptr;
  std::vector<size_t> pxy_off;
  size_t pxy_ts = 0;
  double* d_pyy = nullptr;
  std::map<std::pair<int, int>, std::pair<size_t, size_t>> pyy;  // (d1,d2) -> (offset, tstride)
  double* d_w = nullptr;
  std::vector<size_t> w_off;
  size_t w_ts = 0;
  double* d_wd = nullptr;
  size_t wd_ts = 0;
  double* d_minit = nullptr;
  size_t minit_ts = 0;
  std::vector<int> gen_ny;  // generic: joint neighbour states with neighbour j removed; gen_ny[z] = all neighbours
};

struct MsgStore {
  double* data = nullptr;
  int* bonds = nullptr;
  double* ls = nullptr;
};

struct Arena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  void* take(size_t bytes) {
    size_t a = (used + 255) & ~size_t(255);
    if (a + bytes > cap) return nullptr;
    used = a + bytes;
    return base + a;
  }
};

}  // namespace

struct mpbp_state {
  int device = 0;
  int64_t N = 0, E2 = 0;
  int T = 0, L = 1, dmax = 1, qmax = 1;
  int inf_k = 0;  // > 0: infinite (iid) graph: every node stands for a whole class and sees inf_deg[i] copies of its single in-edge
                  // (InfiniteRegularGraph(k): N = 1, self edge; InfiniteBipartiteRegularGraph((kA,kB)): N = 2, one edge pair)
  std::vector<int> inf_deg;
  bool periodic = false;  // periodic-in-time MPBP (periodic_mpbp, src/mpbp.jl:399-409): ring messages, csrc/periodic.cuh
  std::vector<int64_t> lz_off;  // offset of node i's log z_{i->j} entries in d_logzij
  int node_deg(int64_t i) const { return inf_k > 0 ? inf_deg[i] : (int)(colptr[i + 1] - colptr[i]); }
  int64_t out_edge(int64_t i, int k) const { return inf_k > 0 ? colptr[i] : colptr[i] + k; }
  std::vector<int> q;
  std::vector<int64_t> colptr, dst, rev, src;
  std::vector<int64_t> phi_off, psi_off, marg_off;
  double *d_phi = nullptr, *d_psi = nullptr;
  MsgStore msg[2];
  int cur = 0;
  int64_t slot = 0;  // doubles per message slot
  int sstride = 0;   // doubles per message site
  int* d_qprod = nullptr;
  double *d_marg = nullptr, *d_logzi = nullptr, *d_logzij = nullptr, *d_f = nullptr, *d_means = nullptr;
  int64_t* d_marg_off = nullptr;
  int* d_q = nullptr;
  double* d_delta = nullptr;
  int* d_err = nullptr;
  double* d_flops = nullptr;
  int64_t* d_edge_idx = nullptr;  // device copy of the edge list of the last pack/unpack call
  size_t edge_idx_cap = 0;
  std::vector<NodeClass> classes;
  std::vector<int> class_of_node;
  Arena arena;
  cudaStream_t st = nullptr;
  bool own_stream = true;
  static constexpr int NAUX = 7;
  cudaStream_t aux[NAUX] = {};  // extra streams: op groups of one level run concurrently
  cudaEvent_t ev_fork = nullptr, ev_join[NAUX] = {};
  double nstreams = 4;
  double bulk_split = 4;     // TSQR chunks allowed per tall matrix inside a FULL launch (1 = off)
  double bulk_split_min = 5; // ... for matrices of at least this many times n rows (X = nstates*q >= 6 at full bonds)
  double group_mode = 0;     // how the cost-sorted ops of a round are dealt into stream groups: 0 round-robin (every group the same
                             // mix), 1 contiguous chunks of equal work (homogeneous launches: no intra-launch tail)
  int twovar = 0;            // > 0: maxdist of the two-time marginals computed with every belief (option "twovar")
  double* d_tv = nullptr;    // [N][L][L][qmax*qmax]
  double hub_lane = 0;       // 1: high-degree nodes of a chunk run on their own (high-priority) stream (2: also for tiny chunks, tests).
                             // OFF by default: measured slower (N=256 bench 22.8 s vs 19.0 s per step) -- a hub kernel can only start
                             // when a bulk CTA retires, and the bulk QR CTAs run for 6-30 ms each, so the ~4500 sequential hub
                             // launches of a step queue behind them; kept as an option for graphs whose bulk CTAs are short.
  double hub_frac = 0.3;     // share of the chunk's cost the hub lane may take
  double lanes = 0;          // >= 2: lane mode, the nodes of a chunk are dealt (cost-balanced) into that many lanes, each running
                             // the whole cavity DAG of its nodes on its own stream with no barrier across lanes (1xx: also for
                             // tiny chunks, tests).  OFF by default: measured slower (53.7 / 52.5 edge-updates/s with 4 / 8
                             // lanes vs 55.1, N=256 bench) -- the smaller per-stream launches lose more in their tails
  cudaStream_t hub_st = nullptr;
  cudaEvent_t ev_hub_fork = nullptr, ev_hub_join = nullptr;
  double svd_mode = 2;       // truncating SVD of large matrices: 2 (default) un-squared block iteration, blocks orthonormalised by
                             // Cholesky-QR2 with the Householder route as fallback; 1 same with Householder only; 0 round-1
                             // squared iteration with Jacobi orthonormalisation when the block fits shared memory
  double tri_merge = 1;      // 1: the TSQR merge skips the zero panels of the stacked triangular chunk factors
  double kron_mma = 1;       // 1: DMMA Kronecker-carry kernel (k_kron_carry_mma), 0: scalar FP64 kernel (k_kron_carry)
  double outlier_split = 1.7;  // > 0: ops costing more than this multiple of the mean of their launch group run in a
                             // group of their own, TSQR-split, next to the other groups (0 = off).  1.7: N=256 bench 26.5 -> 23.4 s per step
  double level_balance = 1;  // stagger the cavity levels of the nodes of a chunk so that every round carries similar work
  double damp = 0.0;  // set by mpbp_iterate for the duration of the call
  // options
  double arena_gb = 0;       // 0 = auto
  double qr_fill = 148;      // CTAs that fill the GPU for the QR kernel (H = 64: one per SM); fewer ops per launch -> TSQR split
  double max_group_ops = 1e9;
  int profile = 0;
  // counters
  double n_launch = 0, qr_ms = 0, n_ops = 0, n_edge_updates = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  std::vector<int> ev_tag;
  size_t ev_used = 0;
  double fam_ms[16] = {0};
  int max_smem = 0;
};

namespace {

template <class T>
int upload(T** dptr, const T* host, size_t n) {
  if (n == 0) n = 1;
  CUDA_OK(cudaMalloc((void**)dptr, n * sizeof(T)));
  if (host) CUDA_OK(cudaMemcpy(*dptr, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int alloc_msg_store(mpbp_state* h, MsgStore& m) {
  const int64_t ne = std::max<int64_t>(h->E2, 1);
  CUDA_OK(cudaMalloc((void**)&m.data, sizeof(double) * h->slot * ne));
  CUDA_OK(cudaMalloc((void**)&m.bonds, sizeof(int) * (h->L + 1) * ne));
  CUDA_OK(cudaMalloc((void**)&m.ls, sizeof(double) * ne));
  return 0;
}

TTRef msg_ref(const mpbp_state* h, const MsgStore& m, int64_t e, int P) {
  TTRef r;
  r.data = m.data + e * h->slot;
  r.bonds = m.bonds + e * (h->L + 1);
  r.ls = m.ls + e;
  r.stride = h->sstride;
  r.P = P;
  return r;
}

int flat_messages(mpbp_state* h, MsgStore& m) {
  if (h->E2 == 0) return 0;  // a graph without edges is legal (the reference accepts it): nothing to initialise
  k_flat_messages<<<(unsigned)h->E2, 128, 0, h->st>>>(m.data, m.bonds, m.ls, h->d_qprod, h->slot, h->L, h->E2, h->sstride);
  h->n_launch++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int common_init(mpbp_state* h) {
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamCreate(&h->st));
  for (int k = 0; k < mpbp_state::NAUX; ++k) {
    CUDA_OK(cudaStreamCreateWithFlags(&h->aux[k], cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&h->ev_join[k], cudaEventDisableTiming));
  }
  CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  {
    // highest priority: the hub lane is the critical path of a step; its (few, small) launches must get the SMs that the
    // bulk kernels free, ahead of the bulk's own queued CTAs
    int lo = 0, hi = 0;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_OK(cudaStreamCreateWithPriority(&h->hub_st, cudaStreamNonBlocking, hi));
  }
  CUDA_OK(cudaEventCreateWithFlags(&h->ev_hub_fork, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->ev_hub_join, cudaEventDisableTiming));
  CUDA_OK(cudaDeviceGetAttribute(&h->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  h->max_smem -= 2048;  // head-room for the kernels' static shared memory
  {
    const int ms = h->max_smem;
    CUDA_OK(cudaFuncSetAttribute(k_kron_carry<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_kron_carry<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_kron_carry_mma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_kron_carry_mma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_kron_carry_mma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_qr_ft<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_qr_ft_merge<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_qr_small<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_qr_ft<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_qr_ft<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_qr_ft_merge<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_qr_ft_merge<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_kron_proj, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_qr_small<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_qr_small<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_jacobi_project, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_damp, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_belief, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_pair_belief, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
    CUDA_OK(cudaFuncSetAttribute(k_alt_marginal, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
  }
  const int L = h->L;
  h->qmax = *std::max_element(h->q.begin(), h->q.end());
  h->sstride = h->dmax * h->dmax * h->qmax * h->qmax;
  h->slot = (int64_t)L * h->sstride;
  h->phi_off.resize(h->N + 1);
  h->marg_off.resize(h->N + 1);
  h->phi_off[0] = 0;
  for (int64_t i = 0; i < h->N; ++i) h->phi_off[i + 1] = h->phi_off[i] + (int64_t)L * h->q[i];
  h->marg_off = h->phi_off;
  h->psi_off.resize(h->E2 + 1);
  h->psi_off[0] = 0;
  std::vector<int> qprod(h->E2);
  for (int64_t e = 0; e < h->E2; ++e) {
    qprod[e] = h->q[h->src[e]] * h->q[h->dst[e]];
    h->psi_off[e + 1] = h->psi_off[e] + (int64_t)L * qprod[e];
  }
  std::vector<double> ones(std::max(h->phi_off[h->N], h->psi_off[h->E2]), 1.0);
  if (upload(&h->d_phi, ones.data(), h->phi_off[h->N])) return 1;
  if (upload(&h->d_psi, ones.data(), h->psi_off[h->E2])) return 1;
  if (upload(&h->d_qprod, qprod.data(), h->E2)) return 1;
  if (upload(&h->d_q, h->q.data(), h->N)) return 1;
  if (upload(&h->d_marg_off, h->marg_off.data(), h->N + 1)) return 1;
  // beliefs start uniform (flat_mpem1), means accordingly
  std::vector<double> marg(h->marg_off[h->N]);
  std::vector<double> means(h->N * L);
  for (int64_t i = 0; i < h->N; ++i) {
    for (int64_t k = h->marg_off[i]; k < h->marg_off[i + 1]; ++k) marg[k] = 1.0 / h->q[i];
    for (int t = 0; t < L; ++t) means[i * L + t] = 0.5 * (h->q[i] + 1);
  }
  if (upload(&h->d_marg, marg.data(), marg.size())) return 1;
  if (upload(&h->d_means, means.data(), means.size())) return 1;
  h->lz_off.assign(h->N + 1, 0);
  for (int64_t i = 0; i < h->N; ++i) h->lz_off[i + 1] = h->lz_off[i] + h->node_deg(i);
  const int64_t nzij = h->lz_off[h->N];
  std::vector<double> zeros(std::max<int64_t>(std::max<int64_t>(h->N, nzij), 32), 0.0);
  if (upload(&h->d_logzi, zeros.data(), h->N)) return 1;
  if (upload(&h->d_logzij, zeros.data(), nzij)) return 1;
  if (upload(&h->d_f, zeros.data(), h->N)) return 1;
  if (upload(&h->d_delta, zeros.data(), 1)) return 1;
  if (upload(&h->d_flops, zeros.data(), 24)) return 1;
  int zero = 0;
  if (upload(&h->d_err, &zero, 1)) return 1;
  if (alloc_msg_store(h, h->msg[0])) return 1;
  if (flat_messages(h, h->msg[0])) return 1;
  CUDA_OK(cudaStreamSynchronize(h->st));
  return 0;
}

int ensure_arena(mpbp_state* h) {
  if (h->arena.base) return 0;
  size_t freeb = 0, total = 0;
  CUDA_OK(cudaMemGetInfo(&freeb, &total));
  size_t want = h->arena_gb > 0 ? (size_t)(h->arena_gb * 1e9) : (size_t)(freeb * 0.80);
  if (h->arena_gb <= 0) {
    // leave room for a second message buffer (Jacobi schedule)
    size_t msgb = sizeof(double) * h->slot * h->E2;
    if (want > msgb + (size_t(1) << 30)) want -= msgb;
  }
  CUDA_OK(cudaMalloc((void**)&h->arena.base, want));
  h->arena.cap = want;
  return 0;
}

struct Plan {
  std::vector<BtJob> bt;
  std::vector<GenJob> gen;
  std::vector<InitJob> init;
  std::vector<std::vector<OpDesc>> levels;  // ops by level (scratch pointers filled per group)
  std::vector<std::vector<int>> capA, capB;  // bond capacities of the operands per op (1 or dmax)
  // hub lane: the ops of the few high-degree nodes of the chunk, whose level chains are the critical path; they run on
  // their own stream, concurrently with the rounds of the bulk
  struct Lane {
    std::vector<std::vector<OpDesc>> levels;
    std::vector<std::vector<int>> capA, capB;
  };
  std::vector<Lane> lanes;  // index = lane id - 1 (lane id 0 = the bulk rounds above)
  std::vector<FinJob> fin;
  std::vector<FinJob> gfin;    // generic path: the dummy-neighbour message of every generic node (compressed before it is
  std::vector<MargJob> gmarg;  // marginalised into the belief, src/mpbp.jl:145-154), and the marginals read from it
  std::vector<DampJob> damp;  // one per FinJob when damp > 0 (same order)
  std::vector<BelJob> bel;
  std::vector<FJob> fj;
};

TTRef arena_tt(mpbp_state* h, int cap, int P, bool& ok) {
  TTRef r;
  r.stride = cap * cap * P;
  r.P = P;
  r.data = (double*)h->arena.take(sizeof(double) * (size_t)h->L * r.stride);
  r.bonds = (int*)h->arena.take(sizeof(int) * (h->L + 1));
  r.ls = (double*)h->arena.take(sizeof(double));
  ok = ok && r.data && r.bonds && r.ls;
  return r;
}

// bytes of persistent (per-node) arena storage needed by one node update
size_t node_bytes(const mpbp_state* h, int64_t i) {
  const NodeClass& c = h->classes[h->class_of_node[i]];
  const int z = c.z, q = c.q, d = h->dmax, L = h->L;
  auto tt = [&](int cap, int ny) { return (size_t)L * cap * cap * ny * q * 8 + 4 * (L + 1) + 8 + 3 * 256; };
  size_t b = 0;
  if (h->twovar > 0) b += 8 * ((size_t)L * d * d * q * q + (size_t)L * d * q) + 2 * 256;
  if (c.generic) {
    for (int j = 0; j <= z; ++j) b += tt(d, c.gen_ny[j]);
    const int qjm = h->qmax;
    const size_t fin = (size_t)(L + 1) * (d * q) * (d * q) + (size_t)d * d * q * q * q * qjm + (size_t)d * d * q * q * qjm +
                       (size_t)d * d * q * qjm + 2 * (size_t)d * d * q + (L + 1) + (size_t)(d * q * qjm) * (d * q * qjm);
    b += (size_t)z * (fin * 8 + 8 * 256) + ((size_t)L * d * q + (size_t)d * d * q * q) * 8 + 2 * 256;
    b += fin * 8 + h->slot * 8 + 4 * (L + 1) + 8 + (size_t)(L + 1) * d * 8 + 16 * 256;  // dummy-neighbour message + marginals
    if (h->inf_k > 0) b += (size_t)z * (h->slot * 8 + 4 * (L + 1) + 8 + 3 * 256);
    if (h->damp > 0.0) {
      const size_t b2 = 2 * (size_t)d, Pm = (size_t)h->qmax * h->qmax;
      b += (size_t)z * (h->slot * 8 + 8 * ((size_t)(L + 1) * b2 * b2 + b2 * Pm * b2 + 2 * d * b2 * Pm + (d * Pm) * (d * Pm) + 2 * d * b2) + 4 * (L + 1) * 2 + 16 * 256);
    }
    return b;
  }
  if (z > 0) b += (size_t)z * tt(d, c.ny[1]);
  b += tt(1, c.ny[0]);
  for (int k = 1; k < z; ++k) b += tt(d, c.ny[k + 1]) + tt(d, c.ny[z - k]) + tt(d, c.ny[z - 1]);
  b += tt(d, c.ny[z]);
  if (z == 0) return b + ((size_t)L * d * q + (size_t)d * d * q * q) * 8 + 4 * 256;
  // finalize + belief scratch
  const int qjm = h->qmax;
  size_t fin = (size_t)(L + 1) * (d * q) * (d * q) + (size_t)d * d * q * q * q * qjm + (size_t)d * d * q * q * qjm +
               (size_t)d * d * q * qjm + 2 * (size_t)d * d * q + (L + 1) + (size_t)(d * q * qjm) * (d * q * qjm);
  b += (size_t)z * (fin * 8 + 8 * 256);
  b += ((size_t)L * d * q + (size_t)d * d * q * q) * 8 + 2 * 256;
  if (h->inf_k > 0) b += (size_t)z * (h->slot * 8 + 4 * (L + 1) + 8 + 3 * 256);
  if (h->damp > 0.0) {
    const size_t b2 = 2 * (size_t)d, Pm = (size_t)h->qmax * h->qmax;
    b += (size_t)z * (h->slot * 8 + 8 * ((size_t)(L + 1) * b2 * b2 + b2 * Pm * b2 + 2 * d * b2 * Pm + (d * Pm) * (d * Pm) + 2 * d * b2) + 4 * (L + 1) * 2 + 16 * 256);
  }
  return b;
}

// doubles of the tall sweep-1 matrix M of an op; sweep 2 reuses it as the global scratch of the truncating SVD, so it is at
// least svd_scratch_doubles(p <= dX, n <= D)
size_t op_M_doubles(const mpbp_state* h, size_t D, int X) {
  const size_t d = h->dmax;
  return std::max(D * X * D, svd_scratch_doubles((int)(d * X), (int)D));
}
size_t op_scratch_bytes(const mpbp_state* h, int capA, int capB, int X) {
  const size_t D = (size_t)capA * capB, d = h->dmax, L = h->L;
  size_t dbl = L * D * D + op_M_doubles(h, D, X) + QR_NSPLIT_MAX * D * D + d * D * X + D * d * X + (d * X) * (d * X) + 2 * d * D;
  return dbl * 8 + 4 * (L + 1) + 10 * 256;
}

// damp > 0: the finalised message goes to a scratch slot, then k_damp combines it with the old message IN PLACE on the
// write-buffer slot (which holds the old message: same buffer for the sequential schedule, a copy for the Jacobi one)
void add_damp_job(mpbp_state* h, Plan& P, FinJob& fj, const TTRef& dest, bool& ok) {
  const int L = h->L, d = h->dmax;
  const int P_ = fj.q * fj.qj;
  TTRef scratch;
  scratch.stride = h->sstride;
  scratch.P = P_;
  scratch.data = (double*)h->arena.take(sizeof(double) * h->slot);
  scratch.bonds = (int*)h->arena.take(sizeof(int) * (L + 1));
  scratch.ls = (double*)h->arena.take(sizeof(double));
  ok = ok && scratch.data && scratch.bonds && scratch.ls;
  fj.out = scratch;
  DampJob dj;
  memset(&dj, 0, sizeof dj);
  dj.a = scratch;
  dj.b = dest;
  dj.out = dest;
  dj.q = fj.q;
  dj.qj = fj.qj;
  dj.coef = h->damp / (1.0 - h->damp);
  const size_t b2 = 2 * (size_t)d;
  dj.lstride = (int)(b2 * b2);
  dj.Lbuf = (double*)h->arena.take(8 * (size_t)(L + 1) * dj.lstride);
  dj.r = (int*)h->arena.take(4 * (L + 1));
  dj.S = (double*)h->arena.take(8 * b2 * P_ * b2);
  dj.G = (double*)h->arena.take(8 * (size_t)d * b2 * P_);
  dj.M2 = (double*)h->arena.take(8 * (size_t)d * P_ * b2);
  dj.R2 = (double*)h->arena.take(8 * (size_t)(d * P_) * (d * P_));
  dj.Pc[0] = (double*)h->arena.take(8 * (size_t)d * b2);
  dj.Pc[1] = (double*)h->arena.take(8 * (size_t)d * b2);
  ok = ok && dj.Lbuf && dj.r && dj.S && dj.G && dj.M2 && dj.R2 && dj.Pc[0] && dj.Pc[1];
  P.damp.push_back(dj);
}

// Work (~ D^2 X, the QR and SVD cost of an op up to a constant) that a degree-z node of class c puts on each
// level of its cavity DAG; mirrors the level assignment of build_plan below.
std::vector<double> node_level_cost(const NodeClass& c, int d) {
  const int z = c.z, q = c.q;
  std::vector<double> w(z + 1, 0.0);
  auto cost = [&](int ca, int cb, int ny) { return (double)ca * cb * ca * cb * ny * q; };
  if (z == 1) w[1] += cost(d, 1, c.ny[1]);
  if (z < 2) return w;
  for (int k = 1; k < z; ++k) w[k] += cost(d, d, c.ny[k + 1]);
  w[z] += cost(d, 1, c.ny[z]);
  for (int k = z - 1; k >= 1; --k) w[z - k] += cost(d, k == z - 1 ? 1 : d, c.ny[z - k]);
  for (int k = 1; k < z; ++k) w[std::max(k - 1, z - k - 1) + 1] += cost(d, k == z - 1 ? 1 : d, c.ny[z - 1]);
  return w;
}

// Level offsets: the cavity DAG of a degree-z node has z levels, and the nodes of a chunk are independent, so node i
// may run its level l in round l + off[i] for any 0 <= off[i] <= zmax - z_i.  Greedy, deepest nodes first: each node
// takes the offset where its work overlaps least with the work already placed, so that the deep levels of the few
// high-degree nodes share their rounds with the bulk of the low-degree ones instead of running alone on an
// under-filled GPU.  The per-node order of operations (hence the result) is unchanged.
std::vector<int> plan_level_offsets(const mpbp_state* h, const std::vector<int64_t>& nodes) {
  std::vector<int> off(nodes.size(), 0);
  if (h->level_balance <= 0 || nodes.size() < 2) return off;
  int zmax = 0;
  std::vector<size_t> order;
  for (size_t k = 0; k < nodes.size(); ++k) {
    const int ci = h->class_of_node[nodes[k]];
    if (ci < 0 || ci >= (int)h->classes.size() || h->classes[ci].generic) continue;
    zmax = std::max(zmax, h->classes[ci].z);
    order.push_back(k);
  }
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
    return h->classes[h->class_of_node[nodes[a]]].z > h->classes[h->class_of_node[nodes[b]]].z;
  });
  std::vector<double> load(zmax + 1, 0.0);
  std::map<int, std::vector<double>> wcache;
  for (size_t k : order) {
    const int ci = h->class_of_node[nodes[k]];
    const NodeClass& c = h->classes[ci];
    auto it = wcache.find(ci);
    if (it == wcache.end()) it = wcache.emplace(ci, node_level_cost(c, h->dmax)).first;
    const std::vector<double>& w = it->second;
    int best = 0;
    double bestv = -1.0;
    for (int o = 0; o + c.z <= zmax; ++o) {
      double v = 0.0;
      for (int l = 1; l <= c.z; ++l) v += load[l + o] * w[l];
      if (bestv < 0.0 || v < bestv) { bestv = v; best = o; }
    }
    off[k] = best;
    for (int l = 1; l <= c.z; ++l) load[l + best] += w[l];
  }
  return off;
}

int build_plan(mpbp_state* h, const std::vector<int64_t>& nodes, const std::vector<char>& hub, int rb, int wb, Plan& P) {
  const int L = h->L, d = h->dmax;
  bool ok = true;
  // level offsets are planned per lane (each lane balances its own rounds)
  std::vector<int> lev_off(nodes.size(), 0);
  {
    int nl = 0;
    for (char c : hub) nl = std::max(nl, (int)c);
    P.lanes.resize(nl);
    for (int lane = 0; lane <= nl; ++lane) {
      std::vector<int64_t> nn;
      std::vector<size_t> ii;
      for (size_t k = 0; k < nodes.size(); ++k)
        if (hub[k] == lane) { nn.push_back(nodes[k]); ii.push_back(k); }
      const std::vector<int> o = plan_level_offsets(h, nn);
      for (size_t k = 0; k < ii.size(); ++k) lev_off[ii[k]] = o[k];
    }
  }
  for (size_t inode = 0; inode < nodes.size(); ++inode) {
    const int64_t i = nodes[inode];
    const int loff = lev_off[inode];
    const int lane_id = hub[inode];
    const int ci = h->class_of_node[i];
    if (ci < 0 || ci >= (int)h->classes.size()) return fail("node %lld has no factor class", (long long)i);
    const NodeClass& c = h->classes[ci];
    const int z = c.z, q = c.q;
    if (c.generic) {
      // exhaustive-trace path (src/mpbp.jl:117-154, src/bp_core.jl:18-93): Kronecker of the other neighbours' messages,
      // then the same MPEM3->MPEM2 / compress / normalize kernel with the dense factor table as W
      const int degg = h->node_deg(i);
      if (degg != z) return fail("node %lld has degree %d but its class has z=%d", (long long)i, degg, z);
      if (q != h->q[i]) return fail("node %lld: class q mismatch", (long long)i);
      auto make_gen = [&](int skip, int ny) -> TTRef {
        GenJob gj;
        memset(&gj, 0, sizeof gj);
        gj.nk = z;
        gj.skip = skip;
        gj.q = q;
        for (int k = 0; k < z; ++k) {
          const int64_t eout = h->out_edge(i, k);
          const int64_t ein = h->rev[eout];
          gj.qk[k] = c.qn[k];
          gj.msg[k] = msg_ref(h, h->msg[rb], ein, c.qn[k] * q);
          gj.psi[k] = h->d_psi + h->psi_off[eout];
        }
        gj.out = arena_tt(h, d, ny * q, ok);
        P.gen.push_back(gj);
        return gj.out;
      };
      for (int j = 0; j < z; ++j) {
        const int64_t eout = h->out_edge(i, j);
        const int qj = c.qn[j];
        FinJob fj;
        memset(&fj, 0, sizeof fj);
        fj.c = make_gen(j, c.gen_ny[j]);
        fj.q = q;
        fj.qj = qj;
        // (damp is ignored on the generic path, exactly like src/mpbp.jl:117-138)
        if (h->inf_k > 0 && j < z - 1) {
          TTRef scratch;
          scratch.stride = h->sstride;
          scratch.P = q * qj;
          scratch.data = (double*)h->arena.take(sizeof(double) * h->slot);
          scratch.bonds = (int*)h->arena.take(sizeof(int) * (L + 1));
          scratch.ls = (double*)h->arena.take(sizeof(double));
          ok = ok && scratch.data && scratch.bonds && scratch.ls;
          fj.out = scratch;
        } else {
          fj.out = msg_ref(h, h->msg[wb], eout, q * qj);
        }
        fj.nyc = c.gen_ny[j];
        fj.q = q;
        fj.qj = qj;
        fj.W = c.d_w + c.w_off[j];
        fj.w_tstride = (int)c.w_ts;
        fj.phi = h->d_phi + h->phi_off[i];
        fj.logz_out = h->d_logzij + h->lz_off[i] + j;
        const size_t dq = (size_t)d * q;
        fj.rstride = (int)(dq * dq);
        fj.Rbuf = (double*)h->arena.take(8 * (size_t)(L + 1) * fj.rstride);
        fj.kdim = (int*)h->arena.take(4 * (L + 1));
        fj.Bt = (double*)h->arena.take(8 * (size_t)d * d * q * qj * q);
        fj.S = (double*)h->arena.take(8 * (size_t)d * d * q * q * q * qj);
        fj.H = (double*)h->arena.take(8 * (size_t)d * d * q * qj);
        fj.R2 = (double*)h->arena.take(8 * (size_t)(d * q * qj) * (d * q * qj));
        fj.Pr[0] = (double*)h->arena.take(8 * (size_t)d * d * q);
        fj.Pr[1] = (double*)h->arena.take(8 * (size_t)d * d * q);
        ok = ok && fj.Rbuf && fj.kdim && fj.Bt && fj.S && fj.H && fj.R2 && fj.Pr[0] && fj.Pr[1];
        P.fin.push_back(fj);
      }
      BelJob bj;
      memset(&bj, 0, sizeof bj);
      bj.full = make_gen(-1, c.gen_ny[z]);
      bj.ny = c.gen_ny[z];
      bj.q = q;
      bj.Wd = c.d_wd;
      bj.w_tstride = (int)c.wd_ts;
      bj.phi = h->d_phi + h->phi_off[i];
      bj.marg = h->d_marg + h->marg_off[i];
      bj.logz = h->d_logzi + i;
      bj.bw = (double*)h->arena.take(8 * (size_t)L * d * q);
      bj.Bt = (double*)h->arena.take(8 * (size_t)d * d * q * q);
      ok = ok && bj.bw && bj.Bt;
      if (h->twovar > 0) {
        bj.tv = h->d_tv + (size_t)i * L * L * h->qmax * h->qmax;
        bj.Btall = (double*)h->arena.take(8 * (size_t)(L - 1 > 0 ? L - 1 : 1) * d * d * q * q);
        bj.fwall = (double*)h->arena.take(8 * (size_t)L * d * q);
        ok = ok && bj.Btall && bj.fwall;
      }
      P.bel.push_back(bj);
      {
        // bp.b[i] = marginalize(compress!(mpem2(f_bp_dummy_neighbor))) : same kernel as an outgoing message with a
        // one-state dummy neighbour (q_j = 1) and the full table, written to a scratch slot; log z_i = its normalisation
        FinJob fj;
        memset(&fj, 0, sizeof fj);
        fj.c = bj.full;
        fj.nyc = c.gen_ny[z];
        fj.q = q;
        fj.qj = 1;
        fj.W = c.d_wd;
        fj.w_tstride = (int)c.wd_ts;
        fj.phi = h->d_phi + h->phi_off[i];
        fj.logz_out = h->d_logzi + i;
        TTRef slot;
        slot.stride = h->sstride;
        slot.P = q;
        slot.data = (double*)h->arena.take(sizeof(double) * h->slot);
        slot.bonds = (int*)h->arena.take(sizeof(int) * (L + 1));
        slot.ls = (double*)h->arena.take(sizeof(double));
        fj.out = slot;
        const size_t dq = (size_t)d * q;
        fj.rstride = (int)(dq * dq);
        fj.Rbuf = (double*)h->arena.take(8 * (size_t)(L + 1) * fj.rstride);
        fj.kdim = (int*)h->arena.take(4 * (L + 1));
        fj.Bt = (double*)h->arena.take(8 * (size_t)d * d * q * q);
        fj.S = (double*)h->arena.take(8 * (size_t)d * d * q * q * q);
        fj.H = (double*)h->arena.take(8 * (size_t)d * d * q);
        fj.R2 = (double*)h->arena.take(8 * (size_t)(d * q) * (d * q));
        fj.Pr[0] = (double*)h->arena.take(8 * (size_t)d * d * q);
        fj.Pr[1] = (double*)h->arena.take(8 * (size_t)d * d * q);
        ok = ok && slot.data && slot.bonds && slot.ls && fj.Rbuf && fj.kdim && fj.Bt && fj.S && fj.H && fj.R2 && fj.Pr[0] && fj.Pr[1];
        P.gfin.push_back(fj);
        MargJob mj;
        mj.tt = slot;
        mj.q = q;
        mj.qj = 1;
        mj.marg = h->d_marg + h->marg_off[i];
        mj.rv = (double*)h->arena.take(8 * (size_t)(L + 1) * d);
        ok = ok && mj.rv;
        P.gmarg.push_back(mj);
      }
      FJob f;
      f.logzi = h->d_logzi + i;
      f.logzij = h->d_logzij + h->lz_off[i];
      f.z = z;
      f.f = h->d_f + i;
      P.fj.push_back(f);
      continue;
    }
    const int deg = h->node_deg(i);
    if (deg != z) return fail("node %lld has degree %d but its class has z=%d", (long long)i, deg, z);
    if (q != h->q[i]) return fail("node %lld: class q mismatch", (long long)i);
    // B~_k
    std::vector<TTRef> src(z);
    for (int k = 0; k < z; ++k) {
      const int64_t eout = h->out_edge(i, k);
      const int64_t ein = h->rev[eout];
      const int qk = h->q[h->dst[eout]];
      if (qk != c.qn[k]) return fail("node %lld neighbour %d: class qn mismatch", (long long)i, k);
      BtJob jb;
      jb.msg = msg_ref(h, h->msg[rb], ein, qk * q);
      jb.out = arena_tt(h, d, c.ny[1] * q, ok);
      jb.psi = h->d_psi + h->psi_off[eout];
      jb.pxy = c.d_pxy + c.pxy_off[k];
      jb.pxy_tstride = (int)c.pxy_ts;
      jb.qk = qk;
      jb.qi = q;
      jb.ny1 = c.ny[1];
      src[k] = jb.out;
      P.bt.push_back(jb);
    }
    // init
    InitJob ij;
    ij.out = arena_tt(h, 1, c.ny[0] * q, ok);
    ij.minit = c.d_minit;
    ij.tstride = (int)c.minit_ts;
    ij.n = c.ny[0] * q;
    P.init.push_back(ij);
    const TTRef init = ij.out;
    auto add_op = [&](int level, const TTRef& a, int da, const TTRef& b, int db, int ca, int cb, TTRef& out) -> int {
      level += loff;
      auto it = c.pyy.find({da, db});
      if (it == c.pyy.end()) return fail("class %d lacks the prob_yy table for (d1,d2)=(%d,%d)", ci, da, db);
      OpDesc op;
      memset(&op, 0, sizeof op);
      op.a = a;
      op.b = b;
      op.ny1 = c.ny[da];
      op.ny2 = c.ny[db];
      op.nyo = c.ny[da + db];
      op.q = q;
      out = arena_tt(h, d, op.nyo * q, ok);
      op.o = out;
      op.pyy = c.d_pyy + it->second.first;
      op.pyy_tstride = (int)it->second.second;
      auto& LV = lane_id ? P.lanes[lane_id - 1].levels : P.levels;
      auto& CA = lane_id ? P.lanes[lane_id - 1].capA : P.capA;
      auto& CB = lane_id ? P.lanes[lane_id - 1].capB : P.capB;
      if ((int)LV.size() <= level) {
        LV.resize(level + 1);
        CA.resize(level + 1);
        CB.resize(level + 1);
      }
      LV[level].push_back(op);
      CA[level].push_back(ca);
      CB[level].push_back(cb);
      return 0;
    };
    std::vector<TTRef> dest(z);
    TTRef full;
    if (z == 0) {
      full = init;  // cavity of an empty neighbourhood: (∅, init)
    } else if (z == 1) {
      dest[0] = init;
      if (add_op(1, src[0], 1, init, 0, d, 1, full)) return 1;
    } else {
      std::vector<TTRef> p(z), s(z + 1);
      p[0] = src[0];
      for (int k = 1; k < z; ++k)
        if (add_op(k, p[k - 1], k, src[k], 1, d, d, p[k])) return 1;
      if (add_op(z, p[z - 1], z, init, 0, d, 1, full)) return 1;
      s[z] = init;
      for (int k = z - 1; k >= 1; --k)
        if (add_op(z - k, src[k], 1, s[k + 1], z - 1 - k, d, k == z - 1 ? 1 : d, s[k])) return 1;
      for (int k = 1; k < z; ++k)
        if (add_op(std::max(k - 1, z - k - 1) + 1, p[k - 1], k, s[k + 1], z - 1 - k, d, k == z - 1 ? 1 : d, dest[k]))
          return 1;
      dest[0] = s[1];
    }
    // outgoing messages
    for (int j = 0; j < z; ++j) {
      const int64_t eout = h->out_edge(i, j);
      const int qj = c.qn[j];
      FinJob fj;
      memset(&fj, 0, sizeof fj);
      fj.c = dest[j];
      fj.q = q;
      fj.qj = qj;
      if (h->damp > 0.0) {
        add_damp_job(h, P, fj, msg_ref(h, h->msg[wb], eout, q * qj), ok);
      } else if (h->inf_k > 0 && j < z - 1) {
        // only the last recomputation stays in bp.mu[1] (src/infinite_graph.jl + recursive_bp_factor.jl:154-158)
        TTRef scratch;
        scratch.stride = h->sstride;
        scratch.P = q * qj;
        scratch.data = (double*)h->arena.take(sizeof(double) * h->slot);
        scratch.bonds = (int*)h->arena.take(sizeof(int) * (L + 1));
        scratch.ls = (double*)h->arena.take(sizeof(double));
        ok = ok && scratch.data && scratch.bonds && scratch.ls;
        fj.out = scratch;
      } else {
        fj.out = msg_ref(h, h->msg[wb], eout, q * qj);
      }
      fj.nyc = c.ny[z - 1];
      fj.q = q;
      fj.qj = qj;
      fj.W = c.d_w + c.w_off[j];
      fj.w_tstride = (int)c.w_ts;
      fj.phi = h->d_phi + h->phi_off[i];
      fj.logz_out = h->d_logzij + h->lz_off[i] + j;
      const size_t dq = (size_t)d * q;
      fj.rstride = (int)(dq * dq);
      fj.Rbuf = (double*)h->arena.take(8 * (size_t)(L + 1) * fj.rstride);
      fj.kdim = (int*)h->arena.take(4 * (L + 1));
      fj.Bt = (double*)h->arena.take(8 * (size_t)d * d * q * qj * q);
      fj.S = (double*)h->arena.take(8 * (size_t)d * d * q * q * q * qj);
      fj.H = (double*)h->arena.take(8 * (size_t)d * d * q * qj);
      fj.R2 = (double*)h->arena.take(8 * (size_t)(d * q * qj) * (d * q * qj));
      fj.Pr[0] = (double*)h->arena.take(8 * (size_t)d * d * q);
      fj.Pr[1] = (double*)h->arena.take(8 * (size_t)d * d * q);
      ok = ok && fj.Rbuf && fj.kdim && fj.Bt && fj.S && fj.H && fj.R2 && fj.Pr[0] && fj.Pr[1];
      P.fin.push_back(fj);
    }
    BelJob bj;
    memset(&bj, 0, sizeof bj);
    bj.full = full;
    bj.ny = c.ny[z];
    bj.q = q;
    bj.Wd = c.d_wd;
    bj.w_tstride = (int)c.wd_ts;
    bj.phi = h->d_phi + h->phi_off[i];
    bj.marg = h->d_marg + h->marg_off[i];
    bj.logz = h->d_logzi + i;
    bj.bw = (double*)h->arena.take(8 * (size_t)L * d * q);
    bj.Bt = (double*)h->arena.take(8 * (size_t)d * d * q * q);
    ok = ok && bj.bw && bj.Bt;
    if (h->twovar > 0) {
      bj.tv = h->d_tv + (size_t)i * L * L * h->qmax * h->qmax;
      bj.Btall = (double*)h->arena.take(8 * (size_t)(L - 1 > 0 ? L - 1 : 1) * d * d * q * q);
      bj.fwall = (double*)h->arena.take(8 * (size_t)L * d * q);
      ok = ok && bj.Btall && bj.fwall;
    }
    P.bel.push_back(bj);
    FJob f;
    f.logzi = h->d_logzi + i;
    f.logzij = h->d_logzij + h->lz_off[i];
    f.z = z;
    f.f = h->d_f + i;
    P.fj.push_back(f);
  }
  if (!ok) return fail("arena exhausted while planning %zu nodes (internal sizing error)", nodes.size());
  return 0;
}

template <class J>
int upload_jobs(mpbp_state* h, const std::vector<J>& v, J** d) {
  *d = (J*)h->arena.take(sizeof(J) * std::max<size_t>(v.size(), 1));
  if (!*d) return fail("arena exhausted (job descriptors)");
  if (!v.empty()) CUDA_OK(cudaMemcpyAsync(*d, v.data(), sizeof(J) * v.size(), cudaMemcpyHostToDevice, h->st));
  return 0;
}

enum { F_QR = 0, F_KC = 1, F_KP = 2, F_GEMM = 3, F_QRS = 4, F_JAC = 5, F_FIN = 6, F_BEL = 7, F_BT = 8, F_NFAM = 9 };
void ev_begin(mpbp_state* h, int tag, cudaStream_t st) {
  if (!h->profile) return;
  if (h->ev_used == h->ev_pool.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    h->ev_pool.push_back({a, b});
    h->ev_tag.push_back(0);
  }
  h->ev_tag[h->ev_used] = tag;
  cudaEventRecord(h->ev_pool[h->ev_used].first, st);
}
void ev_flush(mpbp_state* h) {
  if (!h->profile || h->ev_used == 0) return;
  cudaStreamSynchronize(h->st);
  for (int k = 0; k < mpbp_state::NAUX; ++k) cudaStreamSynchronize(h->aux[k]);
  cudaStreamSynchronize(h->hub_st);
  for (size_t k = 0; k < h->ev_used; ++k) {
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev_pool[k].first, h->ev_pool[k].second);
    h->fam_ms[h->ev_tag[k]] += ms;
    if (h->ev_tag[k] == F_QR) h->qr_ms += ms;
  }
  h->ev_used = 0;
}
void ev_end(mpbp_state* h, cudaStream_t st) {
  if (!h->profile) return;
  cudaEventRecord(h->ev_pool[h->ev_used].second, st);
  h->ev_used++;
}

// one group of ops of one level (scratch already assigned) and the stream it runs on
struct GroupRun {
  const OpDesc* d_ops;
  int nops, maxD, maxX, maxNy, maxq, maxNyS;
  int kc_mma_rb;  // 16 / 8 / 4: right-bond columns per CTA of the DMMA Kronecker-carry kernel; 0 = scalar kernel
  cudaStream_t st;
  size_t kc_smem, ft_big, ft_small;
  int kc_rb;
  int bigH, smallH;  // row-block height of the flat-tree QR: 64 (one CTA/SM, D >= 200), 32 (two CTAs/SM) or 16
  int dXcap;
  size_t jac_doubles, jac_smem;  // shared memory of the truncating-SVD kernel: only what the group's matrices need
  int nsplit;  // TSQR chunks allowed per matrix in this group's QR launches
  int split_min;  // ... for matrices of at least split_min * n rows (0: every matrix, the rule of an under-filled launch)
};

// run the groups of one level concurrently: the per-site launch sequences of the groups are issued interleaved on
// their own streams, so the tail of one group's launch is filled by the other groups' kernels
int run_op_groups(mpbp_state* h, std::vector<GroupRun>& groups_in, const Trunc& tr) {
  const int L = h->L, d = h->dmax;
  std::vector<GroupRun> groups;  // (a group can come out empty of the contiguous dealing: nothing to launch for it)
  for (auto& g : groups_in)
    if (g.nops > 0) groups.push_back(g);
  if (groups.empty()) return 0;
  const size_t kp_smem = (size_t)d * d * d * 8;
  const size_t jac_fixed = 3 * SUB_BMAX;
  const size_t jac_doubles = (size_t)h->max_smem / 8 - jac_fixed;
  if (kp_smem > (size_t)h->max_smem) return fail("bond capacity %d exceeds the shared-memory tiling of k_kron_proj", d);
  for (auto& g : groups) {
    g.kc_smem = ((size_t)g.maxD + (size_t)d * d * g.maxNy) * 8;
    g.kc_rb = (4 * g.kc_smem <= (size_t)h->max_smem / 2) ? 4 : 1;  // keep >= 2 CTAs/SM
    g.kc_smem *= g.kc_rb;
    if (g.kc_smem > (size_t)h->max_smem)
      return fail("bond capacity %d / nstates %d exceed the shared-memory tiling of the contraction kernels", d, g.maxNy);
    g.kc_mma_rb = 0;
    if (h->kron_mma > 0) {
      // the widest tile that fits (the kernel keeps its A fragments and 4 n-tiles of accumulators in registers: one CTA per SM)
      for (int rb : {16, 8, 4})
        if (!g.kc_mma_rb && kc_mma_smem_doubles(rb, g.maxD, d, g.maxNyS, g.maxNy) * 8 <= (size_t)h->max_smem) g.kc_mma_rb = rb;
    }
    g.dXcap = d * g.maxX;
    {
      // small problems (every matrix of the group goes through the direct Jacobi path with a p x c block, c <= 64): ask for
      // the block only, so that several CTAs share an SM; otherwise the whole opt-in shared memory (subspace iteration)
      const size_t cmax = (size_t)std::min(g.dXcap, g.maxD);
      g.jac_doubles = jac_doubles;
      if (cmax <= (size_t)SUB_BMAX && (size_t)g.dXcap * cmax + 64 <= jac_doubles) g.jac_doubles = std::max<size_t>((size_t)g.dXcap * cmax + 64, 1024);
      g.jac_smem = (jac_fixed + g.jac_doubles) * 8;
    }
    auto pickH = [&](int n, size_t& bytes) -> int {
      const size_t s64 = ft_smem_doubles<64>(n) * 8, s32 = ft_smem_doubles<32>(n) * 8, s16 = ft_smem_doubles<16>(n) * 8;
      if (n >= 200 && s64 <= (size_t)h->max_smem) { bytes = s64; return 64; }
      if (s32 <= (size_t)h->max_smem) { bytes = s32; return 32; }
      bytes = s16;
      return s16 <= (size_t)h->max_smem ? 16 : 0;
    };
    g.bigH = pickH(g.maxD, g.ft_big);
    g.smallH = pickH(g.dXcap, g.ft_small);
    if (!g.bigH || !g.smallH)
      return fail("bond capacity %d (D=%d, d*X=%d) exceeds the shared-memory row block of the QR kernel", d, g.maxD, g.dXcap);
    k_op_setup<<<(g.nops + 127) / 128, 128, 0, g.st>>>(g.d_ops, g.nops, L);
    h->n_launch++;
  }
  // ---- sweep 1 (R->L) ----
  for (int t = L - 1; t >= 1; --t) {
    for (auto& g : groups) {
      dim3 g1(g.nops, g.maxq, (g.maxD + KC_RC - 1) / KC_RC);
      ev_begin(h, F_KC, g.st);
      if (g.kc_mma_rb) {
        const int rb = g.kc_mma_rb;
        dim3 gm(g.nops, g.maxq, (g.maxD + rb - 1) / rb);
        const size_t sm = kc_mma_smem_doubles(rb, g.maxD, d, g.maxNyS, g.maxNy) * 8;
        if (rb == 16) k_kron_carry_mma<16><<<gm, NT, sm, g.st>>>(g.d_ops, t, L, g.maxNyS, g.maxNy, h->d_flops + 17);
        else if (rb == 8) k_kron_carry_mma<8><<<gm, NT, sm, g.st>>>(g.d_ops, t, L, g.maxNyS, g.maxNy, h->d_flops + 17);
        else k_kron_carry_mma<4><<<gm, NT, sm, g.st>>>(g.d_ops, t, L, g.maxNyS, g.maxNy, h->d_flops + 17);
      } else if (g.kc_rb == 4) k_kron_carry<4><<<g1, NT, g.kc_smem, g.st>>>(g.d_ops, t, L);
      else k_kron_carry<1><<<g1, NT, g.kc_smem, g.st>>>(g.d_ops, t, L);
      ev_end(h, g.st);
      h->n_launch++;
      ev_begin(h, F_QR, g.st);
      const int nsplit = g.nsplit;
      dim3 gq(g.nops, nsplit);
      if (g.bigH == 64) k_qr_ft<64><<<gq, NT, g.ft_big, g.st>>>(g.d_ops, t, nsplit, h->d_flops, g.split_min);
      else if (g.bigH == 32) k_qr_ft<32><<<gq, NT, g.ft_big, g.st>>>(g.d_ops, t, nsplit, h->d_flops, g.split_min);
      else k_qr_ft<16><<<gq, NT, g.ft_big, g.st>>>(g.d_ops, t, nsplit, h->d_flops, g.split_min);
      h->n_launch++;
      if (nsplit > 1) {
        if (g.bigH == 64) k_qr_ft_merge<64><<<g.nops, NT, g.ft_big, g.st>>>(g.d_ops, t, nsplit, h->d_flops, (int)h->tri_merge, g.split_min);
        else if (g.bigH == 32) k_qr_ft_merge<32><<<g.nops, NT, g.ft_big, g.st>>>(g.d_ops, t, nsplit, h->d_flops, (int)h->tri_merge, g.split_min);
        else k_qr_ft_merge<16><<<g.nops, NT, g.ft_big, g.st>>>(g.d_ops, t, nsplit, h->d_flops, (int)h->tri_merge, g.split_min);
        h->n_launch++;
      }
      ev_end(h, g.st);
    }
  }
  // ---- sweep 2 (L->R) ----
  for (int t = 0; t < L; ++t) {
    for (auto& g : groups) {
      dim3 g3(g.nops, g.maxq, g.maxNy);
      ev_begin(h, F_KP, g.st);
      k_kron_proj<<<g3, NT, kp_smem, g.st>>>(g.d_ops, t);
      ev_end(h, g.st);
      h->n_launch++;
      if (t < L - 1) {
        dim3 g4(g.nops, (g.dXcap + 31) / 32, (g.maxD + 31) / 32);
        ev_begin(h, F_GEMM, g.st);
        k_gemm_m2t<<<g4, NT, 0, g.st>>>(g.d_ops, t);
        ev_end(h, g.st);
        ev_begin(h, F_QRS, g.st);
        if (g.smallH == 64) k_qr_small<64><<<g.nops, NT, g.ft_small, g.st>>>(g.d_ops, t, (int)g.jac_doubles);
        else if (g.smallH == 32) k_qr_small<32><<<g.nops, NT, g.ft_small, g.st>>>(g.d_ops, t, (int)g.jac_doubles);
        else k_qr_small<16><<<g.nops, NT, g.ft_small, g.st>>>(g.d_ops, t, (int)g.jac_doubles);
        ev_end(h, g.st);
        ev_begin(h, F_JAC, g.st);
        k_jacobi_project<<<g.nops, NT, g.jac_smem, g.st>>>(g.d_ops, t, tr, d, (int)g.jac_doubles, h->d_err, h->d_flops + 1, (int)h->svd_mode);
        ev_end(h, g.st);
        h->n_launch += 3;
      } else {
        k_op_last<<<g.nops, NT, 0, g.st>>>(g.d_ops, t);
        h->n_launch++;
      }
    }
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

// scratch of one heavy op from the arena (false: does not fit)
bool alloc_op_scratch(mpbp_state* h, OpDesc& op, int ca, int cb) {
  const int L = h->L, d = h->dmax;
  const int X = op.nyo * op.q;
  if (h->arena.used + op_scratch_bytes(h, ca, cb, X) > h->arena.cap) return false;
  const size_t D = (size_t)ca * cb;
  op.r = (int*)h->arena.take(4 * (L + 1));
  op.Lstride = (long long)(D * D);
  op.Lbuf = (double*)h->arena.take(8 * (size_t)L * D * D);
  op.M = (double*)h->arena.take(8 * op_M_doubles(h, D, X));
  op.Ms = (double*)h->arena.take(8 * (size_t)QR_NSPLIT_MAX * D * D);
  op.G = (double*)h->arena.take(8 * (size_t)d * D * X);
  op.M2T = (double*)h->arena.take(8 * D * (size_t)d * X);
  op.R2 = (double*)h->arena.take(8 * (size_t)(d * X) * (d * X));
  op.Pc[0] = (double*)h->arena.take(8 * (size_t)d * D);
  op.Pc[1] = (double*)h->arena.take(8 * (size_t)d * D);
  return op.r && op.Lbuf && op.M && op.Ms && op.G && op.M2T && op.R2 && op.Pc[0] && op.Pc[1];
}

// Which nodes of a chunk go to the hub lane: the highest degrees, as long as they hold at most hub_frac of the chunk's
// cost.  Their cavity chains (z sequential levels, each 2L sites deep, the heaviest ops of the graph) are the critical
// path of a step; run next to the bulk instead of inside its rounds they no longer stretch every round.
std::vector<char> pick_hub_nodes(const mpbp_state* h, const std::vector<int64_t>& nodes) {
  std::vector<char> hub(nodes.size(), 0);
  if (h->lanes >= 2 && !h->profile && h->inf_k == 0 && (nodes.size() >= 64 || h->lanes >= 100)) {
    // lane mode: EVERY node goes to one of G lanes (longest-processing-time greedy on the node cost); a lane runs the whole
    // cavity DAG of its nodes on its own stream, level after level, with no barrier across lanes
    const int G = std::min((int)h->lanes % 100, 1 + mpbp_state::NAUX);
    std::vector<std::pair<double, size_t>> cost(nodes.size());
    for (size_t k = 0; k < nodes.size(); ++k) {
      const int ci = h->class_of_node[nodes[k]];
      double w = 1.0;
      if (ci >= 0 && ci < (int)h->classes.size() && !h->classes[ci].generic)
        for (double v : node_level_cost(h->classes[ci], h->dmax)) w += v;
      cost[k] = {w, k};
    }
    std::stable_sort(cost.begin(), cost.end(), [](const std::pair<double, size_t>& a, const std::pair<double, size_t>& b) { return a.first > b.first; });
    std::vector<double> load(G, 0.0);
    for (auto& c : cost) {
      const int g = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      hub[c.second] = (char)(g + 1);
      load[g] += c.first;
    }
    return hub;
  }
  if (h->hub_lane <= 0 || (nodes.size() < 64 && h->hub_lane < 2) || h->inf_k > 0) return hub;
  std::map<int, double> cost_by_z;
  std::vector<int> zs(nodes.size(), -1);
  double total = 0.0;
  for (size_t k = 0; k < nodes.size(); ++k) {
    const int ci = h->class_of_node[nodes[k]];
    if (ci < 0 || ci >= (int)h->classes.size() || h->classes[ci].generic) continue;
    const NodeClass& c = h->classes[ci];
    double w = 0.0;
    for (double v : node_level_cost(c, h->dmax)) w += v;
    zs[k] = c.z;
    cost_by_z[c.z] += w;
    total += w;
  }
  int zhub = 1 << 30;
  double acc = 0.0;
  for (auto it = cost_by_z.rbegin(); it != cost_by_z.rend(); ++it) {
    if (it->first < 4 || acc + it->second > h->hub_frac * total) break;
    acc += it->second;
    zhub = it->first;
  }
  size_t nh = 0;
  for (size_t k = 0; k < nodes.size(); ++k)
    if (zs[k] >= zhub) { hub[k] = 1; ++nh; }
  if (nh == nodes.size()) std::fill(hub.begin(), hub.end(), 0);
  return hub;
}

int run_nodes_chunk(mpbp_state* h, const std::vector<int64_t>& nodes, int rb, int wb, const Trunc& tr) {
  const int L = h->L, d = h->dmax;
  h->arena.used = 0;
  Plan P;
  const std::vector<char> hub = pick_hub_nodes(h, nodes);
  if (build_plan(h, nodes, hub, rb, wb, P)) return 1;
  cudaStream_t st = h->st;
  BtJob* d_bt;
  InitJob* d_init;
  FinJob* d_fin;
  BelJob* d_bel;
  FJob* d_fj;
  if (upload_jobs(h, P.bt, &d_bt) || upload_jobs(h, P.init, &d_init) || upload_jobs(h, P.fin, &d_fin) ||
      upload_jobs(h, P.bel, &d_bel) || upload_jobs(h, P.fj, &d_fj))
    return 1;
  if (!P.bt.empty()) {
    dim3 g((unsigned)P.bt.size(), L);
    k_btilde<<<g, NT, 0, st>>>(d_bt, L);
    h->n_launch++;
  }
  if (!P.gen.empty()) {
    GenJob* d_gen;
    if (upload_jobs(h, P.gen, &d_gen)) return 1;
    dim3 gg((unsigned)P.gen.size(), L);
    k_generic_kron<<<gg, NT, 0, st>>>(d_gen, L, d, h->d_err);
    h->n_launch++;
  }
  if (!P.init.empty()) {
    k_init_tt<<<(unsigned)P.init.size(), 64, 0, st>>>(d_init, (int)P.init.size(), L);
    h->n_launch++;
  }
  // ---- lanes: every level of a lane's nodes is enqueued back to back on the lane's stream (no host sync: the lane's scratch
  // region is reused level after level in stream order); the kernels of the lanes are issued interleaved site by site ----
  size_t persistent = h->arena.used;
  if (!P.lanes.empty()) {
    const int NL = (int)P.lanes.size();
    // scratch region of a lane = its largest level
    std::vector<size_t> need(NL, 0);
    size_t need_all = 0, maxlev = 0;
    for (int g = 0; g < NL; ++g) {
      auto& LN = P.lanes[g];
      maxlev = std::max(maxlev, LN.levels.size());
      for (size_t lev = 1; lev < LN.levels.size(); ++lev) {
        size_t nl = sizeof(OpDesc) * LN.levels[lev].size() + 4096;
        for (size_t k = 0; k < LN.levels[lev].size(); ++k)
          nl += op_scratch_bytes(h, LN.capA[lev][k], LN.capB[lev][k], LN.levels[lev][k].nyo * LN.levels[lev][k].q) + 4096;
        need[g] = std::max(need[g], nl);
      }
      need_all += need[g];
    }
    const bool all_lanes = P.levels.empty();  // lane mode proper: nothing left for the bulk rounds
    if (need_all > (size_t)((all_lanes ? 0.95 : 0.4) * (double)(h->arena.cap - persistent))) {
      // does not fit: merge the lanes into the bulk rounds (level by level; the result does not depend on the lane)
      for (int g = 0; g < NL; ++g) {
        auto& LN = P.lanes[g];
        for (size_t lev = 1; lev < LN.levels.size(); ++lev) {
          if (P.levels.size() <= lev) { P.levels.resize(lev + 1); P.capA.resize(lev + 1); P.capB.resize(lev + 1); }
          for (size_t k = 0; k < LN.levels[lev].size(); ++k) {
            P.levels[lev].push_back(LN.levels[lev][k]);
            P.capA[lev].push_back(LN.capA[lev][k]);
            P.capB[lev].push_back(LN.capB[lev][k]);
          }
        }
      }
      P.lanes.clear();
    } else {
      // streams: the hub lane (one lane next to the bulk rounds) runs on its high-priority stream; in lane mode lane g runs
      // on the main stream (g = 0) or an auxiliary one; in profile mode everything runs inline on the main stream
      std::vector<cudaStream_t> ls(NL);
      for (int g = 0; g < NL; ++g) ls[g] = h->profile ? st : (all_lanes ? (g == 0 ? st : h->aux[g - 1]) : h->hub_st);
      // scratch pointers are a pure function of the plan: build the descriptors of EVERY (lane, level) first and upload
      // them with one copy (an async copy from pageable memory synchronises its stream, which would put a host-side barrier
      // between the levels of a lane)
      size_t ndesc = 0;
      for (int g = 0; g < NL; ++g)
        for (size_t lev = 1; lev < P.lanes[g].levels.size(); ++lev) ndesc += P.lanes[g].levels[lev].size();
      h->arena.used = persistent;
      OpDesc* d_all = (OpDesc*)h->arena.take(sizeof(OpDesc) * std::max<size_t>(ndesc, 1));
      if (!d_all) return fail("arena exhausted (lane descriptors)");
      std::vector<size_t> lo(NL);
      size_t acc = h->arena.used;
      for (int g = 0; g < NL; ++g) { lo[g] = acc; acc += need[g]; }
      if (acc > h->arena.cap) return fail("arena exhausted (lane scratch)");
      std::vector<OpDesc> all_desc;
      all_desc.reserve(ndesc);
      std::vector<std::vector<GroupRun>> level_groups(maxlev);
      for (size_t lev = 1; lev < maxlev; ++lev) {
        for (int g = 0; g < NL; ++g) {
          auto& LN = P.lanes[g];
          if (lev >= LN.levels.size() || LN.levels[lev].empty()) continue;
          auto& ops = LN.levels[lev];
          h->arena.used = lo[g];
          bool fit = true;
          GroupRun gr;
          memset(&gr, 0, sizeof gr);
          gr.nops = (int)ops.size(); gr.maxD = 1; gr.maxX = 1; gr.maxNy = 1; gr.maxq = 1; gr.maxNyS = 1;
          gr.st = ls[g];
          // heaviest first (LPT), as in the bulk launches
          std::vector<size_t> idx(ops.size());
          for (size_t k = 0; k < idx.size(); ++k) idx[k] = k;
          auto cost = [&](size_t k) { return (double)LN.capA[lev][k] * LN.capB[lev][k] * ops[k].nyo * ops[k].q; };
          std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return cost(a) > cost(b); });
          gr.d_ops = d_all + all_desc.size();
          for (size_t k = 0; k < idx.size() && fit; ++k) {
            OpDesc op = ops[idx[k]];
            const int ca = LN.capA[lev][idx[k]], cb = LN.capB[lev][idx[k]];
            fit = alloc_op_scratch(h, op, ca, cb);
            all_desc.push_back(op);
            gr.maxD = std::max(gr.maxD, ca * cb);
            gr.maxX = std::max(gr.maxX, op.nyo * op.q);
            gr.maxNy = std::max(gr.maxNy, std::max(op.nyo, std::max(op.ny1, op.ny2)));
            gr.maxq = std::max(gr.maxq, op.q);
            gr.maxNyS = std::max(gr.maxNyS, std::min(op.ny1, op.ny2));
          }
          if (!fit || h->arena.used > lo[g] + need[g]) return fail("arena exhausted in lane %d (internal sizing error)", g);
          const int fill_split = std::max(1, std::min(QR_NSPLIT_MAX, (int)(h->qr_fill / std::max(gr.nops * (all_lanes ? NL : 1), 1))));
          gr.nsplit = std::max(fill_split, std::min(QR_NSPLIT_MAX, (int)h->bulk_split));
          gr.split_min = fill_split >= 2 ? 0 : (int)h->bulk_split_min;
          level_groups[lev].push_back(gr);
          h->n_ops += gr.nops;
        }
      }
      CUDA_OK(cudaMemcpyAsync(d_all, all_desc.data(), sizeof(OpDesc) * all_desc.size(), cudaMemcpyHostToDevice, st));
      CUDA_OK(cudaEventRecord(h->ev_hub_fork, st));
      for (int g = 0; g < NL; ++g)
        if (ls[g] != st) CUDA_OK(cudaStreamWaitEvent(ls[g], h->ev_hub_fork, 0));
      for (size_t lev = 1; lev < maxlev; ++lev) {
        if (level_groups[lev].empty()) continue;
        if (run_op_groups(h, level_groups[lev], tr)) {
          for (int k = 0; k < mpbp_state::NAUX; ++k) cudaStreamSynchronize(h->aux[k]);
          cudaStreamSynchronize(h->hub_st);
          cudaStreamSynchronize(st);
          return 1;
        }
      }
      persistent = acc;
      h->arena.used = persistent;
      if (h->profile) {
        CUDA_OK(cudaStreamSynchronize(st));
        ev_flush(h);
      } else {
        // join: the main stream waits for every lane stream
        for (int g = 0; g < NL; ++g)
          if (ls[g] != st) {
            cudaEvent_t ev = all_lanes ? h->ev_join[g - 1] : h->ev_hub_join;
            CUDA_OK(cudaEventRecord(ev, ls[g]));
            CUDA_OK(cudaStreamWaitEvent(st, ev, 0));
          }
      }
    }
  }
  // ---- cavity levels of the bulk ----
  for (size_t lev = 1; lev < P.levels.size(); ++lev) {
    auto& ops = P.levels[lev];
    {
      // longest-processing-time first: CTAs are dispatched in blockIdx order, so the heaviest matrices
      // (largest D*X) start first and the light ones fill the tail of the launch
      std::vector<size_t> idx(ops.size());
      for (size_t k = 0; k < idx.size(); ++k) idx[k] = k;
      auto cost = [&](size_t k) { return (double)P.capA[lev][k] * P.capB[lev][k] * ops[k].nyo * ops[k].q; };
      std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return cost(a) > cost(b); });
      std::vector<OpDesc> o2(ops.size());
      std::vector<int> a2(ops.size()), b2(ops.size());
      for (size_t k = 0; k < idx.size(); ++k) { o2[k] = ops[idx[k]]; a2[k] = P.capA[lev][idx[k]]; b2[k] = P.capB[lev][idx[k]]; }
      ops.swap(o2);
      P.capA[lev].swap(a2);
      P.capB[lev].swap(b2);
    }
    size_t i0 = 0;
    while (i0 < ops.size()) {
      h->arena.used = persistent;
      size_t i1 = i0;
      // reserve room for the descriptor array first
      OpDesc* d_ops = (OpDesc*)h->arena.take(sizeof(OpDesc) * (ops.size() - i0));
      if (!d_ops) return fail("arena exhausted (op descriptors)");
      while (i1 < ops.size() && (double)(i1 - i0) < h->max_group_ops) {
        OpDesc& op = ops[i1];
        if (!alloc_op_scratch(h, op, P.capA[lev][i1], P.capB[lev][i1])) break;
        ++i1;
      }
      if (i1 == i0) return fail("arena too small for a single op (need %.1f MB)", op_scratch_bytes(h, d, d, ops[i0].nyo * ops[i0].q) / 1e6);
      const int nops = (int)(i1 - i0);
      // deal the (cost-sorted) ops round-robin into up to `nstreams` groups that run on concurrent streams
      const int G = std::max(1, std::min((int)h->nstreams, nops / 48));
      std::vector<std::vector<OpDesc>> gops(G);
      std::vector<GroupRun> groups(G);
      for (int gi = 0; gi < G; ++gi) {
        GroupRun& gr = groups[gi];
        gr.nops = 0; gr.maxD = 1; gr.maxX = 1; gr.maxNy = 1; gr.maxq = 1; gr.maxNyS = 1;
        gr.st = gi == 0 ? st : h->aux[gi - 1];
      }
      // optional: the few ops far above the mean cost of this launch (they come first after the LPT sort) bound the
      // launch by themselves; give them group 0 alone, where the small op count makes the TSQR split kick in
      int nhead = 0;
      if (h->outlier_split > 0 && G >= 2) {
        auto cst = [&](size_t k) { return (double)P.capA[lev][k] * P.capB[lev][k] * ops[k].nyo * ops[k].q; };
        double mean = 0.0;
        for (int k = 0; k < nops; ++k) mean += cst(i0 + k);
        mean /= nops;
        while (nhead < nops && cst(i0 + nhead) >= h->outlier_split * mean) ++nhead;
        if (nhead * 4 > (int)h->qr_fill || nops - nhead < G - 1) nhead = 0;  // too many to be outliers / nothing left for the other groups
      }
      std::vector<int> gchunk(nops, 0);
      if (h->group_mode >= 1) {
        // contiguous chunks of (nearly) equal work over the groups that take the non-outlier ops
        auto cst = [&](size_t k) { return (double)P.capA[lev][k] * P.capB[lev][k] * ops[k].nyo * ops[k].q; };
        const int g0 = nhead > 0 ? 1 : 0, ng = G - g0;
        double tot = 0.0;
        for (int k = nhead; k < nops; ++k) tot += cst(i0 + k);
        double acc = 0.0;
        for (int k = nhead; k < nops; ++k) {
          gchunk[k] = g0 + std::min(ng - 1, (int)(acc / std::max(tot, 1e-300) * ng));
          acc += cst(i0 + k);
        }
      }
      for (int k = 0; k < nops; ++k) {
        const OpDesc& op = ops[i0 + k];
        const int gsel = (h->group_mode >= 1 && k >= nhead) ? gchunk[k] : (nhead > 0 ? (k < nhead ? 0 : 1 + (k - nhead) % (G - 1)) : k % G);
        GroupRun& gr = groups[gsel];
        gops[gsel].push_back(op);
        gr.nops++;
        gr.maxD = std::max(gr.maxD, P.capA[lev][i0 + k] * P.capB[lev][i0 + k]);
        gr.maxX = std::max(gr.maxX, op.nyo * op.q);
        gr.maxNy = std::max(gr.maxNy, std::max(op.nyo, std::max(op.ny1, op.ny2)));
        gr.maxq = std::max(gr.maxq, op.q);
        gr.maxNyS = std::max(gr.maxNyS, std::min(op.ny1, op.ny2));
      }
      {
        size_t off = 0;
        for (int gi = 0; gi < G; ++gi) {
          groups[gi].d_ops = d_ops + off;
          CUDA_OK(cudaMemcpyAsync(d_ops + off, gops[gi].data(), sizeof(OpDesc) * gops[gi].size(), cudaMemcpyHostToDevice, st));
          off += gops[gi].size();
        }
      }
      if (G > 1) {
        CUDA_OK(cudaEventRecord(h->ev_fork, st));
        for (int gi = 1; gi < G; ++gi) CUDA_OK(cudaStreamWaitEvent(h->aux[gi - 1], h->ev_fork, 0));
      }
      // under-filled launch: split every tall matrix (sqrt rule); full launch: only the matrices of >= bulk_split_min * n
      // rows, the long poles of their launch, into at most bulk_split chunks
      const int fill_split = std::max(1, std::min(QR_NSPLIT_MAX, (int)(h->qr_fill / std::max(nops, 1))));
      const int nsplit = std::max(fill_split, std::min(QR_NSPLIT_MAX, (int)h->bulk_split));
      for (int gi = 0; gi < G; ++gi) {
        groups[gi].nsplit = nsplit;
        groups[gi].split_min = fill_split >= 2 ? 0 : (int)h->bulk_split_min;
      }
      if (nhead > 0) {
        groups[0].nsplit = std::max(nsplit, std::min(QR_NSPLIT_MAX, (int)(h->qr_fill / nhead)));
        groups[0].split_min = 0;
      }
      if (run_op_groups(h, groups, tr)) {
        for (int k = 0; k < mpbp_state::NAUX; ++k) cudaStreamSynchronize(h->aux[k]);  // do not leave forked streams running
        cudaStreamSynchronize(st);
        return 1;
      }
      for (int gi = 1; gi < G; ++gi) {
        CUDA_OK(cudaEventRecord(h->ev_join[gi - 1], h->aux[gi - 1]));
        CUDA_OK(cudaStreamWaitEvent(st, h->ev_join[gi - 1], 0));
      }
      // descriptors / scratch are reused by the next group: wait for the stream
      CUDA_OK(cudaStreamSynchronize(st));
      ev_flush(h);
      h->n_ops += nops;
      i0 = i1;
    }
  }
  h->arena.used = persistent;
  // ---- outgoing messages, beliefs, free energy ----
  {
    int qm = h->qmax;
    const int rows_cap = d * qm * qm * qm;
    const int vrows = std::max(64, std::min(QR_MAX_M, rows_cap));
    const int ccap = d * qm * qm;  // max(c, p) of the sweep-B matrices
    const size_t jac_fixed = (size_t)ccap + (ccap + 1) / 2 + 1;
    size_t jac_doubles = std::max<size_t>((size_t)d * qm * qm * d * qm, 2 * (size_t)(d + 1));
    const size_t qrd = qr_shared_doubles(vrows);
    if ((qrd + jac_fixed + jac_doubles) * 8 > (size_t)h->max_smem - 2048) {
      if ((qrd + jac_fixed + 2 * (d + 1)) * 8 > (size_t)h->max_smem - 2048) return fail("finalize kernel: shared memory too small");
      jac_doubles = ((size_t)h->max_smem - 2048) / 8 - qrd - jac_fixed;
    }
    const size_t smem = (qrd + jac_fixed + jac_doubles) * 8;
    if (!P.fin.empty()) {
      ev_begin(h, F_FIN, st);
      k_finalize<<<(unsigned)P.fin.size(), NT, smem, st>>>(d_fin, L, tr, d, vrows, (int)jac_doubles, h->d_err);
      ev_end(h, st);
      h->n_launch++;
    }
    if (!P.damp.empty()) {
      DampJob* d_damp;
      if (upload_jobs(h, P.damp, &d_damp)) return 1;
      const int P2 = qm * qm;
      const int dvrows = std::max(64, std::min(QR_MAX_M, 2 * d * P2));
      const int cpcap = std::max(2 * d, d * P2);
      const size_t dfixed = (size_t)cpcap + (cpcap + 1) / 2 + 1;
      size_t djac = std::max<size_t>((size_t)d * P2 * 2 * d, 2 * (size_t)(d + 1));
      const size_t dqrd = qr_shared_doubles(dvrows);
      if ((dqrd + dfixed + djac) * 8 > (size_t)h->max_smem) djac = (size_t)h->max_smem / 8 - dqrd - dfixed;
      const size_t dsm = (dqrd + dfixed + djac) * 8;
      if (h->inf_k > 0) {
        // the k recomputed messages are damped one after the other against the evolving bp.mu[1] (set_msg! inside the j loop)
        for (size_t j = 0; j < P.damp.size(); ++j) k_damp<<<1, NT, dsm, st>>>(d_damp + j, L, tr, d, dvrows, (int)djac, h->d_err);
        h->n_launch += (double)P.damp.size();
      } else {
        k_damp<<<(unsigned)P.damp.size(), NT, dsm, st>>>(d_damp, L, tr, d, dvrows, (int)djac, h->d_err);
        h->n_launch++;
      }
    }
    const size_t bsm = 2 * (size_t)d * qm * 8;
    if (!P.bel.empty()) {
      ev_begin(h, F_BEL, st);
      k_belief<<<(unsigned)P.bel.size(), NT, bsm, st>>>(d_bel, L, d, h->d_err);
      ev_end(h, st);
    }
    if (!P.gfin.empty()) {
      // generic nodes: the belief comes from the COMPRESSED dummy-neighbour message (overwrites what k_belief wrote)
      FinJob* d_gfin;
      MargJob* d_gmarg;
      if (upload_jobs(h, P.gfin, &d_gfin) || upload_jobs(h, P.gmarg, &d_gmarg)) return 1;
      k_finalize<<<(unsigned)P.gfin.size(), NT, smem, st>>>(d_gfin, L, tr, d, vrows, (int)jac_doubles, h->d_err);
      k_msg_marginals<<<(unsigned)P.gmarg.size(), NT, 2 * (size_t)d * 8, st>>>(d_gmarg, L, d);
      h->n_launch += 2;
    }
    if (h->twovar > 0 && !P.bel.empty()) {
      const size_t tsm = 2 * (size_t)qm * d * qm * 8;
      k_twovar<<<(unsigned)P.bel.size(), NT, tsm, st>>>(d_bel, L, d, h->twovar, qm * qm);
      h->n_launch++;
    }
    if (!P.fj.empty()) k_free_energy<<<(unsigned)(P.fj.size() + 127) / 128, 128, 0, st>>>(d_fj, (int)P.fj.size());
    h->n_launch += 2;
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(st));
  ev_flush(h);
  h->n_edge_updates += (double)P.fin.size();
  return 0;
}

// ---- periodic-in-time MPBP: one CTA per node runs the whole ring update (csrc/periodic.cuh) ----
__global__ void __launch_bounds__(NT) k_periodic_nodes(const mpbp_per::PerNode* nodes) {
  mpbp_per::per_node_update(nodes[blockIdx.x]);
}

__global__ void __launch_bounds__(NT) k_periodic_pairs(const mpbp_per::PerPair* jobs) {
  mpbp_per::per_pair_belief(jobs[blockIdx.x]);
}

int run_nodes_periodic(mpbp_state* h, const std::vector<int64_t>& nodes, int rb, int wb, const Trunc& tr) {
  using namespace mpbp_per;
  if (ensure_arena(h)) return 1;
  const int L = h->L;
  auto as_ptt = [](const TTRef& r) { return PTT{r.data, r.bonds, r.ls, r.stride, r.P}; };
  auto take = [&](size_t bytes) -> void* { return h->arena.take(bytes); };
  std::vector<PerNode> batch;
  auto flush = [&]() -> int {
    if (batch.empty()) return 0;
    PerNode* d_nodes = nullptr;
    CUDA_OK(cudaMalloc((void**)&d_nodes, sizeof(PerNode) * batch.size()));
    cudaError_t e = cudaMemcpyAsync(d_nodes, batch.data(), sizeof(PerNode) * batch.size(), cudaMemcpyHostToDevice, h->st);
    if (e == cudaSuccess) {
      k_periodic_nodes<<<(unsigned)batch.size(), NT, 0, h->st>>>(d_nodes);
      h->n_launch++;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    cudaFree(d_nodes);
    if (e != cudaSuccess) return fail("CUDA error %s in the periodic node update", cudaGetErrorString(e));
    for (const PerNode& nd : batch) h->n_edge_updates += nd.z;
    batch.clear();
    h->arena.used = 0;
    return 0;
  };
  h->arena.used = 0;
  for (size_t k = 0; k < nodes.size(); ++k) {
    const int64_t i = nodes[k];
    const int ci = h->class_of_node[i];
    if (ci < 0 || ci >= (int)h->classes.size()) return fail("node %lld has no factor class", (long long)i);
    const NodeClass& c = h->classes[ci];
    if (c.generic) return fail("the periodic path needs RecursiveBPFactors (node %lld has a generic BPFactor)", (long long)i);
    const int z = c.z, q = c.q;
    if (z > PER_MAXZ) return fail("the periodic path supports degrees up to %d (node %lld has %d)", PER_MAXZ, (long long)i, z);
    if (h->node_deg(i) != z || q != h->q[i]) return fail("node %lld does not match its class (degree / states)", (long long)i);
    PerClassView v;
    v.z = z;
    v.q = q;
    v.qn = c.qn.data();
    v.ny = c.ny.data();
    v.pxy = c.d_pxy;
    v.pxy_off = c.pxy_off.data();
    v.pxy_ts = c.pxy_ts;
    v.pyy = [&c](int d1, int d2, const double** p, size_t* ts) {
      auto it = c.pyy.find({d1, d2});
      if (it == c.pyy.end()) return false;
      *p = c.d_pyy + it->second.first;
      *ts = it->second.second;
      return true;
    };
    v.w = c.d_w;
    v.w_off = c.w_off.data();
    v.w_ts = c.w_ts;
    v.wd = c.d_wd;
    v.wd_ts = c.wd_ts;
    v.minit = c.d_minit;
    v.minit_ts = c.minit_ts;
    for (int attempt = 0;; ++attempt) {
      PerNode nd;
      memset((void*)&nd, 0, sizeof nd);
      const size_t mark = h->arena.used;
      bool ok = per_plan_node(v, L, h->dmax, take, nd);
      // the infinite graph keeps only the last recomputed message (src/infinite_graph.jl): the others go to scratch slots
      for (int j = 0; j < z && ok; ++j) {
        const int64_t eout = h->out_edge(i, j);
        if (h->q[h->dst[eout]] != c.qn[j]) return fail("node %lld neighbour %d: class qn mismatch", (long long)i, j);
        nd.msg_in[j] = as_ptt(msg_ref(h, h->msg[rb], h->rev[eout], c.qn[j] * q));
        nd.psi[j] = h->d_psi + h->psi_off[eout];
        if (h->inf_k > 0 && j < z - 1 && h->damp <= 0.0) {  // (damped: every recomputation is damped into the evolving bp.mu[1])
          PTT sc;
          sc.stride = h->sstride;
          sc.X = q * c.qn[j];
          sc.data = (double*)take(sizeof(double) * h->slot);
          sc.bonds = (int*)take(sizeof(int) * (L + 1));
          sc.ls = (double*)take(sizeof(double));
          ok = ok && sc.data && sc.bonds && sc.ls;
          nd.msg_out[j] = sc;
        } else {
          nd.msg_out[j] = as_ptt(msg_ref(h, h->msg[wb], eout, q * c.qn[j]));
        }
      }
      if (!ok) {
        // arena full: run what is planned, then retry this node on an empty arena (a node that does not fit alone is an error)
        h->arena.used = mark;
        if (batch.empty() || attempt > 0) return fail("arena too small for one periodic node update (dmax=%d, degree %d)", h->dmax, z);
        if (flush()) return 1;
        continue;
      }
      nd.tr = PTrunc{tr.kind, tr.d, tr.eps};
      nd.damp = h->damp;
      nd.phi = h->d_phi + h->phi_off[i];
      nd.marg = h->d_marg + h->marg_off[i];
      nd.logzi = h->d_logzi + i;
      nd.logzij = h->d_logzij + h->lz_off[i];
      nd.f = h->d_f + i;
      if (h->twovar > 0 && h->d_tv) {
        nd.tv = h->d_tv + (size_t)i * L * L * h->qmax * h->qmax;
        nd.tv_maxdist = h->twovar;
        nd.tv_q2cap = h->qmax * h->qmax;
      }
      nd.err = h->d_err;
      batch.push_back(nd);
      break;
    }
  }
  return flush();
}

// update a set of pairwise independent-or-double-buffered nodes, chunked by arena capacity
int run_nodes(mpbp_state* h, const std::vector<int64_t>& nodes, int rb, int wb, const Trunc& tr) {
  if (h->periodic) return run_nodes_periodic(h, nodes, rb, wb, tr);
  if (ensure_arena(h)) return 1;
  // fraction of the arena given to persistent per-node storage; the rest is op scratch
  const size_t budget = (size_t)(h->arena.cap * 0.45);
  size_t i0 = 0;
  while (i0 < nodes.size()) {
    size_t used = 0, i1 = i0;
    while (i1 < nodes.size()) {
      const size_t nb = node_bytes(h, nodes[i1]) + 4096;
      if (used + nb > budget && i1 > i0) break;
      used += nb;
      ++i1;
    }
    std::vector<int64_t> chunk(nodes.begin() + i0, nodes.begin() + i1);
    if (run_nodes_chunk(h, chunk, rb, wb, tr)) return 1;
    i0 = i1;
  }
  return 0;
}

int check_err(mpbp_state* h) {
  int e = 0;
  CUDA_OK(cudaMemcpy(&e, h->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) {
    int zero = 0;
    cudaMemcpy(h->d_err, &zero, sizeof(int), cudaMemcpyHostToDevice);
    return fail("device error flags 0x%x:%s%s%s", e, (e & ERR_BOND_OVERFLOW) ? " bond dimension exceeds dmax" : "",
                (e & ERR_NAN) ? " NaN/non-positive normalisation in tensor train" : "",
                (e & ERR_JACOBI_NOCONV) ? " Jacobi SVD did not converge" : "");
  }
  return 0;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* mpbp_last_error(void) { return g_err.c_str(); }
int mpbp_version(void) { return 100; }

int mpbp_create(int64_t N, int64_t E2, int T, const int32_t* q, const int64_t* colptr, const int64_t* dst,
                const int64_t* rev, int dmax, int device, mpbp_handle* out) {
  if (!out) return fail("null out");
  if (N <= 0 || E2 < 0 || T < 0 || dmax < 1) return fail("invalid sizes N=%lld E2=%lld T=%d dmax=%d", (long long)N, (long long)E2, T, dmax);
  if (dmax > 30) return fail("dmax=%d exceeds the supported bond capacity 30 (shared-memory tiling of the contraction kernels)", dmax);
  mpbp_state* h = new mpbp_state();
  h->device = device;
  h->N = N;
  h->E2 = E2;
  h->T = T;
  h->L = T + 1;
  h->dmax = dmax;
  h->q.assign(q, q + N);
  h->colptr.assign(colptr, colptr + N + 1);
  h->dst.assign(dst, dst + E2);
  h->rev.assign(rev, rev + E2);
  h->src.resize(E2);
  for (int64_t i = 0; i < N; ++i) {
    if (h->q[i] < 1 || h->q[i] > 8) { delete h; return fail("q[%lld]=%d out of range 1..8", (long long)i, q[i]); }
    for (int64_t e = colptr[i]; e < colptr[i + 1]; ++e) h->src[e] = i;
  }
  if (h->colptr[0] != 0 || h->colptr[N] != E2) { delete h; return fail("colptr does not span the edges"); }
  for (int64_t e = 0; e < E2; ++e) {
    const int64_t r = rev[e];
    if (r < 0 || r >= E2 || h->rev[r] != e || h->src[r] != h->dst[e] || h->dst[r] != h->src[e]) {
      delete h;
      return fail("rev[%lld] is not the reverse edge (graph must be symmetric, src/mpbp.jl:18)", (long long)e);
    }
  }
  h->class_of_node.assign(N, -1);
  if (common_init(h)) { delete h; return 1; }
  *out = h;
  return 0;
}

int mpbp_create_periodic(int64_t N, int64_t E2, int T, const int32_t* q, const int64_t* colptr, const int64_t* dst,
                         const int64_t* rev, int dmax, int device, mpbp_handle* out) {
  if (dmax > 16) return fail("dmax=%d exceeds the bond capacity 16 of the periodic path (its Kronecker workspace holds bond dmax^2)", dmax);
  if (mpbp_create(N, E2, T, q, colptr, dst, rev, dmax, device, out)) return 1;
  (*out)->periodic = true;  // flat_periodic_mpem2 with d = 1 is the flat open message: nothing else to initialise
  return 0;
}

int mpbp_create_infinite(int k, int T, int q, int dmax, int device, mpbp_handle* out) {
  if (!out) return fail("null out");
  if (k < 1 || T < 0 || dmax < 1 || q < 1 || q > 8) return fail("invalid arguments");
  if (dmax > 30) return fail("dmax=%d exceeds the supported bond capacity 30 (shared-memory tiling of the contraction kernels)", dmax);
  mpbp_state* h = new mpbp_state();
  h->device = device;
  h->N = 1;
  h->E2 = 1;
  h->T = T;
  h->L = T + 1;
  h->dmax = dmax;
  h->inf_k = k;
  h->inf_deg = {k};
  h->q = {q};
  h->colptr = {0, 1};
  h->dst = {0};
  h->rev = {0};
  h->src = {0};
  h->class_of_node.assign(1, -1);
  if (common_init(h)) { delete h; return 1; }
  *out = h;
  return 0;
}

// InfiniteBipartiteRegularGraph((kA, kB)) (src/infinite_graph.jl:62-122): two node classes, node 0 (degree kA, q = qA) and
// node 1 (degree kB, q = qB), one message per direction: edge 0 = (0 -> 1), edge 1 = (1 -> 0).  Node i sees inf_deg[i]
// copies of its single incoming message; of its recomputed outgoing messages the last one stays, as in the reference.
int mpbp_create_infinite_bipartite(int kA, int kB, int T, int qA, int qB, int dmax, int device, mpbp_handle* out) {
  if (!out) return fail("null out");
  if (kA < 1 || kB < 1 || T < 0 || dmax < 1 || qA < 1 || qA > 8 || qB < 1 || qB > 8) return fail("invalid arguments");
  if (dmax > 30) return fail("dmax=%d exceeds the supported bond capacity 30 (shared-memory tiling of the contraction kernels)", dmax);
  mpbp_state* h = new mpbp_state();
  h->device = device;
  h->N = 2;
  h->E2 = 2;
  h->T = T;
  h->L = T + 1;
  h->dmax = dmax;
  h->inf_k = std::max(kA, kB);
  h->inf_deg = {kA, kB};
  h->q = {qA, qB};
  h->colptr = {0, 1, 2};
  h->dst = {1, 0};
  h->rev = {1, 0};
  h->src = {0, 1};
  h->class_of_node.assign(2, -1);
  if (common_init(h)) { delete h; return 1; }
  *out = h;
  return 0;
}

int mpbp_destroy(mpbp_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->st);
  for (auto& c : h->classes) {
    cudaFree(c.d_pxy); cudaFree(c.d_pyy); cudaFree(c.d_w); cudaFree(c.d_wd); cudaFree(c.d_minit);
  }
  for (int b = 0; b < 2; ++b) { cudaFree(h->msg[b].data); cudaFree(h->msg[b].bonds); cudaFree(h->msg[b].ls); }
  cudaFree(h->d_phi); cudaFree(h->d_psi); cudaFree(h->d_qprod); cudaFree(h->d_marg); cudaFree(h->d_logzi);
  cudaFree(h->d_logzij); cudaFree(h->d_f); cudaFree(h->d_means); cudaFree(h->d_marg_off); cudaFree(h->d_q);
  cudaFree(h->d_delta); cudaFree(h->d_err); cudaFree(h->d_flops); cudaFree(h->d_edge_idx); cudaFree(h->arena.base); cudaFree(h->d_tv);
  for (auto& e : h->ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  if (h->own_stream) cudaStreamDestroy(h->st);
  for (int k = 0; k < mpbp_state::NAUX; ++k) { if (h->aux[k]) cudaStreamDestroy(h->aux[k]); if (h->ev_join[k]) cudaEventDestroy(h->ev_join[k]); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->hub_st) cudaStreamDestroy(h->hub_st);
  if (h->ev_hub_fork) cudaEventDestroy(h->ev_hub_fork);
  if (h->ev_hub_join) cudaEventDestroy(h->ev_hub_join);
  delete h;
  return 0;
}

int mpbp_add_node_class(mpbp_handle h, int z, int q, const int32_t* qn, int nt, const int32_t* ny, const double* pxy,
                        int npairs, const int32_t* pair_d1, const int32_t* pair_d2, const double* pyy,
                        const double* w, const double* wd, const double* minit, int32_t* class_id) {
  if (!h) return fail("null handle");
  if (z < 0) return fail("z must be >= 0");
  if (nt != 1 && nt != h->L) return fail("nt must be 1 or T+1");
  CUDA_OK(cudaSetDevice(h->device));
  NodeClass c;
  c.z = z;
  c.q = q;
  c.nt = nt;
  if (z > 0) c.qn.assign(qn, qn + z);
  c.ny.assign(ny, ny + z + 1);
  for (int l = 0; l <= z; ++l)
    if (c.ny[l] < 1 || c.ny[l] > 64) return fail("nstates(w,%d)=%d out of range", l, c.ny[l]);
  const bool td = nt > 1;
  // pxy
  size_t tot = 0;
  c.pxy_off.resize(z);
  for (int k = 0; k < z; ++k) { c.pxy_off[k] = tot; tot += (size_t)c.ny[1] * qn[k] * q; }
  c.pxy_ts = td ? tot : 0;
  if (upload(&c.d_pxy, pxy, tot * nt)) return 1;
  // pyy
  size_t off = 0;
  for (int p = 0; p < npairs; ++p) {
    const int d1 = pair_d1[p], d2 = pair_d2[p];
    if (d1 < 0 || d2 < 0 || d1 + d2 > z) return fail("invalid (d1,d2)=(%d,%d)", d1, d2);
    const size_t sz = (size_t)c.ny[d1 + d2] * c.ny[d1] * c.ny[d2] * q;
    c.pyy[{d1, d2}] = {off, td ? sz : 0};
    off += sz * nt;
  }
  if (upload(&c.d_pyy, pyy, off)) return 1;
  // w
  tot = 0;
  c.w_off.resize(z);
  for (int j = 0; j < z; ++j) { c.w_off[j] = tot; tot += (size_t)q * q * qn[j] * c.ny[z - 1]; }
  c.w_ts = td ? tot : 0;
  if (upload(&c.d_w, w, tot * nt)) return 1;
  const size_t wds = (size_t)q * q * c.ny[z];
  c.wd_ts = td ? wds : 0;
  if (upload(&c.d_wd, wd, wds * nt)) return 1;
  const size_t mis = (size_t)c.ny[0] * q;
  c.minit_ts = td ? mis : 0;
  // Minit must cover all L sites even when time-independent: expand on the host
  std::vector<double> mi((size_t)h->L * mis);
  for (int t = 0; t < h->L; ++t) memcpy(mi.data() + t * mis, minit + (td ? t * mis : 0), mis * 8);
  c.minit_ts = mis;
  if (upload(&c.d_minit, mi.data(), mi.size())) return 1;
  h->classes.push_back(c);
  if (class_id) *class_id = (int)h->classes.size() - 1;
  return 0;
}

int mpbp_add_generic_class(mpbp_handle h, int z, int q, const int32_t* qn, int nt, const double* wtab, int32_t* class_id) {
  if (!h || !wtab) return fail("null argument");
  if (z < 1 || z > GEN_MAXZ) return fail("generic BPFactor classes support degrees 1..%d (the trace is exponential in z)", GEN_MAXZ);
  if (nt != 1 && nt != h->L) return fail("nt must be 1 or T+1");
  CUDA_OK(cudaSetDevice(h->device));
  NodeClass c;
  c.generic = true;
  c.z = z;
  c.q = q;
  c.nt = nt;
  c.qn.assign(qn, qn + z);
  const bool td = nt > 1;
  size_t Yall = 1;
  for (int k = 0; k < z; ++k) Yall *= qn[k];
  const size_t wsz = (size_t)q * Yall * q;  // [x', x_0..x_{z-1}, x]
  c.gen_ny.resize(z + 1);
  for (int j = 0; j < z; ++j) c.gen_ny[j] = (int)(Yall / qn[j]);
  c.gen_ny[z] = (int)Yall;
  // W_j[x' + q*(x + q*(xj + qj*y))], y = (x_k)_{k != j} first fastest ;  Wd[x' + q*(x + q*yall)]
  c.w_off.resize(z);
  size_t tot = 0;
  for (int j = 0; j < z; ++j) { c.w_off[j] = tot; tot += (size_t)q * q * qn[j] * c.gen_ny[j]; }
  c.w_ts = td ? tot : 0;
  std::vector<double> W(tot * nt), Wd((size_t)q * q * Yall * nt);
  for (int t = 0; t < nt; ++t) {
    const double* wt = wtab + (size_t)t * wsz;
    std::vector<int> xs(z);
    for (size_t yall = 0; yall < Yall; ++yall) {
      size_t r = yall;
      for (int k = 0; k < z; ++k) { xs[k] = (int)(r % qn[k]); r /= qn[k]; }
      for (int x = 0; x < q; ++x)
        for (int xn = 0; xn < q; ++xn) {
          const double v = wt[xn + (size_t)q * (yall + Yall * x)];
          Wd[(size_t)t * q * q * Yall + xn + q * (x + (size_t)q * yall)] = v;
          for (int j = 0; j < z; ++j) {
            size_t y = 0, mul = 1;
            for (int k = 0; k < z; ++k)
              if (k != j) { y += mul * xs[k]; mul *= qn[k]; }
            W[(size_t)t * tot + c.w_off[j] + xn + q * (x + (size_t)q * (xs[j] + (size_t)qn[j] * y))] = v;
          }
        }
    }
  }
  if (upload(&c.d_w, W.data(), W.size())) return 1;
  c.wd_ts = td ? (size_t)q * q * Yall : 0;
  if (upload(&c.d_wd, Wd.data(), Wd.size())) return 1;
  h->classes.push_back(c);
  if (class_id) *class_id = (int)h->classes.size() - 1;
  return 0;
}

int mpbp_clear_node_classes(mpbp_handle h) {
  if (!h) return fail("null handle");
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamSynchronize(h->st));
  for (auto& c : h->classes) {
    cudaFree(c.d_pxy); cudaFree(c.d_pyy); cudaFree(c.d_w); cudaFree(c.d_wd); cudaFree(c.d_minit);
  }
  h->classes.clear();
  std::fill(h->class_of_node.begin(), h->class_of_node.end(), -1);
  return 0;
}

int mpbp_set_node_classes(mpbp_handle h, const int32_t* cls) {
  if (!h) return fail("null handle");
  for (int64_t i = 0; i < h->N; ++i) {
    if (cls[i] < -1 || cls[i] >= (int)h->classes.size()) return fail("class_of_node[%lld]=%d unknown", (long long)i, cls[i]);
    h->class_of_node[i] = cls[i];
  }
  return 0;
}

int mpbp_set_phi(mpbp_handle h, const double* phi) {
  if (!h || !phi) return fail("null argument");
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaMemcpyAsync(h->d_phi, phi, sizeof(double) * h->phi_off[h->N], cudaMemcpyHostToDevice, h->st));
  CUDA_OK(cudaStreamSynchronize(h->st));
  return 0;
}
int mpbp_set_psi(mpbp_handle h, const double* psi) {
  if (!h || !psi) return fail("null argument");
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaMemcpyAsync(h->d_psi, psi, sizeof(double) * h->psi_off[h->E2], cudaMemcpyHostToDevice, h->st));
  CUDA_OK(cudaStreamSynchronize(h->st));
  return 0;
}

int mpbp_get_message(mpbp_handle h, int64_t e, int32_t* bonds, double* data, int64_t cap, int64_t* needed) {
  if (!h || e < 0 || e >= h->E2) return fail("bad edge index");
  CUDA_OK(cudaSetDevice(h->device));
  const int L = h->L;
  const MsgStore& m = h->msg[h->cur];
  std::vector<int> b(L + 1);
  CUDA_OK(cudaMemcpy(b.data(), m.bonds + e * (L + 1), sizeof(int) * (L + 1), cudaMemcpyDeviceToHost));
  const int P = h->q[h->src[e]] * h->q[h->dst[e]];
  int64_t tot = 0;
  for (int t = 0; t < L; ++t) tot += (int64_t)b[t] * b[t + 1] * P;
  if (needed) *needed = tot;
  if (bonds) memcpy(bonds, b.data(), sizeof(int) * (L + 1));
  if (!data) return 0;
  if (cap < tot) return fail("buffer too small: need %lld doubles", (long long)tot);
  double ls;
  CUDA_OK(cudaMemcpy(&ls, m.ls + e, sizeof(double), cudaMemcpyDeviceToHost));
  const double f = std::exp(ls / L);
  int64_t off = 0;
  for (int t = 0; t < L; ++t) {
    const int64_t n = (int64_t)b[t] * b[t + 1] * P;
    CUDA_OK(cudaMemcpy(data + off, m.data + e * h->slot + (int64_t)t * h->sstride, sizeof(double) * n, cudaMemcpyDeviceToHost));
    for (int64_t k = 0; k < n; ++k) data[off + k] *= f;
    off += n;
  }
  return 0;
}

int mpbp_set_message(mpbp_handle h, int64_t e, const int32_t* bonds, const double* data) {
  if (!h || e < 0 || e >= h->E2 || !bonds || !data) return fail("bad argument");
  CUDA_OK(cudaSetDevice(h->device));
  const int L = h->L;
  if (h->periodic) {
    if (bonds[0] != bonds[L]) return fail("periodic message: the first and the last bond must coincide (src/mpems.jl:100-101)");
  } else if (bonds[0] != 1 || bonds[L] != 1) return fail("first/last bond must be 1 (src/mpems.jl:41)");
  for (int t = 0; t <= L; ++t)
    if (bonds[t] < 1 || bonds[t] > h->dmax) return fail("bond %d exceeds dmax=%d", bonds[t], h->dmax);
  MsgStore& m = h->msg[h->cur];
  const int P = h->q[h->src[e]] * h->q[h->dst[e]];
  int64_t off = 0;
  for (int t = 0; t < L; ++t) {
    const int64_t n = (int64_t)bonds[t] * bonds[t + 1] * P;
    CUDA_OK(cudaMemcpy(m.data + e * h->slot + (int64_t)t * h->sstride, data + off, sizeof(double) * n, cudaMemcpyHostToDevice));
    off += n;
  }
  std::vector<int> b(bonds, bonds + L + 1);
  CUDA_OK(cudaMemcpy(m.bonds + e * (L + 1), b.data(), sizeof(int) * (L + 1), cudaMemcpyHostToDevice));
  const double zero = 0.0;
  CUDA_OK(cudaMemcpy(m.ls + e, &zero, sizeof(double), cudaMemcpyHostToDevice));
  return 0;
}

int mpbp_reset_messages(mpbp_handle h) {
  if (!h) return fail("null handle");
  CUDA_OK(cudaSetDevice(h->device));
  if (flat_messages(h, h->msg[h->cur])) return 1;
  CUDA_OK(cudaStreamSynchronize(h->st));
  return 0;
}

int mpbp_iterate(mpbp_handle h, int maxiter, int trunc_kind, int trunc_d, double trunc_eps, double tol, double damp,
                 int schedule, const int64_t* nodes_in, int64_t n_nodes, const int64_t* order, const double* obs,
                 int* iters, double* deltas) {
  if (!h) return fail("null handle");
  if (!(damp >= 0.0 && damp < 1.0)) return fail("damp must satisfy 0 <= damp < 1 (src/recursive_bp_factor.jl:169)");
  if (h->L > DAMP_MAXL && damp > 0.0) return fail("damping supports T+1 <= %d", DAMP_MAXL);
  h->damp = damp;
  if (trunc_kind < 0 || trunc_kind > 2) return fail("unknown truncation kind %d", trunc_kind);
  if ((trunc_kind == MPBP_TRUNC_BOND || trunc_kind == MPBP_TRUNC_BOND_THRESH) && trunc_d > h->dmax)
    return fail("TruncBond(%d) exceeds the device bond capacity dmax=%d", trunc_d, h->dmax);
  if (trunc_d < 1 && trunc_kind != MPBP_TRUNC_THRESH) return fail("TruncBond(d) needs d >= 1");
  CUDA_OK(cudaSetDevice(h->device));
  Trunc tr{trunc_kind, trunc_d, trunc_eps};
  std::vector<int64_t> base;
  if (nodes_in) base.assign(nodes_in, nodes_in + n_nodes);
  else { base.resize(h->N); for (int64_t i = 0; i < h->N; ++i) base[i] = i; n_nodes = h->N; }
  for (int64_t i : base) if (i < 0 || i >= h->N) return fail("node index %lld out of range", (long long)i);
  struct DevBuf {  // freed on every exit path (the loud device errors below are expected in normal use)
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
  } obs_buf, nodes_buf;
  double* d_obs = nullptr;
  if (obs) { if (upload(&d_obs, obs, (size_t)h->N * h->qmax)) return 1; }
  obs_buf.p = d_obs;
  int64_t* d_nodes = nullptr;
  if (upload(&d_nodes, base.data(), base.size())) return 1;
  nodes_buf.p = d_nodes;
  if (schedule == MPBP_SCHEDULE_PARALLEL && !h->msg[1].data) {
    if (alloc_msg_store(h, h->msg[1])) return 1;
  }
  // CB_BP builds its baseline m = means(f, bp) from the beliefs as they are NOW and with the caller's f
  // (src/mpbp.jl:165-171): recompute it here, the delta of this pass is discarded
  if (!base.empty()) {
    k_means_delta<<<(unsigned)base.size(), 64, 0, h->st>>>(h->d_marg, h->d_marg_off, h->d_q, d_obs, h->qmax, d_nodes,
                                                           (long long)base.size(), h->L, h->d_means, h->d_delta);
    h->n_launch++;
  }
  int done = 0;
  for (int it = 0; it < maxiter; ++it) {
    std::vector<int64_t> ord = order ? std::vector<int64_t>(order + (size_t)it * n_nodes, order + (size_t)(it + 1) * n_nodes) : base;
    if (schedule == MPBP_SCHEDULE_PARALLEL) {
      const int rb = h->cur, wb = 1 - h->cur;
      // messages of nodes outside `nodes` are carried over unchanged
      CUDA_OK(cudaMemcpyAsync(h->msg[wb].data, h->msg[rb].data, sizeof(double) * h->slot * h->E2, cudaMemcpyDeviceToDevice, h->st));
      CUDA_OK(cudaMemcpyAsync(h->msg[wb].bonds, h->msg[rb].bonds, sizeof(int) * (h->L + 1) * h->E2, cudaMemcpyDeviceToDevice, h->st));
      CUDA_OK(cudaMemcpyAsync(h->msg[wb].ls, h->msg[rb].ls, sizeof(double) * h->E2, cudaMemcpyDeviceToDevice, h->st));
      if (run_nodes(h, ord, rb, wb, tr)) return 1;
      h->cur = wb;
    } else {
      // level scheduling of the in-place sweep: level(i) = 1 + max level of earlier-visited neighbours
      std::vector<int> pos(h->N, -1), level(h->N, 0);
      for (size_t k = 0; k < ord.size(); ++k) pos[ord[k]] = (int)k;
      int nlev = 0;
      std::vector<std::vector<int64_t>> levels;
      for (size_t k = 0; k < ord.size(); ++k) {
        const int64_t i = ord[k];
        int lv = 0;
        // (infinite graphs: the single neighbour class is the destination of the node's one edge)
          for (int64_t e = h->colptr[i]; e < h->colptr[i + 1]; ++e) {
            const int64_t j = h->dst[e];
            if (pos[j] >= 0 && pos[j] < (int)k) lv = std::max(lv, level[j] + 1);
          }
        level[i] = lv;
        if (lv >= nlev) { nlev = lv + 1; levels.resize(nlev); }
        levels[lv].push_back(i);
      }
      for (auto& lvn : levels)
        if (run_nodes(h, lvn, h->cur, h->cur, tr)) return 1;
    }
    if (check_err(h)) return 1;
    // CB_BP: Delta = max |means_new - means_old| (src/mpbp.jl:174-183)
    CUDA_OK(cudaMemsetAsync(h->d_delta, 0, sizeof(double), h->st));
    if (!base.empty()) {
      k_means_delta<<<(unsigned)base.size(), 64, 0, h->st>>>(h->d_marg, h->d_marg_off, h->d_q, d_obs, h->qmax, d_nodes,
                                                             (long long)base.size(), h->L, h->d_means, h->d_delta);
      h->n_launch++;
    }
    double delta = 0;
    CUDA_OK(cudaMemcpyAsync(&delta, h->d_delta, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CUDA_OK(cudaStreamSynchronize(h->st));
    if (deltas) deltas[it] = delta;
    done = it + 1;
    if (delta < tol) break;
  }
  ev_flush(h);
  if (iters) *iters = done;
  return 0;
}

int mpbp_beliefs(mpbp_handle h, double* out) {
  if (!h || !out) return fail("null argument");
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaMemcpy(out, h->d_marg, sizeof(double) * h->marg_off[h->N], cudaMemcpyDeviceToHost));
  return 0;
}

int mpbp_twovar_marginals(mpbp_handle h, double* out) {
  if (!h || !out) return fail("null argument");
  if (h->twovar <= 0 || !h->d_tv) return fail("two-time marginals are off: mpbp_set_option(h, \"twovar\", maxdist) before iterating");
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaMemcpy(out, h->d_tv, sizeof(double) * (size_t)h->N * h->L * h->L * h->qmax * h->qmax, cudaMemcpyDeviceToHost));
  return 0;
}

int mpbp_free_energy(mpbp_handle h, double* f) {
  if (!h || !f) return fail("null argument");
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaMemcpy(f, h->d_f, sizeof(double) * h->N, cudaMemcpyDeviceToHost));
  return 0;
}

// pair_beliefs(bp) on periodic messages: one CTA per directed edge (csrc/periodic.cuh: per_pair_belief)
static int pair_beliefs_periodic(mpbp_state* h, double* out, double* logz) {
  using namespace mpbp_per;
  CUDA_OK(cudaSetDevice(h->device));
  if (ensure_arena(h)) return 1;
  const int L = h->L, D2 = h->dmax * h->dmax;
  const MsgStore& m = h->msg[h->cur];
  double* d_out;
  double* d_lz;
  CUDA_OK(cudaMalloc((void**)&d_out, sizeof(double) * std::max<int64_t>(h->psi_off[h->E2], 1)));
  CUDA_OK(cudaMalloc((void**)&d_lz, sizeof(double) * std::max<int64_t>(h->E2, 1)));
  struct Guard { double *a, *b; ~Guard() { cudaFree(a); cudaFree(b); } } guard{d_out, d_lz};
  const size_t per = 8 * ((size_t)(2 * L + 3) * D2 * D2 + PER_MAXW + 8) + 1024;
  int64_t e0 = 0;
  while (e0 < h->E2) {
    h->arena.used = 0;
    const int64_t maxn = std::max<int64_t>(1, (int64_t)((h->arena.cap / 2) / (per + sizeof(PerPair))));
    const int64_t e1 = std::min(h->E2, e0 + maxn);
    std::vector<PerPair> jobs;
    for (int64_t e = e0; e < e1; ++e) {
      PerPair jb;
      memset((void*)&jb, 0, sizeof jb);
      const int qs = h->q[h->src[e]], qd = h->q[h->dst[e]];
      const TTRef a = msg_ref(h, m, e, qs * qd), b = msg_ref(h, m, h->rev[e], qs * qd);
      jb.a = PTT{a.data, a.bonds, a.ls, a.stride, a.P};
      jb.b = PTT{b.data, b.bonds, b.ls, b.stride, b.P};
      jb.psi = h->d_psi + h->psi_off[e];
      jb.qi = qs;
      jb.qj = qd;
      jb.L = L;
      jb.out = d_out + h->psi_off[e];
      jb.logz = d_lz + e;
      jb.wcap = D2;
      jb.tm = (double*)h->arena.take(8 * (size_t)(2 * L + 3) * D2 * D2);
      jb.red = (double*)h->arena.take(8 * (PER_MAXW + 8));
      jb.err = h->d_err;
      if (!jb.tm || !jb.red) return fail("arena exhausted (periodic pair beliefs)");
      jobs.push_back(jb);
    }
    PerPair* d_jobs;
    if (upload_jobs(h, jobs, &d_jobs)) return 1;
    k_periodic_pairs<<<(unsigned)jobs.size(), NT, 0, h->st>>>(d_jobs);
    h->n_launch++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(h->st));
    e0 = e1;
  }
  if (check_err(h)) return 1;
  CUDA_OK(cudaMemcpy(out, d_out, sizeof(double) * h->psi_off[h->E2], cudaMemcpyDeviceToHost));
  std::vector<double> lz(std::max<int64_t>(h->E2, 1));
  CUDA_OK(cudaMemcpy(lz.data(), d_lz, sizeof(double) * h->E2, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < h->N; ++i) logz[i] = 0.0;
  if (h->inf_k > 0) {
    for (int64_t e = 0; e < h->E2; ++e) logz[h->dst[e]] = (1.0 / (h->inf_deg[h->dst[e]] - 1) - 0.5) * lz[e];
  } else {
    for (int64_t e = 0; e < h->E2; ++e) {
      const int64_t j = h->src[e];
      const double dj = (double)(h->colptr[j + 1] - h->colptr[j]);
      logz[j] += (1.0 / dj - 0.5) * lz[e];
    }
  }
  return 0;
}

int mpbp_pair_beliefs(mpbp_handle h, double* out, double* logz) {
  if (!h || !out || !logz) return fail("null argument");
  if (h->periodic) return pair_beliefs_periodic(h, out, logz);
  CUDA_OK(cudaSetDevice(h->device));
  if (ensure_arena(h)) return 1;
  const int L = h->L, d = h->dmax;
  const MsgStore& m = h->msg[h->cur];
  double* d_out;
  double* d_lz;
  CUDA_OK(cudaMalloc((void**)&d_out, sizeof(double) * h->psi_off[h->E2]));
  CUDA_OK(cudaMalloc((void**)&d_lz, sizeof(double) * h->E2));
  const size_t per = 8 * (size_t)(L + 1) * d * d + 512;
  int64_t e0 = 0;
  while (e0 < h->E2) {
    h->arena.used = 0;
    const int64_t maxn = std::max<int64_t>(1, (int64_t)((h->arena.cap / 2) / (per + sizeof(PairJob))));
    const int64_t e1 = std::min(h->E2, e0 + maxn);
    std::vector<PairJob> jobs;
    for (int64_t e = e0; e < e1; ++e) {
      PairJob jb;
      const int qs = h->q[h->src[e]], qd = h->q[h->dst[e]];
      jb.A = msg_ref(h, m, e, qs * qd);
      jb.B = msg_ref(h, m, h->rev[e], qs * qd);
      jb.psi = h->d_psi + h->psi_off[e];
      jb.qs = qs;
      jb.qd = qd;
      jb.out = d_out + h->psi_off[e];
      jb.logz = d_lz + e;
      jb.Renv = (double*)h->arena.take(8 * (size_t)(L + 1) * d * d);
      if (!jb.Renv) return fail("arena exhausted (pair beliefs)");
      jobs.push_back(jb);
    }
    PairJob* d_jobs;
    if (upload_jobs(h, jobs, &d_jobs)) return 1;
    k_pair_belief<<<(unsigned)jobs.size(), NT, 3 * (size_t)d * d * 8, h->st>>>(d_jobs, L, d);
    h->n_launch++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(h->st));
    e0 = e1;
  }
  CUDA_OK(cudaMemcpy(out, d_out, sizeof(double) * h->psi_off[h->E2], cudaMemcpyDeviceToHost));
  std::vector<double> lz(h->E2);
  CUDA_OK(cudaMemcpy(lz.data(), d_lz, sizeof(double) * h->E2, cudaMemcpyDeviceToHost));
  cudaFree(d_out);
  cudaFree(d_lz);
  // logz[j] += (1/d_j - 1/2) log z_ij   (src/mpbp.jl:230 ; infinite graph: src/infinite_graph.jl:40)
  for (int64_t i = 0; i < h->N; ++i) logz[i] = 0.0;
  if (h->inf_k > 0) {
    // src/infinite_graph.jl:37-43,110-118: logz[i] = (1/(k_i - 1) - 1/2) log z of the pair built on the message INTO i
    for (int64_t e = 0; e < h->E2; ++e) logz[h->dst[e]] = (1.0 / (h->inf_deg[h->dst[e]] - 1) - 0.5) * lz[e];
  } else {
    for (int64_t e = 0; e < h->E2; ++e) {
      const int64_t j = h->src[e];
      const double dj = (double)(h->colptr[j + 1] - h->colptr[j]);
      logz[j] += (1.0 / dj - 0.5) * lz[e];
    }
  }
  return 0;
}

int mpbp_alternate_marginals(mpbp_handle h, double* out) {
  if (!h || !out) return fail("null argument");
  if (h->periodic) return fail("alternate marginals are not available on the periodic path");
  CUDA_OK(cudaSetDevice(h->device));
  if (ensure_arena(h)) return 1;
  const int L = h->L, d = h->dmax;
  const size_t smem = (4 + (size_t)h->qmax) * d * d * 8;
  if (smem > (size_t)h->max_smem) return fail("alternate marginals: bond capacity %d exceeds the shared-memory tiling", d);
  const MsgStore& m = h->msg[h->cur];
  double* d_out;
  CUDA_OK(cudaMalloc((void**)&d_out, sizeof(double) * std::max<int64_t>(h->psi_off[h->E2], 1)));
  const size_t per = 8 * (size_t)(L + 1) * d * d + 512;
  int64_t e0 = 0;
  while (e0 < h->E2) {
    h->arena.used = 0;
    const int64_t maxn = std::max<int64_t>(1, (int64_t)((h->arena.cap / 2) / (per + sizeof(PairJob))));
    const int64_t e1 = std::min(h->E2, e0 + maxn);
    std::vector<PairJob> jobs;
    for (int64_t e = e0; e < e1; ++e) {
      PairJob jb;
      memset(&jb, 0, sizeof jb);
      const int qs = h->q[h->src[e]], qd = h->q[h->dst[e]];
      jb.A = msg_ref(h, m, e, qs * qd);
      jb.B = msg_ref(h, m, h->rev[e], qs * qd);
      jb.psi = h->d_psi + h->psi_off[e];
      jb.qs = qs;
      jb.qd = qd;
      jb.out = d_out + h->psi_off[e];
      jb.Renv = (double*)h->arena.take(8 * (size_t)(L + 1) * d * d);
      if (!jb.Renv) { cudaFree(d_out); return fail("arena exhausted (alternate marginals)"); }
      jobs.push_back(jb);
    }
    PairJob* d_jobs;
    if (upload_jobs(h, jobs, &d_jobs)) { cudaFree(d_out); return 1; }
    k_alt_marginal<<<(unsigned)jobs.size(), NT, smem, h->st>>>(d_jobs, L, d);
    h->n_launch++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(h->st));
    e0 = e1;
  }
  CUDA_OK(cudaMemcpy(out, d_out, sizeof(double) * h->psi_off[h->E2], cudaMemcpyDeviceToHost));
  cudaFree(d_out);
  return 0;
}

// forward sample of the prior dynamics on the device (reference onesample!, src/sampling.jl:30-59).
// X: N x (T+1) int32, row i = trajectory of node i, states 0-based.  Deterministic in `seed` (counter-based RNG).
int mpbp_sample_prior(mpbp_handle h, uint64_t seed, int32_t* X) {
  if (!h || !X) return fail("null argument");
  if (h->inf_k > 0) return fail("sampling is defined on finite graphs (src/sampling.jl iterates the nodes of bp.g)");
  // (periodic handles: like the reference's onesample!, the forward simulation starts from phi^0 and ignores the wrap-around
  // factor w^T -- it only generates observations, test/periodic.jl:31)
  CUDA_OK(cudaSetDevice(h->device));
  std::vector<SampCls> sc(std::max<size_t>(h->classes.size(), 1));
  for (size_t ci = 0; ci < h->classes.size(); ++ci) {
    const NodeClass& c = h->classes[ci];
    SampCls& o = sc[ci];
    memset(&o, 0, sizeof o);
    if (c.z > SAMP_MAXZ) return fail("sampler supports degrees up to %d", SAMP_MAXZ);
    o.z = c.z;
    o.q = c.q;
    o.generic = c.generic ? 1 : 0;
    for (int k = 0; k < c.z; ++k) o.qn[k] = c.qn[k];
    o.wd = c.d_wd;
    o.wd_ts = (long long)c.wd_ts;
    if (c.generic) continue;
    for (int l = 0; l <= c.z; ++l) {
      if (c.ny[l] > SAMP_MAXNY) return fail("sampler supports nstates up to %d", SAMP_MAXNY);
      o.ny[l] = c.ny[l];
    }
    o.pxy = c.d_pxy;
    o.pxy_ts = (long long)c.pxy_ts;
    for (int k = 0; k < c.z; ++k) o.pxy_off[k] = (long long)c.pxy_off[k];
    o.pyy = c.d_pyy;
    for (int k = 1; k <= c.z; ++k) {
      const std::pair<int, int> key = k < c.z ? std::make_pair(k, 1) : std::make_pair(c.z, 0);
      auto it = c.pyy.find(key);
      if (it == c.pyy.end()) return fail("class %zu lacks the prob_yy table for (%d,%d)", ci, key.first, key.second);
      o.pyy_off[k] = (long long)it->second.first;
      o.pyy_ts[k] = (long long)it->second.second;
    }
    o.minit = c.d_minit;
    o.minit_ts = (long long)c.minit_ts;
  }
  SampCls* d_sc = nullptr;
  int* d_cls = nullptr;
  int64_t *d_colptr = nullptr, *d_dst = nullptr;
  int* d_X = nullptr;
  struct Guard {
    std::vector<void*> p;
    ~Guard() { for (void* q : p) cudaFree(q); }
  } guard;
  if (upload(&d_sc, sc.data(), sc.size())) return 1;
  guard.p.push_back(d_sc);
  if (upload(&d_cls, h->class_of_node.data(), (size_t)h->N)) return 1;
  guard.p.push_back(d_cls);
  if (upload(&d_colptr, h->colptr.data(), (size_t)h->N + 1)) return 1;
  guard.p.push_back(d_colptr);
  if (upload(&d_dst, h->dst.data(), (size_t)h->E2)) return 1;
  guard.p.push_back(d_dst);
  CUDA_OK(cudaMalloc((void**)&d_X, sizeof(int) * (size_t)h->N * h->L));
  guard.p.push_back(d_X);
  const unsigned nb = (unsigned)((h->N + 127) / 128);
  for (int t = -1; t < h->T; ++t) {
    k_sample_step<<<nb, 128, 0, h->st>>>(d_sc, d_cls, d_colptr, d_dst, h->d_phi, h->d_marg_off, h->d_q, (long long)h->N, h->L, t, seed,
                                         d_X, h->d_err);
    h->n_launch++;
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->st));
  if (check_err(h)) return 1;
  CUDA_OK(cudaMemcpy(X, d_X, sizeof(int) * (size_t)h->N * h->L, cudaMemcpyDeviceToHost));
  return 0;
}

int64_t mpbp_message_slot_bytes(mpbp_handle h) {
  if (!h) return 0;
  // data + bonds + ls, padded to 8 bytes
  const int64_t b = (int64_t)(sizeof(double) * h->slot + sizeof(double) * ((h->L + 1 + 1) / 2) + sizeof(double));
  return (b + 15) & ~int64_t(15);  // records stay 16-byte aligned in a packed buffer (128-bit copies)
}

}  // extern "C"
// halo payload: one fixed-size record per message = [slot doubles][L+1 bond ints, padded to 8 bytes][log-scale]
// PACK: message store -> records ; !PACK: records -> message store.  grid (n, segments), 128-bit copies.
template <bool PACK>
__global__ void __launch_bounds__(NT) k_pack_messages(double* mdata, int* mbonds, double* mls, const int64_t* edges, char* buf,
                                                      long long slot, int L, long long sb) {
  const long long e = edges[blockIdx.x];
  double2* m2 = reinterpret_cast<double2*>(mdata + e * slot);  // slot is a multiple of 2 doubles whenever dmax*q is even;
  double2* b2 = reinterpret_cast<double2*>(buf + (long long)blockIdx.x * sb);
  const long long n2 = slot / 2;
  if ((slot & 1) == 0 && ((reinterpret_cast<uintptr_t>(m2) | reinterpret_cast<uintptr_t>(b2)) & 15) == 0) {
    for (long long i = (long long)blockIdx.y * NT + threadIdx.x; i < n2; i += (long long)gridDim.y * NT) {
      if (PACK) b2[i] = m2[i]; else m2[i] = b2[i];
    }
  } else {
    double* m1 = mdata + e * slot;
    double* b1 = reinterpret_cast<double*>(buf + (long long)blockIdx.x * sb);
    for (long long i = (long long)blockIdx.y * NT + threadIdx.x; i < slot; i += (long long)gridDim.y * NT) {
      if (PACK) b1[i] = m1[i]; else m1[i] = b1[i];
    }
  }
  if (blockIdx.y == 0) {
    int* bb = reinterpret_cast<int*>(buf + (long long)blockIdx.x * sb + sizeof(double) * slot);
    double* bl = reinterpret_cast<double*>(buf + (long long)blockIdx.x * sb + sb - sizeof(double));
    for (int i = threadIdx.x; i <= L; i += NT) {
      if (PACK) bb[i] = mbonds[e * (L + 1) + i]; else mbonds[e * (L + 1) + i] = bb[i];
    }
    if (threadIdx.x == 0) {
      if (PACK) *bl = mls[e]; else mls[e] = *bl;
    }
  }
}
namespace {
int pack_unpack(mpbp_state* h, int64_t n, const int64_t* edges, void* dev_buf, bool pack) {
  if (!h || (!edges && n) || (!dev_buf && n)) return fail("null argument");
  if (n == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  for (int64_t k = 0; k < n; ++k)
    if (edges[k] < 0 || edges[k] >= h->E2) return fail("bad edge index");
  if ((int64_t)h->edge_idx_cap < n) {
    cudaFree(h->d_edge_idx);
    h->d_edge_idx = nullptr;
    CUDA_OK(cudaMalloc((void**)&h->d_edge_idx, sizeof(int64_t) * n));
    h->edge_idx_cap = (size_t)n;
  }
  CUDA_OK(cudaMemcpyAsync(h->d_edge_idx, edges, sizeof(int64_t) * n, cudaMemcpyHostToDevice, h->st));
  const int64_t sb = mpbp_message_slot_bytes(h);
  MsgStore& m = h->msg[h->cur];
  const int segs = (int)std::max<int64_t>(1, std::min<int64_t>(64, h->slot / (8 * NT)));
  dim3 grid((unsigned)n, segs);
  if (pack) k_pack_messages<true><<<grid, NT, 0, h->st>>>(m.data, m.bonds, m.ls, h->d_edge_idx, (char*)dev_buf, h->slot, h->L, sb);
  else k_pack_messages<false><<<grid, NT, 0, h->st>>>(m.data, m.bonds, m.ls, h->d_edge_idx, (char*)dev_buf, h->slot, h->L, sb);
  h->n_launch++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->st));
  return 0;
}
}  // namespace
extern "C" {

// One gather kernel on the engine's stream, synchronised before returning: the buffer is complete when the call returns.
int mpbp_pack_messages_dev(mpbp_handle h, int64_t n, const int64_t* edges, void* dev_buf) {
  return pack_unpack(h, n, edges, dev_buf, true);
}

// One scatter kernel on the engine's stream.  The CALLER orders the producer of dev_buf (e.g. the NCCL stream of an
// all_to_all) before this call: the engine's stream does not know about it.
int mpbp_unpack_messages_dev(mpbp_handle h, int64_t n, const int64_t* edges, const void* dev_buf) {
  return pack_unpack(h, n, edges, const_cast<void*>(dev_buf), false);
}

int mpbp_counters(mpbp_handle h, double* out8, int reset) {
  if (!h || !out8) return fail("null argument");
  CUDA_OK(cudaSetDevice(h->device));
  double fl5[6] = {0, 0, 0, 0, 0, 0};
  CUDA_OK(cudaMemcpy(fl5, h->d_flops, 6 * sizeof(double), cudaMemcpyDeviceToHost));
  const double fl = fl5[0];
  out8[0] = h->n_launch;
  out8[1] = fl;
  out8[2] = fl5[1];  // subspace-SVD calls
  out8[3] = h->qr_ms;
  out8[4] = h->n_ops;
  out8[5] = h->n_edge_updates;
  out8[6] = fl5[2];  // subspace-SVD iterations (sum)
  out8[7] = fl5[5];  // subspace-SVD calls that hit the iteration cap (resolved by the exact Jacobi fallback)
  if (reset) {
    h->n_launch = h->qr_ms = h->n_ops = h->n_edge_updates = 0;
    CUDA_OK(cudaMemset(h->d_flops, 0, 24 * sizeof(double)));
  }
  return 0;
}

int mpbp_family_flops(mpbp_handle h, double* out4) {
  if (!h || !out4) return fail("null argument");
  CUDA_OK(cudaSetDevice(h->device));
  double fl[24];
  CUDA_OK(cudaMemcpy(fl, h->d_flops, 24 * sizeof(double), cudaMemcpyDeviceToHost));
  out4[0] = fl[0];   // sweep-1 QR, algorithmic (unsplit 2mn^2 - 2/3 n^3)
  out4[1] = fl[16];  // executed on top of that by the TSQR splits (chunk triangles + merges)
  out4[2] = fl[17];  // Kronecker carry (k_kron_carry_mma), flops of the structured two-stage contraction
  out4[3] = fl[7];   // blocked subspace SVDs (k_jacobi_project): GEMMs + block orthonormalisations of the iterations run
  return 0;
}

int mpbp_set_stream(mpbp_handle h, void* cuda_stream) {
  if (!h) return fail("null handle");
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamSynchronize(h->st));
  if (h->own_stream) cudaStreamDestroy(h->st);
  h->st = (cudaStream_t)cuda_stream;
  h->own_stream = false;
  return 0;
}

// FP64 tensor-pipe (DMMA m8n8k4) throughput of this device, measured live: the roofline denominator of the
// QR/GEMM kernels (MEASURED_PEAKS.json carries no FP64 figure).  Returns TFLOP/s.
__global__ void k_peak_dmma(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = 0.0;
  const double a = threadIdx.x * 1e-9, b = 1.0000001;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[2 * j]), "+d"(c[2 * j + 1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int mpbp_measure_fp64_peak(int device, double* tflops) {
  CUDA_OK(cudaSetDevice(device));
  int sms = 0;
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int tpb = 256, blocks = sms * 8, iters = 20000;
  double* out;
  CUDA_OK(cudaMalloc((void**)&out, sizeof(double) * tpb * blocks));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    k_peak_dmma<<<blocks, tpb>>>(out, iters);
    cudaEventRecord(e1);
    CUDA_OK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = 2.0 * 8 * 8 * 4 * 8 * (double)iters * blocks * (tpb / 32) / best / 1e9;
  return 0;
}

// device ms per kernel family since the last reset (only while option "profile" = 1):
// [0] sweep-1 QR, [1] kron_carry, [2] kron_proj, [3] gemm_m2t, [4] qr_small, [5] jacobi_project, [6] finalize, [7] belief
int mpbp_kernel_times(mpbp_handle h, double* out, int n, int reset) {
  if (!h || !out) return fail("null argument");
  for (int i = 0; i < n; ++i) out[i] = i < 16 ? h->fam_ms[i] : 0.0;
  if (reset) for (int i = 0; i < 16; ++i) h->fam_ms[i] = 0.0;
  return 0;
}

int mpbp_set_option(mpbp_handle h, const char* name, double value) {
  if (!h || !name) return fail("null argument");
  std::string n(name);
  if (n == "arena_gb") {
    if (h->arena.base) return fail("arena already allocated");
    h->arena_gb = value;
  } else if (n == "max_group_ops") h->max_group_ops = value;
  else if (n == "profile") h->profile = (int)value;
  else if (n == "qr_fill") h->qr_fill = value;
  else if (n == "nstreams") h->nstreams = std::max(1.0, std::min(1.0 + mpbp_state::NAUX, value));
  else if (n == "group_mode") h->group_mode = value;
  else if (n == "bulk_split") h->bulk_split = std::max(1.0, value);
  else if (n == "bulk_split_min") h->bulk_split_min = value;
  else if (n == "level_balance") h->level_balance = value;
  else if (n == "outlier_split") h->outlier_split = value;
  else if (n == "kron_mma") h->kron_mma = value;
  else if (n == "hub_lane") h->hub_lane = value;
  else if (n == "tri_merge") h->tri_merge = value;
  else if (n == "svd_mode") h->svd_mode = value;
  else if (n == "hub_frac") h->hub_frac = value;
  else if (n == "lanes") h->lanes = value;
  else if (n == "periodic") {
    // periodic-in-time messages on any handle (periodic_mpbp_infinite_graph: mpbp_create_infinite + this option); set before
    // the first iteration
    if (value != 0 && h->dmax > 16) return fail("the periodic path supports dmax <= 16");
    h->periodic = value != 0;
  }
  else if (n == "twovar") {
    // two-time marginals of every belief computed from now on, for time distances up to `value` (0 = off)
    h->twovar = value > 0 ? (int)std::min<double>(value, h->L) : 0;
    if (h->twovar > 0 && !h->d_tv) {
      const size_t nb = sizeof(double) * (size_t)h->N * h->L * h->L * h->qmax * h->qmax;
      CUDA_OK(cudaSetDevice(h->device));
      CUDA_OK(cudaMalloc((void**)&h->d_tv, nb));
      CUDA_OK(cudaMemset(h->d_tv, 0, nb));
    }
  }
  else return fail("unknown option %s", name);
  return 0;
}

// ---- test hooks (unit tests of the two numerical kernels through the C ABI) ----
}  // extern "C"
__global__ void __launch_bounds__(NT) k_test_qr(double* A, int m, int n, double* R, int vrows) {
  extern __shared__ double smem[];
  qr_r_cta(A + (size_t)blockIdx.x * m * n, m, n, n, R + (size_t)blockIdx.x * min(m, n) * n, n, false, smem, vrows);
}
extern "C" {
int mpbp_test_qr(const double* A, int batch, int m, int n, double* R) {
  if (m > QR_MAX_M || n > NT * QR_MAX_CPT + QB) return fail("test_qr: size out of range");
  double *dA, *dR;
  const int k = std::min(m, n);
  CUDA_OK(cudaMalloc((void**)&dA, sizeof(double) * batch * m * n));
  CUDA_OK(cudaMalloc((void**)&dR, sizeof(double) * batch * k * n));
  CUDA_OK(cudaMemcpy(dA, A, sizeof(double) * batch * m * n, cudaMemcpyHostToDevice));
  const int vrows = std::min(2048, std::max(64, m));
  const size_t sm = qr_shared_doubles(vrows) * 8;
  CUDA_OK(cudaFuncSetAttribute(k_test_qr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  k_test_qr<<<batch, NT, sm>>>(dA, m, n, dR, vrows);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpy(R, dR, sizeof(double) * batch * k * n, cudaMemcpyDeviceToHost));
  cudaFree(dA);
  cudaFree(dR);
  return 0;
}
}  // extern "C"
template <int H>
__global__ void __launch_bounds__(NT, (H == 32 ? 2 : 1)) k_test_qr_ft(const double* A, int m, int n, double* R) {
  extern __shared__ double smem[];
  qr_ft_cta<H>(A + (size_t)blockIdx.x * m * n, m, n, n, R + (size_t)blockIdx.x * n * n, n, false, smem);
}
extern "C" {
// flat-tree DMMA QR: R is n x n per matrix (rows >= min(m,n) are noise); ms = device time of the launch
int mpbp_test_qr_ft(const double* A, int batch, int m, int n, int H, double* R, double* ms) {
  double *dA, *dR;
  CUDA_OK(cudaMalloc((void**)&dA, sizeof(double) * (size_t)batch * m * n));
  CUDA_OK(cudaMalloc((void**)&dR, sizeof(double) * (size_t)batch * n * n));
  CUDA_OK(cudaMemcpy(dA, A, sizeof(double) * (size_t)batch * m * n, cudaMemcpyHostToDevice));
  const size_t sm = (H == 64 ? ft_smem_doubles<64>(n) : H == 32 ? ft_smem_doubles<32>(n) : ft_smem_doubles<16>(n)) * 8;
  int maxs = 0;
  CUDA_OK(cudaDeviceGetAttribute(&maxs, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0));
  if (sm > (size_t)maxs) return fail("test_qr_ft: n too large for H=%d", H);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (H == 64) {
      CUDA_OK(cudaFuncSetAttribute(k_test_qr_ft<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_test_qr_ft<64><<<batch, NT, sm>>>(dA, m, n, dR);
    } else if (H == 32) {
      CUDA_OK(cudaFuncSetAttribute(k_test_qr_ft<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_test_qr_ft<32><<<batch, NT, sm>>>(dA, m, n, dR);
    } else {
      CUDA_OK(cudaFuncSetAttribute(k_test_qr_ft<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_test_qr_ft<16><<<batch, NT, sm>>>(dA, m, n, dR);
    }
    cudaEventRecord(e1);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventSynchronize(e1));
    float t;
    cudaEventElapsedTime(&t, e0, e1);
    if (t < best) best = t;
  }
  if (ms) *ms = best;
  CUDA_OK(cudaMemcpy(R, dR, sizeof(double) * (size_t)batch * n * n, cudaMemcpyDeviceToHost));
  cudaFree(dA);
  cudaFree(dR);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}
__global__ void __launch_bounds__(NT) k_test_svd(const double* M, int p, int n, Trunc tr, int jac_doubles, double* scratch,
                                                 size_t scratch_per, double* U, double* S, int* err, double* stats, int mode) {
  extern __shared__ double smem_tsvd[];
  double* smem = smem_tsvd;
  __shared__ int flag, s_done;
  __shared__ double red[NW + 1];
  const double* Mb = M + (size_t)blockIdx.x * p * n;
  double* sc = scratch + (size_t)blockIdx.x * scratch_per;
  double* R2 = sc;                                  // p*p (direct path, n > p)
  double* Mcopy = R2 + (size_t)p * p;               // n*p row-major copy for the QR pre-reduction (destroyed)
  double* svdscr = Mcopy + (size_t)p * n;            // svd_scratch_doubles(p, n)
  if (n > p && svd_direct(p, n, jac_doubles)) {
    for (int i = threadIdx.x; i < p * n; i += NT) Mcopy[i] = Mb[i];
    __syncthreads();
    qr_ft_cta<16>(Mcopy, n, p, p, R2, p, false, smem);  // M^T (n x p row-major == M col-major) -> R (p x p)
  }
  const SvdLeft sv = svd_left_cta(Mb, R2, svdscr, p, n, tr, tr.d, jac_doubles, smem, &flag, &s_done, red, err, stats, mode);
  const double* sig = smem;
  const int* order = reinterpret_cast<const int*>(smem + SUB_BMAX);
  const int keep = min(tr.d, sv.ceff);
  for (int idx = threadIdx.x; idx < p * tr.d; idx += NT) {
    const int a = idx % p, kk = idx / p;
    U[(size_t)blockIdx.x * p * tr.d + idx] = kk < keep ? sv.A[a + (size_t)order[kk] * p] : 0.0;
  }
  for (int kk = threadIdx.x; kk < tr.d; kk += NT) S[(size_t)blockIdx.x * tr.d + kk] = kk < keep ? sig[order[kk]] : 0.0;
}
// leading-d left singular vectors of `batch` column-major p x n matrices through the op-truncation SVD core
// (direct Jacobi or blocked subspace iteration, chosen as in the engine).  U: [batch][p x d], S: [batch][d].
// stats[0..4] = subspace calls, iterations, block sizes (sum), max sweeps (sum), exact fallbacks.  *ms = best of 3.
int mpbp_test_svd(const double* M, int batch, int p, int n, int d, double* U, double* S, double* stats5, double* ms) {
  int maxs = 0;
  CUDA_OK(cudaDeviceGetAttribute(&maxs, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0));
  maxs -= 2048;
  const size_t jac_fixed = 3 * SUB_BMAX, jac_doubles = (size_t)maxs / 8 - jac_fixed;
  CUDA_OK(cudaFuncSetAttribute(k_test_svd, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));
  const size_t per = (size_t)p * p + (size_t)p * n + svd_scratch_doubles(p, n) + 64;
  double *dM, *dU, *dS, *dscr, *dst;
  int* derr;
  CUDA_OK(cudaMalloc((void**)&dM, sizeof(double) * (size_t)batch * p * n));
  CUDA_OK(cudaMalloc((void**)&dU, sizeof(double) * (size_t)batch * p * d));
  CUDA_OK(cudaMalloc((void**)&dS, sizeof(double) * (size_t)batch * d));
  CUDA_OK(cudaMalloc((void**)&dscr, sizeof(double) * per * batch));
  CUDA_OK(cudaMalloc((void**)&dst, sizeof(double) * 16));
  CUDA_OK(cudaMalloc((void**)&derr, sizeof(int)));
  CUDA_OK(cudaMemset(derr, 0, sizeof(int)));
  CUDA_OK(cudaMemcpy(dM, M, sizeof(double) * (size_t)batch * p * n, cudaMemcpyHostToDevice));
  Trunc tr{0, d, 0.0};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CUDA_OK(cudaMemset(dst, 0, sizeof(double) * 16));
    cudaEventRecord(e0);
    k_test_svd<<<batch, NT, (size_t)maxs>>>(dM, p, n, tr, (int)jac_doubles, dscr, per, dU, dS, derr, dst,
                                            getenv("MPBP_SVD_MODE") ? atoi(getenv("MPBP_SVD_MODE")) : 2);
    cudaEventRecord(e1);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventSynchronize(e1));
    float t;
    cudaEventElapsedTime(&t, e0, e1);
    best = std::min(best, t);
  }
  if (ms) *ms = best;
  int herr = 0;
  CUDA_OK(cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(U, dU, sizeof(double) * (size_t)batch * p * d, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(S, dS, sizeof(double) * (size_t)batch * d, cudaMemcpyDeviceToHost));
  if (stats5) CUDA_OK(cudaMemcpy(stats5, dst, sizeof(double) * 5, cudaMemcpyDeviceToHost));
  {
    double ph[6];
    CUDA_OK(cudaMemcpy(ph, dst + 8, sizeof(double) * 6, cudaMemcpyDeviceToHost));
    double fb = 0;
    CUDA_OK(cudaMemcpy(&fb, dst + 5, sizeof(double), cudaMemcpyDeviceToHost));
    if (getenv("MPBP_SVD_PHASES")) printf("householder orthonormalisations per call (mode 2: Cholesky-QR fallbacks) %.2f\n", fb / batch);
    if (getenv("MPBP_SVD_PHASES")) printf("svd phases (Mcycles/call): final %.2f  select %.2f  start-orth %.2f  iter-gemms %.2f  iter-orth %.2f  iter-ritz %.2f\n", ph[0] / batch / 1e6, ph[1] / batch / 1e6, ph[2] / batch / 1e6, ph[3] / batch / 1e6, ph[4] / batch / 1e6, ph[5] / batch / 1e6);
  }
  cudaFree(dM); cudaFree(dU); cudaFree(dS); cudaFree(dscr); cudaFree(dst); cudaFree(derr);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (herr) return fail("test_svd: device error flags 0x%x", herr);
  return 0;
}
__global__ void __launch_bounds__(NT) k_test_jacobi(double* A, int p, int c, double* sig, int* order) {
  __shared__ int flag;
  extern __shared__ double smem_tjac[];
  double* smem = smem_tjac;
  double* a = A + (size_t)blockIdx.x * p * c;
  jacobi_cols(a, p, c, p, &flag);
  double* s = smem;
  int* o = reinterpret_cast<int*>(smem + c);
  jacobi_sort(a, p, c, p, s, o);
  for (int i = threadIdx.x; i < c; i += NT) {
    sig[(size_t)blockIdx.x * c + i] = s[o[i]];
    order[(size_t)blockIdx.x * c + i] = o[i];
  }
}
int mpbp_test_jacobi(double* A, int batch, int p, int c, double* sig, int32_t* order) {
  double *dA, *dS;
  int* dO;
  CUDA_OK(cudaMalloc((void**)&dA, sizeof(double) * batch * p * c));
  CUDA_OK(cudaMalloc((void**)&dS, sizeof(double) * batch * c));
  CUDA_OK(cudaMalloc((void**)&dO, sizeof(int) * batch * c));
  CUDA_OK(cudaMemcpy(dA, A, sizeof(double) * batch * p * c, cudaMemcpyHostToDevice));
  k_test_jacobi<<<batch, NT, (c + (c + 1) / 2 + 1) * 8>>>(dA, p, c, dS, dO);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpy(A, dA, sizeof(double) * batch * p * c, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(sig, dS, sizeof(double) * batch * c, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(order, dO, sizeof(int) * batch * c, cudaMemcpyDeviceToHost));
  cudaFree(dA);
  cudaFree(dS);
  cudaFree(dO);
  return 0;
}

}  // extern "C"
