/*
 * mpbp.h -- C-ABI of the B200-native MPBP message-update engine (libmpbp_b200.so).
 *
 * The reference (stecrotti/MatrixProductBP.jl) has no FFI layer: the seam is Julia multiple dispatch on
 * the message-store type parameter M2 of MPBP{G,F,V,M2,M1} (src/mpbp.jl:1).  A CUDA-backed message store
 * dispatches the calls below through `ccall` (see INTEGRATION.md and
 * matrixproductbp.jl_b200/julia/MatrixProductBPCUDA.jl).  Every entry point cites the reference code it
 * replaces.  All arrays are caller-owned host buffers unless the name says `dev`; multi-dimensional
 * arrays are COLUMN-MAJOR with the index order of the Julia arrays they mirror.  All functions return 0
 * on success, non-zero on error; mpbp_last_error() returns the message.  A handle is NOT re-entrant.
 *
 * Indices are 0-based at this boundary (the Julia glue subtracts 1).
 */
#ifndef MPBP_B200_H
#define MPBP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mpbp_state* mpbp_handle;

/* truncation policies: TensorTrains.TruncBond / TruncThresh / TruncBondThresh (TruncBondMax == TruncBond),
 * re-exported at src/MatrixProductBP.jl:42,69 and threaded through iterate!(...; svd_trunc) */
enum { MPBP_TRUNC_BOND = 0, MPBP_TRUNC_THRESH = 1, MPBP_TRUNC_BOND_THRESH = 2 };

/* update schedules.  SEQUENTIAL = the reference's in-place sweep in `order` (src/mpbp.jl:189-192 with one
 * thread), executed level by level (nodes of one level are pairwise non-adjacent, so the result is
 * identical to the serial sweep).  PARALLEL = Jacobi: every node reads the previous iteration's messages. */
enum { MPBP_SCHEDULE_SEQUENTIAL = 0, MPBP_SCHEDULE_PARALLEL = 1 };

const char* mpbp_last_error(void);
int mpbp_version(void);

/* ---- construction: replaces mpbp(g, w, q, T; ...) src/mpbp.jl:60-70 and the MPBP struct src/mpbp.jl:1-33 ----
 * Graph in the reference's IndexedBiDiGraph order: directed edge e = position in the CSC of the adjacency
 * matrix (source = column), i.e. sorted by (src,dst); out-edges of node i are colptr[i] .. colptr[i+1]-1,
 * dst[e] their destinations, rev[e] the index of the reverse edge (g.X of src/mpbp.jl:40-58).
 * dmax = capacity of every bond dimension held on the device (<= 30).  device = CUDA ordinal.
 * Messages start as flat_mpem2(q_i,q_j,T; d=1) (src/mpems.jl:20), beliefs uniform, f = 0. */
int mpbp_create(int64_t N, int64_t E2, int T, const int32_t* q, const int64_t* colptr, const int64_t* dst,
                const int64_t* rev, int dmax, int device, mpbp_handle* out);

/* InfiniteRegularGraph(k) / mpbp_infinite_graph, src/infinite_graph.jl:8-35: one node, one stored message
 * that plays the role of all k incoming ones. */
int mpbp_create_infinite(int k, int T, int q, int dmax, int device, mpbp_handle* out);
/* InfiniteBipartiteRegularGraph((kA, kB)) (src/infinite_graph.jl:62-122): node 0 = class A (degree kA, qA states), node 1 =
 * class B (degree kB, qB states); edge 0 = message A -> B, edge 1 = message B -> A (reverse of each other).  Everything
 * else (factor classes, phi, psi per edge, iterate, read-outs) works as on a 2-node graph; the Bethe free energy of the
 * infinite graph is (f[0]*kB + f[1]*kA)/(kA+kB) (src/infinite_graph.jl:120-122, host side). */
int mpbp_create_infinite_bipartite(int kA, int kB, int T, int qA, int qB, int dmax, int device, mpbp_handle* out);

/* periodic_mpbp(g, w, q, T) (src/mpbp.jl:399-409): same arguments as mpbp_create; the messages are PeriodicMPEM2s
 * (src/mpems.jl:96-155: the matrix product is closed by a trace, the factor of the last time maps (x^T_neigh, x^T_i) to x^0_i),
 * starting from flat_periodic_mpem2 with d = 1.  mpbp_iterate runs onebpiter! on ring tensor trains (recursive factors only,
 * degree <= 10, dmax <= 16; csrc/periodic.cuh); mpbp_beliefs / mpbp_free_energy / mpbp_pair_beliefs / get / set_message work as on
 * an open handle (bonds[0] == bonds[T+1] is the closing bond); damping is the ring sum + compress! + normalize! of set_msg!.
 * Two-time marginals (option "twovar") and the forward sampler (which, like the reference's onesample!, ignores the
 * wrap-around factor) work as on an open handle; alternate marginals are not defined on this path (loud error).  periodic_mpbp_infinite_graph = mpbp_create_infinite followed by
 * mpbp_set_option(h, "periodic", 1) before the first iteration. */
int mpbp_create_periodic(int64_t N, int64_t E2, int T, const int32_t* q, const int64_t* colptr, const int64_t* dst,
                         const int64_t* rev, int dmax, int device, mpbp_handle* out);

int mpbp_destroy(mpbp_handle h);

/* ---- factors: the host-tabulated values of a RecursiveBPFactor (src/recursive_bp_factor.jl:6-61) ----
 * One "node class" = all tables of one node type (degree z, q states, neighbour state counts qn[z]).
 * nt = 1 for time-independent factors, T+1 otherwise (slice t is used for site t; the last slice of `w`/`wd`
 * is never read, exactly like _f_bp_partial src/recursive_bp_factor.jl:76-84).
 *   ny[l]    = nstates(w, l), l = 0..z                                  (:11)
 *   pxy      = for t, for k<z : [ny[1] x qn[k] x q]   prob_xy(w,y,xk,xi,k)               (:110-111)
 *   pyy      = for p<npairs, for t : [ny[d1+d2] x ny[d1] x ny[d2] x q]  prob_yy(w,y,y1,y2,xi,d1,d2) (:120-121)
 *   w        = for t, for j<z : [q x q x qn[j] x ny[z-1]]  prob_y_partial(w,x',x,xj,y,z-1,j)   (:49-54,79-80)
 *   wd       = for t : [q x q x ny[z]]                     prob_y(w,x',x,y,z)  (dummy neighbour, :59-61)
 *   minit    = for t : [ny[0] x q]                         prob_y0(w,y,xi)     (:27,133-137)
 * The (d1,d2) pairs must cover the cavity recursion of degree z: (i,1) i=1..z-1, (z,0),
 * (1,l) l=0..z-2, (i,z-1-i) i=1..z-1. */
int mpbp_add_node_class(mpbp_handle h, int z, int q, const int32_t* qn, int nt, const int32_t* ny,
                        const double* pxy, int npairs, const int32_t* pair_d1, const int32_t* pair_d2,
                        const double* pyy, const double* w, const double* wd, const double* minit,
                        int32_t* class_id);
/* drop every node class uploaded so far (device tables freed, all nodes unassigned): called by the host layer before it
 * re-tabulates the factors (bp.w edited in place), so that repeated syncs do not accumulate tables */
int mpbp_clear_node_classes(mpbp_handle h);
int mpbp_set_node_classes(mpbp_handle h, const int32_t* class_of_node /* N */);

/* Generic BPFactor (src/bp_core.jl:1-57): dense table per node, for t<T+1 (nt = 1 or T+1):
 *   wtab = for t : [q x qn[0] x ... x qn[z-1] x q]   w(x', x_neighbours, x).   Exhaustive-trace path: degree 1..8, the
 * product of the other neighbours' bond dimensions must fit dmax (it is d^(z-1), src/bp_core.jl:34). */
int mpbp_add_generic_class(mpbp_handle h, int z, int q, const int32_t* qn, int nt, const double* wtab,
                           int32_t* class_id);

/* ---- reweightings: bp.phi / bp.psi of src/mpbp.jl:4-5 ----
 * phi: for i<N, for t<=T : [q_i]        psi: for e<E2, for t<=T : [q_src x q_dst] */
int mpbp_set_phi(mpbp_handle h, const double* phi);
int mpbp_set_psi(mpbp_handle h, const double* psi);

/* ---- messages: bp.mu[e] as MPEM2 (src/mpems.jl:15-16): site t is [bond[t] x bond[t+1] x q_src x q_dst] ----
 * get: bonds[T+2]; data = concatenation over t of the site tensors, normalisation folded in (z = 1).
 * data_capacity in doubles (query with data == NULL -> *needed).  Used for checkpoint/resume and parity. */
int mpbp_get_message(mpbp_handle h, int64_t e, int32_t* bonds, double* data, int64_t data_capacity, int64_t* needed);
int mpbp_set_message(mpbp_handle h, int64_t e, const int32_t* bonds, const double* data);
int mpbp_reset_messages(mpbp_handle h); /* reset_messages!, src/mpbp.jl:72-80 */

/* ---- the hot path: iterate!(bp; maxiter, svd_trunc, tol, damp, nodes, shuffle_nodes) src/mpbp.jl:185-198 ----
 * nodes/n_nodes: the `nodes` keyword (NULL = all).  order: optional [maxiter x n_nodes] visiting orders
 * (row it = permutation used at iteration it; NULL = `nodes` as given every iteration, i.e.
 * shuffle_nodes=false).  obs: optional [N x qmax] observable f(x,i) for the convergence callback CB_BP
 * (src/mpbp.jl:157-183; NULL = (x,i)->x with x numbered from 1).  deltas[maxiter] receives CB_BP.Δs.
 * Returns the number of iterations run in *iters (stops when Δ < tol). */
int mpbp_iterate(mpbp_handle h, int maxiter, int trunc_kind, int trunc_d, double trunc_eps, double tol,
                 double damp, int schedule, const int64_t* nodes, int64_t n_nodes, const int64_t* order,
                 const double* obs, int* iters, double* deltas);

/* ---- read-outs (north-star item 3) ----
 * beliefs(bp) src/mpbp.jl:237 :            out = for i, for t : [q_i]
 * pair_beliefs(bp) src/mpbp.jl:202-235 :   out = for e, for t : [q_src x q_dst];  logz[N]
 * bethe_free_energy contributions bp.f, src/mpbp.jl:298 / recursive_bp_factor.jl:163 : f[N] */
int mpbp_beliefs(mpbp_handle h, double* out);
int mpbp_pair_beliefs(mpbp_handle h, double* out, double* logz);
int mpbp_free_energy(mpbp_handle h, double* f);

/* two-time marginals b_i(x^t, x^u) of every node's belief (reference: beliefs_tu / twovar_marginals.(bp.b),
 * src/mpbp.jl:239; feeds autocorrelations / autocovariances :245-255,289-296).  Computed together with the beliefs
 * once mpbp_set_option(h, "twovar", maxdist) is set (maxdist >= 1; T = all pairs).
 * out[((i*L + t)*L + u)*Q + x_t + q_i*x_u], L = T+1, Q = qmax*qmax; entries with t >= u or u - t > maxdist are 0. */
int mpbp_twovar_marginals(mpbp_handle h, double* out);

/* alternate marginals p(x_i^t, x_j^{t+1}) of every directed edge i->j from the current messages (reference:
 * alternate_marginals, src/mpbp.jl:270-280).  Same layout and size as the pair beliefs of mpbp_pair_beliefs:
 * per edge e (T+1) blocks of q_src*q_dst doubles, out[...][t][x_i^t + q_src*x_j^{t+1}] for t < T; the block t = T is 0. */
int mpbp_alternate_marginals(mpbp_handle h, double* out);

/* forward sample of the prior dynamics on the device (reference: onesample!, src/sampling.jl:30-59, the input generator of
 * draw_node_observations!, :191-210): x_i^0 ~ phi_i^0/sum, x_i^{t+1} ~ w_i^t(. | x_neighbours^t, x_i^t) evaluated from the
 * uploaded factor tables (recursive classes fold the neighbours like the factor's functor, src/recursive_bp_factor.jl:34-46;
 * generic classes index their dense table).  X: N x (T+1) int32, states 0-based, row i = node i.  Deterministic in `seed`
 * (counter-based RNG documented in csrc/kernels.cuh: samp_uniform), finite graphs only. */
int mpbp_sample_prior(mpbp_handle h, uint64_t seed, int32_t* X);

/* ---- multi-GPU plumbing (no reference counterpart; see DESIGN.md "multi-GPU") ----
 * pack/unpack the fixed-capacity device slots of `n` messages into/from one contiguous DEVICE buffer so that
 * the host layer can exchange cut-edge messages with one collective.  Record size from mpbp_message_slot_bytes
 * (a multiple of 16).  Each call is ONE gather/scatter kernel on the engine's stream, synchronised before it returns.
 * unpack: the caller must have ordered the producer of dev_buf (e.g. the collective's stream) before the call. */
int64_t mpbp_message_slot_bytes(mpbp_handle h);
int mpbp_pack_messages_dev(mpbp_handle h, int64_t n, const int64_t* edges, void* dev_buf);
int mpbp_unpack_messages_dev(mpbp_handle h, int64_t n, const int64_t* edges, const void* dev_buf);

/* ---- introspection for bench / roofline accounting ----
 * counters accumulated since the last reset: [0] kernel launches, [1] ALGORITHMIC FLOPs of the sweep-1 Q-less QRs
 * (2mn^2 - 2/3 n^3 of the unsplit matrices; TSQR chunk/merge overhead is not counted), [2] subspace-SVD calls,
 * [3] device ms in the sweep-1 QR kernels (CUDA events, only when profiling is on), [4] heavy ops run,
 * [5] edge updates, [6] subspace-SVD iterations, [7] subspace-SVD calls resolved by the exact Jacobi fallback. */
int mpbp_counters(mpbp_handle h, double* out8, int reset);
/* device-counted FLOPs per kernel family since the last mpbp_counters(reset = 1): [0] sweep-1 QR, algorithmic (= counters[1]);
 * [1] executed on top of [0] by TSQR splits (chunk triangles + merges; not credited to the roofline); [2] Kronecker carry
 * (structured two-stage contraction, counted from the runtime dims and the non-zero prob_yy pairs); [3] blocked subspace SVDs
 * (the two tall GEMMs and the block orthonormalisations of every iteration run; direct small Jacobi SVDs are not counted).
 * Together with mpbp_kernel_times they give one roofline fraction per kernel family (bench.py: roofline.families). */
int mpbp_family_flops(mpbp_handle h, double* out4);
/* engine tuning knobs (none changes a result bit): "arena_gb" scratch arena size, "max_group_ops" ops per launch group,
 * "nstreams" (1..4) concurrent streams per cavity round, "qr_fill" CTAs below which tall QRs are TSQR-split,
 * "level_balance" (default 1) stagger the cavity levels of independent nodes so that every round carries similar
 * work, "profile" (0/1) per-kernel-family CUDA-event timing, "twovar" (maxdist, 0 = off) also compute the two-time
 * marginals read by mpbp_twovar_marginals, "periodic" (0/1, before the first iteration) periodic-in-time messages on this
 * handle (see mpbp_create_periodic). */
int mpbp_set_option(mpbp_handle h, const char* name, double value);
/* device ms per kernel family since the last reset (option "profile" = 1): [0] sweep-1 QR, [1] kron_carry,
 * [2] kron_proj, [3] gemm_m2t, [4] qr_small, [5] jacobi_project, [6] finalize, [7] belief */
int mpbp_kernel_times(mpbp_handle h, double* out, int n, int reset);
/* run all engine work on a caller-owned CUDA stream (cudaStream_t), e.g. torch's current stream */
int mpbp_set_stream(mpbp_handle h, void* cuda_stream);
/* FP64 tensor-pipe (DMMA) peak of `device` in TFLOP/s, measured live (roofline denominator) */
int mpbp_measure_fp64_peak(int device, double* tflops);

/* ---- test hooks: the two numerical building blocks, callable on raw host matrices ----
 * mpbp_test_qr: `batch` row-major m x n matrices -> R factors (min(m,n) x n, row-major) of the Q-less QR.
 * mpbp_test_jacobi: `batch` column-major p x c matrices, orthogonalised in place by one-sided Jacobi;
 *   sig = column norms sorted descending, order = the matching column indices. */
int mpbp_test_qr(const double* A, int batch, int m, int n, double* R);
/* flat-tree DMMA QR (the sweep-1 kernel): R is n x n per matrix; H = 32 or 16; *ms = best device time of 3 launches */
int mpbp_test_qr_ft(const double* A, int batch, int m, int n, int H, double* R, double* ms);
/* leading-d left singular vectors of `batch` column-major p x n matrices through the op-truncation SVD core (direct
 * Jacobi or blocked subspace iteration, chosen as in the engine).  U: [batch][p x d], S: [batch][d]; stats5 =
 * {subspace calls, iterations, sum of block sizes, sum of max sweeps, calls that hit the iteration cap}. */
int mpbp_test_svd(const double* M, int batch, int p, int n, int d, double* U, double* S, double* stats5, double* ms);
int mpbp_test_jacobi(double* A, int batch, int p, int c, double* sig, int32_t* order);

#ifdef __cplusplus
}
#endif
#endif
