# MatrixProductBPCUDA.jl -- Julia glue that selects the CUDA-backed MPBP state in place of the CPU one.
#
# It keeps the reference API intact: `mpbp(::Glauber/SIS/SIRS)`, the `BPFactor` / `RecursiveBPFactor` definitions and the
# `SVDTrunc` objects are the reference's own; only the message store (the `M2` type parameter of
# `MPBP{G,F,V,M2,M1}`, /root/reference/src/mpbp.jl:1) is replaced by a device handle, and `iterate!`, `beliefs`,
# `pair_beliefs`, `bethe_free_energy`, `reset_messages!` dispatch to `ccall`s into libmpbp_b200.so (include/mpbp.h).
# No CUDA.jl, no kernel DSL, no CPU fallback.  (There is no `julia` binary in the build image: this file is reviewed,
# not executed, there; every entry point it binds is exercised through the same C-ABI by the Python tests.)
module MatrixProductBPCUDA

using MatrixProductBP
using MatrixProductBP: MPBP, BPFactor, RecursiveBPFactor, nstates, prob_y, prob_xy, prob_yy, prob_y0,
    prob_y_partial, getT, CB_BP, expectation, covariance
using IndexedGraphs, TensorTrains, SparseArrays
import MatrixProductBP: iterate!, beliefs, pair_beliefs, bethe_free_energy, reset_messages!, means, beliefs_tu,
    autocorrelations, autocovariances, alternate_marginals

const LIB = get(ENV, "MPBP_B200_LIB", joinpath(@__DIR__, "..", "libmpbp_b200.so"))

check(status) = status == 0 || error(unsafe_string(ccall((:mpbp_last_error, LIB), Cstring, ())))

"Device message store: what replaces `Vector{MPEM2}` as the `M2` parameter."
mutable struct CuMPBP{G,F,V}
    g::G; w::Vector{V}; ϕ::Vector{Vector{Vector{F}}}; ψ::Vector{Vector{Matrix{F}}}
    q::Vector{Int32}; T::Int; dmax::Int
    h::Ptr{Cvoid}
    classes_dirty::Bool
end

function CuMPBP(bp::MPBP{G,F}; dmax::Int=16, device::Int=0) where {G<:IndexedBiDiGraph,F}
    g = bp.g; N = nv(g); T = getT(bp)
    q = Int32[nstates(bp, i) for i in 1:N]
    colptr = Int64.(g.A.colptr .- 1)                  # out-edges of node i = CSC column i (src = column)
    dst = Int64.(rowvals(g.A) .- 1)
    rev = Int64.(nonzeros(g.X) .- 1)                   # index of the reverse edge, src/mpbp.jl:40-58
    h = Ref{Ptr{Cvoid}}(C_NULL)
    # periodic_mpbp states (src/mpbp.jl:399-409) carry PeriodicMPEM2 messages: the ring engine of csrc/periodic.cuh
    periodic = eltype(bp.μ) <: PeriodicMPEM2
    if periodic
        check(ccall((:mpbp_create_periodic, LIB), Cint,
            (Int64, Int64, Cint, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Cint, Cint, Ref{Ptr{Cvoid}}),
            N, ne(g), T, q, colptr, dst, rev, min(dmax, 16), device, h))
    else
    check(ccall((:mpbp_create, LIB), Cint,
        (Int64, Int64, Cint, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Cint, Cint, Ref{Ptr{Cvoid}}),
        N, ne(g), T, q, colptr, dst, rev, dmax, device, h))
    end
    cu = CuMPBP{G,F,eltype(bp.w)}(g, bp.w, bp.ϕ, bp.ψ, q, T, dmax, h[], true)
    finalizer(x -> ccall((:mpbp_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), cu)
    sync_reweightings!(cu)
    return cu
end

"""
`CuMPBP(bp)` for `InfiniteRegularGraph` / `InfiniteBipartiteRegularGraph` states (src/infinite_graph.jl:8-122): one node per
class.  Bipartite: the reference's slot `e` ("message into node e") is the engine's edge `3 - e` (edge 1 = A→B, edge 2 = B→A),
so ψ is uploaded in that order; `bethe_free_energy` re-weights the two blocks on the host.
"""
function CuMPBP(bp::MPBP{G,F}; dmax::Int=16, device::Int=0) where {G<:Union{InfiniteRegularGraph,InfiniteBipartiteRegularGraph},F}
    g = bp.g; T = getT(bp); q = Int32[nstates(bp, i) for i in 1:nv(g)]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    if g isa InfiniteRegularGraph
        check(ccall((:mpbp_create_infinite, LIB), Cint, (Cint, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}), g.k, T, q[1], dmax, device, h))
        # periodic_mpbp_infinite_graph (test/periodic.jl:78-93): the same handle with ring messages
        eltype(bp.μ) <: PeriodicMPEM2 && check(ccall((:mpbp_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Cdouble), h[], "periodic", 1.0))
        ψ = bp.ψ
    else
        check(ccall((:mpbp_create_infinite_bipartite, LIB), Cint, (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
            g.k[1], g.k[2], T, q[1], q[2], dmax, device, h))
        ψ = bp.ψ[[2, 1]]                                # engine edge order
    end
    cu = CuMPBP{G,F,eltype(bp.w)}(g, bp.w, bp.ϕ, ψ, q, T, dmax, h[], true)
    finalizer(x -> ccall((:mpbp_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), cu)
    sync_reweightings!(cu)
    return cu
end

"`sample_prior(cu; seed)` -> X (N × (T+1), states from 1): forward simulation of the prior on the device (src/sampling.jl:30-59)"
function sample_prior(cu::CuMPBP; seed::Integer=rand(UInt64))
    cu.classes_dirty && sync_factors!(cu)
    X = zeros(Int32, cu.T + 1, nv(cu.g))               # C layout [i][t] = Julia (t, i)
    check(ccall((:mpbp_sample_prior, LIB), Cint, (Ptr{Cvoid}, UInt64, Ptr{Int32}), cu.h, UInt64(seed), X))
    permutedims(X) .+ 1
end

function sync_reweightings!(cu::CuMPBP)
    ϕ = reduce(vcat, (reduce(vcat, ϕᵢ) for ϕᵢ in cu.ϕ))            # [i][t][x]
    ψ = reduce(vcat, (reduce(vcat, vec.(ψₑ)) for ψₑ in cu.ψ))      # [e][t][x_src, x_dst] column-major
    GC.@preserve ϕ ψ begin
        check(ccall((:mpbp_set_phi, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), cu.h, ϕ))
        check(ccall((:mpbp_set_psi, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), cu.h, ψ))
    end
end

"(d1,d2) operand sizes met by CavityTools.cavity for a degree-z node"
function cavity_pairs(z)
    z == 0 && return [(0, 0)]
    z == 1 && return [(1, 0)]
    unique(vcat([(k, 1) for k in 1:z-1], [(z, 0)], [(1, z - 1 - k) for k in z-1:-1:1], [(k, z - 1 - k) for k in 1:z-1]))
end

"Tabulate a RecursiveBPFactor (host side, A10 of SURVEY.md) and upload one class per distinct (factor, degree)."
function sync_factors!(cu::CuMPBP)
    N = nv(cu.g); T = cu.T
    cache = Dict{Any,Int32}(); cls = zeros(Int32, N)
    check(ccall((:mpbp_clear_node_classes, LIB), Cint, (Ptr{Cvoid},), cu.h))   # re-sync replaces every table
    for i in 1:N
        wᵢ = cu.w[i]; z = length(outedges(cu.g, i)); qi = Int(cu.q[i])
        qn = Int32[cu.q[dst(e)] for e in outedges(cu.g, i)]
        same = all(w -> w == wᵢ[1], wᵢ)
        key = (same ? wᵢ[1] : objectid(wᵢ), z, qi, qn)
        if haskey(cache, key); cls[i] = cache[key]; continue; end
        ws = same ? wᵢ[1:1] : wᵢ
        ny = Int32[nstates(ws[1], l) for l in 0:z]
        pairs = cavity_pairs(z)
        # column-major fills ([y, xk, xi] etc.) -- see include/mpbp.h for the exact layouts
        pxy = [prob_xy(w, y, xk, xi, k) for w in ws for k in 1:z for xi in 1:qi for xk in 1:qn[k] for y in 1:(z > 0 ? ny[2] : 0)]
        wj  = [prob_y_partial(w, xn, x, xj, y, z - 1, j) for w in ws for j in 1:z for y in 1:ny[z] for xj in 1:qn[j] for x in 1:qi for xn in 1:qi]
        wd  = [prob_y(w, xn, x, y, z) for w in ws for y in 1:ny[z+1] for x in 1:qi for xn in 1:qi]
        mi  = [float(prob_y0(w, y, x)) for w in ws for x in 1:qi for y in 1:ny[1]]
        pyy = [float(prob_yy(w, y, y1, y2, x, d1, d2)) for (d1, d2) in pairs for w in ws
               for x in 1:qi for y2 in 1:ny[d2+1] for y1 in 1:ny[d1+1] for y in 1:ny[d1+d2+1]]
        d1s = Int32[p[1] for p in pairs]; d2s = Int32[p[2] for p in pairs]
        id = Ref{Int32}(0)
        check(ccall((:mpbp_add_node_class, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Cint, Ptr{Int32}, Ptr{Float64}, Cint, Ptr{Int32}, Ptr{Int32},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
            cu.h, z, qi, qn, length(ws), ny, pxy, length(pairs), d1s, d2s, pyy, wj, wd, mi, id))
        cache[key] = id[]; cls[i] = id[]
    end
    check(ccall((:mpbp_set_node_classes, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), cu.h, cls))
    cu.classes_dirty = false
end

trunc_args(t::TruncBond) = (Cint(0), Cint(t.mprime), 0.0)
trunc_args(t::TruncBondMax) = (Cint(0), Cint(t.mprime), 0.0)
trunc_args(t::TruncThresh) = (Cint(1), Cint(0), Float64(t.ε))
trunc_args(t::TruncBondThresh) = (Cint(2), Cint(t.mprime), Float64(t.ε))

"""
Convergence record, the device-side twin of `CB_BP` (src/mpbp.jl:157-183): `Δs` per iteration, `f` the observable
`f(x, i)` whose means define Δ (default: the state index, as `CB_BP`'s default).  `iterate!` returns `(iters, cb)`.
"""
mutable struct CB_CuBP{TF}
    f::TF
    Δs::Vector{Float64}
end
CB_CuBP(cu::CuMPBP; f=(x, i) -> x) = CB_CuBP(f, Float64[])

"`iterate!(bp; maxiter, svd_trunc, tol, damp, nodes, shuffle_nodes, cb)` -- same keywords and return value `(iters, cb)` as src/mpbp.jl:185-198"
function iterate!(cu::CuMPBP; maxiter::Integer=5, svd_trunc=TruncThresh(1e-6), tol=1e-10, damp=0.0,
        nodes=collect(vertices(cu.g)), shuffle_nodes::Bool=true, schedule::Symbol=:sequential, showprogress=false,
        cb=CB_CuBP(cu))
    cu.classes_dirty && sync_factors!(cu)
    kind, d, ε = trunc_args(svd_trunc)
    nd = Int64.(nodes .- 1)
    qmax = maximum(Int.(cu.q))
    obs = zeros(qmax, nv(cu.g))                       # obs[x, i] = f(x, i): row-major [i][x] on the C side
    for i in 1:nv(cu.g), x in 1:Int(cu.q[i]); obs[x, i] = cb.f(x, i); end
    iters = Ref{Cint}(0); Δ = zeros(1)
    call(nodes_now, n_it, Δs) = check(ccall((:mpbp_iterate, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Cint, Float64, Float64, Float64, Cint, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Float64}, Ref{Cint}, Ptr{Float64}),
        cu.h, n_it, kind, d, ε, tol, damp, schedule == :parallel ? 1 : 0, nodes_now, length(nodes_now), C_NULL, obs, iters, Δs))
    if shuffle_nodes && schedule == :sequential
        # src/mpbp.jl:188-196: first sweep in the given order, then `sample!(nodes, vertices(bp.g), replace=false)` after
        # every sweep.  One C call per iteration with that iteration's list (nothing of size maxiter x N is
        # materialised); Δ and the tol test are the library's.
        for it in 1:maxiter
            call(nd, 1, Δ)
            push!(cb.Δs, Δ[1])
            Δ[1] < tol && return it, cb
            nd = Int64.(randperm(nv(cu.g))[1:length(nd)] .- 1)
        end
        return maxiter, cb
    end
    Δs = zeros(maxiter)
    call(nd, maxiter, Δs)
    append!(cb.Δs, Δs[1:iters[]])
    return Int(iters[]), cb
end

function beliefs(cu::CuMPBP{G,F}) where {G,F}
    out = zeros(sum(Int.(cu.q)) * (cu.T + 1))
    check(ccall((:mpbp_beliefs, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), cu.h, out))
    off = 0
    map(1:nv(cu.g)) do i
        b = [out[off+(t-1)*cu.q[i]+1:off+t*cu.q[i]] for t in 1:cu.T+1]; off += cu.q[i] * (cu.T + 1); b
    end
end

means(f, cu::CuMPBP) = [[expectation(x -> f(x, i), bt) for bt in b] for (i, b) in enumerate(beliefs(cu))]

# two-time marginals: switch on with enable_twovar!(cu; maxdist) BEFORE iterate! (they are computed with the beliefs)
enable_twovar!(cu::CuMPBP; maxdist::Integer=cu.T) =
    (check(ccall((:mpbp_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Float64), cu.h, "twovar", Float64(maxdist))); nothing)

function beliefs_tu(cu::CuMPBP)
    L = cu.T + 1; Q = maximum(Int.(cu.q))^2
    out = zeros(Q, L, L, nv(cu.g))           # [x_t + q*x_u, u, t, i] in memory order
    check(ccall((:mpbp_twovar_marginals, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), cu.h, out))
    map(1:nv(cu.g)) do i
        qi = Int(cu.q[i])
        [t < u ? reshape(out[1:qi*qi, u, t, i], qi, qi) : zeros(qi, qi) for t in 1:L, u in 1:L]
    end
end

function autocorrelations(f, cu::CuMPBP)
    map(enumerate(beliefs_tu(cu))) do (i, tv)
        expectation.(x -> f(x, i), tv)
    end
end

function autocovariances(f, cu::CuMPBP)
    μ = means(f, cu)
    covariance.(autocorrelations(f, cu), μ)
end

function pair_beliefs(cu::CuMPBP)
    sizes = [Int(cu.q[i]) * Int(cu.q[j]) for (i, j) in edges(cu.g)]
    out = zeros(sum(sizes) * (cu.T + 1)); logz = zeros(nv(cu.g))
    check(ccall((:mpbp_pair_beliefs, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), cu.h, out, logz))
    off = 0
    b = map(edges(cu.g)) do (i, j)
        qi, qj = Int(cu.q[i]), Int(cu.q[j])
        bt = [reshape(out[off+(t-1)*qi*qj+1:off+t*qi*qj], qi, qj) for t in 1:cu.T+1]; off += qi * qj * (cu.T + 1); bt
    end
    b, logz
end

function alternate_marginals(cu::CuMPBP)
    sizes = [Int(cu.q[i]) * Int(cu.q[j]) for (i, j) in edges(cu.g)]
    out = zeros(sum(sizes) * (cu.T + 1))
    check(ccall((:mpbp_alternate_marginals, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), cu.h, out))
    off = 0
    map(edges(cu.g)) do (i, j)
        qi, qj = Int(cu.q[i]), Int(cu.q[j])
        am = [reshape(out[off+(t-1)*qi*qj+1:off+t*qi*qj], qi, qj) for t in 1:cu.T]; off += qi * qj * (cu.T + 1); am
    end
end

function bethe_free_energy(cu::CuMPBP)
    f = zeros(nv(cu.g))
    check(ccall((:mpbp_free_energy, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), cu.h, f))
    if cu.g isa InfiniteBipartiteRegularGraph          # src/infinite_graph.jl:120-122
        k = cu.g.k
        return (f[1] * k[2] + f[2] * k[1]) / sum(k)
    end
    sum(f)
end

reset_messages!(cu::CuMPBP) = (check(ccall((:mpbp_reset_messages, LIB), Cint, (Ptr{Cvoid},), cu.h)); nothing)

"checkpoint / resume: message `e` as a TensorTrains MPEM2 (z = 1), and back"
function get_message(cu::CuMPBP, e::Integer)
    bonds = zeros(Int32, cu.T + 2); need = Ref{Int64}(0)
    check(ccall((:mpbp_get_message, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Float64}, Int64, Ref{Int64}), cu.h, e - 1, bonds, C_NULL, 0, need))
    data = zeros(need[])
    check(ccall((:mpbp_get_message, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Float64}, Int64, Ref{Int64}), cu.h, e - 1, bonds, data, need[], need))
    (i, j) = collect(edges(cu.g))[e][1:2]; qi, qj = Int(cu.q[i]), Int(cu.q[j]); off = 0
    tensors = map(1:cu.T+1) do t
        n = bonds[t] * bonds[t+1] * qi * qj
        A = reshape(data[off+1:off+n], Int(bonds[t]), Int(bonds[t+1]), qi, qj); off += n; A
    end
    TensorTrain(tensors)
end

export CuMPBP, CB_CuBP, sync_factors!, sync_reweightings!, get_message, sample_prior
end # module
