# QaintensorCUDA.jl -- `ccall` glue that lets Qaintensor.jl drive libqaintensor_cuda.
#
# Drop this file next to the reference's sources and apply the edits listed in INTEGRATION.md.
# It could not be executed in the build environment (no `julia` binary); every call below is the
# 1:1 transcription of the ctypes signatures in qaintensor.jl_b200/_lib.py, which the GPU tests run.
module QaintensorCUDA

const LIB = get(ENV, "QAINTENSOR_CUDA_LIB", "libqaintensor_cuda")
const QTN_C128 = Cint(0)
const QTN_EDOMAIN = Cint(-6)
const QTN_EBUSY = Cint(-7)   # a second host thread entered a device entry point (the library is not re-entrant)

struct QtnError <: Exception
    code::Cint
    msg::String
end

last_error() = unsafe_string(ccall((:qtn_last_error, LIB), Cstring, ()))

function check(rc::Cint)
    rc == 0 && return nothing
    msg = last_error()
    # QTN_EDOMAIN carries the reference's own error() strings (src/svd.jl:9, :16): re-raise unchanged
    rc == QTN_EDOMAIN ? error(msg) : throw(QtnError(rc, msg))
end

function __init__()
    # no CPU fallback: this throws QtnError(-2, ...) when no sm_100 device is usable
    check(ccall((:qtn_init, LIB), Cint, (Cint,), parse(Cint, get(ENV, "LOCAL_RANK", "0"))))
end

# ---- TensorOperations.ncon replacement (src/contract.jl:257, :263) ---------------------------------
function ncon(tensors::Vector{<:Array{ComplexF64}}, network::Vector{Vector{Int}}; order=nothing)
    nt = length(tensors)
    ranks = Cint[ndims(t) for t in tensors]
    dims = [Int64[size(t)...] for t in tensors]
    labs = [Cint.(l) for l in network]
    nout = 1
    for (t, l) in zip(tensors, network), j in 1:ndims(t)
        l[j] < 0 && (nout *= size(t, j))
    end
    out = Vector{ComplexF64}(undef, max(nout, 1))
    rank = Ref{Cint}(0)
    odims = zeros(Int64, 64)
    ord = order === nothing ? Cint[] : Cint.(order)
    tptr = Ptr{Cvoid}[pointer(t) for t in tensors]
    dptr = Ptr{Int64}[pointer(d) for d in dims]
    lptr = Ptr{Cint}[pointer(l) for l in labs]
    GC.@preserve tensors dims labs ord out begin
        check(ccall((:qtn_contract, LIB), Cint,
            (Cint, Ptr{Ptr{Cvoid}}, Ptr{Cint}, Ptr{Ptr{Int64}}, Ptr{Ptr{Cint}}, Ptr{Cint}, Cint, Cint,
             Ptr{Cvoid}, Ref{Cint}, Ptr{Int64}),
            nt, tptr, ranks, dptr, lptr, order === nothing ? C_NULL : pointer(ord), length(ord), QTN_C128,
            out, rank, odims))
    end
    r = Int(rank[])
    r == 0 ? fill(out[1]) : reshape(out[1:nout], Tuple(odims[1:r]))
end

# ---- contraction_order(net) replacement (src/network2graph.jl:429-446) --------------------------------
function treewidth_perm(ntensors::Integer, pairs::Vector{NTuple{4,Int}})
    nc = length(pairs)
    flat = Cint[x for p in pairs for x in p]
    perm = Vector{Cint}(undef, max(nc, 1))
    tw = Ref{Cint}(0)
    check(ccall((:qtn_order_treewidth, LIB), Cint, (Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ref{Cint}),
                ntensors, nc, flat, perm, tw))
    Int.(perm[1:nc]), Int(tw[])
end

# Drop-in for `Qaintensor.optimize_contraction_order!(net)` (src/network2graph.jl:473-479) on the reference's own
# network types (duck-typed: `tensors`, `contractions::Vector{Summation}` with `idx::Vector{Pair}`, `openidx`): the
# permutation comes from the bit-exact C++ restatement, the warning and error texts are the reference's
# (src/network2graph.jl:474 and :122 -- `line_graph` warns too -- and :59).
function optimize_contraction_order!(net)
    (length(net.openidx) == 0) || @warn("For TensorNetworks with open indices the treewidth algorithm is unlikely to optimize performance")
    (length(net.openidx) == 0) || @warn("All open indices are disregarded")
    pairs = NTuple{4,Int}[]
    for s in net.contractions
        length(s.idx) == 2 || error("Contractions of more than 2 tensors not supported")
        push!(pairs, (Int(s.idx[1].first), Int(s.idx[1].second), Int(s.idx[2].first), Int(s.idx[2].second)))
    end
    perm, _ = treewidth_perm(length(net.tensors), pairs)
    net.contractions = net.contractions[perm]
    nothing
end

# ---- EXTENSION: searched contraction order (opt-in replacement of optimize_contraction_order!'s heuristic) ----
# `network` as for `ncon` (positive labels = contractions); returns the label sequence to pass as `order`,
# equivalently the permutation `perm` with `net.contractions = net.contractions[perm]`.
function search_order(dims::Vector{Vector{Int64}}, network::Vector{Vector{Int}}; ntrials::Integer=256, seed::Integer=0,
                      max_log2_elems::Integer=-1)
    nt = length(dims)
    ranks = Cint[length(d) for d in dims]
    labs = [Cint.(l) for l in network]
    ncap = max(sum(length, network), 1)
    order = Vector{Cint}(undef, ncap)
    n = Ref{Cint}(0)
    cost = zeros(Float64, 4)
    dptr = Ptr{Int64}[pointer(d) for d in dims]
    lptr = Ptr{Cint}[pointer(l) for l in labs]
    GC.@preserve dims labs begin
        check(ccall((:qtn_order_search, LIB), Cint,
            (Cint, Ptr{Cint}, Ptr{Ptr{Int64}}, Ptr{Ptr{Cint}}, Cint, UInt64, Cint, Ptr{Cint}, Ref{Cint}, Ptr{Float64}),
            nt, ranks, dptr, lptr, ntrials, UInt64(seed), max_log2_elems, order, n, cost))
    end
    Int.(order[1:n[]]), (total_flops=cost[1], flops_per_slice=cost[2], nslices=cost[3], log2_max_elems=cost[4])
end

# ---- library-side network builder (qtn_net_*, csrc/network.cpp): circuit -> amplitude without Julia-side symbolics ----
mutable struct NativeNetwork
    h::Ptr{Cvoid}
    function NativeNetwork(h::Ptr{Cvoid})
        n = new(h)
        finalizer(x -> ccall((:qtn_net_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), n)
        n
    end
end

# GeneralTensorNetwork(tensors, contractions, openidx): pairs = [(t1, l1, t2, l2), ...], openidx = [(t, l), ...] (1-based)
function native_network(tensors::Vector{<:Array{ComplexF64}}, pairs::Vector{NTuple{4,Int}}, openidx::Vector{NTuple{2,Int}})
    ranks = Cint[ndims(t) for t in tensors]
    dims = [Int64[size(t)...] for t in tensors]
    tptr = Ptr{Cvoid}[pointer(t) for t in tensors]
    dptr = Ptr{Int64}[pointer(d) for d in dims]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve tensors dims begin
        check(ccall((:qtn_net_create, LIB), Cint,
            (Cint, Ptr{Ptr{Cvoid}}, Ptr{Cint}, Ptr{Ptr{Int64}}, Cint, Ptr{Cint}, Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}),
            length(tensors), tptr, ranks, dptr, length(pairs), Cint[x for p in pairs for x in p],
            length(openidx), Cint[x for p in openidx for x in p], h))
    end
    NativeNetwork(h[])
end

# tensor_circuit!(psi, cgc), non-decomposed branch (src/tensor_circuit.jl:44-51): gates = [(iwire::Tuple, matrix), ...]
function tensor_circuit!(net::NativeNetwork, gates::Vector{<:Tuple})
    mats = [Matrix{ComplexF64}(g[2]) for g in gates]
    nw = Cint[length(g[1]) for g in gates]
    wires = Cint[w for g in gates for w in g[1]]
    mptr = Ptr{Cvoid}[pointer(m) for m in mats]
    GC.@preserve mats check(ccall((:qtn_net_tensor_circuit, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Ptr{Cvoid}}),
                                  net.h, length(gates), nw, wires, mptr))
    net
end

close_wires!(net::NativeNetwork, bits) = (check(ccall((:qtn_net_close, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}), net.h, Cint.(bits))); net)

# optimize_contraction_order!(net): method = :treewidth (reference, bit-exact) or :search (extension)
function optimize_contraction_order!(net::NativeNetwork; method::Symbol=:treewidth, ntrials::Integer=256, seed::Integer=0,
                                     max_log2_elems::Integer=-1)
    check(ccall((:qtn_net_optimize_order, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, UInt64, Cint),
                net.h, method == :search ? 1 : 0, ntrials, UInt64(seed), max_log2_elems))
    net
end

# contract(net) (src/contract.jl:242-264); nout = product of the open legs' extents (1 for a closed network)
function contract(net::NativeNetwork, nout::Integer=1; max_log2_elems::Integer=-1)
    out = Vector{ComplexF64}(undef, max(nout, 1))
    rank = Ref{Cint}(0)
    odims = zeros(Int64, 64)
    check(ccall((:qtn_net_contract, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ref{Cint}, Ptr{Int64}),
                net.h, QTN_C128, max_log2_elems, out, rank, odims))
    r = Int(rank[])
    r == 0 ? fill(out[1]) : reshape(out, Tuple(odims[1:r]))
end

# ---- LinearAlgebra.svd + tail-norm rule (src/svd.jl:26-33) + max-bond cap (extension) ------------------
function svd_trunc(A::Matrix{ComplexF64}; er::Float64=-1.0, maxdim::Integer=0)
    m, n = size(A)
    r = min(m, n)
    U = Matrix{ComplexF64}(undef, m, r)
    S = Vector{Float64}(undef, r)
    Vh = Matrix{ComplexF64}(undef, r, n)
    k = Ref{Int64}(0)
    check(ccall((:qtn_svd_trunc, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Cdouble, Int64, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cvoid}, Ref{Int64}),
        A, m, n, er, maxdim, U, S, Vh, k))
    U, S, Vh, Int(k[])
end

# ---- contract_svd (src/svd.jl:7-38) -------------------------------------------------------------------
function contract_svd(T1::Array{ComplexF64}, T2::Array{ComplexF64}, indx::NTuple{2,Int}; er=0.0)
    i1, i2 = indx
    d1, d2 = Int64[size(T1)...], Int64[size(T2)...]
    newdim = (d1[1:i1-1]..., d1[i1+1:end]..., d2[1:i2-1]..., d2[i2+1:end]...)
    out = Array{ComplexF64}(undef, newdim...)
    check(ccall((:qtn_contract_svd, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Int64}, Cint, Ptr{Cvoid}, Cint, Ptr{Int64}, Cint, Cdouble, Ptr{Cvoid}),
        T1, ndims(T1), d1, i1, T2, ndims(T2), d2, i2, Float64(er), out))
    out
end

# ---- permutedims (src/contract.jl:244, src/switch.jl:29-35) ----------------------------------------------
function permutedims_gpu(A::Array{ComplexF64}, perm)
    p = collect(Int, perm)
    out = Array{ComplexF64}(undef, size(A)[p]...)
    check(ccall((:qtn_permutedims, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Cint}, Cint, Ptr{Cvoid}),
                A, ndims(A), Int64[size(A)...], Cint.(p), QTN_C128, out))
    out
end

# ---- device-resident chains (round 2): one library call each, the running tensor never returns to the host --------

# contract_svd_mps (src/mps.jl:190-201): fold of contract_svd over the tensors of an open-boundary MPS.  `tensors` are the
# Arrays of mps.tensors in order; the caller keeps the periodic-boundary check of src/mps.jl:195.
function contract_svd_mps(tensors::Vector{<:Array{ComplexF64}}; er=0.0)
    n = length(tensors)
    shape = n > 1 ? collect(size(tensors[1])[1:end-1]) : collect(size(tensors[1]))
    for j in 2:n
        sz = size(tensors[j])
        append!(shape, j < n ? sz[2:end-1] : sz[2:end])
    end
    out = Array{ComplexF64}(undef, shape...)
    ptrs = [pointer(t) for t in tensors]
    GC.@preserve tensors check(ccall((:qtn_contract_svd_fold, LIB), Cint,
        (Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Cdouble, Ptr{Cvoid}, Int64),
        n, ptrs, Int64[length(t) for t in tensors], Int64[size(t, 1) for t in tensors],
        Int64[size(t, ndims(t)) for t in tensors], Float64(er), out, length(out)))
    out
end

# arithmetic of switch!(mps, i) (src/switch.jl:18-56): returns the new (T1, T2) data of the two neighbours
function switch_adjacent(T1::Array{ComplexF64}, T2::Array{ComplexF64})
    l1 = ndims(T1) == 3 ? size(T1, 1) : 0
    r2 = ndims(T2) == 3 ? size(T2, 3) : 0
    b = size(T1, ndims(T1))
    bond = min(2 * max(l1, 1), 2 * max(r2, 1))
    U = l1 == 0 ? Array{ComplexF64}(undef, 2, bond) : Array{ComplexF64}(undef, l1, 2, bond)
    V = r2 == 0 ? Array{ComplexF64}(undef, bond, 2) : Array{ComplexF64}(undef, bond, 2, r2)
    kb = Ref{Int64}(0)
    check(ccall((:qtn_mps_switch_adjacent, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Int64}), T1, l1, b, T2, r2, U, V, kb))
    U, V
end

# SVD chain of MPO(m) (src/mpo.jl:40-75; entry = :qtn_mpo_from_matrix) and decompose!(cg) (src/decompose.jl:17-48;
# entry = :qtn_decompose): the M site tensors (2, 2, b1), (b_i, 2, 2, b_{i+1}), ..., (b_{M-1}, 2, 2)
function operator_chain(m::Matrix{ComplexF64}, M::Integer; entry::Symbol=:qtn_mpo_from_matrix)
    caps = Int64[]
    b = 1
    for i in 1:M-1
        b = min(4b, 4^(M - i))
        push!(caps, b)
    end
    shapes = [(2, 2, caps[1]); [(caps[i-1], 2, 2, caps[i]) for i in 2:M-1]; (caps[end], 2, 2)]
    sites = [Array{ComplexF64}(undef, sh...) for sh in shapes]
    ptrs = [pointer(t) for t in sites]
    bonds = Vector{Int64}(undef, max(M - 1, 1))
    GC.@preserve sites begin
        rc = entry === :qtn_decompose ?
            ccall((:qtn_decompose, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}), m, M, ptrs, bonds) :
            ccall((:qtn_mpo_from_matrix, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}), m, M, ptrs, bonds)
        check(rc)
    end
    sites
end

end # module
