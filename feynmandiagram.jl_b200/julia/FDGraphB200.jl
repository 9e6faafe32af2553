# FDGraphB200.jl -- the Julia side of the drop-in: `ccall` shim over libfdgraph.so (include/fdgraph.h).
#
# NEVER EXECUTED: there is no julia binary in the build image or on the GPU box (SURVEY.md D2).  What is checked instead:
# the struct layouts and the call sequence of compile() / leafgen() are restated in C (tests/abi/julia_layout.c: the
# fieldoffsets of the structs below are static-asserted against include/fdgraph.h and the calls are made in this file's
# order), and the FDGRAPH file written by save_graph() is the format tests/test_lowering.py and tests/test_gpu_parity.py
# read back.  The shim is kept deliberately thin -- it only flattens Graph objects through the four getters the reference emitter itself uses
# (id / operator / subgraphs / subgraph_factors, src/backend/static.jl:106-124) and forwards pointers -- and it is
# mirrored 1:1 by the Python ctypes binding (feynmandiagram.jl_b200/_capi.py), which IS tested against the library.
#
# Usage (drop-in for Compilers.compile / eval_graph!, src/backend/static.jl:221-227):
#     using FeynmanDiagram, CUDA
#     include("FDGraphB200.jl")
#     eval_graph!, leafmap = FDGraphB200.compile(diags)          # same (callable, leafmap::Dict{Int,Graph})
#     leafVal = CUDA.rand(Float64, B, length(leafmap)) .+ 0.5    # B x L, column-major: batch index unit-stride
#     root    = CUDA.zeros(Float64, B, length(diags))
#     eval_graph!(root, leafVal)                                 # all B samples in one call
module FDGraphB200

using FeynmanDiagram
import FeynmanDiagram.ComputationalGraphs: AbstractGraph, id, operator, subgraphs, subgraph_factors, Sum, Prod, Power, Unitary

const LIB = get(ENV, "FDGRAPH_LIB", joinpath(@__DIR__, "..", "libfdgraph.so"))

struct GraphDesc
    n_nodes::Int64
    n_edges::Int64
    node_id::Ptr{Int64}
    node_op::Ptr{Int32}
    node_pow::Ptr{Int32}
    child_ptr::Ptr{Int64}
    child_node::Ptr{Int32}
    child_factor::Ptr{Float64}
    n_graphs::Int64
    graphs::Ptr{Int32}
    n_roots::Int64
    root_id::Ptr{Int64}
end

struct Options
    dtype::Int32
    max_slots::Int32
    prefetch::Int32
    schedule::Int32
    backend::Int32
    jit_segment::Int32
    cse::Int32   # 0 = automatic (the planner's model decides), 1 = always, -1 = never merge common sub-expressions
    fma::Int32   # 1 = multiplies may be fused into adds (opt-in, not bit-identical)
end

opcode(::Type{Unitary}) = Int32(0)
opcode(::Type{Sum}) = Int32(1)
opcode(::Type{Prod}) = Int32(2)
opcode(::Type{Power{N}}) where {N} = Int32(3)
opcode(op) = error("Static representation for computational graph nodes with operator $(op) not yet implemented!")
pow_n(::Type{Power{N}}) where {N} = Int32(N)
pow_n(_) = Int32(0)

check(rc) = rc == 0 || error("libfdgraph: " * unsafe_string(ccall((:fdg_last_error, LIB), Cstring, ())))

mutable struct Evaluator
    handle::Ptr{Cvoid}
    n_leaves::Int
    n_roots::Int
    last_root::Int
    function Evaluator(h, L, R, last)
        ev = new(h, L, R, last)
        finalizer(e -> ccall((:fdg_destroy, LIB), Cint, (Ptr{Cvoid},), e.handle), ev)
        return ev
    end
end

# one entry per node OBJECT, children before parents (any order is accepted by the library); returns the arrays of
# fdg_graph_desc plus the node objects behind them
function _flatten(graphs::AbstractVector{G}) where {G<:AbstractGraph}
    nodes = G[]
    index = IdDict{G,Int32}()
    function visit(g)
        haskey(index, g) && return
        for s in subgraphs(g)
            visit(s)
        end
        push!(nodes, g)
        index[g] = Int32(length(nodes) - 1)
    end
    foreach(visit, graphs)
    node_id = Int64[id(g) for g in nodes]
    node_op = Int32[isempty(subgraphs(g)) && !(operator(g) <: Union{Sum,Prod,Power,Unitary}) ? 0 : opcode(operator(g)) for g in nodes]
    node_pow = Int32[pow_n(operator(g)) for g in nodes]
    child_ptr = Int64[0]
    child_node = Int32[]
    child_factor = Float64[]
    for g in nodes
        for (s, f) in zip(subgraphs(g), subgraph_factors(g))
            push!(child_node, index[s])
            push!(child_factor, Float64(f))
        end
        push!(child_ptr, length(child_node))
    end
    gidx = Int32[index[g] for g in graphs]
    return node_id, node_op, node_pow, child_ptr, child_node, child_factor, gidx, nodes
end

"""
    compile(graphs; root=[id(g) for g in graphs], dtype=Float64) -> (eval_graph!, leafmap)

Same contract as `Compilers.compile` (static.jl:221-227): `leafmap[k]` is the leaf graph whose value is column `k`
of `leafVal` (1-based, identical numbering to the reference).
"""
function compile(graphs::AbstractVector{G}; root::AbstractVector{Int}=[id(g) for g in graphs], dtype::DataType=Float64) where {G<:AbstractGraph}
    node_id, node_op, node_pow, child_ptr, child_node, child_factor, gidx, nodes = _flatten(graphs)
    rootid = Int64.(root)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve node_id node_op node_pow child_ptr child_node child_factor gidx rootid begin
        desc = Ref(GraphDesc(length(nodes), length(child_node), pointer(node_id), pointer(node_op), pointer(node_pow),
            pointer(child_ptr), pointer(child_node), pointer(child_factor), length(gidx), pointer(gidx), length(rootid), pointer(rootid)))
        opts = Ref(Options(dtype <: Complex ? 1 : 0, 0, 0, 0, 0, 0, 0, 0))
        check(ccall((:fdg_compile, LIB), Cint, (Ref{GraphDesc}, Ref{Options}, Ref{Ptr{Cvoid}}), desc, opts, h))
    end
    stats = zeros(Int64, 16)
    check(ccall((:fdg_stats, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), h[], stats))
    L, R = Int(stats[1]), Int(stats[3])
    leaf_node = zeros(Int32, max(L, 1))
    check(ccall((:fdg_leafmap, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), h[], leaf_node))
    last = Ref{Int32}(0)
    check(ccall((:fdg_last_root, LIB), Cint, (Ptr{Cvoid}, Ref{Int32}), h[], last))
    leafmap = Dict{Int,G}(k => nodes[leaf_node[k]+1] for k in 1:L)
    ev = Evaluator(h[], L, R, Int(last[]) + 1)
    return ev, leafmap   # `ev(root, leafVal)` is the generated function (callable struct below)
end

"""
    save_graph(graphs, path; root=[id(g) for g in graphs])

Writes the flattened graph as an FDGRAPH file (include/fdgraph.h: fdg_graph_write / fdg_compile_file), so that a graph
built here can be evaluated on a machine that has the GPU but no Julia (`fdg_compile_file`, `fd.compile_file` in Python).
"""
function save_graph(graphs::AbstractVector{G}, path::AbstractString; root::AbstractVector{Int}=[id(g) for g in graphs]) where {G<:AbstractGraph}
    node_id, node_op, node_pow, child_ptr, child_node, child_factor, graph_idx, _ = _flatten(graphs)
    open(path, "w") do io
        write(io, UInt8['F', 'D', 'G', 'R', 'A', 'P', 'H', 0x01])
        write(io, Int64(length(node_id)), Int64(length(child_node)), Int64(length(graph_idx)), Int64(length(root)))
        write(io, node_id, node_op, node_pow, child_ptr, child_node, child_factor, graph_idx, Int64.(root))
    end
    return path
end

# the value returned by compile() is called like the reference's generated function
(ev::Evaluator)(root, leafVal) = eval_graph!(ev, root, leafVal)

# device matrices (CuArray{T,2}, B x L and B x R, column-major): one launch sequence for the whole batch
function eval_graph!(ev::Evaluator, root::AbstractMatrix, leafVal::AbstractMatrix; stream::Ptr{Cvoid}=C_NULL)
    B = size(leafVal, 1)
    size(root, 1) == B || throw(DimensionMismatch("root and leafVal must have the same number of rows (samples)"))
    (size(leafVal, 2) >= ev.n_leaves && size(root, 2) >= ev.n_roots) || throw(BoundsError())
    # the raw pointers handed to ccall do not keep the arrays alive: GC.@preserve does, for the duration of the call
    GC.@preserve root leafVal begin
        if leafVal isa Array   # host matrices: fdg_eval_host does H2D / kernel / D2H
            check(ccall((:fdg_eval_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64),
                ev.handle, pointer(leafVal), stride(leafVal, 2), pointer(root), stride(root, 2), B))
        else                   # CuArray: pointer() gives a CuPtr, reinterpret as the raw device address
            check(ccall((:fdg_eval, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}),
                ev.handle, reinterpret(Ptr{Cvoid}, pointer(leafVal)), stride(leafVal, 2), reinterpret(Ptr{Cvoid}, pointer(root)),
                stride(root, 2), B, stream))
        end
    end
    return ev.last_root >= 1 ? view(root, :, ev.last_root) : nothing
end

# vectors are the B = 1 case of the reference's generated function: returns the last root value (static.jl:127,131)
function eval_graph!(ev::Evaluator, root::AbstractVector, leafVal::AbstractVector)
    r = reshape(root, 1, :)
    eval_graph!(ev, r, reshape(leafVal, 1, :))
    return ev.last_root >= 1 ? root[ev.last_root] : nothing
end

# ---- leaf values from (K, tau) on the device (include/fdgraph.h, fdg_leafgen_*; SURVEY §8f N1) -----------------------
struct LeafGenDesc
    n_leaves::Int64
    leaf_type::Ptr{Int32}
    leaf_order::Ptr{Int32}
    tau_in::Ptr{Int32}
    tau_out::Ptr{Int32}
    loop_index::Ptr{Int32}
    n_basis::Int64
    n_loops::Int64
    dim::Int64
    n_tau::Int64
    loop_basis::Ptr{Float64}
    kF::Float64
    beta::Float64
    lambda::Float64
end

mutable struct LeafGen
    handle::Ptr{Cvoid}
    n_loops::Int
    dim::Int
    n_tau::Int
end

"""
    leafgen(leafstat, loopbasis; partition=1, dim=3, kF, β, λ)

`leafstat, loopbasis = FrontEnds.leafstates(leaf_maps, maxloopNum)` (src/frontend/frontends.jl:175-232) as returned by the
reference; 1-based Julia indices are shifted to the library's 0-based ones here.
"""
function leafgen(leafstat, loopbasis::Vector{Vector{Float64}}; partition::Int=1, dim::Int=3, kF::Float64, β::Float64, λ::Float64)
    _, leafType, leafOrders, leafInTau, leafOutTau, leafLoopIndex = leafstat
    L = length(leafType[partition])
    ltype = Int32.(leafType[partition])
    lorder = Int32[get(leafOrders[partition][l], c, 0) for c in 1:2, l in 1:L]   # 2 x L, column-major == [l*2 + c]
    tin = Int32.(leafInTau[partition] .- 1)
    tout = Int32.(leafOutTau[partition] .- 1)
    lidx = Int32.(leafLoopIndex[partition] .- 1)
    basis = reduce(hcat, loopbasis)                                               # n_loops x n_basis == [i*n_loops + j]
    n_tau = max(maximum(tin), maximum(tout)) + 1
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve ltype lorder tin tout lidx basis begin
        desc = Ref(LeafGenDesc(L, pointer(ltype), pointer(lorder), pointer(tin), pointer(tout), pointer(lidx),
            size(basis, 2), size(basis, 1), dim, n_tau, pointer(basis), kF, β, λ))
        check(ccall((:fdg_leafgen_create, LIB), Cint, (Ref{LeafGenDesc}, Ref{Ptr{Cvoid}}), desc, h))
    end
    g = LeafGen(h[], size(basis, 1), dim, n_tau)
    finalizer(x -> ccall((:fdg_leafgen_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), g)
    return g
end

# host matrices K (B x dim*n_loops, column (j-1)*dim + c) and T (B x n_tau) -> the R per-root sums over the B samples
function eval_generated(ev::Evaluator, g::LeafGen, K::Matrix{Float64}, T::Matrix{Float64})
    B = size(K, 1)
    (size(K, 2) == g.dim * g.n_loops && size(T) == (B, g.n_tau)) || throw(DimensionMismatch("K must be B x dim*n_loops, T must be B x n_tau"))
    acc = zeros(Float64, ev.n_roots)
    check(ccall((:fdg_eval_generated_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Ptr{Float64}),
        ev.handle, g.handle, K, T, B, B, acc))
    return acc
end

end # module
