# QInchwormCUDA.jl — thin `ccall` shim that routes QInchworm.jl's three qMC worker functions to
# libqinchworm_cuda.so (include/qinchworm.h).
#
# STATUS: written against the C header, NOT executed in this build environment (the image has no
# Julia toolchain).  The Python host layer `qinchworm.jl_b200/` drives the identical symbols through
# ctypes and is what the tests exercise; this file shows the reference-side binding a maintainer adds.
#
# What it replaces (same signatures and return values as the reference):
#   QInchworm.inchworm.inchworm_step_bare(expansion, c, τ_i, τ_f, top_data)          src/inchworm.jl:228
#   QInchworm.inchworm.inchworm_step(expansion, c, τ_i, τ_w, τ_f, top_data)          src/inchworm.jl:123
#   QInchworm.inchworm.correlator_2p(expansion, grid, A_B_pair_idx, τ, top_data)     src/inchworm.jl:805
# Everything else (inchworm!, correlator_2p drivers, set_ppgf!, normalize!, logging) is untouched.
module QInchwormCUDA

using Keldysh; kd = Keldysh
using MPI
using QInchworm
using QInchworm.expansion: Expansion
using QInchworm.sector_block_matrix: SectorBlockMatrix
using QInchworm.scrambled_sobol: ScrambledSobolSeq
using QInchworm.spline_gf: SplineInterpolatedGF
using QInchworm.inchworm: TopologiesInputData
using LinearAlgebra: tr

const lib = "libqinchworm_cuda"
const Ctx = Ptr{Cvoid}

struct QiwOptions
    device::Int32
    warps_per_block::Int32
    reserved::NTuple{6, Int32}
end

check(ctx, rc) = rc == 0 ? nothing :
    error("libqinchworm_cuda: ", unsafe_string(ccall((:qiw_last_error, lib), Cstring, (Ctx,), ctx)))

mutable struct Session
    ctx::Ctx
    expansion::Expansion
    n_tau::Int
    bsize::Int
    dims::Vector{Int}
    entry_ids::Dict{UInt, Int32}   # objectid(td) => compiled entry
    next_entry::Int32
end

"Create a context on `device`, upload the model of `expansion` and join the NCCL communicator."
function Session(expansion::Expansion, grid::kd.ImaginaryTimeGrid; device::Integer = -1)
    ctxref = Ref{Ctx}(C_NULL)
    opts = Ref(QiwOptions(Int32(device), 0, ntuple(_ -> Int32(0), 6)))
    rc = ccall((:qiw_create, lib), Cint, (Ref{QiwOptions}, Ref{Ctx}), opts, ctxref)
    rc == 0 || error("qiw_create failed: no CUDA device (the library has no CPU path)")
    ctx = ctxref[]
    ed = expansion.ed
    dims = Int32[length(s) for s in ed.subspaces]
    S = length(dims)
    energies = Float64[]
    for P0s in expansion.P0
        append!(energies, P0s.E)                      # E .+ λ, src/exact_atomic_ppgf.jl:133
    end
    # operators as sector block matrices: pairs (operator_i, operator_f), then correlator (A, B)
    sbms = SectorBlockMatrix[]
    for (op_i, op_f) in expansion.pair_operator_mat
        push!(sbms, op_i); push!(sbms, op_f)
    end
    n_pair_ops = length(sbms)
    for (A, B) in expansion.corr_operators_mat
        push!(sbms, A); push!(sbms, B)
    end
    n_ops = length(sbms)
    op_target = fill(Int32(-1), S, n_ops)              # column-major == C [n_ops][S]
    op_off = zeros(Int64, S, n_ops)
    pool = ComplexF64[]
    for (o, sbm) in enumerate(sbms), (s_i, (s_f, mat)) in sbm
        op_target[s_i, o] = s_f - 1
        op_off[s_i, o] = length(pool)
        append!(pool, vec(mat))
    end
    n_pairs = length(expansion.pairs)
    # distinct propagator tables
    tables = Tuple{Int32, Vector{ComplexF64}}[]
    pair_table = Int32[]
    for p in expansion.pairs
        g = p.propagator
        kind, data = g isa SplineInterpolatedGF ? (Int32(1), g.GF.mat.data[1, 1, :]) : (Int32(0), g.mat.data[1, 1, :])
        idx = findfirst(t -> t[1] == kind && t[2] == data, tables)
        if isnothing(idx)
            push!(tables, (kind, data)); idx = length(tables)
        end
        push!(pair_table, idx - 1)
    end
    pair_op_i = Int32[2 * (k - 1) for k in 1:n_pairs]
    pair_op_f = Int32[2 * (k - 1) + 1 for k in 1:n_pairs]
    n_corr = length(expansion.corr_operators_mat)
    corr_A = Int32[n_pair_ops + 2 * (k - 1) for k in 1:n_corr]
    corr_B = Int32[n_pair_ops + 2 * (k - 1) + 1 for k in 1:n_corr]
    check(ctx, ccall((:qiw_set_model, lib), Cint,
        (Ctx, Int32, Ptr{Int32}, Ptr{Float64}, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{ComplexF64}, Int32,
         Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}),
        ctx, S, dims, energies, n_ops, op_target, op_off, pool, n_pairs, pair_op_i, pair_op_f, pair_table,
        n_corr, corr_A, corr_B))
    β = grid.contour.β
    for (t, (kind, data)) in enumerate(tables)
        check(ctx, ccall((:qiw_set_delta, lib), Cint, (Ctx, Int32, Int32, Int32, Float64, Ptr{ComplexF64}),
                         ctx, t - 1, kind, length(data), β, data))
    end
    check(ctx, ccall((:qiw_set_grid, lib), Cint, (Ctx, Int32, Float64), ctx, length(grid), β))
    # one process per GPU: rank 0 creates the NCCL id, MPI ships it
    comm = MPI.COMM_WORLD
    if MPI.Comm_size(comm) > 1
        id = zeros(UInt8, 128)
        MPI.Comm_rank(comm) == 0 && ccall((:qiw_comm_unique_id, lib), Cint, (Ptr{UInt8},), id)
        MPI.Bcast!(id, 0, comm)
        check(ctx, ccall((:qiw_comm_init, lib), Cint, (Ctx, Int32, Int32, Ptr{UInt8}),
                         ctx, MPI.Comm_size(comm), MPI.Comm_rank(comm), id))
    end
    s = Session(ctx, expansion, length(grid), sum(d^2 for d in dims), Int.(dims), Dict{UInt, Int32}(), 0)
    finalizer(x -> ccall((:qiw_destroy, lib), Cint, (Ctx,), x.ctx), s)
    return s
end

"Packed block vector of the current P table (layout (d_s, d_s, n_τ) per sector → rows per grid point)."
function upload_P!(s::Session)
    rows = zeros(ComplexF64, s.bsize, s.n_tau)
    off = 0
    for (sec, P_s) in enumerate(s.expansion.P)
        data = hasproperty(P_s, :GF) ? P_s.GF.mat.data : P_s.mat.data
        d = s.dims[sec]
        rows[off+1:off+d*d, :] = reshape(data, d * d, s.n_tau)
        off += d * d
    end
    check(s.ctx, ccall((:qiw_set_P, lib), Cint, (Ctx, Int32, Int32, Ptr{ComplexF64}), s.ctx, 0, s.n_tau, rows))
end

function entry_id!(s::Session, td::TopologiesInputData, mode::Integer, corr_idx::Integer = 0)
    get!(s.entry_ids, hash((objectid(td), mode, corr_idx))) do
        id = s.next_entry
        s.next_entry += 1
        n_top = length(td.topologies)
        pairs = Int32[x for top in td.topologies for p in top.pairs for x in (p.first, p.second)]
        parity = Int32[top.parity for top in td.topologies]
        check(s.ctx, ccall((:qiw_set_topologies, lib), Cint,
            (Ctx, Int32, Int32, Int32, Int32, Int32, Int32, Ptr{Int32}, Ptr{Int32}),
            s.ctx, id, mode, td.order, td.n_pts_after, corr_idx, n_top, pairs, parity))
        id
    end
end

unpack(s::Session, v::AbstractVector{ComplexF64}) = begin
    sbm = SectorBlockMatrix(); off = 0
    for (sec, d) in enumerate(s.dims)
        sbm[sec] = (sec, reshape(v[off+1:off+d*d], d, d)); off += d * d
    end
    sbm
end

"One library call per scrambled sequence (mean_std_from_randomization stays in Julia, src/randomization.jl:86)."
function eval_entries(s::Session, mode, t_i, t_w, t_f, top_data; corr_idx = 0)
    ids = Int32[entry_id!(s, td, mode, corr_idx) for td in top_data]
    N = top_data[1].N_samples
    rp = top_data[1].rand_params
    samples = Matrix{ComplexF64}[]
    for _ in 1:rp.N_seqs
        m = UInt32[]; x0 = UInt32[]
        for td in top_data
            seq = ScrambledSobolSeq(2 * td.order, scramble_rng = rp.rng)   # consumes the user's RNG stream
            append!(m, vec(permutedims(seq.m))); append!(x0, seq.x)        # C layout m[D][32]
        end
        out = zeros(ComplexF64, s.bsize, length(ids))
        check(s.ctx, ccall((:qiw_eval, lib), Cint,
            (Ctx, Float64, Float64, Float64, Int32, Ptr{Int32}, Ptr{UInt32}, Ptr{UInt32}, UInt64, Ptr{ComplexF64}),
            s.ctx, t_i, t_w, t_f, length(ids), ids, m, x0, N, out))
        push!(samples, out)
    end
    return samples
end

# The three worker overloads build (sum, Dict order => SBM, Dict order => SBM_std) / (ComplexF64, ComplexF64)
# from `eval_entries` exactly as src/inchworm.jl:193-204, 296-304, 881-889 do; omitted here for brevity
# of the untested shim — see qinchworm.jl_b200/inchworm.py (_order_sums, correlator_2p) for the tested
# statement of the same glue.

end # module
