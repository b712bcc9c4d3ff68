# QInchwormCUDA.jl — thin `ccall` shim that routes QInchworm.jl's three qMC worker functions to
# libqinchworm_cuda.so (include/qinchworm.h).
#
# STATUS: EXPERIMENTAL — written against the C header, NOT executed in this build environment (the image has
# no Julia toolchain).  The Python host layer `qinchworm.jl_b200/` drives the identical symbols through
# ctypes and is what the tests exercise; tests/test_bindings.py checks statically that every ccall below names
# a symbol of include/qinchworm.h with the right number of arguments.  This file shows the reference-side
# binding a maintainer adds.
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
    # (mode, order, n_pts_after, corr_idx) => compiled entry.  The reference builds fresh TopologiesInputData on
    # every inchworm! / correlator_2p call (src/inchworm.jl:380-447), so the key is what determines the compiled
    # program, not the object: entries are compiled once per Session and reused across calls.
    entry_ids::Dict{NTuple{4, Int}, Int32}
    n_top::Dict{NTuple{4, Int}, Int}   # topology count the entry was compiled from (guards against a different list)
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
    s = Session(ctx, expansion, length(grid), sum(d^2 for d in dims), Int.(dims), Dict{NTuple{4, Int}, Int32}(),
                Dict{NTuple{4, Int}, Int}(), 0)
    finalizer(x -> ccall((:qiw_destroy, lib), Cint, (Ctx,), x.ctx), s)
    init_peer!(s)      # all-reduce inside the step kernel over peer memory (no-op on one rank)
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
    key = (Int(mode), Int(td.order), Int(td.n_pts_after), Int(corr_idx))
    id = get!(s.entry_ids, key) do
        id = s.next_entry
        s.next_entry += 1
        id
    end
    if get(s.n_top, key, -1) != length(td.topologies)      # first use of the key (or a different list): compile in place
        n_top = length(td.topologies)
        pairs = Int32[x for top in td.topologies for p in top.pairs for x in (p.first, p.second)]
        parity = Int32[top.parity for top in td.topologies]
        check(s.ctx, ccall((:qiw_set_topologies, lib), Cint,
            (Ctx, Int32, Int32, Int32, Int32, Int32, Int32, Ptr{Int32}, Ptr{Int32}),
            s.ctx, id, mode, td.order, td.n_pts_after, corr_idx, n_top, pairs, parity))
        s.n_top[key] = n_top
    end
    return id
end

unpack(s::Session, v::AbstractVector{ComplexF64}) = begin
    sbm = SectorBlockMatrix(); off = 0
    for (sec, d) in enumerate(s.dims)
        sbm[sec] = (sec, reshape(v[off+1:off+d*d], d, d)); off += d * d
    end
    sbm
end

"""
`mean_std_from_randomization` (src/randomization.jl:86-100) for all entries of a step.
target_std = 0 (the default): all N_seqs scrambled sequences in ONE library call (qiw_eval_seqs); the sequences are
constructed in the reference's order — entry by entry, N_seqs sequences each, because the reference calls
mean_std_from_randomization per entry (src/inchworm.jl:142,174) — so the user's RNG stream is consumed identically.
target_std > 0: the reference's early stop (:93-99), one library call per sequence, stopping as soon as the largest
standard deviation over the entries is below the target.
"""
function eval_entries(s::Session, mode, t_i, t_w, t_f, top_data; corr_idx = 0)
    ids = Int32[entry_id!(s, td, mode, corr_idx) for td in top_data]
    N = top_data[1].N_samples
    rp = top_data[1].rand_params
    call_seqs(seqs_per_entry, n_seqs) = begin
        m = UInt32[]; x0 = UInt32[]
        for q in 1:n_seqs, j in 1:length(top_data)
            seq = seqs_per_entry[j][q]
            append!(m, vec(permutedims(seq.m))); append!(x0, seq.x)          # C layout m[D][32]
        end
        out = zeros(ComplexF64, s.bsize, length(ids), n_seqs)
        check(s.ctx, ccall((:qiw_eval_seqs, lib), Cint,
            (Ctx, Float64, Float64, Float64, Int32, Int32, Ptr{Int32}, Ptr{UInt32}, Ptr{UInt32}, UInt64, Ptr{ComplexF64}),
            s.ctx, t_i, t_w, t_f, n_seqs, length(ids), ids, m, x0, N, out))
        [out[:, :, q] for q in 1:n_seqs]
    end
    if rp.target_std <= 0 || rp.N_seqs == 1
        per_entry = [[ScrambledSobolSeq(2 * td.order, scramble_rng = rp.rng) for _ in 1:rp.N_seqs] for td in top_data]
        return call_seqs(per_entry, rp.N_seqs)
    end
    samples = Matrix{ComplexF64}[]
    for q in 1:rp.N_seqs
        one = [[ScrambledSobolSeq(2 * td.order, scramble_rng = rp.rng)] for td in top_data]
        append!(samples, call_seqs(one, 1))
        if q > 1
            _, σ = mean_std(samples)
            maximum(abs, σ) <= rp.target_std && break
        end
    end
    return samples
end

"mean / std over the randomised sequences (src/randomization.jl:93-99): std of a single sequence is NaN."
function mean_std(samples::Vector{Matrix{ComplexF64}})
    n = length(samples)
    μ = sum(samples) / n
    σ = n > 1 ? sqrt.(sum(abs2.(x .- μ) for x in samples) / (n - 1)) .+ 0im : fill(ComplexF64(NaN), size(μ))
    return μ, σ
end

"(sum, Dict order => SBM, Dict order => SBM_std), exactly as src/inchworm.jl:193-204 / :296-304 build it."
function order_sums(s::Session, top_data, μ, σ)
    orders = unique(td.order for td in top_data)
    contribs = Dict(o => zeros(SectorBlockMatrix, s.expansion.ed) for o in orders)
    contribs_std = Dict(o => zeros(SectorBlockMatrix, s.expansion.ed) for o in orders)
    for (j, td) in enumerate(top_data)
        contribs[td.order] += unpack(s, μ[:, j])
        contribs_std[td.order] += td.order == 0 ? zeros(SectorBlockMatrix, s.expansion.ed) : unpack(s, σ[:, j])
    end
    return sum(values(contribs)), contribs, contribs_std
end

"Drop-in for QInchworm.inchworm.inchworm_step_bare (src/inchworm.jl:228)."
function inchworm_step_bare(s::Session, τ_i::kd.TimeGridPoint, τ_f::kd.TimeGridPoint, top_data)
    t_i, t_f = -imag(τ_i.bpoint.val), -imag(τ_f.bpoint.val)      # imaginary-time branch: val = -iτ
    μ, σ = mean_std(eval_entries(s, 0, t_i, t_i, t_f, top_data))
    return order_sums(s, top_data, μ, σ)
end

"""
Drop-in for QInchworm.inchworm.inchworm_step (src/inchworm.jl:123).  The device's P table must be current: call
`upload_P!` once before the first step, and `scale_P!` after every host-side `set_ppgf!` + `normalize!` (it sends the
new row and λ instead of the whole table).
"""
function inchworm_step(s::Session, τ_i::kd.TimeGridPoint, τ_w::kd.TimeGridPoint, τ_f::kd.TimeGridPoint, top_data)
    t_i, t_w, t_f = (-imag(τ.bpoint.val) for τ in (τ_i, τ_w, τ_f))
    μ, σ = mean_std(eval_entries(s, 1, t_i, t_w, t_f, top_data))
    return order_sums(s, top_data, μ, σ)
end

"""
The step seam (qiw_scale_P): after `set_ppgf!(P, τ_i, τ_f, result)` and `normalize!(P, τ_f)` on the host
(src/inchworm.jl:486-488), bring the device's table up to date with ONE row and λ: row k_f := `result` (packed), then
every stored row k times exp(-λ τ_k) — λ as normalize! computes it (src/ppgf.jl:649-650), 0 for no rescaling.
"""
function scale_P!(s::Session, k_f::Integer, result::SectorBlockMatrix, λ::Real)
    row = ComplexF64[]
    for sec in 1:length(s.dims)
        append!(row, vec(result[sec][2]))
    end
    check(s.ctx, ccall((:qiw_scale_P, lib), Cint, (Ctx, Int32, Ptr{ComplexF64}, Float64), s.ctx, k_f - 1, row, λ))
end

"Drop-in for the single-τ QInchworm.inchworm.correlator_2p (src/inchworm.jl:805): returns (value, std)."
function correlator_2p(s::Session, grid::kd.ImaginaryTimeGrid, A_B_pair_idx::Integer, τ::kd.TimeGridPoint, top_data)
    t_i, t_f = 0.0, grid.contour.β
    t_w = -imag(τ.bpoint.val)
    μ, σ = mean_std(eval_entries(s, 2, t_i, t_w, t_f, top_data; corr_idx = A_B_pair_idx - 1))
    Z = QInchworm.ppgf.partition_function(s.expansion.P)
    tr_of(v) = sum(tr(unpack(s, v)[sec][2]) for sec in 1:length(s.dims))
    return sum(tr_of(μ[:, j]) for j in 1:size(μ, 2)) / Z, sum(tr_of(σ[:, j]) for j in 1:size(σ, 2)) / Z
end

"""
All grid points of one correlator in ONE launch (qiw_eval_batch): replaces the loop over τ of the
correlator_2p driver (src/inchworm.jl:1035-1046) when the default RandomizationParams are used.
Returns the vector of tr(...)/Z for grid points 2:n_τ (the τ = 0 point has order 0 only).
"""
function correlator_2p_all_tau(s::Session, grid::kd.ImaginaryTimeGrid, A_B_pair_idx::Integer, top_data)
    ids = Int32[entry_id!(s, td, 2, A_B_pair_idx - 1) for td in top_data]
    β = grid.contour.β
    times = Float64[x for k in 2:length(grid) for x in (0.0, -imag(grid[k].bpoint.val), β)]
    n_times = length(grid) - 1
    out = zeros(ComplexF64, s.bsize, length(ids), n_times)
    check(s.ctx, ccall((:qiw_eval_batch, lib), Cint,
        (Ctx, Int32, Ptr{Float64}, Int32, Ptr{Int32}, Ptr{UInt32}, Ptr{UInt32}, UInt64, Ptr{ComplexF64}),
        s.ctx, n_times, times, length(ids), ids, C_NULL, C_NULL, top_data[1].N_samples, out))
    Z = QInchworm.ppgf.partition_function(s.expansion.P)
    tr_of(v) = sum(tr(unpack(s, v)[sec][2]) for sec in 1:length(s.dims))
    return [sum(tr_of(out[:, j, k]) for j in 1:length(ids)) / Z for k in 1:n_times]
end

"""
The whole loop of inchworm! on the device (qiw_inchworm_run, src/inchworm.jl:400-493): set_ppgf! and
normalize! run on the GPU between steps.  `bare`/`bold` are the TopologiesInputData vectors inchworm!
builds (:380-447).  On return expansion.P holds the final table; the per-entry contributions are
returned as an array (bsize, n_entries, n_τ).
"""
function inchworm_run!(s::Session, bare, bold, N_samples::Integer)
    upload_P!(s)
    bare_ids = Int32[entry_id!(s, td, 0) for td in bare]
    bold_ids = Int32[entry_id!(s, td, 1) for td in bold]
    hist = zeros(ComplexF64, s.bsize, length(bare_ids) + length(bold_ids), s.n_tau)
    check(s.ctx, ccall((:qiw_inchworm_run, lib), Cint,
        (Ctx, Int32, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{UInt32}, Ptr{UInt32}, UInt64, Ptr{ComplexF64}),
        s.ctx, length(bare_ids), bare_ids, length(bold_ids), bold_ids, C_NULL, C_NULL, N_samples, hist))
    rows = zeros(ComplexF64, s.bsize, s.n_tau)
    check(s.ctx, ccall((:qiw_get_P, lib), Cint, (Ctx, Int32, Int32, Ptr{ComplexF64}), s.ctx, 0, s.n_tau, rows))
    off = 0
    for (sec, P_s) in enumerate(s.expansion.P)
        data = hasproperty(P_s, :GF) ? P_s.GF.mat.data : P_s.mat.data
        d = s.dims[sec]
        data[:, :, :] = reshape(rows[off+1:off+d*d, :], d, d, s.n_tau)
        off += d * d
    end
    return hist
end

"Map every rank's mailbox into this process so that the all-reduce runs inside the step kernel (qiw_peer_*)."
function init_peer!(s::Session)
    comm = MPI.COMM_WORLD
    MPI.Comm_size(comm) > 1 || return
    mine = zeros(UInt8, 64)
    check(s.ctx, ccall((:qiw_peer_handle, lib), Cint, (Ctx, Ptr{UInt8}), s.ctx, mine))
    all = MPI.Allgather(mine, comm)
    check(s.ctx, ccall((:qiw_peer_init, lib), Cint, (Ctx, Int32, Int32, Ptr{UInt8}),
                       s.ctx, MPI.Comm_size(comm), MPI.Comm_rank(comm), all))
end

end # module
