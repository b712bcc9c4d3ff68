/* qinchworm.h — C ABI of libqinchworm_cuda.so: the B200 (sm_100a) implementation of the quasi
 * Monte Carlo diagram-evaluation hot path of QInchworm.jl.
 *
 * The reference (pure Julia + MPI.jl) has no FFI seam of its own; this header defines the one a
 * thin Julia `ccall` shim binds (julia/QInchwormCUDA.jl).  Each entry point cites the
 * reference interface it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain C linkage, plain pointers and sizes; no exceptions cross the boundary;
 *   - complex numbers are passed as interleaved (re, im) doubles == Julia ComplexF64 / C99 double
 *     _Complex; matrices are column-major (Julia layout);
 *   - every pointer argument is HOST memory borrowed for the duration of the call, unless the
 *     name starts with `dev_`;
 *   - a "packed block vector" is the concatenation over sectors s = 0..S-1 of the d_s x d_s
 *     column-major block of sector s (what src/mpi.jl:104-127 packs for its all-reduce);
 *   - sectors, operators, pairs, tables, entries are 0-based; topology vertices are 1-based as
 *     in src/diagrammatics.jl (vertex 1 is the latest time);
 *   - a context is bound to ONE CUDA device and is not thread-safe; calls are synchronous
 *     (results are in host memory on return) unless stated otherwise;
 *   - return value 0 = success, otherwise one of qiw_status; qiw_last_error() gives the text.
 *
 * There is no CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef QINCHWORM_H
#define QINCHWORM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qiw_context qiw_context;

typedef enum {
    QIW_OK = 0,
    QIW_ERR_BAD_ARG = 1,
    QIW_ERR_CUDA = 2,
    QIW_ERR_NCCL = 3,
    /* the `@assert s_i == s_f` of src/topology_eval.jl:462 */
    QIW_ERR_BLOCK_OFFDIAGONAL = 4,
    /* e.g. seq_type = RandomSeq (src/inchworm.jl:339): only Sobol sequences run on the device */
    QIW_ERR_UNSUPPORTED = 5
} qiw_status;

/* How a scalar propagator Delta(t_f, t_i) is evaluated between grid points. */
typedef enum {
    /* plain kd.ImaginaryTimeGF: Keldysh.jl's generic bilinear / triangular grid interpolation
     * (call site src/topology_eval.jl:414; rule in DESIGN.md) */
    QIW_DELTA_GRID_BILINEAR = 0,
    /* SplineInterpolatedGF: natural cubic spline in t_f - t_i (src/spline_gf.jl:193-219) */
    QIW_DELTA_CUBIC_SPLINE = 1
} qiw_delta_kind;

/* What kind of diagrams an entry (one TopologiesInputData, src/inchworm.jl:60-98) holds. */
typedef enum {
    QIW_MODE_BARE = 0, /* inchworm_step_bare: bare propagators P0, RootTransform        (:228) */
    QIW_MODE_BOLD = 1, /* inchworm_step: bold propagators P, DoubleSimplexRootTransform  (:123) */
    QIW_MODE_CORR = 2  /* correlator_2p at one tau: operator nodes A, B; trace           (:805) */
} qiw_mode;

#define QIW_DEVICE_CURRENT (-1)
/* Planning-only context: no CUDA call is ever made; qiw_set_model, qiw_set_topologies,
 * qiw_entry_stats and qiw_entry_program work (host-side compilation of the traversal programs),
 * every compute entry point returns QIW_ERR_CUDA.  Used to inspect programs on a GPU-less host. */
#define QIW_DEVICE_NONE (-2)

typedef struct {
    int32_t device;          /* CUDA device ordinal; QIW_DEVICE_CURRENT; QIW_DEVICE_NONE */
    int32_t warps_per_block; /* 0 = library default */
    int32_t reserved[6];
} qiw_options;

/* ---- life cycle -------------------------------------------------------------------------- */

int qiw_create(const qiw_options* opts, qiw_context** ctx);
int qiw_destroy(qiw_context* ctx);
const char* qiw_last_error(const qiw_context* ctx);
/* "major.minor.patch" of the library. */
const char* qiw_version(void);

/* ---- problem definition: the fields of `Expansion` (src/expansion.jl:110-134) ---------------- */

/* Sectors, atomic energies, operators and interaction pairs.
 *   n_sectors, sector_dims[S]       : invariant subspaces of EDCore and their dimensions
 *   energies[sum_s d_s]             : E_s[i] + lambda of ExactAtomicPPGF (src/exact_atomic_ppgf.jl:133)
 *   n_ops                           : number of operators given as sector-block matrices
 *   op_target[n_ops][S]             : target sector of the block leaving sector s, or -1
 *                                     (SectorBlockMatrix: right index => (left index, block),
 *                                      src/sector_block_matrix.jl:45)
 *   op_mat_off[n_ops][S]            : offset (in complex elements) of that block in op_pool
 *   op_pool                         : blocks, each d_target x d_source column-major complex
 *   n_pairs, pair_op_i/pair_op_f    : InteractionPair.operator_i / operator_f as operator indices
 *                                     (exp.pair_operator_mat, src/expansion.jl:169-173)
 *   pair_table[n_pairs]             : which Delta table the pair's propagator is
 *   n_corr, corr_op_A/corr_op_B     : exp.corr_operators_mat (src/expansion.jl:174-178)
 * subspace_attachable_pairs (src/expansion.jl:180-183) is derived inside. */
int qiw_set_model(qiw_context* ctx, int32_t n_sectors, const int32_t* sector_dims,
                  const double* energies, int32_t n_ops, const int32_t* op_target,
                  const int64_t* op_mat_off, const double* op_pool, int32_t n_pairs,
                  const int32_t* pair_op_i, const int32_t* pair_op_f, const int32_t* pair_table,
                  int32_t n_corr, const int32_t* corr_op_A, const int32_t* corr_op_B);

/* Imaginary-time grid of the bold propagators: kd.ImaginaryTimeGrid(contour, n_tau). */
int qiw_set_grid(qiw_context* ctx, int32_t n_tau, double beta);

/* One scalar propagator table: values[n_knots] complex = GF.mat.data[1,1,:] on [0, beta]. */
int qiw_set_delta(qiw_context* ctx, int32_t table_id, int32_t kind, int32_t n_knots, double beta,
                  const double* values);

/* set_ppgf! + normalize! across the step seam without re-sending the table (src/ppgf.jl:495-504,646-668): writes row
 * k_f of the device's P table (`row`: bsize complex values; NULL keeps the device's row) and then multiplies every
 * stored row k by exp(-lambda * tau_k) (lambda = 0: no rescaling).  The host applies the same lambda to its own copy
 * (normalize!), so the two tables agree to rounding; qiw_set_P re-synchronises them exactly.  Asynchronous: no host
 * synchronisation, the caller's `row` may be reused at return. */
int qiw_scale_P(qiw_context* ctx, int32_t k_f, const double* row, double lambda);

/* Rows [first, first+count) of the bold propagator table: for each grid point the packed block
 * vector of P_s(tau_k) (P[s].mat.data, layout (d_s, d_s, n_tau)).  Replaces nothing in the
 * reference: it is how `expansion.P` reaches the device after set_ppgf!/normalize!
 * (src/ppgf.jl:474-504,646-668). */
int qiw_set_P(qiw_context* ctx, int32_t first, int32_t count, const double* rows);
int qiw_get_P(qiw_context* ctx, int32_t first, int32_t count, double* rows);

/* One TopologiesInputData (src/inchworm.jl:60-98): `n_top` topologies of `order` arcs,
 * pairs[n_top][order][2] (arc a => b, 1-based vertices, a < b) and parity[n_top], straight from
 * get_topologies_at_order (src/diagrammatics.jl:322-337).  The library compiles them against the
 * model into its traversal programs (replaces the TopologyEvaluator constructor,
 * src/topology_eval.jl:249-330, and the run-time pruning of _traverse_configuration_tree!,
 * :454-556).  For QIW_MODE_CORR pass corr_idx = index of the (A, B) pair, else 0.
 * Must be called after qiw_set_model. */
int qiw_set_topologies(qiw_context* ctx, int32_t entry_id, int32_t mode, int32_t order,
                       int32_t n_pts_after, int32_t corr_idx, int32_t n_top, const int32_t* pairs,
                       const int32_t* parity);

/* Work counters of a compiled entry: topologies, surviving configurations (leaves), live tree
 * edges and the chain FLOPs per sample of the reference algorithm. */
int qiw_entry_stats(qiw_context* ctx, int32_t entry_id, int64_t* n_top, int64_t* n_leaves,
                    int64_t* n_edges, double* flops_per_sample);

/* Disassembly of a compiled entry (inspection / tests).  Call with NULL arrays to get the sizes.
 *   words[n_words]      : pre-order node stream (layout: csrc/qiw_host.hpp make_word)
 *   tree_off[n_trees+1] : word offset of every tree root
 *   coefs[n_coefs]      : distinct leaf coefficients (complex)
 *   dslots[n_dslots][3] : (tail position, head position, Delta table) of every pair-interaction slot
 *   pos_src[20]         : per backbone position (1-based): >= 0 index of the sample's time,
 *                         -1 t_i, -2 t_w, -3 t_f
 *   info[4]             : n_nodes, number of propagator slots nP, S, scalar flag */
int qiw_entry_program(qiw_context* ctx, int32_t entry_id, int64_t* n_words, uint64_t* words,
                      int64_t* n_trees, uint32_t* tree_off, int64_t* n_coefs, double* coefs,
                      int64_t* n_dslots, int32_t* dslots, int32_t* pos_src, int32_t* info);

/* Factorised configuration records of a compiled entry of a 1x1-block model — what the step kernel
 * executes (layout: csrc/qiw_host.hpp EntryProgram::rec2).  Call with NULL arrays to get the sizes.
 *   info[8]  : K (segments), L2 (= K + order operands per configuration), n_leaves, nSeg,
 *              seg_stride, nP, nD, 0
 *   rec2[n_leaves][L2 + 1], segdef[nSeg][seg_stride] */
int qiw_entry_records(qiw_context* ctx, int32_t entry_id, int32_t* info, uint32_t* rec2, uint16_t* segdef);

/* Lane program of a compiled entry of a 1x1-block model — what the step kernel executes with lane = sample
 * (layout: csrc/qiw_host.hpp EntryProgram::lane_*).  Configurations sharing all pair-interaction operands and the
 * initial sector form groups, cut into records of M = 4, 2 or 1 members:
 *     record value = prod(T[Delta slots]) * sum_members prod(T[segment slots of the member])
 * with every member's coefficient folded into its first segment product.  Call with NULL arrays to get the sizes.
 *   info[8]             : n_sections, n_items, nSegL, seg_stride, K, order, first slot of the segment table, cost
 *   sections[n_sections][4] : sector code, M, number of records, first item; sector code = initial sector a |
 *                         (sector b + 1) << 16: with a second sector (records shared by two initial sectors, from
 *                         order 5 on; QIW_LANE_DUAL=0 / 1 overrides) the first M / 2 members of a record are summed
 *                         for sector a and the others for sector b, the Delta operands being loaded once for both
 *   items[n_items]      : per record `order` Delta slots, then M * K segment slots, padded to a multiple of 8
 *   segdef[nSegL][seg_stride] : propagator slots of every segment product (0xFFFF = unused)
 *   seg_coef[nSegL]     : index of the coefficient folded into the product (0xFFFF = none) */
int qiw_entry_lane_program(qiw_context* ctx, int32_t entry_id, int32_t* info, int32_t* sections, uint32_t* items,
                           uint16_t* segdef, uint16_t* seg_coef);

/* Walk units of a compiled entry of a sector-block model (blocks larger than 1x1) — what block_walk_kernel
 * replays: the pruned configuration trees of src/topology_eval.jl:454-556 cut into sub-trees of bounded cost,
 * each emitted once per group of at most two columns of the running product (word layout: csrc/qiw_kernels.cu,
 * "Expanded program word").  Call with NULL arrays to get the sizes.
 *   unit_off[n_units + 1] : first word of every unit;  words[n_words][4] */
int qiw_entry_walk_units(qiw_context* ctx, int32_t entry_id, int64_t* n_units, int64_t* n_words, uint32_t* unit_off,
                         uint32_t* words);

/* ---- the hot path ---------------------------------------------------------------------------- */

/* Evaluate `n_entries` entries at the fixed times (t_i, t_w, t_f) with N_total Sobol points each.
 * Replaces, per entry, the body of
 *     mean_std_from_randomization(...) do seq ... contour_integral(t -> eval(td.topologies, t), ...)
 * i.e. src/inchworm.jl:174-191 (bold), :278-296 (bare), :853-879 (correlator), including
 * scrambled_sobol.next!/skip! (src/scrambled_sobol.jl:158-197), the Root / DoubleSimplexRoot maps
 * (src/qmc_integrate.jl:225-235,425-463), contour_integral / qmc_integral (:497-507,597-612), the
 * TopologyEvaluator call (src/topology_eval.jl:350-437), rank_sub_range (src/mpi.jl:49-54) and
 * all_reduce! (src/mpi.jl:104-127).
 *   t_i, t_w, t_f : imaginary times (BARE ignores t_w; CORR: t_i = 0, t_w = tau, t_f = beta)
 *   sobol_m       : per entry, concatenated: direction numbers m[2*order][32] (uint32, bit 31 =
 *                   most significant digit) of that entry's ScrambledSobolSeq; NULL = unscrambled
 *   sobol_x0      : per entry, concatenated: first point x[2*order]; NULL = zeros
 *   N_total       : td.N_samples.  This rank evaluates its rank_sub_range of 0..N_total-1.
 *   out           : [n_entries] packed block vectors (complex): the qMC estimate of each entry,
 *                   including (-i)^d, the Jacobian and 1/N, summed over all ranks of the
 *                   communicator.  Order-0 entries are evaluated exactly (no sampling).
 * For QIW_MODE_CORR the caller takes the trace (src/inchworm.jl:869) and divides by Z (:889). */
int qiw_eval(qiw_context* ctx, double t_i, double t_w, double t_f, int32_t n_entries,
             const int32_t* entry_ids, const uint32_t* sobol_m, const uint32_t* sobol_x0,
             uint64_t N_total, double* out);

/* Batched form of qiw_eval: the same entries at `n_times` independent time triples
 * times[n_times][3] = (t_i, t_w, t_f) in ONE launch; out[n_times][n_entries] packed block vectors.
 * Replaces the loop over grid points of the correlator_2p driver (src/inchworm.jl:1035-1046), whose
 * (operator pair, tau) evaluations are independent because P is final.  The same Sobol sequence is
 * used for every triple (RandomizationParams() default). */
int qiw_eval_batch(qiw_context* ctx, int32_t n_times, const double* times, int32_t n_entries,
                   const int32_t* entry_ids, const uint32_t* sobol_m, const uint32_t* sobol_x0,
                   uint64_t N_total, double* out);

/* Randomised-qMC form of qiw_eval: the same entries at the same times with `n_seqs` independently scrambled
 * Sobol sequences in ONE launch — the loop of mean_std_from_randomization (src/randomization.jl:93-99) when
 * no early stop is requested (target_std = 0).  sobol_m / sobol_x0 hold, per sequence, the per-entry
 * concatenated parameters exactly as in qiw_eval; out[n_seqs][n_entries] packed block vectors, from which the
 * host forms mean and standard deviation. */
int qiw_eval_seqs(qiw_context* ctx, double t_i, double t_w, double t_f, int32_t n_seqs, int32_t n_entries,
                  const int32_t* entry_ids, const uint32_t* sobol_m, const uint32_t* sobol_x0, uint64_t N_total,
                  double* out);

/* Same, but evaluates only Sobol indices [start, start+count) and does NOT all-reduce: the
 * rank-local partial sum (already divided by N_total).  Used by hosts that own the collective. */
int qiw_eval_range(qiw_context* ctx, double t_i, double t_w, double t_f, int32_t n_entries,
                   const int32_t* entry_ids, const uint32_t* sobol_m, const uint32_t* sobol_x0,
                   uint64_t start, uint64_t count, uint64_t N_total, double* out);

/* Per-sample evaluator values at explicit times (no transform, Jacobian or 1/N): the return
 * value of (eval::TopologyEvaluator)(topologies, times), src/topology_eval.jl:350, for
 * n_samples time tuples times[n_samples][2*order] (descending).  out[n_samples] packed block
 * vectors.  This is what test/topology_eval.jl:137-141 records. */
int qiw_eval_at_times(qiw_context* ctx, int32_t entry_id, double t_i, double t_w, double t_f,
                      int32_t n_samples, const double* times, double* out);

/* Device time (ms, CUDA events) spent in the kernels of the last qiw_eval* call, so that the
 * host's timer sections (TimerOutputs "Evaluation", src/inchworm.jl:171) keep their meaning. */
int qiw_last_device_ms(qiw_context* ctx, double* ms);
/* Number of kernel launches issued by this context so far. */
int qiw_launch_count(qiw_context* ctx, int64_t* n);

/* Per-kernel device timing (CUDA events on the launching stream around every launch).  Classes:
 * 0 step kernel in complex arithmetic, 1 step kernel in real arithmetic (1x1-block models), 2 persistent run kernel
 * (all bold steps of qiw_inchworm_run in one launch), 3 FP64 tensor-core kernel for sector blocks of 5 to 8 rows, 4 reduction,
 * 5 per-step state update, 6 NCCL all-reduce, 7 step kernel for sector blocks larger than 1x1.  Profiling serialises nothing but adds two event records per launch; keep it
 * off for timed runs.  qiw_profile_read synchronises, returns accumulated ms and launch counts per
 * class (arrays of QIW_PROFILE_CLASSES) and optionally resets them. */
#define QIW_PROFILE_CLASSES 8
int qiw_profile_enable(qiw_context* ctx, int32_t on);
int qiw_profile_read(qiw_context* ctx, double* ms, int64_t* launches, int32_t reset);

/* ---- whole inchworm run on the device (replaces the loop of inchworm!, src/inchworm.jl:474-493) -- */

/* Runs the bare step (entries bare_ids at grid[0] -> grid[1]) followed by the bold steps
 * n = 2 .. n_tau-1 (entries bold_ids), with set_ppgf! and normalize!(P, tau_f)
 * (src/ppgf.jl:495-504,646-668) executed on the device between steps, so no host round trip
 * separates the n_tau-1 sequential steps.  The P table must hold the initial (atomic) values.
 * On return the device P table is final; read it with qiw_get_P.
 *   order_contribs : optional [n_tau][n_bare+n_bold entries] packed block vectors: the
 *                    per-entry contribution to P(tau_k) *before* later normalisations (what
 *                    inchworm! accumulates into P_orders, src/inchworm.jl:489-492); may be NULL.
 *   sobol_m/x0     : per entry (bare entries first, then bold), as in qiw_eval; NULL = unscrambled.
 *                    The same sequence is reused at every step (RandomizationParams() default,
 *                    src/randomization.jl:57). */
int qiw_inchworm_run(qiw_context* ctx, int32_t n_bare, const int32_t* bare_ids, int32_t n_bold,
                     const int32_t* bold_ids, const uint32_t* sobol_m, const uint32_t* sobol_x0,
                     uint64_t N_total, double* order_contribs);

/* ---- Sobol sequence (src/scrambled_sobol.jl) -------------------------------------------------- */

/* Host: direction numbers m[D][32] of the unscrambled sequence (ScrambledSobolSeq(D), :66-100). */
int qiw_sobol_direction_numbers(int32_t D, uint32_t* m);
/* Host: LMS + shift scrambling from explicit random bits (:104-139): shift_bits[D][32] and
 * ltm_bits[D][32][32] in Julia (column-major) memory order, i.e. exactly what
 * rand(rng, UInt32[0,1], D, 32) and rand(rng, UInt32[0,1], D, 32, 32) returned. */
int qiw_sobol_scramble(int32_t D, uint32_t* m, uint32_t* x0, const uint8_t* shift_bits,
                       const uint8_t* ltm_bits);
/* Device: points start .. start+count-1 (0-based; point 0 is x0) by Gray-code random access,
 * equivalent to skip!(seq, start, exact=true) followed by `count` next! calls (:158-197).
 * points_u32[count][D] raw integers; value = integer * 2^-32. */
int qiw_sobol_points(qiw_context* ctx, int32_t D, const uint32_t* m, const uint32_t* x0,
                     uint64_t start, uint64_t count, uint32_t* points_u32);

/* ---- topologies (src/diagrammatics.jl), host ----------------------------------------------------- */

/* get_topologies_at_order(order, k; with_external_arc) (:322-337); k < 0 = `nothing`.
 * Returns the count through n_top; fills pairs[n_top][order][2] / parity[n_top] when non-NULL
 * (call once with NULL to size the buffers).  Same enumeration order and parity as the reference. */
int qiw_topologies(int32_t order, int32_t k, int32_t with_external_arc, int64_t* n_top,
                   int32_t* pairs, int32_t* parity);

/* ---- sample partitioning and multi-GPU (src/mpi.jl) ------------------------------------------------ */

/* rank_sub_range (src/mpi.jl:49-54) / split_count (src/utility.jl:164-167), 0-based start. */
int qiw_rank_sub_range(uint64_t N, int32_t n_ranks, int32_t rank, uint64_t* start, uint64_t* count);

#define QIW_UNIQUE_ID_BYTES 128
/* One process per GPU.  Rank 0 calls qiw_comm_unique_id and ships the 128 bytes to the other
 * ranks by any host-side means (MPI.bcast in Julia, torch.distributed here); every rank then
 * calls qiw_comm_init.  Afterwards qiw_eval / qiw_inchworm_run shard Sobol indices by
 * rank_sub_range and combine the packed block sums with ONE ncclAllReduce per call / per inchworm
 * step (replaces all_reduce!, src/mpi.jl:104-127). */
int qiw_comm_unique_id(uint8_t id[QIW_UNIQUE_ID_BYTES]);
int qiw_comm_init(qiw_context* ctx, int32_t n_ranks, int32_t rank, const uint8_t id[QIW_UNIQUE_ID_BYTES]);
int qiw_comm_destroy(qiw_context* ctx);

#define QIW_PEER_HANDLE_BYTES 64
/* Peer-memory all-reduce (preferred when the GPUs of the job see each other over NVLink/NVSwitch):
 * the reduction of the per-GPU block sums is performed INSIDE the step kernel's tail — every rank
 * stores its sums into all peers' mailboxes, raises a flag and adds up what the peers stored, in rank
 * order — so an inchworm step stays a single kernel launch on every GPU.
 * Every rank calls qiw_peer_handle (allocates its mailbox, returns the CUDA IPC handle), the host
 * all-gathers the 64-byte handles (MPI.Allgather in Julia, torch.distributed here) and passes all of
 * them, in rank order, to qiw_peer_init.  Payloads larger than 64 KiB per call and block models fall
 * back to the NCCL communicator of qiw_comm_init (which must then exist).  If any rank fails to map the
 * mailboxes, ALL ranks must call qiw_peer_init(ctx, 0, 0, NULL) to switch the peer path off again. */
int qiw_peer_handle(qiw_context* ctx, uint8_t handle[QIW_PEER_HANDLE_BYTES]);
int qiw_peer_init(qiw_context* ctx, int32_t n_ranks, int32_t rank, const uint8_t* handles);

/* ---- measurement helpers ---------------------------------------------------------------------------- */

/* Runs a DFMA-saturating kernel and reports the measured FP64 FMA throughput (TFLOP/s, 2 flops
 * per DFMA): the roofline denominator for this path (MEASURED_PEAKS.json has no FP64 entry). */
int qiw_measure_fp64_peak(qiw_context* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* QINCHWORM_H */
