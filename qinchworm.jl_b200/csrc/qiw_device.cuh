// qiw_device.cuh — device-side structures and kernels' declarations of libqinchworm_cuda.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace qiw {

constexpr int kDevMaxNodes = 19;
constexpr int kWalkMaxSp = 10;     // branch-point stack frames of the block walker (order + 2 <= 10)
constexpr int kWalkPrefetch = 24;   // words the block walker prefetches ahead of its program counter
constexpr int kDevMaxDim = 16;      // Sobol dimensions handled per entry (2 * order <= 16)
constexpr int kMaxTables = 64;
constexpr int kInlineTables = 8;     // propagator tables described directly in the kernel parameters

// One scalar propagator table resident in HBM.
struct DevDelta {
    const double2* y;   // grid values
    const double2* M;   // spline second derivatives (kind == 1)
    int kind, n;
    double h;           // beta / (n - 1)
    double inv_h;
};

// Static (per compiled entry) description read by every CTA working on the entry.
struct DevEntry {
    int mode, order, n_nodes, D;
    int d_after, d_before;
    int nP, nD;                 // table slots: propagators, pair interactions
    int exact;                  // order 0: one deterministic evaluation
    int pos_src[kDevMaxNodes + 1];
    int n_coefs;
    // lane program (EntryProgram::lane_*): what the step kernel executes with lane = sample
    int K, nSegL, seg_stride;   // segments per configuration; entries / propagator slots per entry of the segment table
    const uint4* lane_items;    // records: `order` Delta slots then M * K segment slots, 16 bits each, padded to 8
    const uint4* lane_segdef4;  // [nSegL][seg_stride > 7 ? 2 : 1] packed 16-bit fields: index of the coefficient folded into
                                // the product (0xFFFF = none), then its propagator slots (0xFFFF = unused)
    const double2* coefs;
    const int4* dslots;         // (pos_tail, pos_head, table, 0)
};

// Per call, per entry.
struct DevEntryDyn {
    const uint32_t* sobol;      // m[D][32] followed by x0[D]
    double* ucache;             // cached simplex roots [D][count] of this entry, or null
    int ucache_valid, pad_;
    double weight;              // 1/N_total (sampled entries); 1 or 0 (exact entries, by rank)
    unsigned long long start;   // first Sobol index evaluated by this rank
    unsigned long long count;   // number of Sobol points evaluated by this rank (1 if exact)
    int entry;                  // id of the compiled entry
    int out_index;              // row of the output
    int item0, n_items;         // this entry's CTA jobs (consecutive rows of the partials buffer)
};

// One CTA job: up to `warps` consecutive chunks of one entry's lane program (one chunk per warp).
struct WorkItem {
    int entry;        // id of the compiled entry (index into the DevEntry table)
    int slot;         // index of the entry within the call (DevEntryDyn)
    int chunk0;       // first chunk of the job: index into StepParams::chunk_off
    int n_chunks;     // chunks in the job (<= warps per CTA)
    int n_chunks_total;  // block models: chunks the entry's trees are split into
    int partial0;     // first row of this item in the partials buffer
};

// A run of consecutive records of one section (same shape, same initial sector) inside a chunk.
//   x = first 128-bit word of the run in DevEntry::lane_items, y = number of records,
//   z = initial sector | (second sector + 1) << 16 (0: none; else the first M / 2 members of a record belong to the
//       initial sector, the others to the second one), w = shape code: order * 16 + (K - 1) * 4 + (0, 1, 2 for M = 1, 2, 4)
typedef uint4 LaneRun;

struct StepParams {
    const DevEntry* entries;
    const DevEntryDyn* dyn;
    const WorkItem* items;
    const uint32_t* chunk_off;     // scalar models: [chunks + 1] first run of every chunk
    const LaneRun* runs;
    // model
    const double2* P;              // [n_tau][bsize]
    const double* E;               // [S] (scalar models) energies + lambda
    const DevDelta* deltas;
    DevDelta deltas_inline[kInlineTables];
    int max_dslots;
    int S, bsize, n_tau;
    double h, inv_h;               // beta / (n_tau - 1) and its reciprocal
    // times; when `times_dev` is non-null the triple is read from device memory (run-level API)
    double t_i, t_w, t_f;
    const double* times_dev;
    int max_slots;                 // operands per sample reserved in shared memory
    int aux_off, red_off;          // shared-memory offsets: times / grid cells (dead once the propagator and interaction
                                   // slots are filled: they overlap the segment-product rows of T), per-warp sums
    int max_nodes1;                // largest number of backbone positions of the launch's entries, plus one
    int spb, spb_log2;             // samples per CTA pass (power of two <= 32)
    int sobol_z_stride;            // > 0: blockIdx.z selects a Sobol sequence (words between the sequences' parameter blocks)
    int tables_on_grid;            // every Delta table is a plain grid function on the P grid and described inline
    int allow_overlap;             // launch with programmatic stream serialization (QIW_PDL=1 enables; off by default)
    // explicit-times mode (qiw_eval_at_times): times[count][D], per-sample output, no reduction
    const double* explicit_times;
    double2* per_sample_out;       // [count][S]
    // output
    double2* partials;             // [n partial rows][gridDim.x][S]
    unsigned long long* trace;     // diagnostics: 12 words per CTA, or null
    // fused tail: the last CTA to finish reduces the per-CTA partial sums of every entry of the call
    // (deterministic order) and, in the device-resident loop, applies set_ppgf! + normalize!
    unsigned int* done_counter;    // null = no fused tail (separate reduction kernel)
    int n_call_entries;            // entries of this call (rows of `dyn`)
    double2* out;                  // [n_call_entries][S] reduced results
    int finish_k_f;                // >= 0: write P(tau_k_f) = sum of the entries' results ...
    int finish_normalize;          // ... and rescale the table (src/ppgf.jl:646-668)
    double2* finish_P;             // the P table to update (== P)
    const int* finish_diag;
    int finish_n_diag;
    double2* finish_hist;          // optional per-entry contributions of this step
    // peer-memory all-reduce inside the fused tail (one process per GPU, mailboxes mapped with CUDA IPC):
    // every rank stores its reduced block sums into all peers' mailboxes over NVLink, raises a flag,
    // waits for the peers' flags and sums the contributions in rank order (identical on all ranks)
    int peer_ranks, peer_rank;     // peer_ranks <= 1: no exchange
    unsigned long long peer_seq;   // sequence number of this collective (flags carry it; parity selects the buffer)
    unsigned char* const* peer_mail;   // [peer_ranks] base address of every rank's mailbox (own included)
    int* peer_status;              // set to 1 if a peer did not answer within the time-out
    unsigned long long peer_timeout_ns;   // how long a rank waits for its peers (QIW_PEER_TIMEOUT_S, default 60 s)
};

// One CTA of the persistent run kernel: a work item (entry + chunks of its lane program) on 32 * n_sub consecutive
// samples starting at sample block sb0, for every step of the run.
struct RunJob { int item, sb0, n_sub, aux_off, row, stash_off, flags, pad_; };   // row: this job's row of the partial sums;
                                                                              // stash_off: on-chip copy of its program, or -1;
                                                                              // flags & 1: the staged partial rows overlap neither its
                                                                              // pair-interaction rows nor its times (next step's may be prepared early)

struct RunParams {
    StepParams sp;                 // tables, entries, work items, runs; finish_P = the global P table; peer_* as in a step
    const RunJob* jobs;            // [n_jobs], grouped by CTA
    const int* cta_job0;           // [gridDim.x + 1] first job of every CTA (jobs are packed at equal estimated time)
    const int* entry_job0;         // [n_call_entries + 1] first partial row of every entry of the call (rows of an entry are consecutive)
    int n_jobs;
    int k_first, n_steps;          // step i: tau_w = grid[k_first + i], tau_f = grid[k_first + i + 1]
    void* partials;                // [2][n_jobs][S] per-job sums (kernel arithmetic), alternating with the step's parity
    unsigned int* barrier;         // arrival counter of the grid barrier, zero at launch
    unsigned int* sm_map;          // null, or [1024] CTAs arrived per SM, [1024] bin claimed by the SM (+1), [1] bins claimed; zero at launch
    int post_warps;                // warps of a CTA that reduce / exchange / update P after the barrier while the others prepare the next step
    int ctas_per_sm;               // job lists per bin (cta_job0 is then indexed by bin * ctas_per_sm + arrival order on the SM)
    int n_tables;                  // pair-interaction tables staged in shared memory
    int D_table_off[kInlineTables];   // first element of every table in the staged array (used unless sp.tables_on_grid)
    int rows_staged;               // the partial rows of a step fit the operand table's space: reduce them from shared memory
    int ok_off, pw_off, red_off, ds_off, P_off, D_off, out_off;   // shared-memory layout (bytes)
    double2* hist;                 // optional per-entry contributions: hist[k_f * hist_stride + hist_off + entry * S + s]
    size_t hist_stride, hist_off;
};

// Mailbox layout (per rank): flags[kMaxPeers][2] (uint64), then data[kMaxPeers][2][kPeerSlotBytes].
constexpr int kMaxPeers = 16;
constexpr size_t kPeerSlotBytes = 64 * 1024;
constexpr size_t kPeerFlagBytes = (size_t)kMaxPeers * 2 * sizeof(unsigned long long);
constexpr size_t kPeerMailBytes = kPeerFlagBytes + (size_t)kMaxPeers * 2 * kPeerSlotBytes;

}  // namespace qiw

namespace qiw {

// Model tables for sector blocks larger than 1x1 (block kernel).
struct DevModel {
    const int* dim;          // [S]
    const int* boff;         // [S] offsets into the packed block vector
    const int* eoff;         // [S] offsets into E
    const int* op_target;    // [n_ops][S]
    const long long* op_off; // [n_ops][S] into pool
    const double2* pool;     // operator blocks, column-major
    int S, bsize, maxdim, n_ops;
};

// Real-arithmetic tree replay for block models (block_walk_kernel).
struct BlockWalkParams {
    const double* pool_re;        // operator blocks, real parts, same offsets as DevModel::pool
    int pool_n;                   // > 0: number of doubles of the pool, staged in shared memory by the walker
    const uint4* const* xwords;   // per compiled entry id: expanded program words (layout: qiw_kernels.cu)
    const uint32_t* const* unit_off;   // per compiled entry id: [n_units + 1] word offsets of the walk units
    const int* chunk_bounds;      // [n_items][warps + 1] unit ranges of every warp of every CTA job
    int warps, max_sp;            // warps per CTA; stack frames reserved per warp
    int nI_max, nD_max;           // table sizes reserved in shared memory
};

struct BlockParams {
    DevModel m;
    const uint64_t* const* words;      // per compiled entry id: tree word stream
    const uint32_t* const* tree_off;   // per compiled entry id: [n_trees + 1]
    const int* n_trees;                // per compiled entry id
    double2* scratch;                  // per thread: propagator blocks, interaction factors
    size_t scratch_per_thread;         // in double2
};

}  // namespace qiw
