// qiw_host.hpp — host-side data structures shared by the compiler (qiw_compile.cpp) and the
// runtime (qiw_api.cu) of libqinchworm_cuda.so.
#pragma once

#include <complex>
#include <cstdint>
#include <string>
#include <vector>

namespace qiw {

typedef std::complex<double> cplx;

constexpr int kMaxNodes = 19;    // 2*order + 3 with order <= 8
constexpr int kMaxOrder = 8;

// Flattened `Expansion` (reference: src/expansion.jl:110-134).
struct HostModel {
    int S = 0;
    std::vector<int> dim, boff, eoff;
    int bsize = 0, maxdim = 0;
    bool scalar = false;                  // every sector block is 1x1
    std::vector<double> energies;         // E + lambda, concatenated
    int n_ops = 0;
    std::vector<int> op_target;           // [n_ops][S]
    std::vector<int64_t> op_off;          // [n_ops][S] into pool
    std::vector<cplx> pool;
    std::vector<int> pair_op_i, pair_op_f, pair_table;
    std::vector<std::vector<int>> attachable;  // src/expansion.jl:180-183
    std::vector<int> corr_A, corr_B;
    int target(int op, int s) const { return op_target[(size_t)op * S + s]; }
    const cplx* block(int op, int s) const { return pool.data() + op_off[(size_t)op * S + s]; }
};

// Where the time of a backbone position comes from.
enum : int { SRC_TI = -1, SRC_TW = -2, SRC_TF = -3 };

struct DeltaSlot { int pos_tail, pos_head, table; };

// Program word layout (64 bit), one word per node of the pruned configuration tree, pre-order.
//   [ 0,12) slotA : index of the propagator factor in the per-sample table
//                   scalar program: (level-2)*S + sector ; block program: source sector
//   [12,24) slotB : absolute table index of the pair-interaction factor (0 = none)
//   [24,32) nchild: number of children (0 = this node is the last position -> leaf)
//   [32,48) aux   : leaf: coefficient index; root: initial sector
//   [48,60) op    : block program only: operator index + 1 (0 = identity)
inline uint64_t make_word(uint32_t slotA, uint32_t slotB, uint32_t nchild, uint32_t aux, uint32_t op = 0) {
    return (uint64_t)(slotA & 0xFFF) | ((uint64_t)(slotB & 0xFFF) << 12) | ((uint64_t)(nchild & 0xFF) << 24) |
           ((uint64_t)(aux & 0xFFFF) << 32) | ((uint64_t)(op & 0xFFF) << 48);
}

struct EntryProgram {
    int mode = 0, order = 0, n_pts_after = 0, corr_idx = 0;
    int n_nodes = 0, D = 0;
    bool scalar = false;
    int pos_src[kMaxNodes + 1];            // 1-based position -> time source
    int fixed_op[kMaxNodes + 1];           // operator index of fixed operator nodes, else -1
    std::vector<uint64_t> words;           // all trees, pre-order, roots included, +1 pad word
    std::vector<uint32_t> tree_off;        // [n_trees + 1] word offsets
    std::vector<uint32_t> tree_cost;       // edges per tree (load balancing)
    std::vector<cplx> coefs;               // distinct leaf coefficients
    std::vector<DeltaSlot> dslots;         // pair-interaction factors referenced by the program
    int nP = 0;                            // number of propagator slots ((n_nodes-1) * S)
    // Leaf records (scalar models): one fixed-length record per surviving configuration, in the
    // tree's leaf order.  rec[0] = coefficient index | initial sector << 16; rec[1..L] = table slot
    // of every factor of the configuration's weight in the per-sample table; padded
    // with zeros to RL words (RL % 4 == 0).  L = (n_nodes - 1) propagators + `order` interactions.
    std::vector<uint32_t> records;
    int L = 0, RL = 0;
    // Factorised records (scalar models; what the step kernel executes).  The (n_nodes - 1)
    // propagator factors of a configuration are grouped into K contiguous segments of backbone
    // intervals; every distinct sector sub-sequence of a segment becomes one entry of a per-sample
    // table of segment products, so a configuration needs only K + order operands:
    //   rec2[0] = coefficient index | initial sector << 16
    //   rec2[1..K]       = slot of the segment product  (nP + nD + entry index)
    //   rec2[K+1..K+order] = slot of the pair-interaction factor (nP + dslot index)
    // segdef[nSeg][seg_stride]: propagator slots multiplied into each segment product (0xFFFF = unused).
    std::vector<uint32_t> rec2;
    // Lane program (scalar models; what the step kernel executes with lane = sample).  Configurations that share
    // all pair-interaction operands and the initial sector differ only in the flavours running along the backbone,
    // i.e. in their segment products; they form a GROUP:
    //     sum over the group = prod(Delta operands) * sum_members prod(segment products of the member).
    // The coefficient of a member is folded into its first segment product (a table entry of its own per distinct
    // (coefficient, segment) combination: lane_segdef / lane_seg_coef), so a member is a plain product of table slots.
    // Groups are cut into fixed-shape records of M = 4, 2 or 1 members and sorted into sections of equal
    // (M, initial sector); a record is `order` Delta slots followed by M * K segment slots, 16 bits each (plain slot
    // numbers of the per-sample table), padded to a multiple of 8 items (one warp-uniform 128-bit load per 8 items:
    // the record stream shares the load pipe with the operand loads, so its width matters).
    // s_b < 0: all M members of a record belong to initial sector s_i; else the first M / 2 to s_i, the others to s_b
    struct LaneSection { int32_t s_i, s_b, M; uint32_t rec0, n_rec; uint32_t chunk0, cost; };   // chunk0: first 128-bit word; cost per record
    std::vector<LaneSection> lane_sections;
    std::vector<uint16_t> lane_items;
    std::vector<uint16_t> lane_segdef;     // [nSegL][seg_stride] propagator slots (0xFFFF = unused)
    std::vector<uint16_t> lane_seg_coef;   // [nSegL] index of the folded coefficient, 0xFFFF = none
    int nSegL = 0;
    int64_t lane_cost = 0;                 // sum over sections of n_rec * cost (operand loads per sample)
    std::vector<uint16_t> segdef;
    int K = 0, L2 = 0, nSeg = 0, seg_stride = 0;
    // statistics
    int64_t n_top = 0, n_leaves = 0, n_edges = 0;
    double flops_per_sample = 0;
};

// Compile one TopologiesInputData against the model.  Returns 0 or a qiw_status.
int compile_entry(const HostModel& m, int mode, int order, int n_pts_after, int corr_idx, int n_top,
                  const int32_t* pairs, const int32_t* parity, EntryProgram& out, std::string& err);
// Regroup the flat leaf records into factorised records with `K` segments (0 = choose).
void factorise_records(EntryProgram& e, int S, int K);
// Group the factorised records into the lane program (EntryProgram::lane_*).
void build_lane_program(EntryProgram& e);
inline int lane_record_items(int order, int K, int M) { return ((order + M * K + 7) / 8) * 8; }

// Host Sobol / topology helpers (qiw_seq.cpp)
int sobol_direction_numbers(int D, uint32_t* m);
int sobol_scramble(int D, uint32_t* m, uint32_t* x0, const uint8_t* shift_bits, const uint8_t* ltm_bits);
int64_t enumerate_topologies(int order, int k, bool external_arc, int32_t* pairs, int32_t* parity);
void rank_sub_range(uint64_t N, int n_ranks, int rank, uint64_t* start, uint64_t* count);

// Natural cubic spline second derivatives for knots j*h.
void natural_spline_second_derivatives(int n, double h, const cplx* y, cplx* M);

}  // namespace qiw
