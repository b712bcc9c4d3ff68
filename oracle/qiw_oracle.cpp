// qiw_oracle.cpp — CPU ORACLE (test infrastructure, NOT the product).
//
// A plain, sequential C++ restatement of the reference's quasi Monte Carlo diagram-evaluation
// hot path (krivenko/QInchworm.jl v0.5.0).  It exists to *check* the CUDA library and to serve as
// the CPU baseline in bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it; the product (libqinchworm_cuda.so and the qinchworm_b200
// host layer) never does.
//
// Parity status: PINNED.  tests/test_oracle_golden.py checks this file against the reference's
// own golden vectors (tests/golden/*.json extracted from test/inchworm.h5, test/topology_eval.h5,
// test/scrambled_sobol.jl, README.md topology counts).  Blocks with d_s > 1 are unpinned at the
// reference level (no golden exists); they are self-validated by basis-rotation invariance and a
// brute-force Fock-space evaluation in tests/.
//
// Every function cites the reference file:line it restates (paths relative to /root/reference).
// Third-party arithmetic that is not vendored in the reference (Keldysh.jl grid interpolation,
// Interpolations.jl natural cubic spline, Sobol.jl Joe-Kuo table) is restated from its published
// algorithm and pinned through the reference's golden files.

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "joe_kuo_table.h"

typedef std::complex<double> cplx;
static const cplx I_(0.0, 1.0);

// ------------------------------------------------------------------------------------------------
// Sobol sequence (src/scrambled_sobol.jl)
// ------------------------------------------------------------------------------------------------

// Direction numbers m[d][j], j = 0..31, pre-shifted so that bit 31 is the most significant
// binary digit.  src/scrambled_sobol.jl:69-100.
static int sobol_direction_numbers(int D, uint32_t* m /* [D][32] */) {
    if (D < 0 || D > QIW_JK_NDIM) return 1;
    for (int d = 0; d < D; ++d)
        for (int j = 0; j < 32; ++j) m[d * 32 + j] = 1u;  // :69
    for (int d = 1; d < D; ++d) {                          // :78 (d is 0-based here)
        uint32_t a = qiw_jk_a[d - 1];                      // :79
        int deg = 31 - __builtin_clz(a);                   // :80 floor(log2(a))
        uint32_t* md = m + d * 32;
        for (int j = 0; j < deg; ++j) md[j] = qiw_jk_minit[d - 1][j];  // :83
        for (int j = deg; j < 32; ++j) {                   // :85
            uint32_t ac = a;
            md[j] = md[j - deg];
            for (int k = 0; k < deg; ++k) {                // :88-92
                md[j] ^= ((ac & 1u) * md[j - deg + k]) << (deg - k);
                ac >>= 1;
            }
        }
    }
    for (int d = 0; d < D; ++d)
        for (int j = 0; j < 32; ++j) m[d * 32 + j] <<= (31 - j);  // :98-100
    return 0;
}

// LMS + digital shift scrambling from explicit random bits.
//   shift_bits: what `rand(rng, UInt32[0,1], D, 32)` returned, column-major (D fastest)   (:106)
//   ltm_bits  : what `rand(rng, UInt32[0,1], D, 32, 32)` returned, column-major           (:110)
// src/scrambled_sobol.jl:104-139.
static void sobol_scramble(int D, uint32_t* m, uint32_t* x0, const uint8_t* shift_bits,
                           const uint8_t* ltm_bits) {
    for (int d = 0; d < D; ++d) {
        uint32_t x = 0;
        for (int j = 0; j < 32; ++j) x += (uint32_t)shift_bits[d + D * j] << j;  // :106
        x0[d] = x;
    }
    for (int d = 0; d < D; ++d) {
        // Lower-triangular 32x32 bit matrix with unit diagonal (:109-118); ltm[d, p, q] lives at
        // ltm_bits[d + D*(p + 32*q)], p, q 0-based.
        uint32_t row[32];  // row[p] = sum_q ltm[p][q] * 2^(31-q)  (:127)
        for (int p = 0; p < 32; ++p) {
            uint32_t r = 0;
            for (int q = 0; q < 32; ++q) {
                uint32_t bit = (q < p) ? ltm_bits[d + D * (p + 32 * q)] : (q == p ? 1u : 0u);
                r |= (bit & 1u) << (31 - q);
            }
            row[p] = r;
        }
        for (int j = 0; j < 32; ++j) {  // :122-138
            uint32_t mdj = m[d * 32 + j], t2 = 0, l = 1;
            for (int p = 31; p >= 0; --p) {
                uint32_t t1 = (uint32_t)__builtin_popcount(row[p] & mdj) & 1u;
                t2 += t1 * l;
                l <<= 1;
            }
            m[d * 32 + j] = t2;
        }
    }
}

// Sequential generator state and next!()  (src/scrambled_sobol.jl:50-57,158-173)
struct SobolSeq {
    int D;
    std::vector<uint32_t> m, x;
    uint32_t n;
    void next(double* out) {
        if (n == 0) {
            n += 1;
        } else {
            n += 1;
            int c = __builtin_ctz(n - 1);  // :168
            for (int d = 0; d < D; ++d) x[d] ^= m[d * 32 + c];
        }
        for (int d = 0; d < D; ++d) out[d] = std::ldexp((double)x[d], -32);  // :163,170
    }
    void skip_exact(uint64_t k) {  // skip!(…, exact=true) :189-197
        std::vector<double> buf(D);
        for (uint64_t i = 0; i < k; ++i) next(buf.data());
    }
};

// ------------------------------------------------------------------------------------------------
// Topologies (src/diagrammatics.jl)
// ------------------------------------------------------------------------------------------------

struct Topology {
    std::vector<int> a, b;  // arc k connects vertices a[k] => b[k] (1-based, a < b)
    int parity;
};

// generate_topologies_impl, src/diagrammatics.jl:288-302
static void gen_topologies_impl(Topology& partial, std::vector<int>& unpaired,
                                std::vector<Topology>& out) {
    if (unpaired.empty()) { out.push_back(partial); return; }
    for (size_t i = 1; i < unpaired.size(); ++i) {  // Julia i = 2:length  → (i+1) is 1-based
        int va = unpaired[0], vb = unpaired[i];
        std::vector<int> rest;
        for (size_t k = 1; k < unpaired.size(); ++k) if (k != i) rest.push_back(unpaired[k]);
        int save = partial.parity;
        partial.parity = save * (((i + 1) % 2 == 0) ? 1 : -1);  // (-1)^i, 1-based i  (:294)
        partial.a.push_back(va); partial.b.push_back(vb);
        gen_topologies_impl(partial, rest, out);
        partial.a.pop_back(); partial.b.pop_back();
        partial.parity = save;
    }
}

static bool arcs_cross(int a1, int b1, int a2, int b2) {  // iscrossing :109-115
    if (a1 > b1) std::swap(a1, b1);
    if (a2 > b2) std::swap(a2, b2);
    if (a1 < a2 && a2 < b1 && b1 < b2) return true;
    if (a2 < a1 && a1 < b2 && b2 < b1) return true;
    return false;
}

// is_doubly_k_connected(t, k), src/diagrammatics.jl:231-235 with :148-159,197-213
static bool is_doubly_k_connected(const Topology& t, int k) {
    int n = (int)t.a.size();
    std::vector<int> state(n, 0);  // 0 = disconnected, 1 = connected
    std::vector<int> stack;
    for (int i = 0; i < n; ++i) {
        bool dk = (t.a[i] <= k && t.b[i] > k) || (t.b[i] <= k && t.a[i] > k);
        if (dk) { state[i] = 1; stack.push_back(i); }
    }
    while (!stack.empty()) {
        int c = stack.back(); stack.pop_back();
        for (int i = 0; i < n; ++i)
            if (!state[i] && arcs_cross(t.a[c], t.b[c], t.a[i], t.b[i])) {
                state[i] = 1; stack.push_back(i);
            }
    }
    for (int i = 0; i < n; ++i) if (!state[i]) return false;
    return true;
}

// get_topologies_at_order, src/diagrammatics.jl:322-337.  k < 0 means `nothing`.
static std::vector<Topology> get_topologies_at_order(int order, int k, bool with_external_arc) {
    std::vector<Topology> all, out;
    Topology empty; empty.parity = 1;
    std::vector<int> verts(2 * order);
    for (int i = 0; i < 2 * order; ++i) verts[i] = i + 1;
    gen_topologies_impl(empty, verts, all);
    if (k < 0) return all;
    for (auto& t : all) if (is_doubly_k_connected(t, k)) out.push_back(t);
    if (with_external_arc)
        for (auto& t : out) t.parity *= (k % 2 == 0) ? 1 : -1;  // (-1)^k  (:333)
    return out;
}

// ------------------------------------------------------------------------------------------------
// Model tables (flattened src/expansion.jl Expansion)
// ------------------------------------------------------------------------------------------------

enum { QO_DELTA_GRID = 0, QO_DELTA_SPLINE = 1 };
enum { QO_MODE_BARE = 0, QO_MODE_BOLD = 1, QO_MODE_CORR = 2 };

struct DeltaTable {
    int kind = 0, n = 0;
    double beta = 0;
    std::vector<cplx> y;   // grid values
    std::vector<cplx> M;   // natural-spline second derivatives (kind == SPLINE)
};

struct Operator {
    std::vector<int> target;             // per source sector: target sector or -1
    std::vector<std::vector<cplx>> mat;  // per source sector: d_target x d_source, column-major
};

struct Entry {
    int mode = 0, order = 0, n_pts_after = 0;
    std::vector<Topology> tops;
};

struct Oracle {
    int S = 0;
    std::vector<int> dim, boff;  // block offsets into the packed sum_s d_s^2 layout
    int bsize = 0;               // sum_s d_s^2
    std::vector<std::vector<double>> E;  // E_s[i] + lambda0 (src/exact_atomic_ppgf.jl:133)
    std::vector<Operator> ops;
    std::vector<int> pair_op_i, pair_op_f, pair_table;
    std::vector<std::vector<int>> attachable;  // src/expansion.jl:180-183
    std::vector<int> corr_A, corr_B;
    std::vector<DeltaTable> deltas;
    int n_tau = 0;
    double beta = 0;
    std::vector<cplx> P;  // [k][bsize]
    std::vector<Entry> entries;
    int n_threads = 1;
    // host threads: a persistent pool whose workers keep one evaluator per entry across calls (the moral
    // equivalent of persistent MPI ranks, src/mpi.jl:49-54); use_pool = 0 restores one std::thread per
    // (entry, call) with a fresh evaluator each, as the reference rebuilds its TopologyEvaluator every step
    int use_pool = 1;
    uint64_t cache_epoch = 0;      // bumped whenever the model or an entry changes: cached evaluators are stale
    struct Pool* pool = nullptr;
    std::string err;
    // flop counter of the last evaluation (8*m*k*n per live edge + 8*d_i*d_f per leaf)
    double last_flops_per_sample = 0, last_leaves = 0, last_edges = 0;
};

// ------------------------------------------------------------------------------------------------
// Interpolation
// ------------------------------------------------------------------------------------------------

// Keldysh.jl generic grid interpolation of a time-translation-invariant imaginary-time function
// stored as D[k] = G(k h) (third party; rule stated in SURVEY.md §8a a11, pinned by
// test/topology_eval.h5 and test/inchworm.h5).  Call sites: src/topology_eval.jl:368,414.
// `stride` lets the same routine read one matrix element out of the packed P table.
static inline cplx grid_interp(const cplx* D, int stride, int n, double h, double t_f, double t_i) {
    int a = (int)std::floor(t_f / h), b = (int)std::floor(t_i / h);
    if (a > n - 2) a = n - 2;
    if (b > n - 2) b = n - 2;
    if (a < 0) a = 0;
    if (b < 0) b = 0;
    double w1 = t_f / h - a, w2 = t_i / h - b;
    if (a == b) {
        cplx d0 = D[0], d1 = D[stride];
        return d0 + (w1 - w2) * (d1 - d0);
    }
    int k = a - b;
    cplx dk = D[(size_t)k * stride], dkp = D[(size_t)(k + 1) * stride], dkm = D[(size_t)(k - 1) * stride];
    return (1 - w1) * (1 - w2) * dk + w1 * (1 - w2) * dkp + (1 - w1) * w2 * dkm + w1 * w2 * dk;
}

// Natural cubic spline through (j h, y_j): second derivatives by the Thomas algorithm.
// Restates Interpolations.jl BSpline(Cubic(Line(OnGrid()))) used at src/spline_gf.jl:193-201.
static void spline_prepare(DeltaTable& t) {
    int n = t.n;
    double h = t.beta / (n - 1);
    t.M.assign(n, cplx(0));
    if (n < 3) return;
    int m = n - 2;
    std::vector<double> c(m);
    std::vector<cplx> d(m);
    for (int j = 0; j < m; ++j) d[j] = 6.0 * (t.y[j] - 2.0 * t.y[j + 1] + t.y[j + 2]) / (h * h);
    // tridiagonal (1, 4, 1)
    c[0] = 1.0 / 4.0; d[0] = d[0] / 4.0;
    for (int j = 1; j < m; ++j) {
        double den = 4.0 - c[j - 1];
        c[j] = 1.0 / den;
        d[j] = (d[j] - d[j - 1]) / den;
    }
    for (int j = m - 2; j >= 0; --j) d[j] -= c[j] * d[j + 1];
    for (int j = 0; j < m; ++j) t.M[j + 1] = d[j];
}

// interpolate(G_int, t1, t2) for t1 >= t2: spline evaluated at the time difference,
// src/spline_gf.jl:208-219.
static inline cplx spline_eval(const DeltaTable& t, double dt) {
    int n = t.n;
    double h = t.beta / (n - 1);
    int j = (int)std::floor(dt / h);
    if (j > n - 2) j = n - 2;
    if (j < 0) j = 0;
    double xa = dt - j * h, xb = (j + 1) * h - dt;
    return t.M[j] * (xb * xb * xb) / (6 * h) + t.M[j + 1] * (xa * xa * xa) / (6 * h) +
           (t.y[j] / h - t.M[j] * h / 6.0) * xb + (t.y[j + 1] / h - t.M[j + 1] * h / 6.0) * xa;
}

static inline cplx delta_eval(const DeltaTable& t, double t_f, double t_i) {
    if (t.kind == QO_DELTA_SPLINE) return spline_eval(t, t_f - t_i);
    return grid_interp(t.y.data(), 1, t.n, t.beta / (t.n - 1), t_f, t_i);
}

// ------------------------------------------------------------------------------------------------
// TopologyEvaluator (src/topology_eval.jl)
// ------------------------------------------------------------------------------------------------

enum NodeKind { PAIR = 1, IDENT = 2, INCH = 3, OPER = 4 };  // :56

struct Node { int kind, arc_index, operator_index; };

// LazyMatrixProduct: A_N ... A_1 with cached partial products (src/utility.jl:234-323)
struct LazyProduct {
    struct Mat { int r, c; const cplx* ptr; };
    std::vector<Mat> mats;
    std::vector<std::vector<cplx>> partial;
    std::vector<std::pair<int, int>> pdim;
    int n_mats = 0, n_prods = 0;
    double* flops;
    void init(int maxn, int maxd) {
        mats.resize(maxn); partial.assign(maxn, std::vector<cplx>(maxd * maxd)); pdim.resize(maxn);
        n_mats = n_prods = 0;
    }
    void push(int r, int c, const cplx* p) { mats[n_mats++] = {r, c, p}; }  // pushfirst! :274
    void pop() { --n_mats; if (n_prods > n_mats) n_prods = n_mats; }         // popfirst! :288
    // eval! :301-323
    const cplx* eval(int& r, int& c) {
        if (n_prods == 0) {
            std::copy(mats[0].ptr, mats[0].ptr + mats[0].r * mats[0].c, partial[0].begin());
            pdim[0] = {mats[0].r, mats[0].c};
            n_prods = 1;
        }
        for (int n = n_prods; n < n_mats; ++n) {
            const Mat& A = mats[n];
            int K = pdim[n - 1].first, C = pdim[n - 1].second;  // A.c == K
            cplx* out = partial[n].data();
            const cplx* B = partial[n - 1].data();
            for (int j = 0; j < C; ++j)
                for (int i = 0; i < A.r; ++i) {
                    cplx acc = 0;
                    for (int k = 0; k < K; ++k) acc += A.ptr[i + A.r * k] * B[k + K * j];
                    out[i + A.r * j] = acc;
                }
            pdim[n] = {A.r, C};
            if (flops) *flops += 8.0 * A.r * K * C;
        }
        n_prods = n_mats;
        r = pdim[n_mats - 1].first; c = pdim[n_mats - 1].second;
        return partial[n_mats - 1].data();
    }
};

struct Evaluator {
    const Oracle& o;
    int mode, order, n_nodes;
    bool use_bold;
    std::vector<Node> conf;
    std::vector<int> fixed_kind;      // 0 = free
    std::vector<int> fixed_op;        // operator index for OPER nodes
    std::vector<double> times;        // per position (1-based -> index pos-1)
    std::vector<int> top_to_conf_pos, var_time_pos;  // :269-275
    std::vector<std::vector<cplx>> ppgf;             // [interval][packed blocks]  i*P
    // node_mats[pos][op][source sector] = op_block * iP  (:376-388)
    std::vector<std::vector<std::vector<std::vector<cplx>>>> node_mats;
    std::vector<std::vector<cplx>> ident_mats;       // [pos] packed iP (finite-time identity)
    std::vector<cplx> pair_ints;                     // [arc][pair]
    std::vector<int> selected;
    std::vector<std::vector<cplx>> top_result;       // per sector d x d
    std::vector<LazyProduct> prods;
    std::vector<cplx> result;                        // packed
    double flops = 0, leaves = 0;
    bool offdiag_error = false;

    // TopologyEvaluator(exp, order, use_bold_prop, fixed_nodes)  :249-330
    Evaluator(const Oracle& o_, int mode_, int order_, int n_pts_after, double t_i, double t_w,
              double t_f, int corr_idx)
        : o(o_), mode(mode_), order(order_) {
        use_bold = (mode != QO_MODE_BARE);
        int n = order;
        if (mode == QO_MODE_BARE) n_nodes = 2 * n + 2; else n_nodes = 2 * n + 3;   // :254
        conf.assign(n_nodes, Node{0, -1, -1});
        fixed_kind.assign(n_nodes, 0); fixed_op.assign(n_nodes, -1);
        times.assign(n_nodes, 0.0);
        auto fix = [&](int pos, int kind, double t, int op) {
            fixed_kind[pos - 1] = kind; times[pos - 1] = t; fixed_op[pos - 1] = op;
            conf[pos - 1] = Node{kind, -1, -1};
        };
        if (mode == QO_MODE_BARE) {           // src/inchworm.jl:260,274
            fix(1, IDENT, t_i, -1); fix(2 * n + 2, IDENT, t_f, -1);
        } else if (mode == QO_MODE_BOLD) {    // src/inchworm.jl:150,163
            int d_before = 2 * n - n_pts_after;
            fix(1, IDENT, t_i, -1); fix(d_before + 2, INCH, t_w, -1); fix(2 * n + 3, IDENT, t_f, -1);
        } else {                              // src/inchworm.jl:817-819,835,849
            int d_before = 2 * n - n_pts_after;
            fix(1, OPER, t_i, o.corr_B[corr_idx]); fix(d_before + 2, OPER, t_w, o.corr_A[corr_idx]);
            fix(2 * n + 3, IDENT, t_f, -1);
        }
        for (int pos = n_nodes; pos >= 1; --pos)
            if (!fixed_kind[pos - 1]) { top_to_conf_pos.push_back(pos); var_time_pos.push_back(pos); }
        ppgf.assign(n_nodes - 1, std::vector<cplx>(o.bsize));
        pair_ints.assign((size_t)std::max(order, 1) * std::max<size_t>(o.pair_table.size(), 1), cplx(0));
        selected.assign(std::max(order, 1), -1);
        int maxd = *std::max_element(o.dim.begin(), o.dim.end());
        top_result.resize(o.S);
        prods.resize(o.S);
        for (int s = 0; s < o.S; ++s) {
            top_result[s].assign(o.dim[s] * o.dim[s], cplx(0));
            prods[s].init(2 * n_nodes - 1, maxd);
            prods[s].flops = &flops;
        }
        result.assign(o.bsize, cplx(0));
        node_mats.resize(n_nodes);
        for (int pos = 0; pos < n_nodes; ++pos) {
            node_mats[pos].resize(o.ops.size());
            for (size_t op = 0; op < o.ops.size(); ++op) {
                node_mats[pos][op].resize(o.S);
                for (int s = 0; s < o.S; ++s) node_mats[pos][op][s] = o.ops[op].mat[s];  // bare copy
            }
        }
        ident_mats.assign(n_nodes, std::vector<cplx>(o.bsize));
    }

    // a cached evaluator is re-used at the next step: only the times of the fixed nodes change
    void set_fixed_times(int n_pts_after, double t_i, double t_w, double t_f) {
        const int n = order;
        if (mode == QO_MODE_BARE) { times[0] = t_i; times[2 * n + 1] = t_f; }
        else { times[0] = t_i; times[2 * n - n_pts_after + 1] = t_w; times[2 * n + 2] = t_f; }
        flops = 0; leaves = 0;
    }

    // i * P_s(t_f, t_i) for every sector  (:357-374; bare: src/exact_atomic_ppgf.jl:109-118)
    void fill_ppgf(int i, double t_i, double t_f) {
        if (t_f < t_i) t_f = t_i;  // :362-364
        std::vector<cplx>& out = ppgf[i];
        double h = o.beta / (o.n_tau - 1);
        for (int s = 0; s < o.S; ++s) {
            int d = o.dim[s];
            cplx* blk = out.data() + o.boff[s];
            if (use_bold) {
                for (int e = 0; e < d * d; ++e)
                    blk[e] = I_ * grid_interp(o.P.data() + o.boff[s] + e, o.bsize, o.n_tau, h, t_f, t_i);
            } else {
                for (int e = 0; e < d * d; ++e) blk[e] = 0;
                for (int k = 0; k < d; ++k)  // -i exp(-i dz E), dz = -i (t_f - t_i)
                    blk[k + d * k] = I_ * (-I_ * std::exp(-(t_f - t_i) * o.E[s][k]));
            }
        }
    }

    // update_finite_time!: node matrix = op_block * iP_s  (:162-166,376-388)
    void update_node_mats() {
        for (int pos = 1; pos < n_nodes; ++pos) {  // pos is 0-based here; skip position 1 (:377)
            const std::vector<cplx>& ip = ppgf[pos - 1];
            for (size_t op = 0; op < o.ops.size(); ++op)
                for (int s = 0; s < o.S; ++s) {
                    int tgt = o.ops[op].target[s];
                    if (tgt < 0) continue;
                    int r = o.dim[tgt], c = o.dim[s];
                    const cplx* A = o.ops[op].mat[s].data();
                    const cplx* B = ip.data() + o.boff[s];
                    cplx* out = node_mats[pos][op][s].data();
                    for (int j = 0; j < c; ++j)
                        for (int i2 = 0; i2 < r; ++i2) {
                            cplx acc = 0;
                            for (int k = 0; k < c; ++k) acc += A[i2 + r * k] * B[k + c * j];
                            out[i2 + r * j] = acc;
                        }
                }
            ident_mats[pos] = ip;  // identity operator times iP
        }
    }

    // (eval::TopologyEvaluator)(topologies, times)  :350-437
    const std::vector<cplx>& operator()(const std::vector<Topology>& tops, const double* t) {
        for (size_t k = 0; k < var_time_pos.size(); ++k) times[var_time_pos[k] - 1] = t[k];  // :354
        for (int i = 0; i < n_nodes - 1; ++i) fill_ppgf(i, times[i], times[i + 1]);
        update_node_mats();
        std::fill(result.begin(), result.end(), cplx(0));
        int np = (int)o.pair_table.size();
        for (const Topology& top : tops) {  // :392
            for (int a = 0; a < order; ++a) {  // :397-416
                int pos_head = top_to_conf_pos[top.b[a] - 1];
                int pos_tail = top_to_conf_pos[top.a[a] - 1];
                conf[pos_head - 1] = Node{PAIR, a, 1};
                conf[pos_tail - 1] = Node{PAIR, a, 2};
                double ti = times[pos_head - 1], tf = times[pos_tail - 1];
                if (tf < ti) tf = ti;
                for (int p = 0; p < np; ++p)
                    pair_ints[a * np + p] = I_ * delta_eval(o.deltas[o.pair_table[p]], tf, ti);
            }
            for (int s = 0; s < o.S; ++s) std::fill(top_result[s].begin(), top_result[s].end(), cplx(0));
            for (int s = 0; s < o.S; ++s) traverse(1, s, s, cplx(1.0));  // :423-429
            cplx coef = -I_ * (double)top.parity * ((order % 2 == 0) ? 1.0 : -1.0);  // :431
            for (int s = 0; s < o.S; ++s)
                for (int e = 0; e < o.dim[s] * o.dim[s]; ++e) result[o.boff[s] + e] += coef * top_result[s][e];
        }
        return result;
    }

    // _traverse_configuration_tree!  :454-556
    void traverse(int pos, int s_i, int s_f, cplx w) {
        LazyProduct& lp = prods[s_f];
        if (pos > n_nodes) {  // leaf :461-467
            if (s_i != s_f) { offdiag_error = true; return; }
            int r, c;
            const cplx* m = lp.eval(r, c);
            for (int e = 0; e < r * c; ++e) top_result[s_i][e] += w * m[e];
            flops += 8.0 * r * c; leaves += 1;
            return;
        }
        const Node& node = conf[pos - 1];
        if (node.kind == PAIR) {
            if (node.operator_index == 1) {  // head :475-493
                for (int p : o.attachable[s_i]) {
                    selected[node.arc_index] = p;
                    int op = o.pair_op_i[p];
                    int s_next = o.ops[op].target[s_i];
                    lp.push(o.dim[s_next], o.dim[s_i], node_mats[pos - 1][op][s_i].data());
                    traverse(pos + 1, s_next, s_f, w);
                    lp.pop();
                }
            } else {  // tail :495-518
                int p = selected[node.arc_index];
                int op = o.pair_op_f[p];
                int s_next = o.ops[op].target[s_i];
                if (s_next >= 0) {
                    lp.push(o.dim[s_next], o.dim[s_i], node_mats[pos - 1][op][s_i].data());
                    cplx w2 = pair_ints[node.arc_index * o.pair_table.size() + p] * w;
                    traverse(pos + 1, s_next, s_f, w2);
                    lp.pop();
                }
            }
        } else if (node.kind == OPER) {  // :520-536
            int op = fixed_op[pos - 1];
            int s_next = o.ops[op].target[s_i];
            if (s_next >= 0) {
                lp.push(o.dim[s_next], o.dim[s_i], node_mats[pos - 1][op][s_i].data());
                traverse(pos + 1, s_next, s_f, w);
                lp.pop();
            }
        } else {  // identity / inch :538-550
            if (pos != 1) lp.push(o.dim[s_i], o.dim[s_i], ident_mats[pos - 1].data() + o.boff[s_i]);
            traverse(pos + 1, s_i, s_f, w);
            if (pos != 1) lp.pop();
        }
    }
};

// ------------------------------------------------------------------------------------------------
// qMC integration (src/qmc_integrate.jl)
// ------------------------------------------------------------------------------------------------

static double simplex_volume(int d, double edge) {  // :46
    double v = 1.0;
    for (int i = 1; i <= d; ++i) v *= edge / i;
    return v;
}

// make_trans_f(RootTransform) :225-235
static void root_transform(int d, double u_i, double u_diff, const double* x, double* u) {
    if (d == 0) return;
    u[0] = std::pow(x[0], 1.0 / d);
    for (int i = 1; i < d; ++i) u[i] = u[i - 1] * std::pow(x[i], 1.0 / (d - i));
    for (int i = 0; i < d; ++i) u[i] = u_i + u[i] * u_diff;
}

// make_trans_f(DoubleSimplexRootTransform) :425-449
static void double_simplex_transform(int d_lesser, int d_greater, double u_i, double u_w,
                                     double diff_wi, double diff_fw, const double* x, double* u) {
    int d = d_lesser + d_greater;
    u[0] = std::pow(x[0], 1.0 / d_greater);
    for (int i = 1; i < d_greater; ++i) u[i] = u[i - 1] * std::pow(x[i], 1.0 / (d_greater - i));
    for (int i = 0; i < d_greater; ++i) { u[i] *= diff_fw; u[i] += u_w; }
    if (d_lesser == 0) return;
    u[d_greater] = std::pow(x[d_greater], 1.0 / d_lesser);
    for (int i = d_greater + 1; i < d; ++i) u[i] = u[i - 1] * std::pow(x[i], 1.0 / (d - i));
    for (int i = d_greater; i < d; ++i) { u[i] *= diff_wi; u[i] += u_i; }
}

// One qMC block for one entry over the Sobol index range [start, start+count) of a sequence of
// N_total points: returns sum / N_total, i.e. rank_weight * contour_integral(...) of
// src/inchworm.jl:176-188 (bold), :281-293 (bare), :855-875 (correlator, trace taken by caller).
// contour_integral / qmc_integral: src/qmc_integrate.jl:497-507,597-612.
static bool eval_entry_range_core(Oracle& o, const Entry& e, double t_i, double t_w, double t_f, int corr_idx,
                             const uint32_t* m, const uint32_t* x0, uint64_t start, uint64_t count,
                             uint64_t N_total, cplx* out, double* flops, double* leaves, Evaluator* cached = nullptr) {
    int n = e.order, d = 2 * n;
    std::unique_ptr<Evaluator> own;
    if (!cached) own.reset(new Evaluator(o, e.mode, n, e.n_pts_after, t_i, t_w, t_f, corr_idx));
    else cached->set_fixed_times(e.n_pts_after, t_i, t_w, t_f);
    Evaluator& ev = cached ? *cached : *own;
    std::vector<cplx> acc(o.bsize, cplx(0));
    if (n == 0) {  // exact, no sampling: src/inchworm.jl:148-156,258-266,833-841
        const std::vector<cplx>& r = ev(e.tops, nullptr);
        for (int k = 0; k < o.bsize; ++k) out[k] = r[k];
        if (flops) *flops = ev.flops;
        if (leaves) *leaves = ev.leaves;
        return ev.offdiag_error;
    }
    SobolSeq seq;
    seq.D = d; seq.n = 0;
    seq.m.assign(m, m + d * 32);
    seq.x.assign(x0, x0 + d);
    seq.skip_exact(start);  // skip!(seq, first(N_range) - 1, exact=true)
    std::vector<double> x(d), u(d);
    double jac;
    int d_after = e.n_pts_after, d_before = d - d_after;
    if (e.mode == QO_MODE_BARE) jac = simplex_volume(d, t_f - t_i);
    else jac = simplex_volume(d_before, t_w - t_i) * simplex_volume(d_after, t_f - t_w);  // :458-463
    // prod of branch directions: (-i)^d  (:565-569,610)
    cplx dir = 1.0;
    for (int k = 0; k < d; ++k) dir *= -I_;
    for (uint64_t i = 0; i < count; ++i) {
        seq.next(x.data());
        if (e.mode == QO_MODE_BARE) root_transform(d, t_i, t_f - t_i, x.data(), u.data());
        else double_simplex_transform(d_before, d_after, t_i, t_w, t_w - t_i, t_f - t_w, x.data(), u.data());
        bool ok = true;
        for (int k = 0; k < d; ++k) if (!(u[k] >= 0.0)) ok = false;  // :608
        if (!ok) continue;
        const std::vector<cplx>& r = ev(e.tops, u.data());
        for (int k = 0; k < o.bsize; ++k) acc[k] += (dir * r[k]) * jac;  // :504,610
    }
    for (int k = 0; k < o.bsize; ++k) out[k] = acc[k] / (double)N_total;
    if (flops) *flops = count ? ev.flops / (double)count : 0;
    if (leaves) *leaves = count ? ev.leaves / (double)count : 0;
    return ev.offdiag_error;
}

static void eval_entry_range(Oracle& o, const Entry& e, double t_i, double t_w, double t_f, int corr_idx,
                             const uint32_t* m, const uint32_t* x0, uint64_t start, uint64_t count,
                             uint64_t N_total, cplx* out, double* flops, double* leaves) {
    if (eval_entry_range_core(o, e, t_i, t_w, t_f, corr_idx, m, x0, start, count, N_total, out, flops, leaves))
        o.err = "block off-diagonal contribution (src/topology_eval.jl:462)";
}
// for the pool's workers: the sector-mismatch flag stays in the evaluator (the caller collects it)
static void eval_entry_range_quiet(Oracle& o, const Entry& e, double t_i, double t_w, double t_f, int corr_idx,
                                   const uint32_t* m, const uint32_t* x0, uint64_t start, uint64_t count,
                                   uint64_t N_total, cplx* out, double* flops, double* leaves, Evaluator* ev) {
    eval_entry_range_core(o, e, t_i, t_w, t_f, corr_idx, m, x0, start, count, N_total, out, flops, leaves, ev);
}

// split_count / range_from_chunks_and_idx: src/utility.jl:164-179 (0-based start returned)
static void rank_sub_range(uint64_t N, int n_ranks, int rank, uint64_t* start, uint64_t* count) {
    uint64_t q = N / n_ranks, r = N % n_ranks;
    uint64_t s = 0;
    for (int i = 0; i < rank; ++i) s += (uint64_t)i < r ? q + 1 : q;
    *start = s;
    *count = (uint64_t)rank < r ? q + 1 : q;
}

// Persistent host threads.  Worker r plays MPI rank r of the reference (src/mpi.jl:49-54): in every call it
// evaluates its rank_sub_range of every sampled entry, one after the other, with evaluators it keeps from call to
// call (keyed by entry id and correlator index; dropped when the model or an entry changes).
struct Pool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    uint64_t gen = 0;
    int pending = 0;
    bool stop = false;
    std::function<void(int)> job;
    struct Cache { uint64_t epoch = ~0ull; std::map<std::pair<int, int>, std::unique_ptr<Evaluator>> ev; };
    std::vector<Cache> cache;
    explicit Pool(int nt) : cache(nt) {
        for (int r = 0; r < nt; ++r)
            th.emplace_back([this, r]() {
                uint64_t seen = 0;
                for (;;) {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_go.wait(lk, [&] { return stop || gen != seen; });
                    if (stop) return;
                    seen = gen;
                    lk.unlock();
                    job(r);
                    lk.lock();
                    if (--pending == 0) cv_done.notify_one();
                }
            });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_go.notify_all();
        for (auto& t : th) t.join();
    }
    void run(const std::function<void(int)>& f) {
        std::unique_lock<std::mutex> lk(mu);
        job = f; pending = (int)th.size(); ++gen;
        cv_go.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

// Whole range [start, start+count) split over host threads with the same rule the reference uses
// over MPI ranks (src/mpi.jl:49-54), partial sums added in rank order (src/mpi.jl:104-127).
// use_pool = 0: one std::thread per (entry, call), each with a fresh evaluator.
static void eval_entry(Oracle& o, const Entry& e, double t_i, double t_w, double t_f, int corr_idx,
                       const uint32_t* m, const uint32_t* x0, uint64_t start, uint64_t count,
                       uint64_t N_total, cplx* out) {
    int nt = std::max(1, o.n_threads);
    if (e.order == 0 || nt == 1 || count < (uint64_t)nt * 4) {
        eval_entry_range(o, e, t_i, t_w, t_f, corr_idx, m, x0, start, count, N_total, out,
                         &o.last_flops_per_sample, &o.last_leaves);
        return;
    }
    std::vector<std::vector<cplx>> parts(nt, std::vector<cplx>(o.bsize));
    std::vector<double> fl(nt, 0), lv(nt, 0);
    std::vector<std::thread> th;
    for (int r = 0; r < nt; ++r)
        th.emplace_back([&, r]() {
            uint64_t s, c;
            rank_sub_range(count, nt, r, &s, &c);
            eval_entry_range(o, e, t_i, t_w, t_f, corr_idx, m, x0, start + s, c, N_total,
                             parts[r].data(), &fl[r], &lv[r]);
        });
    for (auto& t : th) t.join();
    for (int k = 0; k < o.bsize; ++k) { out[k] = 0; for (int r = 0; r < nt; ++r) out[k] += parts[r][k]; }
    o.last_flops_per_sample = fl[0]; o.last_leaves = lv[0];
}

// All entries of one call on the persistent pool; `out` is [n_entries][bsize].
static void eval_entries_pooled(Oracle& o, int n_entries, const int32_t* ids, double t_i, double t_w, double t_f, int corr_idx,
                                const uint32_t* sobol_m, const uint32_t* sobol_x0, uint64_t start, uint64_t count,
                                uint64_t N_total, cplx* out) {
    const int nt = std::max(1, o.n_threads);
    if (!o.pool || (int)o.pool->th.size() != nt) { delete o.pool; o.pool = new Pool(nt); }
    std::vector<size_t> moff(n_entries), xoff(n_entries);
    size_t mo = 0, xo = 0;
    for (int i = 0; i < n_entries; ++i) { moff[i] = mo; xoff[i] = xo; mo += (size_t)2 * o.entries[ids[i]].order * 32; xo += 2 * o.entries[ids[i]].order; }
    std::vector<cplx> parts((size_t)nt * n_entries * o.bsize, cplx(0));
    std::vector<double> fl((size_t)nt * n_entries, 0), lv((size_t)nt * n_entries, 0);
    std::vector<char> bad(nt, 0);
    const uint64_t epoch = o.cache_epoch;
    o.pool->run([&](int r) {
        Pool::Cache& c = o.pool->cache[r];
        if (c.epoch != epoch) { c.ev.clear(); c.epoch = epoch; }
        uint64_t s, n;
        rank_sub_range(count, nt, r, &s, &n);
        for (int i = 0; i < n_entries; ++i) {
            const Entry& e = o.entries[ids[i]];
            if (e.order == 0) continue;   // exact entries: evaluated once by the caller below
            std::unique_ptr<Evaluator>& ev = c.ev[std::make_pair((int)ids[i], corr_idx)];
            if (!ev) ev.reset(new Evaluator(o, e.mode, e.order, e.n_pts_after, t_i, t_w, t_f, corr_idx));
            ev->offdiag_error = false;
            eval_entry_range_quiet(o, e, t_i, t_w, t_f, corr_idx, sobol_m + moff[i], sobol_x0 + xoff[i], start + s, n, N_total,
                                   parts.data() + ((size_t)r * n_entries + i) * o.bsize, &fl[(size_t)r * n_entries + i],
                                   &lv[(size_t)r * n_entries + i], ev.get());
            if (ev->offdiag_error) bad[r] = 1;
        }
    });
    for (int i = 0; i < n_entries; ++i) {
        const Entry& e = o.entries[ids[i]];
        cplx* dst = out + (size_t)i * o.bsize;
        if (e.order == 0) {
            eval_entry_range(o, e, t_i, t_w, t_f, corr_idx, sobol_m + moff[i], sobol_x0 + xoff[i], start, count, N_total, dst,
                             &o.last_flops_per_sample, &o.last_leaves);
            continue;
        }
        for (int k = 0; k < o.bsize; ++k) {
            cplx v = 0;
            for (int r = 0; r < nt; ++r) v += parts[((size_t)r * n_entries + i) * o.bsize + k];
            dst[k] = v;
        }
        o.last_flops_per_sample = fl[i]; o.last_leaves = lv[i];
    }
    for (int r = 0; r < nt; ++r) if (bad[r]) o.err = "block off-diagonal contribution (src/topology_eval.jl:462)";
}

// ------------------------------------------------------------------------------------------------
// C interface (ctypes) — test harness only
// ------------------------------------------------------------------------------------------------

extern "C" {

int qo_sobol_direction_numbers(int D, uint32_t* m) { return sobol_direction_numbers(D, m); }

void qo_sobol_scramble(int D, uint32_t* m, uint32_t* x0, const uint8_t* shift_bits, const uint8_t* ltm_bits) {
    sobol_scramble(D, m, x0, shift_bits, ltm_bits);
}

// Sequential generation of `count` points after skipping `skip` points (exact).  Writes raw
// 32-bit integers (xi) and/or the Float64 values (xf = xi * 2^-32), row-major [count][D].
void qo_sobol_points(int D, const uint32_t* m, const uint32_t* x0, uint64_t skip, uint64_t count,
                     uint32_t* xi, double* xf) {
    SobolSeq seq; seq.D = D; seq.n = 0; seq.m.assign(m, m + D * 32); seq.x.assign(x0, x0 + D);
    seq.skip_exact(skip);
    std::vector<double> buf(D);
    for (uint64_t i = 0; i < count; ++i) {
        seq.next(buf.data());
        for (int d = 0; d < D; ++d) { if (xi) xi[i * D + d] = seq.x[d]; if (xf) xf[i * D + d] = buf[d]; }
    }
}

// Returns the number of topologies; fills pairs[n][order][2] and parity[n] when non-null.
int qo_topologies(int order, int k, int with_external_arc, int32_t* pairs, int32_t* parity) {
    std::vector<Topology> t = get_topologies_at_order(order, k, with_external_arc != 0);
    if (pairs)
        for (size_t i = 0; i < t.size(); ++i) {
            for (int a = 0; a < order; ++a) { pairs[(i * order + a) * 2] = t[i].a[a]; pairs[(i * order + a) * 2 + 1] = t[i].b[a]; }
            if (parity) parity[i] = t[i].parity;
        }
    return (int)t.size();
}

void qo_transform(int mode, int d_before, int d_after, double t_i, double t_w, double t_f,
                  const double* x, double* u, double* jac) {
    if (mode == QO_MODE_BARE) {
        int d = d_before + d_after;
        root_transform(d, t_i, t_f - t_i, x, u);
        *jac = simplex_volume(d, t_f - t_i);
    } else {
        double_simplex_transform(d_before, d_after, t_i, t_w, t_w - t_i, t_f - t_w, x, u);
        *jac = simplex_volume(d_before, t_w - t_i) * simplex_volume(d_after, t_f - t_w);
    }
}

void qo_rank_sub_range(uint64_t N, int n_ranks, int rank, uint64_t* start, uint64_t* count) {
    rank_sub_range(N, n_ranks, rank, start, count);
}

Oracle* qo_create() { return new Oracle(); }
void qo_destroy(Oracle* o) { delete o->pool; delete o; }
const char* qo_last_error(Oracle* o) { return o->err.c_str(); }
void qo_set_threads(Oracle* o, int n) { o->n_threads = n; }
// 1 (default): persistent worker pool with cached evaluators; 0: one std::thread per (entry, call)
void qo_set_pool(Oracle* o, int on) { o->use_pool = on; }

// Model: sectors, P0 energies (E + lambda0 per state), operators as sector-block matrices,
// interaction pairs (operator_i, operator_f, Delta table), correlator operator pairs (A, B).
// Mirrors the fields of `Expansion` (src/expansion.jl:110-134,168-183).
int qo_set_model(Oracle* o, int S, const int32_t* dims, const double* energies, int n_ops,
                 const int32_t* op_target, const int64_t* op_mat_off, const double* op_pool,
                 int n_pairs, const int32_t* pair_op_i, const int32_t* pair_op_f,
                 const int32_t* pair_table, int n_corr, const int32_t* corr_A, const int32_t* corr_B) {
    ++o->cache_epoch;
    o->S = S;
    o->dim.assign(dims, dims + S);
    o->boff.assign(S, 0);
    int off = 0, eoff = 0;
    o->E.resize(S);
    for (int s = 0; s < S; ++s) {
        o->boff[s] = off; off += dims[s] * dims[s];
        o->E[s].assign(energies + eoff, energies + eoff + dims[s]); eoff += dims[s];
    }
    o->bsize = off;
    o->ops.resize(n_ops);
    for (int op = 0; op < n_ops; ++op) {
        o->ops[op].target.assign(op_target + (size_t)op * S, op_target + (size_t)(op + 1) * S);
        o->ops[op].mat.resize(S);
        for (int s = 0; s < S; ++s) {
            int t = o->ops[op].target[s];
            if (t < 0) continue;
            int sz = dims[t] * dims[s];
            const double* p = op_pool + 2 * op_mat_off[(size_t)op * S + s];
            o->ops[op].mat[s].resize(sz);
            for (int e = 0; e < sz; ++e) o->ops[op].mat[s][e] = cplx(p[2 * e], p[2 * e + 1]);
        }
    }
    o->pair_op_i.assign(pair_op_i, pair_op_i + n_pairs);
    o->pair_op_f.assign(pair_op_f, pair_op_f + n_pairs);
    o->pair_table.assign(pair_table, pair_table + n_pairs);
    o->attachable.assign(S, {});
    for (int s = 0; s < S; ++s)  // findall(op -> haskey(op[1], s), pair_operator_mat)
        for (int p = 0; p < n_pairs; ++p)
            if (o->ops[pair_op_i[p]].target[s] >= 0) o->attachable[s].push_back(p);
    o->corr_A.assign(corr_A, corr_A + n_corr);
    o->corr_B.assign(corr_B, corr_B + n_corr);
    return 0;
}

int qo_set_delta(Oracle* o, int table_id, int kind, int n, double beta, const double* data) {
    if ((int)o->deltas.size() <= table_id) o->deltas.resize(table_id + 1);
    DeltaTable& t = o->deltas[table_id];
    t.kind = kind; t.n = n; t.beta = beta;
    t.y.resize(n);
    for (int k = 0; k < n; ++k) t.y[k] = cplx(data[2 * k], data[2 * k + 1]);
    if (kind == QO_DELTA_SPLINE) spline_prepare(t);
    return 0;
}

int qo_set_grid(Oracle* o, int n_tau, double beta) {
    o->n_tau = n_tau; o->beta = beta;
    o->P.assign((size_t)n_tau * o->bsize, cplx(0));
    return 0;
}

// P table rows [first, first+count): packed blocks per grid point (column-major inside a block).
int qo_set_P(Oracle* o, int first, int count, const double* data) {
    for (size_t k = 0; k < (size_t)count * o->bsize; ++k)
        o->P[(size_t)first * o->bsize + k] = cplx(data[2 * k], data[2 * k + 1]);
    return 0;
}

int qo_get_P(Oracle* o, int first, int count, double* data) {
    for (size_t k = 0; k < (size_t)count * o->bsize; ++k) {
        data[2 * k] = o->P[(size_t)first * o->bsize + k].real();
        data[2 * k + 1] = o->P[(size_t)first * o->bsize + k].imag();
    }
    return 0;
}

int qo_set_topologies(Oracle* o, int entry_id, int mode, int order, int n_pts_after, int n_top,
                      const int32_t* pairs, const int32_t* parity) {
    ++o->cache_epoch;
    if ((int)o->entries.size() <= entry_id) o->entries.resize(entry_id + 1);
    Entry& e = o->entries[entry_id];
    e.mode = mode; e.order = order; e.n_pts_after = n_pts_after;
    e.tops.resize(n_top);
    for (int i = 0; i < n_top; ++i) {
        e.tops[i].a.resize(order); e.tops[i].b.resize(order);
        for (int a = 0; a < order; ++a) {
            e.tops[i].a[a] = pairs[((size_t)i * order + a) * 2];
            e.tops[i].b[a] = pairs[((size_t)i * order + a) * 2 + 1];
        }
        e.tops[i].parity = parity[i];
    }
    return 0;
}

// Evaluate given entries at fixed (t_i, t_w, t_f) over the Sobol index range [start, start+count)
// of N_total points.  sobol_m / sobol_x0 hold, per entry, the direction numbers [2*order][32] and
// the first point [2*order] (concatenated).  out: [n_entries][bsize] complex.
int qo_eval(Oracle* o, double t_i, double t_w, double t_f, int corr_idx, int n_entries,
            const int32_t* entry_ids, const uint32_t* sobol_m, const uint32_t* sobol_x0,
            uint64_t start, uint64_t count, uint64_t N_total, double* out) {
    o->err.clear();
    size_t moff = 0, xoff = 0;
    std::vector<cplx> buf(o->bsize);
    if (o->use_pool && o->n_threads > 1 && count >= (uint64_t)o->n_threads * 4) {
        std::vector<cplx> res((size_t)n_entries * o->bsize);
        eval_entries_pooled(*o, n_entries, entry_ids, t_i, t_w, t_f, corr_idx, sobol_m, sobol_x0, start, count, N_total, res.data());
        for (size_t k = 0; k < res.size(); ++k) { out[2 * k] = res[k].real(); out[2 * k + 1] = res[k].imag(); }
        return o->err.empty() ? 0 : 4;
    }
    for (int i = 0; i < n_entries; ++i) {
        const Entry& e = o->entries[entry_ids[i]];
        int d = 2 * e.order;
        eval_entry(*o, e, t_i, t_w, t_f, corr_idx, sobol_m + moff, sobol_x0 + xoff, start, count, N_total, buf.data());
        for (int k = 0; k < o->bsize; ++k) { out[2 * ((size_t)i * o->bsize + k)] = buf[k].real(); out[2 * ((size_t)i * o->bsize + k) + 1] = buf[k].imag(); }
        moff += (size_t)d * 32; xoff += d;
    }
    return o->err.empty() ? 0 : 4;
}

// Per-sample evaluator values at explicit times (no transform, no Jacobian): what
// test/topology_eval.jl:137-141 records.  times: [n_samples][2*order] descending.
int qo_eval_at_times(Oracle* o, int entry_id, double t_i, double t_w, double t_f, int corr_idx,
                     int n_samples, const double* times, double* out) {
    const Entry& e = o->entries[entry_id];
    Evaluator ev(*o, e.mode, e.order, e.n_pts_after, t_i, t_w, t_f, corr_idx);
    for (int i = 0; i < n_samples; ++i) {
        const std::vector<cplx>& r = ev(e.tops, times + (size_t)i * 2 * e.order);
        for (int k = 0; k < o->bsize; ++k) { out[2 * ((size_t)i * o->bsize + k)] = r[k].real(); out[2 * ((size_t)i * o->bsize + k) + 1] = r[k].imag(); }
    }
    o->last_flops_per_sample = n_samples ? ev.flops / n_samples : 0;
    o->last_leaves = n_samples ? ev.leaves / n_samples : 0;
    return ev.offdiag_error ? 4 : 0;
}

// Chain FLOPs and leaves per sample of the last qo_eval entry / qo_eval_at_times call.
void qo_last_counts(Oracle* o, double* flops_per_sample, double* leaves_per_sample) {
    *flops_per_sample = o->last_flops_per_sample;
    *leaves_per_sample = o->last_leaves;
}

// set_ppgf! + normalize!(P, tau_f): src/ppgf.jl:495-504,646-668.  Returns lambda.
double qo_set_and_normalize(Oracle* o, int k_f, const double* value, int do_normalize) {
    for (int e = 0; e < o->bsize; ++e) o->P[(size_t)k_f * o->bsize + e] = cplx(value[2 * e], value[2 * e + 1]);
    if (!do_normalize) return 0.0;
    double h = o->beta / (o->n_tau - 1);
    double pmax = -1e300;
    for (int s = 0; s < o->S; ++s)
        for (int i = 0; i < o->dim[s]; ++i)
            pmax = std::max(pmax, -o->P[(size_t)k_f * o->bsize + o->boff[s] + i + o->dim[s] * i].imag());
    double tau_f = k_f * h;
    double lambda = std::log(pmax) / tau_f;
    for (int k = 0; k < o->n_tau; ++k) {
        double f = std::exp(-(k * h) * lambda);
        for (int e = 0; e < o->bsize; ++e) o->P[(size_t)k * o->bsize + e] *= f;
    }
    return lambda;
}

}  // extern "C"
