// qiw_compile.cpp — turns one TopologiesInputData + the model into a traversal program.
//
// The reference rebuilds a TopologyEvaluator at every step (src/inchworm.jl:165 ->
// src/topology_eval.jl:249-330) and, for every sample and topology, walks the configuration tree
// recursively with run-time pruning by sector selection rules (_traverse_configuration_tree!,
// :454-556).  Which branches survive depends only on (model, topology, initial sector) — not on the
// sample — so here the walk is done ONCE on the host: dead branches are dropped and the surviving
// tree is emitted as a flat pre-order word stream that every sample (= one GPU lane) replays in
// lock step.  Node order inside a tree is the reference's own DFS order.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <system_error>
#include <thread>
#include <tuple>

#include "qiw_host.hpp"

namespace qiw {

namespace {

enum Kind { K_FREE = 0, K_IDENT = 1, K_INCH = 2, K_OPER = 3, K_HEAD = 4, K_TAIL = 5 };

struct Builder {
    const HostModel& m;
    EntryProgram& e;
    int n_nodes;
    int kind[kMaxNodes + 2];
    int arc_of[kMaxNodes + 2];      // arc index of a head/tail position
    int head_pos_of_arc[kMaxOrder + 1];
    int selected[kMaxOrder + 1];    // pair chosen at the head of each arc
    int s_init = 0;
    cplx top_sign;
    bool offdiag = false;
    std::map<std::pair<double, double>, int> coef_index;
    std::map<std::tuple<int, int, int>, int> dslot_index;
    // statistics of the reference algorithm for the current chain
    int64_t leaves = 0, edges = 0;
    double flops = 0;
    std::vector<uint32_t> path;   // table slots of the factors collected along the current branch

    Builder(const HostModel& m_, EntryProgram& e_) : m(m_), e(e_), n_nodes(e_.n_nodes) {}

    int coef_id(cplx c) {
        // canonicalise signed zeros so that +0 / -0 do not create distinct entries
        double re = c.real() == 0.0 ? 0.0 : c.real(), im = c.imag() == 0.0 ? 0.0 : c.imag();
        auto key = std::make_pair(re, im);
        auto it = coef_index.find(key);
        if (it != coef_index.end()) return it->second;
        int id = (int)e.coefs.size();
        e.coefs.push_back(cplx(re, im));
        coef_index[key] = id;
        return id;
    }

    int dslot_id(int pos_tail, int pos_head, int table) {
        auto key = std::make_tuple(pos_tail, pos_head, table);
        auto it = dslot_index.find(key);
        if (it != dslot_index.end()) return it->second;
        int id = (int)e.dslots.size();
        e.dslots.push_back(DeltaSlot{pos_tail, pos_head, table});
        dslot_index[key] = id;
        return id;
    }

    // Emits the sibling nodes of position `pos` (one per live alternative: a head position has one
    // alternative per attachable pair, every other position has one), entered in sector `s` with
    // the product `coef` of operator matrix elements collected so far (scalar models only).
    // `first` = no matrix has been pushed yet (the reference's LazyMatrixProduct copies its first
    // factor instead of multiplying, src/utility.jl:304-308).
    // Returns the number of leaves below and, through n_emitted, the number of nodes emitted at
    // this level; dead branches emit nothing.
    int64_t node(int pos, int s, cplx coef, bool first, uint32_t& n_emitted) {
        int ops[256], pair_of[256], nops = 0;
        const int k = kind[pos];
        if (k == K_IDENT || k == K_INCH) { ops[nops] = -1; pair_of[nops++] = -1; }
        else if (k == K_OPER) { ops[nops] = e.fixed_op[pos]; pair_of[nops++] = -1; }
        else if (k == K_HEAD) {  // :475-493
            for (int p : m.attachable[s]) { ops[nops] = m.pair_op_i[p]; pair_of[nops++] = p; }
        } else {  // tail: the pair was chosen at the head (:497)
            const int p = selected[arc_of[pos]];
            ops[nops] = m.pair_op_f[p]; pair_of[nops++] = p;
        }
        int64_t total = 0;
        n_emitted = 0;
        for (int alt = 0; alt < nops; ++alt) {
            const int op = ops[alt];
            int s_next = s;
            cplx c2 = coef;
            if (op >= 0) {
                s_next = m.target(op, s);
                if (s_next < 0) continue;  // pruned: no block leaves the current sector (:501,524)
                if (e.scalar) c2 *= m.block(op, s)[0];
            }
            if (k == K_HEAD) selected[arc_of[pos]] = pair_of[alt];
            const uint32_t slotA = e.scalar ? (uint32_t)((pos - 2) * m.S + s) : (uint32_t)s;
            const size_t my = e.words.size();
            const int64_t my_edges = edges;
            const double my_flops = flops;
            const size_t my_dslots = e.dslots.size();
            uint32_t slotB = 0;
            if (k == K_TAIL)
                slotB = (uint32_t)(e.nP + dslot_id(pos, head_pos_of_arc[arc_of[pos]], m.pair_table[pair_of[alt]]));
            e.words.push_back(0);  // placeholder, patched below
            // cost of this edge in the reference algorithm: (d_next x d_s) times (d_s x d_init)
            edges += 1;
            if (!first) flops += 8.0 * m.dim[s_next] * m.dim[s] * m.dim[s_init];
            int64_t below = 0;
            uint32_t nchild = 0;
            const size_t path_mark = path.size();
            path.push_back(slotA);
            if (slotB) path.push_back(slotB);
            if (pos == n_nodes) {
                if (s_next != s_init) offdiag = true;  // the @assert at :462
                else {
                    below = 1;
                    flops += 8.0 * m.dim[s_init] * m.dim[s_init];
                    if (e.scalar) {   // leaf record
                        const uint32_t cid = (uint32_t)coef_id(c2 * top_sign);
                        e.records.push_back(cid | ((uint32_t)s_init << 16));
                        for (uint32_t sl : path) e.records.push_back(sl);
                        for (int k = (int)path.size() + 1; k < e.RL; ++k) e.records.push_back(0u);
                    }
                }
            } else {
                below = node(pos + 1, s_next, c2, false, nchild);
            }
            path.resize(path_mark);
            if (below == 0) {  // dead branch: roll back everything it emitted
                e.words.resize(my);
                edges = my_edges;
                flops = my_flops;
                while (e.dslots.size() > my_dslots) {
                    const DeltaSlot& d = e.dslots.back();
                    dslot_index.erase(std::make_tuple(d.pos_tail, d.pos_head, d.table));
                    e.dslots.pop_back();
                }
                continue;
            }
            uint32_t aux = 0;
            if (pos == n_nodes) aux = (uint32_t)coef_id(e.scalar ? c2 * top_sign : top_sign);
            e.words[my] = make_word(slotA, slotB, nchild, aux, e.scalar ? 0u : (uint32_t)(op + 1));
            total += below;
            ++n_emitted;
        }
        return total;
    }

    // All trees with initial sector s_i, in the reference's topology order.  Returns non-zero on a malformed topology.
    int walk_sector(int s_i, int n_top, const int32_t* pairs, const int32_t* parity, const int* kind0, const int* vertex_pos,
                    std::string& err) {
        Builder& b = *this;
        const int n = e.order;
        for (int t = 0; t < n_top; ++t) {
            for (int pos = 1; pos <= e.n_nodes; ++pos) { b.kind[pos] = kind0[pos]; b.arc_of[pos] = -1; }
            for (int a = 0; a < n; ++a) {
                const int va = pairs[((size_t)t * n + a) * 2], vb = pairs[((size_t)t * n + a) * 2 + 1];
                if (va < 1 || vb < 1 || va > 2 * n || vb > 2 * n || va >= vb) { err = "bad topology pair"; return 1; }
                const int pos_tail = vertex_pos[va], pos_head = vertex_pos[vb];  // :398-403
                b.kind[pos_tail] = K_TAIL; b.arc_of[pos_tail] = a;
                b.kind[pos_head] = K_HEAD; b.arc_of[pos_head] = a;
                b.head_pos_of_arc[a] = pos_head;
            }
            // result += -i * parity * (-1)^order * top_result  (src/topology_eval.jl:431)
            b.top_sign = cplx(0, -1) * (double)parity[t] * ((n & 1) ? -1.0 : 1.0);
            b.s_init = s_i;
            const size_t root = e.words.size();
            const int64_t edges0 = b.edges;
            e.words.push_back(0);
            // position 1 is always a fixed node: identity (no factor) or operator B (bare matrix)
            cplx coef(1.0, 0.0);
            int s_next = s_i;
            uint32_t rootop = 0;
            bool first = true;
            if (kind0[1] == K_OPER) {
                const int op = e.fixed_op[1];
                s_next = m.target(op, s_i);
                if (s_next < 0) { e.words.resize(root); continue; }
                if (e.scalar) coef *= m.block(op, s_i)[0];
                rootop = (uint32_t)(op + 1);
                b.edges += 1;  // pushed as the (free) first factor
                first = false;
            }
            uint32_t nchild = 0;
            const double flops0 = b.flops;
            b.path.clear();
            const int64_t below = b.node(2, s_next, coef, first, nchild);
            if (below == 0) { e.words.resize(root); b.edges = edges0; b.flops = flops0; continue; }
            e.words[root] = make_word((uint32_t)s_next, 0, nchild, (uint32_t)s_i, rootop);
            b.leaves += below;
            e.tree_off.push_back((uint32_t)root);
            e.tree_cost.push_back((uint32_t)(b.edges - edges0));
        }
        return 0;
    }
};

}  // namespace

int compile_entry(const HostModel& m, int mode, int order, int n_pts_after, int corr_idx, int n_top,
                  const int32_t* pairs, const int32_t* parity, EntryProgram& e, std::string& err) {
    if (order < 0 || order > kMaxOrder) { err = "order out of range"; return 1; }
    if (mode < 0 || mode > 2) { err = "bad mode"; return 1; }
    if (mode == 2 && (corr_idx < 0 || corr_idx >= (int)m.corr_A.size())) { err = "bad corr_idx"; return 1; }
    e = EntryProgram();
    e.mode = mode; e.order = order; e.n_pts_after = n_pts_after; e.corr_idx = corr_idx;
    e.scalar = m.scalar;
    e.D = 2 * order;
    const int n = order;
    const int d_after = (mode == 0) ? 2 * n : n_pts_after;
    const int d_before = 2 * n - d_after;
    if (order > 0 && mode != 0 && (d_after < 1 || d_after > 2 * n)) { err = "bad n_pts_after"; return 1; }
    e.n_nodes = (mode == 0) ? 2 * n + 2 : 2 * n + 3;  // src/topology_eval.jl:254
    e.nP = (e.n_nodes - 1) * m.S;
    e.L = e.n_nodes - 1 + order;
    e.RL = ((e.L + 1 + 3) / 4) * 4;
    int kind[kMaxNodes + 2];
    for (int p = 0; p <= kMaxNodes; ++p) { e.pos_src[p] = 0; e.fixed_op[p] = -1; kind[p] = K_FREE; }
    // fixed nodes: src/inchworm.jl:150,163 (bold) :260,274 (bare) :817-819,835,849 (correlator)
    if (mode == 0) {
        kind[1] = K_IDENT; e.pos_src[1] = SRC_TI;
        kind[2 * n + 2] = K_IDENT; e.pos_src[2 * n + 2] = SRC_TF;
    } else {
        const int pw = (order == 0) ? 2 : d_before + 2;
        kind[1] = (mode == 1) ? K_IDENT : K_OPER; e.pos_src[1] = SRC_TI;
        kind[pw] = (mode == 1) ? K_INCH : K_OPER; e.pos_src[pw] = SRC_TW;
        kind[2 * n + 3] = K_IDENT; e.pos_src[2 * n + 3] = SRC_TF;
        if (mode == 2) { e.fixed_op[1] = m.corr_B[corr_idx]; e.fixed_op[pw] = m.corr_A[corr_idx]; }
    }
    // free positions, highest first, receive the sample's times in the order the transform
    // produces them; topology vertex v is the v-th of them (src/topology_eval.jl:269-275)
    int vertex_pos[2 * kMaxOrder + 1];
    {
        int v = 0;
        for (int pos = e.n_nodes; pos >= 1; --pos)
            if (kind[pos] == K_FREE) { e.pos_src[pos] = v; vertex_pos[++v] = pos; }
        if (v != 2 * n) { err = "internal: free position count"; return 1; }
    }

    const auto t_start = std::chrono::steady_clock::now();
    e.n_top = n_top;
    e.tree_off.clear(); e.tree_cost.clear();
    // Trees are grouped by initial sector (outer) so that an executor flushes its accumulator
    // rarely; inside a sector group the reference's topology order is kept.
    // The sector groups are independent of each other except for the numbering of the distinct coefficients and
    // pair-interaction slots (first seen, first numbered).  Big entries are therefore walked one sector group per
    // host thread into private fragments, which are then appended in sector order with their local numbers mapped to
    // the global first-seen numbering — the resulting program is identical, bit for bit, to the sequential one
    // (QIW_COMPILE_THREADS=1 forces the sequential walk; the host-logic tests compare both).
    int n_threads = (int)std::min<unsigned>((unsigned)m.S, std::max(1u, std::thread::hardware_concurrency()));
    {   // small entries: a thread costs more than the walk.  Work estimate: trees x (attachable pairs per sector)^order
        double attach = 0;
        for (int s = 0; s < m.S; ++s) attach += (double)m.attachable[s].size();
        attach = std::max(1.0, attach / std::max(1, m.S));
        double est = (double)n_top * m.S;
        for (int a = 0; a < order; ++a) est *= attach;
        if (est < 2e4) n_threads = 1;
    }
    if (const char* env = getenv("QIW_COMPILE_THREADS")) n_threads = std::max(1, std::min(atoi(env), m.S));   // an explicit request wins
    Builder b(m, e);
    if (n_threads <= 1) {
        for (int s_i = 0; s_i < m.S; ++s_i)
            if (b.walk_sector(s_i, n_top, pairs, parity, kind, vertex_pos, err)) return 1;
    } else {
        struct Fragment { EntryProgram e; std::unique_ptr<Builder> b; std::string err; int rc = 0; };
        std::vector<std::unique_ptr<Fragment>> frag(m.S);
        for (int s_i = 0; s_i < m.S; ++s_i) {
            frag[s_i].reset(new Fragment());
            EntryProgram& fe = frag[s_i]->e;
            fe.mode = e.mode; fe.order = e.order; fe.n_pts_after = e.n_pts_after; fe.corr_idx = e.corr_idx;
            fe.scalar = e.scalar; fe.D = e.D; fe.n_nodes = e.n_nodes; fe.nP = e.nP; fe.L = e.L; fe.RL = e.RL; fe.n_top = n_top;
            for (int p = 0; p <= kMaxNodes; ++p) { fe.pos_src[p] = e.pos_src[p]; fe.fixed_op[p] = e.fixed_op[p]; }
            frag[s_i]->b.reset(new Builder(m, fe));
        }
        auto work = [&](int w) {
            for (int s_i = w; s_i < m.S; s_i += n_threads)
                frag[s_i]->rc = frag[s_i]->b->walk_sector(s_i, n_top, pairs, parity, kind, vertex_pos, frag[s_i]->err);
        };
        std::vector<std::thread> pool;
        int started = 0;
        try {
            for (; started < n_threads - 1; ++started) pool.emplace_back(work, started);
        } catch (const std::system_error&) {}   // no more threads to be had: the caller does the rest itself
        for (int w = started; w < n_threads; ++w) work(w);
        for (auto& t : pool) t.join();
        for (int s_i = 0; s_i < m.S; ++s_i)
            if (frag[s_i]->rc) { err = frag[s_i]->err; return 1; }
        for (int s_i = 0; s_i < m.S; ++s_i) {
            const EntryProgram& fe = frag[s_i]->e;
            const Builder& fb = *frag[s_i]->b;
            if (fe.nP + (int)fe.dslots.size() > 4095 || fe.coefs.size() > 65535) { err = "program table overflow"; return 5; }
            std::vector<uint32_t> cmap(fe.coefs.size()), dmap(fe.dslots.size());
            for (size_t i = 0; i < fe.coefs.size(); ++i) cmap[i] = (uint32_t)b.coef_id(fe.coefs[i]);
            for (size_t i = 0; i < fe.dslots.size(); ++i)
                dmap[i] = (uint32_t)b.dslot_id(fe.dslots[i].pos_tail, fe.dslots[i].pos_head, fe.dslots[i].table);
            if (e.nP + (int)e.dslots.size() > 4095 || e.coefs.size() > 65535) { err = "program table overflow"; return 5; }
            const size_t base = e.words.size();
            for (uint32_t off : fe.tree_off) e.tree_off.push_back((uint32_t)(base + off));
            e.tree_cost.insert(e.tree_cost.end(), fe.tree_cost.begin(), fe.tree_cost.end());
            size_t next_root = 0;   // index into fe.tree_off: root words carry the initial sector in aux and are left alone
            for (size_t i = 0; i < fe.words.size(); ++i) {
                uint64_t w = fe.words[i];
                if (next_root < fe.tree_off.size() && fe.tree_off[next_root] == i) { ++next_root; e.words.push_back(w); continue; }
                const uint32_t slotB = (uint32_t)(w >> 12) & 0xFFFu, nchild = (uint32_t)(w >> 24) & 0xFFu;
                if (slotB) w = (w & ~((uint64_t)0xFFF << 12)) | ((uint64_t)((uint32_t)fe.nP + dmap[slotB - (uint32_t)fe.nP]) << 12);
                if (nchild == 0) w = (w & ~((uint64_t)0xFFFF << 32)) | ((uint64_t)cmap[(uint32_t)(w >> 32) & 0xFFFFu] << 32);
                e.words.push_back(w);
            }
            const size_t r0 = e.records.size();
            e.records.insert(e.records.end(), fe.records.begin(), fe.records.end());
            for (size_t r = r0; r < e.records.size(); r += (size_t)e.RL) {
                uint32_t* rec = e.records.data() + r;
                rec[0] = cmap[rec[0] & 0xFFFFu] | (rec[0] & 0xFFFF0000u);
                for (int q = 1; q <= e.L; ++q)
                    if ((int)rec[q] >= e.nP) rec[q] = (uint32_t)e.nP + dmap[rec[q] - (uint32_t)e.nP];
            }
            b.leaves += fb.leaves; b.edges += fb.edges; b.flops += fb.flops;
            b.offdiag = b.offdiag || fb.offdiag;
        }
    }
    if (b.offdiag) {
        err = "a block off-diagonal contribution to a pseudo-particle propagator detected "
              "(src/topology_eval.jl:462): pass more symmetry breakers to the ED";
        return 4;
    }
    e.tree_off.push_back((uint32_t)e.words.size());
    e.words.push_back(0);  // pad: executors prefetch two words ahead
    e.words.push_back(0);
    e.n_leaves = b.leaves; e.n_edges = b.edges; e.flops_per_sample = b.flops;
    if (e.nP + (int)e.dslots.size() > 4095 || e.coefs.size() > 65535) { err = "program table overflow"; return 5; }
    const auto t_walk = std::chrono::steady_clock::now();
    if (e.scalar) {
        int K = 0;
        if (const char* env = getenv("QIW_SEGMENTS")) K = atoi(env);
        factorise_records(e, m.S, K);
        if (e.nP + (int)e.dslots.size() + e.nSeg > 65535) { err = "segment table overflow"; return 5; }
    }
    if (getenv("QIW_COMPILE_TIMING")) {
        const std::chrono::duration<double, std::milli> walk = t_walk - t_start, fact = std::chrono::steady_clock::now() - t_walk;
        fprintf(stderr, "qiw compile: mode %d order %d k %d: %lld configurations, tree walk %.1f ms on %d thread(s), factorise %.1f ms\n",
                mode, order, n_pts_after, (long long)e.n_leaves, walk.count(), n_threads, fact.count());
    }
    return 0;
}

// Groups the propagator factors of every configuration into K segments of consecutive backbone
// intervals and tabulates the distinct sector sub-sequences per segment.  For 1x1 blocks all factors
// commute, so  weight = coef * prod_segments[prod_{iv in seg} iP(iv, s_iv)] * prod_arcs[i Delta]:
// the segment products are shared by all configurations that run through the same sectors on that
// stretch of the backbone (and they are many: the sector sequence is fixed by the operator sequence,
// whereas configurations also differ by which vertices the arcs connect).
void factorise_records(EntryProgram& e, int S, int K_req) {
    const int nI = e.n_nodes - 1, n = e.order, nP = e.nP, nD = (int)e.dslots.size();
    const int64_t nl = e.n_leaves;
    // candidate segmentations: K equal parts; keep the cheapest in shared-memory operations per sample (see `build`)
    const int Kmin = (nI + 8) / 9;   // the kernel handles segments of up to 9 intervals
    const int Kmax = std::max(Kmin, std::min(nI, 4));
    const int K_lo = K_req > 0 ? std::max(Kmin, std::min(K_req, nI)) : Kmin, K_hi = K_req > 0 ? K_lo : Kmax;
    struct Candidate { std::vector<uint32_t> rec; std::vector<uint16_t> def; int nseg = 0, stride = 0, n_lane_seg = 0; double cost = 1e300; };
    std::vector<Candidate> cand(K_hi - K_lo + 1);
    auto build = [&](int K, Candidate& c) {
        std::vector<int> bound(K + 1);
        for (int g = 0; g <= K; ++g) bound[g] = (int)((int64_t)g * nI / K);
        int stride = 0;
        for (int g = 0; g < K; ++g) stride = std::max(stride, bound[g + 1] - bound[g]);
        std::vector<std::map<std::vector<uint16_t>, int>> index(K);
        std::vector<uint16_t>& def = c.def;
        std::vector<uint32_t>& rec = c.rec;
        rec.assign((size_t)nl * (K + n + 1), 0u);
        int nseg = 0;
        for (int64_t l = 0; l < nl; ++l) {
            const uint32_t* src = e.records.data() + (size_t)l * e.RL;
            uint32_t* dst = rec.data() + (size_t)l * (K + n + 1);
            dst[0] = src[0];
            // split the flat factor list into propagator slots (in interval order) and interaction slots
            uint16_t ps[64];
            int np = 0, nd = 0;
            for (int q = 1; q <= e.L; ++q) {
                if ((int)src[q] < nP) ps[np++] = (uint16_t)src[q];
                else dst[1 + K + nd++] = src[q];
            }
            for (int g = 0; g < K; ++g) {
                std::vector<uint16_t> key(ps + bound[g], ps + bound[g + 1]);
                auto it = index[g].find(key);
                int id;
                if (it == index[g].end()) {
                    id = nseg++;
                    index[g][key] = id;
                    for (int i = 0; i < stride; ++i) def.push_back(i < (int)key.size() ? key[i] : (uint16_t)0xFFFF);
                } else id = it->second;
                dst[1 + g] = (uint32_t)(nP + nD + id);
            }
        }
        // Cost of the candidate in shared-memory operations per sample, as the step kernel executes it (lane program):
        // records of M = 4, 2, 1 configurations sharing their pair-interaction operands cost order + M K loads (+ 2 for
        // the record fetch), a segment-table entry its operands + 3 (definition, coefficient, store) — and the table
        // is rebuilt for every sample, so its entries weigh as much as the records.
        double rec_cost = 0, tab_cost = 0;
        {
            std::map<std::vector<uint32_t>, int> members;                 // (initial sector, sorted interaction slots) -> configurations
            std::map<std::pair<uint32_t, uint32_t>, int> first_seg;       // (coefficient, first segment entry): entries of their own
            std::vector<uint32_t> key;
            for (int64_t l = 0; l < nl; ++l) {
                const uint32_t* r = rec.data() + (size_t)l * (K + n + 1);
                key.assign(r + 1 + K, r + 1 + K + n);
                std::sort(key.begin(), key.end());
                key.push_back(r[0] >> 16);
                ++members[key];
                first_seg[std::make_pair(r[0] & 0xFFFFu, r[1])] = 1;
            }
            for (const auto& kv : members) {
                const int nm = kv.second;
                rec_cost += (nm / 4) * (n + 4 * K + 2) + ((nm % 4) / 2) * (n + 2 * K + 2) + (nm % 2) * (n + K + 2);
            }
            std::vector<char> is_first(nseg, 0);
            for (const auto& kv : first_seg) is_first[kv.first.second - (uint32_t)(nP + nD)] = 1;
            auto len_of = [&](int id) { int len = 0; for (int i = 0; i < stride; ++i) len += def[(size_t)id * stride + i] != 0xFFFF; return len; };
            for (int id = 0; id < nseg; ++id) if (!is_first[id]) tab_cost += len_of(id) + 3;
            for (const auto& kv : first_seg) tab_cost += len_of((int)(kv.first.second - (uint32_t)(nP + nD))) + 3;
            c.n_lane_seg = (int)first_seg.size();
            for (int id = 0; id < nseg; ++id) c.n_lane_seg += !is_first[id];
        }
        // slots per sample bound the CTA's shared memory: penalise tables beyond ~4 KB of doubles per sample
        const double row = nP + nD + c.n_lane_seg;
        c.cost = rec_cost + tab_cost + (row > 512 ? 1e6 * (row - 512) : 0.0);
        c.nseg = nseg; c.stride = stride;
    };
    // the candidates are independent: one host thread each when the entry is big enough to pay for it
    bool threaded = cand.size() > 1 && nl >= 4096 && std::thread::hardware_concurrency() > 1;
    if (const char* env = getenv("QIW_COMPILE_THREADS")) threaded = threaded && atoi(env) > 1;
    if (threaded) {
        std::vector<std::thread> pool;
        int K = K_lo;
        try {
            for (; K < K_hi; ++K) pool.emplace_back([&, K]() { build(K, cand[K - K_lo]); });
        } catch (const std::system_error&) {}   // no more threads to be had: the caller does the rest itself
        for (; K <= K_hi; ++K) build(K, cand[K - K_lo]);
        for (auto& t : pool) t.join();
    } else {
        for (int K = K_lo; K <= K_hi; ++K) build(K, cand[K - K_lo]);
    }
    int bestK = K_lo;
    for (int K = K_lo; K <= K_hi; ++K)          // first of the cheapest, as the sequential search picks
        if (cand[K - K_lo].cost < cand[bestK - K_lo].cost) bestK = K;
    std::vector<uint32_t>& best_rec = cand[bestK - K_lo].rec;
    std::vector<uint16_t>& best_def = cand[bestK - K_lo].def;
    const int best_nseg = cand[bestK - K_lo].nseg, best_stride = cand[bestK - K_lo].stride;
    e.K = bestK; e.L2 = bestK + n; e.nSeg = best_nseg; e.seg_stride = std::max(best_stride, 1);
    e.rec2.swap(best_rec); e.segdef.swap(best_def);
    if (e.segdef.empty()) e.segdef.assign(1, 0xFFFF);
    build_lane_program(e);
}

void build_lane_program(EntryProgram& e) {
    const int K = e.K, n = e.order, RL = e.L2 + 1, nP = e.nP, nD = (int)e.dslots.size(), stride = e.seg_stride;
    const int64_t nl = e.n_leaves;
    e.lane_sections.clear(); e.lane_items.clear(); e.lane_segdef.clear(); e.lane_seg_coef.clear(); e.nSegL = 0; e.lane_cost = 0;
    // 1. segment table with the coefficient folded into the first segment of every configuration: entries are
    //    (coefficient or none, entry of the plain table), numbered in order of first use
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> seg_index;
    auto seg_id = [&](uint32_t coef, uint32_t plain) {
        auto key = std::make_pair(coef, plain);
        auto it = seg_index.find(key);
        if (it != seg_index.end()) return it->second;
        const uint32_t id = (uint32_t)e.nSegL++;
        seg_index[key] = id;
        const uint16_t* def = e.segdef.data() + (size_t)(plain - (uint32_t)(nP + nD)) * stride;
        e.lane_segdef.insert(e.lane_segdef.end(), def, def + stride);
        e.lane_seg_coef.push_back((uint16_t)coef);
        return id;
    };
    // 2. groups: sorted Delta slots -> members by initial sector, in order of first appearance.  Configurations that
    //    differ only in the initial sector (spin-symmetric sectors, typically) share all pair-interaction operands.
    struct Group { std::vector<uint32_t> dsl; std::map<uint32_t, std::vector<uint32_t>> members; };   // sector -> K slots per member
    std::vector<Group> groups;
    std::map<std::vector<uint32_t>, size_t> gindex;
    std::vector<uint32_t> key;
    for (int64_t l = 0; l < nl; ++l) {
        const uint32_t* r = e.rec2.data() + (size_t)l * RL;
        key.assign(r + 1 + K, r + 1 + K + n);
        std::sort(key.begin(), key.end());
        auto it = gindex.find(key);
        size_t g;
        if (it == gindex.end()) {
            g = groups.size();
            gindex[key] = g;
            groups.emplace_back();
            groups[g].dsl = key;
        } else g = it->second;
        std::vector<uint32_t>& mem = groups[g].members[r[0] >> 16];
        for (int q = 0; q < K; ++q)
            mem.push_back((uint32_t)(nP + nD) + seg_id(q == 0 ? (r[0] & 0xFFFFu) : 0xFFFFu, r[1 + q]));
    }
    // 3. fixed-shape records.  Per group and sector: records of four members; of the remainder, a unit of two and/or a
    //    unit of one.  Units of equal size from two sectors of the same group share ONE record (first half of the members
    //    -> sector a, second half -> sector b: the Delta operands are loaded once for both), what is left over becomes a
    //    record of its own.  Sections = (members per record, sector a[, sector b]).  Shared records are used from order 5
    //    on: 16 % fewer operand loads and a third fewer records (orders 0:6 at N = 2^14: 21.0 -> 17.8 ms), but two more
    //    record shapes per entry, and at order <= 4 a warp runs so few records per shape that the additional code
    //    costs more than the loads save (README run 3.52 -> 3.9 ms).  QIW_LANE_DUAL=0 / 1 overrides.
    bool dual_ok = n >= 5;
    if (const char* env = getenv("QIW_LANE_DUAL")) dual_ok = env[0] != '0';
    struct SecKey { int cls; int s_a, s_b; bool operator<(const SecKey& o) const { return std::tie(cls, s_a, s_b) < std::tie(o.cls, o.s_a, o.s_b); } };
    std::map<SecKey, std::vector<uint16_t>> sec_items;     // records of every section, in group order
    auto emit = [&](int cls, int M, int s_a, int s_b, const Group& g, const uint32_t* ma, const uint32_t* mb) {
        std::vector<uint16_t>& v = sec_items[SecKey{cls, s_a, s_b}];
        const size_t base = v.size();
        for (uint32_t d : g.dsl) v.push_back((uint16_t)d);
        const int half = (s_b < 0) ? M : M / 2;
        for (int q = 0; q < half * K; ++q) v.push_back((uint16_t)ma[q]);
        if (s_b >= 0) for (int q = 0; q < half * K; ++q) v.push_back((uint16_t)mb[q]);
        v.resize(base + lane_record_items(n, K, M), 0);
    };
    for (const Group& g : groups) {
        std::vector<std::pair<int, const uint32_t*>> twos, ones;     // (sector, first member) of the remainder units
        for (const auto& kv : g.members) {
            const int s = (int)kv.first, nm = (int)(kv.second.size() / std::max(K, 1));
            const uint32_t* m = kv.second.data();
            for (int c = 0; c < nm / 4; ++c) emit(0, 4, s, -1, g, m + (size_t)c * 4 * K, nullptr);
            int at = (nm / 4) * 4;
            if (nm - at >= 2) { twos.push_back(std::make_pair(s, m + (size_t)at * K)); at += 2; }
            if (nm - at >= 1) ones.push_back(std::make_pair(s, m + (size_t)at * K));
        }
        size_t k = 0;
        if (dual_ok) for (; k + 1 < twos.size(); k += 2) emit(1, 4, twos[k].first, twos[k + 1].first, g, twos[k].second, twos[k + 1].second);
        for (; k < twos.size(); ++k) emit(2, 2, twos[k].first, -1, g, twos[k].second, nullptr);
        k = 0;
        if (dual_ok) for (; k + 1 < ones.size(); k += 2) emit(3, 2, ones[k].first, ones[k + 1].first, g, ones[k].second, ones[k + 1].second);
        for (; k < ones.size(); ++k) emit(4, 1, ones[k].first, -1, g, ones[k].second, nullptr);
    }
    const int cls_M[5] = {4, 4, 2, 2, 1};
    for (const auto& kv : sec_items) {
        EntryProgram::LaneSection sec;
        sec.s_i = kv.first.s_a; sec.s_b = kv.first.s_b; sec.M = cls_M[kv.first.cls]; sec.rec0 = 0;
        sec.n_rec = (uint32_t)(kv.second.size() / (size_t)lane_record_items(n, K, sec.M));
        sec.chunk0 = (uint32_t)(e.lane_items.size() / 8);
        sec.cost = (uint32_t)(n + sec.M * K + 2);
        e.lane_items.insert(e.lane_items.end(), kv.second.begin(), kv.second.end());
        e.lane_sections.push_back(sec);
        e.lane_cost += (int64_t)sec.n_rec * sec.cost;
    }
    uint32_t rec0 = 0;
    for (auto& sec : e.lane_sections) { sec.rec0 = rec0; rec0 += sec.n_rec; }
    if (e.lane_segdef.empty()) e.lane_segdef.assign(std::max(stride, 1), 0xFFFF);
}

}  // namespace qiw
