// qiw_seq.cpp — host-side integer work of the hot path: Sobol direction numbers and scrambling,
// topology enumeration, sample partitioning, spline preparation.
//
// Reference behaviour reproduced bit-exactly (checked against the reference's golden tables in
// tests/): src/scrambled_sobol.jl:66-143 (constructor), src/diagrammatics.jl:288-337
// (generate_topologies / get_topologies_at_order), src/utility.jl:164-179 (split_count).
#include "qiw_host.hpp"

#include <cstring>

#include "joe_kuo_table.h"

namespace qiw {

// ScrambledSobolSeq(D) without scrambling: m[d][j] = direction integer j of dimension d, left
// aligned so that bit 31 is the first binary digit (src/scrambled_sobol.jl:69-100).
int sobol_direction_numbers(int D, uint32_t* m) {
    if (D < 0 || D > QIW_JK_NDIM) return 1;
    for (int d = 0; d < D; ++d) {
        uint32_t v[32];
        if (d == 0) {
            for (int j = 0; j < 32; ++j) v[j] = 1u;
        } else {
            const uint32_t poly = qiw_jk_a[d - 1];
            int deg = 0;
            while ((poly >> (deg + 1)) != 0) ++deg;
            for (int j = 0; j < deg; ++j) v[j] = qiw_jk_minit[d - 1][j];
            for (int j = deg; j < 32; ++j) {
                // m_j = m_{j-deg} xor sum_k a_k (m_{j-deg+k} << (deg-k)), a_k = bit k of poly
                uint32_t acc = v[j - deg];
                for (int k = 0; k < deg; ++k)
                    if ((poly >> k) & 1u) acc ^= v[j - deg + k] << (deg - k);
                v[j] = acc;
            }
        }
        for (int j = 0; j < 32; ++j) m[d * 32 + j] = v[j] << (31 - j);
    }
    return 0;
}

// Linear matrix scramble + digital shift (src/scrambled_sobol.jl:104-139).  The random bits are an
// input so that the host language's RNG stream stays the host's business.
int sobol_scramble(int D, uint32_t* m, uint32_t* x0, const uint8_t* shift_bits, const uint8_t* ltm_bits) {
    for (int d = 0; d < D; ++d) {
        uint32_t shift = 0;
        for (int j = 0; j < 32; ++j) shift |= (uint32_t)(shift_bits[d + (size_t)D * j] & 1u) << j;
        x0[d] = shift;
        // rows of the unit lower-triangular matrix as bit masks over input digits; digit q
        // (q = 0 is the most significant) is bit 31-q of a direction integer
        uint32_t rowmask[32];
        for (int p = 0; p < 32; ++p) {
            uint32_t mask = 1u << (31 - p);  // unit diagonal
            for (int q = 0; q < p; ++q)
                if (ltm_bits[d + (size_t)D * (p + 32 * (size_t)q)] & 1u) mask |= 1u << (31 - q);
            rowmask[p] = mask;
        }
        for (int j = 0; j < 32; ++j) {
            const uint32_t in = m[d * 32 + j];
            uint32_t out = 0;
            for (int p = 0; p < 32; ++p)
                out |= (uint32_t)(__builtin_popcount(rowmask[p] & in) & 1) << (31 - p);
            m[d * 32 + j] = out;
        }
    }
    return 0;
}

// ---- topologies -------------------------------------------------------------------------------

namespace {

struct TopoGen {
    int order, k;
    bool filter, external;
    int32_t* pairs;
    int32_t* parity;
    int64_t count = 0;
    int a[kMaxOrder + 2], b[kMaxOrder + 2];

    static bool cross(int a1, int b1, int a2, int b2) {  // a < b for both arcs
        return (a1 < a2 && a2 < b1 && b1 < b2) || (a2 < a1 && a1 < b2 && b2 < b1);
    }

    // every connected component of the crossing graph must contain an arc straddling k
    bool doubly_k_connected() const {
        uint32_t reached = 0, frontier = 0;
        for (int i = 0; i < order; ++i)
            if (a[i] <= k && b[i] > k) frontier |= 1u << i;
        reached = frontier;
        while (frontier) {
            int c = __builtin_ctz(frontier);
            frontier &= frontier - 1;
            for (int i = 0; i < order; ++i)
                if (!((reached >> i) & 1u) && cross(a[c], b[c], a[i], b[i])) {
                    reached |= 1u << i;
                    frontier |= 1u << i;
                }
        }
        return reached == ((order == 32) ? 0xFFFFFFFFu : ((1u << order) - 1u));
    }

    void emit(int sign) {
        if (filter && !doubly_k_connected()) return;
        if (pairs) {
            for (int i = 0; i < order; ++i) {
                pairs[(count * order + i) * 2] = a[i];
                pairs[(count * order + i) * 2 + 1] = b[i];
            }
        }
        if (parity) parity[count] = (external && (k & 1)) ? -sign : sign;
        ++count;
    }

    // `free_mask`: bit v-1 set = vertex v unpaired.  Pair the lowest free vertex with each other
    // free vertex in increasing order; the r-th candidate (r = 2, 3, ... counting the lowest vertex
    // as 1) multiplies the parity by (-1)^r  (src/diagrammatics.jl:288-302).
    void rec(uint32_t free_mask, int depth, int sign) {
        if (!free_mask) { emit(sign); return; }
        const int v = __builtin_ctz(free_mask);
        uint32_t rest = free_mask & (free_mask - 1);
        int r = 2;
        for (uint32_t it = rest; it; it &= it - 1, ++r) {
            const int u = __builtin_ctz(it);
            a[depth] = v + 1;
            b[depth] = u + 1;
            rec(rest & ~(1u << u), depth + 1, (r & 1) ? -sign : sign);
        }
    }
};

}  // namespace

int64_t enumerate_topologies(int order, int k, bool external_arc, int32_t* pairs, int32_t* parity) {
    if (order < 0 || order > kMaxOrder) return -1;
    TopoGen g;
    g.order = order; g.k = k; g.filter = (k >= 0); g.external = external_arc && (k >= 0);
    g.pairs = pairs; g.parity = parity;
    const uint32_t all = (order == 0) ? 0u : ((2 * order == 32) ? 0xFFFFFFFFu : ((1u << (2 * order)) - 1u));
    g.rec(all, 0, 1);
    return g.count;
}

// split_count + range_from_chunks_and_idx: the first N % n ranks own one extra sample.
void rank_sub_range(uint64_t N, int n_ranks, int rank, uint64_t* start, uint64_t* count) {
    const uint64_t q = N / (uint64_t)n_ranks, r = N % (uint64_t)n_ranks, k = (uint64_t)rank;
    *start = k * q + (k < r ? k : r);
    *count = q + (k < r ? 1 : 0);
}

// Natural cubic spline (zero second derivative at both ends) through equidistant knots: the
// interpolant Interpolations.jl builds for BSpline(Cubic(Line(OnGrid()))) (src/spline_gf.jl:193-201).
void natural_spline_second_derivatives(int n, double h, const cplx* y, cplx* M) {
    for (int j = 0; j < n; ++j) M[j] = cplx(0.0, 0.0);
    if (n < 3) return;
    const int u = n - 2;  // unknowns M_1 .. M_{n-2}
    std::vector<double> cp(u);
    std::vector<cplx> dp(u);
    const double s = 6.0 / (h * h);
    for (int j = 0; j < u; ++j) {
        const cplx rhs = s * (y[j] - 2.0 * y[j + 1] + y[j + 2]);
        const double denom = (j == 0) ? 4.0 : 4.0 - cp[j - 1];
        cp[j] = 1.0 / denom;
        dp[j] = (j == 0) ? rhs / denom : (rhs - dp[j - 1]) / denom;
    }
    M[u] = dp[u - 1];
    for (int j = u - 2; j >= 0; --j) M[j + 1] = dp[j] - cp[j] * M[j + 2];
}

}  // namespace qiw
