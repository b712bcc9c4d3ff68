// qiw_scalar.cu — the kernels of libqinchworm_cuda.so for models whose sector blocks are all 1x1 (sm_100a):
//   scalar_step_kernel  one launch evaluates every (entry, sample block) of a call: scrambled-Sobol points, simplex
//                       maps, P / Delta interpolation, segment products and the configuration sums, followed by the
//                       deterministic reduction, the peer-memory all-reduce and (device-resident loop) set_ppgf! +
//                       normalize! in the last CTA;
//   scalar_run_kernel   all bold steps of qiw_inchworm_run in ONE cooperative launch when a step is too small to fill
//                       the machine: one job per CTA for the whole run, P and the pair-interaction tables staged in
//                       shared memory, one grid barrier per step.
// Both are made of the same phases (phase_roots ... phase_walk below).
//
// Mapping: LANE = SAMPLE.  A CTA owns (entry, block of 32 m samples[, part of the entry's lane program][, z]);
// the per-sample operand table lives in shared memory as T[slot][sample], so that a warp reading one slot for
// its 32 samples touches 32 consecutive words (no bank conflicts) and every record of the lane program
// (EntryProgram::lane_*, qiw_host.hpp) is warp-uniform: its slot numbers arrive by one broadcast 128-bit load
// per 8 operands and the per-sector sum is carried in one register per lane, reduced over the lanes once per
// run of records instead of once per 32 configurations.
//
// Reference: src/topology_eval.jl:350-437,454-556 (per-sample evaluation), src/qmc_integrate.jl:363-463,
// 497-507,597-612 (transforms, integral), src/scrambled_sobol.jl:158-197 (points), src/mpi.jl:104-127,
// src/inchworm.jl:474-493 and src/ppgf.jl:495-504,646-668 (the loop over steps and its state update).
#include <cstdio>

#include "qiw_devfn.cuh"

namespace qiw {

// ---- configuration sums (lane program) ---------------------------------------------------------------
// One record = one group of M configurations sharing all pair-interaction operands and the initial sector:
//     prod_{f < ND} T[d_f]  *  sum_{m < M} prod_{q < K} T[s_{m,q}]
// (the coefficient of member m is folded into its first segment product).  ND + M K operand loads and about
// as many FP64 operations per record and sample; the record words of the next record are fetched while the
// current one is evaluated.  `Tl` = table base + this lane's column.
template <int ND, int K, int M, bool REAL>
__device__ __noinline__ typename Num<REAL>::T lane_walk(const uint4* rec, int n_rec, const unsigned char* Tl,
                                                        unsigned row_bytes) {
    typedef typename Num<REAL>::T T;
    typedef Num<REAL> N;
    constexpr int NI = ND + M * K, NC = (NI + 7) / 8;       // 128-bit words per record: eight 16-bit slot numbers each
    T acc = N::zero();
    uint4 nx[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) nx[c] = rec[c];        // generic loads: the records live in global or shared memory
    for (int r = 0; r < n_rec; ++r) {
        uint32_t wd[NC * 4];
#pragma unroll
        for (int c = 0; c < NC; ++c) { wd[4 * c] = nx[c].x; wd[4 * c + 1] = nx[c].y; wd[4 * c + 2] = nx[c].z; wd[4 * c + 3] = nx[c].w; }
        rec += NC;   // the item array is padded: the fetch past the last record stays inside it
#pragma unroll
        for (int c = 0; c < NC; ++c) nx[c] = rec[c];
        uint32_t it[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) it[i] = (i & 1) ? (wd[i >> 1] >> 16) : (wd[i >> 1] & 0xFFFFu);
        T v[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) v[i] = *reinterpret_cast<const T*>(Tl + it[i] * row_bytes);
        T sum = v[ND];
#pragma unroll
        for (int q = 1; q < K; ++q) sum = N::mul(sum, v[ND + q]);
#pragma unroll
        for (int m = 1; m < M; ++m) {
            T s = v[ND + m * K];
#pragma unroll
            for (int q = 1; q < K; ++q) s = N::mul(s, v[ND + m * K + q]);
            sum = N::add(sum, s);
        }
        if constexpr (ND > 0) {
            T d = v[0];
#pragma unroll
            for (int f = 1; f < ND; ++f) d = N::mul(d, v[f]);
            sum = N::mul(d, sum);
        }
        acc = N::add(acc, sum);
    }
    return acc;
}

// shape code = order * 16 + (K - 1) * 4 + (0, 1, 2 for M = 1, 2, 4); order <= 8, K <= 4
template <bool REAL>
__device__ __forceinline__ typename Num<REAL>::T lane_dispatch(unsigned code, const uint4* rec, int n_rec, const unsigned char* Tl,
                                                               unsigned row_bytes) {
#define QIW_LW(D_, K_, MC_, M_) case (D_ * 16 + (K_ - 1) * 4 + MC_): return lane_walk<D_, K_, M_, REAL>(rec, n_rec, Tl, row_bytes);
#define QIW_LK(D_, K_) QIW_LW(D_, K_, 0, 1) QIW_LW(D_, K_, 1, 2) QIW_LW(D_, K_, 2, 4)
#define QIW_LD(D_) QIW_LK(D_, 1) QIW_LK(D_, 2) QIW_LK(D_, 3) QIW_LK(D_, 4)
    switch (code) {
        QIW_LD(0) QIW_LD(1) QIW_LD(2) QIW_LD(3) QIW_LD(4) QIW_LD(5) QIW_LD(6) QIW_LD(7) QIW_LD(8)
        default: break;
    }
#undef QIW_LD
#undef QIW_LK
#undef QIW_LW
    return Num<REAL>::zero();
}

// Records shared by two initial sectors (LaneRun::z): the first M / 2 members are summed for sector a, the others for
// sector b, and both sums take the record's product of pair-interaction operands, which is loaded once for the two.
template <bool REAL> struct Acc2 { typename Num<REAL>::T a, b; };

template <int ND, int K, int M, bool REAL>
__device__ __noinline__ Acc2<REAL> lane_walk2(const uint4* rec, int n_rec, const unsigned char* Tl, unsigned row_bytes) {
    typedef typename Num<REAL>::T T;
    typedef Num<REAL> N;
    constexpr int NI = ND + M * K, NC = (NI + 7) / 8, H = M / 2;
    Acc2<REAL> acc;
    acc.a = N::zero(); acc.b = N::zero();
    uint4 nx[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) nx[c] = rec[c];
    for (int r = 0; r < n_rec; ++r) {
        uint32_t wd[NC * 4];
#pragma unroll
        for (int c = 0; c < NC; ++c) { wd[4 * c] = nx[c].x; wd[4 * c + 1] = nx[c].y; wd[4 * c + 2] = nx[c].z; wd[4 * c + 3] = nx[c].w; }
        rec += NC;
#pragma unroll
        for (int c = 0; c < NC; ++c) nx[c] = rec[c];
        uint32_t it[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) it[i] = (i & 1) ? (wd[i >> 1] >> 16) : (wd[i >> 1] & 0xFFFFu);
        T v[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) v[i] = *reinterpret_cast<const T*>(Tl + it[i] * row_bytes);
        T sum[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            T sm = v[ND + h * H * K];
#pragma unroll
            for (int q = 1; q < K; ++q) sm = N::mul(sm, v[ND + h * H * K + q]);
#pragma unroll
            for (int m = 1; m < H; ++m) {
                T s = v[ND + (h * H + m) * K];
#pragma unroll
                for (int q = 1; q < K; ++q) s = N::mul(s, v[ND + (h * H + m) * K + q]);
                sm = N::add(sm, s);
            }
            sum[h] = sm;
        }
        if constexpr (ND > 0) {
            T d = v[0];
#pragma unroll
            for (int f = 1; f < ND; ++f) d = N::mul(d, v[f]);
            sum[0] = N::mul(d, sum[0]);
            sum[1] = N::mul(d, sum[1]);
        }
        acc.a = N::add(acc.a, sum[0]);
        acc.b = N::add(acc.b, sum[1]);
    }
    return acc;
}

template <bool REAL>
__device__ __forceinline__ Acc2<REAL> lane_dispatch2(unsigned code, const uint4* rec, int n_rec, const unsigned char* Tl, unsigned row_bytes) {
#define QIW_LW(D_, K_, MC_, M_) case (D_ * 16 + (K_ - 1) * 4 + MC_): return lane_walk2<D_, K_, M_, REAL>(rec, n_rec, Tl, row_bytes);
#define QIW_LK(D_, K_) QIW_LW(D_, K_, 1, 2) QIW_LW(D_, K_, 2, 4)
#define QIW_LD(D_) QIW_LK(D_, 1) QIW_LK(D_, 2) QIW_LK(D_, 3) QIW_LK(D_, 4)
    switch (code) {
        QIW_LD(0) QIW_LD(1) QIW_LD(2) QIW_LD(3) QIW_LD(4) QIW_LD(5) QIW_LD(6) QIW_LD(7) QIW_LD(8)
        default: break;
    }
#undef QIW_LD
#undef QIW_LK
#undef QIW_LW
    Acc2<REAL> z;
    z.a = Num<REAL>::zero(); z.b = Num<REAL>::zero();
    return z;
}

// ---- segment products: T[nP + nD + j][sample] = [coef_j *] prod_i T[def_j[i]][sample] ----------------
// One warp per (table entry, 32 samples), lane = sample.  A definition is warp-uniform and arrives by one broadcast
// 128-bit load: eight 16-bit fields — the folded coefficient's index (0xFFFF = none), then the propagator slots
// (0xFFFF = unused); definitions of more than seven slots take a second word.  Four entries are in flight per warp
// (the chain definition -> operands -> product -> store is pure latency).  `defs` and `segcoef` (the coefficients
// by table entry, in the kernel's arithmetic) may live in shared memory (run kernel) or `segcoef` be null.
template <int STRIDE, bool REAL>
__device__ __forceinline__ void segment_products(const DevEntry& e, const uint4* defs, const typename Num<REAL>::T* segcoef,
                                                 unsigned char* Tb, unsigned row_bytes, int warp, int nw, int sub_sh, int lane) {
    typedef typename Num<REAL>::T T;
    typedef Num<REAL> N;
    constexpr int NW = STRIDE > 7 ? 2 : 1, U = 4;
    const int base = e.nP + e.nD, total = e.nSegL << sub_sh;        // units = (table entry, sample sub-block of 32)
    for (int u0 = warp; u0 < total; u0 += U * nw) {
        uint32_t w[U][4 * NW];
        int jj[U]; unsigned col[U];
#pragma unroll
        for (int b = 0; b < U; ++b) {
            const int u = min(u0 + b * nw, total - 1);      // the tail repeats the last unit (same value stored twice)
            jj[b] = u >> sub_sh;
            col[b] = (unsigned)((u - (jj[b] << sub_sh)) * 32 + lane) * (unsigned)sizeof(T);
            const uint4 a = defs[(size_t)jj[b] * NW];
            w[b][0] = a.x; w[b][1] = a.y; w[b][2] = a.z; w[b][3] = a.w;
            if constexpr (NW > 1) { const uint4 c2 = defs[(size_t)jj[b] * NW + 1]; w[b][4] = c2.x; w[b][5] = c2.y; w[b][6] = c2.z; w[b][7] = c2.w; }
        }
        T v[U];
#pragma unroll
        for (int b = 0; b < U; ++b) {
            auto field = [&](int k) -> uint32_t { return (k & 1) ? (w[b][k >> 1] >> 16) : (w[b][k >> 1] & 0xFFFFu); };
            v[b] = *reinterpret_cast<const T*>(Tb + field(1) * row_bytes + col[b]);   // a segment holds at least one interval
#pragma unroll
            for (int i = 1; i < STRIDE; ++i) {
                const uint32_t q = field(1 + i);
                if (q != 0xFFFFu) v[b] = N::mul(v[b], *reinterpret_cast<const T*>(Tb + q * row_bytes + col[b]));
            }
            const uint32_t ci = field(0);
            if (ci != 0xFFFFu) v[b] = N::mul(segcoef ? segcoef[jj[b]] : N::coef_of(__ldg(e.coefs + ci)), v[b]);
        }
#pragma unroll
        for (int b = 0; b < U; ++b) *reinterpret_cast<T*>(Tb + (unsigned)(base + jj[b]) * row_bytes + col[b]) = v[b];
    }
}

// ---- per-CTA view of one job and the phases both kernels are made of -------------------------------------
// A job = (entry, `ns` consecutive samples, chunks of the entry's lane program).  ns is a power of two: <= 32 in the
// step kernel; 32 m in the persistent run kernel, where a CTA takes m sample blocks of a light entry at once.
template <bool REAL>
struct Cta {
    const DevEntry* e;
    int ns, ns_sh;                // samples of the job, log2
    unsigned row_bytes;           // ns * sizeof(T): one row of T[slot][sample]
    unsigned char* Tb;            // operand table
    double* pw;                   // [D][ns] simplex roots (aliases the start of T in the step kernel)
    double* times;                // [n_nodes + 1][ns]
    double* cellw;                // [n_nodes + 1][ns] fractional weight of every time inside its P-grid cell
    int* cella;                   // [n_nodes + 1][ns] P-grid cell of every time
    int* okflag;                  // [ns]
    const uint32_t* dslots_s;     // [nD] pos_tail | pos_head << 8 | table << 16
    double2* red;                 // [S][nw] per-warp sums
    // the job's part of the lane program: global memory, or the run kernel's on-chip copy
    const uint4* items;           // records
    const LaneRun* runs;          // runs of records of equal shape and initial sector
    const uint32_t* chunk_off;    // [..] first run of every chunk; the job's chunks are chunk0 .. chunk0 + n_chunks - 1
    int chunk0, n_chunks;
    const uint4* segdefs;         // packed definitions of the segment-product table
    const typename Num<REAL>::T* segcoef;   // coefficient of every table entry, or null (then read through DevEntry::coefs)
    int lane, warp, nw;
    int tid, nthr;                // this thread's index among the `nthr` threads that execute a phase together
};

// -- 1. Sobol coordinates and the independent roots x_j^(1/(remaining dims)).  The roots depend only on (entry, Sobol
//       sequence, sample), not on the time step: with a cache they are computed once per run and re-read afterwards.
template <bool REAL>
__device__ __forceinline__ void phase_roots(const Cta<REAL>& c, const uint32_t* __restrict__ sm, unsigned long long start,
                                            unsigned long long local0, unsigned long long count, double* uc, bool uc_valid,
                                            bool uc_write) {
    const DevEntry& e = *c.e;
    const int D = e.D, d_after = e.d_after, ns = c.ns;
    for (int smp = c.tid; smp < ns; smp += c.nthr) c.okflag[smp] = (local0 + smp < count) ? 1 : 0;
    for (int task = c.tid; task < D * ns; task += c.nthr) {
        const int j = task >> c.ns_sh, smp = task & (ns - 1);
        const unsigned long long local = local0 + smp;
        const bool active = local < count;
        double r;
        if (uc && uc_valid) {
            r = active ? uc[(size_t)j * count + local] : 0.0;
        } else {
            const uint32_t xi = sobol_coord(sm + j * 32, __ldg(sm + D * 32 + j), (uint32_t)(start + local));
            const double x = (double)xi * 2.3283064365386963e-10;  // ldexp(x, -32), exact
            const int den = (j < d_after) ? (d_after - j) : (D - j);
            r = (den == 1) ? x : pow(x, 1.0 / (double)den);
            if (uc && active && uc_write) uc[(size_t)j * count + local] = r;
        }
        c.pw[j * ns + smp] = r;
    }
}

// -- 2. ordered times of every backbone position: thread = (position, sample).  The running product
//       u_j = ((r_0 r_1) r_2) ... r_j is re-evaluated from the start of its simplex so that positions are independent
//       (same operation order as the sequential map, bit-identical); grid cell and weight of every time on the P grid.
template <bool REAL>
__device__ __forceinline__ void phase_times(const Cta<REAL>& c, double t_i, double t_w, double t_f, int n_tau, double inv_h,
                                            const double* __restrict__ explicit_times, unsigned long long local0,
                                            unsigned long long count) {
    const DevEntry& e = *c.e;
    const int D = e.D, d_after = e.d_after, ns = c.ns, n_nodes = e.n_nodes;
    const double lo_after = (e.mode == 0) ? t_i : t_w, len_after = t_f - lo_after;
    const double len_before = t_w - t_i;
    for (int task = c.tid; task < n_nodes * ns; task += c.nthr) {
        const int pos = 1 + (task >> c.ns_sh), smp = task & (ns - 1);
        const int src = e.pos_src[pos];
        double t;
        if (src == -1) t = t_i;
        else if (src == -2) t = t_w;
        else if (src == -3) t = t_f;
        else if (explicit_times) {
            t = (local0 + smp < count) ? explicit_times[(local0 + smp) * D + src] : 0.0;
        } else {
            const int j0 = (src < d_after) ? 0 : d_after;
            double u = c.pw[j0 * ns + smp];
            for (int j = j0 + 1; j <= src; ++j) u = __dmul_rn(u, c.pw[j * ns + smp]);
            if (src < d_after) t = __dadd_rn(__dmul_rn(u, len_after), lo_after);
            else t = __dadd_rn(__dmul_rn(u, len_before), t_i);
            if (!(t >= 0.0)) c.okflag[smp] = 0;   // all(refs .>= 0) (src/qmc_integrate.jl:608)
        }
        c.times[pos * ns + smp] = t;
        // grid cell and weight of this time on the P grid (Keldysh.jl rule, shared by all slots)
        const double q = t * inv_h;
        const int a = min(max(__double2int_rd(q), 0), n_tau - 2);
        c.cella[pos * ns + smp] = a;
        c.cellw[pos * ns + smp] = q - (double)a;
    }
}

// Where the propagator tables are read from.  Step kernel: the tables in global memory as the host stores them
// (complex, P[k][s]); run kernel: per-CTA copies in shared memory that already hold i * value in the kernel's
// arithmetic (T), pair-interaction table t at Ds + t * n_tau.
template <bool REAL>
struct StagedTables {
    const typename Num<REAL>::T* Ps;     // [n_tau][S]  i P
    const typename Num<REAL>::T* Ds;     // pair-interaction tables, i * value: table t at Ds + Doff[t]; grid values, then (splines) second derivatives
    const int* Doff;                     // null: every table is a plain function on the P grid, table t at Ds + t * n_tau
};

template <bool REAL>
__device__ __forceinline__ typename Num<REAL>::T cell3_apply_staged(const typename Num<REAL>::T* __restrict__ D, int stride, const GridCell3& c) {
    typedef typename Num<REAL>::T T;
    const T dk = D[c.i0 * stride], dp = D[c.ip * stride], dm = D[c.im * stride];
    if constexpr (REAL) return c.ck * dk + c.cp * dp + c.cm * dm;
    else return make_double2(c.ck * dk.x + c.cp * dp.x + c.cm * dm.x, c.ck * dk.y + c.cp * dp.y + c.cm * dm.y);
}

// -- 3a. pair-interaction slots: thread = (slot, sample).  Discarded samples (src/qmc_integrate.jl:503,608) and
//        samples past the range get zero columns.
template <bool REAL, bool STAGED>
__device__ __forceinline__ void phase_fill_delta(const Cta<REAL>& c, const StepParams& p, const StagedTables<REAL>& st) {
    typedef typename Num<REAL>::T T;
    typedef Num<REAL> N;
    const DevEntry& e = *c.e;
    const int nP = e.nP, nD = e.nD, ns = c.ns;
    T* Tt = reinterpret_cast<T*>(c.Tb);
    const double* times = c.times; const double* cellw = c.cellw; const int* cella = c.cella;
    if (p.tables_on_grid) {   // every Delta table is a plain grid function on the P grid: branch-free, cells reused
#pragma unroll 4
        for (int task = c.tid; task < nD * ns; task += c.nthr) {
            const int q = task >> c.ns_sh, smp = task & (ns - 1);
            const uint32_t dsw = c.dslots_s[q];
            const int3 ds = make_int3((int)(dsw & 0xFFu), (int)((dsw >> 8) & 0xFFu), (int)(dsw >> 16));
            const int ih = ds.y * ns + smp;
            const int it2 = (times[ds.x * ns + smp] <= times[ih]) ? ih : ds.x * ns + smp;   // clamp (:407-410)
            const GridCell3 cell = grid_cell3_from(cella[it2], cellw[it2], cella[ih], cellw[ih]);
            T val;
            if constexpr (STAGED) val = cell3_apply_staged<REAL>(st.Ds + ds.z * p.n_tau, 1, cell);
            else val = cell3_apply_i<REAL>(p.deltas_inline[ds.z].y, 1, cell);
            Tt[(nP + q) * ns + smp] = c.okflag[smp] ? val : N::zero();
        }
    } else if constexpr (STAGED) {
        // staged tables of any kind: spline-interpolated functions (natural cubic spline in t_f - t_i,
        // src/spline_gf.jl:208-219) and grid functions on a grid of their own
        for (int task = c.tid; task < nD * ns; task += c.nthr) {
            const int q = task >> c.ns_sh, smp = task & (ns - 1);
            const uint32_t dsw = c.dslots_s[q];
            const int3 ds = make_int3((int)(dsw & 0xFFu), (int)((dsw >> 8) & 0xFFu), (int)(dsw >> 16));
            const double th = times[ds.y * ns + smp];
            double tt = times[ds.x * ns + smp];
            if (tt < th) tt = th;                       // :407-410
            const DevDelta& dt = p.deltas_inline[ds.z];
            const T* tab = st.Ds + st.Doff[ds.z];
            T val;
            if (dt.kind == 1) {
                const double d_t = tt - th, h = dt.h;
                int j = (int)floor(d_t * dt.inv_h);
                j = min(max(j, 0), dt.n - 2);
                const double xa = d_t - (double)j * h, xb = (double)(j + 1) * h - d_t;
                const T y0 = tab[j], y1 = tab[j + 1], m0 = tab[dt.n + j], m1 = tab[dt.n + j + 1];
                const double i6h = dt.inv_h * (1.0 / 6.0), h6 = h * (1.0 / 6.0), ih = dt.inv_h;
                const double ca = xa * xa * xa * i6h, cb = xb * xb * xb * i6h;
                if constexpr (REAL) val = m0 * cb + m1 * ca + (y0 * ih - m0 * h6) * xb + (y1 * ih - m1 * h6) * xa;
                else val = make_double2(m0.x * cb + m1.x * ca + (y0.x * ih - m0.x * h6) * xb + (y1.x * ih - m1.x * h6) * xa,
                                        m0.y * cb + m1.y * ca + (y0.y * ih - m0.y * h6) * xb + (y1.y * ih - m1.y * h6) * xa);
            } else {
                const double qf = tt * dt.inv_h, qi = th * dt.inv_h;
                const int a = min(max(__double2int_rd(qf), 0), dt.n - 2), b = min(max(__double2int_rd(qi), 0), dt.n - 2);
                val = cell3_apply_staged<REAL>(tab, 1, grid_cell3_from(a, qf - (double)a, b, qi - (double)b));
            }
            Tt[(nP + q) * ns + smp] = c.okflag[smp] ? val : N::zero();
        }
    } else {
        for (int task = c.tid; task < nD * ns; task += c.nthr) {
            const int q = task >> c.ns_sh, smp = task & (ns - 1);
            const bool ok = c.okflag[smp] != 0;
            const uint32_t dsw = c.dslots_s[q];
            const int3 ds = make_int3((int)(dsw & 0xFFu), (int)((dsw >> 8) & 0xFFu), (int)(dsw >> 16));
            const double th = times[ds.y * ns + smp];
            double tt = times[ds.x * ns + smp];
            if (tt < th) tt = th;                       // :407-410
            const DevDelta& dt = ds.z < kInlineTables ? p.deltas_inline[ds.z] : p.deltas[ds.z];
            T val;
            if (dt.kind == 0 && dt.n == p.n_tau && dt.inv_h == p.inv_h) {   // table on the P grid: reuse the cells
                const int ih = ds.y * ns + smp, it2 = (tt == th) ? ih : ds.x * ns + smp;
                val = cell_apply_i<REAL>(dt.y, 1, grid_cell_from(cella[it2], cellw[it2], cella[ih], cellw[ih]));
            } else {
                val = delta_apply_i<REAL>(dt, tt, th);
            }
            Tt[(nP + q) * ns + smp] = ok ? val : N::zero();
        }
    }
}

// -- 3b. propagator slots: thread = (backbone interval, sample) evaluates all sectors.
template <bool REAL, bool STAGED>
__device__ __forceinline__ void phase_fill_P(const Cta<REAL>& c, const StepParams& p, const StagedTables<REAL>& st) {
    typedef typename Num<REAL>::T T;
    typedef Num<REAL> N;
    const DevEntry& e = *c.e;
    const int ns = c.ns, nI = e.n_nodes - 1, S = p.S;
    T* Tt = reinterpret_cast<T*>(c.Tb);
    const double* times = c.times; const double* cellw = c.cellw; const int* cella = c.cella;
    for (int task = c.tid; task < nI * ns; task += c.nthr) {
        const int q = task >> c.ns_sh, smp = task & (ns - 1);
        const bool ok = c.okflag[smp] != 0;
        const double ta = times[(q + 1) * ns + smp];
        double tb = times[(q + 2) * ns + smp];
        if (tb < ta) tb = ta;                       // src/topology_eval.jl:362-364
        if (e.mode == 0) {
            for (int s = 0; s < S; ++s) {           // bare: i * (-i) exp(-dt (E + lambda))
                const T val = N::from_real(exp(-(tb - ta) * __ldg(p.E + s)));
                Tt[(q * S + s) * ns + smp] = ok ? val : N::zero();
            }
        } else {
            const bool sw = times[(q + 2) * ns + smp] < ta;   // clamped: both ends in the earlier time's cell
            const int ia = (q + 1) * ns + smp, ib = sw ? ia : ia + ns;
            const GridCell3 cell = grid_cell3_from(cella[ib], cellw[ib], cella[ia], cellw[ia]);
#pragma unroll 4
            for (int s = 0; s < S; ++s) {
                T val;
                if constexpr (STAGED) val = cell3_apply_staged<REAL>(st.Ps + s, S, cell);
                else val = cell3_apply_i<REAL>(p.P + s, p.bsize, cell);
                Tt[(q * S + s) * ns + smp] = ok ? val : N::zero();
            }
        }
    }
}

// -- 4. segment products (with the members' coefficients folded in): warp = (table entry, 32 samples)
template <bool REAL>
__device__ __forceinline__ void phase_segments(const Cta<REAL>& c) {
    const DevEntry& e = *c.e;
    const int sub_sh = max(c.ns_sh - 5, 0);     // ns is a power of two
    if (c.lane >= c.ns) return;
    switch (e.seg_stride) {
#define QIW_SEG(N_) case N_: segment_products<N_, REAL>(e, c.segdefs, c.segcoef, c.Tb, c.row_bytes, c.warp, c.nw, sub_sh, c.lane); break;
        QIW_SEG(1) QIW_SEG(2) QIW_SEG(3) QIW_SEG(4) QIW_SEG(5) QIW_SEG(6) QIW_SEG(7) QIW_SEG(8) QIW_SEG(9)
#undef QIW_SEG
        default: break;
    }
}

// -- 5. configuration sums: the job's work units are (sample sub-block of 32, chunk of the lane program); warp w takes
//       units w, w + nw, ...  A chunk is a list of runs of records of equal shape and initial sector.
template <bool REAL, bool DUAL>
__device__ __forceinline__ void phase_walk(const Cta<REAL>& c, int S, double2* per_sample_out, unsigned long long local0,
                                           unsigned long long count) {
    typedef typename Num<REAL>::T T;
    typedef Num<REAL> N;
    const int sub_sh = max(c.ns_sh - 5, 0), n_units = c.n_chunks << sub_sh;     // ns is a power of two
    const bool lane_on = c.lane < c.ns;
    for (int u = c.warp; u < n_units; u += c.nw) {
        const int ch = u >> sub_sh, sub = u - (ch << sub_sh);
        const unsigned col = (unsigned)(sub * 32 + (c.lane & (c.ns - 1))) * (unsigned)sizeof(T);
        const uint32_t run0 = c.chunk_off[c.chunk0 + ch], run1 = c.chunk_off[c.chunk0 + ch + 1];
        for (uint32_t k = run0; k < run1; ++k) {
            const LaneRun rn = c.runs[k];
            // s2 >= 0: records shared by two sectors (only in the DUAL instantiations of the kernels: programs without
            // such records run the kernels that do not carry the code for them)
            const int s = (int)(rn.z & 0xFFFFu), s2 = DUAL ? (int)(rn.z >> 16) - 1 : -1;
            T acc, acc2 = N::zero();
            if (DUAL && s2 >= 0) {
                const Acc2<REAL> pr = lane_dispatch2<REAL>(rn.w, c.items + rn.x, (int)rn.y, c.Tb + col, c.row_bytes);
                acc = pr.a; acc2 = pr.b;
            } else {
                acc = lane_dispatch<REAL>(rn.w, c.items + rn.x, (int)rn.y, c.Tb + col, c.row_bytes);
            }
            if (!lane_on) { acc = N::zero(); acc2 = N::zero(); }
            if (per_sample_out) {
                // qiw_eval_at_times: the evaluator's value for every sample separately (REAL: value = i * acc)
                const unsigned long long smp = local0 + (unsigned long long)(sub * 32 + c.lane);
                if (lane_on && smp < count) {
                    double2* o = per_sample_out + smp * S + s;
                    if constexpr (REAL) atomicAdd(&o->y, acc);
                    else { atomicAdd(&o->x, acc.x); atomicAdd(&o->y, acc.y); }
                    if (s2 >= 0) {
                        double2* o2 = per_sample_out + smp * S + s2;
                        if constexpr (REAL) atomicAdd(&o2->y, acc2);
                        else { atomicAdd(&o2->x, acc2.x); atomicAdd(&o2->y, acc2.y); }
                    }
                }
            } else {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc = N::add(acc, N::shfl_down(acc, off));
                if (s2 >= 0) {
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) acc2 = N::add(acc2, N::shfl_down(acc2, off));
                }
                if (c.lane == 0) {
                    double2& r = c.red[s * c.nw + c.warp];
                    if constexpr (REAL) r.y += acc;
                    else r = cadd(r, acc);
                    if (s2 >= 0) {
                        double2& r2 = c.red[s2 * c.nw + c.warp];
                        if constexpr (REAL) r2.y += acc2;
                        else r2 = cadd(r2, acc2);
                    }
                }
            }
        }
    }
}

// ---- the step kernel ---------------------------------------------------------------------------------
// CTA = (work item: entry + up to nw chunks of its lane program, sample blocks blockIdx.x, blockIdx.x + gridDim.x, ...
// [, time triple / Sobol sequence blockIdx.z]).
template <bool REAL, bool DUAL>
__global__ void __launch_bounds__(768, 1) scalar_step_kernel(const StepParams p) {
    typedef typename Num<REAL>::T T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // let the next step's grid start as soon as every CTA of this one is running (it waits before it touches P)
    asm volatile("griddepcontrol.launch_dependents;");
    const WorkItem it = p.items[blockIdx.y];
    // the entry's description is read in every phase of every sample block: one copy in shared memory
    __shared__ DevEntry e_s;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.entries + it.entry);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&e_s);
        for (int k = threadIdx.x; k < (int)(sizeof(DevEntry) / 4); k += blockDim.x) dst[k] = __ldg(src + k);
    }
    __syncthreads();
    const DevEntry& e = e_s;
    const DevEntryDyn& dy = p.dyn[it.slot];
    const int S = p.S, nD = e.nD;
    Cta<REAL> c;
    c.e = &e;
    c.lane = threadIdx.x & 31; c.warp = threadIdx.x >> 5; c.nw = blockDim.x >> 5; c.tid = threadIdx.x; c.nthr = blockDim.x;
    c.ns = p.spb; c.ns_sh = p.spb_log2;                          // a power of two <= 32
    // shared memory carve-up (sizes fixed per launch from the largest entry, see host): T[slot][sample] — consecutive
    // lanes = consecutive words.  The roots — dead once the times exist — live at the start of T, which is filled
    // afterwards; the times and grid cells — dead once the propagator and interaction slots are filled — overlap the
    // rows of the segment products, which are written after them (a CTA barrier apart).
    c.row_bytes = (unsigned)c.ns * (unsigned)sizeof(T);
    c.Tb = smem_raw;
    c.red = reinterpret_cast<double2*>(smem_raw + p.red_off);                    // [S][nw]
    uint32_t* dslots_s = reinterpret_cast<uint32_t*>(c.red + (size_t)S * c.nw);  // [max_dslots]
    c.dslots_s = dslots_s;
    c.items = e.lane_items; c.runs = p.runs; c.chunk_off = p.chunk_off; c.chunk0 = it.chunk0; c.n_chunks = it.n_chunks;
    c.segdefs = e.lane_segdef4; c.segcoef = nullptr;
    c.pw = reinterpret_cast<double*>(c.Tb);
    c.times = reinterpret_cast<double*>(smem_raw + p.aux_off);                   // [max_nodes1][ns]
    c.cellw = c.times + p.max_nodes1 * c.ns;
    c.cella = reinterpret_cast<int*>(c.cellw + p.max_nodes1 * c.ns);
    c.okflag = c.cella + p.max_nodes1 * c.ns;                                    // [ns]
    for (int k = threadIdx.x; k < nD; k += c.nthr) { const int4 d = e.dslots[k]; dslots_s[k] = (uint32_t)d.x | ((uint32_t)d.y << 8) | ((uint32_t)d.z << 16); }
    for (int k = threadIdx.x; k < S * c.nw; k += c.nthr) c.red[k] = make_double2(0.0, 0.0);

    double t_i = p.t_i, t_w = p.t_w, t_f = p.t_f;
    // batched evaluation: blockIdx.z selects one (t_i, t_w, t_f) triple of the call
    if (p.times_dev) { const double* tz = p.times_dev + 3 * blockIdx.z; t_i = tz[0]; t_w = tz[1]; t_f = tz[2]; }
    // randomised qMC: blockIdx.z selects one of several scrambled sequences (no root cache then)
    const uint32_t* __restrict__ sm = dy.sobol + (size_t)blockIdx.z * p.sobol_z_stride;
    const unsigned long long count = dy.count;
    const int n_sb = (int)((count + (unsigned long long)c.ns - 1ull) / (unsigned long long)c.ns);
    const StagedTables<REAL> none = {nullptr, nullptr, nullptr};

    // optional per-CTA timeline (diagnostics; compiled in only with -DQIW_TRACE_BUILD because
    // reading %globaltimer costs microseconds): start / tables ready / walk done / end
#ifdef QIW_TRACE_BUILD
    unsigned long long* trace = p.trace ? p.trace + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 12 : nullptr;
#else
    constexpr unsigned long long* trace = nullptr;
#endif
    if (trace && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        trace[7] = globaltimer_ns();   // one wall-clock stamp; phase durations use the SM cycle counter
        trace[0] = clock64(); trace[4] = smid; trace[5] = it.entry; trace[6] = (unsigned long long)it.n_chunks;
    }

    for (int sb = blockIdx.x; sb < n_sb; sb += gridDim.x) {
        const unsigned long long local0 = (unsigned long long)sb * (unsigned long long)c.ns;
        if (p.explicit_times == nullptr)
            phase_roots<REAL>(c, sm, dy.start, local0, count, p.sobol_z_stride ? nullptr : dy.ucache, dy.ucache_valid != 0,
                              (int)blockIdx.y == dy.item0);
        else
            for (int smp = threadIdx.x; smp < c.ns; smp += c.nthr) c.okflag[smp] = (local0 + smp < count) ? 1 : 0;
        __syncthreads();
        if (trace && threadIdx.x == 0) trace[8] = clock64();
        phase_times<REAL>(c, t_i, t_w, t_f, p.n_tau, p.inv_h, p.explicit_times, local0, count);
        __syncthreads();
        if (trace && threadIdx.x == 0) trace[9] = clock64();
        // pair interactions first: they do not depend on the bold propagators, so with programmatic dependent launch
        // this part (like everything above) overlaps the previous step's tail
        phase_fill_delta<REAL, false>(c, p, none);
        // Programmatic dependent launch: this grid may have been started while the previous step's grid
        // was still finishing.  Everything up to here used only data no kernel writes; the P table, the
        // partial-sum rows and the arrival counter belong to the previous grid until it has completed.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        phase_fill_P<REAL, false>(c, p, none);
        __syncthreads();
        if (trace && threadIdx.x == 0) trace[10] = clock64();
        phase_segments<REAL>(c);
        __syncthreads();
        if (trace && threadIdx.x == 0) trace[1] = clock64();
        phase_walk<REAL, DUAL>(c, S, p.per_sample_out, local0, count);
        __syncthreads();
        if (trace && threadIdx.x == 0) trace[2] = clock64();
    }

    if (p.per_sample_out) return;

    // -- 6. CTA result: warps summed in fixed order (CTAs without a sample block have not waited yet) --
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();     // a CTA without a sample block comes straight from zeroing `red` (found by compute-sanitizer racecheck)
    if (c.warp == 0) {
        for (int s = c.lane; s < S; s += 32) {
            double2 v = c.red[s * c.nw];
            for (int w2 = 1; w2 < c.nw; ++w2) v = cadd(v, c.red[s * c.nw + w2]);
            p.partials[((size_t)blockIdx.z * gridDim.y * gridDim.x + (size_t)it.partial0 * gridDim.x + blockIdx.x) * S + s] = v;
        }
    }
    if (trace && threadIdx.x == 0) trace[3] = clock64();

    // -- 7. fused tail: the last CTA to arrive reduces all partial sums (and updates P) ---------------
    if (p.done_counter) {
        __shared__ int is_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned ticket = atomicAdd(p.done_counter + blockIdx.z, 1u);
            is_last = (ticket == gridDim.x * gridDim.y - 1u) ? 1 : 0;
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        fused_tail(p, t_i, t_w, t_f, (int)gridDim.x);
        if (threadIdx.x == 0) p.done_counter[blockIdx.z] = 0u;
        if (trace && threadIdx.x == 0) trace[11] = clock64();
    }
}

// ---- the persistent run kernel -----------------------------------------------------------------------
// inchworm!'s loop over the bold steps (src/inchworm.jl:474-493) in ONE cooperative launch, for runs whose steps are
// too small to fill the machine (README configuration: 1 024 samples per step).  Every CTA owns one job — (entry,
// 32 m samples, part of the lane program) — for the whole run, so everything that does not change from step to step
// stays on chip: the job's simplex roots, its pair-interaction slot list, and per-CTA copies of the P table and of the
// pair-interaction tables in shared memory (as i * value in the kernel's arithmetic), from which the per-sample
// operands are interpolated.  One step = times -> operand table -> segment products -> configuration sums -> one
// partial row per job -> ONE grid barrier -> every CTA reduces all partial rows in the same fixed order, (multi-GPU)
// adds the peers' block sums from its GPU's mailbox in rank order, and applies set_ppgf! + normalize!
// (src/ppgf.jl:495-504,646-668) to its own copy of P.  CTA 0 sends this rank's sums to the peers and keeps the global
// P table and the per-order history up to date.
template <bool REAL, bool DUAL>
__global__ void __launch_bounds__(384, 2) scalar_run_kernel(const RunParams rp) {
    typedef typename Num<REAL>::T T;
    typedef Num<REAL> N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StepParams& p = rp.sp;
    const int S = p.S, n_tau = p.n_tau, n_ent = p.n_call_entries;
    // Job list of this CTA.  The lists are balanced per SM (the step is bound by the SM's shared-memory pipe): the
    // CTAs of one SM claim the lists of one bin, whichever SM the hardware placed them on.
    __shared__ int list_s;
    if (threadIdx.x == 0) {
        int list = (int)blockIdx.x;
        if (rp.sm_map) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            smid &= 1023u;
            const unsigned slot = atomicAdd(rp.sm_map + smid, 1u);
            unsigned bin1;
            if (slot == 0) {
                bin1 = atomicAdd(rp.sm_map + 2048, 1u) + 1u;
                asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(rp.sm_map + 1024 + smid), "r"(bin1) : "memory");
            } else {
                do { asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(bin1) : "l"(rp.sm_map + 1024 + smid) : "memory"); } while (bin1 == 0u);
            }
            list = (int)((bin1 - 1u) * (unsigned)rp.ctas_per_sm + slot);
        }
        list_s = list;
    }
    __syncthreads();
    const int jb0 = rp.cta_job0[list_s], jb1 = rp.cta_job0[list_s + 1];
    const bool single = (jb1 - jb0 == 1);        // the usual case: this CTA's only job keeps its roots and slot list on chip
    Cta<REAL> c;
    c.lane = threadIdx.x & 31; c.warp = threadIdx.x >> 5; c.nw = blockDim.x >> 5; c.tid = threadIdx.x; c.nthr = blockDim.x;
    c.Tb = smem_raw;
    c.okflag = reinterpret_cast<int*>(smem_raw + rp.ok_off);                     // [max ns]
    c.pw = reinterpret_cast<double*>(smem_raw + rp.pw_off);                      // [D][ns]
    c.red = reinterpret_cast<double2*>(smem_raw + rp.red_off);                   // [S][nw]
    uint32_t* dslots_s = reinterpret_cast<uint32_t*>(smem_raw + rp.ds_off);
    c.dslots_s = dslots_s;
    T* Ps = reinterpret_cast<T*>(smem_raw + rp.P_off);                           // [n_tau][S]   i P
    T* Ds = reinterpret_cast<T*>(smem_raw + rp.D_off);                           // [n_tables][n_tau]   i Delta
    double2* outs = reinterpret_cast<double2*>(smem_raw + rp.out_off);           // [n_ent][S] this step's block sums
    double* scales_s = reinterpret_cast<double*>(outs + (size_t)n_ent * S);      // [n_ent] weight * (-i)^d * Jacobian of the step
    int* ejob0_s = reinterpret_cast<int*>(scales_s + n_ent);                     // [n_ent + 1] first partial row of every entry
    for (int k = threadIdx.x; k <= n_ent; k += blockDim.x) ejob0_s[k] = rp.entry_job0[k];
    __shared__ double lambda_s;
    const StagedTables<REAL> st = {Ps, Ds, p.tables_on_grid ? nullptr : rp.D_table_off};
    __shared__ DevEntry e_s;      // the current job's entry description (read in every phase)
    auto view = [&](const RunJob& job, const WorkItem& it, const DevEntry& e) {      // followed by a CTA barrier
        {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(&e);
            uint32_t* dst = reinterpret_cast<uint32_t*>(&e_s);
            for (int k = threadIdx.x; k < (int)(sizeof(DevEntry) / 4); k += blockDim.x) dst[k] = __ldg(src + k);
        }
        c.e = &e_s;
        c.items = e.lane_items; c.runs = p.runs; c.chunk_off = p.chunk_off; c.chunk0 = it.chunk0; c.n_chunks = it.n_chunks;
        c.segdefs = e.lane_segdef4; c.segcoef = nullptr;
        c.ns = 32 * job.n_sub; c.ns_sh = 31 - __clz(c.ns);
        c.row_bytes = (unsigned)c.ns * (unsigned)sizeof(T);
        c.times = reinterpret_cast<double*>(smem_raw + job.aux_off);             // overlaps the segment-product rows
        c.cellw = c.times + (e.n_nodes + 1) * c.ns;
        c.cella = reinterpret_cast<int*>(c.cellw + (e.n_nodes + 1) * c.ns);
    };

    // ---- once per run: tables, and every job's roots (kept on chip if the CTA has one job, else in the root cache) ----
    for (int k = threadIdx.x; k < n_tau * S; k += c.nthr) Ps[k] = N::times_i_of(p.P[(size_t)(k / S) * p.bsize + (k % S)]);
    if (p.tables_on_grid) {
        for (int k = threadIdx.x; k < rp.n_tables * n_tau; k += c.nthr) Ds[k] = N::times_i_of(__ldg(p.deltas_inline[k / n_tau].y + (k % n_tau)));
    } else {
        for (int t = 0; t < rp.n_tables; ++t) {
            const DevDelta& dt = p.deltas_inline[t];
            T* tab = Ds + rp.D_table_off[t];
            for (int k = threadIdx.x; k < dt.n; k += c.nthr) {
                tab[k] = N::times_i_of(__ldg(dt.y + k));
                if (dt.kind == 1) tab[dt.n + k] = N::times_i_of(__ldg(dt.M + k));
            }
        }
    }
    for (int jj = jb0; jj < jb1; ++jj) {
        const RunJob job = rp.jobs[jj];
        const WorkItem it = p.items[job.item];
        const DevEntry& e = p.entries[it.entry];
        const DevEntryDyn& dy = p.dyn[it.slot];
        view(job, it, e);
        __syncthreads();
        phase_roots<REAL>(c, dy.sobol, dy.start, (unsigned long long)job.sb0 * 32ull, dy.count, dy.ucache, dy.ucache_valid != 0, true);
        if (single) {
            for (int k = threadIdx.x; k < e.nD; k += c.nthr) { const int4 d = e.dslots[k]; dslots_s[k] = (uint32_t)d.x | ((uint32_t)d.y << 8) | ((uint32_t)d.z << 16); }
            if (job.stash_off >= 0) {
                // this CTA's only job keeps its part of the lane program on chip for the whole run: records, runs, chunk
                // table, segment definitions and their coefficients (layout mirrored by the host: enqueue_run)
                const uint32_t r0 = __ldg(p.chunk_off + it.chunk0), r1 = __ldg(p.chunk_off + it.chunk0 + it.n_chunks);
                uint4* items_s = reinterpret_cast<uint4*>(smem_raw + job.stash_off);
                uint32_t n4 = 0;
                for (uint32_t k = r0; k < r1; ++k) {
                    const LaneRun rn = __ldg(p.runs + k);
                    const uint32_t code = rn.w, ni = (code >> 4) + ((((code >> 2) & 3u) + 1u) << (code & 3u)), len = rn.y * ((ni + 7u) / 8u);
                    for (uint32_t q = threadIdx.x; q < len; q += c.nthr) items_s[n4 + q] = __ldg(e.lane_items + rn.x + q);
                    n4 += len;
                }
                for (uint32_t q = threadIdx.x; q < 8u; q += c.nthr) items_s[n4 + q] = make_uint4(0u, 0u, 0u, 0u);   // the walk fetches one record ahead
                n4 += 8u;
                LaneRun* runs_s = reinterpret_cast<LaneRun*>(items_s + n4);
                if (threadIdx.x == 0) {
                    uint32_t off = 0;
                    for (uint32_t k = r0; k < r1; ++k) {
                        LaneRun rn = __ldg(p.runs + k);
                        const uint32_t code = rn.w, ni = (code >> 4) + ((((code >> 2) & 3u) + 1u) << (code & 3u)), len = rn.y * ((ni + 7u) / 8u);
                        rn.x = off; off += len;
                        runs_s[k - r0] = rn;
                    }
                }
                uint32_t* choff_s = reinterpret_cast<uint32_t*>(runs_s + (r1 - r0));
                for (int k = threadIdx.x; k <= it.n_chunks; k += c.nthr) choff_s[k] = __ldg(p.chunk_off + it.chunk0 + k) - r0;
                const int NW = e.seg_stride > 7 ? 2 : 1;
                uint4* defs_s = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(choff_s) + (((size_t)(it.n_chunks + 1) * 4 + 15) & ~(size_t)15));
                for (int k = threadIdx.x; k < e.nSegL * NW; k += c.nthr) defs_s[k] = __ldg(e.lane_segdef4 + k);
                T* coef_s = reinterpret_cast<T*>(defs_s + (size_t)e.nSegL * NW);
                for (int k = threadIdx.x; k < e.nSegL; k += c.nthr) {
                    const uint32_t ci = __ldg(reinterpret_cast<const uint32_t*>(e.lane_segdef4 + (size_t)k * NW)) & 0xFFFFu;
                    coef_s[k] = (ci != 0xFFFFu) ? N::coef_of(__ldg(e.coefs + ci)) : N::from_real(1.0);
                }
                c.items = items_s; c.runs = runs_s; c.chunk_off = choff_s; c.chunk0 = 0; c.segdefs = defs_s; c.segcoef = coef_s;
            }
        }
    }
    __syncthreads();
    const double h = p.h;
    const size_t prow = (size_t)rp.n_jobs * S;
    T* partials = reinterpret_cast<T*>(rp.partials);

#ifdef QIW_TRACE_BUILD
    // per-CTA timeline of one step in the middle of the run (diagnostics): SM cycle counter at the phase boundaries
    unsigned long long* trace = p.trace ? p.trace + (size_t)blockIdx.x * 16 : nullptr;
#define QIW_RT(k_) if (trace && step == rp.n_steps / 2 && threadIdx.x == 0) trace[k_] = clock64();
#else
#define QIW_RT(k_)
#endif
    // single-job CTAs: the times and the pair-interaction rows of the COMING step do not depend on this step's result;
    // a part of the CTA's warps prepares them while the others reduce, exchange and update P (`pre`), which takes them —
    // and, on several GPUs, the NVLink round trip of the all-reduce — off the step's critical path
    bool pre = false;
    for (int step = 0; step < rp.n_steps; ++step) {
        const int k_w = rp.k_first + step, k_f = k_w + 1;
        const double t_i = 0.0, t_w = (double)k_w * h, t_f = (double)k_f * h;
        QIW_RT(0)
        // this step's partial rows (buffers alternate with the step's parity: a CTA that is one step ahead never
        // overwrites rows a slower CTA is still reading)
        T* my_rows = partials + (size_t)(step & 1) * prow;
        for (int jj = jb0; jj < jb1; ++jj) {
            const RunJob job = rp.jobs[jj];
            const WorkItem it = p.items[job.item];
            const DevEntry& e = p.entries[it.entry];
            const DevEntryDyn& dy = p.dyn[it.slot];
            const unsigned long long count = dy.count, local0 = (unsigned long long)job.sb0 * 32ull;
            if (!single) { view(job, it, e); __syncthreads(); }
            if (!single) {     // the roots written in the prologue (or recomputed, if the call has no root cache)
                phase_roots<REAL>(c, dy.sobol, dy.start, local0, count, dy.ucache, dy.ucache != nullptr, false);
                for (int k = threadIdx.x; k < e.nD; k += c.nthr) { const int4 d = e.dslots[k]; dslots_s[k] = (uint32_t)d.x | ((uint32_t)d.y << 8) | ((uint32_t)d.z << 16); }
            }
            for (int k = threadIdx.x; k < S * c.nw; k += c.nthr) c.red[k] = make_double2(0.0, 0.0);
            if (!single) __syncthreads();
            if (!pre) {
                phase_times<REAL>(c, t_i, t_w, t_f, n_tau, p.inv_h, nullptr, local0, count);
                __syncthreads();
                phase_fill_delta<REAL, true>(c, p, st);
            }
            if (jj == jb0) { QIW_RT(1) }
            phase_fill_P<REAL, true>(c, p, st);
            __syncthreads();
            if (jj == jb0) { QIW_RT(2) }
            phase_segments<REAL>(c);
            __syncthreads();
            if (jj == jb0) { QIW_RT(3) }
            phase_walk<REAL, DUAL>(c, S, nullptr, local0, count);
            __syncthreads();
            if (jj == jb0) { QIW_RT(4) }
            if (c.warp == 0) {
                for (int s = c.lane; s < S; s += 32) {
                    double2 v = c.red[s * c.nw];
                    for (int w2 = 1; w2 < c.nw; ++w2) v = cadd(v, c.red[s * c.nw + w2]);
                    if constexpr (REAL) my_rows[(size_t)job.row * S + s] = v.y;
                    else my_rows[(size_t)job.row * S + s] = v;
                }
            }
        }
        // ---- grid barrier: all CTAs are co-resident (cooperative launch) ----
        QIW_RT(5)
        // Arrive with release semantics (ordered after this CTA's partial rows through the CTA barrier), spin with
        // relaxed loads, read the rows from L2 (__ldcg): unlike __threadfence() / ld.acquire this does not invalidate
        // the SM's L1, which holds the program data every step re-reads.
        __syncthreads();
        // the entries' scale factors do not depend on the sums: the last warp computes them while thread 0 waits
        for (int i = (int)threadIdx.x - (c.nthr - 32); i >= 0 && i < n_ent; i += 32) scales_s[i] = entry_scale(p.entries[p.dyn[i].entry], p.dyn[i], t_i, t_w, t_f);
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(rp.barrier), "r"(1u) : "memory");
            const unsigned target = (unsigned)(step + 1) * gridDim.x;
            unsigned seen;
            do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(rp.barrier) : "memory"); } while (seen < target);
        }
        __syncthreads();
        QIW_RT(6)
        // Warp specialisation after the barrier.  Group A (the first nA warps): the reduction of the partial rows, the
        // exchange with the peer GPUs and the update of this CTA's copy of P — a chain of latencies, not of work.  Group B
        // (the other warps): the times and the pair-interaction rows of the COMING step, which do not depend on this
        // step's result (`pre`).  The two groups meet at the CTA barrier below.
        const bool do_pre = single && c.nw > rp.post_warps && step + 1 < rp.n_steps && (rp.jobs[jb0].flags & 1);
        const int nA = do_pre ? rp.post_warps : c.nw, nthrA = nA * 32, tidA = (int)threadIdx.x;     // no group B: everybody is group A
        // On several GPUs the whole CTA reduces (the sooner this rank's sums are on the wire the better) and splits
        // afterwards; on one GPU group B starts on the next step right away and group A reduces alone.
        const int nR = (p.peer_ranks > 1) ? c.nw : nA, nthrR = nR * 32, tidR = (int)threadIdx.x;
        auto bar_A = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(nthrA) : "memory"); };
        auto bar_R = [&]() { if (nR == c.nw) __syncthreads(); else asm volatile("bar.sync 1, %0;" ::"r"(nthrR) : "memory"); };
        if (c.warp < nR) {
            // ---- every CTA: block sums of every entry, same fixed order everywhere ----
            // all rows first (independent L2 loads, one latency), into the operand table's space, which is dead by now
            const T* rows = my_rows;
            if (rp.rows_staged) {
                T* rows_s = reinterpret_cast<T*>(smem_raw);
                const int n_val = rp.n_jobs * S;
                for (int k0 = tidR; k0 < n_val; k0 += 10 * nthrR) {      // ten loads in flight per thread: one L2 round trip
                    T tmp[10];
#pragma unroll
                    for (int u = 0; u < 10; ++u) { const int k = k0 + u * nthrR; tmp[u] = (k < n_val) ? __ldcg(my_rows + k) : N::zero(); }
#pragma unroll
                    for (int u = 0; u < 10; ++u) { const int k = k0 + u * nthrR; if (k < n_val) rows_s[k] = tmp[u]; }
                }
                bar_R();
                rows = rows_s;
            }
            // four lanes per (entry, sector) output: lane g adds rows g, g+4, ... in order, then a fixed butterfly
            for (int o0 = 0; o0 < n_ent * S; o0 += nthrR / 4) {
                const int o = o0 + (int)tidR / 4, g = (int)tidR & 3;
                T v = N::zero();
                double scale = 0.0;
                if (o < n_ent * S) {
                    const int i = o / S, s = o - i * S;
                    const int j0 = ejob0_s[i], j1 = ejob0_s[i + 1];
                    if (rp.rows_staged) for (int j = j0 + g; j < j1; j += 4) v = N::add(v, rows[(size_t)j * S + s]);
                    else for (int j = j0 + g; j < j1; j += 4) v = N::add(v, __ldcg(rows + (size_t)j * S + s));
                    scale = scales_s[i];
                }
                v = N::add(v, N::shfl_xor(v, 1));
                v = N::add(v, N::shfl_xor(v, 2));
                if (o < n_ent * S && g == 0) {
                    if constexpr (REAL) outs[o] = make_double2(0.0, scale * v);
                    else outs[o] = cscale(scale, v);
                }
            }
            bar_R();
        }
        if (c.warp < nA) {
            if (p.peer_ranks > 1) {
                // ---- all-reduce over peer memory (protocol: fused_tail).  CTA 0 sends; every CTA of this GPU receives ----
                const unsigned long long seq = p.peer_seq + (unsigned long long)step;
                const int par = (int)(seq & 1ull);
                const unsigned int seq32 = (unsigned int)(seq % 0xFFFFFFFFull) + 1u;
                const size_t my_slot = kPeerFlagBytes + ((size_t)p.peer_rank * 2 + par) * kPeerSlotBytes;
                const int n_dbl = 2 * n_ent * S;
                double* outd = reinterpret_cast<double*>(outs);
                if (blockIdx.x == 0)
                    for (int j = tidA; j < n_dbl; j += nthrA) peer_send(p, my_slot, j, outd[j], seq32);
                const unsigned long long t0 = globaltimer_ns();
                for (int j = tidA; j < n_dbl; j += nthrA)
                    outd[j] = peer_gather(p, par, seq32, j, outd[j], t0);   // element j is read and written by this thread only
                bar_A();
            }
            QIW_RT(7)
            // ---- set_ppgf!(P, tau_f, sum of the entries) and normalize!(P, tau_f) on this CTA's copy ----
            // row sums: warp = sector, lanes = entries, fixed butterfly (the same sum in every CTA and on every rank)
            for (int s = c.warp; s < S; s += nA) {
                double2 v = make_double2(0.0, 0.0);
                for (int j = c.lane; j < n_ent; j += 32) {
                    const double2 cj = outs[j * S + s];
                    v = cadd(v, cj);
                    if (blockIdx.x == 0 && rp.hist) rp.hist[(size_t)k_f * rp.hist_stride + rp.hist_off + (size_t)j * S + s] = cj;
                }
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) {
                    v.x += __shfl_xor_sync(0xFFFFFFFFu, v.x, m);
                    v.y += __shfl_xor_sync(0xFFFFFFFFu, v.y, m);
                }
                if (c.lane == 0) Ps[k_f * S + s] = N::times_i_of(v);
            }
            bar_A();
            // lambda from the diagonal (every block of a scalar model is its own diagonal element)
            if (tidA == 0) {
                double pmax = -1.0e300;
                for (int s = 0; s < S; ++s) {
                    double mip;   // -Im P_s(tau_f)
                    if constexpr (REAL) mip = Ps[k_f * S + s]; else mip = Ps[k_f * S + s].x;
                    pmax = fmax(pmax, mip);
                }
                lambda_s = log(pmax) / ((double)k_f * h);
            }
            bar_A();
            const double lambda = lambda_s;
            for (int k = tidA; k < n_tau; k += nthrA) {      // one exponential per grid point
                const double f = exp(-((double)k * h) * lambda);
                for (int s = 0; s < S; ++s) {
                    if constexpr (REAL) Ps[k * S + s] = f * Ps[k * S + s];
                    else Ps[k * S + s] = cscale(f, Ps[k * S + s]);
                }
            }
            bar_A();
        } else if (do_pre) {
            const RunJob job = rp.jobs[jb0];
            const unsigned long long count = p.dyn[p.items[job.item].slot].count, local0 = (unsigned long long)job.sb0 * 32ull;
            Cta<REAL> cB = c;
            cB.tid = (int)threadIdx.x - nthrA; cB.nthr = c.nthr - nthrA;
            phase_times<REAL>(cB, 0.0, (double)(k_w + 1) * h, (double)(k_f + 1) * h, n_tau, p.inv_h, nullptr, local0, count);
            asm volatile("bar.sync 2, %0;" ::"r"(cB.nthr) : "memory");
            phase_fill_delta<REAL, true>(cB, p, st);
        }
        pre = do_pre;
        __syncthreads();
        QIW_RT(8)
#ifdef QIW_TRACE_BUILD
        if (trace && step == rp.n_steps / 2 && threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            trace[9] = smid; trace[10] = (unsigned long long)(jb1 - jb0); trace[11] = (unsigned long long)p.items[rp.jobs[jb0].item].entry;
            trace[12] = (unsigned long long)rp.jobs[jb0].n_sub; trace[13] = globaltimer_ns();
        }
#endif
    }
    // the P table goes back to global memory once: P = -i * (i P)
    if (blockIdx.x == 0) {
        double2* Pg = p.finish_P;
        for (int idx = threadIdx.x; idx < n_tau * S; idx += c.nthr) {
            double2 v;
            if constexpr (REAL) v = make_double2(0.0, -Ps[idx]);
            else v = make_double2(Ps[idx].y, -Ps[idx].x);
            Pg[(size_t)(idx / S) * p.bsize + (idx % S)] = v;
        }
    }
}

// ---- host-callable launcher --------------------------------------------------------------------------
// The opt-in to more than 48 KB of dynamic shared memory is a per-device function attribute: one bit per
// (device, instantiation), so that contexts on several GPUs of one process all get it.
template <class K>
static cudaError_t optin_smem(K kernel, unsigned long long* mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && ((__atomic_load_n(mask, __ATOMIC_RELAXED) >> dev) & 1ull)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    if (dev < 64) __atomic_fetch_or(mask, 1ull << dev, __ATOMIC_RELAXED);
    return cudaSuccess;
}

template <bool REAL, bool DUAL>
static cudaError_t launch_scalar_t(const StepParams& p, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    static unsigned long long attr_mask = 0ull;
    cudaError_t e = optin_smem(scalar_step_kernel<REAL, DUAL>, &attr_mask);
    if (e != cudaSuccess) return e;
    // programmatic dependent launch: consecutive step kernels of a stream may overlap prologue and tail
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = p.allow_overlap ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, scalar_step_kernel<REAL, DUAL>, p);
}

// `real_mode`: every table and coefficient in use has been verified purely imaginary by the host.
// `dual`: some entry of the launch has records shared by two initial sectors (LaneRun::z).
cudaError_t launch_scalar_step(bool real_mode, bool dual, const StepParams& p, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    if (dual) return real_mode ? launch_scalar_t<true, true>(p, grid, threads, smem, st) : launch_scalar_t<false, true>(p, grid, threads, smem, st);
    return real_mode ? launch_scalar_t<true, false>(p, grid, threads, smem, st) : launch_scalar_t<false, false>(p, grid, threads, smem, st);
}

template <bool REAL, bool DUAL>
static cudaError_t launch_run_t(const RunParams& rp, int n_ctas, int threads, size_t smem, cudaStream_t st) {
    static unsigned long long attr_mask = 0ull;
    cudaError_t e = optin_smem(scalar_run_kernel<REAL, DUAL>, &attr_mask);
    if (e != cudaSuccess) return e;
    void* args[] = {(void*)&rp};
    return cudaLaunchCooperativeKernel((const void*)scalar_run_kernel<REAL, DUAL>, dim3((unsigned)n_ctas), dim3((unsigned)threads), args, smem, st);
}

cudaError_t launch_scalar_run(bool real_mode, bool dual, const RunParams& rp, int n_ctas, int threads, size_t smem, cudaStream_t st) {
    if (dual) return real_mode ? launch_run_t<true, true>(rp, n_ctas, threads, smem, st) : launch_run_t<false, true>(rp, n_ctas, threads, smem, st);
    return real_mode ? launch_run_t<true, false>(rp, n_ctas, threads, smem, st) : launch_run_t<false, false>(rp, n_ctas, threads, smem, st);
}

// CTAs of the run kernel that can be co-resident on the device (cooperative launch limit).
template <bool REAL, bool DUAL>
static int run_max_ctas_t(int threads, size_t smem, int n_sm) {
    static unsigned long long mask = 0ull;
    int per_sm = 0;
    if (optin_smem(scalar_run_kernel<REAL, DUAL>, &mask) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scalar_run_kernel<REAL, DUAL>, threads, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return per_sm * n_sm;
}

int scalar_run_max_ctas(bool real_mode, bool dual, int threads, size_t smem, int n_sm) {
    if (dual) return real_mode ? run_max_ctas_t<true, true>(threads, smem, n_sm) : run_max_ctas_t<false, true>(threads, smem, n_sm);
    return real_mode ? run_max_ctas_t<true, false>(threads, smem, n_sm) : run_max_ctas_t<false, false>(threads, smem, n_sm);
}

}  // namespace qiw
