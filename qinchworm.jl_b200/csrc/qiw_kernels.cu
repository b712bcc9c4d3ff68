// qiw_kernels.cu — hand-written CUDA kernels (sm_100a) of the qMC diagram-evaluation hot path.
//
// scalar_step_kernel (1x1 sector blocks): a CTA owns a block of <= 32 scrambled-Sobol samples
// of one entry (one TopologiesInputData).  All warps first build the per-sample operand tables in shared
// memory T[sample][slot] —
//   Sobol point by Gray-code random access      (src/scrambled_sobol.jl:158-197)
//   cube -> ordered-time simplex                (src/qmc_integrate.jl:225-235,425-449)
//   i*P_s(t_pos, t_pos-1) for every interval/sector, i*Delta for every used arc/table
//                                               (src/topology_eval.jl:357-374,397-416)
//   products of the propagators over segments of the backbone (factorised records, qiw_compile.cpp)
// — and then every LANE takes one pre-compiled, pruned configuration (src/topology_eval.jl:454-556) whose
// record of K + order operand slots it keeps in registers, and loops over the CTA's samples.  The last CTA
// to finish reduces all partial sums in fixed order, exchanges them with the peer GPUs and applies
// set_ppgf! / normalize! (fused tail).  In real mode (everything purely imaginary-time) the same kernel runs
// in real arithmetic.
//
// block_walk_kernel (sector blocks up to 4x4, real arithmetic): lane = sample, warp-uniform replay of the
// pruned configuration tree with the running matrix product in registers.  block_step_kernel is the general
// complex path for block models.
#include <cstdio>

#include "qiw_device.cuh"

namespace qiw {

// ---- small complex helpers -------------------------------------------------------------------

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {  // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 cscale(double a, double2 b) { return make_double2(a * b.x, a * b.y); }
__device__ __forceinline__ double2 times_i(double2 a) { return make_double2(-a.y, a.x); }

// ---- interpolation ---------------------------------------------------------------------------

// Keldysh.jl's generic grid interpolation of a translation-invariant imaginary-time function
// stored as D[k] = G(k h): bilinear on the cell (a, b) of the (t_f, t_i) grid, linear on the
// triangle when both times share a cell (rule: DESIGN.md §2; call sites
// src/topology_eval.jl:368,414).
__device__ __forceinline__ double2 grid_interp(const double2* __restrict__ D, int stride, int n, double inv_h,
                                               double t_f, double t_i) {
    // t * (1/h) instead of t / h: may pick the neighbouring cell when t sits on a grid point to the
    // last bit, where the interpolant is continuous, so the value changes by O(ulp) only
    const double qf = t_f * inv_h, qi = t_i * inv_h;
    int a = (int)floor(qf), b = (int)floor(qi);
    a = min(max(a, 0), n - 2);
    b = min(max(b, 0), n - 2);
    const double w1 = qf - (double)a, w2 = qi - (double)b;
    if (a == b) {
        const double2 d0 = __ldg(D), d1 = __ldg(D + stride);
        const double w = w1 - w2;
        return make_double2(d0.x + w * (d1.x - d0.x), d0.y + w * (d1.y - d0.y));
    }
    const int k = a - b;
    const double2 dk = __ldg(D + (size_t)k * stride), dp = __ldg(D + (size_t)(k + 1) * stride),
                  dm = __ldg(D + (size_t)(k - 1) * stride);
    const double c00 = (1.0 - w1) * (1.0 - w2), c10 = w1 * (1.0 - w2), c01 = (1.0 - w1) * w2, c11 = w1 * w2;
    return make_double2(c00 * dk.x + c10 * dp.x + c01 * dm.x + c11 * dk.x,
                        c00 * dk.y + c10 * dp.y + c01 * dm.y + c11 * dk.y);
}

// Natural cubic spline in dt = t_f - t_i (src/spline_gf.jl:208-219).
__device__ __forceinline__ double2 spline_eval(const DevDelta& t, double dt) {
    const double h = t.h;
    int j = (int)floor(dt * t.inv_h);
    j = min(max(j, 0), t.n - 2);
    const double xa = dt - (double)j * h, xb = (double)(j + 1) * h - dt;
    const double2 y0 = __ldg(t.y + j), y1 = __ldg(t.y + j + 1), m0 = __ldg(t.M + j), m1 = __ldg(t.M + j + 1);
    const double i6h = t.inv_h * (1.0 / 6.0), h6 = h * (1.0 / 6.0), ih = t.inv_h;
    const double ca = xa * xa * xa * i6h, cb = xb * xb * xb * i6h;
    return make_double2(m0.x * cb + m1.x * ca + (y0.x * ih - m0.x * h6) * xb + (y1.x * ih - m1.x * h6) * xa,
                        m0.y * cb + m1.y * ca + (y0.y * ih - m0.y * h6) * xb + (y1.y * ih - m1.y * h6) * xa);
}

__device__ __forceinline__ double2 delta_eval(const DevDelta& t, double t_f, double t_i) {
    if (t.kind == 1) return spline_eval(t, t_f - t_i);
    return grid_interp(t.y, 1, t.n, t.inv_h, t_f, t_i);
}

// ---- Sobol -----------------------------------------------------------------------------------

// Point k (0-based) of a digital sequence: x0 xor the direction numbers selected by gray(k);
// identical to k calls of next! (src/scrambled_sobol.jl:158-173).
__device__ __forceinline__ uint32_t sobol_coord(const uint32_t* __restrict__ m, uint32_t x0, uint32_t k) {
    uint32_t g = k ^ (k >> 1), x = x0;
    int b = 0;
    while (g) {
        if (g & 1u) x ^= __ldg(m + b);
        g >>= 1;
        ++b;
    }
    return x;
}

__global__ void sobol_points_kernel(int D, const uint32_t* __restrict__ m, const uint32_t* __restrict__ x0,
                                    unsigned long long start, unsigned long long count, uint32_t* __restrict__ out) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= count * (unsigned long long)D) return;
    const unsigned long long k = i / D;
    const int d = (int)(i % D);
    out[i] = sobol_coord(m + d * 32, x0[d], (uint32_t)(start + k));
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- arithmetic modes ----------------------------------------------------------------------------
// On the imaginary-time branch every factor of a configuration's weight is i*P or i*Delta with P and
// Delta purely imaginary, and the folded coefficient (operator matrix elements times -i * parity *
// (-1)^order) is purely imaginary: the whole product is (real number) * i, exactly.  When the host has
// verified that for the tables and coefficients in use (DESIGN.md §3 "real mode") the kernel runs in
// real arithmetic — bit-identical results, one FP64 multiply and one 8-byte shared-memory operand per
// factor instead of four and 16 bytes.  Anything else runs the same code in complex arithmetic.
template <bool REAL> struct Num;
template <> struct Num<true> {
    typedef double T;
    static __device__ __forceinline__ T zero() { return 0.0; }
    static __device__ __forceinline__ T mul(T a, T b) { return a * b; }
    static __device__ __forceinline__ T add(T a, T b) { return a + b; }
    static __device__ __forceinline__ T from_real(double x) { return x; }
    static __device__ __forceinline__ T times_i_of(double2 v) { return -v.y; }          // Re(i v), Im(i v) = v.x = 0
    static __device__ __forceinline__ T coef_of(double2 c) { return c.y; }               // coef = i * c.y
    static __device__ __forceinline__ double2 result(T coef, T acc) { return make_double2(0.0, coef * acc); }
    static __device__ __forceinline__ T shfl_down(T v, int off) { return __shfl_down_sync(0xFFFFFFFFu, v, off); }
};
template <> struct Num<false> {
    typedef double2 T;
    static __device__ __forceinline__ T zero() { return make_double2(0.0, 0.0); }
    static __device__ __forceinline__ T mul(T a, T b) { return cmul(a, b); }
    static __device__ __forceinline__ T add(T a, T b) { return cadd(a, b); }
    static __device__ __forceinline__ T from_real(double x) { return make_double2(x, 0.0); }
    static __device__ __forceinline__ T times_i_of(double2 v) { return times_i(v); }
    static __device__ __forceinline__ T coef_of(double2 c) { return c; }
    static __device__ __forceinline__ double2 result(T coef, T acc) { return cmul(coef, acc); }
    static __device__ __forceinline__ T shfl_down(T v, int off) {
        return make_double2(__shfl_down_sync(0xFFFFFFFFu, v.x, off), __shfl_down_sync(0xFFFFFFFFu, v.y, off));
    }
};

// Cell and weights of the Keldysh.jl grid rule for one (t_f, t_i) pair; shared by every sector / table
// evaluated at that pair.  k = 0: both times in one cell (triangular rule).
struct GridCell { int k; double c00, c10, c01, c11; };
__device__ __forceinline__ GridCell grid_cell(int n, double inv_h, double t_f, double t_i) {
    const double qf = t_f * inv_h, qi = t_i * inv_h;
    int a = __double2int_rd(qf), b = __double2int_rd(qi);
    a = min(max(a, 0), n - 2);
    b = min(max(b, 0), n - 2);
    const double w1 = qf - (double)a, w2 = qi - (double)b;
    GridCell c;
    c.k = a - b;
    if (c.k == 0) { c.c00 = w1 - w2; c.c10 = c.c01 = c.c11 = 0.0; }
    else { c.c00 = (1.0 - w1) * (1.0 - w2); c.c10 = w1 * (1.0 - w2); c.c01 = (1.0 - w1) * w2; c.c11 = w1 * w2; }
    return c;
}
// The same from per-time cell indices and fractional weights computed once per backbone position.
__device__ __forceinline__ GridCell grid_cell_from(int a, double w1, int b, double w2) {
    GridCell c;
    c.k = a - b;
    if (c.k == 0) { c.c00 = w1 - w2; c.c10 = c.c01 = c.c11 = 0.0; }
    else { c.c00 = (1.0 - w1) * (1.0 - w2); c.c10 = w1 * (1.0 - w2); c.c01 = (1.0 - w1) * w2; c.c11 = w1 * w2; }
    return c;
}
// Branch-free form used by the table fill: three table indices and three coefficients,
//   value = ck D[i0] + cp D[ip] + cm D[im]
// (k = 0: (1 - w) D[0] + w D[1]; else (c00 + c11) D[k] + c10 D[k+1] + c01 D[k-1]); differs from the reference's
// operation order by O(ulp), and lets the compiler overlap the loads of consecutive table slots.
struct GridCell3 { int i0, ip, im; double ck, cp, cm; };
__device__ __forceinline__ GridCell3 grid_cell3_from(int a, double w1, int b, double w2) {
    GridCell3 c;
    const int k = a - b;
    const bool dg = (k == 0);
    const double u1 = 1.0 - w1, u2 = 1.0 - w2, w = w1 - w2;
    c.i0 = k; c.ip = k + 1; c.im = dg ? 0 : k - 1;
    c.ck = dg ? 1.0 - w : fma(u1, u2, w1 * w2);
    c.cp = dg ? w : w1 * u2;
    c.cm = dg ? 0.0 : u1 * w2;
    return c;
}
template <bool REAL>
__device__ __forceinline__ typename Num<REAL>::T cell3_apply_i(const double2* __restrict__ D, int stride, const GridCell3& c) {
    if constexpr (!REAL) {
        const double2 dk = __ldg(D + (size_t)c.i0 * stride), dp = __ldg(D + (size_t)c.ip * stride), dm = __ldg(D + (size_t)c.im * stride);
        return make_double2(-(c.ck * dk.y + c.cp * dp.y + c.cm * dm.y), c.ck * dk.x + c.cp * dp.x + c.cm * dm.x);
    } else {
        const double dk = __ldg(&D[(size_t)c.i0 * stride].y), dp = __ldg(&D[(size_t)c.ip * stride].y), dm = __ldg(&D[(size_t)c.im * stride].y);
        return -(c.ck * dk + c.cp * dp + c.cm * dm);
    }
}
// i * D(t_f, t_i) for a table D[k] = G(k h) with element stride `stride` (same operation order as
// grid_interp; the real mode works on the imaginary components only).
template <bool REAL>
__device__ __forceinline__ typename Num<REAL>::T cell_apply_i(const double2* __restrict__ D, int stride, const GridCell& c) {
    if constexpr (!REAL) {
        if (c.k == 0) {
            const double2 d0 = __ldg(D), d1 = __ldg(D + stride);
            return make_double2(-(d0.y + c.c00 * (d1.y - d0.y)), d0.x + c.c00 * (d1.x - d0.x));
        }
        const double2 dk = __ldg(D + (size_t)c.k * stride), dp = __ldg(D + (size_t)(c.k + 1) * stride),
                      dm = __ldg(D + (size_t)(c.k - 1) * stride);
        return make_double2(-(c.c00 * dk.y + c.c10 * dp.y + c.c01 * dm.y + c.c11 * dk.y),
                            c.c00 * dk.x + c.c10 * dp.x + c.c01 * dm.x + c.c11 * dk.x);
    } else {
        if (c.k == 0) {
            const double d0 = __ldg(&D[0].y), d1 = __ldg(&D[stride].y);
            return -(d0 + c.c00 * (d1 - d0));
        }
        const double dk = __ldg(&D[(size_t)c.k * stride].y), dp = __ldg(&D[(size_t)(c.k + 1) * stride].y),
                     dm = __ldg(&D[(size_t)(c.k - 1) * stride].y);
        return -(c.c00 * dk + c.c10 * dp + c.c01 * dm + c.c11 * dk);
    }
}

template <bool REAL>
__device__ __forceinline__ typename Num<REAL>::T delta_apply_i(const DevDelta& t, double t_f, double t_i) {
    if constexpr (!REAL) {
        return times_i(delta_eval(t, t_f, t_i));
    } else {
        if (t.kind == 1) {
            const double dt = t_f - t_i, h = t.h;
            int j = (int)floor(dt * t.inv_h);
            j = min(max(j, 0), t.n - 2);
            const double xa = dt - (double)j * h, xb = (double)(j + 1) * h - dt;
            const double y0 = __ldg(&t.y[j].y), y1 = __ldg(&t.y[j + 1].y), m0 = __ldg(&t.M[j].y), m1 = __ldg(&t.M[j + 1].y);
            const double i6h = t.inv_h * (1.0 / 6.0), h6 = h * (1.0 / 6.0), ih = t.inv_h;
            const double ca = xa * xa * xa * i6h, cb = xb * xb * xb * i6h;
            return -(m0 * cb + m1 * ca + (y0 * ih - m0 * h6) * xb + (y1 * ih - m1 * h6) * xa);
        }
        return cell_apply_i<true>(t.y, 1, grid_cell(t.n, t.inv_h, t_f, t_i));
    }
}

// Segment products (EntryProgram::segdef): lane = table entry, loop over the samples of a range; the
// entry's propagator slots live in registers.  Shorter segments are padded with the row's constant-one slot.
template <int LEN, bool REAL>
__device__ __forceinline__ void segment_products(const uint16_t* __restrict__ segdef_t, int n_seg, int out0,
                                                 unsigned char* Tb, int row_bytes, int spb, int warp, int nw, int lane) {
    typedef typename Num<REAL>::T T;
    typedef Num<REAL> N;
    constexpr int SH = REAL ? 3 : 4;
    const int n_groups = (n_seg + 31) >> 5;
    // tasks = (group of 32 entries) x (sample sub-range); split the samples so that every warp has work
    int split = 1;
    while (n_groups * split < nw && split < spb) split <<= 1;
    const int per = spb / split;
    for (int task = warp; task < n_groups * split; task += nw) {
        const int g = task / split, part = task - g * split;
        const int j = g * 32 + lane;
        uint32_t q[LEN];
        const uint16_t* sd = segdef_t + (size_t)g * LEN * 32 + lane;
#pragma unroll
        for (int i = 0; i < LEN; ++i) q[i] = (uint32_t)sd[i * 32] << SH;
        if (j >= n_seg) continue;
        unsigned char* row = Tb + (size_t)part * per * row_bytes;
#pragma unroll 2
        for (int smp = 0; smp < per; ++smp, row += row_bytes) {
            T v = *reinterpret_cast<const T*>(row + q[0]);
#pragma unroll
            for (int i = 1; i < LEN; ++i) v = N::mul(v, *reinterpret_cast<const T*>(row + q[i]));
            reinterpret_cast<T*>(row)[out0 + j] = v;
        }
    }
}

// ---- configuration walk ------------------------------------------------------------------------
// Every surviving configuration of an entry is a fixed-length record (qiw_host.hpp: EntryProgram::rec2)
// of K + order operand slots in the per-sample table: K segment products of propagators and one
// pair-interaction factor per arc; the operator matrix elements and the topology sign
// (-i * parity * (-1)^order, src/topology_eval.jl:431) are folded into the record's coefficient.
//
// Mapping: one LANE owns one configuration (its record lives in registers for the whole sample
// block) and loops over the samples of the CTA, whose tables sit in shared memory as T[sample][slot].
// Per factor that is one LDS feeding one multiply (real mode) or four FP64 instructions (complex).
template <int L, bool REAL, bool PER_SAMPLE>
__device__ __forceinline__ void config_walk(const uint32_t* __restrict__ rec_t, int g0, int g1, const unsigned char* Tb,
                                            int row_bytes, int spb, const typename Num<REAL>::T* coefs_s, double2* red,
                                            int nw, int warp, int lane, int S, double2* sample_out,
                                            unsigned long long local0, unsigned long long count) {
    typedef typename Num<REAL>::T T;
    typedef Num<REAL> N;
    constexpr int SH = REAL ? 3 : 4;
    for (int g = g0; g < g1; ++g) {
        uint32_t w[L + 1];
        const uint32_t* rp = rec_t + (size_t)g * (L + 1) * 32 + lane;
#pragma unroll
        for (int q = 0; q <= L; ++q) w[q] = __ldg(rp + q * 32);
#pragma unroll
        for (int q = 1; q <= L; ++q) w[q] <<= SH;
        const int s_i = (int)(w[0] >> 16);
        const T coef = coefs_s[w[0] & 0xFFFFu];
        const int smin = (int)__reduce_min_sync(0xFFFFFFFFu, (unsigned)s_i);
        const int smax = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)s_i);
        T acc0 = N::zero(), acc1 = N::zero();
        if constexpr (PER_SAMPLE) {
            for (int smp = 0; smp < spb; ++smp) {
                const unsigned char* row = Tb + smp * row_bytes;
                T va = *reinterpret_cast<const T*>(row + w[1]);
#pragma unroll
                for (int f = 2; f <= L; ++f) va = N::mul(va, *reinterpret_cast<const T*>(row + w[f]));
                // qiw_eval_at_times: the evaluator's value for every sample separately
                const double2 c = N::result(coef, va);
                for (int s = smin; s <= smax; ++s) {
                    double2 r = (s_i == s) ? c : make_double2(0.0, 0.0);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        r.x += __shfl_down_sync(0xFFFFFFFFu, r.x, off);
                        r.y += __shfl_down_sync(0xFFFFFFFFu, r.y, off);
                    }
                    if (lane == 0 && local0 + smp < count) {
                        double2* o = sample_out + (local0 + smp) * S + s;
                        *o = cadd(*o, r);
                    }
                }
            }
        } else {
            // four samples in flight per iteration: independent multiply chains
            int smp = 0;
            for (; smp + 3 < spb; smp += 4) {
                const unsigned char* r0 = Tb + smp * row_bytes;
                const unsigned char* r1 = r0 + row_bytes;
                const unsigned char* r2 = r1 + row_bytes;
                const unsigned char* r3 = r2 + row_bytes;
                T va = *reinterpret_cast<const T*>(r0 + w[1]);
                T vb = *reinterpret_cast<const T*>(r1 + w[1]);
                T vc = *reinterpret_cast<const T*>(r2 + w[1]);
                T vd = *reinterpret_cast<const T*>(r3 + w[1]);
#pragma unroll
                for (int f = 2; f <= L; ++f) {
                    va = N::mul(va, *reinterpret_cast<const T*>(r0 + w[f]));
                    vb = N::mul(vb, *reinterpret_cast<const T*>(r1 + w[f]));
                    vc = N::mul(vc, *reinterpret_cast<const T*>(r2 + w[f]));
                    vd = N::mul(vd, *reinterpret_cast<const T*>(r3 + w[f]));
                }
                acc0 = N::add(acc0, N::add(va, vc));
                acc1 = N::add(acc1, N::add(vb, vd));
            }
            for (; smp < spb; ++smp) {
                const unsigned char* r0 = Tb + smp * row_bytes;
                T va = *reinterpret_cast<const T*>(r0 + w[1]);
#pragma unroll
                for (int f = 2; f <= L; ++f) va = N::mul(va, *reinterpret_cast<const T*>(r0 + w[f]));
                acc0 = N::add(acc0, va);
            }
            // sum over the lanes' configurations, separately for every initial sector present
            const double2 c = N::result(coef, N::add(acc0, acc1));
            for (int s = smin; s <= smax; ++s) {
                double2 r = (s_i == s) ? c : make_double2(0.0, 0.0);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    if constexpr (!REAL) r.x += __shfl_down_sync(0xFFFFFFFFu, r.x, off);
                    r.y += __shfl_down_sync(0xFFFFFFFFu, r.y, off);
                }
                if (lane == 0) red[s * nw + warp] = cadd(red[s * nw + warp], r);
            }
        }
    }
}

// Paired records (EntryProgram::rec_pair): two configurations with identical pair-interaction operands and
// initial sector,  prod(Delta) * (coefA * prod(segA) + coefB * prod(segB)):  ND + 2K operands instead of 2 (ND + K).
template <int ND, int K, bool REAL>
__device__ __forceinline__ void pair_walk(const uint32_t* __restrict__ rec_t, int g0, int g1, const unsigned char* Tb,
                                          int row_bytes, int spb, const typename Num<REAL>::T* coefs_s, double2* red,
                                          int nw, int warp, int lane) {
    typedef typename Num<REAL>::T T;
    typedef Num<REAL> N;
    constexpr int SH = REAL ? 3 : 4;
    constexpr int RL = 2 + 2 * K + ND;
    for (int g = g0; g < g1; ++g) {
        uint32_t w[RL];
        const uint32_t* rp = rec_t + (size_t)g * RL * 32 + lane;
#pragma unroll
        for (int q = 0; q < RL; ++q) w[q] = __ldg(rp + q * 32);
#pragma unroll
        for (int q = 2; q < RL; ++q) w[q] <<= SH;
        const int s_i = (int)(w[0] >> 16);
        const T coefA = coefs_s[w[0] & 0xFFFFu], coefB = coefs_s[w[1] & 0xFFFFu];
        const int smin = (int)__reduce_min_sync(0xFFFFFFFFu, (unsigned)s_i);
        const int smax = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)s_i);
        T accA = N::zero(), accB = N::zero();
        // four samples in flight per iteration (spb is a power of two; blocks smaller than 4 take the scalar loop)
        int smp = 0;
        for (; smp + 3 < spb; smp += 4) {
            const unsigned char* row = Tb + smp * row_bytes;
            T d[4], sa[4], sb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned char* r = row + u * row_bytes;
                d[u] = *reinterpret_cast<const T*>(r + w[2 + 2 * K]);
                sa[u] = *reinterpret_cast<const T*>(r + w[2]);
                sb[u] = *reinterpret_cast<const T*>(r + w[2 + K]);
            }
#pragma unroll
            for (int f = 1; f < ND; ++f)
#pragma unroll
                for (int u = 0; u < 4; ++u) d[u] = N::mul(d[u], *reinterpret_cast<const T*>(row + u * row_bytes + w[2 + 2 * K + f]));
#pragma unroll
            for (int f = 1; f < K; ++f)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    sa[u] = N::mul(sa[u], *reinterpret_cast<const T*>(row + u * row_bytes + w[2 + f]));
                    sb[u] = N::mul(sb[u], *reinterpret_cast<const T*>(row + u * row_bytes + w[2 + K + f]));
                }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                accA = N::add(accA, N::mul(d[u], sa[u]));
                accB = N::add(accB, N::mul(d[u], sb[u]));
            }
        }
        for (; smp < spb; ++smp) {
            const unsigned char* row = Tb + smp * row_bytes;
            T d = *reinterpret_cast<const T*>(row + w[2 + 2 * K]);
#pragma unroll
            for (int f = 1; f < ND; ++f) d = N::mul(d, *reinterpret_cast<const T*>(row + w[2 + 2 * K + f]));
            T sa = *reinterpret_cast<const T*>(row + w[2]);
            T sb = *reinterpret_cast<const T*>(row + w[2 + K]);
#pragma unroll
            for (int f = 1; f < K; ++f) {
                sa = N::mul(sa, *reinterpret_cast<const T*>(row + w[2 + f]));
                sb = N::mul(sb, *reinterpret_cast<const T*>(row + w[2 + K + f]));
            }
            accA = N::add(accA, N::mul(d, sa));
            accB = N::add(accB, N::mul(d, sb));
        }
        const double2 c = cadd(N::result(coefA, accA), N::result(coefB, accB));
        for (int s = smin; s <= smax; ++s) {
            double2 r = (s_i == s) ? c : make_double2(0.0, 0.0);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                if constexpr (!REAL) r.x += __shfl_down_sync(0xFFFFFFFFu, r.x, off);
                r.y += __shfl_down_sync(0xFFFFFFFFu, r.y, off);
            }
            if (lane == 0) red[s * nw + warp] = cadd(red[s * nw + warp], r);
        }
    }
}

template <bool REAL>
__device__ __forceinline__ void pair_dispatch(int nd, int K, const uint32_t* rec_t, int g0, int g1, const unsigned char* Tb,
                                              int row_bytes, int spb, const typename Num<REAL>::T* coefs_s, double2* red,
                                              int nw, int warp, int lane) {
#define QIW_PC(D_, K_) case (D_ * 8 + K_): pair_walk<D_, K_, REAL>(rec_t, g0, g1, Tb, row_bytes, spb, coefs_s, red, nw, warp, lane); break;
#define QIW_PK(D_) QIW_PC(D_, 1) QIW_PC(D_, 2) QIW_PC(D_, 3) QIW_PC(D_, 4)
    switch (nd * 8 + K) {
        QIW_PK(1) QIW_PK(2) QIW_PK(3) QIW_PK(4) QIW_PK(5) QIW_PK(6) QIW_PK(7) QIW_PK(8)
        default: break;
    }
#undef QIW_PK
#undef QIW_PC
}

// Record lengths: K + order operands, K <= 4 segments, order <= 8.
template <bool REAL, bool PER_SAMPLE>
__device__ __forceinline__ void walk_dispatch(int L, const uint32_t* rec_t, int g0, int g1, const unsigned char* Tb,
                                              int row_bytes, int spb, const typename Num<REAL>::T* coefs_s, double2* red,
                                              int nw, int warp, int lane, int S, double2* so, unsigned long long local0,
                                              unsigned long long count) {
#define QIW_CASE(N_) case N_: config_walk<N_, REAL, PER_SAMPLE>(rec_t, g0, g1, Tb, row_bytes, spb, coefs_s, red, nw, warp, lane, S, so, local0, count); break;
    switch (L) {
        QIW_CASE(1) QIW_CASE(2) QIW_CASE(3) QIW_CASE(4) QIW_CASE(5) QIW_CASE(6) QIW_CASE(7) QIW_CASE(8)
        QIW_CASE(9) QIW_CASE(10) QIW_CASE(11) QIW_CASE(12)
        default: break;
    }
#undef QIW_CASE
}

// ---- fused tail of the step kernel ---------------------------------------------------------------
// out = weight * (-i)^d * Jacobian * sum(rows): the factors of contour_integral / qmc_integral
// (src/qmc_integrate.jl:497-507,565-569,597-612) and of the simplex maps (:46,458-463).
__device__ __forceinline__ double simplex_volume(int d, double edge) {
    // prod_{i<=d} edge / i (src/qmc_integrate.jl:46) with the reciprocals tabulated: FP64 division costs
    // hundreds of cycles and this sits on the critical path of every step's tail
    const double inv[17] = {1.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10,
                            1.0 / 11, 1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16};
    double v = 1.0;
#pragma unroll
    for (int i = 1; i <= 16; ++i)
        if (i <= d) v *= edge * inv[i];
    return v;
}

__device__ __forceinline__ double entry_scale(const DevEntry& e, const DevEntryDyn& dy, double t_i, double t_w, double t_f) {
    if (e.exact) return dy.weight;
    const double jac = (e.mode == 0) ? simplex_volume(e.D, t_f - t_i)
                                     : simplex_volume(e.d_before, t_w - t_i) * simplex_volume(e.d_after, t_f - t_w);
    const double dir = (e.order & 1) ? -1.0 : 1.0;   // (-i)^(2 order)
    return dir * jac * dy.weight;
}

// Executed by the last CTA of a step launch: every (entry, sector) sum runs over the partial rows in
// fixed order, so the result does not depend on which CTA happens to be last.  Optionally followed by
// set_ppgf!(P, tau_f, result) and normalize!(P, tau_f) (src/ppgf.jl:495-504,646-668).
__device__ void fused_tail(const StepParams& pp, double t_i, double t_w, double t_f, int pitch) {
    StepParams p = pp;
    const int S = p.S, n_out = p.n_call_entries * S;
    p.partials += (size_t)blockIdx.z * gridDim.y * gridDim.x * S;   // this time triple's rows and results
    p.out += (size_t)blockIdx.z * n_out;
    // four lanes per (entry, sector) output: lane g adds rows g, g+4, ... in order, then the four partial sums
    // are combined in a fixed butterfly — the same operation order whichever CTA runs the tail
    for (int o0 = 0; o0 < n_out; o0 += (int)blockDim.x / 4) {
        const int o = o0 + (int)threadIdx.x / 4, g = (int)threadIdx.x & 3;
        double2 v = make_double2(0.0, 0.0);
        double scale = 0.0;
        int oi = 0;
        if (o < n_out) {
            const int i = o / S, s = o - i * S;
            const DevEntryDyn& dy = p.dyn[i];
            const size_t row0 = (size_t)dy.item0 * pitch, nrows = (size_t)dy.n_items * pitch;
            for (size_t r = g; r < nrows; r += 4) v = cadd(v, __ldcg(p.partials + (row0 + r) * S + s));
            if (g == 0) scale = entry_scale(p.entries[dy.entry], dy, t_i, t_w, t_f);
            oi = dy.out_index * S + s;
        }
        v.x += __shfl_xor_sync(0xFFFFFFFFu, v.x, 1); v.y += __shfl_xor_sync(0xFFFFFFFFu, v.y, 1);
        v.x += __shfl_xor_sync(0xFFFFFFFFu, v.x, 2); v.y += __shfl_xor_sync(0xFFFFFFFFu, v.y, 2);
        if (o < n_out && g == 0) p.out[oi] = cscale(scale, v);
    }
    if (p.peer_ranks > 1) {
        // ---- all-reduce over peer memory (replaces all_reduce!, src/mpi.jl:104-127) ----
        // Low-latency protocol: every 8-byte store carries 4 bytes of payload and the 4-byte sequence number of
        // this collective, so the receiver needs no separate flag and the sender no system-wide fence: a word is
        // valid as soon as its flag matches (8-byte stores are single transactions on NVLink).  Each double
        // travels as two such words.  Buffers alternate with the parity of the sequence number: a slot is
        // rewritten two collectives later, after every peer has provably finished reading it.
        __syncthreads();
        const int par = (int)(p.peer_seq & 1ull);
        const unsigned int seq32 = (unsigned int)(p.peer_seq % 0xFFFFFFFFull) + 1u;   // never 0 (the mailbox starts zeroed)
        const size_t my_slot = kPeerFlagBytes + ((size_t)p.peer_rank * 2 + par) * kPeerSlotBytes;
        const int n_dbl = 2 * n_out, n_words = 2 * n_dbl;
        const double* outd = reinterpret_cast<const double*>(p.out);
        for (int k = threadIdx.x; k < n_words; k += blockDim.x) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(outd[k >> 1]);
            const unsigned int half = (k & 1) ? (unsigned int)(bits >> 32) : (unsigned int)bits;
            const uint2 wd = make_uint2(half, seq32);
            for (int q = 0; q < p.peer_ranks; ++q) {
                if (q == p.peer_rank) continue;
                uint2* dst = reinterpret_cast<uint2*>(p.peer_mail[q] + my_slot) + k;
                asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(wd.x), "r"(wd.y) : "memory");
            }
        }
        // receive: poll every word until its flag shows this collective, add the contributions in rank order
        const unsigned char* base = p.peer_mail[p.peer_rank] + kPeerFlagBytes;
        const unsigned long long t0 = globaltimer_ns();
        for (int j = threadIdx.x; j < n_dbl; j += blockDim.x) {
            double v = 0.0;
            for (int q = 0; q < p.peer_ranks; ++q) {
                if (q == p.peer_rank) { v += outd[j]; continue; }
                const uint2* src = reinterpret_cast<const uint2*>(base + ((size_t)q * 2 + par) * kPeerSlotBytes) + 2 * j;
                uint2 lo, hi;
                bool ok = true;
                do {
                    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(lo.x), "=r"(lo.y) : "l"(src) : "memory");
                    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(hi.x), "=r"(hi.y) : "l"(src + 1) : "memory");
                    if ((lo.y != seq32 || hi.y != seq32) && globaltimer_ns() - t0 > 10000000000ull) { *p.peer_status = 1; ok = false; break; }   // 10 s
                } while (lo.y != seq32 || hi.y != seq32);
                if (ok) v += __longlong_as_double((long long)(((unsigned long long)hi.x << 32) | (unsigned long long)lo.x));
            }
            reinterpret_cast<double*>(p.out)[j] = v;   // element j is read and written by this thread only
        }
    }
    if (p.finish_k_f < 0) return;
    __syncthreads();
    __shared__ double lambda_s;
    const int bsize = p.bsize, k_f = p.finish_k_f;
    double2* P = p.finish_P;
    for (int el = threadIdx.x; el < bsize; el += blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        for (int j = 0; j < p.n_call_entries; ++j) {
            const double2 c = p.out[(size_t)j * bsize + el];
            v = cadd(v, c);
            if (p.finish_hist) p.finish_hist[(size_t)j * bsize + el] = c;
        }
        P[(size_t)k_f * bsize + el] = v;
    }
    __syncthreads();
    if (!p.finish_normalize) return;
    if (threadIdx.x == 0) {
        double pmax = -1.0e300;
        for (int i = 0; i < p.finish_n_diag; ++i) pmax = fmax(pmax, -P[(size_t)k_f * bsize + p.finish_diag[i]].y);
        lambda_s = log(pmax) / ((double)k_f * p.h);
    }
    __syncthreads();
    const double lambda = lambda_s;
    for (int idx = threadIdx.x; idx < p.n_tau * bsize; idx += blockDim.x) {
        const int k = idx / bsize;
        const double f = exp(-((double)k * p.h) * lambda);
        P[idx] = cscale(f, P[idx]);
    }
}

// ---- the step kernel (scalar models: every sector block is 1x1) ------------------------------

// PAIRS: the launch contains entries with paired records (their walk code is only in this instantiation, so
// that the common orders <= 4 case keeps a small instruction footprint — the N = 2^10 step is latency-bound).
template <bool REAL, bool PER_SAMPLE, bool PAIRS>
__global__ void __launch_bounds__(256, 3) scalar_step_kernel(const StepParams p) {
    typedef typename Num<REAL>::T T;
    typedef Num<REAL> N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // let the next step's grid start as soon as every CTA of this one is running (it waits before it touches P)
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, nthr = blockDim.x;
    const WorkItem it = p.items[blockIdx.y];
    const DevEntry& e = p.entries[it.entry];
    const DevEntryDyn& dy = p.dyn[it.slot];
    const int S = p.S, D = e.D, n_nodes = e.n_nodes, nP = e.nP, nD = e.nD, nSeg = e.nSeg;
    const int n_slots = nP + nD + nSeg + 1;   // + the constant-one slot that pads short segment products
    const int d_after = e.d_after;
    const int spb = p.spb, spb_sh = p.spb_log2, spb_mask = spb - 1;   // spb is a power of two
    // shared memory carve-up (sizes fixed per launch from the largest entry, see host):
    // T[spb samples][row] with an odd row pitch (in operand units) so that the lanes' stores during
    // the fill hit different banks
    const int row_bytes = (n_slots | 1) * (int)sizeof(T);
    unsigned char* Tb = smem_raw;                                                            // [spb][max_row]
    double2* red = reinterpret_cast<double2*>(smem_raw + (size_t)p.max_slots * spb * sizeof(T));   // [S][nw]
    // The small per-sample arrays are sized for the launch's largest entry (p.max_nodes1 = positions + 1), and the
    // roots pw[D][32] — dead once the times exist — live at the start of the table T, which is filled afterwards:
    // 63 instead of 73 KB per CTA for the README model, so three CTAs leave 60 KB of the SM to L1 instead of 28.
    double* times = reinterpret_cast<double*>(red + (size_t)S * nw);             // [max_nodes1][32]
    double* pw = reinterpret_cast<double*>(Tb);                                  // [D][32], aliases T (phases 1-2 only)
    double* cellw = times + p.max_nodes1 * 32;                                   // [max_nodes1][32] fractional cell weight
    int* cella = reinterpret_cast<int*>(cellw + p.max_nodes1 * 32);              // [max_nodes1][32] grid cell of each time
    int* okflag = cella + p.max_nodes1 * 32;                                     // [32]
    int4* dslots_s = reinterpret_cast<int4*>(okflag + 32);                       // [max_dslots]
    T* coefs_s = reinterpret_cast<T*>(dslots_s + p.max_dslots);                  // [max_coefs + 1]
    const int seg_stride = e.seg_stride;
    for (int c = threadIdx.x; c < nD; c += nthr) dslots_s[c] = e.dslots[c];
    for (int c = threadIdx.x; c < S * nw; c += nthr) red[c] = make_double2(0.0, 0.0);
    for (int c = threadIdx.x; c <= e.n_coefs; c += nthr)
        coefs_s[c] = (c < e.n_coefs) ? N::coef_of(e.coefs[c]) : N::zero();   // last: padding records

    double t_i = p.t_i, t_w = p.t_w, t_f = p.t_f;
    // batched evaluation: blockIdx.z selects one (t_i, t_w, t_f) triple of the call
    if (p.times_dev) { const double* tz = p.times_dev + 3 * blockIdx.z; t_i = tz[0]; t_w = tz[1]; t_f = tz[2]; }
    const double lo_after = (e.mode == 0) ? t_i : t_w, len_after = t_f - lo_after;
    const double len_before = t_w - t_i;

    // randomised qMC: blockIdx.z selects one of several scrambled sequences (no root cache then)
    const uint32_t* __restrict__ sm = dy.sobol + (size_t)blockIdx.z * p.sobol_z_stride;
    const unsigned long long count = dy.count;
    const int n_sb = (int)((count + (unsigned long long)spb - 1ull) / (unsigned long long)spb);

    // this warp's share of the entry's configuration groups (32 records per group): all records for the
    // per-sample evaluation, pairs + leftovers for the summing walk
    int g0 = 0, g1 = 0, pg0 = 0, pg1 = 0;
    if (warp < it.n_chunks) {
        const long long c = it.chunk0 + warp, nct = it.n_chunks_total;
        const long long ng = PAIRS ? e.n_groups_left : e.n_groups, npg = PAIRS ? e.n_groups_pair : 0;
        g0 = (int)(c * ng / nct);
        g1 = (int)((c + 1) * ng / nct);
        // pairs are dealt in the opposite direction so that a warp short on pairs gets more leftovers
        pg0 = (int)((nct - 1 - c) * npg / nct);
        pg1 = (int)((nct - c) * npg / nct);
    }

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
        trace[0] = clock64(); trace[4] = smid; trace[5] = it.entry; trace[6] = (unsigned long long)(g1 - g0);
    }

    for (int sb = blockIdx.x; sb < n_sb; sb += gridDim.x) {
        const unsigned long long local0 = (unsigned long long)sb * (unsigned long long)spb;

        // -- 1. Sobol coordinates and the independent roots x_j^(1/(remaining dims)) ----------
        //       The roots depend only on (entry, Sobol sequence, sample), not on the time step: when the
        //       host provides a cache they are computed once per run and re-read afterwards.
        if ((int)threadIdx.x < spb) okflag[threadIdx.x] = (local0 + threadIdx.x < count) ? 1 : 0;
        if (p.explicit_times == nullptr) {
            double* uc = p.sobol_z_stride ? nullptr : dy.ucache;
            for (int task = threadIdx.x; task < D * spb; task += nthr) {
                const int j = task >> spb_sh, smp = task & spb_mask;
                const unsigned long long local = local0 + smp;
                const bool active = local < count;
                double r;
                if (uc && dy.ucache_valid) {
                    r = active ? uc[(size_t)j * count + local] : 0.0;
                } else {
                    const uint32_t xi = sobol_coord(sm + j * 32, __ldg(sm + D * 32 + j), (uint32_t)(dy.start + local));
                    const double x = (double)xi * 2.3283064365386963e-10;  // ldexp(x, -32), exact
                    const int den = (j < d_after) ? (d_after - j) : (D - j);
                    r = (den == 1) ? x : pow(x, 1.0 / (double)den);
                    if (uc && active && it.chunk0 == 0) uc[(size_t)j * count + local] = r;
                }
                pw[j * 32 + smp] = r;
            }
        }
        __syncthreads();
        if (trace && threadIdx.x == 0) trace[8] = clock64();

        // -- 2. ordered times of every backbone position: thread = (position, sample).  The running
        //       product u_j = ((r_0 r_1) r_2) ... r_j is re-evaluated from the start of its simplex so that
        //       positions are independent (same operation order as the sequential map, bit-identical).
        for (int task = threadIdx.x; task < n_nodes * spb; task += nthr) {
            const int pos = 1 + (task >> spb_sh), smp = task & spb_mask;
            const int src = e.pos_src[pos];
            double t;
            if (src == -1) t = t_i;
            else if (src == -2) t = t_w;
            else if (src == -3) t = t_f;
            else if (p.explicit_times) {
                t = (local0 + smp < count) ? p.explicit_times[(local0 + smp) * D + src] : 0.0;
            } else {
                const int j0 = (src < d_after) ? 0 : d_after;
                double u = pw[j0 * 32 + smp];
                for (int j = j0 + 1; j <= src; ++j) u = __dmul_rn(u, pw[j * 32 + smp]);
                if (src < d_after) t = __dadd_rn(__dmul_rn(u, len_after), lo_after);
                else t = __dadd_rn(__dmul_rn(u, len_before), t_i);
                if (!(t >= 0.0)) okflag[smp] = 0;   // all(refs .>= 0) (src/qmc_integrate.jl:608)
            }
            times[pos * 32 + smp] = t;
            // grid cell and weight of this time on the P grid (Keldysh.jl rule, shared by all slots)
            const double q = t * p.inv_h;
            const int a = min(max(__double2int_rd(q), 0), p.n_tau - 2);
            cella[pos * 32 + smp] = a;
            cellw[pos * 32 + smp] = q - (double)a;
        }
        __syncthreads();
        if (trace && threadIdx.x == 0) trace[9] = clock64();

        // -- 3. per-sample tables.  Propagators: thread = (backbone interval, sample) evaluates all
        //       sectors; pair interactions: thread = (slot, sample).  Discarded samples
        //       (src/qmc_integrate.jl:503,608) and samples past the range get zero rows.
        {
            const int nI = n_nodes - 1;
            // pair interactions first: they do not depend on the bold propagators, so in the device-resident
            // loop this part (like everything above) overlaps the previous step's tail (see below)
            if (p.tables_on_grid) {   // every Delta table is a plain grid function on the P grid: branch-free, cells reused
#pragma unroll 4
                for (int task = threadIdx.x; task < nD * spb; task += nthr) {
                    const int q = task >> spb_sh, smp = task & spb_mask;
                    const int4 ds = dslots_s[q];
                    const int ih = ds.y * 32 + smp;
                    const int it2 = (times[ds.x * 32 + smp] <= times[ih]) ? ih : ds.x * 32 + smp;   // clamp (:407-410)
                    const DevDelta& dt = p.deltas_inline[ds.z];
                    const T val = cell3_apply_i<REAL>(dt.y, 1, grid_cell3_from(cella[it2], cellw[it2], cella[ih], cellw[ih]));
                    reinterpret_cast<T*>(Tb + smp * row_bytes)[nP + q] = okflag[smp] ? val : N::zero();
                }
            } else {
                for (int task = threadIdx.x; task < nD * spb; task += nthr) {
                    const int q = task >> spb_sh, smp = task & spb_mask;
                    const bool ok = okflag[smp] != 0;
                    T* myrow = reinterpret_cast<T*>(Tb + smp * row_bytes);
                    const int4 ds = dslots_s[q];
                    const double th = times[ds.y * 32 + smp];
                    double tt = times[ds.x * 32 + smp];
                    if (tt < th) tt = th;                       // :407-410
                    const DevDelta& dt = ds.z < kInlineTables ? p.deltas_inline[ds.z] : p.deltas[ds.z];
                    T val;
                    if (dt.kind == 0 && dt.n == p.n_tau && dt.inv_h == p.inv_h) {   // table on the P grid: reuse the cells
                        const int ih = ds.y * 32 + smp, it2 = (tt == th) ? ih : ds.x * 32 + smp;
                        val = cell_apply_i<REAL>(dt.y, 1, grid_cell_from(cella[it2], cellw[it2], cella[ih], cellw[ih]));
                    } else {
                        val = delta_apply_i<REAL>(dt, tt, th);
                    }
                    myrow[nP + q] = ok ? val : N::zero();
                }
            }
            // Programmatic dependent launch: this grid may have been started while the previous step's grid
            // was still finishing.  Everything up to here used only data no kernel writes; the P table, the
            // partial-sum rows and the arrival counter belong to the previous grid until it has completed.
            asm volatile("griddepcontrol.wait;" ::: "memory");
            for (int task = threadIdx.x; task < nI * spb; task += nthr) {
                const int q = task >> spb_sh, smp = task & spb_mask;
                const bool ok = okflag[smp] != 0;
                T* myrow = reinterpret_cast<T*>(Tb + smp * row_bytes);
                {
                    const double ta = times[(q + 1) * 32 + smp];
                    double tb = times[(q + 2) * 32 + smp];
                    if (tb < ta) tb = ta;                       // src/topology_eval.jl:362-364
                    if (q == 0) myrow[n_slots - 1] = N::from_real(1.0);
                    if (e.mode == 0) {
                        for (int s = 0; s < S; ++s) {           // bare: i * (-i) exp(-dt (E + lambda))
                            const T val = N::from_real(exp(-(tb - ta) * __ldg(p.E + s)));
                            myrow[q * S + s] = ok ? val : N::zero();
                        }
                    } else {
                        const bool sw = times[(q + 2) * 32 + smp] < ta;   // clamped: both ends in the earlier time's cell
                        const int ia = (q + 1) * 32 + smp, ib = sw ? ia : ia + 32;
                        const GridCell3 cell = grid_cell3_from(cella[ib], cellw[ib], cella[ia], cellw[ia]);
#pragma unroll 4
                        for (int s = 0; s < S; ++s) {
                            const T val = cell3_apply_i<REAL>(p.P + s, p.bsize, cell);
                            myrow[q * S + s] = ok ? val : N::zero();
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (trace && threadIdx.x == 0) trace[10] = clock64();

        // -- 4. segment products -----------------------------------------------------------------------
        switch (seg_stride) {
#define QIW_SEG(N_) case N_: segment_products<N_, REAL>(e.segdef, nSeg, nP + nD, Tb, row_bytes, spb, warp, nw, lane); break;
            QIW_SEG(1) QIW_SEG(2) QIW_SEG(3) QIW_SEG(4) QIW_SEG(5) QIW_SEG(6) QIW_SEG(7) QIW_SEG(8) QIW_SEG(9)
#undef QIW_SEG
            default: break;
        }
        __syncthreads();
        if (trace && threadIdx.x == 0) trace[1] = clock64();

        // -- 5. this warp's configurations -----------------------------------------------------
        if constexpr (PAIRS) {
            if (pg0 < pg1) pair_dispatch<REAL>(e.order, e.K, e.records_pair, pg0, pg1, Tb, row_bytes, spb, coefs_s, red, nw, warp, lane);
        }
        if (g0 < g1) {
            walk_dispatch<REAL, PER_SAMPLE>(e.L2, PAIRS ? e.records_left : e.records, g0, g1, Tb, row_bytes, spb, coefs_s, red,
                                            nw, warp, lane, S, p.per_sample_out, local0, count);
        }
        __syncthreads();
        if (trace && threadIdx.x == 0) trace[2] = clock64();
    }

    if constexpr (PER_SAMPLE) return;

    // -- 6. CTA result: warps summed in fixed order (CTAs without a sample block have not waited yet) --
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (warp == 0) {
        // one load per lane and a fixed butterfly over the warps' sums instead of a chain of dependent loads
        // (the shared-memory pipe is busy with the other CTAs' walks at this point); nw is a power of two <= 8
        const int per = 32 / nw;                       // sectors per pass
        for (int s0 = 0; s0 < S; s0 += per) {
            const int s = s0 + lane / nw;
            double2 v = (s < S) ? red[s * nw + (lane % nw)] : make_double2(0.0, 0.0);
            for (int off = nw >> 1; off > 0; off >>= 1) {
                v.x += __shfl_xor_sync(0xFFFFFFFFu, v.x, off);
                v.y += __shfl_xor_sync(0xFFFFFFFFu, v.y, off);
            }
            if (s < S && (lane % nw) == 0)
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

}  // namespace qiw

namespace qiw {

// ---- deterministic reduction of the per-CTA partial sums -------------------------------------
// One CTA per entry of the call; rows of an entry are consecutive.
// out = weight * (-i)^d * Jacobian * sum(rows): the factors of contour_integral / qmc_integral
// (src/qmc_integrate.jl:497-507,565-569,597-612) and of the simplex maps (:46,458-463).
__global__ void __launch_bounds__(128) reduce_partials_kernel(const DevEntryDyn* __restrict__ dyn,
                                                              const DevEntry* __restrict__ entries,
                                                              const double2* __restrict__ partials, int pitch, int S,
                                                              double t_i, double t_w, double t_f,
                                                              double2* __restrict__ out) {
    __shared__ double2 buf[128];
    const DevEntryDyn& dy = dyn[blockIdx.x];
    const DevEntry& e = entries[dy.entry];
    const double scale = entry_scale(e, dy, t_i, t_w, t_f);
    const size_t row0 = (size_t)dy.item0 * pitch, nrows = (size_t)dy.n_items * pitch;
    for (int s = 0; s < S; ++s) {
        double2 v = make_double2(0.0, 0.0);
        for (size_t r = threadIdx.x; r < nrows; r += blockDim.x) v = cadd(v, partials[(row0 + r) * S + s]);
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int off = 64; off > 0; off >>= 1) {
            if ((int)threadIdx.x < off) buf[threadIdx.x] = cadd(buf[threadIdx.x], buf[threadIdx.x + off]);
            __syncthreads();
        }
        if (threadIdx.x == 0) out[(size_t)dy.out_index * S + s] = cscale(scale, buf[0]);
        __syncthreads();
    }
}

// ---- per-step state update on the device ---------------------------------------------------------
// set_ppgf!(P, tau_i, tau_f, result) followed by normalize!(P, tau_f) (src/ppgf.jl:495-504,646-668):
// P(tau_f) <- sum of the entries' contributions; lambda = log(max_s max diag(-Im P_s(tau_f))) / tau_f;
// every stored grid value is multiplied by exp(-lambda tau_k).  One CTA; the table is tiny.
__global__ void __launch_bounds__(256) finish_step_kernel(double2* __restrict__ P, int n_tau, int bsize,
                                                          const int* __restrict__ diag, int n_diag, double h, int k_f,
                                                          const double2* __restrict__ contribs, int n_contrib,
                                                          int do_normalize, double2* __restrict__ hist) {
    __shared__ double lambda_s;
    for (int e = threadIdx.x; e < bsize; e += blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        for (int j = 0; j < n_contrib; ++j) {
            const double2 c = contribs[(size_t)j * bsize + e];
            v = cadd(v, c);
            if (hist) hist[(size_t)j * bsize + e] = c;
        }
        P[(size_t)k_f * bsize + e] = v;
    }
    __syncthreads();
    if (!do_normalize) return;
    if (threadIdx.x == 0) {
        double pmax = -1.0e300;
        for (int i = 0; i < n_diag; ++i) pmax = fmax(pmax, -P[(size_t)k_f * bsize + diag[i]].y);
        lambda_s = log(pmax) / ((double)k_f * h);
    }
    __syncthreads();
    const double lambda = lambda_s;
    for (int idx = threadIdx.x; idx < n_tau * bsize; idx += blockDim.x) {
        const int k = idx / bsize;
        const double f = exp(-((double)k * h) * lambda);
        P[idx] = cscale(f, P[idx]);
    }
}

// ---- FP64 FMA peak probe ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0,
           a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---- host-callable launchers -----------------------------------------------------------------

template <bool REAL, bool PER_SAMPLE, bool PAIRS>
static cudaError_t launch_scalar_t(const StepParams& p, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(scalar_step_kernel<REAL, PER_SAMPLE, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    // programmatic dependent launch: consecutive step kernels of a stream may overlap prologue and tail
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = p.allow_overlap ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, scalar_step_kernel<REAL, PER_SAMPLE, PAIRS>, p);
}

// `real_mode`: every table and coefficient in use has been verified purely imaginary by the host.
// `pairs`: some entry of the launch has paired records (then every entry is walked as pairs + leftovers).
cudaError_t launch_scalar_step(bool real_mode, bool pairs, const StepParams& p, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    if (p.per_sample_out) {
        return real_mode ? launch_scalar_t<true, true, false>(p, grid, threads, smem, st) : launch_scalar_t<false, true, false>(p, grid, threads, smem, st);
    }
    if (pairs) return real_mode ? launch_scalar_t<true, false, true>(p, grid, threads, smem, st) : launch_scalar_t<false, false, true>(p, grid, threads, smem, st);
    return real_mode ? launch_scalar_t<true, false, false>(p, grid, threads, smem, st) : launch_scalar_t<false, false, false>(p, grid, threads, smem, st);
}

cudaError_t launch_reduce(const DevEntryDyn* dyn, const DevEntry* entries, const double2* partials, int pitch, int S,
                          double t_i, double t_w, double t_f, double2* out, int n_entries, cudaStream_t st) {
    reduce_partials_kernel<<<n_entries, 128, 0, st>>>(dyn, entries, partials, pitch, S, t_i, t_w, t_f, out);
    return cudaGetLastError();
}

cudaError_t launch_finish_step(double2* P, int n_tau, int bsize, const int* diag, int n_diag, double h, int k_f,
                               const double2* contribs, int n_contrib, int do_normalize, double2* hist, cudaStream_t st) {
    finish_step_kernel<<<1, 256, 0, st>>>(P, n_tau, bsize, diag, n_diag, h, k_f, contribs, n_contrib, do_normalize, hist);
    return cudaGetLastError();
}

cudaError_t launch_sobol_points(int D, const uint32_t* m, const uint32_t* x0, unsigned long long start,
                                unsigned long long count, uint32_t* out, cudaStream_t st) {
    const unsigned long long n = count * (unsigned long long)D;
    sobol_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(D, m, x0, start, count, out);
    return cudaGetLastError();
}

cudaError_t launch_dfma_peak(double* out, int blocks, int iters, cudaStream_t st) {
    dfma_peak_kernel<<<blocks, 256, 0, st>>>(out, iters);
    return cudaGetLastError();
}

}  // namespace qiw

namespace qiw {

// ---- the step kernel for sector blocks larger than 1x1 --------------------------------------------
// One THREAD owns one (sample, chunk of configuration trees) and replays the pruned tree of
// src/topology_eval.jl:454-556 with the reference's own prefix sharing: the running product
// A_pos ... A_1 (src/utility.jl:234-323) is kept on a per-thread stack, one (d x d_init) matrix per tree
// level.  Node matrix = operator block times i P_s(t_pos, t_pos-1) (:376-388).  Blocks are at most
// kMaxBlockDim x kMaxBlockDim.  This path favours generality over speed: all BASELINE headline
// configurations have 1x1 blocks and use the scalar kernel.
constexpr int kMaxBlockDim = 4;

__device__ __forceinline__ double2 zero2() { return make_double2(0.0, 0.0); }
// tree word fields (qiw_host.hpp make_word)
__device__ __forceinline__ uint32_t w_slotA(uint64_t w) { return (uint32_t)w & 0xFFFu; }
__device__ __forceinline__ uint32_t w_slotB(uint64_t w) { return ((uint32_t)w >> 12) & 0xFFFu; }
__device__ __forceinline__ uint32_t w_nchild(uint64_t w) { return ((uint32_t)w >> 24) & 0xFFu; }
__device__ __forceinline__ uint32_t w_aux(uint64_t w) { return (uint32_t)(w >> 32) & 0xFFFFu; }

__global__ void __launch_bounds__(64) block_step_kernel(const StepParams p, const BlockParams bp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const WorkItem it = p.items[blockIdx.y];
    const DevEntry& e = p.entries[it.entry];
    const DevEntryDyn& dy = p.dyn[it.slot];
    const DevModel& m = bp.m;
    const int S = m.S, bsize = m.bsize, D = e.D, n_nodes = e.n_nodes, d_after = e.d_after;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* red = reinterpret_cast<double2*>(smem_raw);   // [nw][bsize]

    double t_i = p.t_i, t_w = p.t_w, t_f = p.t_f;
    const double lo_after = (e.mode == 0) ? t_i : t_w, len_after = t_f - lo_after, len_before = t_w - t_i;
    const unsigned long long count = dy.count;
    const int n_sb = (int)((count + blockDim.x - 1) / blockDim.x);
    const size_t tid_global = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    double2* iP = bp.scratch + tid_global * bp.scratch_per_thread;        // [n_nodes-1][bsize]
    double2* dl = iP + (size_t)(n_nodes - 1) * bsize;                      // [nD]
    double2* acc = dl + e.nD;                                              // [bsize]
    const uint64_t* __restrict__ words = bp.words[it.entry];
    const uint32_t* __restrict__ toff = bp.tree_off[it.entry];
    const int n_trees = bp.n_trees[it.entry];
    const long long c = it.chunk0, nct = it.n_chunks_total;
    const int tree0 = (int)(c * n_trees / nct), tree1 = (int)((c + 1) * n_trees / nct);

    for (int k = 0; k < bsize; ++k) acc[k] = zero2();

    for (int sb = blockIdx.x; sb < n_sb; sb += gridDim.x) {
        const unsigned long long local = (unsigned long long)sb * blockDim.x + threadIdx.x;
        const bool active = local < count;
        if (!active) continue;
        // -- times of every backbone position (src/qmc_integrate.jl:225-235,425-449) ---------------
        double times[kDevMaxNodes + 1];
        bool ok = true;
        {
            double u = 1.0;
            const uint32_t kk = (uint32_t)(dy.start + local);
            for (int pos = n_nodes; pos >= 1; --pos) {
                const int src = e.pos_src[pos];
                double t;
                if (src == -1) t = t_i;
                else if (src == -2) t = t_w;
                else if (src == -3) t = t_f;
                else if (p.explicit_times) t = p.explicit_times[local * D + src];
                else {
                    const uint32_t xi = sobol_coord(dy.sobol + src * 32, __ldg(dy.sobol + D * 32 + src), kk);
                    const double x = (double)xi * 2.3283064365386963e-10;
                    const int den = (src < d_after) ? (d_after - src) : (D - src);
                    const double r = (den == 1) ? x : pow(x, 1.0 / (double)den);
                    u = (src == 0 || src == d_after) ? r : __dmul_rn(u, r);
                    t = (src < d_after) ? __dadd_rn(__dmul_rn(u, len_after), lo_after) : __dadd_rn(__dmul_rn(u, len_before), t_i);
                    ok = ok && (t >= 0.0);
                }
                times[pos] = t;
            }
        }
        if (!ok) continue;   // discarded sample still counts in N (src/qmc_integrate.jl:503)
        // -- i P_s(t_pos, t_pos-1) for every interval and sector (:357-374) ------------------------------
        for (int iv = 0; iv < n_nodes - 1; ++iv) {
            const double ta = times[iv + 1];
            double tb = times[iv + 2];
            if (tb < ta) tb = ta;
            for (int s = 0; s < S; ++s) {
                const int d = m.dim[s], bo = m.boff[s];
                for (int el = 0; el < d * d; ++el) {
                    double2 v;
                    if (e.mode == 0) {
                        const int r = el % d, cc = el / d;
                        v = (r == cc) ? make_double2(exp(-(tb - ta) * __ldg(p.E + m.eoff[s] + r)), 0.0) : zero2();
                    } else {
                        v = times_i(grid_interp(p.P + bo + el, bsize, p.n_tau, p.inv_h, tb, ta));
                    }
                    iP[(size_t)iv * bsize + bo + el] = v;
                }
            }
        }
        for (int q = 0; q < e.nD; ++q) {   // i Delta for every used (arc, table)  (:397-416)
            const int4 ds = __ldg(e.dslots + q);
            const double th = times[ds.y];
            double tt = times[ds.x];
            if (tt < th) tt = th;
            dl[q] = times_i(delta_eval(p.deltas[ds.z], tt, th));
        }
        // -- replay the trees --------------------------------------------------------------------------
        double2 V[kDevMaxNodes + 1][kMaxBlockDim * kMaxBlockDim];   // V[level]: (dim_cur x d_init), column-major
        int rem[kDevMaxNodes + 1];
        for (int t = tree0; t < tree1; ++t) {
            uint32_t pc = toff[t];
            const uint64_t root = words[pc++];
            const int s_init = (int)w_aux(root), d0 = m.dim[s_init];
            const int rootop = (int)((root >> 48) & 0xFFF) - 1;
            int depth = 1;
            if (rootop >= 0) {   // operator node at position 1: bare matrix (:377,540)
                const int s_next = (int)w_slotA(root), dr = m.dim[s_next];
                const double2* O = m.pool + m.op_off[(size_t)rootop * S + s_init];
                for (int k = 0; k < dr * d0; ++k) V[1][k] = O[k];
            } else {
                for (int k = 0; k < d0 * d0; ++k) V[1][k] = (k % d0 == k / d0) ? make_double2(1.0, 0.0) : zero2();
            }
            rem[1] = (int)w_nchild(root);
            while (depth >= 1) {
                if (rem[depth] == 0) { --depth; continue; }
                --rem[depth];
                const uint64_t w = words[pc++];
                const int s = (int)w_slotA(w), ds_ = m.dim[s];
                const int op = (int)((w >> 48) & 0xFFF) - 1;
                const int iv = depth - 1;   // the node sits at position depth+1: interval depth-1
                const double2* Pm = iP + (size_t)iv * bsize + m.boff[s];
                const double2* Vp = V[depth];
                double2 tmp[kMaxBlockDim * kMaxBlockDim];
                for (int j = 0; j < d0; ++j)           // tmp = iP_s * V_parent   (d_s x d0)
                    for (int i = 0; i < ds_; ++i) {
                        double2 a = zero2();
                        for (int k = 0; k < ds_; ++k) a = cfma(Pm[i + ds_ * k], Vp[k + ds_ * j], a);
                        tmp[i + ds_ * j] = a;
                    }
                int dr = ds_;
                double2* Vc = V[depth + 1];
                if (op >= 0) {                          // V_child = O * tmp   (d_t x d0)
                    const int tgt = m.op_target[(size_t)op * S + s];
                    dr = m.dim[tgt];
                    const double2* O = m.pool + m.op_off[(size_t)op * S + s];
                    for (int j = 0; j < d0; ++j)
                        for (int i = 0; i < dr; ++i) {
                            double2 a = zero2();
                            for (int k = 0; k < ds_; ++k) a = cfma(O[i + dr * k], tmp[k + ds_ * j], a);
                            Vc[i + dr * j] = a;
                        }
                } else {
                    for (int k = 0; k < ds_ * d0; ++k) Vc[k] = tmp[k];
                }
                const uint32_t sbq = w_slotB(w);
                if (sbq) {                              // interaction weight at the arc's tail (:506-507)
                    const double2 dv = dl[sbq - e.nP];
                    for (int k = 0; k < dr * d0; ++k) Vc[k] = cmul(Vc[k], dv);
                }
                const int nc = (int)w_nchild(w);
                if (nc == 0) {                          // leaf: top_result[s_init] += weight * product (:465)
                    const double2 coef = __ldg(e.coefs + w_aux(w));
                    double2* a = acc + m.boff[s_init];
                    for (int k = 0; k < d0 * d0; ++k) a[k] = cfma(coef, Vc[k], a[k]);
                } else {
                    ++depth;
                    rem[depth] = nc;
                }
            }
        }
        if (p.per_sample_out) {
            for (int k = 0; k < bsize; ++k) { p.per_sample_out[local * bsize + k] = acc[k]; acc[k] = zero2(); }
        }
    }
    if (p.per_sample_out) return;
    // -- CTA reduction ------------------------------------------------------------------------------------
    for (int k = 0; k < bsize; ++k) {
        double2 v = acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            v.x += __shfl_down_sync(0xFFFFFFFFu, v.x, off);
            v.y += __shfl_down_sync(0xFFFFFFFFu, v.y, off);
        }
        if (lane == 0) red[warp * bsize + k] = v;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < bsize; k += blockDim.x) {
        double2 v = zero2();
        for (int w2 = 0; w2 < nw; ++w2) v = cadd(v, red[w2 * bsize + k]);
        p.partials[((size_t)it.partial0 * gridDim.x + blockIdx.x) * bsize + k] = v;
    }
}

cudaError_t launch_block_step(const StepParams& p, const BlockParams& bp, dim3 grid, int threads, size_t smem, cudaStream_t st) {
    block_step_kernel<<<grid, threads, smem, st>>>(p, bp);
    return cudaGetLastError();
}

}  // namespace qiw

namespace qiw {

// ---- sector blocks larger than 1x1, real arithmetic: lane = sample, warp-uniform tree replay ---------
// The pruned configuration trees (src/topology_eval.jl:454-556 with the dead branches removed at
// compile time) are identical for every sample, so a warp takes 32 samples and replays the tree in
// lock step: no divergence, the program words are warp-uniform, the data are per lane.
//   CTA = (entry, 32 samples, group of tree chunks); its W warps walk different chunks of the entry's
//   trees over the SAME samples and share the per-sample tables in shared memory:
//     TP[(interval, element of the packed block vector)][lane] = Re(i P_s(t_pos, t_pos-1))   (:357-374)
//     TD[pair-interaction slot][lane]                          = Re(i Delta(t_tail, t_head))  (:397-416)
//   The running product A_pos ... A_1 (src/utility.jl:234-323) of the current node lives in REGISTERS
//   (dims <= 4 x 4, fully unrolled matrix-vector products specialised on the block shape); it is saved to
//   a small shared-memory stack only at nodes with more than one child — the reference's prefix
//   sharing with a stack as deep as the number of branch points (<= order + 1), not the tree depth.
//   The scalar pair-interaction weights (:506-507) are carried as one running product per lane and
//   applied at the leaf together with the coefficient (topology sign, :431).
// Preconditions checked by the host: operator blocks real, P and Delta purely imaginary, coefficients
// purely imaginary (then every product is real, exactly).  Otherwise block_step_kernel (complex) runs.
// One edge = two stages, each dispatched on its own block shape so that the code the walker touches stays
// small (the SM's instruction cache holds 32 KB): V <- iP_s * V (DS x DS times DS x D0), then, at nodes
// with an operator, V <- O * V (DR x DS times DS x D0).  Both run column by column and in place.
template <int DS, int D0>
__device__ __forceinline__ void block_mul_P(const double* __restrict__ Pm, double (&V)[4 * D0]) {
    double Pv[DS * DS];
#pragma unroll
    for (int k = 0; k < DS * DS; ++k) Pv[k] = Pm[k * 32];
#pragma unroll
    for (int j = 0; j < D0; ++j) {
        double t[DS];
#pragma unroll
        for (int i = 0; i < DS; ++i) {
            double a = Pv[i] * V[4 * j];
#pragma unroll
            for (int k = 1; k < DS; ++k) a = fma(Pv[i + DS * k], V[k + 4 * j], a);
            t[i] = a;
        }
#pragma unroll
        for (int i = 0; i < DS; ++i) V[i + 4 * j] = t[i];
    }
}

template <int DR, int DS, int D0>
__device__ __forceinline__ void block_mul_O(const double* __restrict__ O, double (&V)[4 * D0]) {
    double Ov[DR * DS];
#pragma unroll
    for (int k = 0; k < DR * DS; ++k) Ov[k] = __ldg(O + k);
#pragma unroll
    for (int j = 0; j < D0; ++j) {
        double t[DS];
#pragma unroll
        for (int k = 0; k < DS; ++k) t[k] = V[k + 4 * j];
#pragma unroll
        for (int i = 0; i < DR; ++i) {
            double a = Ov[i] * t[0];
#pragma unroll
            for (int k = 1; k < DS; ++k) a = fma(Ov[i + DR * k], t[k], a);
            V[i + 4 * j] = a;
        }
    }
}

template <int D0>
__device__ __forceinline__ void block_edge_dispatch(int dr, int ds, const double* Pm, const double* O, bool has_op, double (&V)[4 * D0]) {
    switch (ds) {
        case 1: block_mul_P<1, D0>(Pm, V); break;
        case 2: block_mul_P<2, D0>(Pm, V); break;
        case 3: block_mul_P<3, D0>(Pm, V); break;
        default: block_mul_P<4, D0>(Pm, V); break;
    }
    if (!has_op) return;
#define QIW_BE(R_, S_) case (R_ * 8 + S_): block_mul_O<R_, S_, D0>(O, V); break;
    switch (dr * 8 + ds) {
        QIW_BE(1, 1) QIW_BE(1, 2) QIW_BE(1, 3) QIW_BE(1, 4) QIW_BE(2, 1) QIW_BE(2, 2) QIW_BE(2, 3) QIW_BE(2, 4)
        QIW_BE(3, 1) QIW_BE(3, 2) QIW_BE(3, 3) QIW_BE(3, 4) QIW_BE(4, 1) QIW_BE(4, 2) QIW_BE(4, 3) QIW_BE(4, 4)
        default: break;
    }
#undef QIW_BE
}

// Expanded program word of the block walker (built on the host from the tree words and the model, so the
// device loop needs no model look-ups and can prefetch one edge ahead):
//   x = ds | dr << 4 | has_op << 8 | nchild << 16      (source / target block dimension of the edge)
//   y = offset of the source sector's block of the edge's interval in the table TP, in elements
//       (interval * bsize + block offset; folded on the host) | (Delta slot + 1) << 16
//   z = offset of the operator block in pool_re
//   w = leaf: coefficient index; root: element offset of the initial sector's block
//   root only: x also carries the column group of this unit: number of columns << 9 | first column << 11
// The columns of the running product are independent (V <- A V acts column by column), so a tree whose
// initial sector has dimension d0 > 2 is replayed once per group of at most two columns: the walker then
// holds at most 4 x 2 doubles of V per lane whatever the sector sizes (template parameter D0 below = columns
// of the group), which keeps it at 85 registers, 24 warps per SM, and halves the code it touches.
// Replays one unit for the warp's 32 samples and adds the sum over the samples of
// Im(coef) * dprod * chain (columns c0 .. c0 + D0 - 1 of the d0 x d0 result) to wacc (the warp's block sums of
// the tree's initial sector).
template <int D0>
__device__ __forceinline__ void block_walk_tree(const uint4* __restrict__ xw, uint32_t pc, const double* __restrict__ pool_re,
                                                const double2* __restrict__ coefs, const double* TP, const double* TD,
                                                int bsize, int lane, double* wacc) {
    // branch-point stack: per-thread local memory (lane-interleaved by the hardware, L1-resident), so the
    // CTA's shared memory holds only the per-sample tables and several CTAs fit on an SM
    double stackV[kWalkMaxSp * 4 * D0];
    double stackD[kWalkMaxSp];
    int stackRem[kWalkMaxSp];
    double acc[4 * D0];
#pragma unroll
    for (int k = 0; k < 4 * D0; ++k) acc[k] = 0.0;
    const uint4 root = __ldg(xw + pc);
    ++pc;
    const int d0 = (int)(root.x & 0xFu), c0 = (int)((root.x >> 11) & 0x3u);
    double V[4 * D0];
#pragma unroll
    for (int k = 0; k < 4 * D0; ++k) V[k] = 0.0;
    if ((root.x >> 8) & 1u) {   // operator node at position 1: bare matrix (:377,540)
        const int dcur = (int)((root.x >> 4) & 0xFu);
        const double* O = pool_re + root.z;
#pragma unroll
        for (int j = 0; j < D0; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < dcur) V[i + 4 * j] = __ldg(O + i + dcur * (c0 + j));
    } else {
#pragma unroll
        for (int j = 0; j < D0; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) V[i + 4 * j] = (i == c0 + j) ? 1.0 : 0.0;
    }
    double dprod = 1.0;
    int sp = 0;
    int nch = (int)(root.x >> 16);
    uint4 w = __ldg(xw + pc);   // the next edge is always the next word of the pre-order stream
    while (nch > 0) {
        if (nch > 1) {   // branch point: save the running product for the later siblings
#pragma unroll
            for (int k = 0; k < 4 * D0; ++k) stackV[sp * 4 * D0 + k] = V[k];
            stackD[sp] = dprod;
            stackRem[sp] = nch - 1;
            ++sp;
        }
        // one edge
        const uint4 cur = w;
        ++pc;
        w = __ldg(xw + pc);     // one word ahead (the stream is padded by one word) ...
        // ... and its cache line well ahead: the words are read once per block of samples, so every new
        // 128-byte line (8 words) would otherwise come from L2 with the warp waiting on it
        asm volatile("prefetch.global.L1 [%0];" ::"l"(xw + pc + kWalkPrefetch));
        const int ds = (int)(cur.x & 0xFu), dr = (int)((cur.x >> 4) & 0xFu);
        const bool has_op = (cur.x >> 8) & 1u;
        const double* Pm = TP + (cur.y & 0xFFFFu) * 32 + lane;
        block_edge_dispatch<D0>(dr, ds, Pm, pool_re + cur.z, has_op, V);
        const uint32_t sbq = cur.y >> 16;
        if (sbq) dprod *= TD[(size_t)(sbq - 1) * 32 + lane];   // interaction weight at the arc's tail (:506-507)
        nch = (int)(cur.x >> 16);
        if (nch == 0) {   // leaf: top_result[s_init] += weight * product (:465); rows >= d0 of V are stale
            const double c = __ldg(&coefs[cur.w].y) * dprod;
#pragma unroll
            for (int k = 0; k < 4 * D0; ++k) acc[k] = fma(c, ((k & 3) < d0) ? V[k] : 0.0, acc[k]);
            if (sp == 0) break;
            // back to the nearest branch point with children left
            --sp;
            const int rem = stackRem[sp];
#pragma unroll
            for (int k = 0; k < 4 * D0; ++k) V[k] = stackV[sp * 4 * D0 + k];
            dprod = stackD[sp];
            if (rem > 1) {   // more siblings after this one: keep the frame
                stackRem[sp] = rem - 1;
                ++sp;
            }
            nch = 1;   // continue with exactly one child of the restored node (the frame handles the rest)
        }
    }
    // sum over the warp's samples, then one lane adds to the warp's block sums
#pragma unroll
    for (int j = 0; j < D0; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double v = acc[i + 4 * j];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, off);
            if (lane == 0 && i < d0) wacc[i + d0 * (c0 + j)] += v;
        }
}

__global__ void __launch_bounds__(768, 1) block_walk_kernel(const StepParams p, const BlockParams bp, const BlockWalkParams wp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, nthr = blockDim.x;
    const WorkItem it = p.items[blockIdx.y];
    const DevEntry& e = p.entries[it.entry];
    const DevEntryDyn& dy = p.dyn[it.slot];
    const DevModel& m = bp.m;
    const int S = m.S, bsize = m.bsize, D = e.D, n_nodes = e.n_nodes, d_after = e.d_after, nD = e.nD, nI = n_nodes - 1;
    // shared memory carve-up
    double* TP = reinterpret_cast<double*>(smem_raw);                       // [nI_max * bsize][32]
    double* TD = TP + (size_t)wp.nI_max * bsize * 32;                       // [nD_max][32]
    double* accs = TD + (size_t)wp.nD_max * 32;                             // [nw][bsize] per-warp block sums
    double* times = accs + (size_t)nw * bsize;                              // [kDevMaxNodes + 1][32]
    double* pw = times + (kDevMaxNodes + 1) * 32;                           // [kDevMaxDim][32]
    int* okflag = reinterpret_cast<int*>(pw + kDevMaxDim * 32);             // [32]

    const double t_i = p.t_i, t_w = p.t_w, t_f = p.t_f;
    const double lo_after = (e.mode == 0) ? t_i : t_w, len_after = t_f - lo_after, len_before = t_w - t_i;
    const unsigned long long count = dy.count;
    const int n_sb = (int)((count + 31ull) >> 5);
    const uint4* __restrict__ xw = wp.xwords[it.entry];
    const uint32_t* __restrict__ toff = wp.unit_off[it.entry];
    const int* bounds = wp.chunk_bounds + (size_t)blockIdx.y * (wp.warps + 1);
    const int tree0 = bounds[warp], tree1 = bounds[warp + 1];   // contiguous range of about equal cost per warp
    double* my_acc = accs + (size_t)warp * bsize;
    for (int k = lane; k < bsize; k += 32) my_acc[k] = 0.0;
    __syncwarp();

    for (int sb = blockIdx.x; sb < n_sb; sb += gridDim.x) {
        const unsigned long long local0 = (unsigned long long)sb * 32ull;
        // -- 1. roots, 2. times (as in the scalar kernel) ---------------------------------------------------
        if (threadIdx.x < 32) okflag[threadIdx.x] = (local0 + threadIdx.x < count) ? 1 : 0;
        for (int task = threadIdx.x; task < D * 32; task += nthr) {
            const int j = task >> 5, smp = task & 31;
            const uint32_t xi = sobol_coord(dy.sobol + j * 32, __ldg(dy.sobol + D * 32 + j), (uint32_t)(dy.start + local0 + smp));
            const double x = (double)xi * 2.3283064365386963e-10;
            const int den = (j < d_after) ? (d_after - j) : (D - j);
            pw[j * 32 + smp] = (den == 1) ? x : pow(x, 1.0 / (double)den);
        }
        __syncthreads();
        for (int task = threadIdx.x; task < n_nodes * 32; task += nthr) {
            const int pos = 1 + (task >> 5), smp = task & 31;
            const int src = e.pos_src[pos];
            double t;
            if (src == -1) t = t_i;
            else if (src == -2) t = t_w;
            else if (src == -3) t = t_f;
            else {
                const int j0 = (src < d_after) ? 0 : d_after;
                double u = pw[j0 * 32 + smp];
                for (int j = j0 + 1; j <= src; ++j) u = __dmul_rn(u, pw[j * 32 + smp]);
                if (src < d_after) t = __dadd_rn(__dmul_rn(u, len_after), lo_after);
                else t = __dadd_rn(__dmul_rn(u, len_before), t_i);
                if (!(t >= 0.0)) okflag[smp] = 0;
            }
            times[pos * 32 + smp] = t;
        }
        __syncthreads();
        // -- 3. tables: Re(i P) for every interval and block element, Re(i Delta) for every slot -------------
        for (int task = threadIdx.x; task < (nI + nD) * 32; task += nthr) {
            const int q = task >> 5, smp = task & 31;
            const bool ok = okflag[smp] != 0;
            if (q < nI) {
                const double ta = times[(q + 1) * 32 + smp];
                double tb = times[(q + 2) * 32 + smp];
                if (tb < ta) tb = ta;
                double* dst = TP + (size_t)q * bsize * 32 + smp;
                if (e.mode == 0) {
                    for (int s = 0; s < S; ++s) {
                        const int d = m.dim[s], bo = m.boff[s];
                        for (int el = 0; el < d * d; ++el) {
                            const int r = el % d, cc = el / d;
                            const double v = (r == cc) ? exp(-(tb - ta) * __ldg(p.E + m.eoff[s] + r)) : 0.0;
                            dst[(size_t)(bo + el) * 32] = ok ? v : 0.0;
                        }
                    }
                } else {
                    const GridCell cell = grid_cell(p.n_tau, p.inv_h, tb, ta);
                    for (int el = 0; el < bsize; ++el) {
                        const double v = cell_apply_i<true>(p.P + el, bsize, cell);
                        dst[(size_t)el * 32] = ok ? v : 0.0;
                    }
                }
            } else {
                const int4 ds = __ldg(e.dslots + (q - nI));
                const double th = times[ds.y * 32 + smp];
                double tt = times[ds.x * 32 + smp];
                if (tt < th) tt = th;
                const double v = delta_apply_i<true>(p.deltas[ds.z], tt, th);
                TD[(size_t)(q - nI) * 32 + smp] = ok ? v : 0.0;
            }
        }
        __syncthreads();
        // -- 4. this warp's trees ----------------------------------------------------------------------------
        for (int t = tree0; t < tree1; ++t) {
            const uint32_t pc = toff[t];
            const uint4 root = __ldg(xw + pc);
            double* a = my_acc + root.w;
            if (((root.x >> 9) & 0x3u) == 1u) block_walk_tree<1>(xw, pc, wp.pool_re, e.coefs, TP, TD, bsize, lane, a);
            else block_walk_tree<2>(xw, pc, wp.pool_re, e.coefs, TP, TD, bsize, lane, a);
            __syncwarp();
        }
        __syncthreads();
    }
    __syncthreads();
    // -- 5. CTA result: sum over the warps, coefficient's factor i restored ------------------------------------
    for (int k = threadIdx.x; k < bsize; k += nthr) {
        double v = 0.0;
        for (int w2 = 0; w2 < nw; ++w2) v += accs[(size_t)w2 * bsize + k];
        p.partials[((size_t)it.partial0 * gridDim.x + blockIdx.x) * bsize + k] = make_double2(0.0, v);
    }
}

cudaError_t launch_block_walk(const StepParams& p, const BlockParams& bp, const BlockWalkParams& wp, dim3 grid, int threads,
                              size_t smem, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(block_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    block_walk_kernel<<<grid, threads, smem, st>>>(p, bp, wp);
    return cudaGetLastError();
}

}  // namespace qiw
