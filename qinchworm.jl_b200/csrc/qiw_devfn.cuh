// qiw_devfn.cuh — device helpers shared by the kernels of libqinchworm_cuda.so: complex arithmetic, the
// Keldysh.jl grid interpolation rule, the natural cubic spline, Sobol points, real / complex arithmetic modes.
#pragma once

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
    static __device__ __forceinline__ T shfl_xor(T v, int m) { return __shfl_xor_sync(0xFFFFFFFFu, v, m); }
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
    static __device__ __forceinline__ T shfl_xor(T v, int m) {
        return make_double2(__shfl_xor_sync(0xFFFFFFFFu, v.x, m), __shfl_xor_sync(0xFFFFFFFFu, v.y, m));
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

// ---- peer-memory all-reduce: one element on the wire ----------------------------------------------
// Low-latency protocol: a double travels as one 16-byte line {low half, seq, high half, seq}; the receiver needs no
// separate flag and the sender no system-wide fence: each 8-byte half is valid as soon as its flag shows the
// sequence number of the collective (the line is written and read with single 128-bit volatile accesses).
__device__ __forceinline__ void peer_send(const StepParams& p, size_t slot_byte_off, int j, double value, unsigned int seq32) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(value);
    const unsigned int lo = (unsigned int)bits, hi = (unsigned int)(bits >> 32);
    for (int q = 0; q < p.peer_ranks; ++q) {
        if (q == p.peer_rank) continue;
        uint4* dst = reinterpret_cast<uint4*>(p.peer_mail[q] + slot_byte_off) + j;
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(seq32), "r"(hi), "r"(seq32) : "memory");
    }
}

// Element j of the all-reduced vector: this rank's `own` value and the peers' values from this GPU's mailbox, added
// in rank order (every rank forms the same sum).  A peer that does not answer within the time-out raises the status
// flag and the result is NaN, so that a device-resident run cannot continue on a partial sum (the host reports the
// error).  Peer by peer on purpose: requesting all peers' lines at once (eight lines in flight per thread, or one
// thread per (peer, element) through shared memory) was measured on 4 GPUs and gains nothing — the wait is for the
// slowest peer, not for L2 — while the extra live registers cost the run kernel 4-15 % on ONE GPU.
__device__ __forceinline__ double peer_gather(const StepParams& p, int par, unsigned int seq32, int j, double own,
                                              unsigned long long t0) {
    const unsigned char* base = p.peer_mail[p.peer_rank] + kPeerFlagBytes;
    double v = 0.0;
    bool ok = true;
    for (int q = 0; q < p.peer_ranks; ++q) {
        if (q == p.peer_rank) { v += own; continue; }
        const uint4* src = reinterpret_cast<const uint4*>(base + ((size_t)q * 2 + par) * kPeerSlotBytes) + j;
        uint4 w;
        do {
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(src) : "memory");
            if ((w.y != seq32 || w.w != seq32) && globaltimer_ns() - t0 > p.peer_timeout_ns) { *p.peer_status = 1; ok = false; break; }
        } while (w.y != seq32 || w.w != seq32);
        v += __longlong_as_double((long long)(((unsigned long long)w.z << 32) | (unsigned long long)w.x));
    }
    return ok ? v : __longlong_as_double(0x7FF8000000000000ll);
}

// ---- fused tail of the step kernel ---------------------------------------------------------------
// Executed by the last CTA of a step launch: every (entry, sector) sum runs over the partial rows in
// fixed order, so the result does not depend on which CTA happens to be last.  Optionally followed by
// set_ppgf!(P, tau_f, result) and normalize!(P, tau_f) (src/ppgf.jl:495-504,646-668).
// Rows are `bsize` wide (the packed block vector; == S for models of 1x1 blocks).
__device__ inline void fused_tail(const StepParams& pp, double t_i, double t_w, double t_f, int pitch) {
    StepParams p = pp;
    const int S = p.bsize, n_out = p.n_call_entries * S;
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
        // protocol: peer_send / peer_gather above.  Buffers alternate with the parity of the sequence number: a slot
        // is rewritten two collectives later, after every peer has provably finished reading it.
        __syncthreads();
        const int par = (int)(p.peer_seq & 1ull);
        const unsigned int seq32 = (unsigned int)(p.peer_seq % 0xFFFFFFFFull) + 1u;   // never 0 (the mailbox starts zeroed)
        const size_t my_slot = kPeerFlagBytes + ((size_t)p.peer_rank * 2 + par) * kPeerSlotBytes;
        const int n_dbl = 2 * n_out;
        double* outd = reinterpret_cast<double*>(p.out);
        for (int j = threadIdx.x; j < n_dbl; j += blockDim.x) peer_send(p, my_slot, j, outd[j], seq32);
        const unsigned long long t0 = globaltimer_ns();
        for (int j = threadIdx.x; j < n_dbl; j += blockDim.x)
            outd[j] = peer_gather(p, par, seq32, j, outd[j], t0);   // element j is read and written by this thread only
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

}  // namespace qiw
