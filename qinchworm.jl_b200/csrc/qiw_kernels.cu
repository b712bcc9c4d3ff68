// qiw_kernels.cu — hand-written CUDA kernels (sm_100a) of the qMC diagram-evaluation hot path, part 2:
// the kernels for sector blocks larger than 1x1 and the small service kernels.  The step / run kernels of
// 1x1-block models live in qiw_scalar.cu.
//
// block_walk_kernel (sector blocks up to 4x4, real arithmetic): lane = sample, warp-uniform replay of the
// pruned configuration tree with the running matrix product in registers.  block_step_kernel is the general
// complex path for block models.
#include <cstdio>

#include "qiw_devfn.cuh"

namespace qiw {

__global__ void sobol_points_kernel(int D, const uint32_t* __restrict__ m, const uint32_t* __restrict__ x0,
                                    unsigned long long start, unsigned long long count, uint32_t* __restrict__ out) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= count * (unsigned long long)D) return;
    const unsigned long long k = i / D;
    const int d = (int)(i % D);
    out[i] = sobol_coord(m + d * 32, x0[d], (uint32_t)(start + k));
}

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

// normalize!(P, lambda) alone (src/ppgf.jl:662-668): every stored grid value times exp(-lambda tau_k).  Used across the
// step seam (qiw_scale_P) so that a host-stepped loop sends one row and lambda instead of the whole table.
__global__ void __launch_bounds__(256) scale_P_kernel(double2* __restrict__ P, int n_tau, int bsize, double h, double lambda) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_tau * bsize) return;
    const int k = idx / bsize;
    const double f = exp(-((double)k * h) * lambda);
    P[idx] = cscale(f, P[idx]);
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

cudaError_t launch_scale_P(double2* P, int n_tau, int bsize, double h, double lambda, cudaStream_t st) {
    scale_P_kernel<<<(n_tau * bsize + 255) / 256, 256, 0, st>>>(P, n_tau, bsize, h, lambda);
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
// MAXD x MAXD (instantiated for 4 and 8).  This path favours generality over speed: complex hybridisation functions,
// per-sample evaluation (qiw_eval_at_times) and sector blocks of 5 to 8 rows, for which the real-arithmetic walker
// (block_walk_kernel, blocks <= 4x4) has no shapes; all BASELINE configurations run on the scalar kernel or the walker.

__device__ __forceinline__ double2 zero2() { return make_double2(0.0, 0.0); }
// tree word fields (qiw_host.hpp make_word)
__device__ __forceinline__ uint32_t w_slotA(uint64_t w) { return (uint32_t)w & 0xFFFu; }
__device__ __forceinline__ uint32_t w_slotB(uint64_t w) { return ((uint32_t)w >> 12) & 0xFFFu; }
__device__ __forceinline__ uint32_t w_nchild(uint64_t w) { return ((uint32_t)w >> 24) & 0xFFu; }
__device__ __forceinline__ uint32_t w_aux(uint64_t w) { return (uint32_t)(w >> 32) & 0xFFFFu; }

template <int MAXD>
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
        double2 V[kDevMaxNodes + 1][MAXD * MAXD];   // V[level]: (dim_cur x d_init), column-major
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
                double2 tmp[MAXD * MAXD];
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
    if (bp.m.maxdim <= 4) block_step_kernel<4><<<grid, threads, smem, st>>>(p, bp);
    else block_step_kernel<8><<<grid, threads, smem, st>>>(p, bp);
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
    for (int k = 0; k < DR * DS; ++k) Ov[k] = O[k];      // the operator pool is staged in shared memory when it fits (warp-uniform address: broadcast)
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
                if (i < dcur) V[i + 4 * j] = O[i + dcur * (c0 + j)];
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
    // the operator blocks (real parts) are read at warp-uniform addresses on every edge: a copy in shared memory when
    // the pool is small (round 1 read them through L1, whose lines the branch stacks and word streams keep evicting:
    // 1.0e9 global-load instructions per launch, L1 hit rate 64 %)
    const double* pool = wp.pool_re;
    if (wp.pool_n > 0) {
        double* pool_s = reinterpret_cast<double*>(okflag + 32);
        for (int k = threadIdx.x; k < wp.pool_n; k += nthr) pool_s[k] = __ldg(wp.pool_re + k);
        pool = pool_s;
        __syncthreads();
    }

    double t_i = p.t_i, t_w = p.t_w, t_f = p.t_f;
    // batched evaluation: blockIdx.z selects one (t_i, t_w, t_f) triple of the call, or one scrambled Sobol sequence
    if (p.times_dev) { const double* tz = p.times_dev + 3 * blockIdx.z; t_i = tz[0]; t_w = tz[1]; t_f = tz[2]; }
    const uint32_t* __restrict__ sm = dy.sobol + (size_t)blockIdx.z * p.sobol_z_stride;
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
            const uint32_t xi = sobol_coord(sm + j * 32, __ldg(sm + D * 32 + j), (uint32_t)(dy.start + local0 + smp));
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
            if (((root.x >> 9) & 0x3u) == 1u) block_walk_tree<1>(xw, pc, pool, e.coefs, TP, TD, bsize, lane, a);
            else block_walk_tree<2>(xw, pc, pool, e.coefs, TP, TD, bsize, lane, a);
            __syncwarp();
        }
        __syncthreads();
    }
    __syncthreads();
    // -- 5. CTA result: sum over the warps, coefficient's factor i restored ------------------------------------
    for (int k = threadIdx.x; k < bsize; k += nthr) {
        double v = 0.0;
        for (int w2 = 0; w2 < nw; ++w2) v += accs[(size_t)w2 * bsize + k];
        p.partials[((size_t)blockIdx.z * gridDim.y * gridDim.x + (size_t)it.partial0 * gridDim.x + blockIdx.x) * bsize + k] = make_double2(0.0, v);
    }
    // -- 6. fused tail, as in the scalar step kernel: the last CTA to arrive reduces all partial rows in fixed order,
    //       exchanges the block sums with the peer GPUs and (device-resident loop) performs set_ppgf! + normalize! -----
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
    }
}

// per-device opt-in to more than 48 KB of dynamic shared memory (one bit per device)
static cudaError_t optin_smem_block(const void* kernel, unsigned long long* mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && ((__atomic_load_n(mask, __ATOMIC_RELAXED) >> dev) & 1ull)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    if (dev < 64) __atomic_fetch_or(mask, 1ull << dev, __ATOMIC_RELAXED);
    return cudaSuccess;
}

cudaError_t launch_block_walk(const StepParams& p, const BlockParams& bp, const BlockWalkParams& wp, dim3 grid, int threads,
                              size_t smem, cudaStream_t st) {
    static unsigned long long mask = 0ull;
    cudaError_t e = optin_smem_block((const void*)block_walk_kernel, &mask);
    if (e != cudaSuccess) return e;
    block_walk_kernel<<<grid, threads, smem, st>>>(p, bp, wp);
    return cudaGetLastError();
}

}  // namespace qiw
