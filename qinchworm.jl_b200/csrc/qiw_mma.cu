// qiw_mma.cu — sector blocks of 5 to 8 rows on the FP64 tensor cores (sm_100a): block_mma_kernel.
//
// BASELINE.json's north star reserves `mma.sync.m8n8k4.f64` (DMMA) for "sector blocks large enough to be a real
// dense contraction".  Up to 4x4 the walker of qiw_kernels.cu (lane = sample, blocks in registers) is the right
// mapping; from 5 rows on an edge of the configuration tree — V <- O_op (iP_s V), src/topology_eval.jl:376-388,
// 506-537 with the reference's LazyMatrixProduct prefix sharing, src/utility.jl:234-323 — is an (<=8)x(<=8) by
// (<=8)x(<=8) product, i.e. two m8n8k4 tiles per factor when padded to 8, and the lane = sample layout would need
// 64 shared-memory loads of the iP block per lane and edge.  Here WARP = SAMPLE: the running product V lives in the
// accumulator fragment of the warp (2 doubles per lane), every block of iP or of an operator is loaded once per
// product as the A fragment (1 double per lane and k-chunk), and V is re-distributed from the accumulator layout to
// the B layout with two shuffles per k-chunk.  Real arithmetic (operator blocks real, P and Delta purely imaginary:
// checked by the host, as for the walker); anything else takes the general complex kernel.
//
// CTA = (entry, chunk of its trees, W consecutive samples — one per warp[, time triple / Sobol sequence z]).
#include <cstdio>

#include "qiw_devfn.cuh"

namespace qiw {

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// V <- A V for an (rows x cols) block A, column-major with leading dimension `rows`, zero-padded to 8x8; V in the
// accumulator layout: lane l holds V[l / 4][2 (l % 4) + {0, 1}].
__device__ __forceinline__ void mma_apply(const double* __restrict__ A, int rows, int cols, double& v0, double& v1, int lane) {
    const int m = lane >> 2, kq = lane & 3, n = lane >> 2;
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int k = kq + 4 * c;
        // B fragment: V[k][n], held in the accumulator layout by lane 4 k + n / 2, register n % 2
        const int src = 4 * k + (n >> 1);
        const double b0 = __shfl_sync(0xFFFFFFFFu, v0, src), b1 = __shfl_sync(0xFFFFFFFFu, v1, src);
        const double b = (n & 1) ? b1 : b0;
        const double a = (m < rows && k < cols) ? A[m + rows * k] : 0.0;
        dmma884(d0, d1, a, b, d0, d1);
    }
    v0 = d0; v1 = d1;
}

__global__ void __launch_bounds__(256) block_mma_kernel(const StepParams p, const BlockParams bp, const BlockWalkParams wp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const WorkItem it = p.items[blockIdx.y];
    const DevEntry& e = p.entries[it.entry];
    const DevEntryDyn& dy = p.dyn[it.slot];
    const DevModel& m = bp.m;
    const int S = m.S, bsize = m.bsize, D = e.D, n_nodes = e.n_nodes, d_after = e.d_after, nD = e.nD, nI = n_nodes - 1;
    // shared memory: per warp (= sample) the tables Re(i P) [nI_max][bsize], Re(i Delta) [nD_max], the times, and the
    // warp's block sums [bsize]
    const size_t per_warp = (size_t)wp.nI_max * bsize + wp.nD_max + (kDevMaxNodes + 1) + bsize;
    double* TP = reinterpret_cast<double*>(smem_raw) + (size_t)warp * per_warp;
    double* TD = TP + (size_t)wp.nI_max * bsize;
    double* times = TD + wp.nD_max;
    double* my_acc = times + (kDevMaxNodes + 1);

    double t_i = p.t_i, t_w = p.t_w, t_f = p.t_f;
    if (p.times_dev) { const double* tz = p.times_dev + 3 * blockIdx.z; t_i = tz[0]; t_w = tz[1]; t_f = tz[2]; }
    const uint32_t* __restrict__ sm = dy.sobol + (size_t)blockIdx.z * p.sobol_z_stride;
    const double lo_after = (e.mode == 0) ? t_i : t_w, len_after = t_f - lo_after, len_before = t_w - t_i;
    const unsigned long long count = dy.count;
    const int n_sb = (int)((count + (unsigned long long)nw - 1ull) / (unsigned long long)nw);
    const uint64_t* __restrict__ words = bp.words[it.entry];
    const uint32_t* __restrict__ toff = bp.tree_off[it.entry];
    const int n_trees = bp.n_trees[it.entry];
    const long long ch = it.chunk0, nct = it.n_chunks_total;
    const int tree0 = (int)(ch * n_trees / nct), tree1 = (int)((ch + 1) * n_trees / nct);
    for (int k = lane; k < bsize; k += 32) my_acc[k] = 0.0;
    __syncwarp();

    for (int sb = blockIdx.x; sb < n_sb; sb += gridDim.x) {
        const unsigned long long local = (unsigned long long)sb * nw + warp;
        if (local >= count) continue;          // warp-uniform
        // -- times of every backbone position: the lanes compute the roots, lane 0 chains them (src/qmc_integrate.jl:225-235,425-449)
        double root = 1.0;
        if (lane < D) {
            const uint32_t xi = sobol_coord(sm + lane * 32, __ldg(sm + D * 32 + lane), (uint32_t)(dy.start + local));
            const double x = (double)xi * 2.3283064365386963e-10;
            const int den = (lane < d_after) ? (d_after - lane) : (D - lane);
            root = (den == 1) ? x : pow(x, 1.0 / (double)den);
        }
        int ok = 1;
        {
            double u = 1.0;
            for (int pos = n_nodes; pos >= 1; --pos) {
                const int src = e.pos_src[pos];
                double t;
                if (src == -1) t = t_i;
                else if (src == -2) t = t_w;
                else if (src == -3) t = t_f;
                else {
                    const double r = __shfl_sync(0xFFFFFFFFu, root, src);
                    u = (src == 0 || src == d_after) ? r : __dmul_rn(u, r);
                    t = (src < d_after) ? __dadd_rn(__dmul_rn(u, len_after), lo_after) : __dadd_rn(__dmul_rn(u, len_before), t_i);
                    ok = ok && (t >= 0.0);
                }
                if (lane == 0) times[pos] = t;
            }
        }
        __syncwarp();
        if (!ok) continue;   // discarded sample still counts in N (src/qmc_integrate.jl:503)
        // -- tables of this sample: Re(i P_s(t_pos, t_pos-1)) for every interval and block element, Re(i Delta) per slot
        for (int task = lane; task < nI * bsize; task += 32) {
            const int iv = task / bsize, el = task - iv * bsize;
            const double ta = times[iv + 1];
            double tb = times[iv + 2];
            if (tb < ta) tb = ta;
            double v;
            if (e.mode == 0) {
                // bare: diagonal exp(-dt (E + lambda)); find the element's sector and (row, col)
                int s = 0;
                while (s + 1 < S && m.boff[s + 1] <= el) ++s;
                const int d = m.dim[s], r = (el - m.boff[s]) % d, cc = (el - m.boff[s]) / d;
                v = (r == cc) ? exp(-(tb - ta) * __ldg(p.E + m.eoff[s] + r)) : 0.0;
            } else {
                v = cell_apply_i<true>(p.P + el, bsize, grid_cell(p.n_tau, p.inv_h, tb, ta));
            }
            TP[(size_t)iv * bsize + el] = v;
        }
        for (int q = lane; q < nD; q += 32) {
            const int4 ds = __ldg(e.dslots + q);
            const double th = times[ds.y];
            double tt = times[ds.x];
            if (tt < th) tt = th;
            TD[q] = delta_apply_i<true>(p.deltas[ds.z], tt, th);
        }
        __syncwarp();
        // -- replay this CTA's trees: V[level] in the accumulator layout, one (8 x 8 padded) matrix per tree level
        double V0[kDevMaxNodes + 1], V1[kDevMaxNodes + 1];
        int rem[kDevMaxNodes + 1];
        const int r_ = lane >> 2, c_ = 2 * (lane & 3);
        for (int t = tree0; t < tree1; ++t) {
            uint32_t pc = toff[t];
            const uint64_t rootw = words[pc++];
            const int s_init = (int)((uint32_t)(rootw >> 32) & 0xFFFFu), d0 = m.dim[s_init];
            const int rootop = (int)((rootw >> 48) & 0xFFF) - 1;
            int depth = 1;
            if (rootop >= 0) {   // operator node at position 1: bare matrix (:377,540)
                const int s_next = (int)((uint32_t)rootw & 0xFFFu), dr = m.dim[s_next];
                const double* O = wp.pool_re + m.op_off[(size_t)rootop * S + s_init];
                V0[1] = (r_ < dr && c_ < d0) ? __ldg(O + r_ + dr * c_) : 0.0;
                V1[1] = (r_ < dr && c_ + 1 < d0) ? __ldg(O + r_ + dr * (c_ + 1)) : 0.0;
            } else {
                V0[1] = (r_ == c_ && r_ < d0) ? 1.0 : 0.0;
                V1[1] = (r_ == c_ + 1 && r_ < d0) ? 1.0 : 0.0;
            }
            rem[1] = (int)(((uint32_t)rootw >> 24) & 0xFFu);
            double a0 = 0.0, a1 = 0.0;      // this tree's contribution to block s_init, accumulator layout
            while (depth >= 1) {
                if (rem[depth] == 0) { --depth; continue; }
                --rem[depth];
                const uint64_t w = words[pc++];
                const int s = (int)((uint32_t)w & 0xFFFu), ds_ = m.dim[s];
                const int op = (int)((w >> 48) & 0xFFF) - 1;
                const int iv = depth - 1;   // the node sits at position depth + 1: interval depth - 1
                double v0 = V0[depth], v1 = V1[depth];
                mma_apply(TP + (size_t)iv * bsize + m.boff[s], ds_, ds_, v0, v1, lane);       // i P_s V
                if (op >= 0) {
                    const int tgt = m.op_target[(size_t)op * S + s];
                    mma_apply(wp.pool_re + m.op_off[(size_t)op * S + s], m.dim[tgt], ds_, v0, v1, lane);   // O (i P_s V)
                }
                const uint32_t sbq = ((uint32_t)w >> 12) & 0xFFFu;
                if (sbq) { const double dv = TD[sbq - e.nP]; v0 *= dv; v1 *= dv; }   // interaction weight at the arc's tail (:506-507)
                const int nc = (int)(((uint32_t)w >> 24) & 0xFFu);
                if (nc == 0) {                          // leaf: top_result[s_init] += weight * product (:465); coef = i * c.y
                    const double cf = __ldg(e.coefs + ((uint32_t)(w >> 32) & 0xFFFFu)).y;
                    a0 = fma(cf, v0, a0); a1 = fma(cf, v1, a1);
                } else {
                    ++depth;
                    V0[depth] = v0; V1[depth] = v1;
                    rem[depth] = nc;
                }
            }
            double* a = my_acc + m.boff[s_init];
            if (r_ < d0 && c_ < d0) a[r_ + d0 * c_] += a0;
            if (r_ < d0 && c_ + 1 < d0) a[r_ + d0 * (c_ + 1)] += a1;
            __syncwarp();
        }
    }
    __syncthreads();
    // -- CTA result: warps (= samples) summed in fixed order, the coefficient's factor i restored ----------------
    for (int k = threadIdx.x; k < bsize; k += blockDim.x) {
        double v = 0.0;
        for (int w2 = 0; w2 < nw; ++w2) v += reinterpret_cast<double*>(smem_raw)[(size_t)w2 * per_warp + (per_warp - bsize) + k];
        p.partials[((size_t)blockIdx.z * gridDim.y * gridDim.x + (size_t)it.partial0 * gridDim.x + blockIdx.x) * bsize + k] = make_double2(0.0, v);
    }
    // -- fused tail (as in the other step kernels) ----------------------------------------------------------------
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

cudaError_t launch_block_mma(const StepParams& p, const BlockParams& bp, const BlockWalkParams& wp, dim3 grid, int threads,
                             size_t smem, cudaStream_t st) {
    static unsigned long long mask = 0ull;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !((__atomic_load_n(&mask, __ATOMIC_RELAXED) >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(block_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return e;
        if (dev < 64) __atomic_fetch_or(&mask, 1ull << dev, __ATOMIC_RELAXED);
    }
    block_mma_kernel<<<grid, threads, smem, st>>>(p, bp, wp);
    return cudaGetLastError();
}

}  // namespace qiw
