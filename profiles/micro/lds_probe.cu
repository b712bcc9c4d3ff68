// lds_probe.cu — micro-benchmark behind DESIGN.md §"shared-memory operand bandwidth":
// cycles per warp-level LDS for the access patterns the configuration walk can generate, and the
// DFMA / DMUL issue rate next to it.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 lds_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int WIDTH, int PATTERN>
__global__ void __launch_bounds__(256) probe(double* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) reinterpret_cast<double*>(sm)[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // PATTERN 0: all lanes same address; 1: 4 distinct addresses (lane & 3) adjacent; 2: all distinct, consecutive;
    // 3: 4 distinct addresses in the same bank (stride 128 B); 4: 32 distinct, 2 per bank-group random-ish
    int idx;
    if (PATTERN == 0) idx = 0;
    else if (PATTERN == 1) idx = lane & 3;
    else if (PATTERN == 2) idx = lane;
    else if (PATTERN == 3) idx = (lane & 3) * (128 / WIDTH);
    else idx = (lane * 5) & 31;
    const unsigned char* base = sm + idx * WIDTH;
    double a0 = 1.0, a1 = 1.0, a2 = 1.0, a3 = 1.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const unsigned char* p = base + ((it * 16 + u) & 63) * 512;
            if (WIDTH == 4) { float v = *reinterpret_cast<const float*>(p); a0 += v; }
            else if (WIDTH == 8) { double v = *reinterpret_cast<const double*>(p); a0 = fma(a0, 1.0000001, v); }
            else { double2 v = *reinterpret_cast<const double2*>(p); a0 = fma(a0, 1.0000001, v.x); a1 = fma(a1, 1.0000001, v.y); }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// FP64 multiply chain fed from shared memory: `MULS` DMULs per LDS.64 (same-address broadcast), 8 warps
template <int MULS>
__global__ void __launch_bounds__(256) feed(double* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) reinterpret_cast<double*>(sm)[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned char* base = sm + (lane & 3) * 8;
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.0 + k;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double v = *reinterpret_cast<const double*>(base + ((it * 8 + u) & 63) * 512);
#pragma unroll
            for (int k = 0; k < MULS; ++k) a[k] *= v;
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class K>
static void run(const char* name, K kern, int per_iter, int blocks_per_sm) {
    double* out; long long* cyc;
    const int blocks = 148 * blocks_per_sm, iters = 2000;
    cudaMalloc(&out, blocks * 256 * sizeof(double));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    kern<<<blocks, 256, 65536, 0>>>(out, cyc, iters);
    kern<<<blocks, 256, 65536, 0>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148 * 4];
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int b = 0; b < blocks; ++b) avg += (double)h[b];
    avg /= blocks;
    // per SM: blocks_per_sm CTAs x 8 warps each issue iters*per_iter loads in `avg` cycles
    printf("%-44s %8.3f SM-cycles per warp-level load (8 warps x %d CTA/SM)  err=%s\n", name,
           avg / ((double)iters * per_iter * 8 * blocks_per_sm), blocks_per_sm, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run("LDS.32  same address", probe<4, 0>, 16, 1);
    run("LDS.32  32 consecutive", probe<4, 2>, 16, 1);
    run("LDS.64  same address", probe<8, 0>, 16, 1);
    run("LDS.64  4 adjacent addresses", probe<8, 1>, 16, 1);
    run("LDS.64  4 addresses same bank", probe<8, 3>, 16, 1);
    run("LDS.64  32 consecutive", probe<8, 2>, 16, 1);
    run("LDS.64  32 permuted", probe<8, 4>, 16, 1);
    run("LDS.128 same address", probe<16, 0>, 16, 1);
    run("LDS.128 4 adjacent addresses", probe<16, 1>, 16, 1);
    run("LDS.128 4 addresses same bank", probe<16, 3>, 16, 1);
    run("LDS.128 32 consecutive", probe<16, 2>, 16, 1);
    run("LDS.128 32 permuted", probe<16, 4>, 16, 1);
    run("LDS.64 bcast + 1 DMUL per load", feed<1>, 8, 1);
    run("LDS.64 bcast + 2 DMUL per load", feed<2>, 8, 1);
    run("LDS.64 bcast + 4 DMUL per load", feed<4>, 8, 1);
    run("LDS.64 bcast + 8 DMUL per load", feed<8>, 8, 1);
    run("LDS.64 bcast + 1 DMUL per load, 2 CTA/SM", feed<1>, 8, 2);
    run("LDS.64 bcast + 2 DMUL per load, 2 CTA/SM", feed<2>, 8, 2);
    run("LDS.64 bcast + 4 DMUL per load, 2 CTA/SM", feed<4>, 8, 2);
    return 0;
}
