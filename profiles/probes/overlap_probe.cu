// overlap_probe.cu -- do FP64 FMAs and 16 B shared-memory traffic overlap on one SM?  (design probe for k_tile_stream)
// Warps with role 0 run DFMA chains, role 1 run conflict-free LDS.128/STS.128 sweeps over a private 16 KB region,
// role 2 alternates both like a phase of the tile kernel (32 LDS.128, 320 DFMA, 32 STS.128).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void __launch_bounds__(512, 1) k_mix(double* out, int iters, int n_f, int n_s, int n_m) {
    extern __shared__ double2 sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* mine = sm + warp * 1024;
    for (int i = lane; i < 1024; i += 32) mine[i] = make_double2(i, -i);
    __syncthreads();
    double acc = 0;
    if (warp < n_f) {
        double a[16], b[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { a[i] = lane + i; b[i] = 0.5 * i; }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 10; ++rep) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fma(0.3, b[i], a[i]);
#pragma unroll
                for (int i = 0; i < 16; ++i) b[i] = fma(-0.3, a[i], b[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += a[i] + b[i];
    } else if (warp < n_f + n_s) {
        double2 v[32];
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = mine[lane + 32 * k];
#pragma unroll
            for (int k = 0; k < 32; ++k) { v[k].x += 1.0; }
#pragma unroll
            for (int k = 0; k < 32; ++k) mine[lane + 32 * k] = v[k];
            __syncwarp();
        }
        acc = v[3].x;
    } else if (warp < n_f + n_s + n_m) {
        double2 v[32];
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = mine[lane + 32 * k];
#pragma unroll
            for (int lev = 0; lev < 5; ++lev)
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (!((i >> lev) & 1)) {
                        double2 x0 = v[i], x1 = v[i | (1 << lev)];
                        v[i].x = fma(0.3, x1.y, x0.x); v[i].y = fma(-0.3, x1.x, x0.y);
                        v[i | (1 << lev)].x = fma(0.3, x0.y, x1.x); v[i | (1 << lev)].y = fma(-0.3, x0.x, x1.y);
                    }
#pragma unroll
            for (int k = 0; k < 32; ++k) mine[lane + 32 * k] = v[k];
            __syncwarp();
        }
        acc = v[3].x;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    double* out;
    cudaMalloc(&out, 148 * 512 * 8);
    cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 16384);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    struct { int f, s, m; const char* what; } cases[] = {
        {4, 0, 0, "4 DFMA warps (1/SMSP)"}, {0, 4, 0, "4 LDS/STS warps"}, {4, 4, 0, "4 DFMA + 4 LDS/STS warps"},
        {8, 0, 0, "8 DFMA warps"}, {0, 8, 0, "8 LDS/STS warps"}, {4, 8, 0, "4 DFMA + 8 LDS/STS"}, {8, 4, 0, "8 DFMA + 4 LDS/STS"},
        {0, 0, 4, "4 phase-like warps"}, {0, 0, 8, "8 phase-like warps"}, {0, 0, 12, "12 phase-like warps"} };
    for (auto& c : cases) {
        const int threads = 32 * (c.f + c.s + c.m);
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k_mix<<<148, threads, 12 * 16384>>>(out, iters, c.f, c.s, c.m);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        // per iteration per warp: DFMA role 320 DFMA; LDS/STS role 32+32 x 512 B; phase role both
        printf("%-28s %8.3f ms  -> %7.1f ns per iteration\n", c.what, ms, ms * 1e6 / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
