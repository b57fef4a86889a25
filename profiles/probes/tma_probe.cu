// tma_probe.cu -- round-1 hardware probes behind the design of k_tile_stream (not part of the product):
//   (1) where does a 5-D tensor-map box with a 64 B inner row land in shared memory under SWIZZLE_128B?
//   (2) the same for a 2-D box with 128 B rows; round trip through a tensor-map store;
//   (3) FP64 FMA issue rate per SM at 1/2/3 warps per scheduler with 8/16 independent chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe5(const __grid_constant__ CUtensorMap in, const __grid_constant__ CUtensorMap out, double* dump,
                         int c1, int c4) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t dst = (smem_u32(sm) + 1023u) & ~1023u;
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(65536));
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5,%6}], [%7];"
            ::"r"(dst), "l"(&in), "r"(0), "r"(c1), "r"(0), "r"(0), "r"(c4), "r"(b) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b) : "memory");
    const double* s = (const double*)(sm + (dst - smem_u32(sm)));
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) dump[i] = s[i];
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%1,%2,%3,%4,%5}], [%6];"
                     ::"l"(&out), "r"(0), "r"(c1), "r"(0), "r"(0), "r"(c4), "r"(dst) : "memory");
        asm volatile("cp.async.bulk.commit_group;");
        asm volatile("cp.async.bulk.wait_group 0;");
    }
}

__global__ void k_probe2(const __grid_constant__ CUtensorMap in, double* dump, int row0) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t dst = (smem_u32(sm) + 1023u) & ~1023u;
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(32768));
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3}], [%4];"
            ::"r"(dst), "l"(&in), "r"(0), "r"(row0), "r"(b) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b) : "memory");
    const double* s = (const double*)(sm + (dst - smem_u32(sm)));
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) dump[i] = s[i];
}

template <int ILP>
__global__ void k_dfma(double* out, int iters, double t) {
    double a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x + i; b[i] = 0.5 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(t, b[i], a[i]);
#pragma unroll
        for (int i = 0; i < ILP; ++i) b[i] = fma(-t, a[i], b[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    CK(cudaSetDevice(0));
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const int n = 20, g = 10;
    const size_t ne = (size_t)2 << n;     // two "trajectories"
    std::vector<double> h(ne * 2);
    for (size_t i = 0; i < ne; ++i) { h[2 * i] = (double)i; h[2 * i + 1] = -(double)i; }
    double *d, *d2, *dump;
    CK(cudaMalloc(&d, ne * 16));
    CK(cudaMalloc(&d2, ne * 16));
    CK(cudaMemset(d2, 0, ne * 16));
    CK(cudaMalloc(&dump, 65536));
    CK(cudaMemcpy(d, h.data(), ne * 16, cudaMemcpyHostToDevice));
    for (int swz = 0; swz < 2; ++swz) {   // SWIZZLE_128B with a 64 B inner row faults (illegal memory access): measured, excluded
        const CUtensorMapSwizzle sw = swz == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : swz == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
        CUtensorMap tin, tout;
        cuuint64_t dims[5] = {8, 1ull << (g - 2), 32, 32, 2ull << (n - g - 10)};
        cuuint64_t strides[4] = {64, 16ull << g, 16ull << (g + 5), 16ull << (g + 10)};
        cuuint32_t box[5] = {8, 1, 32, 32, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r1 = encode(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = encode(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, d2, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("5d swizzle %d: encode rc %d %d\n", swz, (int)r1, (int)r2);
        if (r1 || r2) continue;
        CK(cudaFuncSetAttribute(k_probe5, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560));
        const int c1 = 5, c4 = 1;     // tile with bits[2,10) = 5, second trajectory
        k_probe5<<<1, 128, 66560>>>(tin, tout, dump, c1, c4);
        CK(cudaDeviceSynchronize());
        std::vector<double> o(8192);
        CK(cudaMemcpy(o.data(), dump, 65536, cudaMemcpyDeviceToHost));
        // expected dense layout: smem amp index l = p | a<<2 (p: bits 0,1; a: bits g..g+9) holds global amp
        // ((c4<<n) | a<<g | c1<<2 | p).  Report the permutation actually found for the first 64 slots and check a model.
        int bad_dense = 0, bad_x = 0;
        for (int slot = 0; slot < 4096; ++slot) {
            const long long gi = (long long)o[2 * slot];
            const long long a = (gi >> g) & 1023, p = gi & 3, rest = (gi >> 2) & 255;
            const int l = (int)(p | (a << 2));
            if (rest != c1 || (gi >> n) != c4) { bad_dense++; bad_x++; continue; }
            if (l != slot) bad_dense++;
            if ((l ^ ((l >> 3) & 7)) != slot) bad_x++;
        }
        printf("  mismatches vs dense: %d, vs 16B-chunk xor (l ^ ((l>>3)&7)): %d\n", bad_dense, bad_x);
        printf("  first slots -> logical l: ");
        for (int slot = 0; slot < 40; ++slot) {
            const long long gi = (long long)o[2 * slot];
            printf("%lld ", ((gi & 3) | (((gi >> g) & 1023) << 2)));
        }
        printf("\n");
        // round trip
        std::vector<double> back(ne * 2);
        CK(cudaMemcpy(back.data(), d2, ne * 16, cudaMemcpyDeviceToHost));
        long long wrong = 0, written = 0;
        for (size_t i = 0; i < ne; ++i) {
            const bool in_tile = ((i >> 2) & 255) == (size_t)c1 && (i >> n) == (size_t)c4;
            if (back[2 * i] != 0.0 || back[2 * i + 1] != 0.0) written++;
            if (in_tile && i != 0 && back[2 * i] != (double)i) wrong++;
            if (!in_tile && back[2 * i] != 0.0) wrong++;
        }
        printf("  store round trip: %lld amplitudes written, %lld wrong\n", written, wrong);
        CK(cudaMemset(d2, 0, ne * 16));
    }
    {   // 2-D, 128 B rows
        CUtensorMap tin;
        cuuint64_t dims[2] = {16, ne / 8};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {16, 256};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("2d swizzle128: encode rc %d\n", (int)r);
        if (!r) {
            CK(cudaFuncSetAttribute(k_probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 34816));
            k_probe2<<<1, 128, 34816>>>(tin, dump, 512 * 3);
            CK(cudaDeviceSynchronize());
            std::vector<double> o(4096);
            CK(cudaMemcpy(o.data(), dump, 32768, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int slot = 0; slot < 2048; ++slot) {
                const long long l = (long long)o[2 * slot] - 4096 * 3;
                if ((l ^ ((l >> 3) & 7)) != slot) bad++;
            }
            printf("  mismatches vs xor model: %d\n", bad);
        }
    }
    {   // FP64 issue rate
        double* out;
        CK(cudaMalloc(&out, 148 * 1024 * 8));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int iters = 20000;
        for (int ilp = 8; ilp <= 16; ilp += 8)
            for (int wps = 1; wps <= 4; ++wps) {
                const int threads = 128 * wps;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    if (ilp == 8) k_dfma<8><<<148, threads>>>(out, iters, 0.3);
                    else k_dfma<16><<<148, threads>>>(out, iters, 0.3);
                    cudaEventRecord(e1);
                    CK(cudaEventSynchronize(e1));
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                const double fma = 2.0 * ilp * iters * (double)threads * 148;
                printf("dfma ilp %d warps/sched %d: %.3f ms, %.2f T DFMA/s (%.1f per SM per ns)\n", ilp, wps, ms, fma / ms * 1e-9,
                       fma / ms * 1e-6 / 148);
            }
    }
    return 0;
}
