// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps issuing them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I enerf_b200/csrc tools/tmem_bench.cu -o tools/bin/tmem_bench
#include "tc_common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace enerf::tc;

template <int MODE>   // 0: ld x32 (+wait), 1: st x16 (+wait), 2: ld x32 then st x16 (epilogue pattern), 3: ld x16
__global__ void k_bench(int iters, unsigned long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t q = warp & 3, group = warp >> 2;
    const uint32_t t = tmem_base + ((q * 32u) << 16) + (group * 96u) % 448u;
    uint32_t acc[32], p[16], x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = threadIdx.x + i;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
            tmem_ld32(t, acc);
            tc_wait_ld();
            x += acc[0] ^ acc[31];
            tmem_ld32(t + 32, acc);
            tc_wait_ld();
            x += acc[1] ^ acc[30];
        }
        if (MODE == 3) {
            uint32_t a16[16];
            tmem_ld16(t, a16);
            tc_wait_ld();
            x += a16[0] ^ a16[15];
            tmem_ld16(t + 16, a16);
            tc_wait_ld();
            x += a16[1] ^ a16[14];
        }
        if (MODE == 1 || MODE == 2) {
            p[0] = x;
            tmem_st16(t + 64, p);
            tmem_st16(t + 80, p);
            tc_wait_st();
        }
    }
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (x == 0x12345678u) sink[0] = x;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
    unsigned long long* d_c;
    uint32_t* d_s;
    cudaMalloc(&d_c, 148 * 8);
    cudaMalloc(&d_s, 4);
    const int iters = 2000;
    const char* names[4] = {"ld 2x(32x32b.x32) = 64 fp32 cols/row", "st 2x(32x32b.x16) = 32 cols/row", "ld 64 cols + st 32 cols", "ld 2x(32x32b.x16) = 32 cols/row"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {4, 8, 12, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k_bench<0><<<148, warps * 32>>>(iters, d_c, d_s);
                if (mode == 1) k_bench<1><<<148, warps * 32>>>(iters, d_c, d_s);
                if (mode == 2) k_bench<2><<<148, warps * 32>>>(iters, d_c, d_s);
                if (mode == 3) k_bench<3><<<148, warps * 32>>>(iters, d_c, d_s);
                cudaDeviceSynchronize();
            }
            unsigned long long c;
            cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
            const double per_iter = (double)c / iters;
            const double ld_bytes = (mode == 0 || mode == 2) ? warps * 32.0 * 64 * 4 : (mode == 3 ? warps * 32.0 * 32 * 4 : 0);
            const double st_bytes = (mode == 1 || mode == 2) ? warps * 32.0 * 32 * 4 : 0;
            printf("%-40s warps=%2d  %.0f cycles/iter  ld %.1f B/cyc/SM  st %.1f B/cyc/SM  (%s)\n", names[mode], warps, per_iter, ld_bytes / per_iter, st_bytes / per_iter,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
