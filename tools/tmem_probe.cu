// TMEM read-rate probe: how many cycles a 128-row x 64-column accumulator tile costs to read with tcgen05.ld, as fp32
// (2 x .32x32b.x32) and as fp16 accumulators packed two per register (.32x32b.x32.pack::16b), with 4..16 warps per SM reading at once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/tmem_probe tools/tmem_probe.cu && tools/_build/tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define R8(r, o) "=r"(r[o]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
#define REGS32 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
#define REGS16 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"

__device__ __forceinline__ void ld32(uint32_t a, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " REGS32 : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld32p(uint32_t a, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " REGS32 : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld16p(uint32_t a, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 " REGS16 : R8(r, 0), R8(r, 8) : "r"(a) : "memory");
}
__device__ __forceinline__ void st16(uint32_t a, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
                 "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE 0: 64 fp32 columns (2 loads), 1: 64 packed fp16 columns (1 load, 32 regs), 2: 32 fp32 columns (1 load), 3: 64 packed as 2 x x16
// 4: MODE 0 + 16-register tcgen05.st + wait::st per iteration (what a hidden-layer epilogue does)
template <int MODE>
__global__ void k_probe(unsigned long long* cycles, uint32_t* sink, int iters) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0 = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 96 % 416);
    // give the cells defined contents
    {
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0x3c003c00u + i;
        for (int c = 0; c < 96; c += 16) st16(t0 + c, z);
        wait_st();
    }
    __syncthreads();
    uint32_t acc = 0;
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 4) {
            uint32_t a[32], b[32];
            ld32(t0, a);
            ld32(t0 + 32, b);
            wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= a[i] + b[i];
            if (MODE == 4) {
                uint32_t p[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) p[i] = a[i] ^ b[i + 16];
                st16(t0 + 64, p);
                wait_st();
            }
        } else if (MODE == 1) {
            uint32_t a[32];
            ld32p(t0, a);
            wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= a[i];
        } else if (MODE == 2) {
            uint32_t a[32];
            ld32(t0, a);
            wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= a[i];
        } else {
            uint32_t a[16], b[16];
            ld16p(t0, a);
            ld16p(t0 + 32, b);
            wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) acc ^= a[i] + b[i];
        }
    }
    const long long c1 = clock64();
    if (lane == 0) cycles[blockIdx.x * 32 + warp] = (unsigned long long)(c1 - c0);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

// what does pack::16b return?  cells hold 0x3c003c00 + i in column i (16 columns repeated): print lane 0's registers
__global__ void k_layout(uint32_t* out) {
    __shared__ uint32_t tbase;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = ((0xA000u + i) << 16) | (0x1000u + i + (threadIdx.x << 8));
    st16(tbase, z);
    st16(tbase + 16, z);
    wait_st();
    uint32_t a[16];
    ld16p(tbase, a);
    wait_ld();
    if (threadIdx.x == 1)
        for (int i = 0; i < 16; ++i) out[i] = a[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(64u) : "memory");
}

template <int MODE>
static void run(const char* name, int warps, unsigned long long* d_cyc, uint32_t* d_sink) {
    const int iters = 4000, grid = 148;
    k_probe<MODE><<<grid, warps * 32>>>(d_cyc, d_sink, iters);
    cudaDeviceSynchronize();
    k_probe<MODE><<<grid, warps * 32>>>(d_cyc, d_sink, iters);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148 * 32];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0;
    for (int b = 0; b < grid; ++b)
        for (int w = 0; w < warps; ++w) mx = h[b * 32 + w] > mx ? h[b * 32 + w] : mx;
    // per iteration a warp reads its quarter of one tile; a "tile read" for the SM = 4 warps
    printf("{\"mode\": \"%s\", \"warps\": %d, \"cycles_per_warp_iter\": %.1f, \"sm_cycles_per_128row_tile\": %.1f, \"err\": \"%s\"}\n", name, warps,
           (double)mx / iters, (double)mx / iters / (warps / 4.0), cudaGetErrorString(e));
}

int main() {
    unsigned long long* d_cyc;
    uint32_t* d_sink;
    cudaMalloc(&d_cyc, 148 * 32 * 8);
    cudaMalloc(&d_sink, 148 * 1024 * 4);
    uint32_t* d_out;
    cudaMalloc(&d_out, 64);
    k_layout<<<1, 128>>>(d_out);
    uint32_t ho[16];
    cudaMemcpy(ho, d_out, 64, cudaMemcpyDeviceToHost);
    printf("pack::16b x16 of 32 columns holding (0xA000+i)<<16 | 0x1100+i, lane 1:");
    for (int i = 0; i < 16; ++i) printf(" %08x", ho[i]);
    printf("\n");
    for (int w = 4; w <= 16; w += 4) {
        run<0>("64 cols fp32 (2 x ld.x32)", w, d_cyc, d_sink);
        run<1>("64 cols fp16 packed (1 x ld.x32.pack::16b)", w, d_cyc, d_sink);
        run<3>("64 cols fp16 packed (2 x ld.x16.pack::16b)", w, d_cyc, d_sink);
        run<2>("32 cols fp32 (1 x ld.x32)", w, d_cyc, d_sink);
        run<4>("64 cols fp32 + st.x16 + wait::st", w, d_cyc, d_sink);
    }
    return 0;
}
