// Shared host/device helpers for the enerf_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/enerf_b200.h"

namespace enerf {

// ---- error state (thread-local message + launch counter) ---------------------------------
void set_error(const char* fmt, ...);
void count_launch(unsigned n = 1);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Checks the launch that just happened; returns 0 / non-zero like the ABI.
#define ENERF_CHECK_LAUNCH(name)                                                     \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            ::enerf::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__)); \
            return (int)e__ ? (int)e__ : -1;                                         \
        }                                                                            \
        ::enerf::count_launch();                                                     \
    } while (0)

#define ENERF_CUDA(call, name)                                                       \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            ::enerf::set_error("%s: %s failed: %s", name, #call, cudaGetErrorString(e__)); \
            return (int)e__;                                                         \
        }                                                                            \
    } while (0)

#define ENERF_REQUIRE(cond, name, msg)                                               \
    do {                                                                             \
        if (!(cond)) {                                                               \
            ::enerf::set_error("%s: %s", name, msg);                                 \
            return -2;                                                               \
        }                                                                            \
    } while (0)

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

// SMs of the current device (148 on B200), queried once per device through cudaDeviceGetAttribute: persistent grids are sized from it
int num_sms();

// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

// 10-bit -> 30-bit interleave (Morton); the bit pattern is fixed by the density-grid file
// format (raymarching.cu:58-83), the implementation is the classic magic-number spread.
__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {
    v = (v | (v << 16)) & 0xFF0000FFu;  // v < 1024 so the multiply form and the or form agree
    v = (v | (v << 8)) & 0x0F00F00Fu;
    v = (v | (v << 4)) & 0xC30C30C3u;
    v = (v | (v << 2)) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xC30C30C3u;
    x = (x | (x >> 4)) & 0x0F00F00Fu;
    x = (x | (x >> 8)) & 0xFF0000FFu;
    x = (x | (x >> 16)) & 0x0000FFFFu;
    return x;
}
__host__ __device__ __forceinline__ uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}

// PCG32 (O'Neill) — must be bit-exact with raymarching/src/pcg32.h:57-72,107-116 because the
// marcher's jitter is part of its output.
struct Pcg32 {
    uint64_t state, inc;
    __host__ __device__ Pcg32(uint64_t initstate, uint64_t initseq) {
        state = 0u;
        inc = (initseq << 1u) | 1u;
        next_uint();
        state += initstate;
        next_uint();
    }
    __host__ __device__ uint32_t next_uint() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31));
    }
    __host__ __device__ float next_float() {
        uint32_t u = (next_uint() >> 9) | 0x3f800000u;
#ifdef __CUDA_ARCH__
        return __uint_as_float(u) - 1.0f;
#else
        float f;
        memcpy(&f, &u, 4);
        return f - 1.0f;
#endif
    }
};

}  // namespace enerf
