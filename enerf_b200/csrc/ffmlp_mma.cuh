// Warp-level building blocks of the fully-fused MLP: legacy tensor-core path (mma.sync
// m16n8k16, fp16 in / fp32 accumulate).  Used for every (hidden_dim, num_layers) shape; the
// 64-wide networks E-NeRF runs take the tcgen05/TMEM kernels in ffmlp_tc.cu instead.
#pragma once
#include "common.cuh"

namespace enerf {
namespace mlp {

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t lds32(const __half* p) { return *reinterpret_cast<const uint32_t*>(p); }

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// ffmlp/src/utils.h:427-470 (forward) — value semantics only
__device__ __forceinline__ float act_fwd(uint32_t act, float x) {
    switch (act) {
        case ENERF_ACT_RELU: return x > 0.f ? x : 0.f;
        case ENERF_ACT_EXPONENTIAL: return expf(x);
        case ENERF_ACT_SINE: return sinf(x);
        case ENERF_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
        case ENERF_ACT_SQUAREPLUS: { const float v = x * 10.0f; return 0.5f * (v + sqrtf(v * v + 4.0f)) / 10.0f; }
        case ENERF_ACT_SOFTPLUS: return logf(expf(x * 10.0f) + 1.0f) / 10.0f;
        default: return x;
    }
}
// ffmlp/src/utils.h:540-578 (backward through the activation, given the stored POST-activation y)
__device__ __forceinline__ float act_bwd(uint32_t act, float g, float y) {
    switch (act) {
        case ENERF_ACT_RELU: return y > 0.f ? g : 0.f;
        case ENERF_ACT_EXPONENTIAL: return g * y;
        case ENERF_ACT_SINE: return g;  // the reference leaves the gradient untouched (needs pre-activations)
        case ENERF_ACT_SIGMOID: return g * (y * (1.0f - y));
        case ENERF_ACT_SQUAREPLUS: { const float v = y * 10.0f; return g * (v * v / (v * v + 1.0f)); }
        case ENERF_ACT_SOFTPLUS: return g * (1.0f - expf(-y * 10.0f));
        default: return g;
    }
}

// acc[MT][NT] (+)= A[16*MT rows, K] * B^T, A = activations in smem (row-major, lda halves),
// Bs[n][k] = weights in smem (row n = output neuron, ldb halves).  Only n-tiles < nt_valid are
// computed.  Fragment layouts: PTX ISA, mma.m16n8k16 .f16.
template <int MT, int NT>
__device__ __forceinline__ void warp_gemm(float (&acc)[MT][NT][4], const __half* __restrict__ As, int lda,
                                          const __half* __restrict__ Bs, int ldb, int K, int nt_valid) {
    const unsigned lane = threadIdx.x & 31u, g = lane >> 2, tg = lane & 3u;
    for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t a[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const __half* p = As + (mt * 16 + g) * lda + k0 + 2 * tg;
            a[mt][0] = lds32(p);
            a[mt][1] = lds32(p + 8 * lda);
            a[mt][2] = lds32(p + 8);
            a[mt][3] = lds32(p + 8 * lda + 8);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            if (nt < nt_valid) {
                const __half* q = Bs + (nt * 8 + g) * ldb + k0 + 2 * tg;
                const uint32_t b0 = lds32(q), b1 = lds32(q + 8);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) mma_16816(acc[mt][nt], a[mt], b0, b1);
            }
        }
    }
}

template <int MT, int NT>
__device__ __forceinline__ void zero_acc(float (&acc)[MT][NT][4]) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
}

// warp-private copy of `rows` rows x `cols` halves (cols % 8 == 0) between global (ld = ldg) and
// shared (ld = lds); 16-byte accesses, fully coalesced on the global side.
__device__ __forceinline__ void warp_copy_g2s(__half* __restrict__ s, int lds_, const __half* __restrict__ g, size_t ldg,
                                              int rows, int cols) {
    const int vec_per_row = cols >> 3, total = rows * vec_per_row;
    for (int i = threadIdx.x & 31; i < total; i += 32) {
        const int r = i / vec_per_row, c = (i - r * vec_per_row) << 3;
        *reinterpret_cast<int4*>(s + r * lds_ + c) = __ldg(reinterpret_cast<const int4*>(g + (size_t)r * ldg + c));
    }
}
__device__ __forceinline__ void warp_copy_s2g(__half* __restrict__ g, size_t ldg, const __half* __restrict__ s, int lds_,
                                              int rows, int cols) {
    const int vec_per_row = cols >> 3, total = rows * vec_per_row;
    for (int i = threadIdx.x & 31; i < total; i += 32) {
        const int r = i / vec_per_row, c = (i - r * vec_per_row) << 3;
        *reinterpret_cast<int4*>(g + (size_t)r * ldg + c) = *reinterpret_cast<const int4*>(s + r * lds_ + c);
    }
}

}  // namespace mlp
}  // namespace enerf
