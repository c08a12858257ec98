// Level geometry and the per-(sample, level) gather of the multiresolution hash grid, shared by gridencoder.cu and the fused
// inference kernel (field_infer.cu).  Arithmetic follows gridencoder/src/gridencoder.cu (index math :53-71, position :123-136,
// corner order and the per-corner rounding of the accumulator :143-168).
#pragma once
#include "common.cuh"
#include <math.h>
#include <type_traits>

namespace enerf {

// ---- element-type helpers: the accumulator is rounded to T after every corner -----------
template <typename T> struct Elem;
template <> struct Elem<float> {
    static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ float acc(float r, float w, float g) { return __fmaf_rn(w, g, r); }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float to_f(float v) { return v; }
    static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Elem<__half> {
    static __device__ __forceinline__ __half ld(const __half* p) { return __ldg(p); }
    // gridencoder.cu:164 with c10::Half: the fp32 product is rounded to half, added in fp32,
    // rounded to half again.
    static __device__ __forceinline__ __half acc(__half r, float w, __half g) {
        const __half p = __float2half_rn(w * __half2float(g));
        return __float2half_rn(__half2float(r) + __half2float(p));
    }
    static __device__ __forceinline__ __half sub(__half a, __half b) {
        return __float2half_rn(__half2float(a) - __half2float(b));
    }
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};

struct LevelGeom {
    float scale;
    uint32_t resolution, hashmap_size, offset;
    bool use_hash, pow2;
};

__device__ __forceinline__ LevelGeom level_geom(const int32_t* __restrict__ offsets, uint32_t level, float S,
                                                uint32_t H, uint32_t gridtype, int D) {
    LevelGeom g;
    g.offset = (uint32_t)offsets[level];
    g.hashmap_size = (uint32_t)offsets[level + 1] - g.offset;
    g.scale = exp2f((float)level * S) * (float)H - 1.0f;       // gridencoder.cu:124
    g.resolution = (uint32_t)ceilf(g.scale) + 1;               // gridencoder.cu:125
    uint32_t stride = 1;
    for (int d = 0; d < D && stride <= g.hashmap_size; ++d) stride *= (g.resolution + 1);
    g.use_hash = (gridtype == 0) && (stride > g.hashmap_size);
    g.pow2 = (g.hashmap_size & (g.hashmap_size - 1)) == 0;
    return g;
}

// gridencoder.cu:53-71 (element index of channel 0)
template <int D>
__device__ __forceinline__ uint32_t grid_index(const LevelGeom& g, const uint32_t (&p)[D]) {
    uint32_t index;
    if (g.use_hash) {
        constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
        index = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) index ^= p[d] * primes[d];
    } else {
        uint32_t stride = 1;
        index = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            if (stride <= g.hashmap_size) {
                index += p[d] * stride;
                stride *= (g.resolution + 1);
            }
        }
    }
    if (g.pow2) return index & (g.hashmap_size - 1);
    return index < g.hashmap_size ? index : index % g.hashmap_size;
}

// The positions of a launch: [B, D] fp32, either already in [0, 1] (mul == 0: the reference's contract) or raw world coordinates that
// the kernel maps itself, x = (raw + add) * mul — GridEncoder.forward's `(inputs + bound) / (2 * bound)` (grid.py:144) with ATen's own
// operation order and roundings (a division by a Python scalar is a multiplication by its fp32 reciprocal), which saves two
// elementwise passes over the samples per call.
struct Inputs {
    const float* p;
    float add, mul;
};
template <int D>
__device__ __forceinline__ bool load_pos(const Inputs& inputs, uint32_t b, float (&x)[D]) {
    bool oob = false;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        x[d] = __ldg(inputs.p + (size_t)b * D + d);
        if (inputs.mul != 0.0f) x[d] = __fmul_rn(__fadd_rn(x[d], inputs.add), inputs.mul);
        oob |= (x[d] < 0.0f) || (x[d] > 1.0f);
    }
    return oob;
}

// ---- "a warp owns 32 consecutive samples and walks all levels" (k_grid_fwd_w, k_field_infer): per-level table in shared memory,
// one of three branch-free corner-index forms per level, corner values fetched first and blended in the reference's order
struct LevelTabW {
    float scale;
    uint32_t hs;       // table entries
    uint32_t offset;   // first entry of the level
    uint32_t m1, m2;   // index = x*1 (+|^) y*m1 (+|^) z*m2
    uint32_t mode;     // 0: hashed, pow2 table; 1: dense, index < hs always; 2: generic (bit0 of flags = hashed, bit1 = pow2)
    uint32_t flags;
    uint32_t pad;
};

template <int C, int MODE>
__device__ __forceinline__ void corner_index(const LevelTabW& lt, const uint32_t (&pg)[3], uint32_t (&e)[8]) {
    const uint32_t ax[2] = {pg[0], pg[0] + 1};
    const uint32_t ay[2] = {pg[1] * lt.m1, pg[1] * lt.m1 + lt.m1};
    const uint32_t az[2] = {pg[2] * lt.m2, pg[2] * lt.m2 + lt.m2};
    if (MODE == 0) {
        const uint32_t mask = lt.hs - 1;
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) e[idx] = ((ax[idx & 1] ^ ay[(idx >> 1) & 1] ^ az[idx >> 2]) & mask) * C;
    } else if (MODE == 1) {
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) e[idx] = (ax[idx & 1] + ay[(idx >> 1) & 1] + az[idx >> 2]) * C;
    } else {
        const bool hashed = lt.flags & 1u, pow2 = lt.flags & 2u;
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) {
            const uint32_t a = ax[idx & 1], bq = ay[(idx >> 1) & 1], c = az[idx >> 2];
            uint32_t i = hashed ? (a ^ bq ^ c) : (a + bq + c);
            i = pow2 ? (i & (lt.hs - 1)) : (i < lt.hs ? i : i % lt.hs);
            e[idx] = i * C;
        }
    }
}

// one (sample, level): corner values fetched (`fetch`), then blended in the reference's order (`blend`)
template <typename T, int C>
struct LevelWork {
    float wxy[4], wz[2];
    T v[8][C];
    template <int MODE>
    __device__ __forceinline__ void fetch(const LevelTabW& lt, const T* __restrict__ grid, const float (&x)[3]) {
        const T* __restrict__ tab = grid + (size_t)lt.offset * C;
        float fr[3];
        uint32_t pg[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float p = __fmaf_rn(x[d], lt.scale, 0.5f);
            const float fl = floorf(p);
            pg[d] = (uint32_t)fl;
            fr[d] = p - fl;
        }
        uint32_t e[8];
        corner_index<C, MODE>(lt, pg, e);
        if (C == 2 && sizeof(T) == 2) {
#pragma unroll
            for (int idx = 0; idx < 8; ++idx) {
                const __half2 h2 = __ldg(reinterpret_cast<const __half2*>(tab + e[idx]));
                v[idx][0] = *reinterpret_cast<const T*>(&h2.x);
                v[idx][C - 1] = *reinterpret_cast<const T*>(&h2.y);
            }
        } else {
#pragma unroll
            for (int idx = 0; idx < 8; ++idx)
#pragma unroll
                for (int ch = 0; ch < C; ++ch) v[idx][ch] = Elem<T>::ld(tab + e[idx] + ch);
        }
        const float wx0 = 1.0f - fr[0], wy0 = 1.0f - fr[1];
        wxy[0] = wx0 * wy0; wxy[1] = fr[0] * wy0; wxy[2] = wx0 * fr[1]; wxy[3] = fr[0] * fr[1];
        wz[0] = 1.0f - fr[2]; wz[1] = fr[2];
    }
    __device__ __forceinline__ void blend(uint32_t (&ow)[(C * (int)sizeof(T)) / 4], bool zero) const {
        constexpr int WPL = (C * (int)sizeof(T)) / 4;
        if (C == 2 && sizeof(T) == 2) {
            __half2 res2 = __floats2half2_rn(0.f, 0.f);
#pragma unroll
            for (int idx = 0; idx < 8; ++idx) {
                const float w = wxy[idx & 3] * wz[idx >> 2];
                const float gx = Elem<T>::to_f(v[idx][0]), gy = Elem<T>::to_f(v[idx][C - 1]);
                res2 = __hadd2(res2, __floats2half2_rn(w * gx, w * gy));
            }
            ow[0] = zero ? 0u : *reinterpret_cast<const uint32_t*>(&res2);
        } else {
            T res[C];
#pragma unroll
            for (int ch = 0; ch < C; ++ch) res[ch] = Elem<T>::from_f(0.f);
#pragma unroll
            for (int idx = 0; idx < 8; ++idx) {
                const float w = wxy[idx & 3] * wz[idx >> 2];
#pragma unroll
                for (int ch = 0; ch < C; ++ch) res[ch] = Elem<T>::acc(res[ch], w, v[idx][ch]);
            }
            const uint32_t* rwords = reinterpret_cast<const uint32_t*>(res);
#pragma unroll
            for (int w = 0; w < WPL; ++w) ow[w] = zero ? 0u : rwords[w];
        }
    }
};

// Fills ltab[0..L) (L <= 64) and plan = {number of leading dense (mode 1) levels, 1 if some other level needs the generic form};
// every thread of the CTA calls it (two CTA-wide barriers inside).
__device__ __forceinline__ void build_level_table(LevelTabW* ltab, uint32_t* plan, const int32_t* __restrict__ offsets, uint32_t L, float S, uint32_t H,
                                                  uint32_t gridtype, uint32_t tid) {
    if (tid < L) {
        const LevelGeom g = level_geom(offsets, tid, S, H, gridtype, 3);
        LevelTabW t;
        t.scale = g.scale;
        t.hs = g.hashmap_size;
        t.offset = g.offset;
        const uint32_t r1 = g.resolution + 1;
        bool nowrap = false;
        if (g.use_hash) {
            t.m1 = 2654435761u;
            t.m2 = 805459861u;
        } else {                                 // the stride loop of gridencoder.cu:58-62, dimension by dimension
            t.m1 = (r1 <= g.hashmap_size) ? r1 : 0u;
            t.m2 = (t.m1 != 0u && r1 * r1 <= g.hashmap_size) ? r1 * r1 : 0u;
            // corner coordinates are <= resolution, so the largest dense index is resolution*(1+m1+m2)
            nowrap = (uint64_t)g.resolution * (1ull + t.m1 + t.m2) < (uint64_t)g.hashmap_size;
        }
        t.flags = (g.use_hash ? 1u : 0u) | (g.pow2 ? 2u : 0u);
        t.mode = (g.use_hash && g.pow2) ? 0u : (nowrap ? 1u : 2u);
        t.pad = 0;
        ltab[tid] = t;
    }
    __syncthreads();
    if (tid == 0) {
        // the usual table is "dense levels first, hashed power-of-two levels after": two branch-free loops
        uint32_t nd = 0, generic = 0;
        while (nd < L && ltab[nd].mode == 1u) ++nd;
        for (uint32_t l = nd; l < L; ++l) generic |= (ltab[l].mode != 0u) ? 1u : 0u;
        plan[0] = nd;
        plan[1] = generic;
    }
    __syncthreads();
}

}  // namespace enerf
