// Multiresolution hash-grid encoder (gather forward, scatter-add backward) for sm_100a.
//
// Replaces the reference's `_gridencoder` extension (gridencoder/src/gridencoder.h:12-13).
// Arithmetic follows gridencoder/src/gridencoder.cu (index math :53-71, position :123-136,
// corner order and the per-corner rounding of the accumulator :143-168, backward :258-310).
//
// Execution model (differs from the reference's one-thread-per-(sample,level) 1-D blocks):
//   * a CTA owns 32 consecutive samples and ALL levels: blockDim = (32 lanes, W warps), warp w
//     walks levels w, w+W, ...  Consecutive samples come from the same ray (marcher order), so a
//     warp's 32 lanes hit the same or neighbouring cells of one level: the 8 corner gathers of a
//     warp collapse into a few 32-B sectors on the coarse levels.
//   * the 16 warps of a CTA read the sample coordinates once from HBM (the other 15 hit L1).
//   * outputs are staged in shared memory and written as [B, L*C] rows with fully coalesced
//     stores — the reference writes [L,B,C] and pays a separate permute copy (grid.py:52).
//   * backward reads grad rows in the same [B, L*C] layout (no permute/contiguous copy,
//     grid.py:70) and scatters with vector reductions (red.global.add.v2.f32 / .noftz.f16x2).
#include "common.cuh"
#include "grid_levels.cuh"
#include <math.h>
#include <type_traits>

namespace enerf {

static constexpr int kSamplesPerCta = 32;
static int g_fwd_fast = 1;   // D = 3 without input gradients: 1 = k_grid_fwd_w (warp walks the levels), 2 = k_grid_fwd3; 0 = always the generic kernel
constexpr int kBwdBlockDefault = 128;   // measured (3.29 M samples): 256: 0.685, 192: 0.688, 128: 0.675, 64: 0.673 ms
static int g_bwd_block = kBwdBlockDefault;  // threads per CTA of the walking scatter (enerf_grid_set_backward_block)
static int g_bwd_walk = 1;   // 1: walking scatter (register aggregation along rays), 0: one reduction per corner


// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <typename T, int D, int C, bool BLC>
__global__ void __launch_bounds__(512)
k_grid_fwd(const Inputs inputs, const T* __restrict__ grid, const int32_t* __restrict__ offsets,
           T* __restrict__ outputs, uint32_t B, uint32_t L, float S, uint32_t H, bool calc_grad_inputs,
           T* __restrict__ dy_dx, uint32_t gridtype, uint32_t row_stride /* staging row, in T */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* stage = reinterpret_cast<T*>(smem_raw);

    const unsigned lane = threadIdx.x;
    const uint32_t b0 = blockIdx.x * kSamplesPerCta;
    const uint32_t b = b0 + lane;
    const bool active = b < B;

    float x[D];
    bool oob = true;
    if (active) oob = load_pos<D>(inputs, b, x);

    for (uint32_t level = threadIdx.y; level < L; level += blockDim.y) {
        T res[C];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) res[ch] = Elem<T>::from_f(0.f);

        if (active && !oob) {
            const LevelGeom g = level_geom(offsets, level, S, H, gridtype, D);
            const T* __restrict__ tab = grid + (size_t)g.offset * C;
            float pos[D];
            uint32_t pg[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                pos[d] = __fmaf_rn(x[d], g.scale, 0.5f);
                const float fl = floorf(pos[d]);
                pg[d] = (uint32_t)fl;
                pos[d] -= fl;
            }
            // gather all corners first (independent loads in flight), then blend in order
            T v[1 << D][C];
            float w[1 << D];
#pragma unroll
            for (int idx = 0; idx < (1 << D); ++idx) {
                float ww = 1.0f;
                uint32_t pl[D];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    if ((idx & (1 << d)) == 0) { ww *= 1.0f - pos[d]; pl[d] = pg[d]; }
                    else { ww *= pos[d]; pl[d] = pg[d] + 1; }
                }
                w[idx] = ww;
                const uint32_t e = grid_index<D>(g, pl) * C;
                if (C == 2 && sizeof(T) == 2) {
                    const __half2 h2 = __ldg(reinterpret_cast<const __half2*>(tab + e));
                    v[idx][0] = *reinterpret_cast<const T*>(&h2.x);
                    v[idx][C - 1] = *reinterpret_cast<const T*>(&h2.y);
                } else {
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) v[idx][ch] = Elem<T>::ld(tab + e + ch);
                }
            }
#pragma unroll
            for (int idx = 0; idx < (1 << D); ++idx)
#pragma unroll
                for (int ch = 0; ch < C; ++ch) res[ch] = Elem<T>::acc(res[ch], w[idx], v[idx][ch]);

            if (calc_grad_inputs) {
                // gridencoder.cu:178-220: d(out)/d(x_gd), layout [B, L, D, C]
                T* __restrict__ dd = dy_dx + ((size_t)b * L + level) * D * C;
#pragma unroll
                for (int gd = 0; gd < D; ++gd) {
                    T rg[C];
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) rg[ch] = Elem<T>::from_f(0.f);
#pragma unroll
                    for (int idx = 0; idx < (1 << (D - 1)); ++idx) {
                        float ww = g.scale;
                        uint32_t pl[D];
#pragma unroll
                        for (int nd = 0; nd < D - 1; ++nd) {
                            const int d = (nd >= gd) ? (nd + 1) : nd;
                            if ((idx & (1 << nd)) == 0) { ww *= 1.0f - pos[d]; pl[d] = pg[d]; }
                            else { ww *= pos[d]; pl[d] = pg[d] + 1; }
                        }
                        pl[gd] = pg[gd];
                        const uint32_t el = grid_index<D>(g, pl) * C;
                        pl[gd] = pg[gd] + 1;
                        const uint32_t er = grid_index<D>(g, pl) * C;
#pragma unroll
                        for (int ch = 0; ch < C; ++ch)
                            rg[ch] = Elem<T>::acc(rg[ch], ww, Elem<T>::sub(Elem<T>::ld(tab + er + ch), Elem<T>::ld(tab + el + ch)));
                    }
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) dd[gd * C + ch] = rg[ch];
                }
            }
        } else if (active && calc_grad_inputs) {
            T* __restrict__ dd = dy_dx + ((size_t)b * L + level) * D * C;
            for (int i = 0; i < D * C; ++i) dd[i] = Elem<T>::from_f(0.f);
        }

        if (BLC) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) stage[lane * row_stride + level * C + ch] = res[ch];
        } else if (active) {
            T* __restrict__ o = outputs + ((size_t)level * B + b) * C;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) o[ch] = res[ch];
        }
    }

    if (BLC) {
        __syncthreads();
        // [32, L*C] tile -> contiguous global rows; 4-byte words when the row is word-sized
        const uint32_t row_elems = L * C;
        const uint32_t n_rows = min((uint32_t)kSamplesPerCta, B - b0);
        const uint32_t tid = threadIdx.y * 32 + lane, nthr = blockDim.y * 32;
        T* __restrict__ out = outputs + (size_t)b0 * row_elems;
        if (((row_elems * sizeof(T)) & 3u) == 0 && ((row_stride * sizeof(T)) & 3u) == 0) {
            const uint32_t rw = row_elems * sizeof(T) / 4, sw = row_stride * sizeof(T) / 4;
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(stage);
            uint32_t* o32 = reinterpret_cast<uint32_t*>(out);
            for (uint32_t k = tid; k < n_rows * rw; k += nthr) o32[k] = s32[(k / rw) * sw + (k % rw)];
        } else {
            for (uint32_t k = tid; k < n_rows * row_elems; k += nthr) out[k] = stage[(k / row_elems) * row_stride + (k % row_elems)];
        }
    }
}

// ------------------------------------------------------------------------------------------
// forward, hot path (D = 3, no input gradients): same arithmetic as k_grid_fwd with the per-level
// constants computed once per CTA (shared memory), per-axis index terms and weights hoisted out
// of the corner loop (w = (wx*wy)*wz in the reference's order), and packed fp16x2 arithmetic for
// the fp16/C=2 table: cvt.rn.f16x2.f32 of the two fp32 products + one HADD2 per corner is
// bit-identical to the reference's "round the product, add in fp32, round again" because the sum
// of two fp16 numbers is either exact in fp32 or dominated by the larger operand.
// ------------------------------------------------------------------------------------------
struct LevelTab {
    float scale;
    uint32_t hs, offset, m1, m2, flags;   // flags: bit0 = hashed, bit1 = table size is a power of two
};

template <typename T, int C, bool BLC>
__global__ void __launch_bounds__(512)
k_grid_fwd3(const Inputs inputs, const T* __restrict__ grid, const int32_t* __restrict__ offsets,
            T* __restrict__ outputs, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, uint32_t row_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* stage = reinterpret_cast<T*>(smem_raw);
    __shared__ LevelTab ltab[64];

    const unsigned lane = threadIdx.x;
    const uint32_t tid = threadIdx.y * 32 + lane;
    if (tid < L) {
        const LevelGeom g = level_geom(offsets, tid, S, H, gridtype, 3);
        LevelTab t;
        t.scale = g.scale;
        t.hs = g.hashmap_size;
        t.offset = g.offset;
        const uint32_t r1 = g.resolution + 1;
        if (g.use_hash) {
            t.m1 = 2654435761u;
            t.m2 = 805459861u;
        } else {                                 // the stride loop of gridencoder.cu:58-62, dimension by dimension
            t.m1 = (r1 <= g.hashmap_size) ? r1 : 0u;
            t.m2 = (t.m1 != 0u && r1 * r1 <= g.hashmap_size) ? r1 * r1 : 0u;
        }
        t.flags = (g.use_hash ? 1u : 0u) | (g.pow2 ? 2u : 0u);
        ltab[tid] = t;
    }
    __syncthreads();

    const uint32_t b0 = blockIdx.x * kSamplesPerCta;
    const uint32_t b = b0 + lane;
    const bool active = b < B;
    float x[3];
    bool oob = true;
    if (active) oob = load_pos<3>(inputs, b, x);

    for (uint32_t level = threadIdx.y; level < L; level += blockDim.y) {
        T res[C];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) res[ch] = Elem<T>::from_f(0.f);
        __half2 res2 = __floats2half2_rn(0.f, 0.f);

        if (active && !oob) {
            const LevelTab lt = ltab[level];
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
            const float wx[2] = {1.0f - fr[0], fr[0]}, wy[2] = {1.0f - fr[1], fr[1]}, wz[2] = {1.0f - fr[2], fr[2]};
            const float wxy[4] = {wx[0] * wy[0], wx[1] * wy[0], wx[0] * wy[1], wx[1] * wy[1]};
            const uint32_t ax[2] = {pg[0], pg[0] + 1};
            const uint32_t ay[2] = {pg[1] * lt.m1, (pg[1] + 1) * lt.m1};
            const uint32_t az[2] = {pg[2] * lt.m2, (pg[2] + 1) * lt.m2};
            const bool hashed = lt.flags & 1u, pow2 = lt.flags & 2u;
            uint32_t e[8];
#pragma unroll
            for (int idx = 0; idx < 8; ++idx) {
                const uint32_t a = ax[idx & 1], bq = ay[(idx >> 1) & 1], c = az[idx >> 2];
                uint32_t i = hashed ? (a ^ bq ^ c) : (a + bq + c);
                i = pow2 ? (i & (lt.hs - 1)) : (i < lt.hs ? i : i % lt.hs);
                e[idx] = i * C;
            }
            if (C == 2 && sizeof(T) == 2) {
                __half2 v[8];
#pragma unroll
                for (int idx = 0; idx < 8; ++idx) v[idx] = __ldg(reinterpret_cast<const __half2*>(tab + e[idx]));
#pragma unroll
                for (int idx = 0; idx < 8; ++idx) {
                    const float w = wxy[idx & 3] * wz[idx >> 2];
                    const float2 g = __half22float2(v[idx]);
                    res2 = __hadd2(res2, __floats2half2_rn(w * g.x, w * g.y));
                }
            } else {
                T v[8][C];
#pragma unroll
                for (int idx = 0; idx < 8; ++idx)
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) v[idx][ch] = Elem<T>::ld(tab + e[idx] + ch);
#pragma unroll
                for (int idx = 0; idx < 8; ++idx) {
                    const float w = wxy[idx & 3] * wz[idx >> 2];
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) res[ch] = Elem<T>::acc(res[ch], w, v[idx][ch]);
                }
            }
        }
        if (C == 2 && sizeof(T) == 2) {
            res[0] = *reinterpret_cast<const T*>(&res2.x);
            res[C - 1] = *reinterpret_cast<const T*>(&res2.y);
        }

        if (BLC) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) stage[lane * row_stride + level * C + ch] = res[ch];
        } else if (active) {
            T* __restrict__ o = outputs + ((size_t)level * B + b) * C;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) o[ch] = res[ch];
        }
    }

    if (BLC) {
        __syncthreads();
        const uint32_t row_elems = L * C;
        const uint32_t n_rows = min((uint32_t)kSamplesPerCta, B - b0);
        const uint32_t nthr = blockDim.y * 32;
        T* __restrict__ out = outputs + (size_t)b0 * row_elems;
        if (((row_elems * sizeof(T)) & 3u) == 0 && ((row_stride * sizeof(T)) & 3u) == 0) {
            const uint32_t rw = row_elems * sizeof(T) / 4, sw = row_stride * sizeof(T) / 4;
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(stage);
            uint32_t* o32 = reinterpret_cast<uint32_t*>(out);
            for (uint32_t k = tid; k < n_rows * rw; k += nthr) o32[k] = s32[(k / rw) * sw + (k % rw)];
        } else {
            for (uint32_t k = tid; k < n_rows * row_elems; k += nthr) out[k] = stage[(k / row_elems) * row_stride + (k % row_elems)];
        }
    }
}

// ------------------------------------------------------------------------------------------
// forward, hot path v2 (D = 3, no input gradients): "a warp owns 32 consecutive samples and walks
// all levels".  Differences to k_grid_fwd3 (one warp per (32 samples, level)):
//   * persistent CTAs: the level table is built once per CTA, not once per 32 samples, and there is
//     no CTA-wide barrier in the sample loop (staging tiles are warp-private);
//   * sample coordinates are read once per sample instead of once per (sample, level);
//   * the corner index comes from one of three branch-free forms chosen per level (warp-uniform):
//     hashed & power-of-two table (xor, and), dense without wrap-around (add), generic;
//   * two levels are in flight per thread (16 independent gathers) before the first blend.
// Arithmetic (fma position, corner order, per-corner fp16 rounding) is unchanged: bit-identical.
// ------------------------------------------------------------------------------------------

template <typename T, int C, bool BLC>
__global__ void __launch_bounds__(256)
k_grid_fwd_w(const Inputs inputs, const T* __restrict__ grid, const int32_t* __restrict__ offsets,
             T* __restrict__ outputs, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, uint32_t n_groups,
             uint32_t row_words /* staging row stride in 32-bit words (odd) */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ LevelTabW ltab[64];
    __shared__ uint32_t s_plan[2];                     // [0]: number of leading dense (mode 1) levels, [1]: 1 = some level needs the generic form
    constexpr int WPL = (C * (int)sizeof(T)) / 4;      // 32-bit words per level in an output row (>= 1 on this path)

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    build_level_table(ltab, s_plan, offsets, L, S, H, gridtype, tid);
    const uint32_t n_dense = s_plan[1] ? 0u : s_plan[0];
    const bool generic = s_plan[1] != 0u;

    uint32_t* stage = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)warp * 32 * row_words;
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    const uint32_t rw = L * WPL;
    const int rw_shift = ((rw & (rw - 1)) == 0) ? (31 - __clz(rw)) : -1;

    for (uint32_t group = blockIdx.x * (blockDim.x >> 5) + warp; group < n_groups; group += warps_total) {
        const uint32_t b0 = group * 32u;
        const uint32_t b = b0 + lane;
        const bool active = b < B;
        float x[3] = {0.f, 0.f, 0.f};
        bool oob = true;
        if (active) oob = load_pos<3>(inputs, b, x);
        if (oob) x[0] = x[1] = x[2] = 0.f;         // keep the address arithmetic in range; the result is zeroed below

        auto emit = [&](uint32_t level, const uint32_t (&ow)[WPL]) {
            if (BLC) {
#pragma unroll
                for (int w = 0; w < WPL; ++w) stage[lane * row_words + level * WPL + w] = ow[w];
            } else if (active) {
                uint32_t* __restrict__ o = reinterpret_cast<uint32_t*>(outputs + ((size_t)level * B + b) * C);
#pragma unroll
                for (int w = 0; w < WPL; ++w) o[w] = ow[w];
            }
        };
        // levels [lo, hi) with the index form MODE, two levels (16 gathers) in flight per thread
        auto run = [&](auto mode_tag, uint32_t lo, uint32_t hi) {
            constexpr int MODE = decltype(mode_tag)::value;
            uint32_t level = lo;
            for (; level + 1 < hi; level += 2) {
                LevelWork<T, C> wa, wb;
                wa.template fetch<MODE>(ltab[level], grid, x);
                wb.template fetch<MODE>(ltab[level + 1], grid, x);
                uint32_t oa[WPL], ob[WPL];
                wa.blend(oa, oob);
                wb.blend(ob, oob);
                emit(level, oa);
                emit(level + 1, ob);
            }
            if (level < hi) {
                LevelWork<T, C> wa;
                wa.template fetch<MODE>(ltab[level], grid, x);
                uint32_t oa[WPL];
                wa.blend(oa, oob);
                emit(level, oa);
            }
        };
        if (generic) {
            run(std::integral_constant<int, 2>{}, 0u, L);
        } else {
            run(std::integral_constant<int, 1>{}, 0u, n_dense);
            run(std::integral_constant<int, 0>{}, n_dense, L);
        }

        if (BLC) {
            __syncwarp();
            const uint32_t n_words = min(32u, B - b0) * rw;
            uint32_t* __restrict__ o32 = reinterpret_cast<uint32_t*>(outputs + (size_t)b0 * L * C);
            if (rw_shift >= 0) {
                for (uint32_t k = lane; k < n_words; k += 32) o32[k] = stage[(k >> rw_shift) * row_words + (k & (rw - 1))];
            } else {
                for (uint32_t k = lane; k < n_words; k += 32) o32[k] = stage[(k / rw) * row_words + (k % rw)];
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward (scatter-add into the gradient table)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add(float* addr, float a) { atomicAdd(addr, a); }
__device__ __forceinline__ void red_add2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(__half* addr, float a) { atomicAdd(addr, __float2half_rn(a)); }
__device__ __forceinline__ void red_add2(__half* addr, float a, float b) {
    const __half2 v = __halves2half2(__float2half_rn(a), __float2half_rn(b));   // gridencoder.cu:300
    atomicAdd(reinterpret_cast<__half2*>(addr), v);
}

// T = table / grad element type, G = gradient-table element type (T, or float for fp32 accumulation)
template <typename T, typename G, int D, int C, bool BLC>
__global__ void __launch_bounds__(512)
k_grid_bwd(const T* __restrict__ grad, const Inputs inputs, const int32_t* __restrict__ offsets,
           G* __restrict__ grad_grid, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype) {
    const unsigned lane = threadIdx.x;
    const uint32_t b = blockIdx.x * kSamplesPerCta + lane;
    if (b >= B) return;
    float x[D];
    if (load_pos<D>(inputs, b, x)) return;  // gridencoder.cu:250-256

    for (uint32_t level = threadIdx.y; level < L; level += blockDim.y) {
        const LevelGeom g = level_geom(offsets, level, S, H, gridtype, D);
        G* __restrict__ tab = grad_grid + (size_t)g.offset * C;
        float pos[D];
        uint32_t pg[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            pos[d] = __fmaf_rn(x[d], g.scale, 0.5f);
            const float fl = floorf(pos[d]);
            pg[d] = (uint32_t)fl;
            pos[d] -= fl;
        }
        float gr[C];
        const T* __restrict__ gp = BLC ? grad + ((size_t)b * L + level) * C : grad + ((size_t)level * B + b) * C;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) gr[ch] = Elem<T>::to_f(gp[ch]);

#pragma unroll
        for (int idx = 0; idx < (1 << D); ++idx) {
            float w = 1.0f;
            uint32_t pl[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                if ((idx & (1 << d)) == 0) { w *= 1.0f - pos[d]; pl[d] = pg[d]; }
                else { w *= pos[d]; pl[d] = pg[d] + 1; }
            }
            const uint32_t e = grid_index<D>(g, pl) * C;
            if (C == 1) {
                red_add(tab + e, w * gr[0]);
            } else {
#pragma unroll
                for (int ch = 0; ch < C; ch += 2) red_add2(tab + e + ch, w * gr[ch], w * gr[ch + (C > 1 ? 1 : 0)]);
            }
        }
    }
}

// Scatter-add, "walking" formulation.  A thread owns ONE level and a run of SEG consecutive samples.
// Consecutive samples come from the same ray (the marcher emits them in order), so on coarse and
// middle levels they fall into the same cell for many steps: the thread accumulates the 2^D corner
// contributions in registers and only touches memory (2^D vector reductions) when the cell
// changes.  Lanes of a warp are (level = lane % L, run = lane / L): the L lanes of one run read
// one contiguous L*C-element row of `grad` per step (coalesced) and broadcast-load the position.
// Atomic traffic drops by the mean run length per level (~25x on level 0, none on the finest
// hashed levels) and, more importantly, the same-address contention on the small coarse tables
// disappears.  Same sums as k_grid_bwd up to fp32 reassociation.
template <typename T, typename G, int D, int C, bool BLC, int SEG>
__global__ void __launch_bounds__(256)
k_grid_bwd_walk(const T* __restrict__ grad, const Inputs inputs, const int32_t* __restrict__ offsets,
                G* __restrict__ grad_grid, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t level = t % L;
    const uint32_t run = t / L;
    const uint64_t b0 = (uint64_t)run * SEG;
    if (b0 >= B) return;
    const uint32_t b1 = (uint32_t)min((uint64_t)B, b0 + SEG);

    // per-level constants (one level per thread): index = (x*1 (+|^) y*m[1] (+|^) z*m[2]) mod table size
    const LevelGeom g = level_geom(offsets, level, S, H, gridtype, D);
    G* __restrict__ tab = grad_grid + (size_t)g.offset * C;
    uint32_t mult[D];
    {
        constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
        uint32_t stride = 1;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            if (g.use_hash) mult[d] = primes[d];
            else {                                          // the stride loop of gridencoder.cu:58-62
                mult[d] = (stride <= g.hashmap_size) ? stride : 0u;
                if (stride <= g.hashmap_size) stride *= (g.resolution + 1);
            }
        }
    }

    uint32_t cell[D];
#pragma unroll
    for (int d = 0; d < D; ++d) cell[d] = 0xffffffffu;
    bool have = false;
    float acc[1 << D][C];

    auto flush = [&]() {
        uint32_t term[D][2];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            term[d][0] = cell[d] * mult[d];
            term[d][1] = (cell[d] + 1) * mult[d];
        }
        uint32_t e[1 << D];
#pragma unroll
        for (int idx = 0; idx < (1 << D); ++idx) {
            uint32_t i = 0;
#pragma unroll
            for (int d = 0; d < D; ++d) i = g.use_hash ? (i ^ term[d][(idx >> d) & 1]) : (i + term[d][(idx >> d) & 1]);
            i = g.pow2 ? (i & (g.hashmap_size - 1)) : (i < g.hashmap_size ? i : i % g.hashmap_size);
            e[idx] = i * C;
        }
        if (C == 2 && sizeof(G) == 4) {
            // The two corners along x are neighbouring table entries (index ^ 1 when hashed with prime 1, index + 1
            // when dense): whenever they share an aligned 16-byte block one 128-bit reduction updates both.
#pragma unroll
            for (int idx = 0; idx < (1 << D); idx += 2) {
                const uint32_t e0 = e[idx], e1 = e[idx + 1];
                if ((e0 ^ e1) == 2u) {
                    const bool lo = e0 < e1;
                    red_add4(reinterpret_cast<float*>(tab) + (lo ? e0 : e1), lo ? acc[idx][0] : acc[idx + 1][0], lo ? acc[idx][1] : acc[idx + 1][1],
                             lo ? acc[idx + 1][0] : acc[idx][0], lo ? acc[idx + 1][1] : acc[idx][1]);
                } else {
                    red_add2(tab + e0, acc[idx][0], acc[idx][1]);
                    red_add2(tab + e1, acc[idx + 1][0], acc[idx + 1][1]);
                }
            }
        } else {
#pragma unroll
            for (int idx = 0; idx < (1 << D); ++idx) {
                if (C == 1) {
                    red_add(tab + e[idx], acc[idx][0]);
                } else {
#pragma unroll
                    for (int ch = 0; ch < C; ch += 2) red_add2(tab + e[idx] + ch, acc[idx][ch], acc[idx][ch + (C > 1 ? 1 : 0)]);
                }
            }
        }
    };

    for (uint32_t b = (uint32_t)b0; b < b1; ++b) {
        float x[D];
        if (load_pos<D>(inputs, b, x)) continue;         // out-of-range samples contribute nothing
        float fr[D];
        uint32_t pg[D];
        bool same = have;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float p = __fmaf_rn(x[d], g.scale, 0.5f);
            const float fl = floorf(p);
            pg[d] = (uint32_t)fl;
            fr[d] = p - fl;
            same = same && (pg[d] == cell[d]);
        }
        if (!same) {
            if (have) flush();
#pragma unroll
            for (int d = 0; d < D; ++d) cell[d] = pg[d];
#pragma unroll
            for (int idx = 0; idx < (1 << D); ++idx)
#pragma unroll
                for (int ch = 0; ch < C; ++ch) acc[idx][ch] = 0.f;
            have = true;
        }
        float gr[C];
        const T* __restrict__ gp = BLC ? grad + ((size_t)b * L + level) * C : grad + ((size_t)level * B + b) * C;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) gr[ch] = Elem<T>::to_f(gp[ch]);
        // corner weights in the reference's multiplication order: ((w0 * w1) * w2), partial products shared
        float wpart[1 << (D - 1)];
#pragma unroll
        for (int q = 0; q < (1 << (D - 1)); ++q) {
            float w = 1.0f;
#pragma unroll
            for (int d = 0; d < D - 1; ++d) w *= ((q >> d) & 1) ? fr[d] : 1.0f - fr[d];
            wpart[q] = w;
        }
#pragma unroll
        for (int idx = 0; idx < (1 << D); ++idx) {
            const float w = wpart[idx & ((1 << (D - 1)) - 1)] * ((idx >> (D - 1)) ? fr[D - 1] : 1.0f - fr[D - 1]);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) acc[idx][ch] = __fmaf_rn(w, gr[ch], acc[idx][ch]);
        }
    }
    if (have) flush();
}

// gridencoder.cu:314-340
template <typename T, int D, int C, bool BLC>
__global__ void k_grid_input_bwd(const T* __restrict__ grad, const T* __restrict__ dy_dx, T* __restrict__ grad_inputs,
                                 uint32_t B, uint32_t L) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const T* __restrict__ dd = dy_dx + (size_t)b * L * D * C;
    T result = Elem<T>::from_f(0.f);
    for (uint32_t l = 0; l < L; ++l) {
        for (int ch = 0; ch < C; ++ch) {
            const T gv = BLC ? grad[((size_t)b * L + l) * C + ch] : grad[((size_t)l * B + b) * C + ch];
            const float prod = Elem<T>::to_f(gv) * Elem<T>::to_f(dd[(l * D + d) * C + ch]);
            // `result += a * b` in scalar_t: product and sum each rounded to T (fp32: contracted)
            if (sizeof(T) == 4) result = Elem<T>::from_f(__fmaf_rn(Elem<T>::to_f(gv), Elem<T>::to_f(dd[(l * D + d) * C + ch]), Elem<T>::to_f(result)));
            else result = Elem<T>::from_f(Elem<T>::to_f(result) + Elem<T>::to_f(Elem<T>::from_f(prod)));
        }
    }
    grad_inputs[t] = result;
}

// ------------------------------------------------------------------------------------------
// launch plumbing
// ------------------------------------------------------------------------------------------
static inline uint32_t stage_row_stride(uint32_t L, uint32_t C, size_t elt) {
    uint32_t row = L * C;
    // make the row an odd number of 32-bit words so lanes (rows) fall in distinct banks
    if ((row * elt) % 4 == 0) {
        uint32_t words = row * elt / 4;
        if (words % 2 == 0) row += 4 / elt;
    }
    return row;
}

template <typename T, int D, int C>
static int launch_fwd(const Inputs inputs, const T* emb, const int32_t* offsets, T* outputs, uint32_t B, uint32_t L, float S,
                      uint32_t H, bool cg, T* dy_dx, uint32_t gridtype, int out_layout, cudaStream_t st) {
    const dim3 block(32, min(L, 16u));
    const dim3 grid(ceil_div(B, (uint32_t)kSamplesPerCta));
    const bool fast = (D == 3) && !cg && g_fwd_fast;
    constexpr int kWpl = (C * (int)sizeof(T)) / 4;
    if (fast && g_fwd_fast == 1 && kWpl >= 1 && L * kWpl <= 64) {
        // v2: warp walks all levels of its 32 samples (persistent CTAs of 8 warps)
        const uint32_t n_groups = ceil_div(B, 32u);
        const uint32_t row_words = (L * kWpl) | 1u;
        const size_t smem = (out_layout == 1) ? (size_t)8 * 32 * row_words * 4 : 0;
        // persistent grid = exactly the CTAs that are resident at once (a partial second wave would idle most SMs at the end).
        // (Forcing 6 CTAs/SM with __launch_bounds__(256, 6) — 40 registers, 28 bytes of spills — measured slower: 0.418 vs 0.397 ms.)
        constexpr int CCc = (kWpl >= 1 ? C : 2);
        auto launch = [&](auto kern) -> int {
            int per_sm = 0;
            ENERF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem), "grid_encode_forward");
            if (per_sm < 1) per_sm = 1;
            const uint32_t ctas = min(ceil_div(n_groups, 8u), (uint32_t)num_sms() * (uint32_t)per_sm);
            kern<<<ctas, 256, smem, st>>>(inputs, emb, offsets, outputs, B, L, S, H, gridtype, n_groups, row_words);
            return 0;
        };
        // (A variant fetching the two x-neighbour corners with one 8-byte load where they share an aligned block — 6 instead of 8
        // sector accesses per sample-level — was bit-identical but slower on B200, 0.465 vs 0.399 ms: removed, profiles/r2_01.)
        const int rc = (out_layout == 1) ? launch(k_grid_fwd_w<T, CCc, true>) : launch(k_grid_fwd_w<T, CCc, false>);
        if (rc) return rc;
        ENERF_CHECK_LAUNCH("grid_encode_forward");
        return 0;
    }
    if (out_layout == 1) {
        const uint32_t rs = stage_row_stride(L, C, sizeof(T));
        const size_t smem = (size_t)kSamplesPerCta * rs * sizeof(T);
        if (smem > 40 * 1024) { set_error("grid_encode_forward: L*C too large for the staging tile"); return -2; }
        if (fast) k_grid_fwd3<T, C, true><<<grid, block, smem, st>>>(inputs, emb, offsets, outputs, B, L, S, H, gridtype, rs);
        else k_grid_fwd<T, D, C, true><<<grid, block, smem, st>>>(inputs, emb, offsets, outputs, B, L, S, H, cg, dy_dx, gridtype, rs);
    } else {
        if (fast) k_grid_fwd3<T, C, false><<<grid, block, 0, st>>>(inputs, emb, offsets, outputs, B, L, S, H, gridtype, 0);
        else k_grid_fwd<T, D, C, false><<<grid, block, 0, st>>>(inputs, emb, offsets, outputs, B, L, S, H, cg, dy_dx, gridtype, 0);
    }
    ENERF_CHECK_LAUNCH("grid_encode_forward");
    return 0;
}

template <typename T, typename G, int D, int C>
static int launch_bwd(const T* grad, const Inputs inputs, const int32_t* offsets, G* gg, uint32_t B, uint32_t L, float S, uint32_t H,
                      bool cg, const T* dy_dx, T* grad_inputs, uint32_t gridtype, int out_layout, cudaStream_t st) {
    if (g_bwd_walk && L <= 32 && (32 % L) == 0) {
        // samples per thread-run: every run boundary costs one extra flush (8 reductions) per level, longer runs mean fewer threads
        // (measured, 3.29 M samples: 16: 0.768, 32: 0.707, 64: 0.696, 128: 0.715 ms)
        {
            constexpr int SEG = 64;
            const uint64_t threads = (uint64_t)ceil_div(B, (uint32_t)SEG) * L;
            const uint32_t block = (uint32_t)g_bwd_block;
            const dim3 grid((uint32_t)ceil_div(threads, (uint64_t)block));
            if (out_layout == 1) k_grid_bwd_walk<T, G, D, C, true, SEG><<<grid, block, 0, st>>>(grad, inputs, offsets, gg, B, L, S, H, gridtype);
            else k_grid_bwd_walk<T, G, D, C, false, SEG><<<grid, block, 0, st>>>(grad, inputs, offsets, gg, B, L, S, H, gridtype);
        }
    } else {
        const dim3 block(32, min(L, 16u));
        const dim3 grid(ceil_div(B, (uint32_t)kSamplesPerCta));
        if (out_layout == 1) k_grid_bwd<T, G, D, C, true><<<grid, block, 0, st>>>(grad, inputs, offsets, gg, B, L, S, H, gridtype);
        else k_grid_bwd<T, G, D, C, false><<<grid, block, 0, st>>>(grad, inputs, offsets, gg, B, L, S, H, gridtype);
    }
    ENERF_CHECK_LAUNCH("grid_encode_backward");
    if (cg) {
        const uint32_t n = B * D;
        if (out_layout == 1) k_grid_input_bwd<T, D, C, true><<<ceil_div(n, 256u), 256, 0, st>>>(grad, dy_dx, grad_inputs, B, L);
        else k_grid_input_bwd<T, D, C, false><<<ceil_div(n, 256u), 256, 0, st>>>(grad, dy_dx, grad_inputs, B, L);
        ENERF_CHECK_LAUNCH("grid_encode_backward(input)");
    }
    return 0;
}

#define ENERF_DC_SWITCH(D, C, name, CALL)                                                           \
    if (D == 3) {                                                                                   \
        switch (C) {                                                                                \
            case 1: { constexpr int DD = 3, CC = 1; CALL; } break;                                  \
            case 2: { constexpr int DD = 3, CC = 2; CALL; } break;                                  \
            case 4: { constexpr int DD = 3, CC = 4; CALL; } break;                                  \
            case 8: { constexpr int DD = 3, CC = 8; CALL; } break;                                  \
            default: set_error("%s: GridEncoding: C must be 1, 2, 4, or 8.", name); return -2;      \
        }                                                                                           \
    } else if (D == 2) {                                                                            \
        switch (C) {                                                                                \
            case 1: { constexpr int DD = 2, CC = 1; CALL; } break;                                  \
            case 2: { constexpr int DD = 2, CC = 2; CALL; } break;                                  \
            case 4: { constexpr int DD = 2, CC = 4; CALL; } break;                                  \
            case 8: { constexpr int DD = 2, CC = 8; CALL; } break;                                  \
            default: set_error("%s: GridEncoding: C must be 1, 2, 4, or 8.", name); return -2;      \
        }                                                                                           \
    } else {                                                                                        \
        set_error("%s: GridEncoding: D must be 2 or 3.", name);                                     \
        return -2;                                                                                  \
    }

}  // namespace enerf

using namespace enerf;

extern "C" {

int enerf_grid_set_forward_mode(int mode) {
    ENERF_REQUIRE(mode >= 0 && mode <= 2, "grid_set_forward_mode",
                  "mode must be 0 (generic kernel), 1 (warp-walks-levels D=3 kernel) or 2 (per-level D=3 kernel)");
    g_fwd_fast = mode;
    return 0;
}

int enerf_grid_set_backward_block(int threads) {
    ENERF_REQUIRE(threads == 0 || threads == 64 || threads == 128 || threads == 192 || threads == 256, "grid_set_backward_block",
                  "threads must be 0 (default), 64, 128, 192 or 256");
    g_bwd_block = threads ? threads : kBwdBlockDefault;
    return 0;
}

int enerf_grid_set_backward_mode(int mode) {
    ENERF_REQUIRE(mode == 0 || mode == 1, "grid_set_backward_mode", "mode must be 0 (per-corner reductions) or 1 (walking aggregation)");
    g_bwd_walk = mode;
    return 0;
}

int enerf_grid_encode_forward_xf(const float* raw_inputs, float in_add, float in_mul, const void* embeddings, const int32_t* offsets, void* outputs,
                                 uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs,
                                 void* dy_dx, uint32_t gridtype, int dtype, int out_layout, void* stream) {
    if (B == 0) return 0;
    ENERF_REQUIRE(in_mul == 0.0f || !calc_grad_inputs, "grid_encode_forward", "dy_dx is defined for inputs in [0, 1] only");
    const Inputs inputs = {raw_inputs, in_add, in_mul};
    ENERF_REQUIRE(dtype == ENERF_F32 || dtype == ENERF_F16, "grid_encode_forward", "dtype must be ENERF_F32 or ENERF_F16");
    ENERF_REQUIRE(out_layout == 0 || out_layout == 1, "grid_encode_forward", "out_layout must be 0 or 1");
    ENERF_REQUIRE(L >= 1 && L <= 64, "grid_encode_forward", "L must be in [1,64]");
    cudaStream_t st = as_stream(stream);
    int rc = 0;
    if (dtype == ENERF_F16) {
        ENERF_DC_SWITCH(D, C, "grid_encode_forward",
                        rc = (launch_fwd<__half, DD, CC>(inputs, (const __half*)embeddings, offsets, (__half*)outputs, B, L, S, H,
                                                         calc_grad_inputs != 0, (__half*)dy_dx, gridtype, out_layout, st)));
    } else {
        ENERF_DC_SWITCH(D, C, "grid_encode_forward",
                        rc = (launch_fwd<float, DD, CC>(inputs, (const float*)embeddings, offsets, (float*)outputs, B, L, S, H,
                                                        calc_grad_inputs != 0, (float*)dy_dx, gridtype, out_layout, st)));
    }
    return rc;
}

int enerf_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs,
                              uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs,
                              void* dy_dx, uint32_t gridtype, int dtype, int out_layout, void* stream) {
    return enerf_grid_encode_forward_xf(inputs, 0.0f, 0.0f, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx, gridtype, dtype,
                                        out_layout, stream);
}

int enerf_grid_encode_backward_xf(const void* grad, const float* raw_inputs, float in_add, float in_mul, const void* embeddings, const int32_t* offsets,
                                  void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                  int calc_grad_inputs, const void* dy_dx, void* grad_inputs, uint32_t gridtype, int dtype,
                                  int grad_dtype, int out_layout, void* stream) {
    (void)embeddings;
    if (B == 0) return 0;
    ENERF_REQUIRE(in_mul == 0.0f || !calc_grad_inputs, "grid_encode_backward", "input gradients are defined for inputs in [0, 1] only");
    const Inputs inputs = {raw_inputs, in_add, in_mul};
    ENERF_REQUIRE(dtype == ENERF_F32 || dtype == ENERF_F16, "grid_encode_backward", "dtype must be ENERF_F32 or ENERF_F16");
    ENERF_REQUIRE(grad_dtype == ENERF_F32 || grad_dtype == dtype, "grid_encode_backward", "grad_dtype must be ENERF_F32 or equal dtype");
    ENERF_REQUIRE(out_layout == 0 || out_layout == 1, "grid_encode_backward", "out_layout must be 0 or 1");
    ENERF_REQUIRE(L >= 1 && L <= 64, "grid_encode_backward", "L must be in [1,64]");
    cudaStream_t st = as_stream(stream);
    const bool cg = calc_grad_inputs != 0;
    int rc = 0;
    if (dtype == ENERF_F16 && grad_dtype == ENERF_F16) {
        ENERF_DC_SWITCH(D, C, "grid_encode_backward",
                        rc = (launch_bwd<__half, __half, DD, CC>((const __half*)grad, inputs, offsets, (__half*)grad_embeddings, B, L, S, H,
                                                                 cg, (const __half*)dy_dx, (__half*)grad_inputs, gridtype, out_layout, st)));
    } else if (dtype == ENERF_F16) {
        ENERF_DC_SWITCH(D, C, "grid_encode_backward",
                        rc = (launch_bwd<__half, float, DD, CC>((const __half*)grad, inputs, offsets, (float*)grad_embeddings, B, L, S, H,
                                                                cg, (const __half*)dy_dx, (__half*)grad_inputs, gridtype, out_layout, st)));
    } else {
        ENERF_DC_SWITCH(D, C, "grid_encode_backward",
                        rc = (launch_bwd<float, float, DD, CC>((const float*)grad, inputs, offsets, (float*)grad_embeddings, B, L, S, H,
                                                               cg, (const float*)dy_dx, (float*)grad_inputs, gridtype, out_layout, st)));
    }
    return rc;
}
int enerf_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings, const int32_t* offsets,
                               void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                               int calc_grad_inputs, const void* dy_dx, void* grad_inputs, uint32_t gridtype, int dtype,
                               int grad_dtype, int out_layout, void* stream) {
    return enerf_grid_encode_backward_xf(grad, inputs, 0.0f, 0.0f, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, calc_grad_inputs, dy_dx,
                                         grad_inputs, gridtype, dtype, grad_dtype, out_layout, stream);
}

}  // extern "C"
