// The E-NeRF field under torch.no_grad as ONE kernel: hash-grid gather -> sigma-net -> trunc_exp / SH / colour-net input ->
// colour-net -> sigmoid (nerf/network_ff.py:51-73; what NeRFRenderer.run_cuda's inference loop, renderer.py:364-391, evaluates
// once per marching round).  The unfused chain (grid_encode_forward, field_sigma_forward, field_color_forward) writes the
// [S,32] fp16 features and the [S,32] fp16 colour-net inputs to HBM and reads them back — 256 B per sample, 420 MB per
// 3.29 M samples — and runs the L1-bound gather and the tensor-core MLPs one after the other.  Here
//   * gather warps (teams of four, one lane per sample, all 16 levels, software-pipelined by pairs of levels: 16-32 independent
//     4-byte loads in flight per thread) write the feature tile of 128 samples straight into shared memory in the 64-byte TMA swizzle, i.e. as the layer-0 A
//     operand of the sigma-net (K-major), and hand it to the MMA warp through an mbarrier (ring of NX tiles);
//   * the MMA warp and the 4 epilogue warps per slot are those of k_tc_fwd_tma (ffmlp_tc.cu) with a seven-stage schedule per tile:
//     three sigma-net matmuls, then four colour-net matmuls whose first A operand is the colour-net input tile the sigma head
//     leaves in shared memory (same swizzle) — neither tile ever exists in HBM;
//   * what reaches HBM: 12 + 12 B in (position, direction), 4 + 4*n_ch B out (sigma, rgb) per sample.
// Values are the unfused chain's to the bit: the gather is grid_levels.cuh's (same corner order and fp16 rounding), the MMAs, the
// K order and the rounding points are k_tc_fwd_tma's.  L = 16 levels of 2 fp16 features, D = 3, FFMLP 32-64-64-16 and
// 32-64-64-64-16 only (E-NeRF's field); everything else takes the unfused chain.
#include "tc_common.cuh"
#include "grid_levels.cuh"

namespace enerf {
namespace fi {

using namespace tc;

static constexpr int kW = 64;              // hidden width
static constexpr int kTile = 128;          // samples per tile = UMMA M
static constexpr int kSlotCols = 96;       // TMEM columns per slot: D (64, fp32) + A (32 = 64 fp16)
static constexpr int kLevels = 16;
static constexpr int kStages = 7;          // matmuls per tile: sigma-net 3 + colour-net 4
static constexpr uint32_t kXBytes = kTile * 32 * 2;     // one [128 x 32] fp16 tile (64-byte rows)
static constexpr uint32_t kWBytes = 4096 + 8192 + 2048 + 4096 + 2 * 8192 + 2048;      // both nets, canonical layout

// ---- one sample's 16 levels -> its 64-byte feature row in a 64-byte-swizzled tile (the layer-0 A operand) ----
// The eight corner values of one (sample, level) between their loads and the blend: 8 registers (fp16 pairs as loaded).  The
// interpolation weights are recomputed from the position at blend time — nine cheap instructions instead of six live registers per
// level in flight, which is what lets a thread keep three pairs of levels (48 loads) outstanding.  Same operations in the same order
// as LevelWork (grid_levels.cuh): bit-identical features.
struct LevelLite {
    uint32_t v[8];
    template <int MODE>
    __device__ __forceinline__ void fetch(const LevelTabW& lt, const __half* __restrict__ grid, const float (&x)[3]) {
        const __half* __restrict__ tab = grid + (size_t)lt.offset * 2;
        uint32_t pg[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) pg[d] = (uint32_t)floorf(__fmaf_rn(x[d], lt.scale, 0.5f));
        uint32_t e[8];
        corner_index<2, MODE>(lt, pg, e);
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) v[idx] = __ldg(reinterpret_cast<const uint32_t*>(tab + e[idx]));
    }
    __device__ __forceinline__ uint32_t blend(const LevelTabW& lt, const float (&x)[3], bool zero) const {
        float fr[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float p = __fmaf_rn(x[d], lt.scale, 0.5f);
            fr[d] = p - floorf(p);
        }
        const float wx0 = 1.0f - fr[0], wy0 = 1.0f - fr[1];
        const float wxy[4] = {wx0 * wy0, fr[0] * wy0, wx0 * fr[1], fr[0] * fr[1]};
        const float wz[2] = {1.0f - fr[2], fr[2]};
        __half2 res2 = __floats2half2_rn(0.f, 0.f);
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) {
            const float w = wxy[idx & 3] * wz[idx >> 2];
            const float2 g = __half22float2(*reinterpret_cast<const __half2*>(&v[idx]));
            res2 = __hadd2(res2, __floats2half2_rn(w * g.x, w * g.y));
        }
        return zero ? 0u : *reinterpret_cast<const uint32_t*>(&res2);
    }
};

// Levels [0, ND) use the dense index form, the others the hashed power-of-two form (compile-time per unrolled level).  PD pairs of
// levels are live at any time: the 16 corner loads of pair p+PD-1 are issued before pair p is blended, so a warp keeps up to 16*PD
// independent 4-byte gathers in flight while it computes — the dedicated gather kernel gets that overlap from 48 resident warps per
// SM, here there are 16-20.  Four levels (two pairs) make one 16-byte chunk of the row.
template <int ND, int PD>
__device__ __forceinline__ void gather_row(const LevelTabW* ltab, const __half* __restrict__ grid, const float (&x)[3], bool oob, uint8_t* xb, uint32_t r) {
    LevelLite w[PD][2];
    auto fetch_pair = [&](int p) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int level = 2 * p + i;
            if (level < ND) w[p % PD][i].template fetch<1>(ltab[level], grid, x);
            else w[p % PD][i].template fetch<0>(ltab[level], grid, x);
        }
    };
#pragma unroll
    for (int p = 0; p < PD - 1; ++p) fetch_pair(p);
    uint32_t o[4];
#pragma unroll
    for (int p = 0; p < kLevels / 2; ++p) {
        if (p + PD - 1 < kLevels / 2) fetch_pair(p + PD - 1);
#pragma unroll
        for (int i = 0; i < 2; ++i) o[(2 * p + i) & 3] = w[p % PD][i].blend(ltab[2 * p + i], x, oob);
        if (p & 1) *reinterpret_cast<int4*>(xb + sw_off(r, (uint32_t)(p >> 1), 64)) = make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]);
    }
}
// any level table: four levels (one chunk) at a time, the index form chosen per level
__device__ __forceinline__ void gather_row_generic(const LevelTabW* ltab, const __half* __restrict__ grid, const float (&x)[3], bool oob, uint8_t* xb, uint32_t r) {
#pragma unroll 1
    for (uint32_t c = 0; c < 4; ++c) {
        LevelWork<__half, 2> w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const LevelTabW& lt = ltab[4 * c + i];
            if (lt.mode == 0u) w[i].template fetch<0>(lt, grid, x);
            else if (lt.mode == 1u) w[i].template fetch<1>(lt, grid, x);
            else w[i].template fetch<2>(lt, grid, x);
        }
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t ow[1];
            w[i].blend(ow, oob);
            o[i] = ow[0];
        }
        *reinterpret_cast<int4*>(xb + sw_off(r, c, 64)) = make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]);
    }
}

// NSLOTS tiles in flight in the MLP part, NGT gather teams of four warps, NX feature tiles in the ring (a multiple of NGT, so a ring
// buffer is always filled by the same team and the parity of its "empty" barrier cannot be overrun)
template <int NSLOTS, int NGT, int NX, int PD>
__global__ void __launch_bounds__(32 + NSLOTS * 128 + NGT * 128, 1)
k_field_infer(const Inputs inputs, const float* __restrict__ dirs, const __half* __restrict__ grid, const int32_t* __restrict__ offsets, float S,
              uint32_t H, uint32_t gridtype, const __half* __restrict__ Ws, const __half* __restrict__ Wc, float* __restrict__ sigma,
              float* __restrict__ rgb, int n_ch, uint32_t B, uint32_t n_tiles, const int32_t* __restrict__ n_units_dev, uint32_t rows_per_unit) {
    static_assert(NX % NGT == 0, "ring buffers per team");
    if (n_units_dev) {
        // only the first *n_units_dev * rows_per_unit rows are live (the alive rays of a marching round times its steps; the host sized
        // the launch for an older, larger count): every thread reads the same word, so the CTA agrees on its number of tiles
        const int32_t u = __ldg(n_units_dev);
        const uint64_t rows = u > 0 ? (uint64_t)u * rows_per_unit : 0;
        if (rows < (uint64_t)B) {
            B = (uint32_t)rows;
            n_tiles = (B + (uint32_t)kTile - 1u) / (uint32_t)kTile;
        }
    }
    static_assert(NSLOTS * kSlotCols <= 512, "TMEM budget");
    constexpr bool kBothHalves = (1 + 4 * NSLOTS + 4 * NGT) <= 24;      // 80 registers per thread up to 24 warps, 72 beyond
    extern __shared__ uint8_t smem_dyn[];
    __shared__ LevelTabW ltab[64];
    __shared__ uint32_t s_plan[2];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint8_t* xring = smem;                                      // NX feature tiles, 64-byte swizzle
    uint8_t* cst = xring + (size_t)NX * kXBytes;                // NSLOTS colour-net input tiles, 64-byte swizzle
    uint8_t* ws0 = cst + (size_t)NSLOTS * kXBytes;              // sigma-net: [4][64][16 B], [8][64][16 B], [8][16][16 B]
    uint8_t* wsh = ws0 + 4096;
    uint8_t* wsl = wsh + 8192;
    uint8_t* wc0 = wsl + 2048;                                  // colour-net: [4][64][16 B], 2 x [8][64][16 B], [8][16][16 B]
    uint8_t* wch = wc0 + 4096;
    uint8_t* wcl = wch + 2 * 8192;
    uint64_t* a_ready = reinterpret_cast<uint64_t*>(wcl + 2048);
    uint64_t* d_full = a_ready + NSLOTS;
    uint64_t* x_full = d_full + NSLOTS;
    uint64_t* x_empty = x_full + NX;
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(x_empty + NX);

    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t kCols = (NSLOTS * kSlotCols <= 256) ? 256 : 512;

    stage_matrix(ws0, Ws, kW, 32, tid, nthreads);
    stage_matrix(wsh, Ws + kW * 32, kW, kW, tid, nthreads);
    stage_matrix(wsl, Ws + kW * 32 + kW * kW, 16, kW, tid, nthreads);
    stage_matrix(wc0, Wc, kW, 32, tid, nthreads);
    stage_matrix(wch, Wc + kW * 32, kW, kW, tid, nthreads);
    stage_matrix(wch + 8192, Wc + kW * 32 + kW * kW, kW, kW, tid, nthreads);
    stage_matrix(wcl, Wc + kW * 32 + 2 * kW * kW, 16, kW, tid, nthreads);
    if (tid == 0) {
        for (int s = 0; s < NSLOTS; ++s) {
            mbar_init(&a_ready[s], 4);      // one arrival per epilogue warp
            mbar_init(&d_full[s], 1);       // tcgen05.commit
        }
        for (int b = 0; b < NX; ++b) {
            mbar_init(&x_full[b], 4);       // one arrival per gather warp of the team
            mbar_init(&x_empty[b], 1);      // tcgen05.commit after the layer-0 MMAs that read the tile
        }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_base_ptr, kCols);
    build_level_table(ltab, s_plan, offsets, kLevels, S, H, gridtype, (uint32_t)tid);      // two __syncthreads inside
    fence_proxy_async_smem();      // weights written with st.shared are read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *tmem_base_ptr;
    const uint32_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t plan_nd = s_plan[1] ? 0xffffffffu : s_plan[0];      // leading dense levels, or "needs the generic index form"

    if (warp == 0) {
        // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem0, 0);
        const uint32_t xr_b = smem_u32(xring), cs_b = smem_u32(cst);
        const uint32_t ws0b = smem_u32(ws0), wshb = smem_u32(wsh), wslb = smem_u32(wsl), wc0b = smem_u32(wc0), wchb = smem_u32(wch), wclb = smem_u32(wcl);
        uint32_t nt[NSLOTS];
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) nt[s] = (my_tiles > (uint32_t)s) ? (my_tiles - s + NSLOTS - 1) / NSLOTS : 0;
        constexpr uint32_t idesc64 = idesc_f16(kTile, 64, false, false), idesc16 = idesc_f16(kTile, 16, false, false);
        for (uint32_t tl = 0; tl < nt[0]; ++tl) {
#pragma unroll
            for (int t = 0; t < kStages; ++t) {
#pragma unroll
                for (int s = 0; s < NSLOTS; ++s) {
                    if (tl >= nt[s]) continue;
                    const uint32_t d_t = tm + s * kSlotCols, a_t = d_t + 64;
                    const uint32_t ph = tl * kStages;                   // first phase of this tile on a_ready[s] / d_full[s]
                    if (t == 0) {
                        // ---- sigma-net layer 0: A = the gathered feature tile
                        const uint32_t j = tl * NSLOTS + (uint32_t)s, xbuf = j % NX, use = j / NX;
                        if (tl > 0) mbar_wait_sleep(&a_ready[s], (ph - 1u) & 1u);      // the previous tile's last accumulator has been read
                        mbar_wait_sleep(&x_full[xbuf], use & 1u);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t xb = xr_b + xbuf * kXBytes;
#pragma unroll
                            for (int k = 0; k < 2; ++k) mma_ss(d_t, smem_desc_sw(xb + k * 32, 64), smem_desc(ws0b + k * 2 * (kW * 16), kW * 16, 128), idesc64, k > 0);
                            tc_commit(&d_full[s]);
                            tc_commit(&x_empty[xbuf]);                  // the same MMAs have finished reading the ring buffer
                        }
                        __syncwarp();
                    } else {
                        mbar_wait_sleep(&a_ready[s], (ph + (uint32_t)(t - 1)) & 1u);
                        tc_fence_after();
                        if (elect_one()) {
                            if (t == 3) {
                                // ---- colour-net layer 0: A = the colour-net input tile the sigma head left in shared memory
                                const uint32_t cb = cs_b + (uint32_t)s * kXBytes;
#pragma unroll
                                for (int k = 0; k < 2; ++k) mma_ss(d_t, smem_desc_sw(cb + k * 32, 64), smem_desc(wc0b + k * 2 * (kW * 16), kW * 16, 128), idesc64, k > 0);
                            } else if (t == 2 || t == 6) {
                                const uint32_t wb = (t == 2) ? wslb : wclb;
#pragma unroll
                                for (int k = 0; k < 4; ++k) mma_ts(d_t, a_t + k * 8, smem_desc(wb + k * 2 * (16 * 16), 16 * 16, 128), idesc16, k > 0);
                            } else {
                                const uint32_t wb = (t == 1) ? wshb : wchb + (uint32_t)(t - 4) * 8192u;
#pragma unroll
                                for (int k = 0; k < 4; ++k) mma_ts(d_t, a_t + k * 8, smem_desc(wb + k * 2 * (kW * 16), kW * 16, 128), idesc64, k > 0);
                            }
                            tc_commit(&d_full[s]);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp <= 4 * NSLOTS) {
        // ===================== epilogue warps (4 per slot) =====================
        const int s = (warp - 1) >> 2;                 // slot
        const int q = warp & 3;                        // TMEM quarter this warp may access
        const int r = q * 32 + lane;                   // row of the tile
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const uint32_t d_t = tmem0 + lane_sel + s * kSlotCols, a_t = d_t + 64;
        uint8_t* cb = cst + (size_t)s * kXBytes;
        uint32_t tl = 0;
        for (uint32_t j = s; j < my_tiles; j += NSLOTS, ++tl) {
            const size_t row = ((size_t)blockIdx.x + (size_t)j * gridDim.x) * kTile + r;
            const bool valid = row < (size_t)B;
            float dx = 0.f, dy = 0.f, dz = 0.f;        // requested now, used after the sigma-net
            if (valid) {
                dx = __ldg(dirs + row * 3);
                dy = __ldg(dirs + row * 3 + 1);
                dz = __ldg(dirs + row * 3 + 2);
            }
            const uint32_t ph = tl * kStages;
#pragma unroll
            for (int k = 0; k < kStages; ++k) {
                mbar_wait_sleep(&d_full[s], (ph + (uint32_t)k) & 1u);
                tc_fence_after();
                if (k == 2) {
                    // ---- sigma head: sigma = exp(fp16(y[0])) (trunc_exp), colour-net input row [SH_4(fp16(dir)) | y[1:16] | 0] -> shared memory
                    uint32_t acc[16];
                    tmem_ld16(d_t, acc);
                    tc_wait_ld();
                    if (valid) sigma[row] = expf(f16_round(__uint_as_float(acc[0])));
                    float sh[16];
                    sh_deg4(f16_round(dx), f16_round(dy), f16_round(dz), sh);      // directions reach the SH encoder as fp16 (sphere_harmonics.py:16)
                    uint32_t p[16];
#pragma unroll
                    for (int e = 0; e < 8; ++e) p[e] = pack2(sh[2 * e], sh[2 * e + 1]);
#pragma unroll
                    for (int e = 0; e < 7; ++e) p[8 + e] = pack2(__uint_as_float(acc[1 + 2 * e]), __uint_as_float(acc[2 + 2 * e]));
                    p[15] = pack2(__uint_as_float(acc[15]), 0.0f);
                    // the previous tile's colour-net layer 0 (the last reader of this buffer) completed before its epilogue ran
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        *reinterpret_cast<int4*>(cb + sw_off((uint32_t)r, (uint32_t)v, 64)) =
                            make_int4((int)p[4 * v], (int)p[4 * v + 1], (int)p[4 * v + 2], (int)p[4 * v + 3]);
                    fence_proxy_async_smem();
                    tc_fence_before();
                    warp_arrive(&a_ready[s], lane);
                } else if (k == kStages - 1) {
                    // ---- colour head: rgb = fp16(sigmoid(fp16(y[c])))
                    uint32_t acc[16];
                    tmem_ld16(d_t, acc);
                    tc_wait_ld();
                    tc_fence_before();
                    warp_arrive(&a_ready[s], lane);              // accumulator read: the slot can start its next tile
                    if (valid) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (c < n_ch) {
                                const float y = f16_round(__uint_as_float(acc[c]));
                                rgb[row * n_ch + c] = f16_round(1.0f / (1.0f + expf(-y)));
                            }
                    }
                } else {
                    // ---- hidden layer: relu -> fp16 -> TMEM (A operand of the next matmul)
                    if (kBothHalves) {
                        uint32_t acc2[2][32];                    // both halves requested before the first is used (one TMEM round trip)
                        tmem_ld32(d_t, acc2[0]);
                        tmem_ld32(d_t + 32, acc2[1]);
                        tc_wait_ld();
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint32_t p[16];
#pragma unroll
                            for (int e = 0; e < 16; ++e) p[e] = pack2_relu(__uint_as_float(acc2[h][2 * e]), __uint_as_float(acc2[h][2 * e + 1]));
                            tmem_st16(a_t + h * 16, p);
                        }
                    } else {
                        // 64 registers per thread with 29 warps: the row goes through in four 16-column pieces (the kernel is bound by the
                        // gather, not by this chain)
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            uint32_t acc[16];
                            tmem_ld16(d_t + h * 16, acc);
                            tc_wait_ld();
                            uint32_t p[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) p[e] = pack2_relu(__uint_as_float(acc[2 * e]), __uint_as_float(acc[2 * e + 1]));
                            tmem_st8(a_t + h * 8, p);
                        }
                    }
                    tc_wait_st();
                    tc_fence_before();
                    warp_arrive(&a_ready[s], lane);
                }
            }
        }
    } else {
        // ===================== gather warps (teams of 4: one tile of 128 samples per team and pass) =====================
        const int gw = warp - (1 + 4 * NSLOTS);
        const int team = gw >> 2;
        const uint32_t r = (uint32_t)((gw & 3) * 32 + lane);   // row of the tile
        for (uint32_t j = team; j < my_tiles; j += NGT) {
            const uint32_t xbuf = j % NX, use = j / NX;
            const size_t row = ((size_t)blockIdx.x + (size_t)j * gridDim.x) * kTile + r;
            float x[3] = {0.f, 0.f, 0.f};
            bool oob = true;
            if (row < (size_t)B) oob = load_pos<3>(inputs, (uint32_t)row, x);
            if (oob) x[0] = x[1] = x[2] = 0.f;         // keep the address arithmetic in range; the result is zeroed below
            if (use > 0) mbar_wait_sleep(&x_empty[xbuf], (use - 1u) & 1u);      // the layer-0 MMAs of the buffer's previous tile have read it
            uint8_t* xb = xring + (size_t)xbuf * kXBytes;
            // the usual table is "n_dense dense levels, then hashed power-of-two levels": the level loop is then fully unrolled with
            // compile-time index forms and software-pipelined by pairs of levels (below); anything else walks the levels with a
            // (warp-uniform) branch per level
            switch (plan_nd) {
                case 4: gather_row<4, PD>(ltab, grid, x, oob, xb, r); break;
                case 5: gather_row<5, PD>(ltab, grid, x, oob, xb, r); break;
                case 6: gather_row<6, PD>(ltab, grid, x, oob, xb, r); break;
                default: gather_row_generic(ltab, grid, x, oob, xb, r); break;
            }
            fence_proxy_async_smem();
            warp_arrive(&x_full[xbuf], lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, kCols);
}

template <int NSLOTS, int NGT, int NX, int PD>
static int launch(const Inputs& in, const float* dirs, const __half* grid, const int32_t* offsets, float S, uint32_t H, uint32_t gridtype, const __half* Ws,
                  const __half* Wc, float* sigma, float* rgb, int n_ch, uint32_t B, const int32_t* n_units_dev, uint32_t rows_per_unit, cudaStream_t st) {
    size_t smem = 1024 + (size_t)(NX + NSLOTS) * kXBytes + kWBytes + (size_t)(2 * NSLOTS + 2 * NX) * 8 + 16;
    if (smem < 120 * 1024) smem = 120 * 1024;      // one CTA per SM: it owns the SM's tensor memory
    static bool configured = false;
    if (!configured) {
        ENERF_CUDA(cudaFuncSetAttribute(k_field_infer<NSLOTS, NGT, NX, PD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "field_infer");
        configured = true;
    }
    const uint32_t n_tiles = ceil_div(B, (uint32_t)kTile);
    const uint32_t grid_x = n_tiles < (uint32_t)num_sms() ? n_tiles : (uint32_t)num_sms();
    k_field_infer<NSLOTS, NGT, NX, PD><<<grid_x, 32 + NSLOTS * 128 + NGT * 128, smem, st>>>(in, dirs, grid, offsets, S, H, gridtype, Ws, Wc, sigma, rgb, n_ch, B, n_tiles,
                                                                                            n_units_dev, rows_per_unit);
    ENERF_CHECK_LAUNCH("field_infer");
    return 0;
}

}  // namespace fi
}  // namespace enerf

using namespace enerf;

extern "C" int enerf_field_infer_alive(const float* raw_xyz, float in_add, float in_mul, const float* dirs, const uint16_t* embeddings, const int32_t* offsets,
                                       uint32_t L, uint32_t C, float S, uint32_t H, uint32_t gridtype, const uint16_t* w_sigma, uint32_t num_layers,
                                       const uint16_t* w_color, uint32_t num_layers_color, uint32_t B, uint32_t n_ch, float* sigma, float* rgb,
                                       const int32_t* n_units_dev, uint32_t rows_per_unit, void* stream) {
    ENERF_REQUIRE(L == (uint32_t)fi::kLevels && C == 2, "field_infer", "the fused field takes 16 levels of 2 fp16 features");
    ENERF_REQUIRE(num_layers == 2 && num_layers_color == 3, "field_infer", "the fused field takes the FFMLP 32-64-64-16 / 32-64-64-64-16 nets");
    ENERF_REQUIRE(n_ch >= 1 && n_ch <= 4, "field_infer", "n_ch must be in [1,4]");
    ENERF_REQUIRE(((reinterpret_cast<uintptr_t>(w_sigma) | reinterpret_cast<uintptr_t>(w_color)) & 15u) == 0 && (reinterpret_cast<uintptr_t>(embeddings) & 3u) == 0,
                  "field_infer", "weights must be 16-byte aligned, the table 4-byte aligned");
    if (B == 0) return 0;
    const Inputs in = {raw_xyz, in_add, in_mul};
    const __half* grid = reinterpret_cast<const __half*>(embeddings);
    const __half* Ws = reinterpret_cast<const __half*>(w_sigma);
    const __half* Wc = reinterpret_cast<const __half*>(w_color);
    cudaStream_t st = as_stream(stream);
    // Measured on B200 (3.29 M marcher-ordered samples; the unfused chain: 0.556 ms): 2 MLP slots + 4 gather teams, three pairs of levels
    // in flight per thread 0.511 ms; 3 slots + 4 teams 0.515 (two pairs: 0.517); 2 slots + 5 teams 0.528; 3 slots + 2 / 3 teams with
    // two pairs 0.608 / 0.564; without the software pipeline (four levels fetched, then blended) 0.72 / 0.63 / 0.55 with 2 / 3 / 4 teams
    // (profiles/r2_50_field_infer_probe.json).  ncu (r2_49): the gather warps are bound by their own instruction stream (146
    // instructions per sample-level, 4 warps per scheduler) — L1 data pipe 60 %, issue slots 56 %.
    return fi::launch<2, 4, 8, 3>(in, dirs, grid, offsets, S, H, gridtype, Ws, Wc, sigma, rgb, (int)n_ch, B, n_units_dev, rows_per_unit, st);
}

extern "C" int enerf_field_infer(const float* raw_xyz, float in_add, float in_mul, const float* dirs, const uint16_t* embeddings, const int32_t* offsets,
                                 uint32_t L, uint32_t C, float S, uint32_t H, uint32_t gridtype, const uint16_t* w_sigma, uint32_t num_layers,
                                 const uint16_t* w_color, uint32_t num_layers_color, uint32_t B, uint32_t n_ch, float* sigma, float* rgb, void* stream) {
    return enerf_field_infer_alive(raw_xyz, in_add, in_mul, dirs, embeddings, offsets, L, C, S, H, gridtype, w_sigma, num_layers, w_color, num_layers_color, B, n_ch,
                                   sigma, rgb, nullptr, 0, stream);
}
