// Fully-fused 64-wide MLP on the 5th-generation tensor cores (tcgen05) with TMEM-resident
// activations — the path E-NeRF's networks take: FFMLP sigma-net 32-64-64-16 and colour-net
// 32-64-64-64-16 (nerf/network_ff.py:31-49) and the torch-topology nets 32-64-16 and
// 31(+1)-64-64-C (nerf/network.py:40-77).  Other shapes use the mma.sync kernels in ffmlp.cu.
//
// All kernels: CTA = 1 issuing warp + NSLOTS x 4 epilogue warps, persistent over 128-sample
// tiles; NSLOTS tiles are in flight per CTA so the tensor pipe works on one tile while the
// epilogue warps of the others run.  Operand tiles arrive by TMA (cp.async.bulk.tensor) in the
// hardware swizzle; weights are staged once per CTA in shared memory in the no-swizzle canonical
// layout (tc_common.cuh); accumulators are fp32 in TMEM; between layers the activation goes
// tcgen05.ld -> ReLU -> fp16 -> tcgen05.st and never leaves the SM; hand-offs use mbarriers
// (one arrival per epilogue warp: "A ready"; tcgen05.commit: "D full").
#include "tc_common.cuh"
#include <cuda.h>

namespace enerf {
namespace tcm {

using namespace tc;

static constexpr int kW = 64;            // hidden width
static constexpr int kTile = 128;        // samples per tile = UMMA M
static constexpr int kSlotCols = 96;     // TMEM columns per slot: D (64, fp32) + A (32 = 64 fp16)

static constexpr int kGBytes = 128 * 64 * 2;    // one 128-sample x 64-wide fp16 tile

// Optional cycle trace (tools/build_trace.sh builds a second library with -DENERF_TC_TRACE; the product build compiles it out):
// CTA 0 appends (tag, clock64) pairs of slot 0's issuing warp and first epilogue thread to a global buffer.
#ifdef ENERF_TC_TRACE
__device__ unsigned long long* g_trace = nullptr;
// each tracing thread owns a region of the buffer and a private counter (no atomics: a stamp costs a store and a clock read)
#define ENERF_TRACE_DECL(region) unsigned int trace_n__ = 0; const unsigned int trace_base__ = (region) * 4096u
#define ENERF_TRACE(tag)                                                                        \
    do {                                                                                        \
        if (g_trace && blockIdx.x == 0 && trace_n__ < 2048u) {                                  \
            g_trace[2 * (trace_base__ + trace_n__)] = (unsigned long long)(tag);                \
            g_trace[2 * (trace_base__ + trace_n__) + 1] = clock64();                            \
        }                                                                                       \
        ++trace_n__;                                                                            \
    } while (0)
#else
#define ENERF_TRACE_DECL(region) do { } while (0)
#define ENERF_TRACE(tag) do { } while (0)
#endif

struct alignas(64) TmaDesc { uint8_t bytes[128]; };     // CUtensorMap (opaque in device code; encoded on the host)

// ---- host: tensor maps (driver entry point resolved through the runtime; no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// fp16 row-major [rows, cols] matrix, box = [box_rows, cols]; swizzle = the row size in bytes (32 / 64 / 128)
static bool make_tmap_rows(TmaDesc* out, const void* base, uint64_t rows, uint32_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || (reinterpret_cast<uintptr_t>(base) & 15u)) return false;
    const uint32_t row_bytes = cols * 2;
    const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    if (row_bytes != 128 && row_bytes != 64 && row_bytes != 32) return false;
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {row_bytes};
    const cuuint32_t box[2] = {cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "CUtensorMap size");
    return fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


// Optional fused heads of the last layer (the E-NeRF field, nerf/network_ff.py:51-73):
//   HEAD 1 (sigma-net): h = fp16(y); sigma = exp(h[0]) (fp32, `trunc_exp`); the colour-net input row
//           [SH_4(dir) (16) | h[1:16] (15) | 0] is written directly (no SH kernel, no cat, no zeros_like)
//   HEAD 2 (colour-net): rgb[c] = sigmoid(fp16(y[c])) for c < n_ch, written as fp32 [B, n_ch]
//   HEAD 3 (density only): sigma = exp(fp16(y[0])) and nothing else — the occupancy-grid refresh (nerf/renderer.py:510)
//   HEAD 4 (density + features, nerf/network.py:134-151): sigma as above plus the 16 raw outputs h (fp16 rows; geo_feat = h[1:16])
struct HeadArgs {
    const float* dirs;   // [B,3] fp32                    (HEAD 1)
    float* sigma;        // [B] fp32                      (HEAD 1)
    __half* cin;         // [B,32] fp16                   (HEAD 1)
    float* rgb;          // [B,n_ch] fp32                 (HEAD 2)
    int n_ch;
    const int32_t* n_rows_dev;   // optional: the number of valid rows lives on the device (B is then the capacity); tiles beyond it are skipped
};

// the tile count a kernel really processes when the row count is only known on the device (compacted batches, see field.py)
__device__ __forceinline__ uint32_t effective_tiles(uint32_t n_tiles, const int32_t* __restrict__ n_rows_dev) {
    if (!n_rows_dev) return n_tiles;
    const int32_t v = *n_rows_dev;
    return v <= 0 ? 0u : min(n_tiles, ((uint32_t)v + (uint32_t)kTile - 1u) / (uint32_t)kTile);
}

// ================================================================================================
// Forward / inference (k_tc_fwd_tma): 32 or 64 inputs, 0, 1 or 2 hidden-to-hidden matmuls.
//   * the [128 x in_dim] input tile arrives by cp.async.bulk.tensor (two-deep ring per slot) in the TMA swizzle and
//     is the layer-0 A operand straight from shared memory (K-major swizzled descriptor) — no per-thread loads,
//     no row transposition, no tcgen05.st for the inputs;
//   * training: every hidden activation tile is written to forward_buffer by one TMA store from a swizzled,
//     double-buffered staging tile (8 conflict-free STS.128 per thread instead of STS + LDS + STG);
//   * sigma-net head: the colour-net input tile leaves by TMA store as well;
//   * compile-time layer schedule (templated depth), warp-uniform issue, fixed slot order.
//   a_ready[s] counts, per tile, one phase per hidden epilogue ("A operand ready") plus one when the last
//   accumulator has been read ("slot free for the next tile"); d_full[s] one phase per layer.
// ================================================================================================
template <int NSLOTS, int NH, int IN_DIM, int HEAD, bool TRAIN>
__global__ void __launch_bounds__(32 + NSLOTS * 128, 1)
k_tc_fwd_tma(const __grid_constant__ TmaDesc tm_x, const __grid_constant__ TmaDesc tm_fb, const __grid_constant__ TmaDesc tm_cin,
             const __half* __restrict__ W, __half* __restrict__ out, uint32_t n_tiles, uint32_t B, HeadArgs head) {
    constexpr int S = NH + 2;                                   // matmuls per network
    constexpr uint32_t kXSw = IN_DIM * 2;                       // input row bytes = swizzle span (64 or 128)
    constexpr uint32_t kXBytes = kTile * IN_DIM * 2;
    constexpr uint32_t kCinBytes = kTile * 32 * 2;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint8_t* xring = smem;                                                          // NSLOTS x 2 input tiles
    uint8_t* stg = xring + (size_t)NSLOTS * 2 * kXBytes;                            // TRAIN: NSLOTS x 2 activation tiles (128-byte swizzle)
    uint8_t* cst = stg + (TRAIN ? (size_t)NSLOTS * 2 * kGBytes : 0);                // HEAD 1: NSLOTS colour-input tiles (64-byte swizzle)
    uint8_t* w0s = cst + (HEAD == 1 ? (size_t)NSLOTS * kCinBytes : 0);              // [in_dim/8][64][16 B]
    uint8_t* whs = w0s + IN_DIM * 128;                                              // NH x [8][64][16 B]
    uint8_t* wls = whs + NH * 8192;                                                 // [8][16][16 B]
    uint64_t* a_ready = reinterpret_cast<uint64_t*>(wls + 2048);
    uint64_t* d_full = a_ready + NSLOTS;
    uint64_t* x_full = d_full + NSLOTS;                                             // [NSLOTS][2]
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(x_full + 2 * NSLOTS);

    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t kCols = (NSLOTS * kSlotCols <= 256) ? 256 : 512;

    stage_matrix(w0s, W, kW, IN_DIM, tid, nthreads);
    for (int j = 0; j < NH; ++j) stage_matrix(whs + j * 8192, W + kW * IN_DIM + j * kW * kW, kW, kW, tid, nthreads);
    stage_matrix(wls, W + kW * IN_DIM + NH * kW * kW, 16, kW, tid, nthreads);
    if (tid == 0) {
        for (int s = 0; s < NSLOTS; ++s) {
            mbar_init(&a_ready[s], 4);      // one arrival per epilogue warp
            mbar_init(&d_full[s], 1);
            mbar_init(&x_full[2 * s], 1);
            mbar_init(&x_full[2 * s + 1], 1);
        }
        mbar_fence_init();
        tma_prefetch_desc(&tm_x);
        if (TRAIN) tma_prefetch_desc(&tm_fb);
        if (HEAD == 1) tma_prefetch_desc(&tm_cin);
    }
    if (warp == 0) tmem_alloc(tmem_base_ptr, kCols);
    fence_proxy_async_smem();      // weights written with st.shared are read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *tmem_base_ptr;
    n_tiles = effective_tiles(n_tiles, head.n_rows_dev);
    const uint32_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0) {
        // ===================== MMA / TMA-load issuer (warp-uniform, one elected lane issues) =====================
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem0, 0);
        const uint32_t xr_b = smem_u32(xring), w0b = smem_u32(w0s), whb = smem_u32(whs), wlb = smem_u32(wls);
        uint32_t nt[NSLOTS];
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) nt[s] = (my_tiles > (uint32_t)s) ? (my_tiles - s + NSLOTS - 1) / NSLOTS : 0;
        auto issue_x = [&](int s, uint32_t tl) {               // elected lane only
            const uint32_t tile = blockIdx.x + ((uint32_t)s + tl * NSLOTS) * gridDim.x;
            uint64_t* bar = &x_full[2 * s + (tl & 1u)];
            mbar_arrive_expect_tx(bar, kXBytes);
            tma_load_2d(xr_b + ((uint32_t)s * 2 + (tl & 1u)) * kXBytes, &tm_x, 0, (int32_t)(tile * kTile), bar);
        };
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) {
            if (elect_one()) {
                if (nt[s] > 0) issue_x(s, 0);
                if (nt[s] > 1) issue_x(s, 1);
            }
            __syncwarp();
        }
        constexpr uint32_t idesc64 = idesc_f16(kTile, 64, false, false), idesc16 = idesc_f16(kTile, 16, false, false);
        for (uint32_t tl = 0; tl < nt[0]; ++tl) {
#pragma unroll
            for (int L = 0; L < S; ++L) {
#pragma unroll
                for (int s = 0; s < NSLOTS; ++s) {
                    if (tl >= nt[s]) continue;
                    const uint32_t d_t = tm + s * kSlotCols, a_t = d_t + 64;
                    if (L == 0) {
                        if (tl > 0) mbar_wait(&a_ready[s], (tl * S - 1u) & 1u);          // the previous tile's last accumulator has been read
                        mbar_wait(&x_full[2 * s + (tl & 1u)], (tl >> 1) & 1u);
                        tc_fence_after();
                        const uint32_t xb = xr_b + ((uint32_t)s * 2 + (tl & 1u)) * kXBytes;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < IN_DIM / 16; ++k)
                                mma_ss(d_t, smem_desc_sw(xb + k * 32, kXSw), smem_desc(w0b + k * 2 * (kW * 16), kW * 16, 128), idesc64, k > 0);
                            tc_commit(&d_full[s]);
                        }
                        __syncwarp();
                    } else {
                        mbar_wait(&a_ready[s], (tl * S + (uint32_t)(L - 1)) & 1u);
                        tc_fence_after();
                        if (elect_one()) {
                            // layer 0 of this tile has completed (its epilogue ran): the input buffer can take the tile after next
                            if (L == 1 && tl + 2 < nt[s]) issue_x(s, tl + 2);
                            if (L < S - 1) {
                                const uint32_t wb = whb + (uint32_t)(L - 1) * 8192u;
#pragma unroll
                                for (int k = 0; k < 4; ++k) mma_ts(d_t, a_t + k * 8, smem_desc(wb + k * 2 * (kW * 16), kW * 16, 128), idesc64, k > 0);
                            } else {
#pragma unroll
                                for (int k = 0; k < 4; ++k) mma_ts(d_t, a_t + k * 8, smem_desc(wlb + k * 2 * (16 * 16), 16 * 16, 128), idesc16, k > 0);
                            }
                            tc_commit(&d_full[s]);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ===================== epilogue warps (4 per slot) =====================
        const int s = (warp - 1) >> 2;                 // slot
        const int q = warp & 3;                        // TMEM quarter this warp may access
        const int r = q * 32 + lane;                   // row of the tile
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const uint32_t d_t = tmem0 + lane_sel + s * kSlotCols, a_t = d_t + 64;
        const bool issuer = (q == 0 && lane == 0);     // issues this slot's TMA stores
        uint32_t n_store = 0;                          // activation tiles stored so far by this slot (staging buffer = n_store & 1)
        uint32_t tl = 0;
        for (uint32_t j = s; j < my_tiles; j += NSLOTS, ++tl) {
            const size_t tile = (size_t)blockIdx.x + (size_t)j * gridDim.x;
            const size_t row = tile * kTile + r;
            float dx = 0.f, dy = 0.f, dz = 0.f;
            if (HEAD == 1) {                           // requested now, used after the last layer
                dx = __ldg(head.dirs + row * 3);
                dy = __ldg(head.dirs + row * 3 + 1);
                dz = __ldg(head.dirs + row * 3 + 2);
            }
#pragma unroll
            for (int L = 0; L < S; ++L) {
                mbar_wait(&d_full[s], (tl * S + (uint32_t)L) & 1u);
                tc_fence_after();
                if (L < S - 1) {
                    uint8_t* sb = stg + ((size_t)s * 2 + (n_store & 1u)) * kGBytes;
                    // both halves of the accumulator row are requested before the first is used (one TMEM round trip instead of two)
                    uint32_t acc2[2][32];
                    tmem_ld32(d_t, acc2[0]);
                    tmem_ld32(d_t + 32, acc2[1]);
                    tc_wait_ld();
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t (&acc)[32] = acc2[h];
                        uint32_t p[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) p[e] = pack2_relu(__uint_as_float(acc[2 * e]), __uint_as_float(acc[2 * e + 1]));
                        tmem_st16(a_t + h * 16, p);
                        if (TRAIN) {
#pragma unroll
                            for (int v = 0; v < 4; ++v)
                                *reinterpret_cast<int4*>(sb + sw_off((uint32_t)r, (uint32_t)(h * 4 + v), 128)) =
                                    make_int4((int)p[4 * v], (int)p[4 * v + 1], (int)p[4 * v + 2], (int)p[4 * v + 3]);
                        }
                    }
                    tc_wait_st();
                    if (TRAIN) fence_proxy_async_smem();
                    tc_fence_before();
                    warp_arrive(&a_ready[s], lane);
                    if (TRAIN) {
                        // the store issued one layer ago has long finished reading its buffer; waiting for it here (before the barrier)
                        // tells every thread of the slot that the buffer they will fill NEXT is free
                        if (issuer) tma_store_wait_read<0>();
                        named_bar_sync(1 + s, 128);
                        if (issuer) {
                            tma_store_2d(&tm_fb, smem_u32(sb), 0, (int32_t)((uint32_t)L * B + (uint32_t)tile * kTile));
                            tma_store_commit();
                        }
                        ++n_store;
                    }
                } else {
                    uint32_t acc[16];
                    tmem_ld16(d_t, acc);
                    tc_wait_ld();
                    tc_fence_before();
                    warp_arrive(&a_ready[s], lane);              // accumulator read: the slot can start its next tile
                    if (HEAD == 3 || HEAD == 4) head.sigma[row] = expf(f16_round(__uint_as_float(acc[0])));
                    if (HEAD == 0 || HEAD == 4) {
                        // 32-byte output rows: two 16-byte stores per lane (consecutive lanes -> consecutive rows)
                        int4* o = reinterpret_cast<int4*>(out + row * 16);
                        o[0] = make_int4((int)pack2(__uint_as_float(acc[0]), __uint_as_float(acc[1])), (int)pack2(__uint_as_float(acc[2]), __uint_as_float(acc[3])),
                                         (int)pack2(__uint_as_float(acc[4]), __uint_as_float(acc[5])), (int)pack2(__uint_as_float(acc[6]), __uint_as_float(acc[7])));
                        o[1] = make_int4((int)pack2(__uint_as_float(acc[8]), __uint_as_float(acc[9])), (int)pack2(__uint_as_float(acc[10]), __uint_as_float(acc[11])),
                                         (int)pack2(__uint_as_float(acc[12]), __uint_as_float(acc[13])), (int)pack2(__uint_as_float(acc[14]), __uint_as_float(acc[15])));
                    } else if (HEAD == 1) {
                        head.sigma[row] = expf(f16_round(__uint_as_float(acc[0])));
                        // directions reach the SH encoder as fp16 under autocast (sphere_harmonics.py:16)
                        float sh[16];
                        sh_deg4(f16_round(dx), f16_round(dy), f16_round(dz), sh);
                        uint32_t p[16];
#pragma unroll
                        for (int e = 0; e < 8; ++e) p[e] = pack2(sh[2 * e], sh[2 * e + 1]);
#pragma unroll
                        for (int e = 0; e < 7; ++e) p[8 + e] = pack2(__uint_as_float(acc[1 + 2 * e]), __uint_as_float(acc[2 + 2 * e]));
                        p[15] = pack2(__uint_as_float(acc[15]), 0.0f);
                        uint8_t* cb = cst + (size_t)s * kCinBytes;
                        if (issuer) tma_store_wait_read<0>();      // the previous tile's colour-input store has finished reading cb
                        named_bar_sync(1 + s, 128);
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            *reinterpret_cast<int4*>(cb + sw_off((uint32_t)r, (uint32_t)v, 64)) =
                                make_int4((int)p[4 * v], (int)p[4 * v + 1], (int)p[4 * v + 2], (int)p[4 * v + 3]);
                        fence_proxy_async_smem();
                        named_bar_sync(1 + s, 128);
                        if (issuer) {
                            tma_store_2d(&tm_cin, smem_u32(cb), 0, (int32_t)((uint32_t)tile * kTile));
                            tma_store_commit();
                        }
                    } else if (HEAD == 2) {
                        for (int c = 0; c < head.n_ch; ++c) {
                            const float y = f16_round(__uint_as_float(acc[c]));
                            head.rgb[row * head.n_ch + c] = f16_round(1.0f / (1.0f + expf(-y)));
                        }
                    }
                }
            }
        }
        if (issuer) tma_store_wait<0>();               // all bulk stores of this slot have completed before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, kCols);
}

// Persistent grids: one CTA per SM, optionally capped (enerf_ffmlp_set_max_ctas) so that a kernel with a different bottleneck
// (the hash-grid scatter, bound by L2 reductions) can run on the remaining SMs from another stream at the same time.
static int g_max_ctas = 0;    // 0 = one CTA per SM
void tc_set_max_ctas(int n) { g_max_ctas = (n <= 0 || n > num_sms()) ? 0 : n; }
static inline uint32_t tc_grid(uint32_t n_tiles) {
    const uint32_t cap = (uint32_t)(g_max_ctas > 0 ? g_max_ctas : num_sms());
    return n_tiles < cap ? n_tiles : cap;
}

template <int NSLOTS, int NH, int IN_DIM, int HEAD, bool TRAIN>
static int launch_fwd_tma_n(const TmaDesc& tx, const TmaDesc& tfb, const TmaDesc& tcin, const __half* W, __half* out, uint32_t B, HeadArgs head,
                            cudaStream_t st, const char* name) {
    size_t smem = 1024 + (size_t)NSLOTS * 2 * kTile * IN_DIM * 2 + (TRAIN ? (size_t)NSLOTS * 2 * kGBytes : 0) + (HEAD == 1 ? (size_t)NSLOTS * kTile * 64 : 0) +
                  (size_t)IN_DIM * 128 + (size_t)NH * 8192 + 2048 + 4 * NSLOTS * 8 + 16;
    if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM (TMEM)
    if (smem > 227 * 1024) { set_error("%s: shared-memory budget exceeded", name); return -2; }
    static bool configured = false;
    if (!configured) {
        ENERF_CUDA(cudaFuncSetAttribute(k_tc_fwd_tma<NSLOTS, NH, IN_DIM, HEAD, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), name);
        configured = true;
    }
    const uint32_t n_tiles = B / kTile;
    const uint32_t grid = tc_grid(n_tiles);
    k_tc_fwd_tma<NSLOTS, NH, IN_DIM, HEAD, TRAIN><<<grid, 32 + NSLOTS * 128, smem, st>>>(tx, tfb, tcin, W, out, n_tiles, B, head);
    ENERF_CHECK_LAUNCH(name);
    return 0;
}

// n_hidden_mm = hidden-to-hidden matmuls (0, 1 or 2); fwd_buf != NULL stores every hidden activation (the reference's training contract)
#ifndef ENERF_FWD_SLOTS
#define ENERF_FWD_SLOTS 5       // tiles in flight per CTA of the forward without forward_buffer (3: 0.097 / 0.099 ms, 4: 0.088 / 0.096, 5: 0.084 / 0.092; 6 x 96 columns exceed TMEM)
#endif
template <int HEAD>
static int launch_fwd(const __half* in, const __half* W, uint32_t B, int n_hidden_mm, __half* fwd_buf, __half* out, HeadArgs head, cudaStream_t st,
                      const char* name) {
    constexpr int IN_DIM = 32;
    if (n_hidden_mm < 0 || n_hidden_mm > 2 || (uint64_t)(n_hidden_mm + 1) * B >= (1ull << 31)) {
        set_error("%s: the tcgen05 path takes 1 to 3 layers and fewer than 2^31 activation rows", name);
        return -2;
    }
    TmaDesc tx, tfb, tcin;
    if (!make_tmap_rows(&tx, in, B, IN_DIM, kTile)) { set_error("%s: inputs must be 16-byte aligned (TMA)", name); return -2; }
    if ((HEAD == 0 || HEAD == 4) && (reinterpret_cast<uintptr_t>(out) & 15u)) { set_error("%s: outputs must be 16-byte aligned", name); return -2; }
    tfb = tx;
    tcin = tx;
    if (fwd_buf && !make_tmap_rows(&tfb, fwd_buf, (uint64_t)(n_hidden_mm + 1) * B, 64, kTile)) { set_error("%s: bad forward_buffer", name); return -2; }
    if (HEAD == 1 && !make_tmap_rows(&tcin, head.cin, B, 32, kTile)) { set_error("%s: bad colour-input buffer", name); return -2; }
    if (fwd_buf) {
        if constexpr (HEAD <= 2) {
            if (n_hidden_mm == 1) return launch_fwd_tma_n<3, 1, IN_DIM, HEAD, true>(tx, tfb, tcin, W, out, B, head, st, name);
            if (n_hidden_mm == 2) return launch_fwd_tma_n<3, 2, IN_DIM, HEAD, true>(tx, tfb, tcin, W, out, B, head, st, name);
        }
        set_error("%s: forward_buffer is stored for 2- and 3-layer networks only", name);
        return -2;
    }
    if (n_hidden_mm == 0) {
        if constexpr (HEAD == 0 || HEAD == 3 || HEAD == 4) return launch_fwd_tma_n<ENERF_FWD_SLOTS, 0, IN_DIM, HEAD, false>(tx, tfb, tcin, W, out, B, head, st, name);
        set_error("%s: a 1-layer network has no fused colour head", name);
        return -2;
    }
    if (n_hidden_mm == 1) return launch_fwd_tma_n<ENERF_FWD_SLOTS, 1, IN_DIM, HEAD, false>(tx, tfb, tcin, W, out, B, head, st, name);
    if constexpr (HEAD <= 2) return launch_fwd_tma_n<ENERF_FWD_SLOTS, 2, IN_DIM, HEAD, false>(tx, tfb, tcin, W, out, B, head, st, name);
    set_error("%s: density heads take 1- and 2-layer networks", name);
    return -2;
}

int tc_forward(const __half* in, const __half* W, uint32_t B, int in_dim, int n_hidden_mm, __half* fwd_buf, __half* out, cudaStream_t st,
               const char* name) {
    HeadArgs none = {nullptr, nullptr, nullptr, nullptr, 0, nullptr};
    if (in_dim != 32) { set_error("%s: input_dim must be 32 on the tcgen05 path", name); return -2; }
    return launch_fwd<0>(in, W, B, n_hidden_mm, fwd_buf, out, none, st, name);
}

// sigma-net with fused exp / SH / colour-input head (input_dim 32)
int tc_forward_sigma_head(const __half* feat, const __half* W, uint32_t B, int n_hidden_mm, __half* fwd_buf, const float* dirs, float* sigma,
                          __half* cin, cudaStream_t st) {
    HeadArgs h = {dirs, sigma, cin, nullptr, 0, nullptr};
    return launch_fwd<1>(feat, W, B, n_hidden_mm, fwd_buf, nullptr, h, st, "field_sigma_forward");
}
// colour-net with fused sigmoid head (input_dim 32)
int tc_forward_rgb_head(const __half* cin, const __half* W, uint32_t B, int n_hidden_mm, __half* fwd_buf, float* rgb, int n_ch, const int32_t* n_rows_dev,
                        cudaStream_t st) {
    HeadArgs h = {nullptr, nullptr, nullptr, rgb, n_ch, n_rows_dev};
    return launch_fwd<2>(cin, W, B, n_hidden_mm, fwd_buf, nullptr, h, st, "field_color_forward");
}
// density head: sigma (+ the 16 raw outputs when h != NULL) of a 1- or 2-layer sigma-net
int tc_forward_density(const __half* feat, const __half* W, uint32_t B, int n_hidden_mm, float* sigma, __half* h, cudaStream_t st) {
    HeadArgs a = {nullptr, sigma, nullptr, nullptr, 0, nullptr};
    if (h) return launch_fwd<4>(feat, W, B, n_hidden_mm, nullptr, h, a, st, "field_density_forward");
    return launch_fwd<3>(feat, W, B, n_hidden_mm, nullptr, nullptr, a, st, "field_density_forward");
}

// ================================================================================================
// Backward: activation gradients AND weight gradients in one persistent kernel.
//
//   g_n   = (dy  . W_last) * relu'(h_n)              dgrad: A = dy/g in TMEM (K-major), B = the forward
//   g_i-1 = (g_i . W_i)    * relu'(h_i-1)            weight tile read MN-major (same smem bytes)
//   dx    =  g_0 . W_0
//   dW_last^T += h_n^T . dy     dW_i += g_i^T . h_i-1     dW_0 += g_0^T . x
//                                                     wgrad: M = 64, K = the 128 samples of the tile;
//                                                     both operands are MN-major reads of sample-major
//                                                     tiles the epilogue warps leave in shared memory;
//                                                     fp32 accumulators stay in TMEM for the CTA's whole
//                                                     life and are reduced once with red.global.add.
//   Activation gradients never touch HBM unless the caller asks for backward_buffer.
//
// Stage k (k = 0 .. NH+1) of a tile: epilogue E_k prepares operands, the issuing warp launches
// dgrad_k + wgrad_k, tcgen05.commit -> E_k+1 ... (see the schedule in DESIGN.md).
// ================================================================================================

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Optional fused prologues (where dL/dy of the network comes from):
//   PRO 0: dy [B,16] fp16 as given
//   PRO 1 (colour-net): dy[c] = fp16(g_rgb[c]) * rgb[c] * (1 - rgb[c]) for c < n_ch (sigmoid'), 0 otherwise
//   PRO 2 (sigma-net) : dy[0] = g_sigma * exp(clamp(h0, -15, 15)) (trunc_exp', activation.py:15-18; h0 = log sigma),
//                       dy[1:16] = dL/d(colour-net input)[16:31] (the geo_feat columns)
//   PRO 3 (density head, nerf/network.py:134-151): dy[0] = g_sigma * exp(clamp(h0, -15, 15)) + grad[0], dy[1:16] = grad[1:16]
//                       (grad [B,16] fp16 = dL/dh as autograd delivers it for the geo_feat slice; may be NULL = zero)
struct ProArgs {
    const float* g_rgb;     // [B,n_ch]  (PRO 1)
    const float* rgb;       // [B,n_ch]  (PRO 1)
    int n_ch;
    const float* g_sigma;   // [B]       (PRO 2)
    const float* sigma;     // [B]       (PRO 2)
    const __half* dcin;     // [B,32]    (PRO 2)
    const int32_t* n_rows_dev;   // optional device-side row count (see HeadArgs)
};

// ================================================================================================
// Backward from stored activations (k_tc_bwd_tma; the reference's contract: forward_buffer in).  The first version of this
// kernel had every epilogue thread read its own 128-byte activation row (8 strided 16-byte loads per stage: 32 L1 wavefronts
// per request, the LSU data pipe sat at 75-80 % in ncu) and copy it into a canonical operand tile.  Here the issuing warp
// launches ONE cp.async.bulk.tensor per stage: the [128 x 64] fp16 tile of forward_buffer (or the
// [128 x in_dim] input tile for the last stage) lands in shared memory in the 128-byte (64-byte)
// TMA swizzle, which tcgen05.mma reads directly as the MN-major wgrad operand; a RING-deep ring per
// slot keeps the loads of the next RING-1 stages in flight while stage i computes.  Epilogue threads only read their row's
// ReLU mask from that tile (8 conflict-free LDS.128) and never touch global memory for activations.
//   ring protocol (per slot, stage counter i runs across tiles): load(i) -> buffer i%RING, signalled on
//   h_full[slot][i%RING]; the buffer is free again when E_{i+1} has read its mask (a_ready of stage i+1)
//   and wgrad(i) has completed (d_full of stage i, which E_{i+1} waited for) -> the issuing warp launches
//   load(i+RING) right after it observes a_ready(i+1).
// ================================================================================================

template <int NSLOTS, int RING, int NH, int PRO, int IN_DIM>
__global__ void __launch_bounds__(32 + NSLOTS * 128, 1)
k_tc_bwd_tma(const __grid_constant__ TmaDesc tm_h, const __grid_constant__ TmaDesc tm_x, const __half* __restrict__ grad,
             const __half* __restrict__ W, __half* __restrict__ bwd_buf, __half* __restrict__ grad_inputs, float* __restrict__ dW, uint32_t n_tiles,
             uint32_t B, ProArgs pro) {
    // NH = hidden-to-hidden matmuls (compile time): every stage of a tile is unrolled, so descriptors, ring buffers and
    // accumulator columns are constants relative to a handful of uniform registers and the issuing warp spends a few
    // instructions per MMA (with run-time stage dispatch it needed ~300 instructions per stage and was the bottleneck).
    constexpr int in_dim = IN_DIM;
    constexpr int S = NH + 2;                              // stages per tile
    static_assert(S % RING == 0 && RING >= 2, "the ring depth must divide the stage count (static buffer indices)");
    constexpr uint32_t kXSw = IN_DIM * 2;                  // swizzle bytes of the input tile (row bytes): 64 or 128
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint8_t* hring = smem;                                          // NSLOTS x RING x 16 KB, 1024-byte aligned (TMA swizzle atoms)
    uint8_t* gtiles = hring + (size_t)NSLOTS * RING * kGBytes;      // NSLOTS x 16 KB, canonical [chunk][row][16 B]
    uint8_t* w0s = gtiles + (size_t)NSLOTS * kGBytes;               // [in_dim/8][64][16 B]   (forward layout)
    uint8_t* whs = w0s + in_dim * 128;                              // NH x [8][64][16 B]
    uint8_t* wls = whs + NH * 8192;                                 // [8][16][16 B]
    uint64_t* a_ready = reinterpret_cast<uint64_t*>(wls + 2048);
    uint64_t* d_full = a_ready + NSLOTS;
    uint64_t* h_full = d_full + NSLOTS;                             // [NSLOTS][RING]
    uint64_t* flush_bar = h_full + RING * NSLOTS;
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(flush_bar + 1);

    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t kCols = 512;

    stage_matrix(w0s, W, kW, in_dim, tid, nthreads);
    for (int j = 0; j < NH; ++j) stage_matrix(whs + j * 8192, W + kW * in_dim + j * kW * kW, kW, kW, tid, nthreads);
    stage_matrix(wls, W + kW * in_dim + NH * kW * kW, 16, kW, tid, nthreads);
    if (tid == 0) {
        for (int s = 0; s < NSLOTS; ++s) {
            mbar_init(&a_ready[s], 4);      // one arrival per epilogue warp
            mbar_init(&d_full[s], 1);
            for (int r = 0; r < RING; ++r) mbar_init(&h_full[RING * s + r], 1);
        }
        mbar_init(flush_bar, 1);
        mbar_fence_init();
        tma_prefetch_desc(&tm_h);
        tma_prefetch_desc(&tm_x);
    }
    if (warp == 0) tmem_alloc(tmem_base_ptr, kCols);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *tmem_base_ptr;

    n_tiles = effective_tiles(n_tiles, pro.n_rows_dev);
    const uint32_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // TMEM columns: [slot s: D 64 | A 32] x NSLOTS, then the weight-gradient accumulators [dW_last^T : 16][dW_hidden j : 64 each][dW_0 : in_dim]
    constexpr uint32_t kAccLast = NSLOTS * kSlotCols, kAccHid = kAccLast + 16, kAcc0 = kAccHid + NH * 64;
    static_assert(kAcc0 + IN_DIM <= 512, "TMEM budget");

    if (warp == 0) {
        // ===================== MMA / TMA issuer: warp-uniform control flow, one elected lane issues =====================
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem0, 0);
        const uint32_t hring_b = smem_u32(hring), gt_b = smem_u32(gtiles), w0b = smem_u32(w0s), whb = smem_u32(whs), wlb = smem_u32(wls);
        uint32_t nt[NSLOTS];
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) nt[s] = (my_tiles > (uint32_t)s) ? (my_tiles - s + NSLOTS - 1) / NSLOTS : 0;
        // activation tile of stage k of the slot's tl-th tile -> ring buffer k % RING   (elected lane only)
        auto issue_load = [&](int s, uint32_t tl, int k) {
            const uint32_t tile = blockIdx.x + ((uint32_t)s + tl * NSLOTS) * gridDim.x;
            const uint32_t dst = hring_b + (uint32_t)(s * RING + (k % RING)) * kGBytes;
            uint64_t* bar = &h_full[RING * s + (k % RING)];
            if (k < S - 1) {
                mbar_arrive_expect_tx(bar, kGBytes);
                tma_load_2d(dst, &tm_h, 0, (int32_t)((uint32_t)(NH - k) * B + tile * kTile), bar);
            } else {
                mbar_arrive_expect_tx(bar, kTile * in_dim * 2);
                tma_load_2d(dst, &tm_x, 0, (int32_t)(tile * kTile), bar);
            }
        };
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) {
            if (nt[s] > 0 && elect_one()) {
#pragma unroll
                for (int k = 0; k < RING; ++k) issue_load(s, 0, k);
            }
            __syncwarp();
        }
        constexpr uint32_t idD64 = idesc_f16(kTile, 64, false, true), idDx = idesc_f16(kTile, (uint32_t)IN_DIM, false, true);
        constexpr uint32_t idWl = idesc_f16(64, 16, true, true), idWh = idesc_f16(64, 64, true, true), idW0 = idesc_f16(64, (uint32_t)IN_DIM, true, true);
        for (uint32_t tl = 0; tl < nt[0]; ++tl) {
#pragma unroll
            for (int k = 0; k < S; ++k) {
#pragma unroll
                for (int s = 0; s < NSLOTS; ++s) {
                    if (tl >= nt[s]) continue;
                    const uint32_t i = tl * S + k;                       // slot-local stage counter
                    const uint32_t d_t = tm + s * kSlotCols, a_t = d_t + 64;
                    const uint32_t g_s = gt_b + (uint32_t)s * kGBytes;
                    const uint32_t h_s = hring_b + (uint32_t)(s * RING + (k % RING)) * kGBytes;
                    mbar_wait(&a_ready[s], i & 1u);
                    tc_fence_after();
                    // E_i is done: it has read its ReLU mask from the buffer of stage i-1, whose wgrad completed before E_i started,
                    // so that buffer takes load i-1+RING = stage (k-1+RING) % S of tile tl + (k-1+RING) / S
                    const int kl = (k - 1 + RING) % S, dt = (k - 1 + RING) / S;
                    const bool do_load = (i >= 1) && (tl + dt < nt[s]);
                    if (elect_one()) {
                        if (do_load) issue_load(s, tl + dt, kl);
                        if (k == 0) {
                            // dgrad through the output layer: D[128x64] = dy[128x16] . W_last[16x64]
                            mma_ts(d_t, a_t, smem_desc(wlb, 128, 16 * 16), idD64, false);
                        } else if (k <= NH) {
                            const uint32_t wj = whb + (uint32_t)(NH - k) * 8192u;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) mma_ts(d_t, a_t + ks * 8, smem_desc(wj + ks * 256, 128, 64 * 16), idD64, ks > 0);
                        } else {
                            // dx = g_0 . W_0 (issued even when the caller does not ask for grad_inputs: 4 small MMAs)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) mma_ts(d_t, a_t + ks * 8, smem_desc(w0b + ks * 256, 128, 64 * 16), idDx, ks > 0);
                        }
                    }
                    __syncwarp();
                    // wgrad needs the activation tile of this stage: fill number tl*(S/RING) + k/RING of its ring buffer
                    mbar_wait(&h_full[RING * s + (k % RING)], (tl * (S / RING) + (uint32_t)(k / RING)) & 1u);
                    tc_fence_after();
                    const bool acc = !(tl == 0 && s == 0);               // the very first issue on an accumulator overwrites it
                    if (elect_one()) {
                        if (k == 0) {
                            // dW_last^T[64x16] += h_n^T[64x128] . dy[128x16]   (A = activation tile, B = dy tile in the G buffer)
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks)
                                mma_ss(tm + kAccLast, smem_desc_sw(h_s + ks * 2048, 128), smem_desc(g_s + ks * 256, 128, 2048), idWl, acc || ks > 0);
                        } else if (k <= NH) {
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks)
                                mma_ss(tm + kAccHid + (uint32_t)(NH - k) * 64u, smem_desc(g_s + ks * 256, 128, 2048), smem_desc_sw(h_s + ks * 2048, 128), idWh,
                                       acc || ks > 0);
                        } else {
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks)
                                mma_ss(tm + kAcc0, smem_desc(g_s + ks * 256, 128, 2048), smem_desc_sw(h_s + ks * 16 * kXSw, kXSw), idW0, acc || ks > 0);
                        }
                        tc_commit(&d_full[s]);
                    }
                    __syncwarp();
                }
            }
        }
        if (elect_one()) tc_commit(flush_bar);
        __syncwarp();
    } else {
        // ===================== epilogue warps (4 per slot) =====================
        const int s = (warp - 1) >> 2;
        const int q = warp & 3;
        const int r_in_tile = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const uint32_t d_t = tmem0 + lane_sel + s * kSlotCols, a_t = d_t + 64;
        uint8_t* g_tile = gtiles + (size_t)s * kGBytes;

        // dL/dy of this row: raw operands are fetched one tile ahead (their HBM latency would otherwise sit at the head of every tile)
        int4 pv0 = make_int4(0, 0, 0, 0), pv1 = make_int4(0, 0, 0, 0);
        float pf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        auto fetch = [&](size_t row) {
            if (PRO == 0) {
                const int4* src = reinterpret_cast<const int4*>(grad + row * 16);
                pv0 = __ldg(src);
                pv1 = __ldg(src + 1);
            } else if (PRO == 1) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < pro.n_ch) {
                        pf[c] = __ldg(pro.rgb + row * pro.n_ch + c);
                        pf[4 + c] = __ldg(pro.g_rgb + row * pro.n_ch + c);
                    }
            } else if (PRO == 2) {
                pf[0] = __ldg(pro.sigma + row);
                pf[1] = __ldg(pro.g_sigma + row);
                const int4* src = reinterpret_cast<const int4*>(pro.dcin + row * 32 + 16);
                pv0 = __ldg(src);
                pv1 = __ldg(src + 1);
            } else {
                pf[0] = __ldg(pro.sigma + row);
                pf[1] = pro.g_sigma ? __ldg(pro.g_sigma + row) : 0.f;
                if (grad) {
                    const int4* src = reinterpret_cast<const int4*>(grad + row * 16);
                    pv0 = __ldg(src);
                    pv1 = __ldg(src + 1);
                }
            }
        };
        if ((uint32_t)s < my_tiles) fetch(((size_t)blockIdx.x + (size_t)s * gridDim.x) * kTile + r_in_tile);

        uint32_t tl = 0;
        for (uint32_t j = s; j < my_tiles; j += NSLOTS, ++tl) {
            const size_t tile = (size_t)blockIdx.x + (size_t)j * gridDim.x;
            const size_t row = tile * kTile + r_in_tile;
            // ---- E_0: dy -> TMEM A + dy tile (G buffer)
            {
                int4 v0, v1;
                if (PRO == 0) {
                    v0 = pv0;
                    v1 = pv1;
                } else if (PRO == 1) {
                    float dyv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < pro.n_ch) dyv[c] = f16_round(pf[4 + c]) * (1.0f - pf[c]) * pf[c];
                    v0 = make_int4((int)pack2(dyv[0], dyv[1]), (int)pack2(dyv[2], dyv[3]), 0, 0);
                    v1 = make_int4(0, 0, 0, 0);
                } else {
                    const float sg = fminf(fmaxf(pf[0], 3.0590232050182579e-07f), 3269017.3724721107f);   // exp(-15), exp(15)
                    const __half d0 = __float2half_rn(pf[1] * sg);
                    // dcin columns 16..30 -> dy columns 1..15 (shift by one fp16)
                    const uint32_t w[8] = {(uint32_t)pv0.x, (uint32_t)pv0.y, (uint32_t)pv0.z, (uint32_t)pv0.w,
                                           (uint32_t)pv1.x, (uint32_t)pv1.y, (uint32_t)pv1.z, (uint32_t)pv1.w};
                    uint32_t o[8];
                    o[0] = (uint32_t)__half_as_ushort(d0) | (w[0] << 16);
#pragma unroll
                    for (int e = 1; e < 8; ++e) o[e] = (w[e - 1] >> 16) | (w[e] << 16);
                    v0 = make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]);
                    v1 = make_int4((int)o[4], (int)o[5], (int)o[6], (int)o[7]);
                }
                const uint32_t r8[8] = {(uint32_t)v0.x, (uint32_t)v0.y, (uint32_t)v0.z, (uint32_t)v0.w,
                                        (uint32_t)v1.x, (uint32_t)v1.y, (uint32_t)v1.z, (uint32_t)v1.w};
                tmem_st8(a_t, r8);
                *reinterpret_cast<int4*>(g_tile + 0 * 2048 + r_in_tile * 16) = v0;
                *reinterpret_cast<int4*>(g_tile + 1 * 2048 + r_in_tile * 16) = v1;
                tc_wait_st();
                fence_proxy_async_smem();
                tc_fence_before();
                warp_arrive(&a_ready[s], lane);
                if (j + NSLOTS < my_tiles) fetch(((size_t)blockIdx.x + (size_t)(j + NSLOTS) * gridDim.x) * kTile + r_in_tile);
            }
            // ---- E_k, k = 1 .. S-1: g = D * relu'(h of stage k-1) -> TMEM A + G tile
#pragma unroll
            for (int k = 1; k < S; ++k) {
                mbar_wait(&d_full[s], (tl * S + (uint32_t)(k - 1)) & 1u);
                tc_fence_after();
                const int hb = (k - 1) % RING;                                                           // ring buffer of stage k-1
                mbar_wait(&h_full[RING * s + hb], (tl * (S / RING) + (uint32_t)((k - 1) / RING)) & 1u);   // complete long ago; orders our reads after the TMA writes
                const uint8_t* hrow = hring + (size_t)(s * RING + hb) * kGBytes;
                int4 hv[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) hv[c] = *reinterpret_cast<const int4*>(hrow + sw_off((uint32_t)r_in_tile, (uint32_t)c, 128));
                const __half2 zero2 = __floats2half2_rn(0.f, 0.f);
                uint32_t acc2[2][32];
                tmem_ld32(d_t, acc2[0]);
                tmem_ld32(d_t + 32, acc2[1]);
                tc_wait_ld();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t (&acc)[32] = acc2[h];
                    uint32_t p[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const uint32_t hw = reinterpret_cast<const uint32_t*>(hv)[h * 16 + e];
                        const unsigned m = __hgt2_mask(*reinterpret_cast<const __half2*>(&hw), zero2);    // 0xffff per half where h > 0
                        p[e] = pack2(__uint_as_float(acc[2 * e]), __uint_as_float(acc[2 * e + 1])) & m;
                    }
                    tmem_st16(a_t + h * 16, p);
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int4 val = make_int4((int)p[4 * v], (int)p[4 * v + 1], (int)p[4 * v + 2], (int)p[4 * v + 3]);
                        *reinterpret_cast<int4*>(g_tile + (h * 4 + v) * 2048 + r_in_tile * 16) = val;
                        // the reference's `backward_buffer` [num_layers, B, 64] (ffmlp.cu:742-748): written only when a caller asks for it
                        if (bwd_buf) reinterpret_cast<int4*>(bwd_buf + ((size_t)(k - 1) * B + row) * kW)[h * 4 + v] = val;
                    }
                }
                tc_wait_st();
                fence_proxy_async_smem();
                tc_fence_before();
                warp_arrive(&a_ready[s], lane);
            }
            // ---- E_S: dx
            mbar_wait(&d_full[s], (tl * S + (uint32_t)(S - 1)) & 1u);
            tc_fence_after();
            if (grad_inputs) {
#pragma unroll
                for (int c = 0; c < in_dim / 16; ++c) {
                    uint32_t acc[16];
                    tmem_ld16(d_t + c * 16, acc);
                    tc_wait_ld();
                    uint32_t p[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) p[e] = pack2(__uint_as_float(acc[2 * e]), __uint_as_float(acc[2 * e + 1]));
                    int4* dst = reinterpret_cast<int4*>(grad_inputs + row * in_dim + c * 16);
                    dst[0] = make_int4((int)p[0], (int)p[1], (int)p[2], (int)p[3]);
                    dst[1] = make_int4((int)p[4], (int)p[5], (int)p[6], (int)p[7]);
                }
            }
            tc_fence_before();
        }

        // ---- flush the weight-gradient accumulators (slot 0's four warps; M = 64 -> lanes 0..15 of each quarter)
        if (s == 0 && my_tiles > 0) {
            mbar_wait(flush_bar, 0);
            tc_fence_after();
            const int m = q * 16 + lane;
            const uint32_t base = tmem0 + lane_sel;
            float* dW0 = dW;
            float* dWh = dW + kW * in_dim;
            float* dWl = dWh + (size_t)NH * kW * kW;
            {
                uint32_t acc[16];
                tmem_ld16(base + kAccLast, acc);
                tc_wait_ld();
                if (lane < 16)
#pragma unroll
                    for (int nn = 0; nn < 16; ++nn) atomicAdd(dWl + nn * kW + m, __uint_as_float(acc[nn]));
            }
            for (int jj = 0; jj < NH; ++jj)
                for (int c = 0; c < 4; ++c) {
                    uint32_t acc[16];
                    tmem_ld16(base + kAccHid + jj * 64 + c * 16, acc);
                    tc_wait_ld();
                    if (lane < 16) {
                        float* dst = dWh + (size_t)jj * kW * kW + (size_t)m * kW + c * 16;
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            red_add_v4(dst + 4 * v, __uint_as_float(acc[4 * v]), __uint_as_float(acc[4 * v + 1]), __uint_as_float(acc[4 * v + 2]),
                                       __uint_as_float(acc[4 * v + 3]));
                    }
                }
            for (int c = 0; c < in_dim / 16; ++c) {
                uint32_t acc[16];
                tmem_ld16(base + kAcc0 + c * 16, acc);
                tc_wait_ld();
                if (lane < 16) {
                    float* dst = dW0 + (size_t)m * in_dim + c * 16;
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        red_add_v4(dst + 4 * v, __uint_as_float(acc[4 * v]), __uint_as_float(acc[4 * v + 1]), __uint_as_float(acc[4 * v + 2]),
                                   __uint_as_float(acc[4 * v + 3]));
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, kCols);
}

#ifdef ENERF_TC_TRACE
}  // namespace tcm
}  // namespace enerf
extern "C" int enerf_debug_set_trace(unsigned long long* buf) {
    cudaMemcpyToSymbol(enerf::tcm::g_trace, &buf, sizeof(buf));
    return 0;
}
namespace enerf {
namespace tcm {
#endif
template <int NSLOTS, int RING, int NH, int PRO, int IN_DIM>
static int launch_bwd_tma_n(const TmaDesc& th, const TmaDesc& tx, const __half* grad, const __half* W, __half* bwd_buf, __half* grad_inputs, float* dW,
                            uint32_t B, ProArgs pro, cudaStream_t st, const char* name) {
    static_assert((NH + 2) % RING == 0 && NSLOTS * kSlotCols + 16 + 64 * NH + IN_DIM <= 512, "ring / TMEM budget");
    size_t smem = 1024 + (size_t)NSLOTS * (RING + 1) * kGBytes + (size_t)IN_DIM * 128 + (size_t)NH * 8192 + 2048 + ((2 + RING) * NSLOTS + 1) * 8 + 16;
    if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM: it allocates all 512 TMEM columns
    if (smem > 227 * 1024) { set_error("%s: shared-memory budget exceeded", name); return -2; }
    static bool configured = false;
    if (!configured) {
        ENERF_CUDA(cudaFuncSetAttribute(k_tc_bwd_tma<NSLOTS, RING, NH, PRO, IN_DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), name);
        configured = true;
    }
    const uint32_t n_tiles = B / kTile;
    const uint32_t grid = tc_grid(n_tiles);
    k_tc_bwd_tma<NSLOTS, RING, NH, PRO, IN_DIM><<<grid, 32 + NSLOTS * 128, smem, st>>>(th, tx, grad, W, bwd_buf, grad_inputs, dW, n_tiles, B, pro);
    ENERF_CHECK_LAUNCH(name);
    return 0;
}

// backward from stored activations: 3 tiles in flight; 3 stages per tile -> activation ring of 3, 4 stages -> ring of 2
// (measured, 3.29 M samples: 2 slots 0.25 / 0.34 ms, 3 slots 0.22 / 0.30 ms for the 2- / 3-layer net)
template <int PRO>
static int launch_bwd_tma(const __half* grad, const __half* x, const __half* W, const __half* fwd_buf, __half* bwd_buf, __half* grad_inputs, float* dW,
                          uint32_t B, int in_dim, int n_hidden_mm, ProArgs pro, cudaStream_t st, const char* name) {
    if (in_dim != 32 || (n_hidden_mm != 1 && n_hidden_mm != 2) || (uint64_t)(n_hidden_mm + 1) * B >= (1ull << 31)) {
        set_error("%s: the tcgen05 path takes 32 inputs, 2 or 3 layers and fewer than 2^31 activation rows", name);
        return -2;
    }
    TmaDesc th, tx;
    if (!make_tmap_rows(&th, fwd_buf, (uint64_t)(n_hidden_mm + 1) * B, 64, kTile) || !make_tmap_rows(&tx, x, B, (uint32_t)in_dim, kTile)) {
        set_error("%s: inputs / forward_buffer must be 16-byte aligned (TMA)", name);
        return -2;
    }
    if (n_hidden_mm == 1) return launch_bwd_tma_n<3, 3, 1, PRO, 32>(th, tx, grad, W, bwd_buf, grad_inputs, dW, B, pro, st, name);
    return launch_bwd_tma_n<3, 2, 2, PRO, 32>(th, tx, grad, W, bwd_buf, grad_inputs, dW, B, pro, st, name);
}

// ================================================================================================
// Backward with recomputation (k_tc_bwd_rc): the stored-activation kernels above are bound by HBM — the training
// forward writes [num_layers, B, 64] fp16 of forward_buffer and the backward reads it back (2.2 + 2.8 GB per step on
// the bench workload), while re-running the hidden layers of a tile costs a few hundred tensor-core cycles.  This
// kernel takes only the network input and dL/dy:
//   F_0 .. F_NH : h_L = relu(h_{L-1} . W_L^T); the epilogue writes h_L (fp16) to TMEM (A operand of the next forward MMA) and into a
//                 shared-memory tile in the 128-byte swizzle, which the weight-gradient MMA reads as its MN-major operand; the
//                 ReLU mask of h_L stays with the thread as 64 bits in registers (relu_bits below).
//   B_0 .. B_S-1: exactly the stages of k_tc_bwd_tma (dgrad with A in TMEM, wgrad into TMEM accumulators).
// The recomputed activations are bit-identical to what the training forward would have stored (same MMAs, same K
// order, same rounding point), so the gradients match the stored-activation path.  Input width 32.
//   a_ready[s] / d_full[s] advance 2*NH+3 phases per tile (one per MMA stage; a_ready's last one = "accumulator read,
//   slot free for the next tile").
// ================================================================================================
// ReLU masks of the recomputed activations as bits in registers (one 32-bit word per 32 columns of a row: bit e = column 2e is
// positive, bit 16+e = column 2e+1), so that the backward epilogues do not read the activation tile back from shared memory: 16 KB
// per stage and slot less through the shared-memory pipe that bounds this kernel, 32 registers less in the backward epilogue (no
// spills at 128 any more), for ~3 more ALU instructions per column pair.  Measured (3.29 M samples): sigma-net backward 0.292 ->
// 0.249 ms, colour-net backward 0.342 -> 0.339 ms, same bits; requesting the whole accumulator row before its first half is used on
// top of this: 0.251 / 0.346 (not taken).
__device__ __forceinline__ uint32_t relu_bits(uint32_t bits, uint32_t packed_relu, int e) {
    const __half2 zero2 = __floats2half2_rn(0.f, 0.f);
    const uint32_t m = __hgt2_mask(*reinterpret_cast<const __half2*>(&packed_relu), zero2);       // 0xffff per positive half
    return bits | (m & ((1u << e) | (1u << (16 + e))));
}
// the 0xffff-per-half mask of pair e back from the word: shift the pair's two flags onto byte sign bits, replicate them (prmt)
__device__ __forceinline__ uint32_t relu_mask_from_bits(uint32_t bits, int e) {
    const uint32_t t = (e < 8) ? (bits << (7 - e)) : (bits << (15 - e));
    uint32_t m;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(m) : "r"(t), "r"(e < 8 ? 0xAA88u : 0xBB99u));
    return m;
}

// 128 registers per thread is the hardware limit for 13 warps (a scheduler partition owns 16 K registers and hosts four of them);
// requesting both 32-column halves of an accumulator row before using the first costs 16 more and spills (0.36 / 0.33 vs 0.34 / 0.29 ms)
template <int NSLOTS, int NH, int PRO, bool GD, int CH, bool XA, int NI>
__global__ void __launch_bounds__(32 * NI + NSLOTS * 128 * CH, 1)
k_tc_bwd_rc(const __grid_constant__ TmaDesc tm_x, const __half* __restrict__ grad, const __half* __restrict__ W, __half* __restrict__ grad_inputs,
            float* __restrict__ dW, uint32_t n_tiles, uint32_t B, ProArgs pro) {
    // CH = warps per TMEM lane quarter of a slot (1 or 2): with 2, a row's 64 accumulator columns are split between two
    // threads (warps w and w+4 share the quarter w % 4), which halves the serial epilogue work per stage.
    static_assert(CH == 1 || CH == 2, "CH");
    // NI = issuing warps: with 2, warp 0 issues the forward (recompute) stages and warp 1 the backward stages — they never share an
    // accumulator, and each consumes its own "A ready" barrier (a_ready[2s] / a_ready[2s+1]) phase by phase.
    static_assert(NI == 1 || NI == 2, "NI");
    constexpr int in_dim = 32;
    constexpr int CW = 64 / CH;                            // accumulator columns per epilogue thread
    constexpr int S = NH + 2;                              // backward stages per tile
    constexpr int T = 2 * NH + 3;                          // MMA stages per tile (NH+1 forward, S backward)
    constexpr uint32_t kXBytes = kTile * in_dim * 2;
    // GD: two G tiles per slot.  The dgrad of a stage is then committed on its own, so the next epilogue (which writes the other
    // G tile) overlaps the stage's eight weight-gradient MMAs instead of waiting for them.
    constexpr int NG = GD ? 2 : 1;
    // XA: no buffers of its own for the input tile.  It is loaded twice per tile (8 KB, the second time from L2): into the h_0 buffer
    // for F_0 (h_0 overwrites it afterwards) and, once h_NH is dead (after B_0), into the h_NH buffer for
    // the last stage's weight gradient.  16 KB less per slot -> one more tile in flight per SM.
    constexpr uint32_t kXRing = XA ? 0u : 2u * kXBytes;
    constexpr uint32_t kSlotBytes = kXRing + (NH + 1) * kGBytes + NG * kGBytes;     // [x ring], h_0..h_NH, G tile(s)
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint8_t* w0s = smem + (size_t)NSLOTS * kSlotBytes;     // [in_dim/8][64][16 B]   (forward layout)
    uint8_t* whs = w0s + in_dim * 128;                     // NH x [8][64][16 B]
    uint8_t* wls = whs + NH * 8192;                        // [8][16][16 B]
    uint64_t* a_ready = reinterpret_cast<uint64_t*>(wls + 2048);
    uint64_t* d_full = a_ready + 2 * NSLOTS;           // a_ready: [NSLOTS][2] = (forward issuer's, backward issuer's)
    uint64_t* x_full = d_full + NSLOTS;                    // [NSLOTS][2]
    uint64_t* flush_bar = x_full + 2 * NSLOTS;
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(flush_bar + 1);
    // per-slot regions: [x0 | x1 | h_0 .. h_NH | G]
    auto slot_h = [&](int s, int L) { return smem + (size_t)s * kSlotBytes + kXRing + (size_t)L * kGBytes; };
    auto slot_g = [&](int s, int k) { return smem + (size_t)s * kSlotBytes + kXRing + (size_t)(NH + 1 + (GD ? (k & 1) : 0)) * kGBytes; };

    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t kCols = 512;

    stage_matrix(w0s, W, kW, in_dim, tid, nthreads);
    for (int j = 0; j < NH; ++j) stage_matrix(whs + j * 8192, W + kW * in_dim + j * kW * kW, kW, kW, tid, nthreads);
    stage_matrix(wls, W + kW * in_dim + NH * kW * kW, 16, kW, tid, nthreads);
    if (tid == 0) {
        for (int s = 0; s < NSLOTS; ++s) {
            mbar_init(&a_ready[2 * s], 4 * CH);      // one arrival per epilogue warp
            mbar_init(&a_ready[2 * s + 1], 4 * CH);
            mbar_init(&d_full[s], 1);
            mbar_init(&x_full[2 * s], 1);
            mbar_init(&x_full[2 * s + 1], 1);
        }
        mbar_init(flush_bar, 1);
        mbar_fence_init();
        tma_prefetch_desc(&tm_x);
    }
    if (warp == 0) tmem_alloc(tmem_base_ptr, kCols);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *tmem_base_ptr;

    n_tiles = effective_tiles(n_tiles, pro.n_rows_dev);
    const uint32_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    constexpr uint32_t kAccLast = NSLOTS * kSlotCols, kAccHid = kAccLast + 16, kAcc0 = kAccHid + NH * 64;
    static_assert(kAcc0 + in_dim <= 512, "TMEM budget");

    if (warp < NI) {
        // ===================== MMA / TMA issuer(s): warp-uniform control flow, one elected lane issues =====================
        ENERF_TRACE_DECL(0);
        const bool do_f = (NI == 1) || warp == 0, do_b = (NI == 1) || warp == 1;
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem0, 0);
        const uint32_t sm_b = smem_u32(smem), w0b = smem_u32(w0s), whb = smem_u32(whs), wlb = smem_u32(wls);
        uint32_t nt[NSLOTS];
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) nt[s] = (my_tiles > (uint32_t)s) ? (my_tiles - s + NSLOTS - 1) / NSLOTS : 0;
        // input tile of the slot's tl-th tile -> (ring buffer tl & 1) or, with XA, the h_0 buffer (which = 0) / the h_NH buffer (which = 1)
        auto issue_x = [&](int s, uint32_t tl, uint32_t which) {               // elected lane only
            const uint32_t tile = blockIdx.x + ((uint32_t)s + tl * NSLOTS) * gridDim.x;
            uint64_t* bar = &x_full[2 * s + which];
            const uint32_t dst = sm_b + (uint32_t)s * kSlotBytes + (XA ? (which ? (uint32_t)NH * kGBytes : 0u) : which * kXBytes);
            mbar_arrive_expect_tx(bar, kXBytes);
            tma_load_2d(dst, &tm_x, 0, (int32_t)(tile * kTile), bar);
        };
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) {
            if (do_f && nt[s] > 0 && elect_one()) issue_x(s, 0, 0);
            __syncwarp();
        }
        constexpr uint32_t idF = idesc_f16(kTile, 64, false, false);
        constexpr uint32_t idD64 = idesc_f16(kTile, 64, false, true), idDx = idesc_f16(kTile, (uint32_t)in_dim, false, true);
        constexpr uint32_t idWl = idesc_f16(64, 16, true, true), idWh = idesc_f16(64, 64, true, true), idW0 = idesc_f16(64, (uint32_t)in_dim, true, true);
        for (uint32_t tl = 0; tl < nt[0]; ++tl) {
#pragma unroll
            for (int t = 0; t < T; ++t) {
#pragma unroll
                for (int s = 0; s < NSLOTS; ++s) {
                    if (tl >= nt[s]) continue;
                    if ((t <= NH) ? !do_f : !do_b) continue;
                    const uint32_t fb = tl * (NH + 1), bb = tl * S;      // first phase of this tile on the forward / backward "A ready" barrier
                    const uint32_t d_t = tm + s * kSlotCols, a_t = d_t + 64;
                    const uint32_t slot_b = sm_b + (uint32_t)s * kSlotBytes;
                    const uint32_t h0b = slot_b + kXRing;                // h_L at h0b + L*kGBytes
                    // where F_0 / the last stage find the input tile, and which barrier + parity announces it
                    const uint32_t xb_f = XA ? h0b : slot_b + (tl & 1u) * kXBytes;
                    const uint32_t xb_b = XA ? h0b + (uint32_t)NH * kGBytes : xb_f;
                    const uint32_t xw_f = XA ? 0u : (tl & 1u), xp_f = XA ? (tl & 1u) : ((tl >> 1) & 1u);
                    if (t == 0) {
                        // ---- F_0: h_0 pre-activation = x . W_0^T
                        if (tl > 0) mbar_wait(&a_ready[2 * s], (fb - 1u) & 1u);          // the previous tile's dx accumulator has been read
                        mbar_wait(&x_full[2 * s + xw_f], xp_f);
                        tc_fence_after();
                        if (elect_one()) {
                            if (s == 0) ENERF_TRACE(1000 + t);
                            if (!XA && tl + 1 < nt[s]) issue_x(s, tl + 1, (tl + 1) & 1u);       // the other ring buffer belonged to the finished tile tl-1
#pragma unroll
                            for (int k = 0; k < in_dim / 16; ++k)
                                mma_ss(d_t, smem_desc_sw(xb_f + k * 32, 64), smem_desc(w0b + k * 2 * (kW * 16), kW * 16, 128), idF, k > 0);
                            tc_commit(&d_full[s]);
                            if (s == 0) ENERF_TRACE(2000 + t);
                        }
                        __syncwarp();
                    } else if (t <= NH) {
                        // ---- F_t: h_t pre-activation = h_{t-1} . W_t^T   (A = the swizzled activation tile, K-major)
                        mbar_wait(&a_ready[2 * s], (fb + (uint32_t)(t - 1)) & 1u);
                        tc_fence_after();
                        if (elect_one()) {
                            if (s == 0) ENERF_TRACE(1000 + t);
                            // A = h_{t-1} from TMEM (the epilogue stored it there as well): the shared-memory pipe, which bounds this
                            // kernel (ncu: 37 % LSU + 39 % tensor-core operand wavefronts), is spared 16 KB of operand reads per stage
                            const uint32_t wb = whb + (uint32_t)(t - 1) * 8192u;
#pragma unroll
                            for (int k = 0; k < 4; ++k) mma_ts(d_t, a_t + k * 8, smem_desc(wb + k * 2 * (kW * 16), kW * 16, 128), idF, k > 0);
                            tc_commit(&d_full[s]);
                            if (s == 0) ENERF_TRACE(2000 + t);
                        }
                        __syncwarp();
                    } else {
                        // ---- B_k: backward stage k (see k_tc_bwd_tma); its activation tile is h_{NH-k}, or x for the last stage
                        const int k = t - (NH + 1);
                        const uint32_t g_s = h0b + (uint32_t)(NH + 1 + (GD ? (k & 1) : 0)) * kGBytes;
                        constexpr bool kSplit = GD;                      // commit the dgrad before the wgrad (all but the tile's last stage)
                        mbar_wait(&a_ready[2 * s + 1], (bb + (uint32_t)k) & 1u);
                        if (XA && k == S - 1) mbar_wait(&x_full[2 * s + 1], tl & 1u);      // the re-loaded input tile (in the h_NH buffer)
                        tc_fence_after();
                        const bool acc = !(tl == 0 && s == 0);           // the very first issue on an accumulator overwrites it
                        if (elect_one()) {
                            if (s == 0) ENERF_TRACE(1000 + t);
                            if (XA && k == 1) issue_x(s, tl, 1);                              // B_0 (the last reader of h_NH) is complete: h_NH is dead
                            if (XA && k == S - 1 && tl + 1 < nt[s]) issue_x(s, tl + 1, 0);    // E_{S-1} has read h_0 and B_NH is complete: h_0 is dead
                            if (k == 0) {
                                mma_ts(d_t, a_t, smem_desc(wlb, 128, 16 * 16), idD64, false);
                                if (kSplit) tc_commit(&d_full[s]);
                                const uint32_t h_s = h0b + (uint32_t)NH * kGBytes;
#pragma unroll
                                for (int ks = 0; ks < 8; ++ks)
                                    mma_ss(tm + kAccLast, smem_desc_sw(h_s + ks * 2048, 128), smem_desc(g_s + ks * 256, 128, 2048), idWl, acc || ks > 0);
                            } else if (k <= NH) {
                                const uint32_t wj = whb + (uint32_t)(NH - k) * 8192u;
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) mma_ts(d_t, a_t + ks * 8, smem_desc(wj + ks * 256, 128, 64 * 16), idD64, ks > 0);
                                if (kSplit) tc_commit(&d_full[s]);
                                const uint32_t h_s = h0b + (uint32_t)(NH - k) * kGBytes;
#pragma unroll
                                for (int ks = 0; ks < 8; ++ks)
                                    mma_ss(tm + kAccHid + (uint32_t)(NH - k) * 64u, smem_desc(g_s + ks * 256, 128, 2048), smem_desc_sw(h_s + ks * 2048, 128), idWh,
                                           acc || ks > 0);
                            } else {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) mma_ts(d_t, a_t + ks * 8, smem_desc(w0b + ks * 256, 128, 64 * 16), idDx, ks > 0);
#pragma unroll
                                for (int ks = 0; ks < 8; ++ks)
                                    mma_ss(tm + kAcc0, smem_desc(g_s + ks * 256, 128, 2048), smem_desc_sw(xb_b + ks * 16 * 64, 64), idW0, acc || ks > 0);
                            }
                            // the last stage always commits after its wgrad: everything of the tile (G, x, h tiles) is then free
                            if (!kSplit || k == S - 1) tc_commit(&d_full[s]);
                            if (s == 0) ENERF_TRACE(2000 + t);
                        }
                        __syncwarp();
                    }
                }
            }
        }
        if (do_b && elect_one()) tc_commit(flush_bar);
        __syncwarp();
    } else {
        // ===================== epilogue warps (4*CH per slot) =====================
        const int ew = warp - NI;
        const int s = ew / (4 * CH);
        const int q = warp & 3;                            // TMEM lane quarter this warp may access
        const int hf = (ew % (4 * CH)) >> 2;               // which CW-column part of the row this thread owns
        const int r_in_tile = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const uint32_t d_t = tmem0 + lane_sel + s * kSlotCols, a_t = d_t + 64;
        ENERF_TRACE_DECL(1);

        int4 pv0 = make_int4(0, 0, 0, 0), pv1 = make_int4(0, 0, 0, 0);
        float pf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        auto fetch = [&](size_t row) {                     // raw dL/dy operands of a row, one tile ahead
            if (PRO == 0) {
                const int4* src = reinterpret_cast<const int4*>(grad + row * 16);
                pv0 = __ldg(src);
                pv1 = __ldg(src + 1);
            } else if (PRO == 1) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < pro.n_ch) {
                        pf[c] = __ldg(pro.rgb + row * pro.n_ch + c);
                        pf[4 + c] = __ldg(pro.g_rgb + row * pro.n_ch + c);
                    }
            } else if (PRO == 2) {
                pf[0] = __ldg(pro.sigma + row);
                pf[1] = __ldg(pro.g_sigma + row);
                const int4* src = reinterpret_cast<const int4*>(pro.dcin + row * 32 + 16);
                pv0 = __ldg(src);
                pv1 = __ldg(src + 1);
            } else {
                pf[0] = __ldg(pro.sigma + row);
                pf[1] = pro.g_sigma ? __ldg(pro.g_sigma + row) : 0.f;
                if (grad) {
                    const int4* src = reinterpret_cast<const int4*>(grad + row * 16);
                    pv0 = __ldg(src);
                    pv1 = __ldg(src + 1);
                }
            }
        };
        if (hf == 0 && (uint32_t)s < my_tiles) fetch(((size_t)blockIdx.x + (size_t)s * gridDim.x) * kTile + r_in_tile);

        uint32_t tl = 0;
        for (uint32_t j = s; j < my_tiles; j += NSLOTS, ++tl) {
            const size_t tile = (size_t)blockIdx.x + (size_t)j * gridDim.x;
            const size_t row = tile * kTile + r_in_tile;
            const uint32_t tb = tl * T;
            uint32_t mbits[NH + 1][CW / 32];       // ReLU masks of h_0 .. h_NH for this thread's part of the row
            // ---- E(F_L), L = 0 .. NH: h_L = relu(D) -> swizzled activation tile; the last one also prepares dy
#pragma unroll
            for (int L = 0; L <= NH; ++L) {
                mbar_wait(&d_full[s], (tb + (uint32_t)L) & 1u);
                tc_fence_after();
                if (s == 0 && r_in_tile == 0) ENERF_TRACE(3000 + L);
                uint8_t* hb = slot_h(s, L);
#pragma unroll
                for (int h = 0; h < CW / 32; ++h) {
                    uint32_t acc[32];
                    tmem_ld32(d_t + hf * CW + h * 32, acc);
                    tc_wait_ld();
                    uint32_t p[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) p[e] = pack2_relu(__uint_as_float(acc[2 * e]), __uint_as_float(acc[2 * e + 1]));
                    uint32_t mb = 0u;
#pragma unroll
                    for (int e = 0; e < 16; ++e) mb = relu_bits(mb, p[e], e);
                    mbits[L][h] = mb;
                    if (L < NH) tmem_st16(a_t + hf * (CW / 2) + h * 16, p);       // A operand of the next forward stage
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        *reinterpret_cast<int4*>(hb + sw_off((uint32_t)r_in_tile, (uint32_t)(hf * (CW / 8) + h * 4 + v), 128)) =
                            make_int4((int)p[4 * v], (int)p[4 * v + 1], (int)p[4 * v + 2], (int)p[4 * v + 3]);
                }
                if (L < NH) tc_wait_st();
                if (L == NH && hf == 0) {
                    // dy -> TMEM A + dy tile (G buffer): E_0 of the backward
                    int4 v0, v1;
                    if (PRO == 0) {
                        v0 = pv0;
                        v1 = pv1;
                    } else if (PRO == 1) {
                        float dyv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (c < pro.n_ch) dyv[c] = f16_round(pf[4 + c]) * (1.0f - pf[c]) * pf[c];
                        v0 = make_int4((int)pack2(dyv[0], dyv[1]), (int)pack2(dyv[2], dyv[3]), 0, 0);
                        v1 = make_int4(0, 0, 0, 0);
                    } else if (PRO == 2) {
                        const float sg = fminf(fmaxf(pf[0], 3.0590232050182579e-07f), 3269017.3724721107f);   // exp(-15), exp(15)
                        const __half d0 = __float2half_rn(pf[1] * sg);
                        const uint32_t w[8] = {(uint32_t)pv0.x, (uint32_t)pv0.y, (uint32_t)pv0.z, (uint32_t)pv0.w,
                                               (uint32_t)pv1.x, (uint32_t)pv1.y, (uint32_t)pv1.z, (uint32_t)pv1.w};
                        uint32_t o[8];
                        o[0] = (uint32_t)__half_as_ushort(d0) | (w[0] << 16);
#pragma unroll
                        for (int e = 1; e < 8; ++e) o[e] = (w[e - 1] >> 16) | (w[e] << 16);
                        v0 = make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]);
                        v1 = make_int4((int)o[4], (int)o[5], (int)o[6], (int)o[7]);
                    } else {
                        // density head: column 0 carries trunc_exp' (activation.py:15-18) on top of whatever arrived for h[0]
                        const float sg = fminf(fmaxf(pf[0], 3.0590232050182579e-07f), 3269017.3724721107f);   // exp(-15), exp(15)
                        const float g0 = __half2float(__ushort_as_half((unsigned short)((uint32_t)pv0.x & 0xffffu)));
                        const __half d0 = __float2half_rn(pf[1] * sg + g0);
                        v0 = make_int4((int)(((uint32_t)pv0.x & 0xffff0000u) | (uint32_t)__half_as_ushort(d0)), pv0.y, pv0.z, pv0.w);
                        v1 = pv1;
                    }
                    const uint32_t r8[8] = {(uint32_t)v0.x, (uint32_t)v0.y, (uint32_t)v0.z, (uint32_t)v0.w,
                                            (uint32_t)v1.x, (uint32_t)v1.y, (uint32_t)v1.z, (uint32_t)v1.w};
                    tmem_st8(a_t, r8);
                    uint8_t* g_tile = slot_g(s, 0);
                    *reinterpret_cast<int4*>(g_tile + 0 * 2048 + r_in_tile * 16) = v0;
                    *reinterpret_cast<int4*>(g_tile + 1 * 2048 + r_in_tile * 16) = v1;
                    tc_wait_st();
                }
                if (s == 0 && r_in_tile == 0) ENERF_TRACE(4000 + L);
                fence_proxy_async_smem();
                tc_fence_before();
                warp_arrive(&a_ready[2 * s + (L == NH ? 1 : 0)], lane);      // the next stage is forward (L < NH) or the first backward stage
                if (s == 0 && r_in_tile == 0) ENERF_TRACE(5000 + L);
                if (L == NH && hf == 0 && j + NSLOTS < my_tiles) fetch(((size_t)blockIdx.x + (size_t)(j + NSLOTS) * gridDim.x) * kTile + r_in_tile);
            }
            // ---- E_k, k = 1 .. S-1: g = D * relu'(h_{NH-(k-1)}) -> TMEM A + G tile
#pragma unroll
            for (int k = 1; k < S; ++k) {
                mbar_wait(&d_full[s], (tb + (uint32_t)(NH + k)) & 1u);
                tc_fence_after();
                if (s == 0 && r_in_tile == 0) ENERF_TRACE(3000 + NH + k);
                uint8_t* g_tile = slot_g(s, k);
#pragma unroll
                for (int h = 0; h < CW / 32; ++h) {
                    uint32_t acc[32];
                    tmem_ld32(d_t + hf * CW + h * 32, acc);
                    tc_wait_ld();
                    uint32_t p[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        p[e] = pack2(__uint_as_float(acc[2 * e]), __uint_as_float(acc[2 * e + 1])) & relu_mask_from_bits(mbits[NH - (k - 1)][h], e);
                    tmem_st16(a_t + hf * (CW / 2) + h * 16, p);
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        *reinterpret_cast<int4*>(g_tile + (hf * (CW / 8) + h * 4 + v) * 2048 + r_in_tile * 16) =
                            make_int4((int)p[4 * v], (int)p[4 * v + 1], (int)p[4 * v + 2], (int)p[4 * v + 3]);
                }
                if (s == 0 && r_in_tile == 0) ENERF_TRACE(4000 + NH + k);
                tc_wait_st();
                fence_proxy_async_smem();
                tc_fence_before();
                warp_arrive(&a_ready[2 * s + 1], lane);
                if (s == 0 && r_in_tile == 0) ENERF_TRACE(5000 + NH + k);
            }
            // ---- E_S: dx
            mbar_wait(&d_full[s], (tb + (uint32_t)(T - 1)) & 1u);
            tc_fence_after();
            if (s == 0 && r_in_tile == 0) ENERF_TRACE(3000 + T);
            {
                // dx: 32 columns; with CH == 2 each thread of the row takes 16 of them
                constexpr int DXW = 32 / CH;
                uint32_t acc[DXW];
                if (CH == 1) {
                    uint32_t a0[16], a1[16];
                    tmem_ld16(d_t, a0);
                    tmem_ld16(d_t + 16, a1);
                    tc_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) { acc[e % DXW] = a0[e]; }
#pragma unroll
                    for (int e = 0; e < 16; ++e) { acc[(16 + e) % DXW] = a1[e]; }
                } else {
                    uint32_t a0[16];
                    tmem_ld16(d_t + hf * 16, a0);
                    tc_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc[e % DXW] = a0[e];
                }
                tc_fence_before();
                warp_arrive(&a_ready[2 * s], lane);              // accumulator read: the slot can start its next tile
                if (grad_inputs) {
                    uint32_t p[DXW / 2];
#pragma unroll
                    for (int e = 0; e < DXW / 2; ++e) p[e] = pack2(__uint_as_float(acc[2 * e]), __uint_as_float(acc[2 * e + 1]));
                    int4* dst = reinterpret_cast<int4*>(grad_inputs + row * in_dim + hf * DXW);
#pragma unroll
                    for (int v = 0; v < DXW / 8; ++v) dst[v] = make_int4((int)p[4 * v], (int)p[4 * v + 1], (int)p[4 * v + 2], (int)p[4 * v + 3]);
                }
            }
        }

        // ---- flush the weight-gradient accumulators (slot 0's four warps; M = 64 -> lanes 0..15 of each quarter)
        if (s == 0 && hf == 0 && my_tiles > 0) {
            mbar_wait(flush_bar, 0);
            tc_fence_after();
            const int m = q * 16 + lane;
            const uint32_t base = tmem0 + lane_sel;
            float* dW0 = dW;
            float* dWh = dW + kW * in_dim;
            float* dWl = dWh + (size_t)NH * kW * kW;
            {
                uint32_t acc[16];
                tmem_ld16(base + kAccLast, acc);
                tc_wait_ld();
                if (lane < 16)
#pragma unroll
                    for (int nn = 0; nn < 16; ++nn) atomicAdd(dWl + nn * kW + m, __uint_as_float(acc[nn]));
            }
            for (int jj = 0; jj < NH; ++jj)
                for (int c = 0; c < 4; ++c) {
                    uint32_t acc[16];
                    tmem_ld16(base + kAccHid + jj * 64 + c * 16, acc);
                    tc_wait_ld();
                    if (lane < 16) {
                        float* dst = dWh + (size_t)jj * kW * kW + (size_t)m * kW + c * 16;
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            red_add_v4(dst + 4 * v, __uint_as_float(acc[4 * v]), __uint_as_float(acc[4 * v + 1]), __uint_as_float(acc[4 * v + 2]),
                                       __uint_as_float(acc[4 * v + 3]));
                    }
                }
            for (int c = 0; c < in_dim / 16; ++c) {
                uint32_t acc[16];
                tmem_ld16(base + kAcc0 + c * 16, acc);
                tc_wait_ld();
                if (lane < 16) {
                    float* dst = dW0 + (size_t)m * in_dim + c * 16;
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        red_add_v4(dst + 4 * v, __uint_as_float(acc[4 * v]), __uint_as_float(acc[4 * v + 1]), __uint_as_float(acc[4 * v + 2]),
                                   __uint_as_float(acc[4 * v + 3]));
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, kCols);
}

template <int NSLOTS, int NH, int PRO, bool XA>
static int launch_bwd_rc_n(const TmaDesc& tx, const __half* grad, const __half* W, __half* grad_inputs, float* dW, uint32_t B, ProArgs pro,
                           cudaStream_t st, const char* name) {
    constexpr bool GD = false;      // second G tile + early dgrad commit: measured +-0 %
    // two epilogue threads per row (8 warps per slot, 72 registers): since the ReLU masks live in registers this pays for the 3-layer
    // nets — colour-net backward 0.339 -> 0.317 ms — and not for the 2-layer ones (sigma-net 0.251 -> 0.253)
    constexpr int CH = (NH == 2) ? 2 : 1;
    constexpr int NI = 1;           // second issuing warp: +-0 % in round 1; with two epilogue threads per row 0.374 vs 0.316 ms (worse)
    constexpr size_t kSlot = (XA ? 0 : 2 * (size_t)kTile * 32 * 2) + (size_t)(NH + 2 + (GD ? 1 : 0)) * kGBytes;
    size_t smem = 1024 + NSLOTS * kSlot + 32 * 128 + (size_t)NH * 8192 + 2048 + (5 * NSLOTS + 1) * 8 + 16;
    if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM: it allocates all 512 TMEM columns
    if (smem > 227 * 1024) { set_error("%s: shared-memory budget exceeded", name); return -2; }
    static bool configured = false;
    if (!configured) {
        ENERF_CUDA(cudaFuncSetAttribute(k_tc_bwd_rc<NSLOTS, NH, PRO, GD, CH, XA, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), name);
        configured = true;
    }
    const uint32_t n_tiles = B / kTile;
    const uint32_t grid = tc_grid(n_tiles);
    k_tc_bwd_rc<NSLOTS, NH, PRO, GD, CH, XA, NI><<<grid, 32 * NI + NSLOTS * 128 * CH, smem, st>>>(tx, grad, W, grad_inputs, dW, n_tiles, B, pro);
    ENERF_CHECK_LAUNCH(name);
    return 0;
}

// backward from the network input alone (hidden activations recomputed per tile).
// Measured on B200 (3.29 M samples).  2-layer nets: 0.255 ms with 3 slots x 4 epilogue warps (0.28 with 2 x 8; 0.33 with 4 slots and
// the input tile aliased: 96 registers, spills; again with the ReLU masks in registers — 56-64 bytes of spills left at 96 registers —
// the sigma-net backward takes 0.290 ms with 4 slots against 0.249 with 3).  3-layer nets: 0.343 ms with 3 slots, input tile aliased onto dead activation
// buffers (XA); 0.355 with 2 slots x 8 epilogue warps; 0.386 with 2 x 4; with the masks in registers 3 slots x 8 warps: 0.316 (shipped).  1-layer nets (torch topology sigma-net): 3 slots.
template <int PRO>
static int launch_bwd_rc(const __half* grad, const __half* x, const __half* W, __half* grad_inputs, float* dW, uint32_t B, int in_dim, int n_hidden_mm,
                         ProArgs pro, cudaStream_t st, const char* name) {
    if (in_dim != 32 || n_hidden_mm < 0 || n_hidden_mm > 2 || (uint64_t)B >= (1ull << 31)) {
        set_error("%s: recomputation (NULL forward_buffer) takes 32 inputs and 1 to 3 layers", name);
        return -2;
    }
    TmaDesc tx;
    if (!make_tmap_rows(&tx, x, B, 32, kTile)) { set_error("%s: inputs must be 16-byte aligned (TMA)", name); return -2; }
    if (n_hidden_mm == 0) {
        if constexpr (PRO == 0 || PRO == 3) return launch_bwd_rc_n<3, 0, PRO, false>(tx, grad, W, grad_inputs, dW, B, pro, st, name);
        set_error("%s: a 1-layer network has no fused field prologue", name);
        return -2;
    }
    if (n_hidden_mm == 1) return launch_bwd_rc_n<3, 1, PRO, false>(tx, grad, W, grad_inputs, dW, B, pro, st, name);
    if constexpr (PRO != 3) return launch_bwd_rc_n<3, 2, PRO, true>(tx, grad, W, grad_inputs, dW, B, pro, st, name);
    set_error("%s: density heads take 1- and 2-layer networks", name);
    return -2;
}

int tc_backward(const __half* grad, const __half* x, const __half* W, const __half* fwd_buf, __half* bwd_buf, __half* grad_inputs, float* dW,
                uint32_t B, int in_dim, int n_hidden_mm, cudaStream_t st) {
    ProArgs none = {nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr};
    if (!fwd_buf) return launch_bwd_rc<0>(grad, x, W, grad_inputs, dW, B, in_dim, n_hidden_mm, none, st, "ffmlp_backward");
    return launch_bwd_tma<0>(grad, x, W, fwd_buf, bwd_buf, grad_inputs, dW, B, in_dim, n_hidden_mm, none, st, "ffmlp_backward");
}
int tc_backward_rgb(const float* g_rgb, const float* rgb, int n_ch, const __half* cin, const __half* W, const __half* fwd_buf, __half* dcin, float* dW,
                    uint32_t B, int n_hidden_mm, const int32_t* n_rows_dev, cudaStream_t st) {
    ProArgs p = {g_rgb, rgb, n_ch, nullptr, nullptr, nullptr, n_rows_dev};
    if (!fwd_buf) return launch_bwd_rc<1>(nullptr, cin, W, dcin, dW, B, 32, n_hidden_mm, p, st, "field_color_backward");
    return launch_bwd_tma<1>(nullptr, cin, W, fwd_buf, nullptr, dcin, dW, B, 32, n_hidden_mm, p, st, "field_color_backward");
}
int tc_backward_sigma(const float* g_sigma, const float* sigma, const __half* dcin, const __half* feat, const __half* W, const __half* fwd_buf,
                      __half* dfeat, float* dW, uint32_t B, int n_hidden_mm, cudaStream_t st) {
    ProArgs p = {nullptr, nullptr, 0, g_sigma, sigma, dcin, nullptr};
    if (!fwd_buf) return launch_bwd_rc<2>(nullptr, feat, W, dfeat, dW, B, 32, n_hidden_mm, p, st, "field_sigma_backward");
    return launch_bwd_tma<2>(nullptr, feat, W, fwd_buf, nullptr, dfeat, dW, B, 32, n_hidden_mm, p, st, "field_sigma_backward");
}
// density head (always recomputing): g_sigma [B] (may be NULL) and g_h [B,16] fp16 (may be NULL) -> dfeat, dW
int tc_backward_density(const float* g_sigma, const float* sigma, const __half* g_h, const __half* feat, const __half* W, __half* dfeat, float* dW,
                        uint32_t B, int n_hidden_mm, cudaStream_t st) {
    ProArgs p = {nullptr, nullptr, 0, g_sigma, sigma, nullptr, nullptr};
    return launch_bwd_rc<3>(g_h, feat, W, dfeat, dW, B, 32, n_hidden_mm, p, st, "field_density_backward");
}

// ------------------------------------------------------------------------------------------------
// Colour-net input rows of the torch-topology field (nerf/network.py:171-199): for the samples selected by the `weights > 1e-4`
// mask, row i = [SH_4(fp16(dir of sample idx[i])) * sh_scale (16) | h[idx[i], 1:16] (15) | 0] — the masked gather `d[mask]`,
// `geo_feat[mask]`, the SH encoder, `torch.cat` and the 31 -> 32 padding in one pass.  A ray's direction is shared by its
// `dir_div` consecutive samples (dirs [B / dir_div, 3]); rows n .. n_pad-1 are zero (tile padding).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_color_inputs(const float* __restrict__ dirs, uint32_t dir_div, const __half* __restrict__ h, const int32_t* __restrict__ idx, uint32_t n, uint32_t n_pad,
               float sh_scale, __half* __restrict__ cin, const int32_t* __restrict__ n_dev) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) {                                   // row count on the device: n = capacity; rows up to the next tile boundary are zero-filled
        n = min(n, (uint32_t)max(*n_dev, 0));
        n_pad = min(n_pad, (n + (uint32_t)kTile - 1u) / (uint32_t)kTile * (uint32_t)kTile);
    }
    if (i >= n_pad) return;
    int4* o = reinterpret_cast<int4*>(cin + (size_t)i * 32);
    if (i >= n) {
        o[0] = o[1] = o[2] = o[3] = make_int4(0, 0, 0, 0);
        return;
    }
    const uint32_t j = (uint32_t)idx[i];
    const float* d = dirs + (size_t)(j / dir_div) * 3;
    float sh[16];
    sh_deg4(f16_round(__ldg(d)), f16_round(__ldg(d + 1)), f16_round(__ldg(d + 2)), sh);
    const int4 a = __ldg(reinterpret_cast<const int4*>(h + (size_t)j * 16)), b = __ldg(reinterpret_cast<const int4*>(h + (size_t)j * 16) + 1);
    const uint32_t w[8] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w, (uint32_t)b.x, (uint32_t)b.y, (uint32_t)b.z, (uint32_t)b.w};
    uint32_t p[16];
#pragma unroll
    for (int e = 0; e < 8; ++e) p[e] = pack2(sh[2 * e] * sh_scale, sh[2 * e + 1] * sh_scale);
#pragma unroll
    for (int e = 0; e < 7; ++e) p[8 + e] = (w[e] >> 16) | (w[e + 1] << 16);      // h[1+2e], h[2+2e]
    p[15] = w[7] >> 16;                                                          // h[15], 0
#pragma unroll
    for (int v = 0; v < 4; ++v) o[v] = make_int4((int)p[4 * v], (int)p[4 * v + 1], (int)p[4 * v + 2], (int)p[4 * v + 3]);
}
// backward of the above w.r.t. h: g_h[idx[i], 1:16] = dcin[i, 16:31], g_h[idx[i], 0] = 0 (g_h zero-initialised by the caller)
__global__ void __launch_bounds__(256)
k_color_inputs_bwd(const __half* __restrict__ dcin, const int32_t* __restrict__ idx, uint32_t n, __half* __restrict__ g_h,
                   const int32_t* __restrict__ n_dev) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) n = min(n, (uint32_t)max(*n_dev, 0));
    if (i >= n) return;
    const int4* src = reinterpret_cast<const int4*>(dcin + (size_t)i * 32 + 16);
    const int4 a = __ldg(src), b = __ldg(src + 1);
    const uint32_t w[8] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w, (uint32_t)b.x, (uint32_t)b.y, (uint32_t)b.z, (uint32_t)b.w};
    uint32_t o[8];
    o[0] = w[0] << 16;
#pragma unroll
    for (int e = 1; e < 8; ++e) o[e] = (w[e - 1] >> 16) | (w[e] << 16);
    int4* dst = reinterpret_cast<int4*>(g_h + (size_t)idx[i] * 16);
    dst[0] = make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]);
    dst[1] = make_int4((int)o[4], (int)o[5], (int)o[6], (int)o[7]);
}

int tc_color_inputs(const float* dirs, uint32_t dir_div, const __half* h, const int32_t* idx, uint32_t n, uint32_t n_pad, float sh_scale, __half* cin,
                    const int32_t* n_dev, cudaStream_t st) {
    if (n_pad == 0) return 0;
    k_color_inputs<<<ceil_div(n_pad, 256u), 256, 0, st>>>(dirs, dir_div, h, idx, n, n_pad, sh_scale, cin, n_dev);
    ENERF_CHECK_LAUNCH("field_color_inputs");
    return 0;
}
int tc_color_inputs_backward(const __half* dcin, const int32_t* idx, uint32_t n, __half* g_h, const int32_t* n_dev, cudaStream_t st) {
    if (n == 0) return 0;
    k_color_inputs_bwd<<<ceil_div(n, 256u), 256, 0, st>>>(dcin, idx, n, g_h, n_dev);
    ENERF_CHECK_LAUNCH("field_color_inputs_backward");
    return 0;
}

}  // namespace tcm
}  // namespace enerf
