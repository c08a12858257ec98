// tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a) used by the fused-MLP kernels.
//
// Conventions
//   * TMEM address = (lane << 16) | column; an allocation returns (lane 0, first column).
//   * A warp may only touch TMEM lanes 32*(warp_id % 4) .. +31 ("its quarter"); with the
//     32x32b shapes thread `l` of the warp owns lane 32*(warp_id%4) + l and N consecutive columns.
//   * shared-memory operand tiles use the K-major / MN-major INTERLEAVE (no-swizzle) canonical
//     layout: 16-byte chunks; tile[c][r] (chunk c along the contiguous-in-register dimension,
//     row r) lives at byte  c * (rows*16) + r * 16.  Read as a K-major operand (rows = M or N,
//     chunks = K) the descriptor has LBO = rows*16, SBO = 128; read as an MN-major operand
//     (chunks = M or N, rows = K) it has SBO(chunk stride) = rows*16 and LBO(8-row groups) = 128.
#pragma once
#include "common.cuh"

namespace enerf {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
// non-blocking probe (no suspend-time hint): used by the MMA thread to serve whichever slot is ready
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// the same wait with a suspend-time hint (ns): the warp sleeps in the barrier unit (NANOSLEEP.SYNCS, woken by the phase completion)
// instead of coming back every ~100 cycles to poll.  For kernels whose waiting warps share their schedulers with busy ones
// (field_infer.cu: ncu r2_49 counted a third of the issued instructions in the polling loops of the thirteen waiting MLP warps;
// removing them frees issue slots but did not change the kernel's time, 0.511 ms either way).
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 20000u) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
            : "memory");
    } while (!done);
}

// one lane of the (converged) warp; the same lane every time for the full mask.  MMA-issuing warps run their
// control flow warp-uniformly and wrap only the issue itself in `if (elect_one())`: under a divergent
// `if (lane == 0)` ptxas cannot prove the descriptor operands uniform and emits an ELECT / R2UR waterfall loop
// in front of every UTCHMMA (~150 cycles per MMA on the single issuing thread — measured as THE limiter of the
// fused-MLP kernels).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred)::"memory");
    return pred != 0;
}

// ---- TMEM allocation (one full warp) ------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- fences / commit ----------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// all prior tcgen05.mma of this thread complete -> arrive(1) on the mbarrier
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMA (cp.async.bulk.tensor) -------------------------------------------------------------------
// expect `bytes` of async-proxy writes on the barrier and count one arrival (the issuing thread's)
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 2-D tiled load global -> shared; c0 = element index in the row, c1 = row index; completion is signalled on `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, int32_t c0, int32_t c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// 2-D tiled store shared -> global (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N store groups of this thread have finished READING shared memory
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// named barrier over `nthreads` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- descriptors --------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE; lbo/sbo in bytes (see the layout note on top)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version for sm_100
    return d;                 // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}
// shared-memory matrix descriptor for a TMA-swizzled tile (rows of `swizzle_bytes` = 128 / 64 / 32 bytes, densely
// packed, 8-row swizzle atoms; the tile base must be aligned to 8*swizzle_bytes).  Serves both readings of such a
// tile: K-major (rows = M/N, the row bytes = K) and MN-major (rows = K, the row bytes = M/N); in both the stride
// between 8-row atoms is SBO = 8*swizzle_bytes and LBO is not used (one atom wide).
__device__ __forceinline__ uint64_t smem_desc_sw(uint32_t saddr, uint32_t swizzle_bytes) {
    const uint64_t layout = (swizzle_bytes == 128) ? 2ull : (swizzle_bytes == 64) ? 4ull : 6ull;
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                                       // LBO (ignored for one-atom-wide swizzled tiles)
    d |= (uint64_t)(((8u * swizzle_bytes) >> 4) & 0x3FFFu) << 32;  // SBO
    d |= (uint64_t)1 << 46;                                       // descriptor version for sm_100
    d |= layout << 61;
    return d;
}
// byte offset of 16-byte chunk `c` of row `r` inside a TMA-swizzled tile with dense rows of `swizzle_bytes`:
// Swizzle<B,4,3> XORs address bits [7,7+B) into bits [4,4+B), B = log2(swizzle_bytes/16)
__device__ __forceinline__ uint32_t sw_off(uint32_t r, uint32_t c, uint32_t swizzle_bytes) {
    const uint32_t lin = r * swizzle_bytes + (c << 4);
    return lin ^ (((lin >> 7) & (swizzle_bytes / 16u - 1u)) << 4);
}
// instruction descriptor for kind::f16: fp16 A/B, fp32 accumulate
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                      // D format: F32
           | (0u << 7) | (0u << 10)       // A, B format: F16
           | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
           | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- MMA issue (one thread) ------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// ---- TMEM <-> registers (32x32b: one lane per thread, N consecutive 32-bit columns) -----------------
#define ENERF_R8(a, o) "=r"(a[o + 0]), "=r"(a[o + 1]), "=r"(a[o + 2]), "=r"(a[o + 3]), "=r"(a[o + 4]), "=r"(a[o + 5]), "=r"(a[o + 6]), "=r"(a[o + 7])
#define ENERF_W8(a, o) "r"(a[o + 0]), "r"(a[o + 1]), "r"(a[o + 2]), "r"(a[o + 3]), "r"(a[o + 4]), "r"(a[o + 5]), "r"(a[o + 6]), "r"(a[o + 7])

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : ENERF_R8(r, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : ENERF_R8(r, 0), ENERF_R8(r, 8) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : ENERF_R8(r, 0), ENERF_R8(r, 8), ENERF_R8(r, 16), ENERF_R8(r, 24) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), ENERF_W8(r, 0) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), ENERF_W8(r, 0), ENERF_W8(r, 8) : "memory");
}

// ---- warp-cooperative row <-> coalesced transposition through swizzled shared memory ---------------
// A warp owns 32 consecutive rows of NV 16-byte chunks (row-major, contiguous in global memory).
//   coalesced form: piece i of lane l is chunk (i*32 + l) of the 32*NV-chunk block  -> every global
//                   access of the warp covers 512 contiguous bytes (4 L1 wavefronts instead of 32)
//   row form      : lane l holds the NV chunks of row l (what the TMEM epilogue works on)
// `stage` is a warp-private 32*NV*16-byte region; the XOR swizzle makes both access patterns
// bank-conflict free.
template <int NV>
__device__ __forceinline__ uint32_t swz(int row, int c) {
    constexpr int SH = (NV == 8) ? 0 : (NV == 4) ? 1 : 2;
    return (uint32_t)(row * NV + (c ^ ((row >> SH) & (NV - 1)))) * 16u;
}
template <int NV>
__device__ __forceinline__ void ld_coalesced(int4 (&co)[NV], const int4* __restrict__ g, int lane) {
#pragma unroll
    for (int i = 0; i < NV; ++i) co[i] = __ldg(g + i * 32 + lane);
}
template <int NV>
__device__ __forceinline__ void st_coalesced(int4* __restrict__ g, const int4 (&co)[NV], int lane) {
#pragma unroll
    for (int i = 0; i < NV; ++i) g[i * 32 + lane] = co[i];
}
template <int NV>
__device__ __forceinline__ void coalesced_to_rows(const int4 (&co)[NV], int4 (&rows)[NV], uint8_t* stage, int lane) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int idx = i * 32 + lane;
        *reinterpret_cast<int4*>(stage + swz<NV>(idx / NV, idx % NV)) = co[i];
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NV; ++c) rows[c] = *reinterpret_cast<const int4*>(stage + swz<NV>(lane, c));
    __syncwarp();
}
template <int NV>
__device__ __forceinline__ void rows_to_coalesced(const int4 (&rows)[NV], int4 (&co)[NV], uint8_t* stage, int lane) {
#pragma unroll
    for (int c = 0; c < NV; ++c) *reinterpret_cast<int4*>(stage + swz<NV>(lane, c)) = rows[c];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int idx = i * 32 + lane;
        co[i] = *reinterpret_cast<const int4*>(stage + swz<NV>(idx / NV, idx % NV));
    }
    __syncwarp();
}
// rows (this lane's row) -> global, coalesced
template <int NV>
__device__ __forceinline__ void store_rows(int4* __restrict__ g_warp, const int4 (&rows)[NV], uint8_t* stage, int lane) {
    int4 co[NV];
    rows_to_coalesced<NV>(rows, co, stage, lane);
    st_coalesced<NV>(g_warp, co, lane);
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float relu(float v) { return fmaxf(v, 0.0f); }
// fp16(relu(a)), fp16(relu(b)) packed: round first, clamp the packed pair (one F2FP + one HMNMX2; rounding is monotonic and
// sign-preserving, so this equals rounding relu(a), relu(b))
__device__ __forceinline__ uint32_t pack2_relu(float a, float b) {
    const __half2 h = __hmax2(__floats2half2_rn(a, b), __floats2half2_rn(0.f, 0.f));
    return *reinterpret_cast<const uint32_t*>(&h);
}
// one mbarrier arrival per warp (barrier count = number of warps): 32 same-address arrivals per warp serialise (~250 cycles
// for 128 threads, measured); the warp's lanes are ordered by __syncwarp first
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// ---- pieces shared by the fused-MLP kernels (ffmlp_tc.cu) and the fused inference field (field_infer.cu) ----
// copy a row-major [rows, K] fp16 matrix into the canonical chunked layout tile[c][r] (16-B chunks)
__device__ __forceinline__ void stage_matrix(uint8_t* dst, const __half* __restrict__ src, int rows, int K, int tid, int nthreads) {
    const int cpr = K >> 3;   // chunks per row
    for (int i = tid; i < rows * cpr; i += nthreads) {
        const int r = i / cpr, c = i - r * cpr;
        *reinterpret_cast<int4*>(dst + (size_t)c * rows * 16 + (size_t)r * 16) = __ldg(reinterpret_cast<const int4*>(src) + i);
    }
}

// real spherical harmonics up to l = 3 of (x,y,z) — the basis of shencoder.cu:51-69, fp32
__device__ __forceinline__ void sh_deg4(float x, float y, float z, float (&o)[16]) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;                         // 1/(2 sqrt(pi))
    o[1] = -0.48860251190291987f * y;                    // sqrt(3/(4 pi))
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;                     // sqrt(15/(4 pi))
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;   // sqrt(5/(16 pi)) (3 z^2 - 1)
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * (x2 - y2);             // sqrt(15/(16 pi))
    o[9] = 0.59004358992664352f * y * (y2 - 3.0f * x2);  // sqrt(35/(32 pi))
    o[10] = 2.8906114426405538f * xy * z;                // sqrt(105/(4 pi))
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);   // sqrt(21/(32 pi))
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);    // sqrt(7/(16 pi))
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);         // sqrt(105/(16 pi))
    o[15] = 0.59004358992664352f * x * (3.0f * y2 - x2);
}
__device__ __forceinline__ float f16_round(float v) { return __half2float(__float2half_rn(v)); }

}  // namespace tc
}  // namespace enerf
