// Fully-fused MLP (forward / inference / backward) — generic-shape path on mma.sync.
//
// Replaces the reference's `_ffmlp` extension (ffmlp/src/ffmlp.h:8-13).  Semantics follow
// ffmlp/src/ffmlp.cu: y = W_last * act(W_{n-1} * ... act(W_0 * x)), no bias, weights row-major
// [out,in] (:631-634); forward_buffer[k] = post-activation output of hidden matmul k (:367-383);
// backward_buffer[j] = dL/d(pre-activation) of hidden layer (num_layers-1-j) (:431-507);
// dW_k = dH_k^T * H_{k-1} (:800-877); dX = dH_0 * W_0 (:880-887).
//
// Differences by design: fp32 accumulation everywhere (the reference accumulates in fp16);
// every warp owns 16*MT complete rows of the batch tile, so layers chain with __syncwarp only
// (no block barriers); CTAs are persistent over row tiles with all weights resident in shared
// memory; weight gradients are accumulated in registers over the CTA's whole row range and
// reduced once per CTA with fp32 atomics (the reference runs CUTLASS split-K GEMMs plus
// reduction kernels on side streams).
#include "ffmlp_mma.cuh"

namespace enerf {
namespace mlp {

static constexpr int kWarps = 4;

template <int WIDTH> struct Shape {
    static constexpr int MT = (WIDTH >= 256) ? 1 : 2;   // m16 tiles per warp
    static constexpr int ROWS_W = 16 * MT;               // rows per warp
    static constexpr int ROWS = ROWS_W * kWarps;         // rows per CTA tile
    static constexpr int NT = WIDTH / 8;
};

__device__ __forceinline__ void cta_copy_rows(__half* __restrict__ s, int lds_, const __half* __restrict__ g, int rows, int cols) {
    const int vec_per_row = cols >> 3, total = rows * vec_per_row;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int r = i / vec_per_row, c = (i - r * vec_per_row) << 3;
        *reinterpret_cast<int4*>(s + r * lds_ + c) = __ldg(reinterpret_cast<const int4*>(g + (size_t)r * cols + c));
    }
}
// s[i][o] = g[o][i]   (g row-major [outs, ins])
__device__ __forceinline__ void cta_copy_transposed(__half* __restrict__ s, int lds_, const __half* __restrict__ g, int outs, int ins) {
    const int total = outs * ins;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int o = i / ins, k = i - o * ins;
        s[k * lds_ + o] = g[i];
    }
}

// ------------------------------------------------------------------------------------------
// forward / inference
// ------------------------------------------------------------------------------------------
template <int WIDTH>
__global__ void __launch_bounds__(kWarps * 32)
k_mlp_fwd(const __half* __restrict__ in, const __half* __restrict__ W, __half* __restrict__ fwd_buf,
          __half* __restrict__ out, uint32_t B, int in_dim, int n_hidden_mm, uint32_t act, uint32_t out_act) {
    using S = Shape<WIDTH>;
    constexpr int MT = S::MT, NT = S::NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sm = reinterpret_cast<__half*>(smem_raw);

    const int ld0 = in_dim + 8, ldh = WIDTH + 8;
    const int AS = (in_dim > WIDTH ? in_dim : WIDTH) + 8;
    __half* W0s = sm;
    __half* Whs = W0s + WIDTH * ld0;
    __half* Wls = Whs + n_hidden_mm * WIDTH * ldh;
    __half* acts = Wls + 16 * ldh;

    cta_copy_rows(W0s, ld0, W, WIDTH, in_dim);
    for (int j = 0; j < n_hidden_mm; ++j)
        cta_copy_rows(Whs + j * WIDTH * ldh, ldh, W + WIDTH * in_dim + j * WIDTH * WIDTH, WIDTH, WIDTH);
    cta_copy_rows(Wls, ldh, W + WIDTH * in_dim + n_hidden_mm * WIDTH * WIDTH, 16, WIDTH);
    __syncthreads();

    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, g = lane >> 2, tg = lane & 3u;
    __half* act_w = acts + warp * S::ROWS_W * AS;
    const uint32_t n_tiles = B / S::ROWS;

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t row0 = (size_t)tile * S::ROWS + warp * S::ROWS_W;
        __syncwarp();
        warp_copy_g2s(act_w, AS, in + row0 * in_dim, in_dim, S::ROWS_W, in_dim);
        __syncwarp();

        for (int layer = 0; layer <= n_hidden_mm; ++layer) {
            float acc[MT][NT][4];
            zero_acc<MT, NT>(acc);
            if (layer == 0) warp_gemm<MT, NT>(acc, act_w, AS, W0s, ld0, in_dim, NT);
            else warp_gemm<MT, NT>(acc, act_w, AS, Whs + (layer - 1) * WIDTH * ldh, ldh, WIDTH, NT);
            __syncwarp();
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    __half* p = act_w + (mt * 16 + g) * AS + nt * 8 + 2 * tg;
                    *reinterpret_cast<uint32_t*>(p) = pack_half2(act_fwd(act, acc[mt][nt][0]), act_fwd(act, acc[mt][nt][1]));
                    *reinterpret_cast<uint32_t*>(p + 8 * AS) = pack_half2(act_fwd(act, acc[mt][nt][2]), act_fwd(act, acc[mt][nt][3]));
                }
            __syncwarp();
            if (fwd_buf) warp_copy_s2g(fwd_buf + ((size_t)layer * B + row0) * WIDTH, WIDTH, act_w, AS, S::ROWS_W, WIDTH);
        }
        // output layer: 16 (padded) neurons
        float acc2[MT][2][4];
        zero_acc<MT, 2>(acc2);
        warp_gemm<MT, 2>(acc2, act_w, AS, Wls, ldh, WIDTH, 2);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                __half* p = out + (row0 + mt * 16 + g) * 16 + nt * 8 + 2 * tg;
                *reinterpret_cast<uint32_t*>(p) = pack_half2(act_fwd(out_act, acc2[mt][nt][0]), act_fwd(out_act, acc2[mt][nt][1]));
                *reinterpret_cast<uint32_t*>(p + 8 * 16) = pack_half2(act_fwd(out_act, acc2[mt][nt][2]), act_fwd(out_act, acc2[mt][nt][3]));
            }
    }
}

template <int WIDTH>
static size_t fwd_smem_bytes(int in_dim, int n_hidden_mm) {
    using S = Shape<WIDTH>;
    const int AS = (in_dim > WIDTH ? in_dim : WIDTH) + 8;
    return sizeof(__half) * ((size_t)WIDTH * (in_dim + 8) + (size_t)n_hidden_mm * WIDTH * (WIDTH + 8) + 16 * (WIDTH + 8) + (size_t)S::ROWS * AS);
}

// ------------------------------------------------------------------------------------------
// backward: activation gradients (dgrad chain)
// ------------------------------------------------------------------------------------------
template <int WIDTH>
__global__ void __launch_bounds__(kWarps * 32)
k_mlp_bwd(const __half* __restrict__ grad, const __half* __restrict__ W, const __half* __restrict__ fwd_buf,
          __half* __restrict__ bwd_buf, __half* __restrict__ grad_inputs, uint32_t B, int in_dim, int n_hidden_mm,
          uint32_t act) {
    using S = Shape<WIDTH>;
    constexpr int MT = S::MT, NT = S::NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sm = reinterpret_cast<__half*>(smem_raw);

    const int ldl = 16 + 8, ldh = WIDTH + 8;
    const int AS = WIDTH + 8;
    __half* WTl = sm;                                   // [WIDTH][16]   (n = hidden, k = output)
    __half* WTh = WTl + WIDTH * ldl;                    // n_hidden_mm x [WIDTH(in)][WIDTH(out)]
    __half* WT0 = WTh + n_hidden_mm * WIDTH * ldh;      // [in_dim][WIDTH]  (only when grad_inputs)
    __half* acts = WT0 + (grad_inputs ? in_dim * ldh : 0);
    __half* fwds = acts + S::ROWS * AS;

    const __half* Wh = W + WIDTH * in_dim;
    const __half* Wl = Wh + n_hidden_mm * WIDTH * WIDTH;
    cta_copy_transposed(WTl, ldl, Wl, 16, WIDTH);
    for (int j = 0; j < n_hidden_mm; ++j) cta_copy_transposed(WTh + j * WIDTH * ldh, ldh, Wh + j * WIDTH * WIDTH, WIDTH, WIDTH);
    if (grad_inputs) cta_copy_transposed(WT0, ldh, W, WIDTH, in_dim);
    __syncthreads();

    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, g = lane >> 2, tg = lane & 3u;
    __half* act_w = acts + warp * S::ROWS_W * AS;
    __half* fwd_w = fwds + warp * S::ROWS_W * AS;
    const uint32_t n_tiles = B / S::ROWS;

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t row0 = (size_t)tile * S::ROWS + warp * S::ROWS_W;
        __syncwarp();
        warp_copy_g2s(act_w, AS, grad + row0 * 16, 16, S::ROWS_W, 16);

        // step s = 0: through the output layer; s = 1..n_hidden_mm: through hidden matmul (n_hidden_mm - s)
        for (int s = 0; s <= n_hidden_mm; ++s) {
            const int h_index = n_hidden_mm - s;   // forward activation whose transfer applies
            warp_copy_g2s(fwd_w, AS, fwd_buf + ((size_t)h_index * B + row0) * WIDTH, WIDTH, S::ROWS_W, WIDTH);
            __syncwarp();
            float acc[MT][NT][4];
            zero_acc<MT, NT>(acc);
            if (s == 0) warp_gemm<MT, NT>(acc, act_w, AS, WTl, ldl, 16, NT);
            else warp_gemm<MT, NT>(acc, act_w, AS, WTh + (n_hidden_mm - s) * WIDTH * ldh, ldh, WIDTH, NT);
            __syncwarp();
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int off = (mt * 16 + g) * AS + nt * 8 + 2 * tg;
                    const __half2 y0 = *reinterpret_cast<const __half2*>(fwd_w + off);
                    const __half2 y1 = *reinterpret_cast<const __half2*>(fwd_w + off + 8 * AS);
                    *reinterpret_cast<uint32_t*>(act_w + off) =
                        pack_half2(act_bwd(act, acc[mt][nt][0], __low2float(y0)), act_bwd(act, acc[mt][nt][1], __high2float(y0)));
                    *reinterpret_cast<uint32_t*>(act_w + off + 8 * AS) =
                        pack_half2(act_bwd(act, acc[mt][nt][2], __low2float(y1)), act_bwd(act, acc[mt][nt][3], __high2float(y1)));
                }
            __syncwarp();
            if (bwd_buf) warp_copy_s2g(bwd_buf + ((size_t)s * B + row0) * WIDTH, WIDTH, act_w, AS, S::ROWS_W, WIDTH);
        }

        if (grad_inputs) {
            for (int n0 = 0; n0 < in_dim; n0 += WIDTH) {
                const int cols = min(WIDTH, in_dim - n0);
                float acc[MT][NT][4];
                zero_acc<MT, NT>(acc);
                warp_gemm<MT, NT>(acc, act_w, AS, WT0 + n0 * ldh, ldh, WIDTH, cols / 8);
                __syncwarp();
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        if (nt * 8 < cols) {
                            const int off = (mt * 16 + g) * AS + nt * 8 + 2 * tg;
                            *reinterpret_cast<uint32_t*>(fwd_w + off) = pack_half2(acc[mt][nt][0], acc[mt][nt][1]);
                            *reinterpret_cast<uint32_t*>(fwd_w + off + 8 * AS) = pack_half2(acc[mt][nt][2], acc[mt][nt][3]);
                        }
                __syncwarp();
                warp_copy_s2g(grad_inputs + row0 * in_dim + n0, in_dim, fwd_w, AS, S::ROWS_W, cols);
            }
        }
    }
}

template <int WIDTH>
static size_t bwd_smem_bytes(int in_dim, int n_hidden_mm, bool grad_inputs) {
    using S = Shape<WIDTH>;
    return sizeof(__half) * ((size_t)WIDTH * 24 + (size_t)n_hidden_mm * WIDTH * (WIDTH + 8) + (grad_inputs ? (size_t)in_dim * (WIDTH + 8) : 0) +
                             2 * (size_t)S::ROWS * (WIDTH + 8));
}

// ------------------------------------------------------------------------------------------
// backward: weight gradients   dW[M,N] (+)= dY[B,M]^T * X[B,N]   (fp32 atomics into dW)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const __half* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

static constexpr int kWgRows = 64;  // batch rows per staged tile
__global__ void __launch_bounds__(kWarps * 32)
k_mlp_wgrad(const __half* __restrict__ dY, int ldY, const __half* __restrict__ X, int ldX, float* __restrict__ dW,
            uint32_t B, int M, int N) {
    __shared__ __align__(16) __half Ys[kWgRows][64 + 8];
    __shared__ __align__(16) __half Xs[kWgRows][64 + 8];
    const int m_base = blockIdx.y * 64, n_base = blockIdx.z * 64;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, g = lane >> 2, tg = lane & 3u;
    const int q = lane >> 3, r = lane & 7;
    const bool warp_active = (m_base + (int)warp * 16) < M;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const uint32_t n_blocks = B / kWgRows;
    for (uint32_t rb = blockIdx.x; rb < n_blocks; rb += gridDim.x) {
        const size_t row0 = (size_t)rb * kWgRows;
        __syncthreads();
        for (int i = threadIdx.x; i < kWgRows * 8; i += blockDim.x) {
            const int rr = i >> 3, c = (i & 7) << 3;
            int4 vy = make_int4(0, 0, 0, 0), vx = make_int4(0, 0, 0, 0);
            if (m_base + c < M) vy = __ldg(reinterpret_cast<const int4*>(dY + (row0 + rr) * ldY + m_base + c));
            if (n_base + c < N) vx = __ldg(reinterpret_cast<const int4*>(X + (row0 + rr) * ldX + n_base + c));
            *reinterpret_cast<int4*>(&Ys[rr][c]) = vy;
            *reinterpret_cast<int4*>(&Xs[rr][c]) = vx;
        }
        __syncthreads();
        if (warp_active) {
#pragma unroll
            for (int ks = 0; ks < kWgRows / 16; ++ks) {
                uint32_t a[4];
                ldmatrix_x4_trans(a, &Ys[ks * 16 + (q >> 1) * 8 + r][warp * 16 + (q & 1) * 8]);
#pragma unroll
                for (int np = 0; np < 4; ++np) {   // pairs of n-tiles
                    uint32_t b[4];
                    ldmatrix_x4_trans(b, &Xs[ks * 16 + (q & 1) * 8 + r][np * 16 + (q >> 1) * 8]);
                    mma_16816(acc[np * 2], a, b[0], b[1]);
                    mma_16816(acc[np * 2 + 1], a, b[2], b[3]);
                }
            }
        }
    }
    if (warp_active) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int m = m_base + warp * 16 + g, n = n_base + nt * 8 + 2 * tg;
            if (n < N) {
                if (m < M) { atomicAdd(dW + (size_t)m * N + n, acc[nt][0]); atomicAdd(dW + (size_t)m * N + n + 1, acc[nt][1]); }
                if (m + 8 < M) { atomicAdd(dW + (size_t)(m + 8) * N + n, acc[nt][2]); atomicAdd(dW + (size_t)(m + 8) * N + n + 1, acc[nt][3]); }
            }
        }
    }
}

__global__ void k_f32_to_f16(const float* __restrict__ src, __half* __restrict__ dst, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2half_rn(src[i]);
}

// ------------------------------------------------------------------------------------------
template <int WIDTH>
static int run_fwd(const __half* in, const __half* W, uint32_t B, int in_dim, int nhm, uint32_t act, uint32_t out_act,
                   __half* fwd_buf, __half* out, cudaStream_t st, const char* name) {
    const size_t smem = fwd_smem_bytes<WIDTH>(in_dim, nhm);
    if (smem > 227 * 1024) { set_error("%s: network does not fit in shared memory (%zu B)", name, smem); return -2; }
    static size_t configured = 0;
    if (smem > configured) {
        ENERF_CUDA(cudaFuncSetAttribute(k_mlp_fwd<WIDTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), name);
        configured = smem;
    }
    const uint32_t n_tiles = B / Shape<WIDTH>::ROWS;
    const uint32_t grid = n_tiles < (uint32_t)(num_sms() * 4) ? n_tiles : (uint32_t)(num_sms() * 4);
    k_mlp_fwd<WIDTH><<<grid, kWarps * 32, smem, st>>>(in, W, fwd_buf, out, B, in_dim, nhm, act, out_act);
    ENERF_CHECK_LAUNCH(name);
    return 0;
}

template <int WIDTH>
static int run_bwd(const __half* grad, const __half* W, const __half* fwd_buf, __half* bwd_buf, __half* grad_inputs, uint32_t B,
                   int in_dim, int nhm, uint32_t act, cudaStream_t st) {
    const size_t smem = bwd_smem_bytes<WIDTH>(in_dim, nhm, grad_inputs != nullptr);
    if (smem > 227 * 1024) { set_error("ffmlp_backward: network does not fit in shared memory (%zu B)", smem); return -2; }
    static size_t configured = 0;
    if (smem > configured) {
        ENERF_CUDA(cudaFuncSetAttribute(k_mlp_bwd<WIDTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ffmlp_backward");
        configured = smem;
    }
    const uint32_t n_tiles = B / Shape<WIDTH>::ROWS;
    const uint32_t grid = n_tiles < (uint32_t)(num_sms() * 2) ? n_tiles : (uint32_t)(num_sms() * 2);
    k_mlp_bwd<WIDTH><<<grid, kWarps * 32, smem, st>>>(grad, W, fwd_buf, bwd_buf, grad_inputs, B, in_dim, nhm, act);
    ENERF_CHECK_LAUNCH("ffmlp_backward");
    return 0;
}

static int run_wgrad(const __half* dY, int ldY, const __half* X, int ldX, float* dW, uint32_t B, int M, int N, cudaStream_t st) {
    const dim3 grid_mn(1, ceil_div(M, 64), ceil_div(N, 64));
    uint32_t ks = (uint32_t)(2 * num_sms()) / (grid_mn.y * grid_mn.z);
    const uint32_t n_blocks = B / kWgRows;
    if (ks > n_blocks) ks = n_blocks;
    if (ks < 1) ks = 1;
    k_mlp_wgrad<<<dim3(ks, grid_mn.y, grid_mn.z), kWarps * 32, 0, st>>>(dY, ldY, X, ldX, dW, B, M, N);
    ENERF_CHECK_LAUNCH("ffmlp_backward(wgrad)");
    return 0;
}

static int check_dims(const char* name, uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers) {
    ENERF_REQUIRE(hidden_dim == 16 || hidden_dim == 32 || hidden_dim == 64 || hidden_dim == 128 || hidden_dim == 256, name,
                  "hidden_dim should in [16, 32, 64, 128, 256]");
    ENERF_REQUIRE(input_dim > 0 && input_dim % 16 == 0, name, "input_dim should be 16 * m (m > 0)");
    ENERF_REQUIRE(output_dim == 16, name, "output_dim must be padded to 16");
    ENERF_REQUIRE(num_layers >= 2, name, "num_layers should be >= 2");
    ENERF_REQUIRE(B % 128 == 0, name, "batch size must be a multiple of 128");
    return 0;
}

#define ENERF_WIDTH_SWITCH(hidden, CALL)                       \
    switch (hidden) {                                          \
        case 16: { constexpr int WW = 16; CALL; } break;       \
        case 32: { constexpr int WW = 32; CALL; } break;       \
        case 64: { constexpr int WW = 64; CALL; } break;       \
        case 128: { constexpr int WW = 128; CALL; } break;     \
        default: { constexpr int WW = 256; CALL; } break;      \
    }

}  // namespace mlp

namespace tcm {
int tc_forward(const __half* in, const __half* W, uint32_t B, int in_dim, int n_hidden_mm, __half* fwd_buf, __half* out, cudaStream_t st,
               const char* name);
int tc_backward(const __half* grad, const __half* x, const __half* W, const __half* fwd_buf, __half* bwd_buf, __half* grad_inputs, float* dW,
                uint32_t B, int in_dim, int n_hidden_mm, cudaStream_t st);
int tc_forward_sigma_head(const __half* feat, const __half* W, uint32_t B, int n_hidden_mm, __half* fwd_buf, const float* dirs, float* sigma,
                          __half* cin, cudaStream_t st);
int tc_forward_rgb_head(const __half* cin, const __half* W, uint32_t B, int n_hidden_mm, __half* fwd_buf, float* rgb, int n_ch, const int32_t* n_rows_dev,
                        cudaStream_t st);
int tc_backward_rgb(const float* g_rgb, const float* rgb, int n_ch, const __half* cin, const __half* W, const __half* fwd_buf, __half* dcin, float* dW,
                    uint32_t B, int n_hidden_mm, const int32_t* n_rows_dev, cudaStream_t st);
void tc_set_max_ctas(int n);
int tc_forward_density(const __half* feat, const __half* W, uint32_t B, int n_hidden_mm, float* sigma, __half* h, cudaStream_t st);
int tc_backward_density(const float* g_sigma, const float* sigma, const __half* g_h, const __half* feat, const __half* W, __half* dfeat, float* dW,
                        uint32_t B, int n_hidden_mm, cudaStream_t st);
int tc_color_inputs(const float* dirs, uint32_t dir_div, const __half* h, const int32_t* idx, uint32_t n, uint32_t n_pad, float sh_scale, __half* cin,
                    const int32_t* n_dev, cudaStream_t st);
int tc_color_inputs_backward(const __half* dcin, const int32_t* idx, uint32_t n, __half* g_h, const int32_t* n_dev, cudaStream_t st);
int tc_backward_sigma(const float* g_sigma, const float* sigma, const __half* dcin, const __half* feat, const __half* W, const __half* fwd_buf,
                      __half* dfeat, float* dW, uint32_t B, int n_hidden_mm, cudaStream_t st);
}
static int g_mlp_path = 0;   // 0: tcgen05 kernels when eligible, 1: always the generic mma.sync kernels
// the tcgen05 kernels cover E-NeRF's own networks: 32 inputs, 64 wide, ReLU, 2 or 3 layers (3 or 4 matmuls)
static bool tc_eligible(uint32_t input_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation) {
    return g_mlp_path == 0 && hidden_dim == 64 && input_dim == 32 && activation == ENERF_ACT_RELU && output_activation == ENERF_ACT_NONE &&
           num_layers >= 2 && num_layers <= 3;
}
}  // namespace enerf

using namespace enerf;
using namespace enerf::mlp;

extern "C" {

int enerf_ffmlp_forward(const uint16_t* inputs, const uint16_t* weights, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                        uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                        uint16_t* forward_buffer, uint16_t* outputs, void* stream) {
    if (int rc = check_dims("ffmlp_forward", B, input_dim, output_dim, hidden_dim, num_layers)) return rc;
    if (B == 0) return 0;
    ENERF_REQUIRE(forward_buffer != nullptr, "ffmlp_forward", "forward_buffer must not be NULL (use ffmlp_inference)");
    if (tc_eligible(input_dim, hidden_dim, num_layers, activation, output_activation))
        return tcm::tc_forward((const __half*)inputs, (const __half*)weights, B, (int)input_dim, (int)num_layers - 1, (__half*)forward_buffer,
                               (__half*)outputs, as_stream(stream), "ffmlp_forward");
    int rc = 0;
    ENERF_WIDTH_SWITCH(hidden_dim, rc = run_fwd<WW>((const __half*)inputs, (const __half*)weights, B, (int)input_dim, (int)num_layers - 1, activation,
                                                    output_activation, (__half*)forward_buffer, (__half*)outputs, as_stream(stream), "ffmlp_forward"));
    return rc;
}

int enerf_ffmlp_inference(const uint16_t* inputs, const uint16_t* weights, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                          uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                          uint16_t* inference_buffer, uint16_t* outputs, void* stream) {
    (void)inference_buffer;
    if (int rc = check_dims("ffmlp_inference", B, input_dim, output_dim, hidden_dim, num_layers)) return rc;
    if (B == 0) return 0;
    if (tc_eligible(input_dim, hidden_dim, num_layers, activation, output_activation))
        return tcm::tc_forward((const __half*)inputs, (const __half*)weights, B, (int)input_dim, (int)num_layers - 1, nullptr, (__half*)outputs,
                               as_stream(stream), "ffmlp_inference");
    int rc = 0;
    ENERF_WIDTH_SWITCH(hidden_dim, rc = run_fwd<WW>((const __half*)inputs, (const __half*)weights, B, (int)input_dim, (int)num_layers - 1, activation,
                                                    output_activation, nullptr, (__half*)outputs, as_stream(stream), "ffmlp_inference"));
    return rc;
}

int enerf_ffmlp_backward(const uint16_t* grad, const uint16_t* inputs, const uint16_t* weights, const uint16_t* forward_buffer,
                         uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                         uint32_t activation, uint32_t output_activation, int calc_grad_inputs, uint16_t* backward_buffer,
                         uint16_t* grad_inputs, void* grad_weights, int grad_weights_dtype, float* scratch, void* stream) {
    (void)output_activation;  // ignored by the reference backward as well (ffmlp.cu:781)
    if (int rc = check_dims("ffmlp_backward", B, input_dim, output_dim, hidden_dim, num_layers)) return rc;
    ENERF_REQUIRE(grad_weights_dtype == ENERF_F32 || grad_weights_dtype == ENERF_F16, "ffmlp_backward", "bad grad_weights_dtype");
    ENERF_REQUIRE(scratch != nullptr, "ffmlp_backward", "scratch must not be NULL");
    const bool use_tc = tc_eligible(input_dim, hidden_dim, num_layers, activation, ENERF_ACT_NONE);
    ENERF_REQUIRE(use_tc || backward_buffer != nullptr, "ffmlp_backward", "the mma.sync path needs backward_buffer");
    ENERF_REQUIRE(use_tc || forward_buffer != nullptr, "ffmlp_backward", "the mma.sync path needs forward_buffer (recomputation exists on the tcgen05 path only)");
    cudaStream_t st = as_stream(stream);
    const int nhm = (int)num_layers - 1, Wd = (int)hidden_dim, in = (int)input_dim;
    const size_t n_w = (size_t)Wd * (in + (size_t)Wd * nhm + 16);
    ENERF_CUDA(cudaMemsetAsync(scratch, 0, n_w * sizeof(float), st), "ffmlp_backward");
    if (B == 0) return 0;

    int rc = 0;
    if (use_tc) {
        rc = tcm::tc_backward((const __half*)grad, (const __half*)inputs, (const __half*)weights, (const __half*)forward_buffer,
                              (__half*)backward_buffer, calc_grad_inputs ? (__half*)grad_inputs : nullptr, scratch, B, in, nhm, st);
        if (rc) return rc;
        if (grad_weights_dtype == ENERF_F16) {
            k_f32_to_f16<<<ceil_div((uint32_t)n_w, 256u), 256, 0, st>>>(scratch, (__half*)grad_weights, (uint32_t)n_w);
            ENERF_CHECK_LAUNCH("ffmlp_backward(convert)");
        } else if ((void*)scratch != grad_weights) {
            ENERF_CUDA(cudaMemcpyAsync(grad_weights, scratch, n_w * sizeof(float), cudaMemcpyDeviceToDevice, st), "ffmlp_backward");
        }
        return 0;
    }
    ENERF_WIDTH_SWITCH(hidden_dim, rc = run_bwd<WW>((const __half*)grad, (const __half*)weights, (const __half*)forward_buffer,
                                                    (__half*)backward_buffer, calc_grad_inputs ? (__half*)grad_inputs : nullptr, B, in, nhm,
                                                    activation, st));
    if (rc) return rc;

    const __half* fb = (const __half*)forward_buffer;
    const __half* bb = (const __half*)backward_buffer;
    const size_t plane = (size_t)B * Wd;
    // output layer: dW_last[16, W] = grad^T * h_{nl-1}
    if ((rc = run_wgrad((const __half*)grad, 16, fb + (size_t)nhm * plane, Wd, scratch + (size_t)Wd * in + (size_t)nhm * Wd * Wd, B, 16, Wd, st))) return rc;
    // hidden matmul j (weights index j): dW = g_{j+1}^T * h_j ; g_{j+1} lives in backward_buffer[nhm - (j+1)]
    for (int j = nhm - 1; j >= 0; --j)
        if ((rc = run_wgrad(bb + (size_t)(nhm - (j + 1)) * plane, Wd, fb + (size_t)j * plane, Wd, scratch + (size_t)Wd * in + (size_t)j * Wd * Wd, B, Wd, Wd, st))) return rc;
    // input layer: dW_0[W, in] = g_0^T * x ; g_0 = backward_buffer[nhm]
    if ((rc = run_wgrad(bb + (size_t)nhm * plane, Wd, (const __half*)inputs, in, scratch, B, Wd, in, st))) return rc;

    if (grad_weights_dtype == ENERF_F16) {
        k_f32_to_f16<<<ceil_div((uint32_t)n_w, 256u), 256, 0, st>>>(scratch, (__half*)grad_weights, (uint32_t)n_w);
        ENERF_CHECK_LAUNCH("ffmlp_backward(convert)");
    } else if ((void*)scratch != grad_weights) {
        ENERF_CUDA(cudaMemcpyAsync(grad_weights, scratch, n_w * sizeof(float), cudaMemcpyDeviceToDevice, st), "ffmlp_backward");
    }
    return 0;
}

// ---- fused E-NeRF field heads (64-wide ReLU networks, 32 inputs) on the tcgen05 kernels ------------------
static int field_check(const char* name, uint32_t B, uint32_t num_layers) {
    ENERF_REQUIRE(B % 128 == 0, name, "batch size must be a multiple of 128");
    ENERF_REQUIRE(num_layers >= 2 && num_layers <= 3, name, "num_layers must be 2 or 3");
    return 0;
}

int enerf_field_sigma_forward(const uint16_t* feat, const uint16_t* weights, const float* dirs, uint32_t B, uint32_t num_layers,
                              uint16_t* forward_buffer, float* sigma, uint16_t* cin, void* stream) {
    if (int rc = field_check("field_sigma_forward", B, num_layers)) return rc;
    if (B == 0) return 0;
    return tcm::tc_forward_sigma_head((const __half*)feat, (const __half*)weights, B, (int)num_layers - 1, (__half*)forward_buffer, dirs, sigma,
                                      (__half*)cin, as_stream(stream));
}

int enerf_field_color_forward(const uint16_t* cin, const uint16_t* weights, uint32_t B, uint32_t num_layers, uint32_t n_ch,
                              uint16_t* forward_buffer, float* rgb, const int32_t* n_rows_dev, void* stream) {
    if (int rc = field_check("field_color_forward", B, num_layers)) return rc;
    ENERF_REQUIRE(n_ch >= 1 && n_ch <= 4, "field_color_forward", "n_ch must be in [1,4]");
    if (B == 0) return 0;
    return tcm::tc_forward_rgb_head((const __half*)cin, (const __half*)weights, B, (int)num_layers - 1, (__half*)forward_buffer, rgb, (int)n_ch,
                                    n_rows_dev, as_stream(stream));
}

int enerf_field_color_backward(const float* grad_rgb, const float* rgb, uint32_t n_ch, const uint16_t* cin, const uint16_t* weights,
                               const uint16_t* forward_buffer, uint32_t B, uint32_t num_layers, uint16_t* grad_cin, float* grad_weights,
                               const int32_t* n_rows_dev, void* stream) {
    if (int rc = field_check("field_color_backward", B, num_layers)) return rc;
    ENERF_REQUIRE(n_ch >= 1 && n_ch <= 4, "field_color_backward", "n_ch must be in [1,4]");
    const size_t n_w = (size_t)64 * (32 + (size_t)64 * (num_layers - 1) + 16);
    ENERF_CUDA(cudaMemsetAsync(grad_weights, 0, n_w * sizeof(float), as_stream(stream)), "field_color_backward");
    if (B == 0) return 0;
    return tcm::tc_backward_rgb(grad_rgb, rgb, (int)n_ch, (const __half*)cin, (const __half*)weights, (const __half*)forward_buffer,
                                (__half*)grad_cin, grad_weights, B, (int)num_layers - 1, n_rows_dev, as_stream(stream));
}

int enerf_field_sigma_backward(const float* grad_sigma, const float* sigma, const uint16_t* grad_cin, const uint16_t* feat,
                               const uint16_t* weights, const uint16_t* forward_buffer, uint32_t B, uint32_t num_layers, uint16_t* grad_feat,
                               float* grad_weights, void* stream) {
    if (int rc = field_check("field_sigma_backward", B, num_layers)) return rc;
    const size_t n_w = (size_t)64 * (32 + (size_t)64 * (num_layers - 1) + 16);
    ENERF_CUDA(cudaMemsetAsync(grad_weights, 0, n_w * sizeof(float), as_stream(stream)), "field_sigma_backward");
    if (B == 0) return 0;
    return tcm::tc_backward_sigma(grad_sigma, sigma, (const __half*)grad_cin, (const __half*)feat, (const __half*)weights,
                                  (const __half*)forward_buffer, (__half*)grad_feat, grad_weights, B, (int)num_layers - 1, as_stream(stream));
}

// ---- torch-topology field (nerf/network.py): density head and masked colour inputs -------------------------------
int enerf_field_density_forward(const uint16_t* feat, const uint16_t* weights, uint32_t B, uint32_t num_layers, float* sigma, uint16_t* h,
                                void* stream) {
    ENERF_REQUIRE(B % 128 == 0, "field_density_forward", "batch size must be a multiple of 128");
    ENERF_REQUIRE(num_layers >= 1 && num_layers <= 2, "field_density_forward", "num_layers must be 1 (nerf/network.py) or 2 (nerf/network_ff.py)");
    ENERF_REQUIRE(sigma != nullptr, "field_density_forward", "sigma must not be NULL");
    if (B == 0) return 0;
    return tcm::tc_forward_density((const __half*)feat, (const __half*)weights, B, (int)num_layers - 1, sigma, (__half*)h, as_stream(stream));
}

int enerf_field_density_backward(const float* grad_sigma, const float* sigma, const uint16_t* grad_h, const uint16_t* feat, const uint16_t* weights,
                                 uint32_t B, uint32_t num_layers, uint16_t* grad_feat, float* grad_weights, void* stream) {
    ENERF_REQUIRE(B % 128 == 0, "field_density_backward", "batch size must be a multiple of 128");
    ENERF_REQUIRE(num_layers >= 1 && num_layers <= 2, "field_density_backward", "num_layers must be 1 or 2");
    const size_t n_w = (size_t)64 * (32 + (size_t)64 * (num_layers - 1) + 16);
    ENERF_CUDA(cudaMemsetAsync(grad_weights, 0, n_w * sizeof(float), as_stream(stream)), "field_density_backward");
    if (B == 0) return 0;
    return tcm::tc_backward_density(grad_sigma, sigma, (const __half*)grad_h, (const __half*)feat, (const __half*)weights, (__half*)grad_feat,
                                    grad_weights, B, (int)num_layers - 1, as_stream(stream));
}

int enerf_field_color_inputs(const float* dirs, uint32_t dir_div, const uint16_t* h, const int32_t* idx, uint32_t n, uint32_t n_pad, float sh_scale,
                             uint16_t* cin, const int32_t* n_dev, void* stream) {
    ENERF_REQUIRE(dir_div >= 1 && n_pad >= n && n_pad % 128 == 0, "field_color_inputs", "dir_div >= 1, n_pad >= n, n_pad a multiple of 128");
    return tcm::tc_color_inputs(dirs, dir_div, (const __half*)h, idx, n, n_pad, sh_scale, (__half*)cin, n_dev, as_stream(stream));
}

int enerf_field_color_inputs_backward(const uint16_t* grad_cin, const int32_t* idx, uint32_t n, uint16_t* grad_h, const int32_t* n_dev, void* stream) {
    return tcm::tc_color_inputs_backward((const __half*)grad_cin, idx, n, (__half*)grad_h, n_dev, as_stream(stream));
}

int enerf_ffmlp_uses_tcgen05(uint32_t input_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation) {
    return tc_eligible(input_dim, hidden_dim, num_layers, activation, output_activation) ? 1 : 0;
}

int enerf_ffmlp_set_path(int path) {
    ENERF_REQUIRE(path == 0 || path == 1, "ffmlp_set_path", "path must be 0 (auto) or 1 (generic mma.sync kernels)");
    g_mlp_path = path;
    return 0;
}

int enerf_ffmlp_set_max_ctas(int n) {
    ENERF_REQUIRE(n >= 0 && n <= num_sms(), "ffmlp_set_max_ctas", "n must be in [0, SM count] (0 = one CTA per SM)");
    tcm::tc_set_max_ctas(n);
    return 0;
}

int enerf_allocate_splitk(uint64_t size) {
    (void)size;
    return 0;
}
int enerf_free_splitk(void) { return 0; }

}  // extern "C"
