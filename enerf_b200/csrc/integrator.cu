// Fused fixed-step volume-rendering integrator (the torch math of NeRFRenderer.run,
// nerf/renderer.py:230-255) — warp per ray, shuffle scans, no [N,T] temporaries besides the
// weights the colour mask needs.  This is the path every shipped E-NeRF config executes
// (cuda_ray = False), see SURVEY.md §3.2 / K20.
#include "common.cuh"
#include <math.h>

namespace enerf {

static constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float iscan_mul(float v, unsigned lane) {
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const float o = __shfl_up_sync(kFull, v, k);
        if (lane >= (unsigned)k) v *= o;
    }
    return v;
}
__device__ __forceinline__ float iscan_add(float v, unsigned lane) {
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const float o = __shfl_up_sync(kFull, v, k);
        if (lane >= (unsigned)k) v += o;
    }
    return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) v += __shfl_xor_sync(kFull, v, k);
    return v;
}

struct RayStep {
    float alpha, z;
};

// alpha and z of sample i of ray n (i < T)
__device__ __forceinline__ RayStep load_step(const float* __restrict__ sig, const float* __restrict__ zv, uint32_t i, uint32_t T,
                                             float sample_dist, float density_scale) {
    RayStep s;
    s.z = zv[i];
    const float znext = (i + 1 < T) ? zv[i + 1] : 0.f;
    const float delta = (i + 1 < T) ? (znext - s.z) : sample_dist;
    s.alpha = 1.0f - expf(-delta * density_scale * sig[i]);
    return s;
}

__global__ void __launch_bounds__(256)
k_uniform_fwd(const float* __restrict__ sigmas, const float* __restrict__ z_vals, const float* __restrict__ nears,
              const float* __restrict__ fars, uint32_t N, uint32_t T, uint32_t T_dist, float density_scale, float* __restrict__ weights,
              float* __restrict__ weights_sum, float* __restrict__ depth) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const unsigned lane = lane_id();
    const float near = nears[n], far = fars[n];
    const float sample_dist = (far - near) / (float)T_dist;
    const float* sig = sigmas + (size_t)n * T;
    const float* zv = z_vals + (size_t)n * T;
    float* wout = weights + (size_t)n * T;

    float T_carry = 1.0f, acc_ws = 0.f, acc_d = 0.f;
    for (uint32_t base = 0; base < T; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < T;
        RayStep s = {0.f, 0.f};
        if (valid) s = load_step(sig, zv, i, T, sample_dist, density_scale);
        const float keep = valid ? (1.0f - s.alpha + 1e-15f) : 1.0f;
        const float incl = iscan_mul(keep, lane);
        float excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float w = s.alpha * (T_carry * excl);
        if (valid) {
            wout[i] = w;
            acc_ws += w;
            const float oz = fminf(fmaxf((s.z - near) / (far - near), 0.0f), 1.0f);
            acc_d += w * oz;
        }
        T_carry *= __shfl_sync(kFull, incl, 31);
    }
    acc_ws = wsum(acc_ws);
    acc_d = wsum(acc_d);
    if (lane == 0) {
        weights_sum[n] = acc_ws;
        depth[n] = acc_d;
    }
}

// dL/dsigma_i = delta_i*ds*(1-a_i) * ( g_i*T_i - S_i/(1-a_i+eps) ),  S_i = sum_{k>i} g_k*w_k,
// g_i = gw_i + g_ws + g_depth*clamp((z_i-near)/(far-near),0,1)
__global__ void __launch_bounds__(256)
k_uniform_bwd(const float* __restrict__ grad_weights, const float* __restrict__ grad_ws, const float* __restrict__ grad_depth,
              const float* __restrict__ sigmas, const float* __restrict__ z_vals, const float* __restrict__ nears,
              const float* __restrict__ fars, uint32_t N, uint32_t T, uint32_t T_dist, float density_scale, float* __restrict__ grad_sigmas) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const unsigned lane = lane_id();
    const float near = nears[n], far = fars[n];
    const float sample_dist = (far - near) / (float)T_dist;
    const float* sig = sigmas + (size_t)n * T;
    const float* zv = z_vals + (size_t)n * T;
    const float* gw = grad_weights ? grad_weights + (size_t)n * T : nullptr;
    const float g_ws = grad_ws ? grad_ws[n] : 0.f;
    const float g_d = grad_depth ? grad_depth[n] : 0.f;
    float* gout = grad_sigmas + (size_t)n * T;

    // sweep 1: total = sum_k g_k w_k
    float T_carry = 1.0f, total = 0.f;
    for (uint32_t base = 0; base < T; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < T;
        RayStep s = {0.f, 0.f};
        if (valid) s = load_step(sig, zv, i, T, sample_dist, density_scale);
        const float keep = valid ? (1.0f - s.alpha + 1e-15f) : 1.0f;
        const float incl = iscan_mul(keep, lane);
        float excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float w = s.alpha * (T_carry * excl);
        if (valid) {
            const float oz = fminf(fmaxf((s.z - near) / (far - near), 0.0f), 1.0f);
            const float g = (gw ? gw[i] : 0.f) + g_ws + g_d * oz;
            total += g * w;
        }
        T_carry *= __shfl_sync(kFull, incl, 31);
    }
    total = wsum(total);

    // sweep 2: gradients
    T_carry = 1.0f;
    float prefix = 0.f;
    for (uint32_t base = 0; base < T; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < T;
        RayStep s = {0.f, 0.f};
        float delta = 0.f;
        if (valid) {
            s = load_step(sig, zv, i, T, sample_dist, density_scale);
            delta = (i + 1 < T) ? (zv[i + 1] - s.z) : sample_dist;
        }
        const float keep = valid ? (1.0f - s.alpha + 1e-15f) : 1.0f;
        const float incl = iscan_mul(keep, lane);
        float excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float Ti = T_carry * excl;
        const float w = s.alpha * Ti;
        float g = 0.f;
        if (valid) {
            const float oz = fminf(fmaxf((s.z - near) / (far - near), 0.0f), 1.0f);
            g = (gw ? gw[i] : 0.f) + g_ws + g_d * oz;
        }
        const float gw_incl = prefix + iscan_add(g * w, lane);
        const float suffix = total - gw_incl;   // sum over k > i
        if (valid) gout[i] = delta * density_scale * (1.0f - s.alpha) * (g * Ti - suffix / keep);
        prefix = __shfl_sync(kFull, gw_incl, 31);
        T_carry *= __shfl_sync(kFull, incl, 31);
    }
}

// The last two lines of NeRFRenderer.run_cuda (renderer.py:397-398) as one kernel each way instead of eight ATen launches:
//   image += (1 - weights_sum) * bg_color;   depth = clamp(depth - nears, min=0) / (fars - nears)
// evaluated with torch's operation order and roundings (no fused multiply-add), so the results are the ATen ones to the bit.
// bg: NULL (scalar `bg_scalar`), or [n_ch] (bg_per_ray = 0), or [N, n_ch] (bg_per_ray = 1).
__global__ void __launch_bounds__(256)
k_finish_fwd(const float* __restrict__ weights_sum, const float* __restrict__ depth, const float* __restrict__ image,
             const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ bg, int bg_per_ray, float bg_scalar,
             uint32_t N, uint32_t n_ch, float* __restrict__ image_out, float* __restrict__ depth_out) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float rest = __fsub_rn(1.0f, weights_sum[n]);
    for (uint32_t c = 0; c < n_ch; ++c) {
        const float b = bg ? bg[(bg_per_ray ? (size_t)n * n_ch : 0) + c] : bg_scalar;
        image_out[(size_t)n * n_ch + c] = __fadd_rn(image[(size_t)n * n_ch + c], __fmul_rn(rest, b));
    }
    const float near = nears[n];
    float d = __fsub_rn(depth[n], near);
    d = d < 0.0f ? 0.0f : d;                                  // keeps a NaN, like torch.clamp
    depth_out[n] = __fdiv_rn(d, __fsub_rn(fars[n], near));
}
__global__ void __launch_bounds__(256)
k_finish_bwd(const float* __restrict__ g_image, const float* __restrict__ g_depth, const float* __restrict__ depth,
             const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ bg, int bg_per_ray, float bg_scalar,
             uint32_t N, uint32_t n_ch, float* __restrict__ g_weights_sum, float* __restrict__ g_depth_in) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = 0.0f;
    if (g_image)
        for (uint32_t c = 0; c < n_ch; ++c) {
            const float b = bg ? bg[(bg_per_ray ? (size_t)n * n_ch : 0) + c] : bg_scalar;
            acc = __fadd_rn(acc, __fmul_rn(g_image[(size_t)n * n_ch + c], b));
        }
    g_weights_sum[n] = -acc;
    const float near = nears[n];
    const float pass = __fsub_rn(depth[n], near) >= 0.0f ? 1.0f : 0.0f;
    g_depth_in[n] = g_depth ? __fmul_rn(__fdiv_rn(g_depth[n], __fsub_rn(fars[n], near)), pass) : 0.0f;
}

}  // namespace enerf

using namespace enerf;

extern "C" {

int enerf_finish_rays_forward(const float* weights_sum, const float* depth, const float* image, const float* nears, const float* fars,
                              const float* bg, int bg_per_ray, float bg_scalar, uint32_t N, uint32_t n_ch, float* image_out, float* depth_out,
                              void* stream) {
    if (N == 0) return 0;
    k_finish_fwd<<<ceil_div(N, 256u), 256, 0, as_stream(stream)>>>(weights_sum, depth, image, nears, fars, bg, bg_per_ray, bg_scalar, N, n_ch, image_out,
                                                                  depth_out);
    ENERF_CHECK_LAUNCH("finish_rays_forward");
    return 0;
}
int enerf_finish_rays_backward(const float* g_image, const float* g_depth, const float* depth, const float* nears, const float* fars,
                               const float* bg, int bg_per_ray, float bg_scalar, uint32_t N, uint32_t n_ch, float* g_weights_sum, float* g_depth_in,
                               void* stream) {
    if (N == 0) return 0;
    k_finish_bwd<<<ceil_div(N, 256u), 256, 0, as_stream(stream)>>>(g_image, g_depth, depth, nears, fars, bg, bg_per_ray, bg_scalar, N, n_ch, g_weights_sum,
                                                                  g_depth_in);
    ENERF_CHECK_LAUNCH("finish_rays_backward");
    return 0;
}

int enerf_composite_uniform_forward(const float* sigmas, const float* z_vals, const float* nears, const float* fars, uint32_t N,
                                    uint32_t T, uint32_t T_dist, float density_scale, float* weights, float* weights_sum, float* depth,
                                    void* stream) {
    if (N == 0 || T == 0) return 0;
    if (T_dist == 0) T_dist = T;
    k_uniform_fwd<<<ceil_div(N, 8u), 256, 0, as_stream(stream)>>>(sigmas, z_vals, nears, fars, N, T, T_dist, density_scale, weights, weights_sum, depth);
    ENERF_CHECK_LAUNCH("composite_uniform_forward");
    return 0;
}

int enerf_composite_uniform_backward(const float* grad_weights, const float* grad_weights_sum, const float* grad_depth,
                                     const float* sigmas, const float* z_vals, const float* nears, const float* fars, uint32_t N,
                                     uint32_t T, uint32_t T_dist, float density_scale, float* grad_sigmas, void* stream) {
    if (N == 0 || T == 0) return 0;
    if (T_dist == 0) T_dist = T;
    k_uniform_bwd<<<ceil_div(N, 8u), 256, 0, as_stream(stream)>>>(grad_weights, grad_weights_sum, grad_depth, sigmas, z_vals, nears, fars,
                                                                 N, T, T_dist, density_scale, grad_sigmas);
    ENERF_CHECK_LAUNCH("composite_uniform_backward");
    return 0;
}

}  // extern "C"
