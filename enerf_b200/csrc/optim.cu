// Fused Adam step — the optimizer E-NeRF runs after every backward of the path:
// `torch.optim.Adam(model.get_params(lr), betas=(0.9, 0.99), eps=1e-15)` (main_nerf.py:211-214) over the 13 M-entry hash
// table and the two flat MLP weight vectors, under a GradScaler.  One pass per tensor: 16-byte loads of p, g, m, v, the
// update in registers (same operation order as torch's FusedAdamMathFunctor), 16-byte stores of p, m, v; the gradient is
// un-scaled on the fly (grad_scale) and the whole step is skipped on the device when the scaler found an inf/nan
// (found_inf), so nothing here synchronises with the host and the step can live inside a CUDA graph.
// Purely HBM-bound: 28 bytes per parameter (+2 when the caller keeps an fp16 shadow of the parameter: the hash-grid kernels read
// the table in fp16 under autocast, and writing that copy here replaces a separate 52 MB -> 26 MB cast per step).
#include "common.cuh"

namespace enerf {

struct AdamArgs {
    float lr, beta1, beta2, eps, weight_decay;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a, float inv_scale, float step_size, float bc2_sqrt) {
    g *= inv_scale;
    if (a.weight_decay != 0.f) g = __fmaf_rn(a.weight_decay, p, g);
    m = m + (1.0f - a.beta1) * (g - m);                           // lerp(exp_avg, grad, 1 - beta1)
    v = a.beta2 * v + (1.0f - a.beta2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + a.eps;
    p -= step_size * m / denom;
}

// gradient element i as fp32: G = float (the usual case) or __half (gradients that crossed the wire in fp16, see parallel.py)
template <typename G>
__device__ __forceinline__ float4 load_grad4(const G* __restrict__ g, uint64_t i4);
template <>
__device__ __forceinline__ float4 load_grad4<float>(const float* __restrict__ g, uint64_t i4) { return __ldg(reinterpret_cast<const float4*>(g) + i4); }
template <>
__device__ __forceinline__ float4 load_grad4<__half>(const __half* __restrict__ g, uint64_t i4) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(g) + i4);
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float grad_as_float(float g) { return g; }
__device__ __forceinline__ float grad_as_float(__half g) { return __half2float(g); }

template <typename G>
__global__ void __launch_bounds__(256)
k_adam(float* __restrict__ p, const G* __restrict__ g, float* __restrict__ m, float* __restrict__ v, uint64_t n,
       const float* __restrict__ step, AdamArgs a, const float* __restrict__ grad_scale, const float* __restrict__ found_inf,
       float grad_mul, __half* __restrict__ shadow) {
    if (found_inf && *found_inf != 0.f) return;                  // the scaler skips this step (the fp16 shadow stays valid: p is unchanged)
    const float t = *step;                                        // already incremented by the caller
    // gradients arrive multiplied by the GradScaler's scale and, after a sum-reduction over N ranks, by N: both are undone here
    const float inv_scale = (grad_scale ? 1.0f / *grad_scale : 1.0f) * grad_mul;
    const float bc1 = 1.0f - powf(a.beta1, t), bc2 = 1.0f - powf(a.beta2, t);
    const float step_size = a.lr / bc1, bc2_sqrt = sqrtf(bc2);
    const uint64_t n4 = n >> 2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        const float4 Gv = load_grad4<G>(g, i);
        adam_one(P.x, Gv.x, M.x, V.x, a, inv_scale, step_size, bc2_sqrt);
        adam_one(P.y, Gv.y, M.y, V.y, a, inv_scale, step_size, bc2_sqrt);
        adam_one(P.z, Gv.z, M.z, V.z, a, inv_scale, step_size, bc2_sqrt);
        adam_one(P.w, Gv.w, M.w, V.w, a, inv_scale, step_size, bc2_sqrt);
        reinterpret_cast<float4*>(p)[i] = P;
        reinterpret_cast<float4*>(m)[i] = M;
        reinterpret_cast<float4*>(v)[i] = V;
        if (shadow) {                                             // fp16 copy of the updated parameter (what the hash-grid kernels read)
            const __half2 lo = __floats2half2_rn(P.x, P.y), hi = __floats2half2_rn(P.z, P.w);
            reinterpret_cast<uint2*>(shadow)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
    }
    // tail (n not a multiple of 4)
    const uint64_t tail0 = n4 << 2;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n - tail0) {
        const uint64_t i = tail0 + gid;
        adam_one(p[i], grad_as_float(g[i]), m[i], v[i], a, inv_scale, step_size, bc2_sqrt);
        if (shadow) shadow[i] = __float2half_rn(p[i]);
    }
}

// fp32 -> fp16 copy of a gradient for the wire (parallel.ShardedExchange) that also raises *flag (float, set to 1) when a value is
// non-finite or exceeds `limit` in magnitude (limit = 65504 / world: the sum over the ranks then cannot overflow fp16 either)
__global__ void __launch_bounds__(256)
k_grad_to_half(const float* __restrict__ g, __half* __restrict__ out, uint64_t n, float limit, float* __restrict__ flag) {
    const uint64_t n4 = n >> 2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
        bad |= !(fabsf(v.x) <= limit) | !(fabsf(v.y) <= limit) | !(fabsf(v.z) <= limit) | !(fabsf(v.w) <= limit);      // NaN compares false
        const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
        reinterpret_cast<uint2*>(out)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n - (n4 << 2)) {
        const float v = g[(n4 << 2) + gid];
        bad |= !(fabsf(v) <= limit);
        out[(n4 << 2) + gid] = __float2half_rn(v);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31u) == 0) *flag = 1.0f;
}

}  // namespace enerf

using namespace enerf;

extern "C" int enerf_grad_to_half(const float* grad, uint16_t* out, uint64_t n, float limit, float* flag, void* stream) {
    if (n == 0) return 0;
    ENERF_REQUIRE(flag != nullptr, "grad_to_half", "flag must not be NULL");
    ENERF_REQUIRE(((reinterpret_cast<uintptr_t>(grad) & 15u) | (reinterpret_cast<uintptr_t>(out) & 7u)) == 0, "grad_to_half", "grad must be 16-byte, out 8-byte aligned");
    uint64_t blocks = ((n >> 2) + 255) / 256;
    if (blocks > (uint64_t)num_sms() * 16) blocks = (uint64_t)num_sms() * 16;
    if (blocks == 0) blocks = 1;
    k_grad_to_half<<<(uint32_t)blocks, 256, 0, as_stream(stream)>>>(grad, reinterpret_cast<__half*>(out), n, limit, flag);
    ENERF_CHECK_LAUNCH("grad_to_half");
    return 0;
}

extern "C" int enerf_adam_step(float* param, const void* grad, int grad_dtype, float* exp_avg, float* exp_avg_sq, uint64_t n, const float* step, float lr,
                               float beta1, float beta2, float eps, float weight_decay, const float* grad_scale, const float* found_inf,
                               float grad_mul, uint16_t* half_shadow, void* stream) {
    if (n == 0) return 0;
    ENERF_REQUIRE(step != nullptr, "adam_step", "step must be a device pointer to the (already incremented) step count");
    ENERF_REQUIRE(grad_dtype == ENERF_F32 || grad_dtype == ENERF_F16, "adam_step", "grad_dtype must be ENERF_F32 or ENERF_F16");
    ENERF_REQUIRE(((reinterpret_cast<uintptr_t>(param) | (reinterpret_cast<uintptr_t>(grad) << (grad_dtype == ENERF_F16 ? 1 : 0)) | reinterpret_cast<uintptr_t>(exp_avg) |
                    reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u) == 0, "adam_step", "tensors must be 16-byte aligned");
    ENERF_REQUIRE((reinterpret_cast<uintptr_t>(half_shadow) & 7u) == 0, "adam_step", "half_shadow must be 8-byte aligned");
    const AdamArgs a = {lr, beta1, beta2, eps, weight_decay};
    const uint64_t n4 = n >> 2;
    uint64_t blocks = (n4 + 255) / 256;
    if (blocks > (uint64_t)num_sms() * 16) blocks = (uint64_t)num_sms() * 16;
    if (blocks == 0) blocks = 1;
    if (grad_dtype == ENERF_F16)
        k_adam<__half><<<(uint32_t)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<float*>(param), reinterpret_cast<const __half*>(grad), exp_avg, exp_avg_sq, n,
                                                                        step, a, grad_scale, found_inf, grad_mul, reinterpret_cast<__half*>(half_shadow));
    else
        k_adam<float><<<(uint32_t)blocks, 256, 0, as_stream(stream)>>>(param, reinterpret_cast<const float*>(grad), exp_avg, exp_avg_sq, n, step, a, grad_scale,
                                                                       found_inf, grad_mul, reinterpret_cast<__half*>(half_shadow));
    ENERF_CHECK_LAUNCH("adam_step");
    return 0;
}
