// The two callers' steps that sit directly on either side of the volume-rendering path, fused:
//
//   N2  ray generation on the device: nerf/utils.py:110-169 (get_rays, for given pixel indices) and :185-216
//       (get_event_rays: one pixel seen from two poses) — pixel -> unit camera direction -> world direction, written
//       together with the near/far slab test of raymarching.cu:93-158, so the renderer's first kernel disappears and the
//       [N,3] temporaries of the torch formulation (stack, norm, div, expand, matmul) never exist.
//   N1  the event-loss tail after compositing: nerf/utils.py:494-528 with utils/event_utils.py:23-66 — rgb_to_luma,
//       lin_log (or log with a floor), difference of the two renders, fixed-threshold or normalised loss — forward and
//       analytic backward as three small kernels instead of ~30 ATen launches and their autograd graph.
#include "common.cuh"

namespace enerf {

// camera-frame unit direction of pixel (px, py): nerf/utils.py:160-165 / :203-207
__device__ __forceinline__ void pixel_dir(float px, float py, float fx, float fy, float cx, float cy, float (&d)[3]) {
    const float x = (px - cx) / fx, y = (py - cy) / fy;
    const float nrm = sqrtf(x * x + y * y + 1.0f);
    d[0] = x / nrm;
    d[1] = y / nrm;
    d[2] = 1.0f / nrm;
}
// world direction = R . d with R = pose[:3,:3] (row-major, `stride` floats per row): utils.py:166 / :209-210
__device__ __forceinline__ void rotate(const float* __restrict__ pose, int stride, const float (&d)[3], float (&o)[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = d[0] * __ldg(pose + i * stride) + d[1] * __ldg(pose + i * stride + 1) + d[2] * __ldg(pose + i * stride + 2);
}
// raymarching.cu:110-146
__device__ __forceinline__ void slab(const float (&o)[3], const float (&d)[3], const float* __restrict__ aabb, float min_near, float& near_, float& far_) {
    const float rdx = 1.0f / d[0], rdy = 1.0f / d[1], rdz = 1.0f / d[2];
    float near = (__ldg(aabb + 0) - o[0]) * rdx, far = (__ldg(aabb + 3) - o[0]) * rdx;
    if (near > far) { const float t = near; near = far; far = t; }
    float ny = (__ldg(aabb + 1) - o[1]) * rdy, fy = (__ldg(aabb + 4) - o[1]) * rdy;
    if (ny > fy) { const float t = ny; ny = fy; fy = t; }
    const float kMax = 3.402823466e+38f;
    if (near > fy || ny > far) { near_ = far_ = kMax; return; }
    if (ny > near) near = ny;
    if (fy < far) far = fy;
    float nz = (__ldg(aabb + 2) - o[2]) * rdz, fz = (__ldg(aabb + 5) - o[2]) * rdz;
    if (nz > fz) { const float t = nz; nz = fz; fz = t; }
    if (near > fz || nz > far) { near_ = far_ = kMax; return; }
    if (nz > near) near = nz;
    if (fz < far) far = fz;
    if (near < min_near) near = min_near;
    near_ = near;
    far_ = far;
}

// one thread per ray; poses [B,4,4]; pixel index inds[n] (or n itself) -> (i = idx % W, j = idx / W)
__global__ void k_get_rays(const float* __restrict__ poses, float fx, float fy, float cx, float cy, uint32_t W, const int64_t* __restrict__ inds,
                           uint32_t inds_per_pose, uint32_t B, uint32_t N, const float* __restrict__ aabb, float min_near, float* __restrict__ rays_o,
                           float* __restrict__ rays_d, float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * N) return;
    const uint32_t b = t / N, n = t - b * N;
    const int64_t idx = inds ? inds[inds_per_pose ? (size_t)b * N + n : n] : (int64_t)n;
    const float px = (float)(idx % W), py = (float)(idx / W);
    float dc[3], d[3], o[3];
    pixel_dir(px, py, fx, fy, cx, cy, dc);
    const float* P = poses + (size_t)b * 16;
    rotate(P, 4, dc, d);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        o[i] = __ldg(P + i * 4 + 3);
        rays_o[(size_t)t * 3 + i] = o[i];
        rays_d[(size_t)t * 3 + i] = d[i];
    }
    if (aabb) slab(o, d, aabb, min_near, nears[t], fars[t]);
}

// one thread per event: poses c2w_before / c2w_at [Nevs,3,4]
__global__ void k_event_rays(const float* __restrict__ xs, const float* __restrict__ ys, const float* __restrict__ c2w_before,
                             const float* __restrict__ c2w_at, float fx, float fy, float cx, float cy, uint32_t N, const float* __restrict__ aabb,
                             float min_near, float* __restrict__ o1, float* __restrict__ d1, float* __restrict__ o2, float* __restrict__ d2,
                             float* __restrict__ nf1, float* __restrict__ nf2) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    float dc[3];
    pixel_dir(__ldg(xs + t), __ldg(ys + t), fx, fy, cx, cy, dc);
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        const float* P = (v == 0 ? c2w_before : c2w_at) + (size_t)t * 12;
        float* ro = v == 0 ? o1 : o2;
        float* rd = v == 0 ? d1 : d2;
        float d[3], o[3];
        rotate(P, 4, dc, d);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            o[i] = __ldg(P + i * 4 + 3);
            ro[(size_t)t * 3 + i] = o[i];
            rd[(size_t)t * 3 + i] = d[i];
        }
        float* nf = v == 0 ? nf1 : nf2;
        if (aabb && nf) slab(o, d, aabb, min_near, nf[t], nf[N + t]);      // nf = [nears(N) | fars(N)]
    }
}

// ---------------------------------------------------------------------------------- event loss
struct EvCfg {
    uint32_t N, C, Cp;        // rays, image channels, loss channels (1 with luma, else C)
    int use_luma, linlog, normalized;
    float log_thres, c_thres, weight;
};
static constexpr float kLinlogThres = 20.0f;
static constexpr float kLinSlope = 0.14978661367769955f;      // ln(20)/20  (event_utils.py:63)

// intensity (0..1 image value(s) of one ray) -> log-intensity of loss channel c, and d(log-intensity)/d(x)
__device__ __forceinline__ float log_intensity(const EvCfg& cfg, const float* __restrict__ img, uint32_t c, float& dldx) {
    float x;
    if (cfg.use_luma) x = 0.299f * img[0] + 0.587f * img[1] + 0.114f * img[2];      // rgb_to_luma(esim=True), event_utils.py:40-42
    else x = img[c];
    x *= 255.0f;
    if (cfg.linlog) {
        if (x < kLinlogThres) { dldx = kLinSlope * 255.0f; return kLinSlope * x; }
        dldx = 255.0f / x;
        return logf(x);
    }
    if (x > cfg.log_thres) { dldx = 255.0f / x; return logf(x); }
    dldx = 0.0f;                                                  // torch.maximum passes no gradient to the smaller operand
    return logf(cfg.log_thres);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// pass 1: delta[n,c] = L(img2) - L(img1); acc[c] += delta^2, acc[4] += pol^2     (acc: [0..3] delta^2 per channel, [4] pol^2)
__global__ void k_evloss_delta(const float* __restrict__ img1, const float* __restrict__ img2, const float* __restrict__ pols, EvCfg cfg,
                               float* __restrict__ delta, float* __restrict__ acc) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, sp = 0.f;
    if (n < cfg.N) {
        for (uint32_t c = 0; c < cfg.Cp; ++c) {
            float g1, g2;
            const float d = log_intensity(cfg, img2 + (size_t)n * cfg.C, c, g2) - log_intensity(cfg, img1 + (size_t)n * cfg.C, c, g1);
            delta[(size_t)n * cfg.Cp + c] = d;
            s[c] = d * d;
        }
        const float p = pols[n];
        sp = p * p;
    }
    if (cfg.normalized) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float v = warp_sum(s[c]);
            if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(acc + c, v);
        }
        const float v = warp_sum(sp);
        if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(acc + 4, v);
    }
}

// pass 2: acc[5] += sum_c sum_n r^2 (r = residual), acc[8+c] += sum_n r * delta  (cross term of the normalised loss' gradient)
__global__ void k_evloss_reduce(const float* __restrict__ delta, const float* __restrict__ pols, EvCfg cfg, float* __restrict__ acc) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    float r2 = 0.f, cross[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < cfg.N) {
        const float p = pols[n];
        const float pn = cfg.normalized ? p / (sqrtf(acc[4]) + 1e-9f) : p * cfg.c_thres;
        for (uint32_t c = 0; c < cfg.Cp; ++c) {
            const float d = delta[(size_t)n * cfg.Cp + c];
            const float u = cfg.normalized ? d / (sqrtf(acc[c]) + 1e-9f) : d;
            const float r = u - pn;
            r2 += r * r;
            cross[c] = r * d;
        }
    }
    r2 = warp_sum(r2);
    if ((threadIdx.x & 31) == 0 && r2 != 0.f) atomicAdd(acc + 5, r2);
    if (cfg.normalized) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float v = warp_sum(cross[c]);
            if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(acc + 8 + c, v);
        }
    }
}
__global__ void k_evloss_finish(EvCfg cfg, const float* __restrict__ acc, float* __restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) loss[0] = cfg.weight * acc[5] / (float)((uint64_t)cfg.N * cfg.Cp);
}

// backward: d(loss)/d(img1), d(loss)/d(img2)
__global__ void k_evloss_bwd(const float* __restrict__ img1, const float* __restrict__ img2, const float* __restrict__ pols,
                             const float* __restrict__ delta, const float* __restrict__ acc, const float* __restrict__ grad_loss, EvCfg cfg,
                             float* __restrict__ g1, float* __restrict__ g2) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= cfg.N) return;
    const float scale = grad_loss[0] * cfg.weight * 2.0f / (float)((uint64_t)cfg.N * cfg.Cp);
    const float p = pols[n];
    const float pn = cfg.normalized ? p / (sqrtf(acc[4]) + 1e-9f) : p * cfg.c_thres;
    float ga[4] = {0.f, 0.f, 0.f, 0.f}, gb[4] = {0.f, 0.f, 0.f, 0.f};
    for (uint32_t c = 0; c < cfg.Cp; ++c) {
        const float d = delta[(size_t)n * cfg.Cp + c];
        float dd;                                                  // d(loss)/d(delta[n,c]) / scale
        if (cfg.normalized) {
            const float nr = sqrtf(acc[c]), den = nr + 1e-9f;
            const float u = d / den;
            dd = (u - pn) / den - (nr > 0.f ? d * acc[8 + c] / (nr * den * den) : 0.f);
        } else {
            dd = d - pn;
        }
        float s1, s2;
        log_intensity(cfg, img1 + (size_t)n * cfg.C, c, s1);
        log_intensity(cfg, img2 + (size_t)n * cfg.C, c, s2);
        const float a = -scale * dd * s1, b = scale * dd * s2;
        if (cfg.use_luma) {
            ga[0] += 0.299f * a; ga[1] += 0.587f * a; ga[2] += 0.114f * a;
            gb[0] += 0.299f * b; gb[1] += 0.587f * b; gb[2] += 0.114f * b;
        } else {
            ga[c] = a;
            gb[c] = b;
        }
    }
    for (uint32_t c = 0; c < cfg.C; ++c) {
        g1[(size_t)n * cfg.C + c] = ga[c];
        g2[(size_t)n * cfg.C + c] = gb[c];
    }
}

static int make_cfg(const char* name, uint32_t N, uint32_t C, int use_luma, int linlog, float log_thres, float c_thres, float weight, EvCfg& cfg) {
    ENERF_REQUIRE(C >= 1 && C <= 4, name, "C must be in [1,4]");
    ENERF_REQUIRE(!use_luma || C == 3, name, "use_luma needs 3 channels");
    cfg.N = N;
    cfg.C = C;
    cfg.Cp = use_luma ? 1 : C;
    cfg.use_luma = use_luma;
    cfg.linlog = linlog;
    cfg.normalized = (c_thres == -1.0f);
    cfg.log_thres = log_thres;
    cfg.c_thres = c_thres;
    cfg.weight = weight;
    return 0;
}

}  // namespace enerf

using namespace enerf;

extern "C" {

int enerf_get_rays(const float* poses, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W, const int64_t* inds, uint32_t inds_per_pose,
                   uint32_t B, uint32_t N,
                   const float* aabb, float min_near, float* rays_o, float* rays_d, float* nears, float* fars, void* stream) {
    (void)H;
    if ((uint64_t)B * N == 0) return 0;
    ENERF_REQUIRE(!aabb || (nears && fars), "get_rays", "nears/fars must be given with aabb");
    const uint32_t total = B * N;
    k_get_rays<<<ceil_div(total, 256u), 256, 0, as_stream(stream)>>>(poses, fx, fy, cx, cy, W, inds, inds_per_pose, B, N, aabb, min_near, rays_o, rays_d, nears, fars);
    ENERF_CHECK_LAUNCH("get_rays");
    return 0;
}

int enerf_event_rays(const float* xs, const float* ys, const float* c2w_before, const float* c2w_at, float fx, float fy, float cx, float cy, uint32_t N,
                     const float* aabb, float min_near, float* rays_o1, float* rays_d1, float* rays_o2, float* rays_d2, float* near_far1,
                     float* near_far2, void* stream) {
    if (N == 0) return 0;
    k_event_rays<<<ceil_div(N, 256u), 256, 0, as_stream(stream)>>>(xs, ys, c2w_before, c2w_at, fx, fy, cx, cy, N, aabb, min_near, rays_o1, rays_d1, rays_o2,
                                                                 rays_d2, near_far1, near_far2);
    ENERF_CHECK_LAUNCH("event_rays");
    return 0;
}

int enerf_event_loss_forward(const float* img1, const float* img2, const float* pols, uint32_t N, uint32_t C, int use_luma, int linlog, float log_thres,
                             float c_thres, float weight, float* delta, float* acc, float* loss, void* stream) {
    EvCfg cfg;
    if (int rc = make_cfg("event_loss_forward", N, C, use_luma, linlog, log_thres, c_thres, weight, cfg)) return rc;
    cudaStream_t st = as_stream(stream);
    ENERF_CUDA(cudaMemsetAsync(acc, 0, 16 * sizeof(float), st), "event_loss_forward");
    if (N == 0) {
        ENERF_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st), "event_loss_forward");
        return 0;
    }
    const uint32_t g = ceil_div(N, 256u);
    k_evloss_delta<<<g, 256, 0, st>>>(img1, img2, pols, cfg, delta, acc);
    ENERF_CHECK_LAUNCH("event_loss_forward");
    k_evloss_reduce<<<g, 256, 0, st>>>(delta, pols, cfg, acc);
    ENERF_CHECK_LAUNCH("event_loss_forward");
    k_evloss_finish<<<1, 32, 0, st>>>(cfg, acc, loss);
    ENERF_CHECK_LAUNCH("event_loss_forward");
    return 0;
}

int enerf_event_loss_backward(const float* img1, const float* img2, const float* pols, const float* delta, const float* acc, const float* grad_loss,
                              uint32_t N, uint32_t C, int use_luma, int linlog, float log_thres, float c_thres, float weight, float* grad_img1,
                              float* grad_img2, void* stream) {
    EvCfg cfg;
    if (int rc = make_cfg("event_loss_backward", N, C, use_luma, linlog, log_thres, c_thres, weight, cfg)) return rc;
    if (N == 0) return 0;
    k_evloss_bwd<<<ceil_div(N, 256u), 256, 0, as_stream(stream)>>>(img1, img2, pols, delta, acc, grad_loss, cfg, grad_img1, grad_img2);
    ENERF_CHECK_LAUNCH("event_loss_backward");
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------- N3: event-pair sampler
// nerf/provider.py:1364-1405 (`EventNeRFDataset.collate`, accumulate_evs branch) draws, for each of the M pairs of a step, a start
// event, moves it one back if it is the last event of its pixel, draws an end event among its successors at the same pixel and
// sums the polarities in between — a Python loop with one GPU slice + .sum() per pair.  Here the events of a frame stay resident
// ([E,4] = x, y, t, polarity, grouped by pixel as the provider lays them out), the polarity sums come from an exclusive prefix sum
// (double: exact for +-1 polarities at any E), and one thread per pair turns two uniform variates into (start, end, sum_pol) and —
// given the per-event poses [E,3,4] — directly into the two rays of the pair (the get_event_rays kernel above).
namespace enerf {

__global__ void k_sample_event_pairs(const float* __restrict__ events, const double* __restrict__ pol_prefix, const int32_t* __restrict__ num_succ,
                                     const uint8_t* __restrict__ no_succ, uint32_t E, uint32_t M, int32_t acc_max, const float* __restrict__ u_start,
                                     const float* __restrict__ u_end, int64_t* __restrict__ eidx, int64_t* __restrict__ eidx_end,
                                     float* __restrict__ pols, float* __restrict__ xs, float* __restrict__ ys) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M) return;
    // start ~ U{0..E-1} (np.random.randint(0, E)); a pixel's last event has no successor: take its predecessor (provider.py:1369-1371)
    uint32_t s = min((uint32_t)(__ldg(u_start + t) * (float)E), E - 1);
    if (no_succ[s] && s > 0) s -= 1;                                // (layout validated on the host: EventPairSampler.__init__)
    int32_t n = max(num_succ[s], 1);
    if (acc_max > 0) n = min(n, acc_max + 1);                       // provider.py:1378-1379
    // end ~ U{s+1 .. s+n} (np.random.randint(s+1, s+1+n))
    const uint32_t e = min(s + 1 + min((uint32_t)(__ldg(u_end + t) * (float)n), (uint32_t)(n - 1)), E - 1);
    eidx[t] = s;
    eidx_end[t] = e;
    pols[t] = (float)(pol_prefix[e + 1] - pol_prefix[s + 1]);       // sum of events[s+1 .. e, 3]
    xs[t] = __ldg(events + (size_t)s * 4);
    ys[t] = __ldg(events + (size_t)s * 4 + 1);
}

}  // namespace enerf

extern "C" int enerf_sample_event_pairs(const float* events, const double* pol_prefix, const int32_t* num_successors, const uint8_t* no_successor,
                                        uint32_t E, uint32_t M, int32_t acc_max_num_evs, const float* u_start, const float* u_end, int64_t* eidx,
                                        int64_t* eidx_end, float* pols, float* xs, float* ys, void* stream) {
    if (M == 0) return 0;
    ENERF_REQUIRE(E >= 2, "sample_event_pairs", "a frame needs at least two events");
    enerf::k_sample_event_pairs<<<enerf::ceil_div(M, 256u), 256, 0, enerf::as_stream(stream)>>>(events, pol_prefix, num_successors, no_successor, E, M,
                                                                                       acc_max_num_evs, u_start, u_end, eidx, eidx_end, pols, xs, ys);
    ENERF_CHECK_LAUNCH("sample_event_pairs");
    return 0;
}
