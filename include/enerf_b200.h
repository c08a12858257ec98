/*
 * enerf_b200.h — C ABI of the B200-native E-NeRF volume-rendering hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces exactly one function of the
 * reference's four pybind11 extensions (the only FFI the reference has for this path); the
 * reference declaration each one stands in for is cited as file:line relative to the
 * reference repository (knelk/enerf @ 3fb17cd).  The reference passes `at::Tensor` by value
 * and launches on the legacy default stream; here every argument is a plain device pointer
 * or scalar and every call takes the CUDA stream to launch on (`stream` is a cudaStream_t
 * passed as void*; NULL = default stream).  All buffers are caller-owned and pre-allocated,
 * exactly as in the reference (outputs are written in place).
 *
 * Return value: 0 on success, non-zero on failure (bad argument or CUDA error);
 * enerf_last_error() returns a thread-local human-readable message for the last failure.
 * Nothing here falls back to the CPU: without a CUDA device every compute call fails.
 *
 * dtype codes (for entry points whose reference counterpart dispatches on scalar type):
 */
#ifndef ENERF_B200_H
#define ENERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ENERF_F32 0
#define ENERF_F16 1

/* activation codes, ffmlp/ffmlp.py:87-96, ffmlp/src/ffmlp.cu:22-33 */
#define ENERF_ACT_RELU 0
#define ENERF_ACT_EXPONENTIAL 1
#define ENERF_ACT_SINE 2
#define ENERF_ACT_SIGMOID 3
#define ENERF_ACT_SQUAREPLUS 4
#define ENERF_ACT_SOFTPLUS 5
#define ENERF_ACT_NONE 6

const char* enerf_last_error(void);
/* ABI version of this library (bumped on any signature change). */
int enerf_abi_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t enerf_launch_count(void);

/* ------------------------------------------------------------------ raymarching ---- */
/* raymarching/src/raymarching.h:7-18, raymarching/src/raymarching.cu. fp32 only: the
 * reference's Python wrappers cast every float input to fp32 (raymarching.py:21,54,131,...). */

/* raymarching.h:7 near_far_from_aabb; kernel raymarching.cu:93-158 */
int enerf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                             uint32_t N, float min_near, float* nears, float* fars, void* stream);
/* raymarching.h:8 polar_from_ray; raymarching.cu:164-211 */
int enerf_polar_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N,
                         float* coords, void* stream);
/* raymarching.h:9 morton3D; raymarching.cu:216-234 */
int enerf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream);
/* raymarching.h:10 morton3D_invert; raymarching.cu:239-262 */
int enerf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream);
/* raymarching.h:11 packbits; raymarching.cu:269-302. N = number of output bytes. */
int enerf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield,
                   void* stream);
/* raymarching.h:13 march_rays_train; raymarching.cu:313-490.
 * xyzs/dirs [M,3], deltas [M,2], rays [N,3] = (ray id, sample offset, sample count),
 * counter int32[2] += (samples, rays).  Sample ranges are reserved with one atomic per ray, so
 * (as in the reference) the order of ranges is not deterministic. */
int enerf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid,
                           float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                           uint32_t C, uint32_t H, uint32_t M, const float* nears,
                           const float* fars, float* xyzs, float* dirs, float* deltas,
                           int32_t* rays, int32_t* counter, uint32_t perturb, void* stream);
/* raymarching.h:14 composite_rays_train_forward; raymarching.cu:500-589.
 * n_ch: colour channels per sample (the reference hard-wires 3, raymarching.cu:549-551;
 * E-NeRF trains out_dim_color=1, so the channel count is a parameter here). */
int enerf_composite_rays_train_forward(const float* sigmas, const float* rgbs,
                                       const float* deltas, const int32_t* rays, uint32_t M,
                                       uint32_t N, uint32_t n_ch, float* weights_sum,
                                       float* depth, float* image, void* stream);
/* raymarching.h:15 composite_rays_train_backward; raymarching.cu:602-693 */
int enerf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                        const float* sigmas, const float* rgbs,
                                        const float* deltas, const int32_t* rays,
                                        const float* weights_sum, const float* image,
                                        uint32_t M, uint32_t N, uint32_t n_ch,
                                        float* grad_sigmas, float* grad_rgbs, void* stream);
/* raymarching.h:16 march_rays (inference); raymarching.cu:700-813.  The reference's wrapper zero-fills xyzs / dirs / deltas first and the
 * kernel writes the samples; here (and in the _dev / _bounded forms) the kernel writes every row of the n_alive * n_step it is launched
 * for — samples, then zeros — so the buffers may come uninitialised (since ABI v8). */
int enerf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive,
                     const float* rays_t, const float* rays_o, const float* rays_d, float bound,
                     float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                     const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                     float* dirs, float* deltas, uint32_t perturb, void* stream);
/* raymarching.h:17 composite_rays (inference, in place); raymarching.cu:816-909 */
int enerf_composite_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive,
                         float* rays_t, const float* sigmas, const float* rgbs,
                         const float* deltas, uint32_t n_ch, float* weights_sum, float* depth,
                         float* image, void* stream);
/* raymarching.h:18 compact_rays; raymarching.cu:912-939 */
int enerf_compact_rays(uint32_t n_alive, int32_t* rays_alive, const int32_t* rays_alive_old,
                       float* rays_t, const float* rays_t_old, int32_t* alive_counter,
                       void* stream);

/* The three inference-loop primitives with the alive count on the DEVICE (no reference counterpart): `n_alive` is then only an upper
 * bound (any earlier count: rays never come back to life) and slots at or beyond *n_alive_dev are skipped, so NeRFRenderer.run_cuda's
 * loop (renderer.py:364-391) need not read the counter back after every compaction (renderer.py:374).  n_alive_dev == NULL: as above.
 * composite_rays with n_step >= 32 runs one warp per ray (coalesced loads, shuffle scans) instead of the reference's thread per ray. */
int enerf_march_rays_dev(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                         const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                         const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                         uint32_t perturb, const int32_t* n_alive_dev, void* stream);
int enerf_composite_rays_dev(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, float* rays_t, const float* sigmas,
                             const float* rgbs, const float* deltas, uint32_t n_ch, float* weights_sum, float* depth,
                             float* image, const int32_t* n_alive_dev, void* stream);
int enerf_compact_rays_dev(uint32_t n_alive, int32_t* rays_alive, const int32_t* rays_alive_old, float* rays_t,
                           const float* rays_t_old, int32_t* alive_counter, const int32_t* n_alive_dev, void* stream);
/* Marching bounded by the occupied region (no reference counterpart; the samples are those of raymarching.cu:313-480 / :700-813, bit for
 * bit).  The candidate parameters t0, t0 + dt, ... of a ray do not depend on the occupancy (an empty voxel is left by repeated
 * `t += dt`, raymarching.cu:390-398), so candidates outside the box around the occupied cells can be stepped over without a grid
 * lookup and the march can stop behind it.  enerf_occupancy_bounds: bounds = ENERF_OCC_BOUNDS_WORDS(C) 4-byte words, 16-byte aligned:
 * int32 [C][6] (min x, y, z, max x, y, z in cells of each cascade level, rounded outwards to aligned 8 x 4 x 4 blocks; an empty level
 * has min = H, max = -1), then, from the next multiple of four words, float [C][8] = the same box in units of the level's half extent
 * (lo x, y, z, 1 if the level has occupied cells else 0, hi x, y, z, 0), which is what the marchers read (since ABI v8); H a power of
 * two >= 4.  occ_bounds == NULL: the exhaustive march. */
#define ENERF_OCC_BOUNDS_WORDS(C) ((((6u * (C)) + 3u) & ~3u) + 8u * (C))
int enerf_occupancy_bounds(const uint8_t* grid, uint32_t C, uint32_t H, int32_t* bounds, void* stream);
int enerf_march_rays_train_bounded(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                                   uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                                   const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                                   uint32_t perturb, const int32_t* occ_bounds, void* stream);
int enerf_march_rays_bounded(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                             const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                             const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                             uint32_t perturb, const int32_t* n_alive_dev, const int32_t* occ_bounds, void* stream);

/* ------------------------------------------------------------------ gridencoder ---- */
/* gridencoder/src/gridencoder.h:12-13, gridencoder/src/gridencoder.cu.
 * dtype = element type of embeddings/outputs/dy_dx/grad (ENERF_F32 | ENERF_F16); inputs are
 * always fp32 in [0,1].  out_layout: 0 = [L,B,C] (the reference kernel's layout,
 * gridencoder.cu:94), 1 = [B,L*C] (what gridencoder/grid.py:52 returns after its permute). */

/* gridencoder.h:12 grid_encode_forward; gridencoder.cu:74-222,343-370,416-439 */
int enerf_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets,
                              void* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                              float S, uint32_t H, int calc_grad_inputs, void* dy_dx,
                              uint32_t gridtype, int dtype, int out_layout, void* stream);
/* gridencoder.h:13 grid_encode_backward; gridencoder.cu:225-340,372-412,441-471.
 * grad has layout `out_layout`; grad_embeddings must be zero-initialised by the caller (as
 * gridencoder/grid.py:72 does).  grad_dtype: element type of grad_embeddings — the reference
 * uses `dtype`; ENERF_F32 with a half table accumulates in fp32 instead of fp16 atomics. */
int enerf_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                               const int32_t* offsets, void* grad_embeddings, uint32_t B,
                               uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                               int calc_grad_inputs, const void* dy_dx, void* grad_inputs,
                               uint32_t gridtype, int dtype, int grad_dtype, int out_layout,
                               void* stream);

/* The same two kernels fed with RAW positions: x = (raw + in_add) * in_mul is applied inside (in_mul == 0: none).  GridEncoder.forward's
 * `(inputs + bound) / (2 * bound)` (gridencoder/grid.py:144) is in_add = bound, in_mul = fp32(1) / fp32(2 * bound) — ATen's operation order
 * and roundings for that expression, so the features are the same bits — and costs two elementwise passes over the samples less.
 * calc_grad_inputs must be 0 with a transform. */
int enerf_grid_encode_forward_xf(const float* raw_inputs, float in_add, float in_mul, const void* embeddings, const int32_t* offsets,
                                 void* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                 int calc_grad_inputs, void* dy_dx, uint32_t gridtype, int dtype, int out_layout, void* stream);
int enerf_grid_encode_backward_xf(const void* grad, const float* raw_inputs, float in_add, float in_mul, const void* embeddings,
                                  const int32_t* offsets, void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                  uint32_t H, int calc_grad_inputs, const void* dy_dx, void* grad_inputs, uint32_t gridtype, int dtype,
                                  int grad_dtype, int out_layout, void* stream);
/* Scatter strategy of grid_encode_backward (no reference counterpart): 1 (default) = a thread walks
 * 32 consecutive samples of one level and aggregates in registers while they stay in one cell;
 * 0 = one reduction per corner per sample (the reference's strategy).  Same sums either way. */
int enerf_grid_set_backward_mode(int mode);
/* Threads per CTA of the walking scatter: 64, 128 (default; 0 restores it), 192 or 256.  Smaller CTAs fit next to a
 * resident tcgen05 MLP CTA (53 K registers) when the two kernels run concurrently on two streams. */
int enerf_grid_set_backward_block(int threads);
/* Forward kernel selector (tests), D = 3 without input gradients: 1 (default) = a warp walks all levels of its
 * 32 samples (persistent CTAs), 2 = one warp per (32 samples, level), 0 = always the generic kernel.
 * Bit-identical outputs. */
int enerf_grid_set_forward_mode(int mode);

/* -------------------------------------------------------------------- shencoder ---- */
/* shencoder/src/shencoder.h:10,13, shencoder/src/shencoder.cu.  degree C in [1,8], D == 3. */

/* shencoder.h:10 sh_encode_forward; shencoder.cu:27-356,387-419 */
int enerf_sh_encode_forward(const void* inputs, void* outputs, uint32_t B, uint32_t D, uint32_t C,
                            int calc_grad_inputs, void* dy_dx, int dtype, void* stream);
/* shencoder.h:13 sh_encode_backward; shencoder.cu:359-440.  grad_inputs is accumulated into
 * (the reference does `+=` on a zero-initialised buffer, shencoder.cu:378). */
int enerf_sh_encode_backward(const void* grad, const void* inputs, uint32_t B, uint32_t D,
                             uint32_t C, const void* dy_dx, void* grad_inputs, int dtype,
                             void* stream);

/* ------------------------------------------------------------------------ ffmlp ---- */
/* ffmlp/src/ffmlp.h:8-13, ffmlp/src/ffmlp.cu.  All tensors fp16 (uint16_t* here), row-major
 * [B, width]; B must be a multiple of 128 (ffmlp/ffmlp.py:157-159 pads).  weights: flat
 * [hidden,input] + (num_layers-1) x [hidden,hidden] + [output(16),hidden], each row-major
 * [out,in] (ffmlp.cu:631-634).  Accumulation is fp32 (the reference accumulates in fp16). */

/* ffmlp.h:8 ffmlp_forward; ffmlp.cu:635-672.  forward_buffer [num_layers,B,hidden]. */
int enerf_ffmlp_forward(const uint16_t* inputs, const uint16_t* weights, uint32_t B,
                        uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim,
                        uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                        uint16_t* forward_buffer, uint16_t* outputs, void* stream);
/* ffmlp.h:9 ffmlp_inference; ffmlp.cu:674-709.  inference_buffer [B,hidden] is scratch the
 * reference needs; it is accepted and left untouched here. */
int enerf_ffmlp_inference(const uint16_t* inputs, const uint16_t* weights, uint32_t B,
                          uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim,
                          uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                          uint16_t* inference_buffer, uint16_t* outputs, void* stream);
/* ffmlp.h:11 ffmlp_backward; ffmlp.cu:742-894.  backward_buffer [num_layers,B,hidden] may be
 * NULL on the tcgen05 path (the activation gradients then never leave the SM; when given it is filled as the reference does).  forward_buffer may be NULL for 32-input,
 * 64-wide ReLU networks with 2 or 3 layers: the kernel then recomputes the hidden activations of each
 * 128-sample tile from `inputs` (bit-identical to the stored ones) instead of reading them from HBM.  grad_weights: fp16 when
 * grad_weights_dtype == ENERF_F16 (reference), fp32 flat buffer when ENERF_F32; it is
 * overwritten, not accumulated.  `scratch` = caller-owned fp32 buffer with as many elements as
 * `weights` (used as the reduction target; may alias grad_weights when that is fp32). */
int enerf_ffmlp_backward(const uint16_t* grad, const uint16_t* inputs, const uint16_t* weights,
                         const uint16_t* forward_buffer, uint32_t B, uint32_t input_dim,
                         uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                         uint32_t activation, uint32_t output_activation, int calc_grad_inputs,
                         uint16_t* backward_buffer, uint16_t* grad_inputs, void* grad_weights,
                         int grad_weights_dtype, float* scratch, void* stream);
/* ffmlp.h:12-13 allocate_splitk / free_splitk; ffmlp.cu:711-740.  The reference creates one
 * stream+event per layer for its CUTLASS split-K weight-gradient GEMMs.  The weight gradients
 * are fused into the backward kernel here, so these only validate their argument. */
int enerf_allocate_splitk(uint64_t size);
/* Kernel-family selector (no reference counterpart): 0 = automatic — the tcgen05/TMEM kernels for
 * E-NeRF's own shapes (32 inputs, 64 wide, ReLU, 2 or 3 layers; operands staged by TMA), the mma.sync
 * kernels otherwise; 1 = always the generic mma.sync kernels (used by the parity tests to cross-check
 * the two families). */
int enerf_ffmlp_set_path(int path);
/* Cap on the persistent grid of the tcgen05 kernels (no reference counterpart): n CTAs instead of one
 * per SM (148); 0 restores the default.  Used by the pipelined backward (enerf_b200/field.py), which
 * leaves SMs to the hash-grid scatter running on a second stream. */
int enerf_ffmlp_set_max_ctas(int n);
/* 1 when forward/inference/backward of this shape run on the tcgen05 kernels (backward_buffer may
 * then be NULL), else 0.  Not a compute call: usable without a GPU. */
int enerf_ffmlp_uses_tcgen05(uint32_t input_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                             uint32_t output_activation);
int enerf_free_splitk(void);

/* ---------------------------------------------------- fused extras (no reference ABI) ---- */
/* The E-NeRF field of nerf/network_ff.py:51-73 with its glue fused into the MLP kernels (tcgen05
 * path: hidden 64, ReLU, 32 inputs, B % 128 == 0, fp16 operands):
 *   sigma-net head : h = fp16(sigma_net(feat)); sigma = exp(h[0]) (trunc_exp, activation.py:10);
 *                    cin = [SH_4(fp16(dir)) | h[1:16] | 0]   (shencoder + torch.cat + zeros_like)
 *   colour-net head: rgb[c] = sigmoid(fp16(color_net(cin))[c]), c < n_ch, as fp32
 *   backward       : the prologues apply sigmoid' / trunc_exp' and pick the geo_feat columns.
 * forward_buffer [num_layers,B,64] may be NULL in the forward calls (inference, or training with
 * recomputation) and in the backward calls (the hidden activations are then recomputed per tile from
 * cin / feat; same gradients, no forward_buffer traffic).  grad_weights: fp32, overwritten.
 * n_rows_dev (colour-net calls; may be NULL): int32 device scalar with the number of valid rows when only the device knows it (the
 * compacted `weights > 1e-4` batch of NeRFRenderer.run); B is then the capacity and 128-row tiles beyond the count are skipped. */
int enerf_field_sigma_forward(const uint16_t* feat, const uint16_t* weights, const float* dirs, uint32_t B,
                              uint32_t num_layers, uint16_t* forward_buffer, float* sigma, uint16_t* cin,
                              void* stream);
int enerf_field_color_forward(const uint16_t* cin, const uint16_t* weights, uint32_t B, uint32_t num_layers,
                              uint32_t n_ch, uint16_t* forward_buffer, float* rgb, const int32_t* n_rows_dev, void* stream);
int enerf_field_color_backward(const float* grad_rgb, const float* rgb, uint32_t n_ch, const uint16_t* cin,
                               const uint16_t* weights, const uint16_t* forward_buffer, uint32_t B,
                               uint32_t num_layers, uint16_t* grad_cin, float* grad_weights, const int32_t* n_rows_dev,
                               void* stream);
int enerf_field_sigma_backward(const float* grad_sigma, const float* sigma, const uint16_t* grad_cin,
                               const uint16_t* feat, const uint16_t* weights, const uint16_t* forward_buffer,
                               uint32_t B, uint32_t num_layers, uint16_t* grad_feat, float* grad_weights,
                               void* stream);


/* The same field under torch.no_grad as ONE kernel (no reference counterpart as a single call): hash-grid gather -> sigma-net ->
 * trunc_exp / SH / colour-net input -> colour-net -> sigmoid, i.e. the chain gridencoder.h:12 grid_encode_forward -> ffmlp.h:13
 * ffmlp_inference -> nerf/network_ff.py:58-73 -> ffmlp_inference that NeRFRenderer.run_cuda's inference loop (nerf/renderer.py:380-381)
 * runs once per marching round.  Features and colour-net inputs stay in shared memory; sigma [B] fp32 and rgb [B, n_ch] fp32 are the
 * bits the unfused calls produce.  raw_xyz [B,3] fp32 with x = (raw + in_add) * in_mul applied inside (in_mul == 0: already in [0,1]),
 * dirs [B,3] fp32, embeddings: the fp16 table [entries, C]; L = 16, C = 2, D = 3, num_layers = 2, num_layers_color = 3 (E-NeRF's
 * field) — other shapes are refused (-2) and take the unfused chain.  B is arbitrary (no multiple-of-128 rule). */
int enerf_field_infer(const float* raw_xyz, float in_add, float in_mul, const float* dirs, const uint16_t* embeddings,
                      const int32_t* offsets, uint32_t L, uint32_t C, float S, uint32_t H, uint32_t gridtype,
                      const uint16_t* w_sigma, uint32_t num_layers, const uint16_t* w_color, uint32_t num_layers_color,
                      uint32_t B, uint32_t n_ch, float* sigma, float* rgb, void* stream);
/* The same with the number of live rows on the device: only the first min(B, *n_units_dev * rows_per_unit) rows are evaluated and
 * written (n_units_dev: the int32 alive-ray count enerf_compact_rays_dev leaves, rows_per_unit: the round's n_step).  In the inference
 * loop (nerf/renderer.py:364-391) the host then sizes a round by the last count it has read — an upper bound — without paying the field
 * for the rays that died since.  n_units_dev == NULL: all B rows. */
int enerf_field_infer_alive(const float* raw_xyz, float in_add, float in_mul, const float* dirs, const uint16_t* embeddings,
                            const int32_t* offsets, uint32_t L, uint32_t C, float S, uint32_t H, uint32_t gridtype,
                            const uint16_t* w_sigma, uint32_t num_layers, const uint16_t* w_color, uint32_t num_layers_color,
                            uint32_t B, uint32_t n_ch, float* sigma, float* rgb, const int32_t* n_units_dev, uint32_t rows_per_unit,
                            void* stream);

/* The torch-topology field of nerf/network.py:104-199 (what every shipped E-NeRF config runs: sigma-net Linear(32,64)-ReLU-
 * Linear(64,16), colour-net Linear(31,64)-ReLU-Linear(64,64)-ReLU-Linear(64,C), no bias) on the same tcgen05 kernels.
 * Weights are the nn.Linear matrices ([out,in] row-major) concatenated in FFMLP order, the colour-net's first matrix padded with a
 * zero 32nd column and its last one with zero rows up to 16.
 *   density forward : h = fp16(sigma_net(feat)) [B,16]; sigma = exp(h[0]) (trunc_exp); geo_feat = h[1:16].  num_layers = 1 for
 *                     nerf/network.py (2 matmuls), 2 for the FFMLP sigma-net (used by the occupancy-grid refresh).  h may be NULL.
 *   density backward: dL/dh[0] += grad_sigma * exp(clamp(h0, -15, 15)) (activation.py:15-18); hidden activations are recomputed per
 *                     tile; grad_sigma / grad_h may be NULL (= zero); grad_weights fp32, overwritten.
 *   colour inputs   : rows of the samples idx[0..n) (the `weights > 1e-4` mask of renderer.py:236, compacted): [SH_4(fp16(dir)) *
 *                     sh_scale | h[idx,1:16] | 0], dirs [B/dir_div,3] (one direction per dir_div consecutive samples), rows n..n_pad
 *                     zero; then enerf_field_color_forward / _backward with num_layers = 2 run the colour-net on the compact batch.
 *   colour inputs backward: grad_h[idx[i],1:16] = grad_cin[i,16:31] (grad_h [B,16] zero-initialised by the caller).
 *   n_dev (here and in gather_rows / scatter_rows; may be NULL): the count from enerf_compact_mask, still on the device — n / n_pad
 *   are then capacities, rows [count, next multiple of 128) are zero-filled and nothing beyond is touched: no host round trip. */
int enerf_field_density_forward(const uint16_t* feat, const uint16_t* weights, uint32_t B, uint32_t num_layers, float* sigma,
                                uint16_t* h, void* stream);
int enerf_field_density_backward(const float* grad_sigma, const float* sigma, const uint16_t* grad_h, const uint16_t* feat,
                                 const uint16_t* weights, uint32_t B, uint32_t num_layers, uint16_t* grad_feat,
                                 float* grad_weights, void* stream);
int enerf_field_color_inputs(const float* dirs, uint32_t dir_div, const uint16_t* h, const int32_t* idx, uint32_t n,
                             uint32_t n_pad, float sh_scale, uint16_t* cin, const int32_t* n_dev, void* stream);
int enerf_field_color_inputs_backward(const uint16_t* grad_cin, const int32_t* idx, uint32_t n, uint16_t* grad_h,
                                      const int32_t* n_dev, void* stream);
/* Order-preserving compaction: indices[0..count) = { i < n : values[i] > thresh } in increasing order, count[0] = how many —
 * `torch.nonzero(values > thresh)` (renderer.py:236-242 mask, :523 occupied cells) without the host round trip.
 * scratch: int32 [ceil(n/4096)].  indices may be NULL (count only). */
int enerf_compact_greater(const float* values, float thresh, uint32_t n, int32_t* indices, int32_t* count, int32_t* scratch,
                          void* stream);
/* the same for a byte mask (a torch.bool tensor): { i : mask[i] != 0 } */
int enerf_compact_mask(const uint8_t* mask, uint32_t n, int32_t* indices, int32_t* count, int32_t* scratch, void* stream);
/* dst[i] = src[idx[i]] for i < n, zero rows for n <= i < n_pad / dst[idx[i]] = src[i]; rows of row_bytes (multiple of 4). */
int enerf_gather_rows(const void* src, const int32_t* idx, uint32_t n, uint32_t n_pad, uint32_t row_bytes, void* dst,
                      const int32_t* n_dev, void* stream);
int enerf_scatter_rows(const void* src, const int32_t* idx, uint32_t n, uint32_t row_bytes, void* dst, const int32_t* n_dev,
                       void* stream);
/* image[n,c] = sum_t weights[n,t] * rgbs[n,t,c] (renderer.py:255) and its backward (grad_weights / grad_rgbs may be NULL). */
int enerf_weighted_sum_forward(const float* weights, const float* rgbs, uint32_t N, uint32_t T, uint32_t n_ch, float* image,
                               void* stream);
int enerf_weighted_sum_backward(const float* grad_image, const float* weights, const float* rgbs, uint32_t N, uint32_t T,
                                uint32_t n_ch, float* grad_weights, float* grad_rgbs, void* stream);

/* Occupancy-grid maintenance, nerf/renderer.py:408-563, without host synchronisation.
 *   occ_points_full   : jittered query position of EVERY cell, sample t = cascade*H^3 + Morton index (renderer.py:485-515).
 *   occ_points_partial: per cascade n_pick uniformly random cells + n_pick cells drawn from the occupied ones (renderer.py:517-545);
 *                       occ_list [C,H^3] / occ_count [C] from enerf_compact_greater(density_grid[c], 0); indices [C*2*n_pick] out.
 *                       noise ([n,3] uniform variates), rand_coords [C,n_pick,3], rand_occ [C,n_pick]: scripted draws for the parity
 *                       tests; NULL = in-kernel PCG32 streams keyed by (seed, sample).
 *   occ_update        : grid = max(grid*decay, sigma*scale) where both >= 0 (renderer.py:548-549), duplicates resolved "last sample
 *                       wins"; mean of clamp(grid,0) -> *mean_density; bitfield = grid > min(mean, density_thresh) (renderer.py:550-555).
 *                       indices == NULL: full refresh (sigmas [C*H^3] in cell order).  owner: int32 [C*H^3] scratch holding -1 (left so);
 *                       sum: double scratch.
 *   mark_untrained_grid: density -1 for cells outside every camera frustum, poses [B,4,4] cam-to-world (renderer.py:408-471). */
int enerf_occ_points_full(float* xyzs, uint32_t C, uint32_t H, float bound, const float* noise, uint64_t seed, void* stream);
int enerf_occ_points_partial(float* xyzs, int32_t* indices, uint32_t n_pick, uint32_t C, uint32_t H, float bound,
                             const int32_t* occ_list, const int32_t* occ_count, const int32_t* rand_coords,
                             const int32_t* rand_occ, const float* noise, uint64_t seed, void* stream);
int enerf_occ_update(float* density_grid, const float* sigmas, const int32_t* indices, uint32_t per_cascade, uint32_t C,
                     uint32_t H, float decay, float scale, float density_thresh, int32_t* owner, double* sum,
                     uint8_t* bitfield, float* mean_density, void* stream);
int enerf_mark_untrained_grid(float* density_grid, const float* poses, uint32_t B, float fx, float fy, float cx, float cy,
                              uint32_t C, uint32_t H, float bound, void* stream);

/* Fixed-step integrator of NeRFRenderer.run (nerf/renderer.py:230-255), which the reference
 * evaluates as ~25 ATen kernels over [N,T] temporaries.  One warp per ray:
 *   delta_i = z_{i+1}-z_i, last = (far-near)/T_dist      (renderer.py:177,230-231; T_dist = the COARSE step
 *             count `num_steps`, which the reference keeps for the last delta even when upsampling has made
 *             the rows T = num_steps + upsample_steps long; 0 = T)
 *   alpha_i = 1-exp(-delta_i*density_scale*sigma_i)      (renderer.py:232)
 *   w_i = alpha_i * prod_{j<i}(1-alpha_j+1e-15)          (renderer.py:233-234)
 *   weights_sum = sum w, depth = sum w*clamp((z-near)/(far-near),0,1)   (renderer.py:248-252)
 * sigmas, z_vals, weights: [N,T]; nears, fars, weights_sum, depth: [N]. */
int enerf_composite_uniform_forward(const float* sigmas, const float* z_vals, const float* nears,
                                    const float* fars, uint32_t N, uint32_t T, uint32_t T_dist,
                                    float density_scale, float* weights, float* weights_sum,
                                    float* depth, void* stream);
/* d(loss)/d(sigmas) given d(loss)/d(weights) [N,T], d/d(weights_sum) [N], d/d(depth) [N]
 * (each may be NULL = zero). */
int enerf_composite_uniform_backward(const float* grad_weights, const float* grad_weights_sum,
                                     const float* grad_depth, const float* sigmas,
                                     const float* z_vals, const float* nears, const float* fars,
                                     uint32_t N, uint32_t T, uint32_t T_dist, float density_scale,
                                     float* grad_sigmas, void* stream);

/* nerf/renderer.py:397-398, the two lines after the compositing of run_cuda, one kernel each way (ATen: eight launches):
 *   image_out = image + (1 - weights_sum) * bg_color;  depth_out = clamp(depth - nears, min=0) / (fars - nears)
 * in torch's operation order and roundings.  bg: NULL (use bg_scalar), [n_ch] (bg_per_ray = 0) or [N, n_ch] (bg_per_ray = 1).
 * backward: g_weights_sum = -sum_c g_image * bg, g_depth_in = g_depth / (fars - nears) where depth - nears >= 0 (g_image passes through;
 * either incoming gradient may be NULL). */
int enerf_finish_rays_forward(const float* weights_sum, const float* depth, const float* image, const float* nears, const float* fars,
                              const float* bg, int bg_per_ray, float bg_scalar, uint32_t N, uint32_t n_ch, float* image_out,
                              float* depth_out, void* stream);
int enerf_finish_rays_backward(const float* g_image, const float* g_depth, const float* depth, const float* nears, const float* fars,
                               const float* bg, int bg_per_ray, float bg_scalar, uint32_t N, uint32_t n_ch, float* g_weights_sum,
                               float* g_depth_in, void* stream);

/* ------------------------------------------ next rows of the path ---- */
/* N2 — ray generation on the device, fused with near_far_from_aabb.
 * nerf/utils.py:110-169 get_rays for given pixels: poses [B,4,4] cam2world (row-major), pixel n of every pose is
 * inds[n] (flat index j*W+i; NULL = pixel n; inds_per_pose != 0: inds is [B,N], one set per pose), direction = normalize(((i-cx)/fx, (j-cy)/fy, 1)) rotated by pose[:3,:3],
 * origin = pose[:3,3].  rays_o, rays_d: [B,N,3].  aabb (6 floats, device) may be NULL; otherwise nears/fars [B,N] get the
 * slab test of raymarching.h:7 with `min_near`. */
int enerf_get_rays(const float* poses, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                   const int64_t* inds, uint32_t inds_per_pose, uint32_t B, uint32_t N, const float* aabb,
                   float min_near,
                   float* rays_o, float* rays_d, float* nears, float* fars, void* stream);
/* nerf/utils.py:185-216 get_event_rays: pixel (xs[n], ys[n]) seen from c2w_before[n] and c2w_at[n] ([N,3,4] row-major).
 * near_far1 / near_far2: [2,N] = nears then fars of the two ray sets (NULL or aabb == NULL: not computed). */
int enerf_event_rays(const float* xs, const float* ys, const float* c2w_before, const float* c2w_at, float fx,
                     float fy, float cx, float cy, uint32_t N, const float* aabb, float min_near,
                     float* rays_o1, float* rays_d1, float* rays_o2, float* rays_d2, float* near_far1,
                     float* near_far2, void* stream);
/* N1 — the event-loss tail after compositing, nerf/utils.py:494-528 (+ utils/event_utils.py:23-66 rgb_to_luma, lin_log).
 * img1, img2: [N,C] rendered intensities of the two poses of each event pair (fp32); pols: [N] accumulated polarity.
 * use_luma: BT.601 luma of 3 channels first (esim coefficients); linlog: lin_log(255*x, 20), else log(max(255*x, log_thres));
 * c_thres != -1: loss = weight * mean((delta - pols*c_thres)^2); c_thres == -1: normalised loss
 * weight * mean((delta/(||delta||+1e-9) - pols/(||pols||+1e-9))^2), norms over the N rays per channel.
 * delta [N,C'] (C' = 1 with luma, else C) and acc [16] are outputs the backward call needs; loss: 1 float.
 * (The reference's non-linlog luma branch uses pred_luma1 for both renders, utils.py:504-505 — an obvious slip that makes
 * the loss constant; both renders are used here.) */
int enerf_event_loss_forward(const float* img1, const float* img2, const float* pols, uint32_t N, uint32_t C,
                             int use_luma, int linlog, float log_thres, float c_thres, float weight,
                             float* delta, float* acc, float* loss, void* stream);
int enerf_event_loss_backward(const float* img1, const float* img2, const float* pols, const float* delta,
                              const float* acc, const float* grad_loss, uint32_t N, uint32_t C, int use_luma,
                              int linlog, float log_thres, float c_thres, float weight, float* grad_img1,
                              float* grad_img2, void* stream);
/* N4 — the optimizer step that follows every backward of the path: torch.optim.Adam as configured by main_nerf.py:211-214
 * (amsgrad off, maximize off), fused into one pass per tensor.  `step`: device pointer to the step count of this parameter,
 * ALREADY incremented for this step; grad_scale / found_inf: the GradScaler's device scalars (NULL = 1 / 0): gradients are
 * divided by *grad_scale and multiplied by grad_mul on the fly (grad_mul = 1/N after a sum-reduction over N ranks, else 1) and the
 * call is a no-op when *found_inf != 0.  fp32 tensors, 16-byte aligned; grad is fp32 or (grad_dtype == ENERF_F16, 8-byte aligned) fp16 —
 * gradients that crossed the NVLink wire in half precision (enerf_b200/parallel.py).
 * half_shadow (may be NULL): fp16 [n] copy of the UPDATED parameter written in the same pass — gridencoder/grid.py:38-39
 * re-casts the whole table to half on every forward under autocast; with the shadow that 52 MB -> 26 MB pass disappears. */
int enerf_adam_step(float* param, const void* grad, int grad_dtype, float* exp_avg, float* exp_avg_sq, uint64_t n, const float* step,
                    float lr, float beta1, float beta2, float eps, float weight_decay, const float* grad_scale,
                    const float* found_inf, float grad_mul, uint16_t* half_shadow, void* stream);
/* Wire format of the data-parallel gradient exchange (enerf_b200/parallel.py; no reference counterpart — the reference is single-GPU):
 * out = fp16(grad); *flag (device float, zeroed by the caller) is set to 1 when an element is non-finite or larger than `limit` in
 * magnitude (limit = 65504 / world: the fp16 sum over the ranks then cannot overflow either). */
int enerf_grad_to_half(const float* grad, uint16_t* out, uint64_t n, float limit, float* flag, void* stream);
/* N3 — event-pair sampler, nerf/provider.py:1364-1405 (collate, accumulate_evs branch).  events [E,4] fp32 = (x, y, t, polarity),
 * grouped by pixel as the provider stores them; pol_prefix [E+1] double = exclusive prefix sum of the polarity column;
 * num_successors [E] (provider.py:1181-1187), no_successor [E] = 1 for the last event of a pixel (:1177-1178); u_start, u_end [M]
 * uniform variates in [0,1) supplied by the caller (start = floor(u_start*E), end = start+1+floor(u_end*n)).  Outputs: event ids of
 * the pair, the accumulated polarity between them and the pixel of the start event. */
int enerf_sample_event_pairs(const float* events, const double* pol_prefix, const int32_t* num_successors,
                             const uint8_t* no_successor, uint32_t E, uint32_t M, int32_t acc_max_num_evs,
                             const float* u_start, const float* u_end, int64_t* eidx, int64_t* eidx_end, float* pols,
                             float* xs, float* ys, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ENERF_B200_H */
