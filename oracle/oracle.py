"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY — see oracle/enerf_oracle.c).

numpy in, numpy out.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product (enerf_b200/) never does.

Parity pin: the reference has no tests or golden vectors for this path, so the oracle is pinned
by closed forms / known answers (tests/test_oracle.py) and by the reference's own CUDA build
(oracle/_ref, tests/test_ref_parity.py, tests/golden/*.npz generated on a B200).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRC = os.path.join(_HERE, "enerf_oracle.c")
_lib = None


def build(force=False):
    """Compile oracle/enerf_oracle.c -> oracle/liboracle.so (gcc; OpenMP when available)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    base = ["-O2", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas", "-o", _SO, _SRC, "-lm"]
    attempts = [["/usr/bin/gcc", "-fopenmp"], ["gcc", "-fopenmp"], ["/usr/bin/gcc"], ["gcc"]]
    err = None
    for a in attempts:
        try:
            r = subprocess.run(a + base, capture_output=True, text=True)
        except FileNotFoundError as e:
            err = str(e)
            continue
        if r.returncode == 0:
            return _SO
        err = r.stderr
    raise RuntimeError(f"could not build the oracle: {err}")


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_pcg32_first_float.restype = C.c_float
        _lib.oracle_pcg32_first_float.argtypes = [C.c_uint64, C.c_uint64]
        _lib.oracle_pcg32_stream.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_u32, _f = C.c_uint32, C.c_float


# ---------------------------------------------------------------------------------- rng / morton
def pcg32_stream(initstate, initseq, n):
    out = np.empty(n, dtype=np.uint32)
    lib().oracle_pcg32_stream(initstate, initseq, n, _p(out))
    return out


def pcg32_first_float(initstate, initseq=1):
    return float(lib().oracle_pcg32_first_float(initstate, initseq))


def morton3D(coords):
    coords = _i32(coords)
    out = np.empty(coords.shape[0], dtype=np.int32)
    lib().oracle_morton3D(_p(coords), _u32(coords.shape[0]), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i32(indices)
    out = np.empty((indices.shape[0], 3), dtype=np.int32)
    lib().oracle_morton3D_invert(_p(indices), _u32(indices.shape[0]), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f32(grid)
    n = grid.size // 8
    out = np.empty(n, dtype=np.uint8)
    lib().oracle_packbits(_p(grid), _u32(n), _f(thresh), _p(out))
    return out


# ---------------------------------------------------------------------------------- rays
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().oracle_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), _u32(N), _f(min_near), _p(nears), _p(fars))
    return nears, fars


def march_rays_train(rays_o, rays_d, bound, bitfield, C_, H, nears, fars, perturb=False, dt_gamma=0.0, max_steps=1024, M=None):
    """Returns xyzs [M,3], dirs [M,3], deltas [M,2], rays [N,3], counter [2]; ranges in ray order."""
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    nears, fars = _f32(nears), _f32(fars)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    N = rays_o.shape[0]
    if M is None:
        M = N * max_steps
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    rays = np.zeros((N, 3), np.int32)
    counter = np.zeros(2, np.int32)
    lib().oracle_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), _f(bound), _f(dt_gamma), _u32(max_steps), _u32(N), _u32(C_),
                                  _u32(H), _u32(M), _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas), _p(rays), _p(counter),
                                  _u32(int(perturb)))
    return xyzs, dirs, deltas, rays, counter


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C_, H, nears, fars, perturb=0, dt_gamma=0.0,
               max_steps=1024):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    rays_alive, rays_t = _i32(rays_alive), _f32(rays_t)
    nears, fars = _f32(nears), _f32(fars)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    M = n_alive * n_step
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    lib().oracle_march_rays(_u32(n_alive), _u32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d), _f(bound), _f(dt_gamma),
                            _u32(max_steps), _u32(C_), _u32(H), _p(bitfield), _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas),
                            _u32(int(perturb)))
    return xyzs, dirs, deltas


def composite_rays_train_forward(sigmas, rgbs, deltas, rays):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    n_ch = rgbs.shape[1]
    ws, depth, image = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, n_ch), np.float32)
    lib().oracle_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), _u32(M), _u32(N), _u32(n_ch), _p(ws),
                                              _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    grad_ws, grad_image, weights_sum, image = _f32(grad_ws), _f32(grad_image), _f32(weights_sum), _f32(image)
    M, N = sigmas.shape[0], rays.shape[0]
    n_ch = rgbs.shape[1]
    gs, gr = np.zeros(M, np.float32), np.zeros((M, n_ch), np.float32)
    lib().oracle_composite_rays_train_backward(_p(grad_ws), _p(grad_image), _p(sigmas), _p(rgbs), _p(deltas), _p(rays), _p(weights_sum),
                                               _p(image), _u32(M), _u32(N), _u32(n_ch), _p(gs), _p(gr))
    return gs, gr


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
    """Returns updated copies (rays_t, weights_sum, depth, image)."""
    rays_alive = _i32(rays_alive)
    rays_t, weights_sum, depth, image = _f32(rays_t).copy(), _f32(weights_sum).copy(), _f32(depth).copy(), _f32(image).copy()
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    n_ch = image.shape[1]
    lib().oracle_composite_rays(_u32(n_alive), _u32(n_step), _p(rays_alive), _p(rays_t), _p(sigmas), _p(rgbs), _p(deltas), _u32(n_ch),
                                _p(weights_sum), _p(depth), _p(image))
    return rays_t, weights_sum, depth, image


def compact_rays(n_alive, rays_alive_old, rays_t_old):
    rays_alive_old, rays_t_old = _i32(rays_alive_old), _f32(rays_t_old)
    rays_alive, rays_t = np.zeros_like(rays_alive_old), np.zeros_like(rays_t_old)
    counter = np.zeros(1, np.int32)
    lib().oracle_compact_rays(_u32(n_alive), _p(rays_alive), _p(rays_alive_old), _p(rays_t), _p(rays_t_old), _p(counter))
    return rays_alive, rays_t, int(counter[0])


# ---------------------------------------------------------------------------------- hash grid
def grid_offsets(input_dim=3, num_levels=16, per_level_scale=2.0, base_resolution=16, log2_hashmap_size=19):
    """gridencoder/grid.py:113-123"""
    offsets, offset = [], 0
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        n = min(2 ** log2_hashmap_size, (resolution + 1) ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32)


def per_level_scale_for(desired_resolution, base_resolution=16, num_levels=16):
    """gridencoder/grid.py:96-97"""
    return np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))


def grid_encode_forward(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                        level_scales=None):
    """inputs [B,D] in [0,1]; embeddings [n,C] float32 or float16.  Returns outputs [L,B,C]
    (reference kernel layout) and dy_dx [B,L,D,C] (or None), in the embeddings' dtype."""
    inputs = _f32(inputs)
    offsets = _i32(offsets)
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    Cc = embeddings.shape[1]
    S = np.float32(np.log2(per_level_scale))
    half = embeddings.dtype == np.float16
    emb = np.ascontiguousarray(embeddings)
    outputs = np.zeros((L, B, Cc), emb.dtype)
    dy_dx = np.zeros((B, L, D, Cc), emb.dtype) if calc_grad_inputs else np.zeros(1, emb.dtype)
    ls = None if level_scales is None else _f32(level_scales)
    fn = lib().oracle_grid_encode_forward_f16 if half else lib().oracle_grid_encode_forward_f32
    fn(_p(inputs), _p(emb), _p(offsets), _p(outputs), _u32(B), _u32(D), _u32(Cc), _u32(L), _f(S), _u32(base_resolution),
       C.c_int(int(calc_grad_inputs)), _p(dy_dx), _u32(gridtype), None if ls is None else _p(ls))
    return outputs, (dy_dx if calc_grad_inputs else None)


def grid_encode_backward(grad, inputs, offsets, n_entries, Cc, per_level_scale, base_resolution, gridtype=0, half_products=False,
                         level_scales=None):
    """grad [L,B,C] (float32 values) -> grad_embeddings [n_entries, C] in float64 (exact sum)."""
    grad = _f32(grad)
    inputs = _f32(inputs)
    offsets = _i32(offsets)
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    S = np.float32(np.log2(per_level_scale))
    gg = np.zeros((n_entries, Cc), np.float64)
    ls = None if level_scales is None else _f32(level_scales)
    lib().oracle_grid_encode_backward(_p(grad), _p(inputs), _p(offsets), _p(gg), _u32(B), _u32(D), _u32(Cc), _u32(L), _f(S),
                                      _u32(base_resolution), _u32(gridtype), C.c_int(int(half_products)), None if ls is None else _p(ls))
    return gg


def grid_input_backward(grad, dy_dx):
    """grad [L,B,C], dy_dx [B,L,D,C] float32 -> grad_inputs [B,D]"""
    grad, dy_dx = _f32(grad), _f32(dy_dx)
    L, B, Cc = grad.shape
    D = dy_dx.shape[2]
    out = np.zeros((B, D), np.float32)
    lib().oracle_grid_input_backward_f32(_p(grad), _p(dy_dx), _p(out), _u32(B), _u32(D), _u32(Cc), _u32(L))
    return out


# ---------------------------------------------------------------------------------- SH
def sh_encode(dirs, degree=4):
    """degree <= 4: the reference's closed forms (shencoder.cu:51-69) in fp32.
    degree  > 4: scipy's complex harmonics turned into the same real basis in fp64."""
    dirs = _f32(dirs).reshape(-1, 3)
    B = dirs.shape[0]
    if degree <= 4:
        out = np.zeros((B, degree * degree), np.float32)
        lib().oracle_sh_encode_forward_f32(_p(dirs), _p(out), _u32(B), _u32(degree))
        return out
    return sh_encode_scipy(dirs, degree)


def sh_encode_scipy(dirs, degree):
    """Real SH of UNIT vectors from scipy (independent of both implementations)."""
    from scipy.special import sph_harm_y
    d = np.asarray(dirs, dtype=np.float64)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    theta = np.arccos(np.clip(d[:, 2], -1, 1))          # polar
    phi = np.arctan2(d[:, 1], d[:, 0])                  # azimuth
    out = np.zeros((d.shape[0], degree * degree))
    for l in range(degree):
        for m in range(-l, l + 1):
            Y = sph_harm_y(l, abs(m), theta, phi)
            if m == 0:
                v = Y.real
            elif m > 0:
                v = np.sqrt(2) * Y.real
            else:
                v = np.sqrt(2) * Y.imag
            out[:, l * l + l + m] = v
    return out


# ---------------------------------------------------------------------------------- FFMLP
def ffmlp_split(weights, input_dim, hidden_dim, num_layers, out_pad=16):
    """flat vector -> list of [out,in] matrices (ffmlp.cu:631-634, ffmlp.py:118-121)"""
    w = np.asarray(weights)
    mats, o = [], 0
    mats.append(w[o:o + hidden_dim * input_dim].reshape(hidden_dim, input_dim)); o += hidden_dim * input_dim
    for _ in range(num_layers - 1):
        mats.append(w[o:o + hidden_dim * hidden_dim].reshape(hidden_dim, hidden_dim)); o += hidden_dim * hidden_dim
    mats.append(w[o:o + out_pad * hidden_dim].reshape(out_pad, hidden_dim))
    return mats


def ffmlp_forward(x_half, weights_half, input_dim, hidden_dim, num_layers):
    """fp16 operands, exact (float64) accumulation, ReLU, activations rounded to fp16 between
    layers (what is stored in forward_buffer).  Returns (y [B,16] float64 before rounding,
    forward_buffer [num_layers,B,hidden] float16)."""
    mats = [m.astype(np.float64) for m in ffmlp_split(np.asarray(weights_half, np.float16), input_dim, hidden_dim, num_layers)]
    h = np.asarray(x_half, np.float16).astype(np.float64)
    fb = []
    for k in range(num_layers):
        h = np.maximum(h @ mats[k].T, 0.0).astype(np.float16)
        fb.append(h)
        h = h.astype(np.float64)
    y = h @ mats[-1].T
    return y, np.stack(fb)


def ffmlp_backward(grad_half, x_half, weights_half, forward_buffer, input_dim, hidden_dim, num_layers):
    """Returns (grad_inputs [B,in] f64, grad_weights flat f64, backward_buffer [num_layers,B,hidden] f16).
    Activation gradients are rounded to fp16 between layers (they are stored as such)."""
    mats = [m.astype(np.float64) for m in ffmlp_split(np.asarray(weights_half, np.float16), input_dim, hidden_dim, num_layers)]
    g = np.asarray(grad_half, np.float16).astype(np.float64)
    x = np.asarray(x_half, np.float16).astype(np.float64)
    fb = [f.astype(np.float64) for f in forward_buffer]
    dW = [None] * (num_layers + 1)
    dW[num_layers] = g.T @ fb[num_layers - 1]
    bb = []
    gk = ((g @ mats[num_layers]) * (fb[num_layers - 1] > 0)).astype(np.float16)
    bb.append(gk)
    for k in range(num_layers - 1, 0, -1):        # hidden matmul index k maps h_{k-1} -> h_k
        gk64 = gk.astype(np.float64)
        dW[k] = gk64.T @ fb[k - 1]
        gk = ((gk64 @ mats[k]) * (fb[k - 1] > 0)).astype(np.float16)
        bb.append(gk)
    gk64 = gk.astype(np.float64)
    dW[0] = gk64.T @ x
    gx = gk64 @ mats[0]
    return gx, np.concatenate([d.reshape(-1) for d in dW]), np.stack(bb)
