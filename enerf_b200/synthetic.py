"""Seeded synthetic workloads: cameras on a sphere looking at the origin,
pinhole rays, and an analytic-ball occupancy grid.  numpy only — used by tests and bench.py to
build inputs; no dataset is needed and nothing here is on the timed path.
"""
import numpy as np


def look_at_poses(n_poses, radius, seed=0):
    """[n,4,4] camera-to-world matrices on a sphere of `radius`, -z looking at the origin."""
    rng = np.random.default_rng(seed)
    poses = np.zeros((n_poses, 4, 4), dtype=np.float32)
    for i in range(n_poses):
        v = rng.normal(size=3)
        v /= np.linalg.norm(v)
        pos = v * radius
        fwd = -v                                   # camera looks along -z_cam = fwd
        up = np.array([0.0, 1.0, 0.0]) if abs(v[1]) < 0.95 else np.array([1.0, 0.0, 0.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        up2 = np.cross(right, fwd)
        poses[i, :3, 0] = right
        poses[i, :3, 1] = up2
        poses[i, :3, 2] = -fwd
        poses[i, :3, 3] = pos
        poses[i, 3, 3] = 1.0
    return poses


def pinhole_rays(pose, H, W, fovy_deg=50.0, pixels=None):
    """Unit-direction rays of the given pixels (flat indices; default all) — the math of
    get_rays (nerf/utils.py:161-169): dirs = normalize(((i-cx)/fx, (j-cy)/fy, 1)) @ R^T."""
    focal = H / (2 * np.tan(np.radians(fovy_deg) / 2))
    cx, cy = W / 2, H / 2
    if pixels is None:
        pixels = np.arange(H * W)
    j, i = np.divmod(pixels, W)
    xs = (i + 0.5 - cx) / focal
    ys = (j + 0.5 - cy) / focal
    d = np.stack([xs, ys, np.ones_like(xs)], -1)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    # OpenGL-style pose above (looks along -z): flip y,z of the OpenCV-style pixel direction
    d = d * np.array([1.0, -1.0, -1.0])
    rays_d = d @ pose[:3, :3].T
    rays_o = np.broadcast_to(pose[:3, 3], rays_d.shape)
    return rays_o.astype(np.float32).copy(), rays_d.astype(np.float32).copy()


def random_rays(n_rays, bound, seed=0, n_poses=8, res=800):
    """n_rays rays from n_poses cameras at radius 0.6*bound (inside the aabb, outside the ball)."""
    rng = np.random.default_rng(seed)
    poses = look_at_poses(n_poses, 0.6 * bound, seed)
    per = -(-n_rays // n_poses)
    os_, ds_ = [], []
    for p in poses:
        pix = rng.integers(0, res * res, size=per)
        o, d = pinhole_rays(p, res, res, 50.0, pix)
        os_.append(o)
        ds_.append(d)
    o = np.concatenate(os_)[:n_rays]
    d = np.concatenate(ds_)[:n_rays]
    return np.ascontiguousarray(o), np.ascontiguousarray(d)


def _compact3(x):
    x = x & 0x49249249
    x = (x | (x >> 2)) & 0xc30c30c3
    x = (x | (x >> 4)) & 0x0f00f00f
    x = (x | (x >> 8)) & 0xff0000ff
    x = (x | (x >> 16)) & 0x0000ffff
    return x


def ball_density_grid(bound, cascade, H=128, radius_frac=0.5):
    """density_grid [cascade, H^3] (Morton order): 1.0 where the cell centre (renderer.py:499-506
    mapping, no jitter) lies inside the ball of radius radius_frac*bound, else 0.0."""
    idx = np.arange(H ** 3, dtype=np.uint32)
    coords = np.stack([_compact3(idx), _compact3(idx >> 1), _compact3(idx >> 2)], -1).astype(np.float32)
    xyz = 2 * coords / (H - 1) - 1
    grid = np.zeros((cascade, H ** 3), dtype=np.float32)
    for cas in range(cascade):
        b = min(2 ** cas, bound)
        half = b / H
        c = xyz * (b - half)
        grid[cas] = (np.linalg.norm(c, axis=-1) < radius_frac * bound).astype(np.float32)
    return grid


def packbits_np(grid, thresh=0.01):
    bits = (grid.reshape(-1, 8) > thresh).astype(np.uint8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(-1).astype(np.uint8)


EVENT_H, EVENT_W = 260, 346          # DAVIS-346 sensor, as the mocapDesk2 recordings (TUM-VIE) are undistorted to


def event_frame(bound=2, n_pixels=60000, seed=0):
    """One synthetic event frame laid out as `EventNeRFDataset` stores it (nerf/provider.py:1148-1202): events [E,4] = (x, y, t,
    polarity) grouped by pixel with >= 2 events per pixel, the per-event successor counts, the indices of every pixel's last event,
    and per-event camera poses [E,3,4] = a look-at pose jittered by 0.2 degrees / 1 mm (BASELINE configs[2] / SURVEY.md §8d C3)."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(2, 12, n_pixels)
    pix = rng.choice(EVENT_H * EVENT_W, n_pixels, replace=False)
    xs = np.repeat(pix % EVENT_W, counts).astype(np.float32)
    ys = np.repeat(pix // EVENT_W, counts).astype(np.float32)
    E = int(counts.sum())
    ts = rng.random(E).astype(np.float32)
    pol = rng.choice([-1.0, 1.0], E).astype(np.float32)
    ev = np.stack([xs, ys, ts, pol], 1)
    cum = np.cumsum(counts)
    num_succ = (np.repeat(cum, counts) - np.arange(E) - 1).astype(np.int64)
    base = look_at_poses(1, 0.6 * bound, seed=3)[0]
    ang = np.radians(0.2) * rng.normal(size=(E, 3)).astype(np.float32)
    K = np.zeros((E, 3, 3), np.float32)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ang[:, 2], ang[:, 1], ang[:, 2], -ang[:, 0], -ang[:, 1], ang[:, 0]
    R = (np.eye(3, dtype=np.float32)[None] + K) @ (base[:3, :3] * np.array([1, -1, -1], np.float32))       # OpenCV-style camera axes
    t = base[:3, 3][None] + 1e-3 * rng.normal(size=(E, 3)).astype(np.float32)
    poses = np.concatenate([R, t[:, :, None]], -1).astype(np.float32)
    return ev, num_succ, cum - 1, poses
