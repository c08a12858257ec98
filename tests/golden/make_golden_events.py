#!/usr/bin/env python
"""Golden vectors for the next-row functions, produced by the REFERENCE's own Python code.

Runs in the build container only (needs /root/reference): the source of `get_rays`, `get_event_rays`, `custom_meshgrid`
(nerf/utils.py) and `rgb_to_luma`, `lin_log` (utils/event_utils.py) is extracted with `ast` and executed on the CPU (the
modules themselves do not import here: matplotlib, h5py, tensorboardX ... are missing), and the event-loss expressions of
`Trainer.train_step_events` (nerf/utils.py:494-528) are evaluated with those functions.  Writes tests/golden/events.npz.
"""
import ast
import os

import numpy as np
import torch

REF = os.environ.get("ENERF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_functions(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    import packaging.version as pver                # nerf/utils.py:31
    ns = {"torch": torch, "np": np, "pver": pver}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []                      # @torch.cuda.amp.autocast(enabled=False), @torch.jit.script
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def main():
    u = load_functions(os.path.join(REF, "nerf", "utils.py"), ["custom_meshgrid", "get_rays", "get_event_rays"])
    e = load_functions(os.path.join(REF, "utils", "event_utils.py"), ["rgb_to_luma", "lin_log"])
    g = torch.Generator().manual_seed(1234)
    out = {}

    # ---- get_rays: 3 poses, 40x56 image, all pixels and 257 chosen pixels
    H, W = 40, 56
    intr = np.array([61.5, 59.25, 27.3, 19.8], np.float32)
    A = torch.randn(3, 3, 3, generator=g)
    R = torch.linalg.qr(A)[0]
    poses = torch.zeros(3, 4, 4)
    poses[:, :3, :3] = R
    poses[:, :3, 3] = torch.randn(3, 3, generator=g) * 0.5
    poses[:, 3, 3] = 1
    r = u["get_rays"](poses, intr, H, W, -1)
    out.update(gr_poses=poses.numpy(), gr_intr=intr, gr_hw=np.array([H, W]), gr_o_all=r["rays_o"].numpy(), gr_d_all=r["rays_d"].numpy())
    torch.manual_seed(7)
    r = u["get_rays"](poses, intr, H, W, 257)
    out.update(gr_inds=r["inds"][0].numpy(), gr_o_sel=r["rays_o"].numpy(), gr_d_sel=r["rays_d"].numpy())

    # ---- get_event_rays: 1000 events, a pose per event
    N = 1000
    xs = torch.randint(0, 346, (N,), generator=g).float()
    ys = torch.randint(0, 260, (N,), generator=g).float()
    intr_e = np.array([250.1, 249.7, 172.4, 131.9], np.float32)

    def rand_poses(n):
        q = torch.linalg.qr(torch.randn(n, 3, 3, generator=g))[0]
        p = torch.zeros(1, n, 3, 4)
        p[0, :, :, :3] = q
        p[0, :, :, 3] = torch.randn(n, 3, generator=g) * 0.3
        return p
    pb, pa = rand_poses(N), rand_poses(N)
    r = u["get_event_rays"](xs, ys, pb, pa, intr_e)
    out.update(er_xs=xs.numpy(), er_ys=ys.numpy(), er_pb=pb.numpy(), er_pa=pa.numpy(), er_intr=intr_e,
               **{"er_" + k[9:]: v.numpy() for k, v in r.items()})

    # ---- event loss, nerf/utils.py:494-528, for every (use_luma, linlog, C_thres, event_only) the configs can select
    Nl = 512
    img1 = torch.rand(1, Nl, 3, generator=g) * 0.9 + 0.02
    img2 = (img1 + torch.randn(1, Nl, 3, generator=g) * 0.05).clamp(0.005, 1.0)
    img1[0, :64] *= 0.05                                   # some pixels in the linear part of lin_log (255*x < 20)
    pols = torch.randint(-8, 9, (1, Nl), generator=g).float()
    out.update(el_img1=img1.numpy(), el_img2=img2.numpy(), el_pols=pols.numpy())
    log_thres = torch.tensor(20.0)
    cases = []
    for use_luma in (0, 1):
        for linlog in (1, 0):
            for C_thres in (-1.0, 0.25):
                for event_only in (1, 0):
                    if linlog == 0 and use_luma == 1:
                        continue                          # the reference evaluates pred_luma1 twice there (utils.py:504-505): constant loss
                    a = img1.clone().requires_grad_(True)
                    b = img2.clone().requires_grad_(True)
                    if use_luma:
                        l1, l2 = e["rgb_to_luma"](a, esim=True), e["rgb_to_luma"](b, esim=True)
                    else:
                        l1, l2 = a, b
                    if linlog:
                        p1, p2 = e["lin_log"](l1 * 255, linlog_thres=20), e["lin_log"](l2 * 255, linlog_thres=20)
                    else:
                        p1, p2 = torch.log(torch.maximum(l1 * 255, log_thres)), torch.log(torch.maximum(l2 * 255, log_thres))
                    w_evLoss = 1
                    delta_linlog = (p2 - p1)
                    gt_pol = pols[..., None]
                    if C_thres != -1:
                        loss_evs = w_evLoss * torch.mean((delta_linlog - gt_pol * C_thres) ** 2)
                    else:
                        EPS = 1e-9
                        w_evLoss *= 20
                        if not event_only:
                            w_evLoss *= 20
                        dn = delta_linlog / (torch.linalg.norm(delta_linlog, dim=1, keepdim=True) + EPS)
                        pn = gt_pol / (torch.linalg.norm(gt_pol, dim=1, keepdim=True) + EPS)
                        loss_evs = w_evLoss * torch.mean((dn - pn) ** 2)
                    loss_evs.backward()
                    tag = f"el_{use_luma}{linlog}{int(C_thres != -1)}{event_only}"
                    cases.append((use_luma, linlog, C_thres, event_only))
                    out[tag + "_loss"] = np.float64(loss_evs.item())
                    out[tag + "_delta"] = delta_linlog.detach().numpy()
                    out[tag + "_g1"] = a.grad.numpy()
                    out[tag + "_g2"] = b.grad.numpy()
    out["el_cases"] = np.array(cases, np.float64)
    np.savez_compressed(os.path.join(HERE, "events.npz"), **out)
    print("wrote events.npz:", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "events.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
