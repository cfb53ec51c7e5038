"""Replay of the reference driver's per-frame call sequence through the DROP-IN names (test / measurement harness).

What an unchanged `Full_model_inference/Codes/test_online_tra.py` does between frame loading and the video writer
(:284-399), restated call for call - batch-1 `build_SpatialNet` per frame pair, `build_TemporalNet` on the CPU tensor
lists, the tsmotion loop with `torch_tps_transform_point.transformer` per frame and view, `build_SmoothNet` per
7-frame window with torch.cat growth, and `get_stable_sqe`'s loop (global canvas with torch reductions, one
`torch_tps_transform.transformer` call per frame on the concatenated views, the AVERAGE blend as torch ops, a
`.cpu().numpy()` per frame) - with every name resolved the way the driver resolves it: flat imports
(`from spatial_network import ...`, `import utils.torch_tps_transform`), which find stabstitch2_b200/dropin when that
directory is first on sys.path.  The driver-local helpers (rigid / normalised meshes,
the loop bodies) are torch code of the DRIVER, not of the replaced modules, and are torch code here too.

Used by tests/test_gpu_parity.py::test_dropin_replay_matches_batched_path and by bench.py's `dropin_replay` leg (the
fps a maintainer gets with zero edits)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "stabstitch2_b200", "dropin")
FLAT = ("spatial_network", "temporal_network", "smooth_network", "grid_res", "utils", "utils.torch_tps_transform",
        "utils.torch_tps_transform_point", "utils.torch_DLT", "utils.torch_homo_transform")


def import_flat():
    """the driver's import block (test_online_tra.py:7-9,14-15,21-23) with dropin/ first on the path"""
    saved = {k: sys.modules.pop(k) for k in FLAT if k in sys.modules}
    sys.path.insert(0, DROPIN)
    if ROOT not in sys.path:
        sys.path.insert(1, ROOT)
    try:
        from spatial_network import build_SpatialNet, SpatialNet
        from temporal_network import build_TemporalNet, TemporalNet
        from smooth_network import build_SmoothNet, SmoothNet
        import utils.torch_tps_transform as torch_tps_transform
        import utils.torch_tps_transform_point as torch_tps_transform_point
        import grid_res
        names = dict(build_SpatialNet=build_SpatialNet, SpatialNet=SpatialNet, build_TemporalNet=build_TemporalNet,
                     TemporalNet=TemporalNet, build_SmoothNet=build_SmoothNet, SmoothNet=SmoothNet,
                     torch_tps_transform=torch_tps_transform, torch_tps_transform_point=torch_tps_transform_point,
                     grid_res=grid_res)
    finally:
        sys.path.remove(DROPIN)
        for k in FLAT:
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    return names


def _rigid(h, w, gh, gw, dev):
    ys = torch.linspace(0, h, gh + 1, device=dev)[:, None].expand(gh + 1, gw + 1)
    xs = torch.linspace(0, w, gw + 1, device=dev)[None, :].expand(gh + 1, gw + 1)
    return torch.stack([xs, ys], 2)[None]


def _norm(mesh, h, w):
    return torch.stack([mesh[..., 0] * 2.0 / w - 1.0, mesh[..., 1] * 2.0 / h - 1.0], 3).reshape(mesh.shape[0], -1, 2)


def _recover(nm, h, w, gh, gw):
    return torch.stack([(nm[..., 0] + 1) * w / 2.0, (nm[..., 1] + 1) * h / 2.0], 2).reshape(nm.shape[0], gh + 1, gw + 1, 2)


def replay(names, spatial_net, temporal_net, smooth_net, lr1, lr2, hr1, hr2, warp_mode="NORMAL"):
    """lr*, hr*: lists of CPU tensors [1,3,360,480] / [1,3,H,W] exactly like the driver builds them.  Returns
    (list of [Ho,Wo,3] numpy frames, smooth_mesh1, smooth_mesh2 [1,N,7,9,2])."""
    N = len(lr1)
    gh, gw = names["grid_res"].GRID_H, names["grid_res"].GRID_W
    with torch.no_grad():
        sm1, sm2 = [], []
        for k in range(N):                                            # :284-292, batch 1, H2D per frame
            o = names["build_SpatialNet"](spatial_net, lr1[k].cuda(), lr2[k].cuda())
            sm1.append(o["motion1"])
            sm2.append(o["motion2"])
        tm1 = names["build_TemporalNet"](temporal_net, lr1)["motion_list"]   # :295-299, CPU tensor lists
        tm2 = names["build_TemporalNet"](temporal_net, lr2)["motion_list"]
        dev = sm1[0].device
        rigid = _rigid(360, 480, gh, gw, dev)
        nrigid = _norm(rigid, 360, 480)
        meshes, tsm = ([], []), ([], [])
        tpp = names["torch_tps_transform_point"].transformer
        for k in range(N):                                            # :315-347
            for v, (sm, tm) in enumerate(((sm1, tm1), (sm2, tm2))):
                mesh = rigid + sm[k]
                if k == 0:
                    ts = sm[k].clone() * 0
                else:
                    moved = tpp(_norm(rigid + tm[k], 360, 480), nrigid, _norm(rigid + sm[k - 1], 360, 480))
                    ts = _recover(moved, 360, 480, gh, gw) - mesh
                meshes[v].append(mesh)
                tsm[v].append(ts)
        S1 = S2 = None
        for k in range(N - 6):                                        # :359-392
            a1, a2 = tsm[0][k:k + 7], tsm[1][k:k + 7]
            a1[0], a2[0] = a1[0] * 0, a2[0] * 0
            o = names["build_SmoothNet"](smooth_net, a1, a2, meshes[0][k:k + 7], meshes[1][k:k + 7])
            if k == 0:
                S1, S2 = o["smooth_mesh1"], o["smooth_mesh2"]
            else:
                S1 = torch.cat((S1, o["smooth_mesh1"][:, -1:]), 1)
                S2 = torch.cat((S2, o["smooth_mesh2"][:, -1:]), 1)
        # get_stable_sqe (:96-154), AVERAGE fusion
        _, _, H, W = hr2[0].shape
        nrig_hr = _norm(_rigid(H, W, gh, gw, dev), H, W)
        M1 = torch.stack([S1[..., 0] * W / 480, S1[..., 1] * H / 360], 4)
        M2 = torch.stack([S2[..., 0] * W / 480, S2[..., 1] * H / 360], 4)
        wmin = torch.minimum(M1[..., 0].min(), M2[..., 0].min())
        wmax = torch.maximum(M1[..., 0].max(), M2[..., 0].max())
        hmin = torch.minimum(M1[..., 1].min(), M2[..., 1].min())
        hmax = torch.maximum(M1[..., 1].max(), M2[..., 1].max())
        ow, oh = wmax - wmin, hmax - hmin
        tps = names["torch_tps_transform"].transformer
        frames = []
        for i in range(N):
            n1 = _norm(torch.stack([M1[:, i, ..., 0] - wmin, M1[:, i, ..., 1] - hmin], 3), oh, ow)
            n2 = _norm(torch.stack([M2[:, i, ..., 0] - wmin, M2[:, i, ..., 1] - hmin], 3), oh, ow)
            w = tps(torch.cat([hr1[i].cuda(), hr2[i].cuda()], 0), torch.cat([n1, n2], 0), torch.cat([nrig_hr, nrig_hr], 0),
                    (oh.int(), ow.int()), mode=warp_mode)
            s = w[0] + w[1] + 1e-6
            fused = w[0] * (w[0] / s) + w[1] * (w[1] / s)
            frames.append(fused.cpu().numpy().transpose(1, 2, 0))
    return frames, S1, S2
