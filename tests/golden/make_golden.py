"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Every array written here is an output of /root/reference code (imported through
ref_harness.py) on seeded inputs; the inputs of the small op-level cases are stored beside
the outputs, the stream case regenerates its frames from oracle.synth_frame (checksummed).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness  # noqa: E402
from oracle import stabstitch_oracle as O  # noqa: E402
from oracle import weights as Wt  # noqa: E402

STREAM_N, STREAM_H, STREAM_W = 9, 180, 320
MESH_SCALE_S, MESH_SCALE_T = 20.0, 10.0


def ops_case(m):
    g = torch.Generator().manual_seed(1234)
    out = {}
    rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    S = m["spatial_network"]
    # DLT + H2Mesh
    src = torch.tensor([[0.0, 0.0], [480.0, 0.0], [0.0, 360.0], [480.0, 360.0]])[None].repeat(3, 1, 1)
    dst = src + 30.0 * rn(3, 4, 2)
    Hm = m["utils.torch_DLT"].tensor_DLT(src, dst)
    out["dlt_src"], out["dlt_dst"], out["dlt_H"] = src, dst, Hm
    out["h2mesh"] = S.H2Mesh(Hm, S.get_rigid_mesh(3, 360, 480))
    # homography feature warp
    U = rn(2, 8, 45, 60)
    theta = torch.eye(3)[None].repeat(2, 1, 1) + 0.05 * rn(2, 3, 3)
    out["homo_U"], out["homo_theta"] = U, theta
    out["homo_out"] = m["utils.torch_homo_transform"].transformer(U, theta, (45, 60))
    # cost volume
    a, b = rn(1, 16, 12, 14), rn(1, 16, 12, 14)
    out["cv_a"], out["cv_b"] = a, b
    out["cv_sr5"] = S.SpatialNet.cost_volume(a, b, search_range=5, norm=False)
    out["cv_sr3"] = S.SpatialNet.cost_volume(a, b, search_range=3, norm=False)
    # CCL (needs the module only for extract_patches)
    net = S.SpatialNet()
    f1, f2 = rn(1, 32, 9, 11), rn(1, 32, 9, 11)
    out["ccl_f1"], out["ccl_f2"] = f1, f2
    out["ccl_out"] = net.CCL(f1, f2)
    # TPS: point and image
    rig = S.get_norm_mesh(S.get_rigid_mesh(2, 360, 480), 360, 480)
    srcm = rig + 0.03 * rn(2, 63, 2)
    pts = rig + 0.02 * rn(2, 63, 2)
    out["tps_rigid"], out["tps_src"], out["tps_pts"] = rig, srcm, pts
    out["tps_point_out"] = m["utils.torch_tps_transform_point"].transformer(pts, rig, srcm)
    img = 127.5 * (1 + torch.tanh(torch.nn.functional.interpolate(rn(2, 3, 6, 8), size=(48, 64), mode="bicubic")))
    out["tps_img"] = img
    # view placed inside a wider canvas, like get_stable_sqe does
    srcc = torch.stack([srcm[..., 0] * 0.55 + torch.tensor([-0.4, 0.4])[:, None], srcm[..., 1] * 0.9], 2)
    out["tps_src_canvas"] = srcc
    out["tps_warp_normal"] = m["utils.torch_tps_transform"].transformer(img, srcc, rig, (40, 100), mode="NORMAL")
    out["tps_warp_fast"] = m["utils.torch_tps_transform"].transformer(img, srcc, rig, (40, 100), mode="FAST")
    return {k: v.detach().numpy() for k, v in out.items()}


def threeview_inputs():
    """Seeded inputs of the three-view case: four smooth meshes @480x360 (pair (1,2) and pair (2,3) of a
    three-camera rig; the shared middle view appears in both pairs with a different absolute position) and
    three 3-frame 96x128 image streams."""
    g = torch.Generator().manual_seed(33)
    N = 3
    rig = O.rigid_mesh(1, 360, 480)[:, None].expand(1, N, 7, 9, 2)

    def mesh(dx, dy, amp):
        return rig + torch.tensor([dx, dy]) + amp * torch.randn(1, N, 7, 9, 2, generator=g)
    w12m1, w12m2 = mesh(-80.0, 3.0, 2.5), mesh(85.0, -2.0, 2.5)
    w23m1, w23m2 = mesh(-70.0, 6.0, 2.5), mesh(95.0, 1.0, 2.5)
    H, W = 96, 128
    imgs = [[O.synth_frame(t, v, H, W) for t in range(N)] for v in range(3)]
    return w12m1, w12m2, w23m1, w23m2, imgs


def threeview_case(m):
    """Runs the reference's OWN three-view glue: the body of test() in test_online_tra_threeview.py between
    '# resize the mesh to the original resolution' and the video writer is executed verbatim (read from
    /root/reference at generation time, never copied into the repository) on the inputs above."""
    import argparse
    import textwrap
    w12m1, w12m2, w23m1, w23m2, imgs = threeview_inputs()
    path = os.path.join(ref_harness.REF_DIR, "test_online_tra_threeview.py")
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if "resize the mesh to the original resolution" in l)
    stop = next(i for i, l in enumerate(lines) if 'print("begin to write into video")' in l)
    body = textwrap.dedent("\n".join(lines[start:stop]))
    D = m["test_online_tra"]
    ns = {"torch": torch, "get_norm_mesh": D.get_norm_mesh, "recover_mesh": D.recover_mesh,
          "get_rigid_mesh": D.get_rigid_mesh, "torch_tps_transform": m["utils.torch_tps_transform"],
          "torch_tps_transform_point": m["utils.torch_tps_transform_point"],
          "args": argparse.Namespace(fusion_mode="AVERAGE", warp_mode="NORMAL"),
          "img1_list": imgs[0], "img2_list": imgs[1], "img3_list": imgs[2],
          "warp12_mesh1": w12m1.clone(), "warp12_mesh2": w12m2.clone(),
          "warp23_mesh1": w23m1.clone(), "warp23_mesh2": w23m2.clone()}
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(body, path, "exec"), ns)
    # the same source slice once more with --fusion_mode LINEAR (:492-503)
    ns_lin = dict(ns)
    ns_lin.update({"args": argparse.Namespace(fusion_mode="LINEAR", warp_mode="NORMAL"), "linear_blender": D.linear_blender,
                   "warp12_mesh1": w12m1.clone(), "warp12_mesh2": w12m2.clone(),
                   "warp23_mesh1": w23m1.clone(), "warp23_mesh2": w23m2.clone()})
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(body, path, "exec"), ns_lin)
    out = {"frames_linear": torch.stack(ns_lin["stable_list"], 0), "in_w12m1": w12m1, "in_w12m2": w12m2, "in_w23m1": w23m1, "in_w23m2": w23m2,
           "mesh1": ns["warp12_mesh1"], "middle": ns["middle_mesh"], "mesh3": ns["warp23_mesh2"],
           "canvas": torch.stack([ns["width_min"], ns["height_min"], ns["out_width"], ns["out_height"]]),
           "frames": torch.stack(ns["stable_list"], 0)}
    return {k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}


def stream_case(m):
    S, T, Sm, D = (m["spatial_network"], m["temporal_network"], m["smooth_network"], m["test_online_tra"])
    tpp = m["utils.torch_tps_transform_point"]
    sds = Wt.spatial_state_dict(mesh_scale=MESH_SCALE_S)
    sdt = Wt.temporal_state_dict(mesh_scale=MESH_SCALE_T)
    sdm = Wt.smooth_state_dict()
    sn, tn, mn = S.SpatialNet().eval(), T.TemporalNet().eval(), Sm.SmoothNet().eval()
    sn.load_state_dict(sds, strict=True)
    tn.load_state_dict(sdt, strict=True)
    mn.load_state_dict(sdm, strict=True)
    N, H, W = STREAM_N, STREAM_H, STREAM_W
    hr = [[O.synth_frame(t, v, H, W) for t in range(N)] for v in range(2)]
    lr = [[O.lowres(x) for x in hr[v]] for v in range(2)]
    out = {"hr_checksum": np.array([float(sum(x.double().sum() for x in hr[v])) for v in range(2)]),
           "lr_checksum": np.array([float(sum(x.double().abs().sum() for x in lr[v])) for v in range(2)])}
    # ---- mirrors the body of test_online_tra.test(), lines 284-399 ----
    sm = [[], []]
    fwd = []
    for k in range(N):
        fwd.append(sn(lr[0][k], lr[1][k]))
        r = S.build_SpatialNet(sn, lr[0][k], lr[1][k])
        sm[0].append(r["motion1"])
        sm[1].append(r["motion2"])
    tm = [T.build_TemporalNet(tn, lr[0])["motion_list"], T.build_TemporalNet(tn, lr[1])["motion_list"]]
    rigid = D.get_rigid_mesh(1, 360, 480)
    nrig = D.get_norm_mesh(rigid, 360, 480)
    smesh, ts = [[], []], [[], []]
    for v in range(2):
        for k in range(N):
            s = rigid + sm[v][k]
            if k == 0:
                t_ = sm[v][k].clone() * 0
            else:
                sp = rigid + sm[v][k - 1]
                tmesh = rigid + tm[v][k]
                moved = tpp.transformer(D.get_norm_mesh(tmesh, 360, 480), nrig, D.get_norm_mesh(sp, 360, 480))
                t_ = D.recover_mesh(moved, 360, 480) - s
            smesh[v].append(s)
            ts[v].append(t_)
    S1 = S2 = None
    win0 = None
    for k in range(N - 6):
        a1 = ts[0][k:k + 7]
        a1[0] = a1[0] * 0
        a2 = ts[1][k:k + 7]
        a2[0] = a2[0] * 0
        o = Sm.build_SmoothNet(mn, a1, a2, smesh[0][k:k + 7], smesh[1][k:k + 7])
        if k == 0:
            S1, S2 = o["smooth_mesh1"], o["smooth_mesh2"]
            win0 = o
        else:
            S1 = torch.cat((S1, o["smooth_mesh1"][:, -1:]), 1)
            S2 = torch.cat((S2, o["smooth_mesh2"][:, -1:]), 1)
    with contextlib.redirect_stdout(io.StringIO()):
        frames, ow, oh = D.get_stable_sqe(hr[0], hr[1], S1, S2, warp_mode="NORMAL", fusion_mode="AVERAGE")
    out["offset_1"] = torch.cat([f[0] for f in fwd], 0)
    out["offset_2_ref"] = torch.cat([f[1] for f in fwd], 0)
    out["offset_2_tgt"] = torch.cat([f[2] for f in fwd], 0)
    for v in range(2):
        out["smotion%d" % (v + 1)] = torch.cat(sm[v], 0)
        out["tmotion%d" % (v + 1)] = torch.cat(tm[v], 0)
        out["tsmotion%d" % (v + 1)] = torch.cat(ts[v], 0)
    for key in ("ori_path1", "smooth_path1", "ori_mesh1", "smooth_mesh1",
                "ori_path2", "smooth_path2", "ori_mesh2", "smooth_mesh2"):
        out["win0_" + key] = win0[key]
    out["smooth_mesh1"], out["smooth_mesh2"] = S1, S2
    out["canvas_hw"] = np.array([int(oh), int(ow)])
    out["frame0"] = frames[0]
    out["frame_last_rows8"] = frames[-1][::8]
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}


def linear_inputs():
    """Seeded inputs of the clean-mask LINEAR case: two overlapping views on a 72x150 canvas with exact 0/1 masks
    (sheared quadrilaterals, so the overlap is not axis aligned) and smooth 3-channel images."""
    g = torch.Generator().manual_seed(77)
    Ho, Wo = 72, 150
    yy, xx = torch.meshgrid(torch.arange(Ho, dtype=torch.float32), torch.arange(Wo, dtype=torch.float32), indexing="ij")
    m1 = ((xx + 0.15 * yy < 95) & (xx - 0.1 * yy > 4) & (yy > 2) & (yy < 66)).float()[None, None]
    m2 = ((xx - 0.12 * yy > 52) & (xx + 0.05 * yy < 146) & (yy > 6) & (yy < 70)).float()[None, None]
    img = lambda: 127.5 * (1 + torch.tanh(torch.nn.functional.interpolate(torch.randn(1, 3, 5, 9, generator=g), size=(Ho, Wo), mode="bicubic")))  # noqa: E731
    return img() * m1, img() * m2, m1, m2


def linear_case(m):
    """LINEAR fusion (test_online_tra.py:34-58,143-150): (i) the reference's linear_blender on exact 0/1 masks, (ii) its
    get_stable_sqe(..., 'LINEAR') on the small stream's reference meshes; plus the centroids it uses (over
    torch.nonzero of the warped masks, i.e. including the out-of-image rounding residues) next to the centroids of
    the masks thresholded at 0.5 - the data behind DESIGN.md's note on that mode."""
    D = m["test_online_tra"]
    out = {}
    ref, tgt, m1, m2 = linear_inputs()
    out["clean_out"] = D.linear_blender(ref, tgt, m1, m2)
    out["clean_mask1"] = D.linear_blender(ref, tgt, m1, m2, mask=True)
    gs = dict(np.load(os.path.join(HERE, "stream_small.npz")))
    N, H, W = STREAM_N, STREAM_H, STREAM_W
    hr = [[O.synth_frame(t, v, H, W) for t in range(N)] for v in range(2)]
    S1, S2 = torch.from_numpy(gs["smooth_mesh1"]), torch.from_numpy(gs["smooth_mesh2"])
    with contextlib.redirect_stdout(io.StringIO()):
        frames, ow, oh = D.get_stable_sqe(hr[0][:2], hr[1][:2], S1[:, :2], S2[:, :2], warp_mode="NORMAL", fusion_mode="LINEAR")
    out["stream_canvas_hw"] = np.array([int(oh), int(ow)])
    out["stream_frame0"] = frames[0]
    # the warped masks of frame 0 as the reference sees them, and the two kinds of centroid
    M1, M2, wmin, hmin, oww, ohh = O.canvas(S1[:, :2], S2[:, :2], H, W)
    nrig = O.norm_mesh(O.rigid_mesh(1, H, W), H, W)
    tps = m["utils.torch_tps_transform"]
    cents = []
    for Mv in (M1, M2):
        tt = torch.stack([Mv[0, 0, ..., 0] - wmin, Mv[0, 0, ..., 1] - hmin], 2)[None]
        msk = tps.transformer(torch.ones(1, 1, H, W), O.norm_mesh(tt, ohh, oww), nrig, (int(ohh.int()), int(oww.int())))[0, 0]
        r, c = torch.nonzero(msk, as_tuple=True)
        rc, cc = torch.nonzero(msk > 0.5, as_tuple=True)
        cents.append([float(r.float().mean()), float(c.float().mean()), float(rc.float().mean()), float(cc.float().mean()),
                      float(len(r)), float(len(rc))])
    out["stream_centroids"] = np.array(cents)   # per view: (row, col) over nonzero, (row, col) over mask > 0.5, counts
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}


def metric_case(m):
    """Metric path (test_metric_ssd.py): the reference's own inter_grid_loss / intra_grid_loss / l_num_loss on seeded
    meshes, and its get_stable_sqe (per-view C = 6 warp at 360x480) on two frames of the small stream."""
    D = m["test_metric_ssd"]
    g = torch.Generator().manual_seed(21)
    rig = O.rigid_mesh(1, 360, 480)[:, None]
    mesh = rig + torch.tensor([30.0, -4.0]) + 6.0 * torch.randn(1, 12, 7, 9, 2, generator=g)
    mesh[:, 3, 2, 5, 0] += 160.0      # one stretched cell, so that the intra-grid term is not zero
    path = torch.cumsum(1.5 * torch.randn(1, 12, 7, 9, 2, generator=g), 1)
    out = {"mesh": mesh, "path": path,
           "inter": torch.stack([D.inter_grid_loss(mesh[:, k:k + 1]) for k in range(12)]),
           "intra": torch.stack([D.intra_grid_loss(mesh[:, k:k + 1]) for k in range(12)]),
           "l2_lag3": D.l_num_loss(path[:, :-6], path[:, 3:-3], 2)}
    lr = [[O.lowres(O.synth_frame(t, v, STREAM_H, STREAM_W)) for t in range(2)] for v in range(2)]
    gs = dict(np.load(os.path.join(HERE, "stream_small.npz")))
    S1, S2 = torch.from_numpy(gs["smooth_mesh1"])[:, :2], torch.from_numpy(gs["smooth_mesh2"])[:, :2]
    with contextlib.redirect_stdout(io.StringIO()):
        l1, l2 = D.get_stable_sqe(lr[0], lr[1], S1, S2)
    out["warp1_frame0"], out["warp2_frame1_rows4"] = l1[0], l2[1][::4]
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    mods = ref_harness.load()
    only = sys.argv[1:]
    if not only or "ops" in only:
        np.savez_compressed(os.path.join(HERE, "ops.npz"), **ops_case(mods))
    if not only or "stream" in only:
        np.savez_compressed(os.path.join(HERE, "stream_small.npz"), **stream_case(mods))
    if not only or "threeview" in only:
        np.savez_compressed(os.path.join(HERE, "threeview.npz"), **threeview_case(mods))
    if not only or "linear" in only:
        np.savez_compressed(os.path.join(HERE, "linear.npz"), **linear_case(mods))
    if not only or "metric" in only:
        np.savez_compressed(os.path.join(HERE, "metric.npz"), **metric_case(mods))
    for f in ("ops.npz", "stream_small.npz", "threeview.npz", "linear.npz", "metric.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
