"""Pin the CPU oracle (oracle/stabstitch_oracle.py) to the committed golden vectors, which
are outputs of the unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import stabstitch_oracle as O
from oracle import weights as Wt

T = torch.from_numpy
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def close(a, b, tol):
    a = a.detach().numpy() if torch.is_tensor(a) else a
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= tol, np.abs(a - b).max()


def test_dlt_h2mesh(golden_ops):
    g = golden_ops
    H = O.tensor_dlt(T(g["dlt_src"]), T(g["dlt_dst"]))
    close(H, g["dlt_H"], 1e-5)
    close(O.h2mesh(T(g["dlt_H"]), O.rigid_mesh(3, 360, 480)), g["h2mesh"], 1e-3)


def test_homo_warp(golden_ops):
    g = golden_ops
    close(O.homo_warp(T(g["homo_U"]), T(g["homo_theta"]), (45, 60)), g["homo_out"], 1e-5)


def test_cost_volume(golden_ops):
    g = golden_ops
    close(O.cost_volume(T(g["cv_a"]), T(g["cv_b"]), 5), g["cv_sr5"], 1e-6)
    close(O.cost_volume(T(g["cv_a"]), T(g["cv_b"]), 3), g["cv_sr3"], 1e-6)


def test_ccl(golden_ops):
    g = golden_ops
    close(O.ccl(T(g["ccl_f1"]), T(g["ccl_f2"])), g["ccl_out"], 1e-5)


def test_ccl_known_answers():
    # SURVEY.md 8c known-answer identities: CCL(f,f)=0, CCL(f, roll(f,(1,2)))=(2,1) interior
    g = torch.Generator().manual_seed(5)
    f = torch.randn(1, 64, 12, 16, generator=g)
    assert O.ccl(f, f).abs().max() < 1e-5
    fl = O.ccl(f, torch.roll(f, (1, 2), (2, 3)))[0, :, 3:-3, 3:-3]
    assert (fl[0] - 2).abs().max() < 1e-3 and (fl[1] - 1).abs().max() < 1e-3


def test_tps_point_and_warp(golden_ops):
    g = golden_ops
    close(O.tps_point(T(g["tps_pts"]), T(g["tps_rigid"]), T(g["tps_src"])), g["tps_point_out"], 1e-6)
    w = O.tps_warp(T(g["tps_img"]), T(g["tps_src_canvas"]), T(g["tps_rigid"]), (40, 100), "NORMAL")
    close(w, g["tps_warp_normal"], 1e-4)
    w = O.tps_warp(T(g["tps_img"]), T(g["tps_src_canvas"]), T(g["tps_rigid"]), (40, 100), "FAST")
    close(w, g["tps_warp_fast"], 1e-4)


def test_tps_identity():
    rig = O.norm_mesh(O.rigid_mesh(1, 360, 480), 360, 480)
    close(O.tps_point(rig, rig, rig), rig.numpy(), 1e-5)
    assert (O.tensor_dlt(torch.tensor([[[0., 0.], [4, 0], [0, 3], [4, 3]]]),
                         torch.tensor([[[0., 0.], [4, 0], [0, 3], [4, 3]]])) - torch.eye(3)).abs().max() < 1e-5


def test_fp64_arbiter_agrees(golden_ops):
    g = golden_ops
    src, tgt = T(g["tps_src_canvas"]), T(g["tps_rigid"])
    ax, ay = O.tps_source_coords_fp64(src, tgt, 40, 100, 64, 48)
    Tm = O.tps_solve(src, tgt)
    xt = torch.linspace(-1, 1, 100)[None, :].expand(40, 100).reshape(-1)
    yt = torch.linspace(-1, 1, 40)[:, None].expand(40, 100).reshape(-1)
    xs, ys = O.tps_eval(Tm, src, xt, yt)
    assert np.abs((xs.numpy().reshape(2, 40, 100) + 1) * 32 - ax).max() < 1e-3
    assert np.abs((ys.numpy().reshape(2, 40, 100) + 1) * 24 - ay).max() < 1e-3


def test_stream(golden_stream):
    from tests.golden.make_golden import STREAM_N, STREAM_H, STREAM_W, MESH_SCALE_S, MESH_SCALE_T
    g = golden_stream
    N, H, W = STREAM_N, STREAM_H, STREAM_W
    hr = [[O.synth_frame(t, v, H, W) for t in range(N)] for v in range(2)]
    lr = [[O.lowres(x) for x in hr[v]] for v in range(2)]
    cs = np.array([float(sum(x.double().sum() for x in hr[v])) for v in range(2)])
    assert np.allclose(cs, g["hr_checksum"], rtol=1e-9), "synthetic frames differ from the fixture's"
    with torch.no_grad():
        out = O.stitch_stream(Wt.spatial_state_dict(mesh_scale=MESH_SCALE_S),
                              Wt.temporal_state_dict(mesh_scale=MESH_SCALE_T), Wt.smooth_state_dict(),
                              lr[0], lr[1], hr[0], hr[1])
    close(torch.cat(out["smotion1"], 0), g["smotion1"], 2e-4)
    close(torch.cat(out["smotion2"], 0), g["smotion2"], 2e-4)
    close(torch.cat(out["tmotion1"], 0), g["tmotion1"], 2e-4)
    close(torch.cat(out["tsmotion2"], 0), g["tsmotion2"], 5e-4)
    close(out["smooth_mesh1"], g["smooth_mesh1"], 5e-4)
    close(out["smooth_mesh2"], g["smooth_mesh2"], 5e-4)
    assert tuple(g["canvas_hw"]) == out["canvas"]
    d = np.abs(out["frames"][0] - g["frame0"])
    # hard image edges flip single pixels when the mesh moves by 1e-5 px: bound the fraction
    assert (d > 0.05).mean() < 1e-3, (d > 0.05).mean()
    d = np.abs(out["frames"][-1][::8] - g["frame_last_rows8"])
    assert (d > 0.05).mean() < 1e-3


def test_three_view(golden_threeview):
    """the oracle's three-view glue against the reference's own (test_online_tra_threeview.py:345-505,
    executed verbatim by tests/golden/make_golden.py::threeview_case)"""
    from tests.golden.make_golden import threeview_inputs
    g = golden_threeview
    w12m1, w12m2, w23m1, w23m2, imgs = threeview_inputs()
    close(w12m1, g["in_w12m1"], 0.0)
    with torch.no_grad():
        m1, mid, m3, wmin, hmin, ow, oh = O.three_view_meshes(w12m1, w12m2, w23m1, w23m2, 96, 128)
        close(m1, g["mesh1"], 1e-4)
        close(mid, g["middle"], 1e-4)
        close(m3, g["mesh3"], 1e-4)
        assert np.abs(np.array([float(wmin), float(hmin), float(ow), float(oh)]) - g["canvas"]).max() < 1e-4
        for k in range(3):
            f = O.three_view_frame(imgs[0][k], imgs[1][k], imgs[2][k], m1[:, k], mid[:, k], m3[:, k], wmin, hmin, ow, oh)
            d = np.abs(f.numpy() - g["frames"][k])
            assert (d > 0.05).mean() < 1e-3, (d > 0.05).mean()
            assert np.median(d) < 1e-4


def test_host_edges_against_cv2():
    """the uint8 front end of the reference (cv2.resize INTER_LINEAR on uint8, test_online_tra.py:259) restated in
    oracle/host_edges.py: bit-exact against cv2 itself for the reference's down-scaling shapes"""
    cv2 = __import__("pytest").importorskip("cv2")
    from oracle import host_edges as HE
    rng = np.random.default_rng(0)
    for (h, w) in ((720, 1280), (1080, 1920), (361, 483), (480, 640)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(HE.resize_u8_linear(img, 480, 360), cv2.resize(img, (480, 360))), (h, w)
    img = rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    hr, lr = HE.load_frame(img)
    ref_lr = np.transpose(cv2.resize(img, (480, 360)).astype(np.float32), [2, 0, 1]) / 127.5 - 1.0
    assert hr.shape == (1, 3, 720, 1280) and np.array_equal(hr[0, :, 5, 7], img[5, 7].astype(np.float32))
    assert np.array_equal(lr[0], ref_lr.astype(np.float32))
    f = np.array([[[-0.05, 0.0, 0.99]], [[1.5, 254.99, 255.0]]], np.float32)
    assert np.array_equal(HE.to_video_frame(f), f.astype(np.uint8))


def test_linear_fusion_against_reference(golden_stream, golden_threeview):
    """LINEAR fusion: the oracle's verbatim restatement (clean=False) of linear_blender / the LINEAR branches of
    get_stable_sqe (test_online_tra.py:34-58,143-150) and of the three-view driver (test_online_tra_threeview.py:492-503)
    against the reference's own outputs: bit-exact on this torch build.  And the measured distance of the residue-free
    mask semantics (clean=True, what the CUDA path implements) to the reference."""
    from tests.golden.make_golden import linear_inputs, threeview_inputs, STREAM_H, STREAM_W
    gl = dict(np.load(os.path.join(GOLDEN, "linear.npz")))
    ref, tgt, m1, m2 = linear_inputs()
    assert np.array_equal(O.linear_blender(ref, tgt, m1, m2).numpy(), gl["clean_out"])
    assert np.array_equal(O.linear_blender(ref, tgt, m1, m2, mask=True).numpy(), gl["clean_mask1"])
    assert np.array_equal(O.linear_blender(ref, tgt, m1, m2, clean=True).numpy(), gl["clean_out"])   # 0/1 masks: same
    hr = [[O.synth_frame(t, v, STREAM_H, STREAM_W) for t in range(2)] for v in range(2)]
    S1, S2 = T(golden_stream["smooth_mesh1"])[:, :2], T(golden_stream["smooth_mesh2"])[:, :2]
    M1, M2, wmin, hmin, ow, oh = O.canvas(S1, S2, STREAM_H, STREAM_W)
    f = O.stable_frame_linear(hr[0][0], hr[1][0], M1[:, 0], M2[:, 0], wmin, hmin, ow, oh)
    assert np.abs(f.numpy().transpose(1, 2, 0) - gl["stream_frame0"]).max() < 1e-4
    fc = O.stable_frame_linear(hr[0][0], hr[1][0], M1[:, 0], M2[:, 0], wmin, hmin, ow, oh, clean=True)
    d = np.abs(fc.numpy().transpose(1, 2, 0) - gl["stream_frame0"])
    c = gl["stream_centroids"]
    print("residue-free masks vs reference LINEAR frame: max %.3f p99 %.3f; centre columns %.1f/%.1f (nonzero) vs %.1f/%.1f "
          "(mask > 0.5); %d / %d 'mask' pixels of view 1 are residues" % (d.max(), np.percentile(d, 99), c[0, 1], c[1, 1],
                                                                         c[0, 3], c[1, 3], c[0, 4] - c[0, 5], c[0, 4]))
    assert d.max() < 1.0 and np.percentile(d, 99) < 0.5
    w12m1, w12m2, w23m1, w23m2, imgs = threeview_inputs()
    with torch.no_grad():
        a, mid, b, wmin, hmin, ow, oh = O.three_view_meshes(w12m1, w12m2, w23m1, w23m2, 96, 128)
        f3 = O.three_view_frame_linear(imgs[0][0], imgs[1][0], imgs[2][0], a[:, 0], mid[:, 0], b[:, 0], wmin, hmin, ow, oh)
    assert np.abs(f3.numpy() - golden_threeview["frames_linear"][0]).max() < 1e-4


def test_nview_reduces_to_three_view():
    """the N-view middle-plane chain (config 5) is the reference's three-view glue for N = 3: bit-identical meshes,
    canvas and frames; and N = 4 is self-consistent (a chain of three pairs whose shared views already coincide
    leaves them in place)."""
    from tests.golden.make_golden import threeview_inputs
    w12m1, w12m2, w23m1, w23m2, imgs = threeview_inputs()
    with torch.no_grad():
        m1, mid, m3, wmin, hmin, ow, oh = O.three_view_meshes(w12m1, w12m2, w23m1, w23m2, 96, 128)
        out, wmin2, hmin2, ow2, oh2 = O.nview_meshes([(w12m1, w12m2), (w23m1, w23m2)], 96, 128)
        for a, b in zip(out, (m1, mid, m3)):
            assert torch.equal(a, b)
        assert float(wmin) == float(wmin2) and float(hmin) == float(hmin2) and float(ow) == float(ow2) and float(oh) == float(oh2)
        f3 = O.three_view_frame(imgs[0][0], imgs[1][0], imgs[2][0], m1[:, 0], mid[:, 0], m3[:, 0], wmin, hmin, ow, oh)
        fn = O.nview_frame([imgs[v][0] for v in range(3)], [m[:, 0] for m in out], wmin2, hmin2, ow2, oh2)
        assert torch.equal(f3, fn)
        # N = 4: pair (3,4) built so that its instance of view 3 equals pair (2,3)'s up to a constant shift
        g = torch.Generator().manual_seed(5)
        w34m1 = w23m2 + torch.tensor([7.0, -3.0])
        w34m2 = w23m2 + torch.tensor([150.0, 2.0]) + 2.0 * torch.randn(w23m2.shape, generator=g)
        out4, a, b, c, d = O.nview_meshes([(w12m1, w12m2), (w23m1, w23m2), (w34m1, w34m2)], 96, 128)
        assert len(out4) == 4 and all(t.shape == w12m1.shape for t in out4)
        # shared view 3: both instances coincide after the chain shift, so its middle plane is pair (2,3)'s (shifted) mesh
        out3, *_ = O.nview_meshes([(w12m1, w12m2), (w23m1, w23m2)], 96, 128)
        # (the provisional canvases differ, compare shapes via differences that are translation invariant)
        d4 = out4[2] - out4[1]
        d3 = None
        hr = lambda m: torch.stack([m[..., 0] * 128 / 480, m[..., 1] * 96 / 360], 4)  # noqa: E731
        off = (hr(w12m2) - hr(w23m1)).reshape(1, 3, -1, 2).mean(2)[:, :, None, None, :]
        d3 = (hr(w23m2) + off) - (hr(w12m2) + hr(w23m1) + off) / 2.0
        assert (d4 - d3).abs().max() < 2e-4


def test_metric_path_against_reference(golden_stream):
    """metric path (test_metric_ssd.py): the oracle's restatement of the grid losses, the squared-lag loss and the
    per-view C = 6 warp against the reference's own functions; PSNR / SSIM (skimage, absent) against plain numpy."""
    from oracle import metric_oracle as MO
    from tests.golden.make_golden import STREAM_H, STREAM_W
    g = dict(np.load(os.path.join(GOLDEN, "metric.npz")))
    mesh, path = T(g["mesh"]), T(g["path"])
    for k in range(mesh.shape[1]):
        assert abs(float(MO.inter_grid_loss(mesh[:, k:k + 1])) - g["inter"][k]) < 1e-7
        assert abs(float(MO.intra_grid_loss(mesh[:, k:k + 1])) - g["intra"][k]) < 1e-6
    assert g["intra"].max() > 0.1
    assert abs(float(MO.l_num_loss(path[:, :-6], path[:, 3:-3], 2)) - float(g["l2_lag3"])) < 1e-6
    assert MO.distortion_score(mesh) == float(np.max(g["inter"] + g["intra"]))
    lr = [[O.lowres(O.synth_frame(t, v, STREAM_H, STREAM_W)) for t in range(2)] for v in range(2)]
    S1, S2 = T(golden_stream["smooth_mesh1"])[:, :2], T(golden_stream["smooth_mesh2"])[:, :2]
    w1 = MO.metric_warp(lr[0][0], S1[:, 0]).numpy().transpose(1, 2, 0)
    w2 = MO.metric_warp(lr[1][1], S2[:, 1]).numpy().transpose(1, 2, 0)
    assert np.abs(w1 - g["warp1_frame0"]).max() < 1e-4
    assert np.abs(w2[::4] - g["warp2_frame1_rows4"]).max() < 1e-4
    # PSNR / SSIM sanity (skimage itself is not available: parity unpinned, see the module header)
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.uniform(0, 255, (40, 50, 3)), np.ones((40, 50, 3))], 2).astype(np.float32)
    b = a.copy()
    b[..., :3] += rng.normal(0, 5, (40, 50, 3)).astype(np.float32)
    assert abs(MO.psnr_overlap(a, b) - 10 * np.log10(255 ** 2 / np.mean((a[..., :3].astype(np.float64) - b[..., :3]) ** 2))) < 1e-9
    assert 0.5 < MO.ssim_overlap(a, b) < 1.0 and abs(MO.ssim_overlap(a, a) - 1.0) < 1e-12
