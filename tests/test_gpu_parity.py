"""GPU parity tests: the CUDA path (through the Python shims -> C ABI of libss2.so) against
(1) the committed golden vectors, which are outputs of the unmodified reference, and
(2) the CPU oracle on the same seeded inputs.

Tolerances (fp32 path; all stated in the unit of the quantity):
  - geometry ops, TPS coefficients/points: 1e-5 relative-ish absolute on normalised coords
  - warped pixels.  The reference evaluates the TPS field in fp32: against the fp64 arbiter its
    own source coordinate is off by up to ~1.4e-3 px (mean 1e-4 px) on the 720p case below, so
    two correct fp32 implementations differ by (image gradient) x ~2e-3 px.  Hence
      * SOURCE COORDINATES are the primary check: our error against the arbiter must not exceed
        the reference's own (max: + 2e-4 px slack, mean: x1.2 + 2e-5 px);
      * pixels are checked at north_star's 1e-3 abs on SMOOTH frames (gradient <= 0.3 grey/px),
        where 1e-3 is resolvable, and through gradient-scaled bounds on textured frames;
      * pixels on the reference's hard image edge (taps clamp there) flip between
        a value and 0 when the coordinate moves by 1e-4 px: bounded as a fraction (< 5e-4).
  - mesh vertices (networks): MESH_TOL_PX at 480x360 against the reference's CPU fp32 outputs (split-TF32 tensor-core
    convolutions: 2^-21 relative per product + TMEM accumulation order); the measured values are printed by the tests
    and recorded in DESIGN.md section 2.
"""
import os

import numpy as np
import pytest
import torch

from oracle import stabstitch_oracle as O
from oracle import weights as Wt

pytestmark = pytest.mark.gpu

T = torch.from_numpy


def dev(a):
    return (T(a) if isinstance(a, np.ndarray) else a).cuda()


def maxdiff(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else b
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max())


def grad_max(img):
    """largest one-pixel step of an image tensor [..,H,W]: bound of |d value / d px|"""
    t = img if torch.is_tensor(img) else T(img)
    return float(max((t[..., 1:, :] - t[..., :-1, :]).abs().max(), (t[..., :, 1:] - t[..., :, :-1]).abs().max()))


COORD_TOL_PX = 1.5e-3  # two fp32 evaluations of the TPS field (see header)
MESH_TOL_PX = 1e-3     # network outputs (mesh vertices, px at 480x360) against the fp32 CPU reference; measured <= 4.3e-4 (printed)
STREAM_FRAC = 2e-4     # fraction of pixels of a small-stream frame off by > 0.05 grey levels (hard-edge flips); measured 0


@pytest.fixture(scope="module")
def nets():
    from stabstitch2_b200.spatial_network import SpatialNet
    from stabstitch2_b200.temporal_network import TemporalNet
    from stabstitch2_b200.smooth_network import SmoothNet
    from tests.golden.make_golden import MESH_SCALE_S, MESH_SCALE_T
    s, t, m = SpatialNet().cuda().eval(), TemporalNet().cuda().eval(), SmoothNet().cuda().eval()
    s.load_state_dict(Wt.spatial_state_dict(mesh_scale=MESH_SCALE_S), strict=True)
    t.load_state_dict(Wt.temporal_state_dict(mesh_scale=MESH_SCALE_T), strict=True)
    m.load_state_dict(Wt.smooth_state_dict(), strict=True)
    return s, t, m


@pytest.fixture(scope="module")
def stream_inputs():
    from tests.golden.make_golden import STREAM_N, STREAM_H, STREAM_W
    hr = [[O.synth_frame(t, v, STREAM_H, STREAM_W) for t in range(STREAM_N)] for v in range(2)]
    lr = [[O.lowres(x) for x in hr[v]] for v in range(2)]
    return hr, lr


# ---------------------------------------------------------------- ops vs golden (reference outputs)
def test_dlt_golden(golden_ops):
    from stabstitch2_b200.utils.torch_DLT import tensor_DLT
    H = tensor_DLT(dev(golden_ops["dlt_src"]), dev(golden_ops["dlt_dst"]))
    assert maxdiff(H, golden_ops["dlt_H"]) < 2e-5


def test_homo_warp_golden(golden_ops):
    from stabstitch2_b200.utils.torch_homo_transform import transformer
    out = transformer(dev(golden_ops["homo_U"]), dev(golden_ops["homo_theta"]), (45, 60))
    d = np.abs(out.cpu().numpy() - golden_ops["homo_out"])
    assert (d > 1e-4).mean() < 2e-3, ((d > 1e-4).mean(), d.max())


def test_cost_volume_golden(golden_ops):
    from stabstitch2_b200.spatial_network import SpatialNet
    a, b = dev(golden_ops["cv_a"]), dev(golden_ops["cv_b"])
    assert maxdiff(SpatialNet.cost_volume(a, b, 5, norm=False), golden_ops["cv_sr5"]) < 1e-5
    assert maxdiff(SpatialNet.cost_volume(a, b, 3, norm=False), golden_ops["cv_sr3"]) < 1e-5


def test_ccl_golden(golden_ops):
    from stabstitch2_b200.spatial_network import SpatialNet
    net = SpatialNet()
    out = net.CCL(dev(golden_ops["ccl_f1"]), dev(golden_ops["ccl_f2"]))
    assert maxdiff(out, golden_ops["ccl_out"]) < 1e-4


def test_ccl_known_answers():
    from stabstitch2_b200.spatial_network import SpatialNet
    g = torch.Generator().manual_seed(5)
    f = torch.randn(1, 64, 12, 16, generator=g)
    net = SpatialNet()
    assert net.CCL(f, f).abs().max().item() < 1e-4
    fl = net.CCL(f, torch.roll(f, (1, 2), (2, 3)))[0, :, 3:-3, 3:-3].cpu()
    assert (fl[0] - 2).abs().max() < 1e-3 and (fl[1] - 1).abs().max() < 1e-3


def test_tps_point_golden(golden_ops):
    from stabstitch2_b200.utils.torch_tps_transform_point import transformer
    g = golden_ops
    out = transformer(dev(g["tps_pts"]), dev(g["tps_rigid"]), dev(g["tps_src"]))
    assert maxdiff(out, g["tps_point_out"]) < 2e-6
    rig = dev(g["tps_rigid"])
    assert maxdiff(transformer(rig, rig, rig), g["tps_rigid"]) < 2e-6  # identity


@pytest.mark.parametrize("mode", ["NORMAL", "FAST"])
def test_tps_warp_golden(golden_ops, mode):
    from stabstitch2_b200.utils.torch_tps_transform import transformer
    g = golden_ops
    out = transformer(dev(g["tps_img"]), dev(g["tps_src_canvas"]), dev(g["tps_rigid"]), (40, 100), mode=mode)
    d = np.abs(out.cpu().numpy() - g["tps_warp_" + mode.lower()])
    # 64x48 image stretched over a 100x40 canvas with steps of up to ~40 grey levels per pixel
    bound = grad_max(g["tps_img"]) * COORD_TOL_PX + 1e-3
    assert (d > bound).mean() < 2e-3 and np.median(d) < 1e-3, ((d > bound).mean(), d.max(), bound)


def test_tps_warp_tensor_out_size_and_channels(golden_ops):
    """out_size given as 0-dim CUDA int tensors (test_online_tra.py:140) and C=4 (LINEAR's mask channel)."""
    from stabstitch2_b200.utils.torch_tps_transform import transformer
    g = golden_ops
    img = dev(g["tps_img"])
    img4 = torch.cat([img, torch.ones_like(img[:, :1])], 1)
    oh, ow = torch.tensor(40.7).cuda().int(), torch.tensor(100.2).cuda().int()
    out = transformer(img4, dev(g["tps_src_canvas"]), dev(g["tps_rigid"]), (oh, ow))
    assert out.shape == (2, 4, 40, 100)
    d = np.abs(out[:, :3].cpu().numpy() - g["tps_warp_normal"])
    assert (d > grad_max(g["tps_img"]) * COORD_TOL_PX + 1e-3).mean() < 2e-3
    ref_mask = O.tps_warp(torch.ones(2, 1, 48, 64), T(g["tps_src_canvas"]), T(g["tps_rigid"]), (40, 100))
    dm = (out[:, 3:].cpu() - ref_mask).abs()
    assert (dm > 1e-3).float().mean() < 2e-3


def test_empty_and_errors():
    from stabstitch2_b200 import _lib
    from stabstitch2_b200.utils.torch_tps_transform import transformer
    from stabstitch2_b200.utils.torch_DLT import tensor_DLT
    assert tensor_DLT(torch.zeros(0, 4, 2), torch.zeros(0, 4, 2)).shape == (0, 3, 3)
    rig = O.norm_mesh(O.rigid_mesh(1, 360, 480), 360, 480)
    assert transformer(torch.zeros(1, 3, 8, 8), rig, rig, (0, 5)).shape == (1, 3, 0, 5)
    with pytest.raises(ValueError):
        transformer(torch.zeros(1, 3, 8, 8), rig[:, :60], rig, (4, 4))
    ctx = _lib.context()
    rc = ctx.lib.ss2_tps_warp(ctx.handle, None, None, None, 1, 3, 8, 8, 4, 4, 0, 0, None, None)
    assert rc == -1 and b"bad arguments" in ctx.lib.ss2_last_error(ctx.handle)


# ---------------------------------------------------------------- networks vs golden
def test_spatial_forward_golden(nets, stream_inputs, golden_stream):
    s, _, _ = nets
    _, lr = stream_inputs
    a, b = torch.cat(lr[0], 0).cuda(), torch.cat(lr[1], 0).cuda()
    o1, oref, otgt = s(a, b)
    g = golden_stream
    e = (maxdiff(o1, g["offset_1"]), maxdiff(oref, g["offset_2_ref"]), maxdiff(otgt, g["offset_2_tgt"]))
    print("SpatialNet.forward vs reference: offset_1 %.2e, offset_2_ref %.2e, offset_2_tgt %.2e px" % e)
    assert e[0] < MESH_TOL_PX         # 4-pt offsets, px @480 (magnitude ~170)
    assert e[1] < MESH_TOL_PX         # mesh residuals, px (magnitude ~5)
    assert e[2] < MESH_TOL_PX


def test_build_spatial_batch_equals_single(nets, stream_inputs, golden_stream):
    from stabstitch2_b200.spatial_network import build_SpatialNet
    s, _, _ = nets
    _, lr = stream_inputs
    g = golden_stream
    r = build_SpatialNet(s, torch.cat(lr[0], 0), torch.cat(lr[1], 0))  # CPU tensors in, like the driver
    e = (maxdiff(r["motion1"], g["smotion1"]), maxdiff(r["motion2"], g["smotion2"]))
    print("build_SpatialNet vs reference: motion1 %.2e, motion2 %.2e px" % e)
    assert max(e) < MESH_TOL_PX
    r1 = build_SpatialNet(s, lr[0][3], lr[1][3])
    assert maxdiff(r1["motion1"], r["motion1"][3:4]) < 1e-4
    assert maxdiff(r1["motion2"], r["motion2"][3:4]) < 1e-4


def test_temporal_golden(nets, stream_inputs, golden_stream):
    from stabstitch2_b200.temporal_network import build_TemporalNet
    _, t, _ = nets
    _, lr = stream_inputs
    for v in range(2):
        ml = build_TemporalNet(t, lr[v])["motion_list"]
        assert len(ml) == len(lr[v]) and ml[0].abs().max().item() == 0.0
        e = maxdiff(torch.cat(ml, 0), golden_stream["tmotion%d" % (v + 1)])
        print("build_TemporalNet view %d vs reference: %.2e px" % (v + 1, e))
        assert e < MESH_TOL_PX


def test_temporal_pair_bit_identical_to_per_view(nets, stream_inputs):
    """ss2_build_temporal_pair (both views as one batch: what the whole-stream calls use) == two ss2_build_temporal
    calls, bit for bit - the property the sharded stream relies on (batch composition never changes a frame's bits)"""
    from stabstitch2_b200 import _lib
    from stabstitch2_b200.temporal_network import build_TemporalNet
    _, t, _ = nets
    _, lr = stream_inputs
    per_view = [torch.cat(build_TemporalNet(t, lr[v])["motion_list"], 0) for v in range(2)]
    ctx = _lib.context()
    a, b = torch.cat(lr[0], 0).cuda(), torch.cat(lr[1], 0).cuda()
    n = a.shape[0]
    ma, mb = torch.empty(n, 7, 9, 2, device="cuda"), torch.empty(n, 7, 9, 2, device="cuda")
    ctx.check(ctx.lib.ss2_build_temporal_pair(ctx.handle, _lib.ptr(a), _lib.ptr(b), n, _lib.ptr(ma), _lib.ptr(mb), _lib.cur_stream()))
    torch.cuda.synchronize()
    assert torch.equal(ma.cpu(), per_view[0].cpu()) and torch.equal(mb.cpu(), per_view[1].cpu())


def test_smooth_window_golden(nets, golden_stream):
    from stabstitch2_b200.smooth_network import build_SmoothNet
    _, _, m = nets
    g = golden_stream
    rig = O.rigid_mesh(1, 360, 480)
    sm = [[rig + T(g["smotion%d" % v][k:k + 1]) for k in range(7)] for v in (1, 2)]
    ts = [[T(g["tsmotion%d" % v][k:k + 1]) * (0.0 if k == 0 else 1.0) for k in range(7)] for v in (1, 2)]
    o = build_SmoothNet(m, ts[0], ts[1], sm[0], sm[1])
    for key in ("ori_path1", "smooth_path1", "ori_mesh1", "smooth_mesh1",
                "ori_path2", "smooth_path2", "ori_mesh2", "smooth_mesh2"):
        e = maxdiff(o[key], g["win0_" + key])
        print("build_SmoothNet %s vs reference: %.2e px" % (key, e))
        assert e < MESH_TOL_PX, key


def test_stream_golden(nets, stream_inputs, golden_stream):
    """Whole hot path on the small stream: meshes, canvas size, fused frames."""
    from stabstitch2_b200 import pipeline
    s, t, m = nets
    hr, lr = stream_inputs
    g = golden_stream
    fused, s1, s2 = pipeline.stitch_stream(s, t, m, torch.cat(lr[0], 0).cuda(), torch.cat(lr[1], 0).cuda(),
                                           torch.cat(hr[0], 0).cuda(), torch.cat(hr[1], 0).cuda())
    e = (maxdiff(s1[None], g["smooth_mesh1"]), maxdiff(s2[None], g["smooth_mesh2"]))
    print("whole stream vs reference: smooth_mesh1 %.2e, smooth_mesh2 %.2e px" % e)
    assert max(e) < MESH_TOL_PX
    assert tuple(fused.shape[2:]) == tuple(g["canvas_hw"])
    f0 = fused[0].permute(1, 2, 0).cpu().numpy()
    d = np.abs(f0 - g["frame0"])
    # mesh differences of ~1e-3 px move pixel values by gradient*1e-3; edges flip single pixels
    fl = fused[-1].permute(1, 2, 0).cpu().numpy()[::8]
    fr = ((d > 0.05).mean(), (np.abs(fl - g["frame_last_rows8"]) > 0.05).mean())
    print("whole stream vs reference: frame 0 frac(|d| > 0.05) %.2e median %.2e, last frame frac %.2e" % (fr[0], np.median(d), fr[1]))
    assert fr[0] < STREAM_FRAC, (fr[0], d.max())
    assert np.median(d) < 1e-3
    assert fr[1] < STREAM_FRAC


def test_stream_host_call_matches_device_path(nets, stream_inputs):
    from stabstitch2_b200 import pipeline
    s, t, m = nets
    hr, lr = stream_inputs
    ins = [torch.cat(x, 0).contiguous().pin_memory() for x in (lr[0], lr[1], hr[0], hr[1])]
    fused, s1, s2 = pipeline.stitch_stream(s, t, m, *[x.cuda() for x in ins])
    out = torch.empty(fused.numel()).pin_memory()
    ho, wo, m1, m2 = pipeline.stitch_stream_host(s, t, m, *ins, out, want_meshes=True)
    assert (ho, wo) == tuple(fused.shape[2:])
    assert maxdiff(m1, s1) == 0.0 and maxdiff(m2, s2) == 0.0
    assert maxdiff(out.view_as(fused), fused) == 0.0


def test_stream_host_prefetch_pipeline(nets, stream_inputs):
    """two chunks in flight with the next chunk's upload prefetched: every chunk bit-identical to the blocking call,
    also when the prefetched buffers differ from the ones finally submitted (the copy is then repeated)"""
    from stabstitch2_b200 import pipeline
    s, t, m = nets
    hr, lr = stream_inputs
    ins = [torch.cat(x, 0).contiguous().pin_memory() for x in (lr[0], lr[1], hr[0], hr[1])]
    ins_b = [x.flip(0).contiguous().pin_memory() for x in ins]  # a second, different chunk
    fused, _, _ = pipeline.stitch_stream(s, t, m, *[x.cuda() for x in ins])
    fused_b, _, _ = pipeline.stitch_stream(s, t, m, *[x.cuda() for x in ins_b])
    outs = [torch.empty(max(fused.numel(), fused_b.numel()) + 1024).pin_memory() for _ in range(2)]
    chunks = [ins, ins_b, ins, ins_b]
    sizes = []
    for i, c in enumerate(chunks):
        if i + 1 < len(chunks):
            pipeline.stitch_stream_host_prefetch((i + 1) & 1, *chunks[i + 1])
        if i >= 2:  # results of the chunk that used this slot two steps ago
            pipeline.stitch_stream_host_wait(i & 1)
            ref = fused if (i - 2) % 2 == 0 else fused_b
            assert maxdiff(outs[i & 1][:ref.numel()].view_as(ref), ref) == 0.0
        sizes.append(pipeline.stitch_stream_host_async(s, t, m, i & 1, *c, outs[i & 1]))
    for slot, ref in ((0, fused), (1, fused_b)):
        pipeline.stitch_stream_host_wait(slot)
        assert maxdiff(outs[slot][:ref.numel()].view_as(ref), ref) == 0.0
    assert sizes[0] == tuple(fused.shape[2:]) and sizes[1] == tuple(fused_b.shape[2:])
    # prefetch one chunk, submit another: the stale prefetch must not be used
    pipeline.stitch_stream_host_prefetch(0, *ins_b)
    pipeline.stitch_stream_host_async(s, t, m, 0, *ins, outs[0])
    pipeline.stitch_stream_host_wait(0)
    assert maxdiff(outs[0][:fused.numel()].view_as(fused), fused) == 0.0


def test_three_view_and_prefetch_errors():
    from stabstitch2_b200 import _lib
    ctx = _lib.context()
    rc = ctx.lib.ss2_three_view_meshes(ctx.handle, None, None, None, None, 1, 96, 128, None, None, None, None, None)
    assert rc == -1 and b"bad arguments" in ctx.lib.ss2_last_error(ctx.handle)
    import ctypes
    cv = (ctypes.c_float * 4)(0.0, 0.0, 10.0, 10.0)
    rc = ctx.lib.ss2_three_view_frames(ctx.handle, None, None, None, None, None, None, 1, 96, 128, cv, 0, 0, None, None)
    assert rc == -1
    rc = ctx.lib.ss2_three_view_frames(ctx.handle, None, None, None, None, None, None, 0, 96, 128, cv, 0, 0, None, None)
    assert rc == 0  # no frames: nothing to do
    rc = ctx.lib.ss2_stitch_stream_host_prefetch(ctx.handle, 5, None, None, None, None, 8, 96, 128)
    assert rc == -1
    rc = ctx.lib.ss2_stitch_stream_host_prefetch(ctx.handle, 0, None, None, None, None, 8, 96, 128)
    assert rc == -1 and b"bad arguments" in ctx.lib.ss2_last_error(ctx.handle)


def test_get_stable_sqe_dropin(golden_stream, stream_inputs):
    """Reference-shaped call (lists of CPU tensors, [1,N,7,9,2] meshes) with the golden meshes:
    isolates the resampler+blend from the networks."""
    from stabstitch2_b200.pipeline import get_stable_sqe
    hr, _ = stream_inputs
    g = golden_stream
    frames, ow, oh = get_stable_sqe(hr[0], hr[1], T(g["smooth_mesh1"]), T(g["smooth_mesh2"]), "NORMAL", "AVERAGE")
    assert (int(oh), int(ow)) == tuple(g["canvas_hw"]) and len(frames) == len(hr[0])
    d = np.abs(frames[0] - g["frame0"])
    bound = grad_max(hr[0][0]) * COORD_TOL_PX + 1e-3
    assert (d > bound).mean() < 2e-3 and np.median(d) < 1e-3, ((d > bound).mean(), d.max(), bound)
    assert (d > 0.05).mean() < 5e-4


# ---------------------------------------------------------------- full-size checks (720p)
def _canvas_case(H=720, W=1280):
    """Synthetic smooth meshes shaped like SURVEY.md 8d's probe: view 2 ~35% to the right."""
    g = torch.Generator().manual_seed(11)
    rig = O.rigid_mesh(1, 360, 480)
    m1 = rig + torch.tensor([-86.0, 0.0]) + 3.0 * torch.randn(1, 7, 9, 2, generator=g)
    m2 = rig + torch.tensor([86.0, 0.0]) + 3.0 * torch.randn(1, 7, 9, 2, generator=g)
    return m1, m2


def _smooth_frame(seed, H, W):
    """frame with gradients <= ~0.3 grey/px, on which north_star's 1e-3 abs is resolvable"""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(1, 3, 3, 3, generator=g)
    big = torch.nn.functional.interpolate(z, size=(H, W), mode="bicubic", align_corners=True)
    return ((torch.tanh(0.3 * big) + 1.0) * 127.5).contiguous()


@pytest.mark.parametrize("tps,H,W", [("exact", 720, 1280), ("lattice", 720, 1280), ("lattice", 1080, 1920)])
def test_fullsize_frame_vs_oracle_and_arbiter(tps, H, W):
    """BASELINE.json configs 2 and 3 (720p, 1080p): the resampler + blend at full size.  The lattice cases run the
    production kernels (tps_solve / tps_nodes / tps_warp_lattice, one compile-time instantiation per source size)."""
    from stabstitch2_b200 import _lib, pipeline
    from stabstitch2_b200.utils.torch_tps_transform import transformer
    mode = _lib.TPS_EXACT if tps == "exact" else _lib.TPS_LATTICE
    m1, m2 = _canvas_case(H, W)
    hr1, hr2 = O.synth_frame(0, 0, H, W), O.synth_frame(0, 1, H, W)
    M1, M2, wmin, hmin, ow, oh = O.canvas(m1[None], m2[None], H, W)
    fused_ref, warp_ref = O.stable_frame(hr1, hr2, M1[:, 0], M2[:, 0], wmin, hmin, ow, oh)
    mm = pipeline.canvas_minmax(m1, m2, H, W).cpu().tolist()
    assert abs(mm[0] - float(wmin)) < 1e-4 and abs(mm[2] - float(hmin)) < 1e-4
    fused = pipeline.stable_frames(hr1.cuda(), hr2.cuda(), m1, m2, mm, tps=mode)[0]
    assert tuple(fused.shape) == tuple(fused_ref.shape)
    Ho, Wo = fused.shape[1:]
    # ---- coordinate space: ours vs the fp64 arbiter must be as good as the reference vs the arbiter
    nrig = O.norm_mesh(O.rigid_mesh(1, H, W), H, W)
    ramp = torch.stack([torch.arange(W, dtype=torch.float32)[None, :].expand(H, W),
                        torch.arange(H, dtype=torch.float32)[:, None].expand(H, W),
                        torch.zeros(H, W)], 0)[None]
    inside_all = np.ones((Ho, Wo), bool)
    edge_any = np.zeros((Ho, Wo), bool)
    for v, M in enumerate((M1, M2)):
        tt = torch.stack([M[0, 0, ..., 0] - wmin, M[0, 0, ..., 1] - hmin], 2)[None]
        src = O.norm_mesh(tt, oh, ow)
        ax, ay = O.tps_source_coords_fp64(src, nrig, Ho, Wo, W, H)
        got = transformer(ramp.cuda(), src.cuda(), nrig.cuda(), (Ho, Wo), tps=mode).cpu().numpy()[0]
        ref = O.tps_warp(ramp, src, nrig, (Ho, Wo)).numpy()[0]
        inside = (ax[0] > 1) & (ax[0] < W - 2) & (ay[0] > 1) & (ay[0] < H - 2)
        inside_all &= inside
        edge_any |= ((np.abs(ax[0]) < 1) | (np.abs(ax[0] - (W - 1)) < 1) | (np.abs(ay[0]) < 1) | (np.abs(ay[0] - (H - 1)) < 1))
        for got_c, ref_c, a in ((got[0], ref[0], ax[0]), (got[1], ref[1], ay[0])):
            e_got, e_ref = np.abs(got_c - a)[inside], np.abs(ref_c - a)[inside]
            print("view %d %s coord err vs fp64 (px): ours max %.2e mean %.2e | reference max %.2e mean %.2e"
                  % (v, tps, e_got.max(), e_got.mean(), e_ref.max(), e_ref.mean()))
            assert e_got.max() < 1.5 * e_ref.max() + 2e-4   # the max of rounding noise is itself noisy
            assert e_got.mean() < 1.2 * e_ref.mean() + 2e-5
    # ---- the same check through the FUSED kernel: view v carries the coordinate ramp, the other view is black, so
    # fused = a*a/(a+1e-6) ~ a
    zero = torch.zeros_like(ramp)
    for v, M in enumerate((M1, M2)):
        tt = torch.stack([M[0, 0, ..., 0] - wmin, M[0, 0, ..., 1] - hmin], 2)[None]
        src = O.norm_mesh(tt, oh, ow)
        ax, ay = O.tps_source_coords_fp64(src, nrig, Ho, Wo, W, H)
        ref = O.tps_warp(ramp, src, nrig, (Ho, Wo)).numpy()[0]
        a, b = (ramp, zero) if v == 0 else (zero, ramp)
        got = pipeline.stable_frames(a.cuda(), b.cuda(), m1, m2, mm, tps=mode)[0].cpu().numpy()
        inside = (ax[0] > 2) & (ax[0] < W - 2) & (ay[0] > 2) & (ay[0] < H - 2)
        for got_c, ref_c, arb in ((got[0], ref[0], ax[0]), (got[1], ref[1], ay[0])):
            e_got, e_ref = np.abs(got_c - arb)[inside], np.abs(ref_c - arb)[inside]
            print("view %d %s FUSED coord err vs fp64 (px): ours max %.2e mean %.2e | reference max %.2e mean %.2e"
                  % (v, tps, e_got.max(), e_got.mean(), e_ref.max(), e_ref.mean()))
            assert e_got.max() < 1.5 * e_ref.max() + 2e-4
            assert e_got.mean() < 1.2 * e_ref.mean() + 2e-5
    # ---- pixel space on the textured frame: gradient-scaled bound, hard-edge flips as a fraction
    d = (fused.cpu() - fused_ref).abs().numpy()
    bound = max(grad_max(hr1), grad_max(hr2)) * COORD_TOL_PX + 1e-3
    assert (d > bound).mean() < 5e-4, ((d > bound).mean(), d.max(), bound)
    assert np.median(d) < 2e-3 and (d > 0.05).mean() < 5e-4
    # ---- north_star's 1e-3 abs on smooth frames, away from the hard image edges
    s1, s2 = _smooth_frame(1, H, W), _smooth_frame(2, H, W)
    f_ref, _ = O.stable_frame(s1, s2, M1[:, 0], M2[:, 0], wmin, hmin, ow, oh)
    f_got = pipeline.stable_frames(s1.cuda(), s2.cuda(), m1, m2, mm, tps=mode)[0].cpu()
    ds = (f_got - f_ref).abs().numpy()
    keep = ~edge_any
    # Where a view is outside its image the reference's clamped taps cancel only up to a rounding
    # residue of ~ulp(255 * distance) ~ 1e-2 grey levels, which enters its blend;
    # no re-implementation can reproduce those bits.  So: 1e-3 where BOTH views are inside (the
    # overlap, where the blend mixes two warped images), residue-sized bound elsewhere.
    both = keep & inside_all
    print("%s smooth frames (grad_max %.2f): overlap max |diff| %.2e p99.9 %.2e | elsewhere max %.2e"
          % (tps, grad_max(s1), ds[:, both].max(), np.percentile(ds[:, both], 99.9), ds[:, keep].max()))
    assert both.mean() > 0.15
    assert ds[:, both].max() < 1e-3
    # (the residue is a few ulp of 255 x |distance to the image| and grows with the frame: measured 0.04 at 720p,
    # 0.12 at 1080p)
    assert ds[:, keep].max() < (6e-2 if W <= 1280 else 0.15)


def test_fullsize_properties():
    """Size-independent properties at BASELINE's 720p size."""
    from stabstitch2_b200.utils.torch_tps_transform import transformer, warp_blend_average
    H, W = 720, 1280
    nrig = O.norm_mesh(O.rigid_mesh(1, H, W), H, W).cuda()
    img = O.synth_frame(3, 0, H, W).cuda()
    # identity mesh: out == in resampled at (x+1)*W/2 with x on linspace(-1,1,W): compare with the oracle
    out = transformer(img, nrig, nrig, (H, W))
    ref = O.tps_warp(img.cpu(), nrig.cpu(), nrig.cpu(), (H, W))
    d = (out.cpu() - ref).abs()
    bound = grad_max(img) * COORD_TOL_PX + 1e-3
    assert (d > bound).float().mean() < 5e-4 and (d > 0.05).float().mean() < 5e-4, ((d > bound).float().mean(), d.max())
    # linearity in the image
    a = transformer(img * 0.5, nrig, nrig, (300, 500))
    b = transformer(img, nrig, nrig, (300, 500))
    assert (a - 0.5 * b).abs().max().item() < 1e-3
    # blending a view with itself returns the view: a*(a/(2a+eps)) + a*(a/(2a+eps)) ~= a
    st = torch.stack([nrig[0], nrig[0]], 0)[None]
    both = warp_blend_average(img, img, st, st, (300, 500))
    rel = (both - b).abs() / (b.abs() + 1.0)
    assert rel.max().item() < 1e-3
    # fused kernel == generic kernel + torch blend
    m1, m2 = _canvas_case(H, W)
    img2 = O.synth_frame(3, 1, H, W).cuda()
    M1, M2, wmin, hmin, ow, oh = O.canvas(m1[None], m2[None], H, W)
    srcs = []
    for M in (M1, M2):
        t = torch.stack([M[0, 0, ..., 0] - wmin, M[0, 0, ..., 1] - hmin], 2)[None]
        srcs.append(O.norm_mesh(t, oh, ow))
    src = torch.cat(srcs, 0).cuda()
    tgt = torch.cat([nrig, nrig], 0)
    Ho, Wo = int(oh.int()), int(ow.int())
    w = transformer(torch.cat([img, img2], 0), src, tgt, (Ho, Wo))
    fused = warp_blend_average(img, img2, src[None], tgt[None], (Ho, Wo))[0]
    s = w[0] + w[1] + 1e-6
    # The fused kernel evaluates both views of a pixel in one thread, the generic kernel one view per launch
    # slice: two fp32 roundings of the same field (each checked against the fp64
    # arbiter in test_fullsize_frame_vs_oracle_and_arbiter), so the bound is gradient x coordinate noise;
    # a sample that flips across the hard image edge is counted as a fraction.
    d = (fused - (w[0] * (w[0] / s) + w[1] * (w[1] / s))).abs()
    bound = max(grad_max(img), grad_max(img2)) * COORD_TOL_PX + 1e-3
    assert (d > bound).float().mean().item() < 2e-4, ((d > bound).float().mean().item(), d.max().item(), bound)
    assert d.median().item() < 1e-4


@pytest.mark.parametrize("mode", ["NORMAL", "FAST"])
def test_stable_frames_u8_fused_store_bit_identical(mode):
    """ss2_stable_frames_u8 (the uint8 back end fused into the lattice resampler's store, what the uint8 host pipeline
    runs at 720p) == ss2_stable_frames + ss2_frames_to_u8, byte for byte; the unsupported cases fail loudly."""
    from stabstitch2_b200 import _lib, pipeline
    H, W = 720, 1280
    m1, m2 = _canvas_case(H, W)
    m1, m2 = torch.cat([m1, m1 + 0.7], 0), torch.cat([m2, m2 - 0.4], 0)
    hr1 = torch.cat([O.synth_frame(0, 0, H, W), O.synth_frame(1, 0, H, W)], 0).cuda()
    hr2 = torch.cat([O.synth_frame(0, 1, H, W), O.synth_frame(1, 1, H, W)], 0).cuda()
    mm = pipeline.canvas_minmax(m1, m2, H, W).cpu().tolist()
    fused = pipeline.stable_frames(hr1, hr2, m1, m2, mm, mode=mode, tps=_lib.TPS_LATTICE)
    ref = pipeline.frames_to_u8(fused)
    got = pipeline.stable_frames_u8(hr1, hr2, m1, m2, mm, mode=mode, tps=_lib.TPS_LATTICE)
    assert got.dtype == torch.uint8 and tuple(got.shape) == tuple(ref.shape)
    assert torch.equal(got, ref)
    assert float(got.float().mean()) > 20.0   # not a blank canvas
    with pytest.raises(_lib.SS2Error):
        pipeline.stable_frames_u8(hr1, hr2, m1, m2, mm, mode=mode, tps=_lib.TPS_EXACT)
    small = [0.0, 100.0, 0.0, 60.0]   # a 60 x 100 canvas: too coarse for the lattice
    with pytest.raises(_lib.SS2Error):
        pipeline.stable_frames_u8(hr1, hr2, m1, m2, small, mode=mode, tps=_lib.TPS_LATTICE)


@pytest.mark.parametrize("halo", [0, 1])
def test_spatial_temporal_one_call_bit_identical(nets, halo):
    """ss2_build_spatial_temporal (TemporalNet on a stream and workspace arena of its own next to SpatialNet) == the two
    separate calls, bit for bit; repeated to catch a race between the two streams."""
    from stabstitch2_b200 import _lib, pipeline
    from stabstitch2_b200.spatial_network import build_SpatialNet
    s, t, m = nets
    n = 9
    lr1 = torch.cat([O.lowres(O.synth_frame(k, 0, 360, 480)) for k in range(n + halo)], 0).cuda()
    lr2 = torch.cat([O.lowres(O.synth_frame(k, 1, 360, 480)) for k in range(n + halo)], 0).cuda()
    sp = build_SpatialNet(s, lr1[halo:], lr2[halo:])
    ctx = _lib.context()
    tm = [torch.empty(n + halo, 7, 9, 2, device="cuda") for _ in range(2)]
    ctx.check(ctx.lib.ss2_build_temporal_pair(ctx.handle, _lib.ptr(lr1), _lib.ptr(lr2), n + halo, _lib.ptr(tm[0]), _lib.ptr(tm[1]),
                                              _lib.cur_stream()))
    for _ in range(3):
        sm1, sm2, tm1, tm2 = pipeline.build_spatial_temporal(s, t, lr1, lr2, halo)
        assert torch.equal(sm1, sp["motion1"].reshape(n, 7, 9, 2)) and torch.equal(sm2, sp["motion2"].reshape(n, 7, 9, 2))
        assert torch.equal(tm1, tm[0]) and torch.equal(tm2, tm[1])
    assert float(sm1.abs().max()) > 0 and float(tm1[1:].abs().max()) > 0


def test_stream_host_u8_720p_fused_store(nets):
    """The uint8 host pipeline at 720p (lattice resampler, so the fused uint8 store runs) == device-side edges + fp32
    stream + conversion pass, byte for byte."""
    from stabstitch2_b200 import pipeline
    s, t, m = nets
    n, H, W = 7, 720, 1280
    u1, u2 = _synth_u8(2, n, H, W), _synth_u8(3, n, H, W)
    hr1, lr1 = pipeline.load_frames_u8(u1)
    hr2, lr2 = pipeline.load_frames_u8(u2)
    fused, s1, s2 = pipeline.stitch_stream(s, t, m, lr1, lr2, hr1, hr2)
    ref = pipeline.frames_to_u8(fused).cpu()
    out = torch.empty(fused.numel(), dtype=torch.uint8).pin_memory()
    ho, wo, m1, m2 = pipeline.stitch_stream_host_u8(s, t, m, u1.pin_memory(), u2.pin_memory(), out, want_meshes=True)
    assert (ho, wo) == tuple(fused.shape[2:])
    assert maxdiff(m1, s1) == 0.0 and maxdiff(m2, s2) == 0.0
    assert torch.equal(out.reshape(n, ho, wo, 3), ref)


def test_fullsize_fast_mode_lattice():
    """mode='FAST' (F.grid_sample, align_corners=True; utils/torch_tps_transform.py:158-162) with the LATTICE field
    at 720p: fused frame against the oracle, and the source coordinates of the generic single-view FAST kernel
    against the fp64 arbiter."""
    from stabstitch2_b200 import _lib, pipeline
    from stabstitch2_b200.utils.torch_tps_transform import transformer
    H, W = 720, 1280
    m1, m2 = _canvas_case(H, W)
    hr1, hr2 = O.synth_frame(0, 0, H, W), O.synth_frame(0, 1, H, W)
    M1, M2, wmin, hmin, ow, oh = O.canvas(m1[None], m2[None], H, W)
    fused_ref, _ = O.stable_frame(hr1, hr2, M1[:, 0], M2[:, 0], wmin, hmin, ow, oh, mode="FAST")
    mm = pipeline.canvas_minmax(m1, m2, H, W).cpu().tolist()
    fused = pipeline.stable_frames(hr1.cuda(), hr2.cuda(), m1, m2, mm, mode="FAST", tps=_lib.TPS_LATTICE)[0]
    assert tuple(fused.shape) == tuple(fused_ref.shape)
    d = (fused.cpu() - fused_ref).abs().numpy()
    bound = max(grad_max(hr1), grad_max(hr2)) * COORD_TOL_PX + 1e-3
    print("FAST lattice 720p: frac > bound %.2e, median %.2e, max %.2e" % ((d > bound).mean(), np.median(d), d.max()))
    # grid_sample feathers the image border (zeros padding) instead of cutting it: no hard-edge flips here
    assert (d > bound).mean() < 2e-4 and np.median(d) < 2e-3
    Ho, Wo = fused.shape[1:]
    nrig = O.norm_mesh(O.rigid_mesh(1, H, W), H, W)
    ramp = torch.stack([torch.arange(W, dtype=torch.float32)[None, :].expand(H, W),
                        torch.arange(H, dtype=torch.float32)[:, None].expand(H, W),
                        torch.zeros(H, W)], 0)[None]
    for v, M in enumerate((M1, M2)):
        tt = torch.stack([M[0, 0, ..., 0] - wmin, M[0, 0, ..., 1] - hmin], 2)[None]
        src = O.norm_mesh(tt, oh, ow)
        ax, ay = O.tps_source_coords_fp64(src, nrig, Ho, Wo, W - 1, H - 1)   # align_corners: (x+1)/2*(W-1)
        got = transformer(ramp.cuda(), src.cuda(), nrig.cuda(), (Ho, Wo), mode="FAST", tps=_lib.TPS_LATTICE).cpu().numpy()[0]
        ref = O.tps_warp(ramp, src, nrig, (Ho, Wo), mode="FAST").numpy()[0]
        inside = (ax[0] > 1) & (ax[0] < W - 2) & (ay[0] > 1) & (ay[0] < H - 2)
        for got_c, ref_c, a in ((got[0], ref[0], ax[0]), (got[1], ref[1], ay[0])):
            e_got, e_ref = np.abs(got_c - a)[inside], np.abs(ref_c - a)[inside]
            print("view %d FAST coord err vs fp64 (px): ours max %.2e mean %.2e | reference max %.2e mean %.2e"
                  % (v, e_got.max(), e_got.mean(), e_ref.max(), e_ref.mean()))
            assert e_got.max() < 1.5 * e_ref.max() + 2e-4
            assert e_got.mean() < 1.2 * e_ref.mean() + 2e-5


def test_canvas_minmax_bit_exact_and_truncation_window(golden_stream):
    """Canvas extents (test_online_tra.py:103-120) are data dependent and TRUNCATED to the output shape (:140): with
    identical meshes our fp32 min/max/extent must be bit-identical to torch's, also when the extent sits on an integer
    boundary; with OUR meshes (network rounding noise) the shape can only differ from the reference's inside a
    window of the mesh noise around such a boundary - measured here by sliding view 2 across one."""
    from stabstitch2_b200 import pipeline
    g = golden_stream
    H, W = 720, 1280
    ref1, ref2 = T(g["smooth_mesh1"]), T(g["smooth_mesh2"])   # [1,N,7,9,2] reference meshes of the small stream
    # (1) identical inputs: bit-identical extents over a sweep of sub-ulp-scale shifts across an integer boundary
    _, _, wmin, hmin, ow, oh = O.canvas(ref1, ref2, H, W)
    target = float(torch.floor(ow)) + 1.0                      # next integer above the current width
    base_shift = (target - float(ow)) * 480.0 / W              # shift of view 2 (px @480) that lands on it
    flips = 0
    for k in range(-40, 41):
        dx = base_shift + k * 2.5e-5
        s2 = ref2 + torch.tensor([dx, 0.0])
        _, _, wmin_r, hmin_r, ow_r, oh_r = O.canvas(ref1, s2, H, W)
        mm = pipeline.canvas_minmax(ref1[0], s2[0], H, W).cpu()
        assert float(mm[0]) == float(wmin_r) and float(mm[2]) == float(hmin_r)
        assert float(mm[1] - mm[0]) == float(ow_r) and float(mm[3] - mm[2]) == float(oh_r)
        assert pipeline.canvas_size(mm.tolist()) == (int(oh_r.int()), int(ow_r.int()))
        flips += int(ow_r.int()) != int(ow.int())
    assert 0 < flips < 81                                      # the sweep really crossed the boundary
    # (2) our meshes = reference meshes + noise of the size test_stream_golden measures: shape mismatches are confined
    # to shifts within that noise (scaled to hr pixels) of the boundary
    noise = 1.0e-3
    gen = torch.Generator().manual_seed(4)
    ours1 = ref1 + noise * (2 * torch.rand(ref1.shape, generator=gen) - 1)
    ours2 = ref2 + noise * (2 * torch.rand(ref2.shape, generator=gen) - 1)
    bad = []
    for k in range(-200, 201):
        dx = base_shift + k * 2.0e-5
        sh = torch.tensor([dx, 0.0])
        _, _, _, _, ow_r, oh_r = O.canvas(ref1, ref2 + sh, H, W)
        mm = pipeline.canvas_minmax(ours1[0], (ours2 + sh)[0], H, W).cpu().tolist()
        if pipeline.canvas_size(mm) != (int(oh_r.int()), int(ow_r.int())):
            bad.append(k * 2.0e-5)
    width = (max(bad) - min(bad)) if bad else 0.0
    print("canvas shape mismatches for %d of 401 shifts, window %.2e px @480 (mesh noise +-%.0e px)" % (len(bad), width, noise))
    assert width <= 2 * 2 * noise + 1e-4


# ---------------------------------------------------------------- the unchanged driver's call sequence
def test_dropin_replay_matches_batched_path(stream_inputs, golden_stream):
    """tests/dropin_replay.py: the reference driver's per-frame loops (batch-1 calls, per-frame transformer, torch
    blend; test_online_tra.py:284-399) through the flat drop-in names, against the batched whole-stream call and the
    reference's own golden outputs."""
    from stabstitch2_b200 import pipeline
    from tests import dropin_replay as R
    from tests.golden.make_golden import MESH_SCALE_S, MESH_SCALE_T
    n = R.import_flat()
    s, t, m = n["SpatialNet"]().cuda().eval(), n["TemporalNet"]().cuda().eval(), n["SmoothNet"]().cuda().eval()
    s.load_state_dict(Wt.spatial_state_dict(mesh_scale=MESH_SCALE_S), strict=True)
    t.load_state_dict(Wt.temporal_state_dict(mesh_scale=MESH_SCALE_T), strict=True)
    m.load_state_dict(Wt.smooth_state_dict(), strict=True)
    hr, lr = stream_inputs
    frames, S1, S2 = R.replay(n, s, t, m, lr[0], lr[1], hr[0], hr[1])
    g = golden_stream
    assert maxdiff(S1, g["smooth_mesh1"]) < MESH_TOL_PX and maxdiff(S2, g["smooth_mesh2"]) < MESH_TOL_PX
    assert tuple(frames[0].shape[:2]) == tuple(g["canvas_hw"]) and len(frames) == len(hr[0])
    d = np.abs(frames[0] - g["frame0"])
    assert (d > 0.05).mean() < 2e-3 and np.median(d) < 1e-3
    fused, b1, b2 = pipeline.stitch_stream(s, t, m, torch.cat(lr[0], 0).cuda(), torch.cat(lr[1], 0).cuda(),
                                           torch.cat(hr[0], 0).cuda(), torch.cat(hr[1], 0).cuda())
    # batch-1 calls and the batched stream run the same per-image kernels
    assert maxdiff(S1[0], b1) < 1e-4 and maxdiff(S2[0], b2) < 1e-4
    db = np.abs(frames[3] - fused[3].permute(1, 2, 0).cpu().numpy())
    assert (db > 0.05).mean() < 1e-3 and np.median(db) < 1e-3


# ---------------------------------------------------------------- LINEAR fusion
def test_linear_blender_clean_masks_vs_reference(golden_linear):
    """the driver's linear_blender (test_online_tra.py:34-58) on exact 0/1 masks, where the reference's nonzero-based
    centres and the residue-free semantics of the CUDA path coincide: against the reference's own output."""
    from stabstitch2_b200.pipeline import linear_blender
    from tests.golden.make_golden import linear_inputs
    ref, tgt, m1, m2 = linear_inputs()
    out = linear_blender(ref, tgt, m1, m2)
    mk = linear_blender(ref, tgt, m1, m2, mask=True)
    dm, do = maxdiff(mk, golden_linear["clean_mask1"]), maxdiff(out, golden_linear["clean_out"])
    print("linear_blender vs reference (clean masks): mask1 max |diff| %.2e, stitched max |diff| %.2e" % (dm, do))
    assert dm < 2e-5          # separable 2 x 21-tap blur vs torchvision's 441-tap conv2d: summation order only
    assert do < 255 * 2e-5
    # batch of two identical problems == single
    out2 = linear_blender(torch.cat([ref, ref]), torch.cat([tgt, tgt]), torch.cat([m1, m1]), torch.cat([m2, m2]))
    assert maxdiff(out2[0:1], out) == 0.0 and maxdiff(out2[1:2], out) == 0.0


def test_linear_stream_vs_reference_golden(golden_stream, golden_linear, stream_inputs):
    """get_stable_sqe(..., fusion_mode='LINEAR') on the small stream's reference meshes.  Against the oracle with the
    residue-free mask semantics (what the CUDA path implements): tight.  Against the reference's own LINEAR frame,
    whose centres include the rounding residues of out-of-image samples (tests/golden/linear.npz::stream_centroids:
    the centre columns move by 24 px): the measured distance is printed and bounded."""
    from stabstitch2_b200.pipeline import get_stable_sqe
    hr, _ = stream_inputs
    g, gl = golden_stream, golden_linear
    S1, S2 = T(g["smooth_mesh1"])[:, :2], T(g["smooth_mesh2"])[:, :2]
    frames, ow, oh = get_stable_sqe(hr[0][:2], hr[1][:2], S1, S2, "NORMAL", "LINEAR")
    assert (int(oh), int(ow)) == tuple(gl["stream_canvas_hw"])
    H, W = hr[0][0].shape[2:]
    M1, M2, wmin, hmin, oww, ohh = O.canvas(S1, S2, H, W)
    clean = O.stable_frame_linear(hr[0][0], hr[1][0], M1[:, 0], M2[:, 0], wmin, hmin, oww, ohh, clean=True)
    d_clean = np.abs(frames[0] - clean.numpy().transpose(1, 2, 0))
    d_ref = np.abs(frames[0] - gl["stream_frame0"])
    bound = grad_max(hr[0][0]) * COORD_TOL_PX + 1e-3
    print("LINEAR frame 0: vs oracle(clean masks) frac > %.3f: %.2e, median %.2e, max %.2e | vs reference frame: max %.3f, "
          "p99 %.3f, median %.2e" % (bound, (d_clean > bound).mean(), np.median(d_clean), d_clean.max(), d_ref.max(),
                                     np.percentile(d_ref, 99), np.median(d_ref)))
    # hard-edge flips of single samples (mask and image together) as a fraction, like the AVERAGE tests
    assert (d_clean > bound).mean() < 3e-3 and np.median(d_clean) < 1e-3
    # oracle(clean) vs reference measured: max 0.61, p99 0.26 grey levels (make_golden linear case)
    assert np.percentile(d_ref, 99) < 0.6 and np.median(d_ref) < 1e-2


def test_linear_fullsize_720p_vs_oracle():
    """LINEAR fusion at 720p with the lattice field (C = 4 instantiation carries the mask) against the oracle"""
    from stabstitch2_b200 import pipeline
    H, W = 720, 1280
    m1, m2 = _canvas_case(H, W)
    s1, s2 = _smooth_frame(1, H, W), _smooth_frame(2, H, W)
    M1, M2, wmin, hmin, ow, oh = O.canvas(m1[None], m2[None], H, W)
    ref = O.stable_frame_linear(s1, s2, M1[:, 0], M2[:, 0], wmin, hmin, ow, oh, clean=True)
    mm = pipeline.canvas_minmax(m1, m2, H, W).cpu().tolist()
    got = pipeline.stable_frames(s1.cuda(), s2.cuda(), m1, m2, mm, fusion_mode="LINEAR")[0].cpu()
    assert tuple(got.shape) == tuple(ref.shape)
    d = (got - ref).abs().numpy()
    print("LINEAR 720p vs oracle(clean): max %.3e, p99.9 %.3e, median %.2e, frac > 0.05: %.2e"
          % (d.max(), np.percentile(d, 99.9), np.median(d), (d > 0.05).mean()))
    # a sample that flips across the hard image edge changes mask and pixel together: counted as a fraction
    assert (d > 0.05).mean() < 5e-4 and np.median(d) < 1e-3


def test_three_view_linear_vs_oracle(golden_threeview):
    """three-view LINEAR fusion (test_online_tra_threeview.py:492-503) against the oracle with residue-free masks;
    the distance to the reference's own frames (two chained blends on a 96x128 canvas where 12 % of the 'mask'
    pixels are rounding residues) is printed."""
    from stabstitch2_b200 import pipeline
    from tests.golden.make_golden import threeview_inputs
    w12m1, w12m2, w23m1, w23m2, imgs = threeview_inputs()
    frames = pipeline.three_view_stable(imgs[0], imgs[1], imgs[2], w12m1, w12m2, w23m1, w23m2, fusion_mode="LINEAR")
    with torch.no_grad():
        rm1, rmid, rm3, wmin, hmin, ow, oh = O.three_view_meshes(w12m1, w12m2, w23m1, w23m2, 96, 128)
    for k in range(3):
        with torch.no_grad():
            ref = O.three_view_frame_linear(imgs[0][k], imgs[1][k], imgs[2][k], rm1[:, k], rmid[:, k], rm3[:, k], wmin, hmin,
                                            ow, oh, clean=True)
        d = np.abs(frames[k].numpy() - ref.numpy())
        dr = np.abs(frames[k].numpy() - golden_threeview["frames_linear"][k])
        print("three-view LINEAR frame %d: vs oracle(clean) frac > 0.05 %.2e median %.2e | vs reference max %.2f p99 %.2f"
              % (k, (d > 0.05).mean(), np.median(d), dr.max(), np.percentile(dr, 99)))
        assert tuple(frames[k].shape) == tuple(ref.shape)
        assert (d > 0.05).mean() < 1e-2 and np.median(d) < 1e-3
        assert np.percentile(dr, 99) < 6.0   # oracle(clean) vs reference measured: p99 1.0 / 3.6 / 3.9 grey levels


# ---------------------------------------------------------------- uint8 host edges
def _synth_u8(seed, n, H, W):
    """decoded-video-like uint8 BGR frames: band-limited texture + per-pixel noise, all 256 levels present"""
    g = torch.Generator().manual_seed(seed)
    base = torch.cat([O.synth_frame(k, seed % 2, H, W) for k in range(n)], 0)          # [n,3,H,W] 0..255
    noisy = base + 6.0 * torch.randn(base.shape, generator=g)
    return noisy.clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()           # [n,H,W,3]


@pytest.mark.parametrize("H,W", [(720, 1280), (1080, 1920), (96, 130)])
def test_host_edges_u8_bit_exact(H, W):
    """device front end (astype + transpose, cv2.resize INTER_LINEAR, /127.5-1) and back end (astype(uint8)) against
    oracle/host_edges.py, which is pinned bit-exactly to cv2: BIT-EXACT (integer / byte work)."""
    from oracle import host_edges as HE
    from stabstitch2_b200 import pipeline
    u = _synth_u8(3, 2, H, W)
    hr, lr = pipeline.load_frames_u8(u)
    for k in range(u.shape[0]):
        rhr, rlr = HE.load_frame(u[k].numpy())
        assert np.array_equal(hr[k].cpu().numpy(), rhr[0])
        assert np.array_equal(lr[k].cpu().numpy(), rlr[0]), np.abs(lr[k].cpu().numpy() - rlr[0]).max()
    g = torch.Generator().manual_seed(9)
    fused = torch.rand(2, 3, 57, 131, generator=g) * 262.0 - 3.0      # includes slightly negative and > 255 values
    got = pipeline.frames_to_u8(fused.cuda()).cpu().numpy()
    for k in range(2):
        assert np.array_equal(got[k], HE.to_video_frame(fused[k].permute(1, 2, 0).numpy()))


def test_stream_host_u8_matches_fp32_interface(nets):
    """uint8 e2e call == (host edges of the oracle) -> fp32 e2e call -> astype(uint8): bit-identical frames and
    meshes, also with two chunks in flight (prefetch)."""
    from oracle import host_edges as HE
    from stabstitch2_b200 import pipeline
    s, t, m = nets
    n, H, W = 8, 180, 320
    u1, u2 = _synth_u8(0, n, H, W), _synth_u8(1, n, H, W)
    hr1, lr1 = zip(*[HE.load_frame(u1[k].numpy()) for k in range(n)])
    hr2, lr2 = zip(*[HE.load_frame(u2[k].numpy()) for k in range(n)])
    ins = [torch.from_numpy(np.concatenate(x, 0)).contiguous() for x in (lr1, lr2, hr1, hr2)]
    fused, s1, s2 = pipeline.stitch_stream(s, t, m, *[x.cuda() for x in ins])
    ref_u8 = np.stack([HE.to_video_frame(fused[k].permute(1, 2, 0).cpu().numpy()) for k in range(n)], 0)
    out = torch.empty(fused.numel(), dtype=torch.uint8).pin_memory()
    ho, wo, m1, m2 = pipeline.stitch_stream_host_u8(s, t, m, u1.pin_memory(), u2.pin_memory(), out, want_meshes=True)
    assert (ho, wo) == tuple(fused.shape[2:])
    assert maxdiff(m1, s1) == 0.0 and maxdiff(m2, s2) == 0.0
    assert np.array_equal(out.numpy().reshape(n, ho, wo, 3), ref_u8)
    # pipelined: prefetch + async on two slots
    p1, p2 = u1.pin_memory(), u2.pin_memory()
    outs = [torch.empty(fused.numel(), dtype=torch.uint8).pin_memory() for _ in range(2)]
    for i in range(3):
        if i + 1 < 3:
            pipeline.stitch_stream_host_u8_prefetch((i + 1) & 1, p1, p2)
        pipeline.stitch_stream_host_u8_async(s, t, m, i & 1, p1, p2, outs[i & 1])
    for slot in (0, 1):
        pipeline.stitch_stream_host_wait(slot)
        assert np.array_equal(outs[slot].numpy().reshape(n, ho, wo, 3), ref_u8)
    # two-phase calls: chunk i+1 submitted before chunk i is finished (what bench.py's e2e leg does)
    for o in outs:
        o.zero_()
    pipeline.stitch_stream_host_submit(s, t, m, 0, p1, p2)
    for i in range(4):
        if i + 1 < 4:
            pipeline.stitch_stream_host_submit(s, t, m, (i + 1) & 1, p1, p2)
        assert pipeline.stitch_stream_host_finish(i & 1, outs[i & 1]) == (ho, wo)
        if i >= 1:   # the other slot's chunk (i-1) must be complete before its buffer is checked / reused
            pipeline.stitch_stream_host_wait((i - 1) & 1)
            assert np.array_equal(outs[(i - 1) & 1].numpy().reshape(n, ho, wo, 3), ref_u8)
    pipeline.stitch_stream_host_wait(1)
    assert np.array_equal(outs[1].numpy().reshape(n, ho, wo, 3), ref_u8)
    # misuse: finishing a slot with nothing submitted, submitting twice
    from stabstitch2_b200 import _lib
    with pytest.raises(_lib.SS2Error):
        pipeline.stitch_stream_host_finish(0, outs[0])
    pipeline.stitch_stream_host_submit(s, t, m, 0, p1, p2)
    with pytest.raises(_lib.SS2Error):
        pipeline.stitch_stream_host_submit(s, t, m, 0, p1, p2)
    pipeline.stitch_stream_host_finish(0, outs[0])
    pipeline.stitch_stream_host_wait(0)


# ---------------------------------------------------------------- convolution kernels (SIMT fp32 and tcgen05)
CONV_CASES = [
    # B, D, H, W, Cin, Cout, k, stride, pad, kd, pad_d
    (2, 1, 45, 60, 128, 128, 3, 1, 1, 1, 0),    # ResNet layer2 body
    (2, 1, 90, 120, 64, 64, 3, 1, 1, 1, 0),     # ResNet layer1 body (one image row per direct-conv tile)
    (3, 1, 11, 15, 128, 128, 3, 1, 1, 1, 0),    # regressor stage: ragged last tile (11 rows, 7 per tile)
    (3, 1, 90, 120, 64, 128, 3, 2, 1, 1, 0),    # layer2 entry, stride 2
    (3, 1, 90, 120, 64, 128, 1, 2, 0, 1, 0),    # 1x1 stride-2 shortcut
    (2, 1, 23, 30, 256, 256, 3, 1, 1, 1, 0),    # layer3 body
    (5, 1, 5, 7, 128, 256, 3, 1, 1, 1, 0),      # regressor tail: several images per M tile
    (4, 1, 2, 3, 256, 256, 3, 1, 1, 1, 0),
    (2, 1, 45, 60, 128, 64, 3, 1, 1, 1, 0),     # cost volume (121 -> 128 padded) into the regressor
    (3, 7, 7, 9, 128, 128, 3, 1, 1, 5, 2),      # SmoothNet Conv3d (5,3,3)
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("use_tc", [False, True])
def test_conv_kernels_vs_torch(case, use_tc):
    """plain PyTorch fp32 reference of the same op (CPU, fp64 accumulate) vs our kernels"""
    from stabstitch2_b200 import _lib
    B, D, H, W, Cin, Cout, k, stride, pad, kd, pad_d = case
    g = torch.Generator().manual_seed(B * 1000 + H)
    three_d = kd > 1
    x = torch.randn(B, D, H, W, Cin, generator=g)
    w = torch.randn(*( (Cout, Cin, kd, k, k) if three_d else (Cout, Cin, k, k) ), generator=g) / (Cin * k * k * kd) ** 0.5
    b = torch.randn(Cout, generator=g)
    xin = x if three_d else x[:, 0]
    if three_d:
        ref = torch.nn.functional.conv3d(x.permute(0, 4, 1, 2, 3).double(), w.double(), b.double(), 1, (pad_d, pad, pad))
        ref = ref.permute(0, 2, 3, 4, 1)
    else:
        ref = torch.nn.functional.conv2d(xin.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride, pad).permute(0, 2, 3, 1)
    res = torch.randn(*ref.shape, generator=g)
    ref = torch.relu(ref + res.double()).float()
    out = _lib.conv_nhwc(xin.cuda(), w, b, stride=stride, pad=pad, pad_d=pad_d, relu=True, residual=res.cuda(), use_tc=use_tc)
    assert out.shape == ref.shape
    err = (out.cpu() - ref).abs().max().item()
    # SIMT path: fp32 FMA chain.  Tensor-core path: split-TF32 operands (2^-21 relative) but the
    # TMEM accumulator truncates on every one of the K/8 accumulation steps, ~K/8 * 2^-24 * |sum|
    assert err < (3e-4 if use_tc else 2e-5), err


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("mode", [1, 3])
def test_conv_kernels_f16_planes_vs_torch(case, mode, monkeypatch):
    """every production layer shape with fp16 split planes in (mode 1) and in + out (mode 3): the direct 3x3 kernel and the
    implicit-GEMM kernel (stride-2 entries, 1x1 shortcuts, Conv3d, small maps) run kind::f16 MMAs; same bound as the
    split-TF32 path in test_conv_kernels_vs_torch"""
    from stabstitch2_b200 import _lib
    B, D, H, W, Cin, Cout, k, stride, pad, kd, pad_d = case
    g = torch.Generator().manual_seed(B * 1000 + H)
    three_d = kd > 1
    x = torch.randn(B, D, H, W, Cin, generator=g)
    w = torch.randn(*((Cout, Cin, kd, k, k) if three_d else (Cout, Cin, k, k)), generator=g) / (Cin * k * k * kd) ** 0.5
    b = torch.randn(Cout, generator=g)
    xin = x if three_d else x[:, 0]
    if three_d:
        ref = torch.nn.functional.conv3d(x.permute(0, 4, 1, 2, 3).double(), w.double(), b.double(), 1, (pad_d, pad, pad))
        ref = ref.permute(0, 2, 3, 4, 1)
    else:
        ref = torch.nn.functional.conv2d(xin.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride, pad).permute(0, 2, 3, 1)
    res = torch.randn(*ref.shape, generator=g)
    ref = torch.relu(ref + res.double()).float()
    monkeypatch.setenv("SS2_CONV_TEST_F16", str(mode))
    out = _lib.conv_nhwc(xin.cuda(), w, b, stride=stride, pad=pad, pad_d=pad_d, relu=True, residual=res.cuda(), use_tc=True)
    assert out.shape == ref.shape
    err = (out.cpu() - ref).abs().max().item()
    assert err < 3e-4, err


@pytest.mark.parametrize("shape", [(3, 13, 37, 64, 64), (2, 29, 126, 64, 128), (5, 9, 50, 128, 64)])
def test_conv_dc_every_tile_plan(shape, monkeypatch):
    """the direct 3x3 kernel under every tile plan the planner can pick (1-4 column tiles x 1-2 row blocks, forced with
    SS2_DC_PLAN) on ragged shapes (W not a multiple of the column tile, rows not a multiple of the tile), with residual
    and the split output planes written like inside the networks: every plan against the PyTorch fp64 reference"""
    from stabstitch2_b200 import _lib
    B, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(B * 100 + W)
    x = torch.randn(B, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g)
    res = torch.randn(B, H, W, Cout, generator=g)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), 1, 1).permute(0, 2, 3, 1)
    ref = torch.relu(ref + res.double()).float()
    monkeypatch.setenv("SS2_CONV_TEST_SPLIT", "1")
    ran = 0
    for plan in ["", "1,1", "1,2", "2,1", "2,2", "3,1", "3,2", "4,1", "4,2"]:
        if plan:
            monkeypatch.setenv("SS2_DC_PLAN", plan)
        else:
            monkeypatch.delenv("SS2_DC_PLAN", raising=False)
        try:
            out = _lib.conv_nhwc(x.cuda(), w, b, stride=1, pad=1, relu=True, residual=res.cuda(), use_tc=True).cpu()
        except Exception as exc:       # a plan that does not fit shared memory / TMEM for this layer is refused, not mis-run
            assert plan and "conv_dc" in str(exc), (plan, exc)
            continue
        ran += 1
        err = (out - ref).abs().max().item()
        assert err < 3e-4, (plan, err)
    assert ran >= 5


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("shape", [(2, 11, 37, 64, 64), (1, 23, 30, 256, 256), (3, 9, 60, 128, 128), (2, 90, 120, 64, 64)])
def test_conv_dc_f16_planes(shape, mode, monkeypatch):
    """the fp16-plane path of the direct 3x3 kernel (h16 = fp16(v), l16 = fp16((v - h16) * 2048); kind::f16 MMAs with the
    correction products in a 2048-scaled accumulator) against the PyTorch fp64 reference, at the accuracy asked of the
    split-TF32 path: mode 1 = fp16 planes in, 2 = fp16 planes out (merged back), 3 = both; with residual, bias, ReLU.
    Inputs span six orders of magnitude so that values below fp16's normal range are covered by the scaled low plane."""
    from stabstitch2_b200 import _lib
    B, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(B * 1000 + W + mode)
    x = torch.randn(B, H, W, Cin, generator=g)
    x = x * torch.pow(10.0, torch.randint(-5, 2, (B, H, W, 1), generator=g).float())   # per-pixel scales 1e-5 .. 10
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g)
    res = torch.randn(B, H, W, Cout, generator=g)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), 1, 1).permute(0, 2, 3, 1)
    ref = torch.relu(ref + res.double()).float()
    monkeypatch.delenv("SS2_DC_PLAN", raising=False)
    base = _lib.conv_nhwc(x.cuda(), w, b, stride=1, pad=1, relu=True, residual=res.cuda(), use_tc=True).cpu()
    monkeypatch.setenv("SS2_CONV_TEST_F16", str(mode))
    out = _lib.conv_nhwc(x.cuda(), w, b, stride=1, pad=1, relu=True, residual=res.cuda(), use_tc=True).cpu()
    scale = ref.abs().max().item()
    err, err_base = (out - ref).abs().max().item(), (base - ref).abs().max().item()
    assert err < 1e-4 * max(scale, 1.0), (err, err_base, scale)
    assert err < 4 * err_base + 1e-6 * max(scale, 1.0), (err, err_base)    # no worse than the split-TF32 kernel's own error


def test_f16_planes_range_overflow_fails_loudly(monkeypatch):
    """an activation beyond fp16's range (|v| > 65504) written to the fp16 split planes raises the context's range flag: the
    next entry point reports it (once) instead of handing out clamped values silently"""
    from stabstitch2_b200 import _lib
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 8, 16, 64, generator=g)
    w = torch.randn(64, 64, 3, 3, generator=g) / 24.0
    ok = _lib.conv_nhwc(x.cuda(), w, None, stride=1, pad=1, use_tc=True).cpu()
    monkeypatch.setenv("SS2_CONV_TEST_F16", "2")
    _lib.conv_nhwc((x * 1e6).cuda(), w, None, stride=1, pad=1, use_tc=True)     # outputs ~1e6: do not fit
    monkeypatch.delenv("SS2_CONV_TEST_F16")
    with pytest.raises(_lib.SS2Error, match="fp16 range"):
        _lib.conv_nhwc(x.cuda(), w, None, stride=1, pad=1, use_tc=True)
    again = _lib.conv_nhwc(x.cuda(), w, None, stride=1, pad=1, use_tc=True).cpu()   # reported once, the context works on
    assert torch.equal(ok, again)


@pytest.mark.parametrize("B,H", [(1, 8), (2, 44), (3, 360), (33, 360)])
def test_stem_pool_direct_vs_torch(B, H):
    """the fused direct stem kernel (conv 7x7 s2 + bias + ReLU + max-pool 3x3 s2 in one tcgen05 kernel) against a plain
    PyTorch reference of the same ops (CPU, fp64 accumulate) and against the two-kernel paths; B = 33 is the
    TemporalNet halo chunk (band scheduling with a ragged unit count), H = 8 a single band"""
    from stabstitch2_b200 import _lib
    g = torch.Generator().manual_seed(77 + B + H)
    x = torch.rand(B, 3, H, 480, generator=g) * 2 - 1
    w = torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5
    b = torch.randn(64, generator=g) * 0.2
    out = _lib.stem_pool(x.cuda(), w, b, variant=2).cpu()
    if B <= 3:
        ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), 2, 3)
        ref = torch.nn.functional.max_pool2d(torch.relu(ref), 3, 2, 1).permute(0, 2, 3, 1).float()
        assert out.shape == ref.shape
        err = (out - ref).abs().max().item()
        assert err < 1e-4, err           # split TF32: 2^-21 relative per product, K = 147
        simt = _lib.stem_pool(x.cuda(), w, b, variant=0).cpu()
        assert (simt - ref).abs().max().item() < 2e-5
    two = _lib.stem_pool(x.cuda(), w, b, variant=1).cpu()   # same split-TF32 products, other summation order
    assert out.shape == two.shape
    assert (out - two).abs().max().item() < 1e-4


@pytest.mark.parametrize("sr", [5, 3])
def test_cost_volume_c128_tiled_vs_oracle(sr):
    """production shape (C = 128, 45x60 map): the shared-memory tiled kernel against the oracle"""
    from stabstitch2_b200.spatial_network import SpatialNet
    g = torch.Generator().manual_seed(40 + sr)
    a, b = torch.randn(2, 128, 45, 60, generator=g), torch.randn(2, 128, 45, 60, generator=g)
    got = SpatialNet.cost_volume(a.cuda(), b.cuda(), sr, norm=False)
    assert maxdiff(got, O.cost_volume(a, b, sr)) < 2e-6


def test_ccl_c256_tensor_core_vs_oracle():
    """production shape (C = 256, 23x30 map): correlation GEMM on the tensor cores"""
    from stabstitch2_b200.spatial_network import SpatialNet
    g = torch.Generator().manual_seed(77)
    f1 = torch.randn(2, 256, 23, 30, generator=g)
    f2 = torch.roll(f1, (1, -2), (2, 3)) + 0.3 * torch.randn(2, 256, 23, 30, generator=g)
    got = SpatialNet().CCL(f1.cuda(), f2.cuda())
    assert maxdiff(got, O.ccl(f1, f2)) < 2e-4


# ---------------------------------------------------------------- three views (test_online_tra_threeview.py:345-505)
def test_three_view_vs_reference_golden(golden_threeview):
    """middle-plane meshes, canvas and fused frames against the reference's own three-view glue"""
    from stabstitch2_b200 import pipeline
    from tests.golden.make_golden import threeview_inputs
    g = golden_threeview
    w12m1, w12m2, w23m1, w23m2, imgs = threeview_inputs()
    m1, mid, m3, canvas = pipeline.three_view_meshes(w12m1[0], w12m2[0], w23m1[0], w23m2[0], 96, 128)
    for got, key in ((m1, "mesh1"), (mid, "middle"), (m3, "mesh3")):
        assert np.abs(got.cpu().numpy() - g[key][0]).max() < 2e-3, key
    assert np.abs(canvas.cpu().numpy() - g["canvas"]).max() < 2e-3
    frames = pipeline.three_view_stable(imgs[0], imgs[1], imgs[2], w12m1, w12m2, w23m1, w23m2)
    assert len(frames) == 3 and tuple(frames[0].shape) == tuple(g["frames"][0].shape)
    for k in range(3):
        d = np.abs(frames[k].numpy() - g["frames"][k])
        # 96x128 frames: the three hard image borders (where a 1e-4 px coordinate difference flips a sample in or
        # out of a view) are ~6% of this small canvas, so the flip fraction is higher than on full-size frames
        assert (d > 0.05).mean() < 4e-3, (k, (d > 0.05).mean())
        assert np.median(d) < 1e-3 and np.percentile(d, 99) < 2e-2


def test_three_view_720p_vs_oracle():
    """full-size three-view frame (lattice field) against the CPU oracle's restatement"""
    from stabstitch2_b200 import pipeline
    H, W = 720, 1280
    g = torch.Generator().manual_seed(5)
    rig = O.rigid_mesh(1, 360, 480)[:, None]

    def mesh(dx, dy):
        return rig + torch.tensor([dx, dy]) + 2.5 * torch.randn(1, 1, 7, 9, 2, generator=g)
    w12m1, w12m2, w23m1, w23m2 = mesh(-80.0, 3.0), mesh(85.0, -2.0), mesh(-70.0, 6.0), mesh(95.0, 1.0)
    # mid-grey smooth frames: the reference's blend a*(a/(a+b+1e-6)) amplifies the rounding residue of an uncovered
    # view wherever the covered view is dark, which no re-implementation reproduces
    imgs = [_smooth_frame(10 + v, H, W) for v in range(3)]
    with torch.no_grad():
        rm1, rmid, rm3, wmin, hmin, ow, oh = O.three_view_meshes(w12m1, w12m2, w23m1, w23m2, H, W)
        ref = O.three_view_frame(imgs[0], imgs[1], imgs[2], rm1[:, 0], rmid[:, 0], rm3[:, 0], wmin, hmin, ow, oh)
    m1, mid, m3, canvas = pipeline.three_view_meshes(w12m1[0], w12m2[0], w23m1[0], w23m2[0], H, W)
    assert (m1.cpu() - rm1[0]).abs().max() < 5e-3 and (m3.cpu() - rm3[0]).abs().max() < 5e-3
    assert (mid.cpu() - rmid[0]).abs().max() < 1e-3
    cv = canvas.cpu().tolist()
    assert abs(cv[2] - float(ow)) < 5e-3 and abs(cv[3] - float(oh)) < 5e-3
    # use the oracle's canvas so that both sides sample the same grid even if the extents differ in the last bit
    fused = pipeline.three_view_frames(imgs[0].cuda(), imgs[1].cuda(), imgs[2].cuda(), rm1[0].cuda(), rmid[0].cuda(), rm3[0].cuda(),
                                       [float(wmin), float(hmin), float(ow), float(oh)])[0].cpu()
    assert tuple(fused.shape) == tuple(ref.shape)
    # The reference's AVERAGE fusion a*(a/(a+b+1e-6)) is ill-conditioned wherever BOTH of its inputs are rounding
    # residues of uncovered views: in the part of the canvas that only view 3 covers, stage one fuses two residues
    # (|r| <~ 3e-2) and returns up to +-20 grey levels at ~1% of the pixels (measured once with a throwaway probe), which then leak
    # into the output; where nothing covers, values reach 5e5.  No re-implementation reproduces those bits, so:
    #   region A (view 1 or view 2 covers): compare with the reference restatement, residue-sized bounds;
    #   region B (only view 3 covers): our frame must equal view 3 warped alone (c*c/(c+1e-6));
    #   elsewhere: ~0 from us.
    Ho, Wo = ref.shape[1:]
    nrig = O.norm_mesh(O.rigid_mesh(1, H, W), H, W)
    cov, srcs, far = [], [], []
    for M in (rm1, rmid, rm3):
        tt = torch.stack([M[0, 0, ..., 0] - wmin, M[0, 0, ..., 1] - hmin], 2)[None]
        srcs.append(O.norm_mesh(tt, oh, ow))
        ax, ay = O.tps_source_coords_fp64(srcs[-1], nrig, Ho, Wo, W, H)
        cov.append((ax[0] > 1) & (ax[0] < W - 2) & (ay[0] > 1) & (ay[0] < H - 2))
        far.append((ax[0] < -1) | (ax[0] > W) | (ay[0] < -1) | (ay[0] > H))  # certainly outside the image
    region_a = cov[0] | cov[1]
    region_b = cov[2] & far[0] & far[1]
    assert region_a.mean() > 0.4 and region_b.mean() > 0.05
    d = (fused - ref).abs().numpy()
    da = d[:, region_a]
    assert (da > 0.1).mean() < 1e-3, ((da > 0.1).mean(), da.max())
    assert np.median(da) < 2e-3 and np.percentile(da, 99) < 6e-2, (np.median(da), np.percentile(da, 99))
    with torch.no_grad():
        c = O.tps_warp(imgs[2], srcs[2], nrig, (Ho, Wo))[0]
    db = (fused - c * (c / (c + 1e-6))).abs().numpy()[:, region_b]
    assert (db > 0.1).mean() < 1e-3 and np.percentile(db, 99) < 6e-2, ((db > 0.1).mean(), np.percentile(db, 99))
    rest = np.abs(fused.numpy()[:, far[0] & far[1] & far[2]])
    assert rest.max() < 256.0 and np.median(rest) < 1e-3


# ---------------------------------------------------------------- N views (BASELINE.json config 5)
def test_nview_equals_three_view(golden_threeview):
    """the N-view chain is the three-view glue for N = 3: meshes and canvas bit-identical to ss2_three_view_meshes,
    frames (fused 3-view pass) against the three-warps + blend path and the reference's golden frames"""
    from stabstitch2_b200 import pipeline
    from tests.golden.make_golden import threeview_inputs
    g = golden_threeview
    w12m1, w12m2, w23m1, w23m2, imgs = threeview_inputs()
    m1, mid, m3, canvas = pipeline.three_view_meshes(w12m1[0], w12m2[0], w23m1[0], w23m2[0], 96, 128)
    shifted, mids, mm1 = pipeline.nview_align([(w12m1[0], w12m2[0]), (w23m1[0], w23m2[0])], 96, 128)
    meshes, mm2 = pipeline.nview_remap(shifted, mids, mm1.cpu().tolist())
    for got, ref in zip(meshes, (m1, mid, m3)):
        assert maxdiff(got, ref) == 0.0
    c, q = canvas.cpu().tolist(), mm2.cpu().tolist()
    assert (q[0], q[2]) == (c[0], c[1]) and np.float32(q[1]) - np.float32(q[0]) == np.float32(c[2]) \
        and np.float32(q[3]) - np.float32(q[2]) == np.float32(c[3])
    stacks = [torch.cat(imgs[v], 0).cuda() for v in range(3)]
    fused = pipeline.nview_frames(stacks, meshes, q)
    ref3 = pipeline.three_view_frames(stacks[0], stacks[1], stacks[2], m1, mid, m3, c)
    assert tuple(fused.shape) == tuple(ref3.shape)
    assert maxdiff(fused, ref3) < 1e-4          # one fused pass vs three warps + blend: the same per-view arithmetic
    for k in range(3):
        d = np.abs(fused[k].cpu().numpy() - g["frames"][k])
        assert (d > 0.05).mean() < 4e-3 and np.median(d) < 1e-3


def test_nview_4_720p_vs_oracle():
    """four views at 720p (three chained pairs): meshes and canvas against the oracle's N-view restatement, the fused
    4-view lattice pass against the oracle's frame where the result is well conditioned"""
    from stabstitch2_b200 import pipeline
    H, W = 720, 1280
    g = torch.Generator().manual_seed(8)
    rig = O.rigid_mesh(1, 360, 480)[:, None]

    def mesh(dx, dy):
        return rig + torch.tensor([dx, dy]) + 2.5 * torch.randn(1, 1, 7, 9, 2, generator=g)
    pairs = [(mesh(-80.0, 3.0), mesh(85.0, -2.0)), (mesh(-70.0, 6.0), mesh(95.0, 1.0)), (mesh(-75.0, -4.0), mesh(90.0, 2.0))]
    imgs = [_smooth_frame(20 + v, H, W) for v in range(4)]
    with torch.no_grad():
        rmeshes, wmin, hmin, ow, oh = O.nview_meshes(pairs, H, W)
    shifted, mids, mm1 = pipeline.nview_align([(a[0], b[0]) for a, b in pairs], H, W)
    meshes, mm2 = pipeline.nview_remap(shifted, mids, mm1.cpu().tolist())
    for v in range(4):
        assert (meshes[v].cpu() - rmeshes[v][0]).abs().max() < 5e-3, v
    q = mm2.cpu().tolist()
    assert abs(q[0] - float(wmin)) < 5e-3 and abs(q[1] - q[0] - float(ow)) < 5e-3 and abs(q[3] - q[2] - float(oh)) < 5e-3
    # the oracle's canvas and meshes for the frame comparison, so both sides sample the same grid
    ref_mm = [float(wmin), float(wmin + ow), float(hmin), float(hmin + oh)]
    fused = pipeline.nview_frames([t.cuda() for t in imgs], torch.stack([m[0] for m in rmeshes], 0).cuda(), ref_mm)[0].cpu()
    Ho, Wo = fused.shape[1:]
    assert (Ho, Wo) == (int(oh.int()), int(ow.int())) or abs(Ho - int(oh.int())) <= 1
    nrig = O.norm_mesh(O.rigid_mesh(1, H, W), H, W)
    warps, cover, far = [], [], []
    with torch.no_grad():
        for v in range(4):
            M = rmeshes[v]
            tt = torch.stack([M[0, 0, ..., 0] - wmin, M[0, 0, ..., 1] - hmin], 2)[None]
            src = O.norm_mesh(tt, oh, ow)
            ax, ay = O.tps_source_coords_fp64(src, nrig, Ho, Wo, W, H)
            cover.append((ax[0] > 1) & (ax[0] < W - 2) & (ay[0] > 1) & (ay[0] < H - 2))
            far.append((ax[0] < -1) | (ax[0] > W) | (ay[0] < -1) | (ay[0] > H))
            w = O.tps_warp(imgs[v], src, nrig, (Ho, Wo))[0]
            warps.append(w * torch.from_numpy(cover[-1]).float())      # exactly 0 outside, like the lattice resampler
        ref = warps[0]
        for v in range(1, 4):
            ref = O.average_blend(ref, warps[v])
    settled = np.ones((Ho, Wo), bool)
    for v in range(4):
        settled &= cover[v] | far[v]          # every view either clearly inside or clearly outside its image
    d = (fused - ref).abs().numpy()[:, settled]
    print("4-view 720p: canvas %dx%d, settled fraction %.2f, max |diff| %.2e, median %.2e" % (Ho, Wo, settled.mean(), d.max(), np.median(d)))
    assert settled.mean() > 0.9
    assert (d > 5e-3).mean() < 1e-4 and np.median(d) < 1e-3


# ---------------------------------------------------------------- metric path
def test_metric_scores_vs_reference_golden():
    """stability / distortion scores (test_metric_ssd.py:455-479) against the reference's own grid losses and the oracle"""
    from oracle import metric_oracle as MO
    from stabstitch2_b200 import metrics
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metric.npz")))
    mesh, path = T(g["mesh"]), T(g["path"])
    d = float(metrics.distortion_score(mesh[0]))
    assert abs(d - float(np.max(g["inter"] + g["intra"]))) < 2e-6 * max(1.0, d)
    s = float(metrics.stability_score(path[0]))
    assert abs(s - float(MO.stability_score(path))) < 1e-5 * max(1.0, abs(s))
    # per-window paths -> whole-stream paths (:417-436)
    gen = torch.Generator().manual_seed(2)
    wo = torch.cumsum(torch.randn(5, 7, 7, 9, 2, generator=gen), 1)
    ws = wo + 0.3 * torch.randn(5, 7, 7, 9, 2, generator=gen)
    ori, smo = metrics.stream_paths(wo, ws)
    ro, rs = MO.accumulate_paths(wo, ws)
    assert maxdiff(ori[None], ro) == 0.0 and maxdiff(smo[None], rs) == 0.0


def test_metric_warp_psnr_ssim(golden_stream, stream_inputs):
    """the metric script's per-view warp with mask planes (C = 6, :151-181) against the reference's own output, and
    PSNR / SSIM in the overlap against the oracle's restatement of skimage 0.15 (fp64)"""
    from oracle import metric_oracle as MO
    from stabstitch2_b200 import metrics
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metric.npz")))
    _, lr = stream_inputs
    S1, S2 = T(golden_stream["smooth_mesh1"])[:, :2], T(golden_stream["smooth_mesh2"])[:, :2]
    l1, l2 = metrics.get_stable_sqe(lr[0][:2], lr[1][:2], S1, S2)
    assert l1[0].shape == (360, 480, 6)
    d = np.abs(l1[0] - g["warp1_frame0"])
    bound = grad_max((lr[0][0] + 1) * 127.5) * COORD_TOL_PX + 1e-3
    assert (d > bound).mean() < 2e-3 and np.median(d) < 1e-3, ((d > bound).mean(), d.max())
    d2 = np.abs(l2[1][::4] - g["warp2_frame1_rows4"])
    assert (d2 > bound).mean() < 2e-3
    w1 = torch.from_numpy(np.stack(l1, 0)).permute(0, 3, 1, 2).contiguous()
    w2 = torch.from_numpy(np.stack(l2, 0)).permute(0, 3, 1, 2).contiguous()
    ps, ss = metrics.psnr_ssim(w1, w2)
    for k in range(2):
        rp, rs = MO.psnr_overlap(l1[k], l2[k]), MO.ssim_overlap(l1[k], l2[k])
        print("frame %d: psnr %.4f (oracle %.4f), ssim %.6f (oracle %.6f)" % (k, float(ps[k]), rp, float(ss[k]), rs))
        assert abs(float(ps[k]) - rp) < 1e-3 and abs(float(ss[k]) - rs) < 1e-5
