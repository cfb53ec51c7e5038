"""GPU probe (not a pytest file): accuracy of the two TPS field evaluations against the fp64
arbiter on bench-like 720p meshes, and kernel timings.  python tests/probe_warp.py [H W]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import stabstitch_oracle as O  # noqa: E402
from stabstitch2_b200 import _lib, pipeline, synthetic  # noqa: E402
from stabstitch2_b200.smooth_network import SmoothNet  # noqa: E402
from stabstitch2_b200.spatial_network import SpatialNet  # noqa: E402
from stabstitch2_b200.temporal_network import TemporalNet  # noqa: E402
from stabstitch2_b200.utils.torch_tps_transform import transformer  # noqa: E402


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 and sys.argv[1].isdigit() else (720, 1280)
    N = 16
    s, t, m = SpatialNet().cuda().eval(), TemporalNet().cuda().eval(), SmoothNet().cuda().eval()
    s.load_state_dict(synthetic.spatial_state_dict(mesh_scale=20.0), strict=True)
    t.load_state_dict(synthetic.temporal_state_dict(mesh_scale=10.0), strict=True)
    m.load_state_dict(synthetic.smooth_state_dict(), strict=True)
    hr1 = torch.cat([synthetic.synth_frame(k, 0, H, W) for k in range(N)], 0)
    hr2 = torch.cat([synthetic.synth_frame(k, 1, H, W) for k in range(N)], 0)
    lr1, lr2 = synthetic.lowres(hr1).cuda(), synthetic.lowres(hr2).cuda()
    hr1, hr2 = hr1.cuda(), hr2.cuda()
    S1, S2 = pipeline.stream_meshes(s, t, m, lr1, lr2)
    mm = pipeline.canvas_minmax(S1, S2, H, W).cpu().tolist()
    Ho, Wo = pipeline.canvas_size(mm)
    print("canvas", Ho, Wo)
    k = 3
    if "--time-only" in sys.argv:
        return timings(hr1, hr2, S1, S2, mm, N, Ho, Wo, only_lattice=True)
    # coordinate accuracy on frame 3, both views, against the fp64 arbiter
    M1, M2, wmin, hmin, ow, oh = O.canvas(S1.cpu()[None], S2.cpu()[None], H, W)
    nrig = O.norm_mesh(O.rigid_mesh(1, H, W), H, W)
    ramp = torch.stack([torch.arange(W, dtype=torch.float32)[None, :].expand(H, W),
                        torch.arange(H, dtype=torch.float32)[:, None].expand(H, W),
                        torch.zeros(H, W)], 0)[None]
    for v, M in enumerate((M1, M2)):
        tt = torch.stack([M[0, k, ..., 0] - wmin, M[0, k, ..., 1] - hmin], 2)[None]
        src = O.norm_mesh(tt, oh, ow)
        ax, ay = O.tps_source_coords_fp64(src, nrig, Ho, Wo, W, H)
        inside = (ax[0] > 1) & (ax[0] < W - 2) & (ay[0] > 1) & (ay[0] < H - 2)
        ref = O.tps_warp(ramp, src, nrig, (Ho, Wo)).numpy()[0]
        line = "view %d  ref: x max %.2e mean %.2e | y max %.2e mean %.2e" % (
            v, np.abs(ref[0] - ax[0])[inside].max(), np.abs(ref[0] - ax[0])[inside].mean(),
            np.abs(ref[1] - ay[0])[inside].max(), np.abs(ref[1] - ay[0])[inside].mean())
        print(line)
        for name, tps in (("exact", _lib.TPS_EXACT), ("lattice", _lib.TPS_LATTICE)):
            got = transformer(ramp.cuda(), src.cuda(), nrig.cuda(), (Ho, Wo), tps=tps).cpu().numpy()[0]
            ex, ey = np.abs(got[0] - ax[0])[inside], np.abs(got[1] - ay[0])[inside]
            print("view %d  %-7s: x max %.2e mean %.2e | y max %.2e mean %.2e" % (v, name, ex.max(), ex.mean(),
                                                                             ey.max(), ey.mean()))
    # fused frames: exact vs lattice vs oracle
    f_ex = pipeline.stable_frames(hr1[k:k + 1], hr2[k:k + 1], S1[k:k + 1], S2[k:k + 1], mm, tps=_lib.TPS_EXACT)[0].cpu()
    f_la = pipeline.stable_frames(hr1[k:k + 1], hr2[k:k + 1], S1[k:k + 1], S2[k:k + 1], mm, tps=_lib.TPS_LATTICE)[0].cpu()
    f_or, warps = O.stable_frame(hr1[k:k + 1].cpu(), hr2[k:k + 1].cpu(), M1[:, k], M2[:, k], wmin, hmin, ow, oh)
    for name, f in (("exact", f_ex), ("lattice", f_la)):
        d = (f - f_or).abs()
        print("fused %-7s vs oracle: max %.3f median %.2e mean %.2e frac>1e-3 %.3f frac>1e-2 %.4f frac>0.05 %.5f" % (
            name, d.max(), d.median(), d.mean(), (d > 1e-3).float().mean(), (d > 1e-2).float().mean(),
            (d > 0.05).float().mean()))
    d = (f_ex - f_la).abs()
    print("fused exact vs lattice: max %.3f mean %.2e frac>1e-2 %.4f" % (d.max(), d.mean(), (d > 1e-2).float().mean()))
    timings(hr1, hr2, S1, S2, mm, N, Ho, Wo)


def timings(hr1, hr2, S1, S2, mm, N, Ho, Wo, only_lattice=False):
    ctx = _lib.context()
    out = torch.empty(N, 3, Ho, Wo, device="cuda")
    for name, tps in (("exact", _lib.TPS_EXACT), ("lattice", _lib.TPS_LATTICE)):
        if only_lattice and name == "exact":
            continue
        for _ in range(2):
            pipeline.stable_frames(hr1, hr2, S1, S2, mm, tps=tps, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.profile_enable(_lib.PROF_WARP, True)
        e0.record()
        for _ in range(5):
            pipeline.stable_frames(hr1, hr2, S1, S2, mm, tps=tps, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms, nl, by = ctx.profile_read(_lib.PROF_WARP)
        ctx.profile_enable(_lib.PROF_WARP, False)
        print("%-7s: stable_frames %.3f ms / %d frames; warp kernel(s) %.1f us/frame -> %.0f GB/s algorithmic (%.1f%% of 6549)" % (
            name, e0.elapsed_time(e1) / 5, N, 1e3 * ms / nl / N, by / nl / (ms / nl * 1e-3) / 1e9,
            100 * by / nl / (ms / nl * 1e-3) / 1e9 / 6549.4))


if __name__ == "__main__":
    main()
