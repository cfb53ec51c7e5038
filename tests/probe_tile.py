"""GPU probe (not a pytest file): the TMA-staged tile resampler against the direct-load lattice
kernel (same field evaluation, SS2_TPS_TILE=0) and the exact kernel on hand-made meshes, plus
timings.  python tests/probe_tile.py [H W] [N]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stabstitch2_b200 import _lib, pipeline, synthetic  # noqa: E402


def meshes(n, seed, shift, amp, rot=0.0):
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.linspace(0, 360, 7), torch.linspace(0, 480, 9), indexing="ij")
    rigid = torch.stack([xs, ys], -1)[None].repeat(n, 1, 1, 1)
    m = rigid + amp * torch.randn(n, 7, 9, 2, generator=g)
    if rot:
        c, s = torch.cos(torch.tensor(rot)), torch.sin(torch.tensor(rot))
        x, y = m[..., 0] - 240, m[..., 1] - 180
        m = torch.stack([c * x - s * y + 240, s * x + c * y + 180], -1)
    m[..., 0] += shift
    return m


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    a = [x for x in sys.argv[1:] if x.isdigit()]
    H, W = (int(a[0]), int(a[1])) if len(a) >= 2 else (720, 1280)
    N = int(a[2]) if len(a) >= 3 else 8
    hr1 = torch.cat([synthetic.synth_frame(k, 0, H, W) for k in range(N)], 0).cuda()
    hr2 = torch.cat([synthetic.synth_frame(k, 1, H, W) for k in range(N)], 0).cuda()
    for name, kw in (("smooth", dict(shift=170.0, amp=1.0)), ("bench-like", dict(shift=170.0, amp=4.0)), ("rotated 3deg", dict(shift=120.0, amp=3.0, rot=0.052)),
                     ("strong", dict(shift=60.0, amp=12.0, rot=-0.1))):
        m1 = meshes(N, 1, 0.0, kw["amp"], kw.get("rot", 0.0) * 0.5).cuda()
        m2 = meshes(N, 2, kw["shift"], kw["amp"], kw.get("rot", 0.0)).cuda()
        mm = pipeline.canvas_minmax(m1, m2, H, W).cpu().tolist()
        Ho, Wo = pipeline.canvas_size(mm)
        os.environ["SS2_TPS_TILE"] = os.environ.get("TILE_MODE", "1")
        f_tile = pipeline.stable_frames(hr1, hr2, m1, m2, mm, tps=_lib.TPS_LATTICE)
        torch.cuda.synchronize()
        t_tile = timeit(lambda: pipeline.stable_frames(hr1, hr2, m1, m2, mm, tps=_lib.TPS_LATTICE, out=f_tile))
        os.environ["SS2_TPS_TILE"] = "0"
        f_lat = pipeline.stable_frames(hr1, hr2, m1, m2, mm, tps=_lib.TPS_LATTICE)
        t_lat = timeit(lambda: pipeline.stable_frames(hr1, hr2, m1, m2, mm, tps=_lib.TPS_LATTICE, out=f_lat))
        os.environ["SS2_TPS_TILE"] = os.environ.get("TILE_MODE", "1")
        d = (f_tile - f_lat).abs()
        print("%-13s canvas %dx%d  tile vs lattice: max %.3e  mean %.3e  frac>1e-3 %.2e | ms/frame tile %.4f lattice %.4f"
              % (name, Ho, Wo, float(d.max()), float(d.mean()), float((d > 1e-3).float().mean()), t_tile / N, t_lat / N))
        if "--exact" in sys.argv:
            f_ex = pipeline.stable_frames(hr1[:2], hr2[:2], m1[:2], m2[:2], mm, tps=_lib.TPS_EXACT)
            d = (f_tile[:2] - f_ex).abs()
            print("%-13s tile vs exact: max %.3e mean %.3e frac>0.05 %.2e" % (name, float(d.max()), float(d.mean()),
                                                                         float((d > 0.05).float().mean())))


if __name__ == "__main__":
    main()
