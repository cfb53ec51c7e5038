"""Resampler-only timing: ss2_stable_frames (canvas meshes -> TPS solves -> lattice nodes -> fused resample + blend)
on synthetic 720p / 1080p frames with fixed meshes shaped like the bench stream (view 2 ~35 % to the right).

    python profiles/warp_bench.py [--frames 32] [--height 720 --width 1280] [--iters 20]

Prints one JSON line: CUDA-event time of the whole call, the library's own bracket around solve + nodes + resample
kernels (ss2_profile_*), algorithmic GB/s of both.  Inputs (708 MB at 32 x 720p) are larger than L2.
SS2_LIB selects another build of libss2.so (profiles/build_variants.sh)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--tag", default=os.environ.get("SS2_LIB", "default"))
    a = ap.parse_args()
    from stabstitch2_b200 import _lib, pipeline, synthetic
    H, W, F = a.height, a.width, a.frames
    g = torch.Generator().manual_seed(11)
    ys = torch.linspace(0, 360, 7)[:, None].expand(7, 9)
    xs = torch.linspace(0, 480, 9)[None, :].expand(7, 9)
    rig = torch.stack([xs, ys], 2)[None]
    m1 = (rig + torch.tensor([-86.0, 0.0]) + 3.0 * torch.randn(F, 7, 9, 2, generator=g)).cuda()
    m2 = (rig + torch.tensor([86.0, 0.0]) + 3.0 * torch.randn(F, 7, 9, 2, generator=g)).cuda()
    base = [synthetic.synth_frame(k, v, H, W) for k in range(2) for v in range(2)]
    hr1 = torch.cat([base[2 * (k % 2)] for k in range(F)], 0).cuda()
    hr2 = torch.cat([base[2 * (k % 2) + 1] for k in range(F)], 0).cuda()
    mm = pipeline.canvas_minmax(m1, m2, H, W).cpu().tolist()
    Ho, Wo = pipeline.canvas_size(mm)
    out = torch.empty(F, 3, Ho, Wo, device="cuda")
    ctx = _lib.context()
    for _ in range(3):
        pipeline.stable_frames(hr1, hr2, m1, m2, mm, out=out)
    torch.cuda.synchronize()
    ctx.profile_enable(_lib.PROF_WARP, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        pipeline.stable_frames(hr1, hr2, m1, m2, mm, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    wms, wn, wbytes = ctx.profile_read(_lib.PROF_WARP)
    ctx.profile_enable(_lib.PROF_WARP, False)
    bytes_alg = F * (2 * 3 * H * W + 3 * Ho * Wo) * 4
    print(json.dumps({"tag": a.tag, "frames": F, "src": [H, W], "canvas": [Ho, Wo], "call_ms": ms,
                      "call_gbs": bytes_alg / ms / 1e6, "bracket_ms": wms / max(wn, 1),
                      "bracket_gbs": (wbytes / max(wn, 1)) / (wms / max(wn, 1)) / 1e6 if wn else None,
                      "checksum": float(out.double().sum().item()),
                      # the synthetic frames come from torch CPU kernels (bicubic upsampling, tanh), whose vectorised / scalar split
                      # depends on buffer alignment: a few input pixels can differ in the last bit from process to process
                      "input_checksum": float(hr1.double().sum().item() + hr2.double().sum().item()),
                      "mesh_checksum": float(m1.double().sum().item() + m2.double().sum().item())}))


if __name__ == "__main__":
    main()
