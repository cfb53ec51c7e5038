"""Per-layer timing of the direct tcgen05 convolution (conv_dc.cu) on the bench shapes (32 images), through
ss2_conv_nhwc with the residual and the split output planes a layer inside the networks has.
  python profiles/conv_bench.py dbg   [reps]  timing experiments SS2_DC_DBG = 1 (no TMA traffic) / 2 (no MMAs) / 4 (no stores)
  python profiles/conv_bench.py plans [reps]  tile plans SS2_DC_PLAN = "column tiles,blocks per tile" against the planner's choice
One JSON line per measurement."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stabstitch2_b200 import _lib  # noqa: E402

LAYERS = [("layer1 64->64 90x120", 32, 90, 120, 64, 64), ("layer2 128->128 45x60", 32, 45, 60, 128, 128),
          ("layer3 256->256 23x30", 32, 23, 30, 256, 256), ("regressor 128->128 45x60 B16", 16, 45, 60, 128, 128),
          ("regressor 256->256 12x15 B16", 16, 12, 15, 256, 256), ("regressor 64->64 23x30 B16", 16, 23, 30, 64, 64)]


def timed(ctx, reps, fn0):
    def fn():
        try:
            return fn0()
        except _lib.SS2Error as exc:   # the dbg modes compute on stale shared memory: the fp16 range flag may fire
            if "fp16 range" not in str(exc):
                raise
            return fn0()
    fn()
    ctx.profile_enable(_lib.PROF_CONV, True)
    for _ in range(reps):
        fn()
    ms, n, flops = ctx.profile_read(_lib.PROF_CONV)
    ctx.profile_enable(_lib.PROF_CONV, False)
    return 1e3 * ms / max(n, 1), flops / max(ms, 1e-9) / 1e9


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "plans"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    ctx = _lib.context()
    # SS2_CONV_TEST_F16=3 in the environment: fp16 split planes in and out (conv1 of a BasicBlock), kind::f16 MMAs
    if not os.environ.get("SS2_CONV_TEST_F16"):
        os.environ["SS2_CONV_TEST_SPLIT"] = "1"
    g = torch.Generator().manual_seed(0)
    only = os.environ.get("CONV_BENCH_ONLY")
    for name, B, H, W, Cin, Cout in LAYERS:
        if only and only not in name:
            continue
        x = torch.randn(B, H, W, Cin, generator=g).cuda()
        w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
        b = torch.randn(Cout, generator=g)
        res = torch.randn(B, H, W, Cout, generator=g).cuda()
        for residual in (False, True):
            call = lambda: _lib.conv_nhwc(x, w, b, stride=1, pad=1, relu=True, residual=res if residual else None, use_tc=True)
            if what == "dbg":
                for mode in ("0", "1", "2", "4"):
                    os.environ["SS2_DC_DBG"] = mode
                    us, tf = timed(ctx, reps, call)
                    print(json.dumps({"layer": name, "residual": residual, "mode": {"0": "full", "1": "MMA only", "2": "TMA only", "4": "no stores"}[mode],
                                      "us": round(us, 1), "tflops_alg": round(tf, 1)}), flush=True)
                os.environ["SS2_DC_DBG"] = "0"
            else:
                os.environ.pop("SS2_DC_PLAN", None)
                us, tf = timed(ctx, reps, call)
                print(json.dumps({"layer": name, "residual": residual, "plan": "planner", "us": round(us, 1), "tflops_alg": round(tf, 1)}), flush=True)
                for tw in (1, 2, 3, 4):
                    for mt in (2, 1):
                        os.environ["SS2_DC_PLAN"] = "%d,%d" % (tw, mt)
                        try:
                            us, tf = timed(ctx, reps, call)
                        except Exception as exc:      # plan does not fit / not allowed for this layer
                            ctx.profile_enable(_lib.PROF_CONV, False)
                            continue
                        print(json.dumps({"layer": name, "residual": residual, "plan": "%d column tiles x %d blocks" % (tw, mt),
                                          "us": round(us, 1), "tflops_alg": round(tf, 1)}), flush=True)
                os.environ.pop("SS2_DC_PLAN", None)


if __name__ == "__main__":
    main()
