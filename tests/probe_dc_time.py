"""GPU probe (not a pytest file): where the direct convolution's time goes.  SS2_DC_DBG=1 drops all TMA traffic
(MMAs run on stale shared memory), =2 drops the MMAs (TMA + barriers only)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stabstitch2_b200 import _lib  # noqa: E402


def timeit(fn, reps=5):
    """kernel time from the library's CUDA-event brackets around the convolution launch"""
    ctx = _lib.context()
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ctx.profile_enable(_lib.PROF_CONV, True)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    ms, n, _ = ctx.profile_read(_lib.PROF_CONV)
    ctx.profile_enable(_lib.PROF_CONV, False)
    return ms / max(n, 1) * 1e3


def main():
    for (B, H, W, Cin, Cout) in ((32, 90, 120, 64, 64), (32, 45, 60, 128, 128), (32, 23, 30, 256, 256)):
        x = torch.randn(B, H, W, Cin).cuda()
        w = torch.randn(Cout, Cin, 3, 3) / (Cin * 9) ** 0.5
        line = []
        for dbg in ("0", "1", "2"):
            os.environ["SS2_DC_DBG"] = dbg
            line.append("dbg%s %.0f us" % (dbg, timeit(lambda: _lib.conv_nhwc(x, w, None, stride=1, pad=1, relu=True, use_tc=True))))
        os.environ["SS2_DC_DBG"] = "0"
        print("conv %dx%dx%d %d->%d: %s" % (B, H, W, Cin, Cout, "  ".join(line)))


if __name__ == "__main__":
    main()
