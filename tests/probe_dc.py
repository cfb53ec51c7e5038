"""GPU probe (not a pytest file): the direct 3x3 tcgen05 convolution tap by tap against torch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stabstitch2_b200 import _lib  # noqa: E402


def main():
    B, H, W, Cin, Cout = 1, 8, 60, 32, 64
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, H, W, Cin, generator=g)
    for mode in ("-",):
        line = []
        for tap in range(9):
            w = torch.zeros(Cout, Cin, 3, 3)
            w[:, :, tap // 3, tap % 3] = torch.randn(Cout, Cin, generator=torch.Generator().manual_seed(tap)) / Cin ** 0.5
            ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), None, 1, 1).permute(0, 2, 3, 1).float()
            out = _lib.conv_nhwc(x.cuda(), w, None, stride=1, pad=1, relu=False, use_tc=True).cpu()
            e = (out - ref).abs().amax(dim=(0, 3))  # [H, W]
            line.append("%.1e" % e.max().item())
            if tap in (0, 4) and e.max() > 1e-3:
                bad = (e > 1e-3)
                print("mode %s tap %d bad rows %s bad cols(first 16) %s" % (mode, tap, bad.any(1).nonzero().flatten().tolist(),
                                                                           bad.any(0).nonzero().flatten().tolist()[:16]))
        print("bo_mode %s: max err per tap: %s" % (mode, " ".join(line)))


if __name__ == "__main__":
    main()
