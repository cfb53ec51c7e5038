import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import stabstitch_oracle as O
from stabstitch2_b200 import pipeline, _lib
from tests.test_gpu_parity import _smooth_frame
H, W = 720, 1280
g = torch.Generator().manual_seed(5)
rig = O.rigid_mesh(1, 360, 480)[:, None]
def mesh(dx, dy):
    return rig + torch.tensor([dx, dy]) + 2.5 * torch.randn(1, 1, 7, 9, 2, generator=g)
w12m1, w12m2, w23m1, w23m2 = mesh(-80.0, 3.0), mesh(85.0, -2.0), mesh(-70.0, 6.0), mesh(95.0, 1.0)
imgs = [_smooth_frame(10 + v, H, W) for v in range(3)]
with torch.no_grad():
    rm1, rmid, rm3, wmin, hmin, ow, oh = O.three_view_meshes(w12m1, w12m2, w23m1, w23m2, H, W)
    ref = O.three_view_frame(imgs[0], imgs[1], imgs[2], rm1[:, 0], rmid[:, 0], rm3[:, 0], wmin, hmin, ow, oh)
Ho, Wo = ref.shape[1:]
nrig = O.norm_mesh(O.rigid_mesh(1, H, W), H, W)
cov = []
coords = []
for M in (rm1, rmid, rm3):
    tt = torch.stack([M[0, 0, ..., 0] - wmin, M[0, 0, ..., 1] - hmin], 2)[None]
    ax, ay = O.tps_source_coords_fp64(O.norm_mesh(tt, oh, ow), nrig, Ho, Wo, W, H)
    cov.append((ax[0] > 1) & (ax[0] < W - 2) & (ay[0] > 1) & (ay[0] < H - 2)); coords.append((ax[0], ay[0]))
for tps in (_lib.TPS_EXACT, _lib.TPS_LATTICE):
    fused = pipeline.three_view_frames(imgs[0].cuda(), imgs[1].cuda(), imgs[2].cuda(), rm1[0].cuda(), rmid[0].cuda(), rm3[0].cuda(),
                                       [float(wmin), float(hmin), float(ow), float(oh)], tps=tps)[0].cpu()
    d = (fused - ref).abs().numpy().max(0)
    ncov = cov[0].astype(int) + cov[1] + cov[2]
    bad = d > 0.1
    print("tps", tps, "bad", bad.sum(), "by coverage count", [int((bad & (ncov == k)).sum()) for k in range(4)], "pixels by count", [int((ncov == k).sum()) for k in range(4)])
    ys, xs = np.nonzero(bad & (ncov > 0))
    for i in range(0, len(ys), max(1, len(ys) // 8)):
        y, x = ys[i], xs[i]
        print(" px", y, x, "ref", ref[:, y, x].numpy(), "ours", fused[:, y, x].numpy(), "coords", [(round(float(c[0][y, x]), 3), round(float(c[1][y, x]), 3)) for c in coords])
