"""CPU restatement of the reference's METRIC path - TEST INFRASTRUCTURE like the rest of oracle/.

    Full_model_inference/Codes/test_metric_ssd.py
      :37-67    inter_grid_loss      angle preservation between successive mesh edges
      :72-88    intra_grid_loss      edges longer than two rigid cells
      :33-34    l_num_loss
      :151-181  get_stable_sqe       per-view TPS warp of [image, ones x 3] at 360x480 (no canvas, no fusion)
      :431-436  accumulation of the whole-stream original / smoothed path of view 2 from the per-window outputs
      :444-466  stability score      multi-lag squared distance of the smoothed path, weights 0.9 / 0.3 / 0.1
      :470-479  distortion score     max over frames of inter + intra grid loss of the smoothed mesh of view 2
      :513-518  PSNR / SSIM of the two warped views inside their overlap

PSNR / SSIM are skimage 0.15 calls (third party: scikit-image pinned by environment.yml, `compare_psnr`,
`compare_ssim(multichannel=True)`; not vendored, NOT INSTALLED in this image): restated from the published
algorithm (uniform 7x7 window with scipy.ndimage.uniform_filter reflect borders, sample covariance, K1 = 0.01,
K2 = 0.03, float64, mean over the image cropped by 3 pixels, mean over channels) -> "parity unpinned" for these two;
everything else is pinned against the reference's own functions in tests/golden/metric.npz.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import stabstitch_oracle as O

GRID_H, GRID_W = 6, 8


def l_num_loss(a, b, l_num=1):
    return torch.mean(torch.abs((a - b) ** l_num))                                           # :33-34


def inter_grid_loss(mesh):
    """mesh [bs,T,7,9,2] (:37-67)"""
    gw, gh = GRID_W, GRID_H
    w_edges = mesh[:, :, :, 0:gw, :] - mesh[:, :, :, 1:gw + 1, :]
    cos_w = torch.sum(w_edges[:, :, :, 0:gw - 1, :] * w_edges[:, :, :, 1:gw, :], 3) / (
        torch.sqrt(torch.sum(w_edges[:, :, :, 0:gw - 1, :] * w_edges[:, :, :, 0:gw - 1, :], 3)) *
        torch.sqrt(torch.sum(w_edges[:, :, :, 1:gw, :] * w_edges[:, :, :, 1:gw, :], 3)))
    dw = 1 - cos_w
    dw = dw[:, :, 0:gh, :] + dw[:, :, 1:gh + 1, :]
    h_edges = mesh[:, :, 0:gh, :, :] - mesh[:, :, 1:gh + 1, :, :]
    cos_h = torch.sum(h_edges[:, :, 0:gh - 1, :, :] * h_edges[:, :, 1:gh, :, :], 3) / (
        torch.sqrt(torch.sum(h_edges[:, :, 0:gh - 1, :, :] * h_edges[:, :, 0:gh - 1, :, :], 3)) *
        torch.sqrt(torch.sum(h_edges[:, :, 1:gh, :, :] * h_edges[:, :, 1:gh, :, :], 3)))
    dh = 1 - cos_h
    dh = dh[:, :, :, 0:gw] + dh[:, :, :, 1:gw + 1]
    return torch.mean(dw) + torch.mean(dh)


def intra_grid_loss(pts):
    """:72-88"""
    max_w, max_h = 480 / GRID_W * 2, 360 / GRID_H * 2
    dx = pts[:, :, :, 1:GRID_W + 1, 0] - pts[:, :, :, 0:GRID_W, 0]
    dy = pts[:, :, 1:GRID_H + 1, :, 1] - pts[:, :, 0:GRID_H, :, 1]
    return torch.mean(F.relu(dx - max_w)) + torch.mean(F.relu(dy - max_h))


def accumulate_paths(win_ori_path2, win_smooth_path2):
    """:417-436  per-window outputs [nwin,7,7,9,2] -> whole-stream (ori_path2, smooth_path2) [1,nwin+6,7,9,2]"""
    ori, smo = win_ori_path2[0:1].clone(), win_smooth_path2[0:1].clone()
    for k in range(1, win_ori_path2.shape[0]):
        o, s = win_ori_path2[k:k + 1], win_smooth_path2[k:k + 1]
        new_ori = ori[:, -1] + (o[:, -1] - o[:, -2])
        ori = torch.cat((ori, new_ori.unsqueeze(1)), 1)
        new_smo = ori[:, -1] + (s[:, -1] - o[:, -1])
        smo = torch.cat((smo, new_smo.unsqueeze(1)), 1)
    return ori, smo


def stability_score(path):
    """:455-466 on a path [1,N,7,9,2]"""
    mid = path[:, 3:-3]
    s = (l_num_loss(path[:, :-6], mid, 2) + l_num_loss(path[:, 6:], mid, 2)) * 0.1
    s = s + (l_num_loss(path[:, 1:-5], mid, 2) + l_num_loss(path[:, 5:-1], mid, 2)) * 0.3
    s = s + (l_num_loss(path[:, 2:-4], mid, 2) + l_num_loss(path[:, 4:-2], mid, 2)) * 0.9
    return s


def distortion_score(mesh):
    """:470-479  max over frames of inter + intra grid loss; mesh [1,N,7,9,2]"""
    vals = [float(inter_grid_loss(mesh[:, k:k + 1]) + intra_grid_loss(mesh[:, k:k + 1])) for k in range(mesh.shape[1])]
    return max(vals)


def metric_warp(img, mesh):
    """:151-181 for one view and frame: img [1,3,360,480] in [-1,1], mesh [1,7,9,2] @480x360 -> [6,360,480]
    (warped image 0..255, warped ones x 3)."""
    _, _, H, W = img.shape
    nrig = O.norm_mesh(O.rigid_mesh(1, H, W), H, W)
    x = (img + 1) * 127.5
    return O.tps_warp(torch.cat([x, torch.ones_like(x)], 1), O.norm_mesh(mesh, H, W), nrig, (H, W))[0]


def psnr_overlap(w1, w2):
    """:513-516  w [H,W,6] numpy (channels 0:3 image, 3:6 mask).  skimage.measure.compare_psnr(a, b, 255)."""
    ov = w1[..., 3:6] * w2[..., 3:6]
    a, b = (w1[..., 0:3] * ov).astype(np.float64), (w2[..., 0:3] * ov).astype(np.float64)
    mse = np.mean((a - b) ** 2)
    return 10 * np.log10(255.0 ** 2 / mse)


def ssim_overlap(w1, w2):
    """:517  skimage.measure.compare_ssim(a, b, data_range=255, multichannel=True) of skimage 0.15 (defaults: 7x7 uniform
    window, sample covariance, K1 = 0.01, K2 = 0.03)."""
    from scipy.ndimage import uniform_filter
    ov = w1[..., 3:6] * w2[..., 3:6]
    A, B = (w1[..., 0:3] * ov).astype(np.float64), (w2[..., 0:3] * ov).astype(np.float64)
    win, R = 7, 255.0
    C1, C2 = (0.01 * R) ** 2, (0.03 * R) ** 2
    NP = win * win
    cov_norm = NP / (NP - 1.0)
    vals = []
    for ch in range(3):
        X, Y = A[..., ch], B[..., ch]
        ux, uy = uniform_filter(X, size=win), uniform_filter(Y, size=win)
        uxx, uyy, uxy = uniform_filter(X * X, size=win), uniform_filter(Y * Y, size=win), uniform_filter(X * Y, size=win)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        pad = (win - 1) // 2
        vals.append(S[pad:-pad, pad:-pad].mean())
    return float(np.mean(vals))
