"""Metric path of the reference (Full_model_inference/Codes/test_metric_ssd.py) on libss2: the per-view warp with its
mask planes, the whole-stream path of view 2, and the four scores the script prints per video.

  get_stable_sqe(img1_list, img2_list, smooth_mesh1, smooth_mesh2)   same signature / return as :151-181
  stream_paths(win_ori_path2, win_smooth_path2)                      :417-436
  stability_score(path) / distortion_score(mesh)                     :455-466 / :470-479
  psnr_ssim(warp1, warp2)                                            :513-518 (skimage 0.15 semantics)
"""
import torch

from . import _lib


def warp_with_mask(imgs, meshes):
    """imgs [n,3,H,W] in [-1,1] (the network inputs), meshes [n,7,9,2] @ the image size -> [n,6,H,W] CUDA:
    TPS warp of [(img+1)*127.5, ones x 3] onto the image's own grid (:163-173)."""
    from .spatial_network import get_norm_mesh, get_rigid_mesh
    from .utils.torch_tps_transform import transformer
    x = (_lib.dev_f32(imgs) + 1) * 127.5
    n, _, H, W = x.shape
    m = _lib.dev_f32(meshes).reshape(n, 7, 9, 2)
    nrig = get_norm_mesh(get_rigid_mesh(n, H, W), H, W).to(x.device)
    return transformer(torch.cat([x, torch.ones_like(x)], 1), get_norm_mesh(m, H, W), nrig, (H, W), mode="NORMAL")


def get_stable_sqe(img1_list, img2_list, smooth_mesh1, smooth_mesh2):
    """Drop-in for test_metric_ssd.py:151-181: lists of [1,3,360,480] tensors in [-1,1], smooth meshes [1,N,7,9,2]
    -> (list of [H,W,6] numpy, list of [H,W,6] numpy)."""
    w1 = warp_with_mask(torch.cat(list(img1_list), 0), _lib.dev_f32(smooth_mesh1)[0])
    w2 = warp_with_mask(torch.cat(list(img2_list), 0), _lib.dev_f32(smooth_mesh2)[0])
    a, b = w1.permute(0, 2, 3, 1).cpu().numpy(), w2.permute(0, 2, 3, 1).cpu().numpy()
    return [a[k] for k in range(a.shape[0])], [b[k] for k in range(b.shape[0])]


def stream_paths(win_ori_path, win_smooth_path):
    """per-window outputs [nwin,7,7,9,2] of build_SmoothNet ('ori_path2', 'smooth_path2') -> whole-stream
    (ori_path, smooth_path) [nwin+6,7,9,2] (:417-436)."""
    ctx = _lib.context()
    a, b = _lib.dev_f32(win_ori_path), _lib.dev_f32(win_smooth_path)
    nwin = a.shape[0]
    ori = torch.empty(nwin + 6, 7, 9, 2, device=a.device, dtype=torch.float32)
    smo = torch.empty_like(ori)
    ctx.check(ctx.lib.ss2_assemble_paths(ctx.handle, _lib.ptr(a), _lib.ptr(b), nwin, _lib.ptr(ori), _lib.ptr(smo),
                                         _lib.cur_stream()))
    return ori, smo


def _scores(path, mesh):
    ctx = _lib.context()
    p = _lib.dev_f32(path).reshape(-1, 7, 9, 2) if path is not None else None
    m = _lib.dev_f32(mesh).reshape(-1, 7, 9, 2) if mesh is not None else None
    n = (p if p is not None else m).shape[0]
    out = torch.empty(2, device=(p if p is not None else m).device, dtype=torch.float32)
    ctx.check(ctx.lib.ss2_metric_scores(ctx.handle, _lib.ptr(p), _lib.ptr(m), n, _lib.ptr(out), _lib.cur_stream()))
    return out


def stability_score(path):
    """multi-lag squared distance of a path [N,7,9,2] (N >= 7), weights 0.9 / 0.3 / 0.1 (:455-466) -> 0-dim tensor"""
    return _scores(path, None)[0]


def distortion_score(mesh):
    """max over frames of inter_grid_loss + intra_grid_loss of a mesh sequence [N,7,9,2] (:470-479) -> 0-dim tensor"""
    return _scores(None, mesh)[1]


def psnr_ssim(warp1, warp2):
    """warp1, warp2 [n,6,H,W] (warp_with_mask) -> (psnr [n], ssim [n]) of the views inside their overlap (:513-518)."""
    ctx = _lib.context()
    a, b = _lib.dev_f32(warp1), _lib.dev_f32(warp2)
    n, C, H, W = a.shape
    if C != 6 or b.shape != a.shape:
        raise ValueError("psnr_ssim takes two [n,6,H,W] tensors")
    ps = torch.empty(n, device=a.device, dtype=torch.float32)
    ss = torch.empty_like(ps)
    ctx.check(ctx.lib.ss2_metric_psnr_ssim(ctx.handle, _lib.ptr(a), _lib.ptr(b), n, H, W, _lib.ptr(ps), _lib.ptr(ss),
                                           _lib.cur_stream()))
    return ps, ss
