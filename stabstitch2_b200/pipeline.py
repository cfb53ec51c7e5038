"""Sequence glue of the hot path (the functions the reference keeps inside its driver script,
Full_model_inference/Codes/test_online_tra.py) on top of libss2.

  get_stable_sqe(...)      same signature/return as test_online_tra.py:96 (AVERAGE and LINEAR fusion)
  stream_meshes(...)       test_online_tra.py:284-392 for a device-resident stream
  stitch_stream(...)       device-resident whole stream -> fused frames [n,3,Ho,Wo]
  stitch_stream_host(...)  the same through HOST buffers in one C call (H2D + D2H inside)
"""
import ctypes

import numpy as np
import torch

from . import _lib, grid_res
from .smooth_network import smooth_windows

grid_h = grid_res.GRID_H
grid_w = grid_res.GRID_W
WINDOW = 7


def _sync_nets(ctx, spatial_net, temporal_net, smooth_net):
    spatial_net.sync_weights(ctx)
    temporal_net.sync_weights(ctx)
    smooth_net.sync_weights(ctx)


def canvas_minmax(mesh1, mesh2, img_h, img_w):
    """test_online_tra.py:103-117 -> CUDA tensor [4] = (xmin, xmax, ymin, ymax) in hr pixels."""
    ctx = _lib.context()
    m1 = _lib.dev_f32(mesh1).reshape(-1, 7, 9, 2)
    m2 = _lib.dev_f32(mesh2).reshape(-1, 7, 9, 2)
    out = torch.empty(4, device=m1.device, dtype=torch.float32)
    ctx.check(ctx.lib.ss2_canvas_minmax(ctx.handle, _lib.ptr(m1), _lib.ptr(m2), m1.shape[0], int(img_h), int(img_w),
                                        _lib.ptr(out), _lib.cur_stream()))
    return out


def canvas_size(minmax_host):
    ctx_lib = _lib.load_library()
    mm = (ctypes.c_float * 4)(*[float(v) for v in minmax_host])
    h, w = ctypes.c_int(), ctypes.c_int()
    ctx_lib.ss2_canvas_size(mm, ctypes.byref(h), ctypes.byref(w))
    return h.value, w.value


def stable_frames(hr1, hr2, mesh1, mesh2, minmax_host, mode="NORMAL", tps=None, out=None, fusion_mode="AVERAGE"):
    """The get_stable_sqe loop for n frames given the (global) canvas: hr [n,3,H,W] CUDA,
    meshes [n,7,9,2] @480x360 -> fused [n,3,Ho,Wo]; fusion_mode 'AVERAGE' (test_online_tra.py:142) or 'LINEAR'
    (:143-150)."""
    from .utils import torch_tps_transform as tt
    if fusion_mode not in ("AVERAGE", "LINEAR"):
        raise ValueError("fusion_mode must be 'AVERAGE' or 'LINEAR'")
    ctx = _lib.context()
    hr1, hr2 = _lib.dev_f32(hr1), _lib.dev_f32(hr2)
    m1 = _lib.dev_f32(mesh1).reshape(-1, 7, 9, 2)
    m2 = _lib.dev_f32(mesh2).reshape(-1, 7, 9, 2)
    n, _, H, W = hr1.shape
    Ho, Wo = canvas_size(minmax_host)
    if out is None:
        out = torch.empty(n, 3, Ho, Wo, device=hr1.device, dtype=torch.float32)
    mm = (ctypes.c_float * 4)(*[float(v) for v in minmax_host])
    fn = ctx.lib.ss2_stable_frames if fusion_mode == "AVERAGE" else ctx.lib.ss2_stable_frames_linear
    ctx.check(fn(ctx.handle, _lib.ptr(hr1), _lib.ptr(hr2), _lib.ptr(m1), _lib.ptr(m2), n, H, W,
                 mm, _lib.MODE[mode], tt.DEFAULT_TPS if tps is None else tps, _lib.ptr(out), _lib.cur_stream()))
    return out


def stable_frames_u8(hr1, hr2, mesh1, mesh2, minmax_host, mode="NORMAL", tps=None, out=None):
    """stable_frames (AVERAGE) with the driver's `.astype(uint8)` back end (test_online_tra.py:152,414) fused into the
    resampler's store: fused frames [n,Ho,Wo,3] uint8, bit-identical to stable_frames + frames_to_u8.  Lattice resampler
    only (raises SS2Error for canvases too small for it)."""
    from .utils import torch_tps_transform as tt
    ctx = _lib.context()
    hr1, hr2 = _lib.dev_f32(hr1), _lib.dev_f32(hr2)
    m1 = _lib.dev_f32(mesh1).reshape(-1, 7, 9, 2)
    m2 = _lib.dev_f32(mesh2).reshape(-1, 7, 9, 2)
    n, _, H, W = hr1.shape
    Ho, Wo = canvas_size(minmax_host)
    if out is None:
        out = torch.empty(n, Ho, Wo, 3, device=hr1.device, dtype=torch.uint8)
    mm = (ctypes.c_float * 4)(*[float(v) for v in minmax_host])
    ctx.check(ctx.lib.ss2_stable_frames_u8(ctx.handle, _lib.ptr(hr1), _lib.ptr(hr2), _lib.ptr(m1), _lib.ptr(m2), n, H, W,
                                           mm, _lib.MODE[mode], tt.DEFAULT_TPS if tps is None else tps, _lib.ptr(out),
                                           _lib.cur_stream()))
    return out


def linear_blender(ref, tgt, ref_m, tgt_m, mask=False):
    """Drop-in for the driver's linear_blender (test_online_tra.py:34-58): ref, tgt [n,3,Ho,Wo]; ref_m, tgt_m
    [n,1,Ho,Wo] -> stitched [n,3,Ho,Wo] (mask=True: mask1 [n,1,Ho,Wo]).  Masks are thresholded at 0.5 (see
    ss2_linear_blend)."""
    ctx = _lib.context()
    ref, tgt, ref_m, tgt_m = (_lib.dev_f32(t) for t in (ref, tgt, ref_m, tgt_m))
    n, C, Ho, Wo = ref.shape
    if C != 3 or tgt.shape != ref.shape or ref_m.shape != (n, 1, Ho, Wo) or tgt_m.shape != (n, 1, Ho, Wo):
        raise ValueError("linear_blender takes [n,3,Ho,Wo] images and [n,1,Ho,Wo] masks")
    out = None if mask else torch.empty_like(ref)
    m1 = torch.empty_like(ref_m) if mask else None
    ctx.check(ctx.lib.ss2_linear_blend(ctx.handle, _lib.ptr(ref), _lib.ptr(tgt), 3 * Ho * Wo, _lib.ptr(ref_m), _lib.ptr(tgt_m),
                                       Ho * Wo, n, Ho, Wo, _lib.ptr(out), _lib.ptr(m1), _lib.cur_stream()))
    return m1 if mask else out


def three_view_meshes(w12m1, w12m2, w23m1, w23m2, img_h, img_w):
    """Middle-plane alignment of two stitched pairs sharing their middle view
    (test_online_tra_threeview.py:345-455).  Smooth meshes [n,7,9,2] @480x360 (w12m2 and w23m1 are the two
    instances of the shared view) -> (mesh1, middle, mesh3 [n,7,9,2] CUDA, canvas [4] CUDA =
    width_min, height_min, out_width, out_height of the new canvas)."""
    ctx = _lib.context()
    ms = [_lib.dev_f32(t).reshape(-1, 7, 9, 2) for t in (w12m1, w12m2, w23m1, w23m2)]
    n = ms[0].shape[0]
    outs = [torch.empty(n, 7, 9, 2, device=ms[0].device, dtype=torch.float32) for _ in range(3)]
    canvas = torch.empty(4, device=ms[0].device, dtype=torch.float32)
    ctx.check(ctx.lib.ss2_three_view_meshes(ctx.handle, *[_lib.ptr(t) for t in ms], n, int(img_h), int(img_w),
                                            *[_lib.ptr(t) for t in outs], _lib.ptr(canvas), _lib.cur_stream()))
    return outs[0], outs[1], outs[2], canvas


def three_view_frames(imgs1, imgs2, imgs3, mesh1, middle, mesh3, canvas_host, mode="NORMAL", tps=None,
                      fusion_mode="AVERAGE"):
    """The three-image warp + fusion loop (test_online_tra_threeview.py:461-503, AVERAGE or LINEAR): images
    [n,3,H,W] CUDA per view, meshes from three_view_meshes, canvas (4 host floats) -> fused [n,3,Ho,Wo]."""
    from .utils import torch_tps_transform as tt
    ctx = _lib.context()
    a, b, c = _lib.dev_f32(imgs1), _lib.dev_f32(imgs2), _lib.dev_f32(imgs3)
    n, _, H, W = a.shape
    cv = [float(v) for v in canvas_host]
    Ho, Wo = int(cv[3]), int(cv[2])
    out = torch.empty(n, 3, Ho, Wo, device=a.device, dtype=torch.float32)
    hc = (ctypes.c_float * 4)(*cv)
    if fusion_mode not in ("AVERAGE", "LINEAR"):
        raise ValueError("fusion_mode must be 'AVERAGE' or 'LINEAR'")
    fn = ctx.lib.ss2_three_view_frames if fusion_mode == "AVERAGE" else ctx.lib.ss2_three_view_frames_linear
    ctx.check(fn(ctx.handle, _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(_lib.dev_f32(mesh1)),
                 _lib.ptr(_lib.dev_f32(middle)), _lib.ptr(_lib.dev_f32(mesh3)), n, H, W, hc,
                 _lib.MODE[mode], tt.DEFAULT_TPS if tps is None else tps, _lib.ptr(out), _lib.cur_stream()))
    return out


def three_view_stable(img1_list, img2_list, img3_list, w12m1, w12m2, w23m1, w23m2, warp_mode="NORMAL",
                      fusion_mode="AVERAGE"):
    """Drop-in for the tail of test_online_tra_threeview.py's test() (:345-505, AVERAGE or LINEAR fusion): image lists of
    [1,3,H,W] fp32 0..255, smooth meshes [1,N,7,9,2] of the two pairs -> list of N CPU tensors [3,Ho,Wo]."""
    a = torch.cat([_lib.dev_f32(t) for t in img1_list], 0)
    b = torch.cat([_lib.dev_f32(t) for t in img2_list], 0)
    c = torch.cat([_lib.dev_f32(t) for t in img3_list], 0)
    _, _, H, W = a.shape
    m1, mid, m3, canvas = three_view_meshes(w12m1, w12m2, w23m1, w23m2, H, W)
    fused = three_view_frames(a, b, c, m1, mid, m3, canvas.cpu().tolist(), warp_mode, fusion_mode=fusion_mode)
    host = fused.cpu()
    return [host[k] for k in range(host.shape[0])]


# ------------------------------------------------------------------------------------------
# N views (BASELINE.json config 5): chain of pairs (1,2), (2,3), .., (N-1,N); identical to the three-view glue for N = 3
# ------------------------------------------------------------------------------------------
def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def nview_align(pair_meshes, img_h, img_w):
    """pair_meshes: list of N-1 (meshA, meshB) smooth meshes [n,7,9,2] @480x360 -> (shifted [2(N-1),n,7,9,2],
    mids [N-2,n,7,9,2], minmax1 [4] CUDA = provisional canvas (xmin,xmax,ymin,ymax))."""
    ctx = _lib.context()
    flat = [_lib.dev_f32(t).reshape(-1, 7, 9, 2) for pr in pair_meshes for t in pr]
    nviews, n = len(pair_meshes) + 1, flat[0].shape[0]
    dev = flat[0].device
    shifted = torch.empty(2 * (nviews - 1), n, 7, 9, 2, device=dev, dtype=torch.float32)
    mids = torch.empty(max(nviews - 2, 1), n, 7, 9, 2, device=dev, dtype=torch.float32)
    mm = torch.empty(4, device=dev, dtype=torch.float32)
    ctx.check(ctx.lib.ss2_nview_align(ctx.handle, _ptr_array(flat), nviews, n, int(img_h), int(img_w), _lib.ptr(shifted),
                                      _lib.ptr(mids), _lib.ptr(mm), _lib.cur_stream()))
    return shifted, mids, mm


def nview_remap(shifted, mids, minmax1_host):
    """-> (meshes [N,n,7,9,2] in provisional-canvas pixels, minmax2 [4] CUDA = new canvas)."""
    ctx = _lib.context()
    nviews, n = shifted.shape[0] // 2 + 1, shifted.shape[1]
    meshes = torch.empty(nviews, n, 7, 9, 2, device=shifted.device, dtype=torch.float32)
    mm2 = torch.empty(4, device=shifted.device, dtype=torch.float32)
    h = (ctypes.c_float * 4)(*[float(v) for v in minmax1_host])
    ctx.check(ctx.lib.ss2_nview_remap(ctx.handle, nviews, n, _lib.ptr(shifted), _lib.ptr(mids), h, _lib.ptr(meshes),
                                      _lib.ptr(mm2), _lib.cur_stream()))
    return meshes, mm2


def nview_frames(imgs, meshes, minmax2_host, mode="NORMAL", tps=None, out=None):
    """imgs: list of N CUDA tensors [n,3,H,W]; meshes [N,n,7,9,2] -> fused [n,3,Ho,Wo] (one fused pass, 2 <= N <= 4)."""
    from .utils import torch_tps_transform as tt
    ctx = _lib.context()
    imgs = [_lib.dev_f32(t) for t in imgs]
    n, _, H, W = imgs[0].shape
    Ho, Wo = canvas_size(minmax2_host)
    if out is None:
        out = torch.empty(n, 3, Ho, Wo, device=imgs[0].device, dtype=torch.float32)
    h = (ctypes.c_float * 4)(*[float(v) for v in minmax2_host])
    ctx.check(ctx.lib.ss2_nview_frames(ctx.handle, _ptr_array(imgs), _lib.ptr(_lib.dev_f32(meshes)), len(imgs), n, H, W, h,
                                       _lib.MODE[mode], tt.DEFAULT_TPS if tps is None else tps, _lib.ptr(out),
                                       _lib.cur_stream()))
    return out


def nview_stable(img_lists, pair_meshes, warp_mode="NORMAL", group=None):
    """N image streams (lists of CUDA/CPU tensors [n,3,H,W] or stacked tensors) + the smooth meshes of the N-1 stitched
    pairs -> fused [n,3,Ho,Wo].  With a torch.distributed `group` (or a default group initialised) of more than one
    rank, the frames are this rank's temporal shard and the two canvases are all-reduced (allreduce_canvas)."""
    import torch.distributed as dist
    imgs = [torch.cat(list(t), 0) if isinstance(t, (list, tuple)) else t for t in img_lists]
    imgs = [_lib.dev_f32(t) for t in imgs]
    _, _, H, W = imgs[0].shape
    sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    shifted, mids, mm1 = nview_align(pair_meshes, H, W)
    if sharded:
        mm1 = allreduce_canvas(mm1, group)
    meshes, mm2 = nview_remap(shifted, mids, mm1.cpu().tolist())
    if sharded:
        mm2 = allreduce_canvas(mm2, group)
    return nview_frames(imgs, meshes, mm2.cpu().tolist(), warp_mode), meshes, mm2


def get_stable_sqe(img1_list, img2_list, smooth_mesh1, smooth_mesh2, warp_mode="NORMAL", fusion_mode="AVERAGE"):
    """Drop-in for test_online_tra.py:96-154 (AVERAGE and LINEAR fusion): lists of [1,3,H,W] fp32 0..255
    frames, smooth meshes [1,N,7,9,2] -> (list of [Ho,Wo,3] numpy frames, out_width, out_height)."""
    hr1 = torch.cat([_lib.dev_f32(t) for t in img1_list], 0)
    hr2 = torch.cat([_lib.dev_f32(t) for t in img2_list], 0)
    _, _, H, W = hr2.shape
    m1 = _lib.dev_f32(smooth_mesh1).reshape(-1, 7, 9, 2)
    m2 = _lib.dev_f32(smooth_mesh2).reshape(-1, 7, 9, 2)
    mm = canvas_minmax(m1, m2, H, W).cpu().tolist()
    fused = stable_frames(hr1, hr2, m1, m2, mm, warp_mode, fusion_mode=fusion_mode)
    Ho, Wo = fused.shape[2:]
    host = fused.permute(0, 2, 3, 1).contiguous().cpu().numpy()
    return ([host[k] for k in range(host.shape[0])], torch.tensor(Wo, dtype=torch.int32),
            torch.tensor(Ho, dtype=torch.int32))


def stream_meshes(spatial_net, temporal_net, smooth_net, lr1, lr2, want_raw=False):
    """lr1, lr2 [n,3,360,480] CUDA in [-1,1] -> smooth meshes ([n,7,9,2], [n,7,9,2])."""
    ctx = _lib.context()
    _sync_nets(ctx, spatial_net, temporal_net, smooth_net)
    lr1, lr2 = _lib.dev_f32(lr1), _lib.dev_f32(lr2)
    n = lr1.shape[0]
    mk = lambda: torch.empty(n, 7, 9, 2, device=lr1.device, dtype=torch.float32)  # noqa: E731
    s1, s2 = mk(), mk()
    raw = [mk() for _ in range(4)] if want_raw else [None] * 4
    ctx.check(ctx.lib.ss2_stream_meshes(ctx.handle, _lib.ptr(lr1), _lib.ptr(lr2), n, _lib.ptr(s1), _lib.ptr(s2),
                                        *[_lib.ptr(r) for r in raw], _lib.cur_stream()))
    if want_raw:
        return s1, s2, dict(smotion1=raw[0], smotion2=raw[1], tmotion1=raw[2], tmotion2=raw[3])
    return s1, s2


def stitch_stream(spatial_net, temporal_net, smooth_net, lr1, lr2, hr1, hr2, mode="NORMAL", tps=None):
    """Whole hot path, device resident: returns (fused [n,3,Ho,Wo], smooth_mesh1, smooth_mesh2)."""
    s1, s2 = stream_meshes(spatial_net, temporal_net, smooth_net, lr1, lr2)
    _, _, H, W = hr1.shape
    mm = canvas_minmax(s1, s2, H, W).cpu().tolist()  # the one data-dependent host read (16 B)
    return stable_frames(hr1, hr2, s1, s2, mm, mode, tps), s1, s2


def stitch_stream_host_async(spatial_net, temporal_net, smooth_net, slot, lr1, lr2, hr1, hr2, out, mode="NORMAL",
                             tps=None):
    """Pipelined form of stitch_stream_host (ss2_stitch_stream_host_async): returns (Ho, Wo) once the
    chunk's resample+blend and D2H copies are enqueued; `stitch_stream_host_wait(slot)` completes it.
    Two slots (0, 1) overlap the D2H of one chunk with the H2D + networks of the next."""
    from .utils import torch_tps_transform as tt
    ctx = _lib.context()
    _sync_nets(ctx, spatial_net, temporal_net, smooth_net)
    for t in (lr1, lr2, hr1, hr2, out):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("stitch_stream_host takes contiguous fp32 HOST tensors")
    n, _, H, W = hr1.shape
    ho, wo = ctypes.c_int(), ctypes.c_int()
    ctx.check(ctx.lib.ss2_stitch_stream_host_async(ctx.handle, int(slot), _lib.ptr(lr1), _lib.ptr(lr2), _lib.ptr(hr1),
                                                   _lib.ptr(hr2), n, H, W, _lib.MODE[mode],
                                                   tt.DEFAULT_TPS if tps is None else tps, _lib.ptr(out), out.numel(),
                                                   ctypes.byref(ho), ctypes.byref(wo), None, None))
    return ho.value, wo.value


def stitch_stream_host_prefetch(slot, lr1, lr2, hr1, hr2):
    """ss2_stitch_stream_host_prefetch: start the H2D copies of the chunk the next stitch_stream_host_async on `slot`
    will process (same host tensors); returns immediately."""
    ctx = _lib.context()
    for t in (lr1, lr2, hr1, hr2):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("stitch_stream_host takes contiguous fp32 HOST tensors")
    n, _, H, W = hr1.shape
    ctx.check(ctx.lib.ss2_stitch_stream_host_prefetch(ctx.handle, int(slot), _lib.ptr(lr1), _lib.ptr(lr2), _lib.ptr(hr1),
                                                      _lib.ptr(hr2), n, H, W))


def stitch_stream_host_wait(slot):
    ctx = _lib.context()
    ctx.check(ctx.lib.ss2_stitch_stream_host_wait(ctx.handle, int(slot)))


def stitch_stream_host(spatial_net, temporal_net, smooth_net, lr1, lr2, hr1, hr2, out, mode="NORMAL", tps=None,
                       want_meshes=False):
    """ss2_stitch_stream_host: all inputs/outputs are HOST tensors (pin them for full PCIe
    speed); `out` is a flat fp32 host buffer large enough for n*3*Ho*Wo.  Returns (Ho, Wo)
    [, smooth_mesh1, smooth_mesh2 as host tensors]."""
    from .utils import torch_tps_transform as tt
    ctx = _lib.context()
    _sync_nets(ctx, spatial_net, temporal_net, smooth_net)
    for t in (lr1, lr2, hr1, hr2, out):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("stitch_stream_host takes contiguous fp32 HOST tensors")
    n, _, H, W = hr1.shape
    ho, wo = ctypes.c_int(), ctypes.c_int()
    m1 = torch.empty(n, 7, 9, 2) if want_meshes else None
    m2 = torch.empty(n, 7, 9, 2) if want_meshes else None
    ctx.check(ctx.lib.ss2_stitch_stream_host(ctx.handle, _lib.ptr(lr1), _lib.ptr(lr2), _lib.ptr(hr1), _lib.ptr(hr2),
                                             n, H, W, _lib.MODE[mode], tt.DEFAULT_TPS if tps is None else tps,
                                             _lib.ptr(out), out.numel(), ctypes.byref(ho), ctypes.byref(wo),
                                             _lib.ptr(m1), _lib.ptr(m2)))
    if want_meshes:
        return ho.value, wo.value, m1, m2
    return ho.value, wo.value


# ------------------------------------------------------------------------------------------
# uint8 host edges (test_online_tra.py:252-264 before the networks, :152,414 before the video writer)
# ------------------------------------------------------------------------------------------
def load_frames_u8(frames_u8, want_hr=True, want_lr=True):
    """frames [n,H,W,3] uint8 (cv2.imread layout, BGR; CUDA or host) -> (hr [n,3,H,W] fp32 0..255,
    lr [n,3,360,480] fp32 in [-1,1] = cv2.resize(INTER_LINEAR)/127.5-1, bit-exact), on the device."""
    ctx = _lib.context()
    u = torch.as_tensor(frames_u8)
    if u.dtype != torch.uint8 or u.dim() != 4 or u.shape[3] != 3:
        raise ValueError("frames must be uint8 [n,H,W,3]")
    u = u.cuda().contiguous()
    n, H, W, _ = u.shape
    hr = torch.empty(n, 3, H, W, device=u.device, dtype=torch.float32) if want_hr else None
    lr = torch.empty(n, 3, 360, 480, device=u.device, dtype=torch.float32) if want_lr else None
    ctx.check(ctx.lib.ss2_load_frames_u8(ctx.handle, _lib.ptr(u), n, H, W, _lib.ptr(hr), _lib.ptr(lr), _lib.cur_stream()))
    return hr, lr


def frames_to_u8(fused):
    """fused [n,3,Ho,Wo] fp32 CUDA -> [n,Ho,Wo,3] uint8 CUDA (transpose + astype(uint8) of test_online_tra.py:152,414)."""
    ctx = _lib.context()
    f = _lib.dev_f32(fused)
    n, C, Ho, Wo = f.shape
    if C != 3:
        raise ValueError("fused frames must be [n,3,Ho,Wo]")
    out = torch.empty(n, Ho, Wo, 3, device=f.device, dtype=torch.uint8)
    ctx.check(ctx.lib.ss2_frames_to_u8(ctx.handle, _lib.ptr(f), n, Ho, Wo, _lib.ptr(out), _lib.cur_stream()))
    return out


def _check_u8_host(*ts):
    for t in ts:
        if t.is_cuda or t.dtype != torch.uint8 or not t.is_contiguous():
            raise ValueError("the uint8 host interface takes contiguous uint8 HOST tensors")


def stitch_stream_host_u8_async(spatial_net, temporal_net, smooth_net, slot, bgr1, bgr2, out, mode="NORMAL", tps=None):
    """ss2_stitch_stream_host_u8_async: bgr1, bgr2 [n,H,W,3] uint8 host frames (as cv2.imread returns them), `out` a
    flat uint8 host buffer for n*Ho*Wo*3 bytes; returns (Ho, Wo) once the chunk's work and D2H copies are enqueued."""
    from .utils import torch_tps_transform as tt
    ctx = _lib.context()
    _sync_nets(ctx, spatial_net, temporal_net, smooth_net)
    _check_u8_host(bgr1, bgr2, out)
    n, H, W, _ = bgr1.shape
    ho, wo = ctypes.c_int(), ctypes.c_int()
    ctx.check(ctx.lib.ss2_stitch_stream_host_u8_async(ctx.handle, int(slot), _lib.ptr(bgr1), _lib.ptr(bgr2), n, H, W,
                                                      _lib.MODE[mode], tt.DEFAULT_TPS if tps is None else tps,
                                                      _lib.ptr(out), out.numel(), ctypes.byref(ho), ctypes.byref(wo),
                                                      None, None))
    return ho.value, wo.value


def stitch_stream_host_submit(spatial_net, temporal_net, smooth_net, slot, *ins):
    """First half of the pipelined host call: enqueue a chunk's uploads, front end and networks; returns at once.
    ins = (bgr1, bgr2) uint8 [n,H,W,3] or (lr1, lr2, hr1, hr2) fp32 host tensors."""
    ctx = _lib.context()
    _sync_nets(ctx, spatial_net, temporal_net, smooth_net)
    if len(ins) == 2:
        _check_u8_host(*ins)
        n, H, W, _ = ins[0].shape
        ctx.check(ctx.lib.ss2_stitch_stream_host_u8_submit(ctx.handle, int(slot), _lib.ptr(ins[0]), _lib.ptr(ins[1]), n, H, W))
    else:
        lr1, lr2, hr1, hr2 = ins
        for t in ins:
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("stitch_stream_host takes contiguous fp32 HOST tensors")
        n, _, H, W = hr1.shape
        ctx.check(ctx.lib.ss2_stitch_stream_host_submit(ctx.handle, int(slot), _lib.ptr(lr1), _lib.ptr(lr2), _lib.ptr(hr1),
                                                        _lib.ptr(hr2), n, H, W))


def stitch_stream_host_finish(slot, out, mode="NORMAL", tps=None):
    """Second half: wait for the submitted chunk's canvas, enqueue resample + blend and the frame downloads into `out`
    (uint8 after a uint8 submit, fp32 otherwise); returns (Ho, Wo).  stitch_stream_host_wait(slot) completes it."""
    from .utils import torch_tps_transform as tt
    ctx = _lib.context()
    if out.is_cuda or not out.is_contiguous():
        raise ValueError("`out` must be a contiguous HOST tensor")
    ho, wo = ctypes.c_int(), ctypes.c_int()
    ctx.check(ctx.lib.ss2_stitch_stream_host_finish(ctx.handle, int(slot), _lib.MODE[mode], tt.DEFAULT_TPS if tps is None else tps,
                                                    _lib.ptr(out), out.numel(), ctypes.byref(ho), ctypes.byref(wo), None, None))
    return ho.value, wo.value


def stitch_stream_host_u8_prefetch(slot, bgr1, bgr2):
    ctx = _lib.context()
    _check_u8_host(bgr1, bgr2)
    n, H, W, _ = bgr1.shape
    ctx.check(ctx.lib.ss2_stitch_stream_host_u8_prefetch(ctx.handle, int(slot), _lib.ptr(bgr1), _lib.ptr(bgr2), n, H, W))


def stitch_stream_host_u8(spatial_net, temporal_net, smooth_net, bgr1, bgr2, out, mode="NORMAL", tps=None,
                          want_meshes=False):
    """Blocking form: uint8 frames in, uint8 stitched frames out ([n,Ho,Wo,3] inside `out`); returns (Ho, Wo)
    [, smooth_mesh1, smooth_mesh2]."""
    from .utils import torch_tps_transform as tt
    ctx = _lib.context()
    _sync_nets(ctx, spatial_net, temporal_net, smooth_net)
    _check_u8_host(bgr1, bgr2, out)
    n, H, W, _ = bgr1.shape
    ho, wo = ctypes.c_int(), ctypes.c_int()
    m1 = torch.empty(n, 7, 9, 2) if want_meshes else None
    m2 = torch.empty(n, 7, 9, 2) if want_meshes else None
    ctx.check(ctx.lib.ss2_stitch_stream_host_u8(ctx.handle, _lib.ptr(bgr1), _lib.ptr(bgr2), n, H, W, _lib.MODE[mode],
                                                tt.DEFAULT_TPS if tps is None else tps, _lib.ptr(out), out.numel(),
                                                ctypes.byref(ho), ctypes.byref(wo), _lib.ptr(m1), _lib.ptr(m2)))
    if want_meshes:
        return ho.value, wo.value, m1, m2
    return ho.value, wo.value


# ------------------------------------------------------------------------------------------
# temporally sharded stream: one process per GPU, rank r owns frames [r*F, (r+1)*F)
# (new design - the reference has no multi-GPU path)
# ------------------------------------------------------------------------------------------
def shard_plan(rank, world, frames_per_rank):
    """Index arithmetic of the temporal sharding.  Frame k's smoothed mesh comes from the
    window k-6..k (test_online_tra.py:359-392), whose tsmotion needs the raw spatial mesh of
    frame k-7+1.. and the temporal motion of its own frames (test_online_tra.py:324-340):
      start, stop   this rank's frames
      ctx0          first frame whose raw (smotion, tmotion) this rank needs: max(0, start-6)
      prev          frame whose smotion seeds tsmotion[ctx0] (-1: ctx0 is the stream start)
      nwin          SmoothNet windows this rank evaluates
      with_head     rank 0 keeps all 7 meshes of window 0
      input_halo    extra leading input frames TemporalNet needs (frame start-1)"""
    F = int(frames_per_rank)
    if F < WINDOW:
        raise ValueError("a shard needs at least %d frames (got %d)" % (WINDOW, F))
    start, stop = rank * F, (rank + 1) * F
    ctx0 = max(0, start - (WINDOW - 1))
    with_head = ctx0 == 0
    nwin = (stop - ctx0) - (WINDOW - 1)
    if not with_head:
        assert nwin == F
    return dict(start=start, stop=stop, ctx0=ctx0, prev=ctx0 - 1, nwin=nwin, with_head=with_head,
                input_halo=1 if rank > 0 else 0, total=world * F)


def exchange_raw_meshes(raw, group=None):
    """All-gather of the per-frame raw motions: raw [4,F,7,9,2] (smotion1, smotion2, tmotion1,
    tmotion2 of this rank's frames) -> [4, world*F, 7,9,2] in stream order.  ~4 KB per frame:
    latency-bound on NVLink, so one collective for everything."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [torch.empty_like(raw) for _ in range(world)]
    dist.all_gather(parts, raw.contiguous(), group=group)
    return torch.cat(parts, 1)


def allreduce_canvas(minmax, group=None):
    """Global canvas (test_online_tra.py:106-117 over ALL frames): (xmin,xmax,ymin,ymax) per
    rank -> global, as ONE min-all-reduce of (xmin,-xmax,ymin,-ymax) (negation is exact)."""
    import torch.distributed as dist
    sign = torch.tensor([1.0, -1.0, 1.0, -1.0], device=minmax.device, dtype=minmax.dtype)
    v = minmax * sign
    dist.all_reduce(v, op=dist.ReduceOp.MIN, group=group)
    return v * sign


def build_spatial_temporal(spatial_net, temporal_net, lr1, lr2, halo=0):
    """build_SpatialNet over the last n frame pairs and build_TemporalNet over both views of lr1, lr2 [halo+n,3,360,480] in one
    library call (TemporalNet on a stream of its own next to SpatialNet; bit-identical to the separate calls).
    Returns (motion1, motion2 [n,7,9,2], tmotion1, tmotion2 [halo+n,7,9,2])."""
    ctx = _lib.context()
    spatial_net.sync_weights(ctx)
    temporal_net.sync_weights(ctx)
    lr1, lr2 = _lib.dev_f32(lr1), _lib.dev_f32(lr2)
    n = lr1.shape[0] - int(halo)
    dev = lr1.device
    sm = [torch.empty(n, 7, 9, 2, device=dev, dtype=torch.float32) for _ in range(2)]
    tm = [torch.empty(n + int(halo), 7, 9, 2, device=dev, dtype=torch.float32) for _ in range(2)]
    ctx.check(ctx.lib.ss2_build_spatial_temporal(ctx.handle, _lib.ptr(lr1), _lib.ptr(lr2), n, int(halo), _lib.ptr(sm[0]),
                                                 _lib.ptr(sm[1]), _lib.ptr(tm[0]), _lib.ptr(tm[1]), _lib.cur_stream()))
    return sm[0], sm[1], tm[0], tm[1]


def stream_meshes_sharded(spatial_net, temporal_net, smooth_net, lr1, lr2, input_halo, frames, group=None):
    """Smooth meshes of this rank's `frames` frames of a temporally sharded stream.  lr1, lr2
    [input_halo+frames,3,360,480] (rank > 0 also gets frame start-1 for TemporalNet).  One all-gather of the raw
    per-frame motions (4 KB per frame) rebuilds the 6-frame SmoothNet context locally.  Returns (S1, S2 [frames,7,9,2])
    - identical to rows [start, stop) of the single-process result."""
    import torch.distributed as dist
    ctx = _lib.context()
    _sync_nets(ctx, spatial_net, temporal_net, smooth_net)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    F = int(frames)
    plan = shard_plan(rank, world, F)
    if input_halo != plan["input_halo"] or lr1.shape[0] != F + input_halo:
        raise ValueError("rank %d expects %d halo frame(s) in lr1/lr2" % (rank, plan["input_halo"]))
    lr1, lr2 = _lib.dev_f32(lr1), _lib.dev_f32(lr2)
    dev = lr1.device
    st = _lib.cur_stream()
    raw = torch.empty(4, F, 7, 9, 2, device=dev, dtype=torch.float32)
    tms = [torch.empty(F + input_halo, 7, 9, 2, device=dev, dtype=torch.float32) for _ in range(2)]
    # both networks in one call: TemporalNet (halo frame included) runs next to SpatialNet on a stream of its own
    ctx.check(ctx.lib.ss2_build_spatial_temporal(ctx.handle, _lib.ptr(lr1), _lib.ptr(lr2), F, input_halo, _lib.ptr(raw[0]),
                                                 _lib.ptr(raw[1]), _lib.ptr(tms[0]), _lib.ptr(tms[1]), st))
    raw[2], raw[3] = tms[0][input_halo:], tms[1][input_halo:]
    allraw = exchange_raw_meshes(raw, group)  # [4, world*F, 7,9,2]
    c0, stop = plan["ctx0"], plan["stop"]
    n = stop - c0
    smesh, tsm = [], []
    for v in range(2):
        sm = allraw[v, c0:stop].contiguous()
        tmo = allraw[2 + v, c0:stop].contiguous()
        prev = allraw[v, plan["prev"]].contiguous() if plan["prev"] >= 0 else None
        mesh = torch.empty(n, 7, 9, 2, device=dev, dtype=torch.float32)
        ts = torch.empty_like(mesh)
        ctx.check(ctx.lib.ss2_tsmotion(ctx.handle, _lib.ptr(sm), _lib.ptr(tmo), n, 1 if prev is None else 0,
                                       _lib.ptr(prev), _lib.ptr(mesh), _lib.ptr(ts), st))
        smesh.append(mesh)
        tsm.append(ts)
    win = smooth_windows(smooth_net, tsm[0], tsm[1], smesh[0], smesh[1], plan["nwin"],
                         want=("smooth_mesh1", "smooth_mesh2"))
    S = []
    for key in ("smooth_mesh1", "smooth_mesh2"):
        o = torch.empty(F, 7, 9, 2, device=dev, dtype=torch.float32)
        ctx.check(ctx.lib.ss2_assemble_smooth(ctx.handle, _lib.ptr(win[key]), plan["nwin"],
                                              1 if plan["with_head"] else 0, _lib.ptr(o), st))
        S.append(o)
    return S[0], S[1]


def stitch_stream_sharded(spatial_net, temporal_net, smooth_net, lr1, lr2, hr1, hr2, input_halo, mode="NORMAL",
                          tps=None, group=None):
    """This rank's part of a temporally sharded stream.  lr1, lr2 [input_halo+F,3,360,480]
    (rank > 0 also gets frame start-1 for TemporalNet), hr1, hr2 [F,3,H,W]; all CUDA.
    Returns (fused [F,3,Ho,Wo], smooth_mesh1 [F,7,9,2], smooth_mesh2) - identical to the rows
    [start, stop) of the single-process result."""
    S1, S2 = stream_meshes_sharded(spatial_net, temporal_net, smooth_net, lr1, lr2, input_halo, hr1.shape[0], group)
    _, _, H, W = hr1.shape
    mm = allreduce_canvas(canvas_minmax(S1, S2, H, W), group).cpu().tolist()
    return stable_frames(hr1, hr2, S1, S2, mm, mode, tps), S1, S2


def stitch_nview_stream(spatial_net, temporal_net, smooth_net, lrs, hrs, input_halo=0, mode="NORMAL", group=None):
    """BASELINE.json config 5: N views (2 <= N <= 4), N-1 stitched pairs (v, v+1), middle-plane chain, one fused N-image
    resample + AVERAGE blend.  lrs / hrs: lists of N CUDA tensors [input_halo+F,3,360,480] / [F,3,H,W].  Single process,
    or one temporal shard per rank when torch.distributed is initialised with more than one rank (mesh halo all-gather
    per pair, the two canvases all-reduced).  Returns (fused [F,3,Ho,Wo], meshes [N,F,7,9,2])."""
    import torch.distributed as dist
    sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    F = hrs[0].shape[0]
    pairs = []
    for v in range(len(lrs) - 1):
        if sharded:
            pairs.append(stream_meshes_sharded(spatial_net, temporal_net, smooth_net, lrs[v], lrs[v + 1], input_halo, F, group))
        else:
            pairs.append(stream_meshes(spatial_net, temporal_net, smooth_net, lrs[v], lrs[v + 1]))
    fused, meshes, _ = nview_stable(hrs, pairs, warp_mode=mode, group=group)
    return fused, meshes
