"""Drop-in for Full_model_inference/Codes/utils/torch_tps_transform.py (transformer, :7).

`out_size` elements may be python ints or 0-dim (CUDA) int tensors, as the reference's
get_stable_sqe passes them (test_online_tra.py:140)."""
import torch

from .. import _lib

# evaluation of the dense TPS field: _lib.TPS_EXACT (all 63 radial terms per pixel, the
# reference's arithmetic) or _lib.TPS_LATTICE (far field interpolated from a lattice + exact
# near-field corrections, as accurate against the fp64 arbiter as the reference itself; used
# for 3-channel images on canvases large enough for it, EXACT otherwise)
DEFAULT_TPS = _lib.TPS_LATTICE


def transformer(U, source, target, out_size, mode="NORMAL", tps=None):
    if mode not in _lib.MODE:
        raise ValueError("mode must be 'NORMAL' or 'FAST'")
    ctx = _lib.context()
    U, source, target = _lib.dev_f32(U), _lib.dev_f32(source), _lib.dev_f32(target)
    bn, C, H, W = U.shape
    if source.shape != (bn, 63, 2) or target.shape != (bn, 63, 2):
        raise ValueError("source/target must be [bn,63,2]")
    Ho, Wo = int(out_size[0]), int(out_size[1])
    out = torch.empty(bn, C, Ho, Wo, device=U.device, dtype=torch.float32)
    ctx.check(ctx.lib.ss2_tps_warp(ctx.handle, _lib.ptr(U), _lib.ptr(source), _lib.ptr(target), bn, C, H, W,
                                   Ho, Wo, _lib.MODE[mode], DEFAULT_TPS if tps is None else tps,
                                   _lib.ptr(out), _lib.cur_stream()))
    return out


def warp_blend_average(img1, img2, source, target, out_size, mode="NORMAL", tps=None):
    """Fused form of test_online_tra.py:140-142: img1,img2 [n,3,H,W]; source,target [n,2,63,2]
    -> fused [n,3,Ho,Wo]."""
    ctx = _lib.context()
    img1, img2 = _lib.dev_f32(img1), _lib.dev_f32(img2)
    source, target = _lib.dev_f32(source), _lib.dev_f32(target)
    n, C, H, W = img1.shape
    if C != 3 or img2.shape != img1.shape:
        raise ValueError("img1/img2 must be [n,3,H,W] of equal shape")
    if source.shape != (n, 2, 63, 2) or target.shape != (n, 2, 63, 2):
        raise ValueError("source/target must be [n,2,63,2]")
    Ho, Wo = int(out_size[0]), int(out_size[1])
    out = torch.empty(n, 3, Ho, Wo, device=img1.device, dtype=torch.float32)
    ctx.check(ctx.lib.ss2_tps_warp_blend_avg(ctx.handle, _lib.ptr(img1), _lib.ptr(img2), _lib.ptr(source),
                                             _lib.ptr(target), n, H, W, Ho, Wo, _lib.MODE[mode],
                                             DEFAULT_TPS if tps is None else tps, _lib.ptr(out),
                                             _lib.cur_stream()))
    return out
