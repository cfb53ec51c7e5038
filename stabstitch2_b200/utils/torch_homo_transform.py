"""Drop-in for Full_model_inference/Codes/utils/torch_homo_transform.py (transformer, :6)."""
import torch

from .. import _lib


def transformer(U, theta, out_size, **kwargs):
    ctx = _lib.context()
    U, theta = _lib.dev_f32(U), _lib.dev_f32(theta).reshape(-1, 3, 3)
    bn, C, H, W = U.shape
    Ho, Wo = int(out_size[0]), int(out_size[1])
    out = torch.empty(bn, C, Ho, Wo, device=U.device, dtype=torch.float32)
    ctx.check(ctx.lib.ss2_homo_warp(ctx.handle, _lib.ptr(U), _lib.ptr(theta), bn, C, H, W, Ho, Wo,
                                    _lib.ptr(out), _lib.cur_stream()))
    return out
