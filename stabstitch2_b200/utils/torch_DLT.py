"""Drop-in for Full_model_inference/Codes/utils/torch_DLT.py (tensor_DLT, :17-45)."""
import torch

from .. import _lib


def tensor_DLT(src_p, dst_p):
    ctx = _lib.context()
    s, d = _lib.dev_f32(src_p), _lib.dev_f32(dst_p)
    bs = s.shape[0]
    H = torch.empty(bs, 3, 3, device=s.device, dtype=torch.float32)
    ctx.check(ctx.lib.ss2_dlt(ctx.handle, _lib.ptr(s), _lib.ptr(d), bs, _lib.ptr(H), _lib.cur_stream()))
    return H
