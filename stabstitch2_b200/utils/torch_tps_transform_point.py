"""Drop-in for Full_model_inference/Codes/utils/torch_tps_transform_point.py (transformer, :6)."""
import torch

from .. import _lib


def transformer(point, source, target):
    ctx = _lib.context()
    point, source, target = _lib.dev_f32(point), _lib.dev_f32(source), _lib.dev_f32(target)
    bn = point.shape[0]
    if point.shape != (bn, 63, 2) or source.shape != (bn, 63, 2) or target.shape != (bn, 63, 2):
        raise ValueError("point/source/target must be [bn,63,2]")
    out = torch.empty_like(point)
    ctx.check(ctx.lib.ss2_tps_point(ctx.handle, _lib.ptr(point), _lib.ptr(source), _lib.ptr(target), bn,
                                    _lib.ptr(out), _lib.cur_stream()))
    return out
