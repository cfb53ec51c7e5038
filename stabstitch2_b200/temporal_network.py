"""Drop-in for Full_model_inference/Codes/temporal_network.py: TemporalNet, build_TemporalNet."""
import torch

from . import _lib, grid_res
from ._modules import NativeNet, regress_convs, regress_fc, resnet18_feature_extractors

grid_h = grid_res.GRID_H
grid_w = grid_res.GRID_W


def _stack_frames(img_tensor_list):
    frames = [_lib.dev_f32(t) for t in img_tensor_list]
    bs = frames[0].shape[0]
    if bs != 1:
        raise ValueError("the reference drives TemporalNet with batch 1 (test_online_tra.py:256,263)")
    x = torch.cat(frames, 0)
    if tuple(x.shape[1:]) != (3, 360, 480):
        raise ValueError("TemporalNet runs at [1,3,360,480] frames")
    return x


def build_TemporalNet(net, img_tensor_list):
    """temporal_network.py:23-34: list of N frames (CPU or CUDA) -> {'motion_list': N x [1,7,9,2]},
    element 0 zeros, element k the mesh motion of frame k w.r.t. frame k-1."""
    ctx = _lib.context()
    net.sync_weights(ctx)
    x = _stack_frames(img_tensor_list)
    n = x.shape[0]
    out = torch.empty(n, grid_h + 1, grid_w + 1, 2, device=x.device, dtype=torch.float32)
    ctx.check(ctx.lib.ss2_build_temporal(ctx.handle, _lib.ptr(x), n, _lib.ptr(out), _lib.cur_stream()))
    return dict(motion_list=[out[k:k + 1] for k in range(n)])


class TemporalNet(NativeNet):
    NET_ID = _lib.NET_TEMPORAL

    def __init__(self, dropout=0.):
        super().__init__()
        self.regressNet2_part1 = regress_convs(49, (64, 64, 128, 128, 128, 128, 256, 256))
        self.regressNet2_part2 = regress_fc(1536, 1024, 512, (grid_w + 1) * (grid_h + 1) * 2)
        self.init_reference_style()
        self.feature_extractor_stage1, self.feature_extractor_stage2 = resnet18_feature_extractors()

    def forward(self, img_tensor_list):
        """temporal_network.py:120-147 -> list of N-1 mesh motions [1,7,9,2]."""
        return build_TemporalNet(self, img_tensor_list)["motion_list"][1:]
