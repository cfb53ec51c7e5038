"""Drop-in for Full_model_inference/Codes/spatial_network.py: SpatialNet, build_SpatialNet,
H2Mesh, get_rigid_mesh, get_norm_mesh.  forward/build run entirely in libss2 (CUDA)."""
import torch
import torch.nn as nn

from . import _lib, grid_res
from ._modules import NativeNet, regress_convs, regress_fc, resnet18_feature_extractors
from .utils import torch_DLT

grid_h = grid_res.GRID_H
grid_w = grid_res.GRID_W


def get_rigid_mesh(batch_size, height, width):
    """spatial_network.py:39-50 (plain index arithmetic, host-side helper)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu"
    xs = torch.linspace(0.0, float(width), grid_w + 1)
    ys = torch.linspace(0.0, float(height), grid_h + 1)
    m = torch.stack([xs[None, :].expand(grid_h + 1, -1), ys[:, None].expand(-1, grid_w + 1)], 2)
    return m.to(dev).unsqueeze(0).expand(batch_size, -1, -1, -1)


def get_norm_mesh(mesh, height, width):
    """spatial_network.py:53-59."""
    bs = mesh.size()[0]
    mw = mesh[..., 0] * 2.0 / float(width) - 1.0
    mh = mesh[..., 1] * 2.0 / float(height) - 1.0
    return torch.stack([mw, mh], 3).reshape([bs, -1, 2])


def H2Mesh(H, rigid_mesh):
    """spatial_network.py:20-36: apply H^-1 to the rigid vertices (tiny host-side helper kept
    for API completeness; the hot path uses ss2_spatial_tail)."""
    bs = rigid_mesh.size()[0]
    Hinv = torch.inverse(H)
    pts = torch.cat([rigid_mesh.reshape(bs, -1, 2), torch.ones(bs, (grid_h + 1) * (grid_w + 1), 1, device=H.device)], 2)
    q = torch.matmul(Hinv, pts.permute(0, 2, 1))
    return torch.stack([q[:, 0] / q[:, 2], q[:, 1] / q[:, 2]], 2).reshape(bs, grid_h + 1, grid_w + 1, 2)


def build_SpatialNet(net, input1_tensor, input2_tensor):
    """spatial_network.py:63-118 -> {'motion1','motion2'}: [bs,7,9,2] CUDA fp32."""
    ctx = _lib.context()
    net.sync_weights(ctx)
    a, b = _lib.dev_f32(input1_tensor), _lib.dev_f32(input2_tensor)
    bs, c, h, w = a.shape
    if (c, h, w) != (3, 360, 480) or b.shape != a.shape:
        raise ValueError("SpatialNet runs at [bs,3,360,480] (test_online_tra.py:247-248)")
    m1 = torch.empty(bs, grid_h + 1, grid_w + 1, 2, device=a.device, dtype=torch.float32)
    m2 = torch.empty_like(m1)
    ctx.check(ctx.lib.ss2_build_spatial(ctx.handle, _lib.ptr(a), _lib.ptr(b), bs, _lib.ptr(m1), _lib.ptr(m2),
                                        _lib.cur_stream()))
    return dict(motion1=m1, motion2=m2)


class SpatialNet(NativeNet):
    NET_ID = _lib.NET_SPATIAL

    def __init__(self):
        super().__init__()
        self.regressNet1_part1 = regress_convs(2, (64, 64, 128, 128, 128, 128))
        self.regressNet1_part2 = regress_fc(768, 512, 128, 8)
        self.regressNet2_part1_ref = regress_convs(121, (64, 64, 128, 128, 128, 128, 256, 256))
        self.regressNet2_part2_ref = regress_fc(1536, 1024, 512, (grid_w + 1) * (grid_h + 1) * 2)
        self.regressNet2_part1_tgt = regress_convs(121, (64, 64, 128, 128, 128, 128, 256, 256))
        self.regressNet2_part2_tgt = regress_fc(1536, 1024, 512, (grid_w + 1) * (grid_h + 1) * 2)
        self.init_reference_style()
        # the reference downloads ImageNet weights here (spatial_network.py:268); offline we
        # start from random init - real checkpoints overwrite everything via load_state_dict
        self.feature_extractor_stage1, self.feature_extractor_stage2 = resnet18_feature_extractors()

    @staticmethod
    def cost_volume(x1, x2, search_range, norm=True, fast=True):
        """spatial_network.py:333-358 with NCHW tensors [bs,128,h,w] -> [bs,(2sr+1)^2,h,w].
        The inference path always calls it with norm=False; norm=True is not implemented."""
        if norm:
            raise NotImplementedError("cost_volume(norm=True) is never used on the inference path")
        ctx = _lib.context()
        a = _lib.dev_f32(x1).permute(0, 2, 3, 1).contiguous()
        b = _lib.dev_f32(x2).permute(0, 2, 3, 1).contiguous()
        bs, h, w, c = a.shape
        nd = (2 * search_range + 1) ** 2
        cp = (nd + 31) // 32 * 32
        out = torch.empty(bs, h, w, cp, device=a.device, dtype=torch.float32)
        ctx.check(ctx.lib.ss2_cost_volume_nhwc(ctx.handle, _lib.ptr(a), _lib.ptr(b), bs, h, w, c, search_range, cp,
                                               _lib.ptr(out), _lib.cur_stream()))
        return out[..., :nd].permute(0, 3, 1, 2).contiguous()

    def CCL(self, feature_1, feature_2):
        """spatial_network.py:369-425 with NCHW tensors [bs,c,h,w] -> [bs,2,h,w] (flow_w, flow_h)."""
        ctx = _lib.context()
        a = _lib.dev_f32(feature_1).permute(0, 2, 3, 1).contiguous()
        b = _lib.dev_f32(feature_2).permute(0, 2, 3, 1).contiguous()
        bs, h, w, c = a.shape
        out = torch.empty(bs, h, w, 4, device=a.device, dtype=torch.float32)
        ctx.check(ctx.lib.ss2_ccl_nhwc(ctx.handle, _lib.ptr(a), _lib.ptr(b), bs, h, w, c, _lib.ptr(out),
                                       _lib.cur_stream()))
        return out[..., :2].permute(0, 3, 1, 2).contiguous()

    def forward(self, input1_tesnor, input2_tesnor):
        """spatial_network.py:276-331 -> (offset_1 [bs,8], offset_2_ref [bs,126], offset_2_tgt [bs,126])."""
        ctx = _lib.context()
        self.sync_weights(ctx)
        a, b = _lib.dev_f32(input1_tesnor), _lib.dev_f32(input2_tesnor)
        bs, c, h, w = a.shape
        if (c, h, w) != (3, 360, 480) or b.shape != a.shape:
            raise ValueError("SpatialNet runs at [bs,3,360,480]")
        o1 = torch.empty(bs, 8, device=a.device, dtype=torch.float32)
        oref = torch.empty(bs, 126, device=a.device, dtype=torch.float32)
        otgt = torch.empty_like(oref)
        ctx.check(ctx.lib.ss2_spatial_forward(ctx.handle, _lib.ptr(a), _lib.ptr(b), bs, _lib.ptr(o1), _lib.ptr(oref),
                                              _lib.ptr(otgt), _lib.cur_stream()))
        return o1, oref, otgt
