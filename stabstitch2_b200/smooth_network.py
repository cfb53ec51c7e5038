"""Drop-in for Full_model_inference/Codes/smooth_network.py: SmoothNet, build_SmoothNet."""
import torch
import torch.nn as nn

from . import _lib, grid_res
from ._modules import NativeNet

grid_h = grid_res.GRID_H
grid_w = grid_res.GRID_W

KEYS = ("ori_path1", "smooth_path1", "ori_mesh1", "smooth_mesh1",
        "ori_path2", "smooth_path2", "ori_mesh2", "smooth_mesh2")


class MotionPrediction(nn.Module):
    """Parameter container, smooth_network.py:106-136 (embedding2 exists but is unused)."""

    def __init__(self, kernel=5):
        super().__init__()
        self.embedding1 = nn.Sequential(nn.Linear(2, 32), nn.ReLU())
        self.embedding2 = nn.Sequential(nn.Linear(1, 8), nn.ReLU())
        self.embedding3 = nn.Sequential(nn.Linear(2, 32), nn.ReLU())
        self.pad = kernel // 2
        self.MotionConv3D = nn.Sequential(
            nn.Conv3d(128, 128, (kernel, 3, 3), padding=(self.pad, 1, 1)), nn.ReLU(),
            nn.Conv3d(128, 128, (kernel, 3, 3), padding=(self.pad, 1, 1)), nn.ReLU(),
            nn.Conv3d(128, 128, (kernel, 3, 3), padding=(self.pad, 1, 1)), nn.ReLU())
        self.decoding = nn.Sequential(nn.Linear(128, 4))


def smooth_windows(net, tsmotion1, tsmotion2, smesh1, smesh2, nwin, want=KEYS, zero_first=True):
    """nwin sliding 7-frame windows in one call: inputs [nwin+6,7,9,2] frame-major CUDA tensors
    (tsmotion of each window's first frame is zeroed inside, test_online_tra.py:361-365).
    Returns {key: [nwin,7,7,9,2]} for the requested keys."""
    ctx = _lib.context()
    net.sync_weights(ctx)
    ins = [_lib.dev_f32(t).reshape(-1, grid_h + 1, grid_w + 1, 2) for t in (tsmotion1, tsmotion2, smesh1, smesh2)]
    for t in ins:
        if t.shape[0] != nwin + 6:
            raise ValueError("need nwin+6 frames of [7,9,2] meshes, got %s" % (tuple(t.shape),))
    outs = {k: (torch.empty(nwin, 7, grid_h + 1, grid_w + 1, 2, device=ins[0].device, dtype=torch.float32)
                if k in want else None) for k in KEYS}
    ctx.check(ctx.lib.ss2_build_smooth(ctx.handle, *[_lib.ptr(t) for t in ins], nwin, 1 if zero_first else 0,
                                       *[_lib.ptr(outs[k]) for k in KEYS], _lib.cur_stream()))
    return {k: v for k, v in outs.items() if v is not None}


def build_SmoothNet(net, tsmotion_list1, tsmotion_list2, smesh_list1, smesh_list2):
    """smooth_network.py:23-41: four lists of 7 tensors [1,7,9,2] -> the 8-key dict, each
    [1,7,7,9,2].  Like the reference, this does NOT zero tsmotion_list*[0]: its driver does
    that before the call (test_online_tra.py:361-365)."""
    if not (len(tsmotion_list1) == len(tsmotion_list2) == len(smesh_list1) == len(smesh_list2) == 7):
        raise ValueError("SmoothNet windows are 7 frames long (test_online_tra.py:219)")
    cat = lambda lst: torch.cat([_lib.dev_f32(t) for t in lst], 0)  # noqa: E731
    return smooth_windows(net, cat(tsmotion_list1), cat(tsmotion_list2), cat(smesh_list1), cat(smesh_list2), 1,
                          zero_first=False)


class SmoothNet(NativeNet):
    NET_ID = _lib.NET_SMOOTH

    def __init__(self, dropout=0.):
        super().__init__()
        self.MotionPre = MotionPrediction()
        self.init_reference_style()

    def forward(self, smesh_list1, smesh_list2, tsmotion_list1, tsmotion_list2):
        """smooth_network.py:64-101 -> (smesh1, smesh2, tsflow1, tsflow2, delta1, delta2)."""
        o = build_SmoothNet(self, tsmotion_list1, tsmotion_list2, smesh_list1, smesh_list2)
        return (o["ori_mesh1"], o["ori_mesh2"], o["ori_path1"], o["ori_path2"],
                o["smooth_path1"] - o["ori_path1"], o["smooth_path2"] - o["ori_path2"])
