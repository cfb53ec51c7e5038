"""ctypes binding of libss2.so (include/ss2.h).  There is NO fallback: if the CUDA library
is missing or a call fails, this raises."""
import ctypes
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# SS2_LIB: another build of the same library (kernel parameter sweeps under profiles/); default the in-tree build
LIB_PATH = os.environ.get("SS2_LIB") or os.path.join(HERE, "libss2.so")

NET_SPATIAL, NET_TEMPORAL, NET_SMOOTH = 0, 1, 2
MODE = {"NORMAL": 0, "FAST": 1}
TPS_EXACT, TPS_LATTICE = 0, 1
PROF_WARP, PROF_CONV = 0, 1

_vp, _i, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
_fp = ctypes.POINTER(ctypes.c_float)

# name -> (restype, argtypes); every symbol declared in include/ss2.h
SIGNATURES = {
    "ss2_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "ss2_destroy": (None, [_vp]),
    "ss2_last_error": (ctypes.c_char_p, [_vp]),
    "ss2_version": (ctypes.c_char_p, []),
    "ss2_launch_count": (_i64, [_vp, _i]),
    "ss2_profile_enable": (_i, [_vp, _i, _i]),
    "ss2_profile_read": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i64), ctypes.POINTER(ctypes.c_double)]),
    "ss2_load_tensor": (_i, [_vp, _i, ctypes.c_char_p, _vp, ctypes.POINTER(_i64), _i]),
    "ss2_finalize_weights": (_i, [_vp, _i]),
    "ss2_dlt": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "ss2_homo_warp": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ss2_tps_point": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "ss2_tps_warp": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ss2_tps_warp_blend_avg": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ss2_cost_volume_nhwc": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ss2_ccl_nhwc": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ss2_conv_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, ctypes.POINTER(_i64), _i, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "ss2_stem_pool": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "ss2_spatial_forward": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "ss2_build_spatial": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "ss2_spatial_tail": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "ss2_build_temporal": (_i, [_vp, _vp, _i, _vp, _vp]),
    "ss2_build_temporal_pair": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "ss2_build_spatial_temporal": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "ss2_tsmotion": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "ss2_build_smooth": (_i, [_vp] + [_vp] * 4 + [_i, _i] + [_vp] * 8 + [_vp]),
    "ss2_canvas_minmax": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "ss2_stable_frames": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _fp, _i, _i, _vp, _vp]),
    "ss2_stable_frames_u8": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _fp, _i, _i, _vp, _vp]),
    "ss2_three_view_meshes": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "ss2_three_view_frames": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _fp, _i, _i, _vp, _vp]),
    "ss2_canvas_size": (_i, [_fp, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "ss2_assemble_smooth": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "ss2_stream_meshes": (_i, [_vp, _vp, _vp, _i] + [_vp] * 6 + [_vp]),
    "ss2_stitch_stream_host": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i64,
                                    ctypes.POINTER(_i), ctypes.POINTER(_i), _vp, _vp]),
    "ss2_stitch_stream_host_async": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i64,
                                          ctypes.POINTER(_i), ctypes.POINTER(_i), _vp, _vp]),
    "ss2_stitch_stream_host_wait": (_i, [_vp, _i]),
    "ss2_stitch_stream_host_prefetch": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "ss2_linear_blend": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp]),
    "ss2_stable_frames_linear": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _fp, _i, _i, _vp, _vp]),
    "ss2_three_view_frames_linear": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _fp, _i, _i, _vp, _vp]),
    "ss2_nview_align": (_i, [_vp, ctypes.POINTER(_vp), _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ss2_nview_remap": (_i, [_vp, _i, _i, _vp, _vp, _fp, _vp, _vp, _vp]),
    "ss2_nview_frames": (_i, [_vp, ctypes.POINTER(_vp), _vp, _i, _i, _i, _i, _fp, _i, _i, _vp, _vp]),
    "ss2_assemble_paths": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "ss2_metric_scores": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "ss2_metric_psnr_ssim": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "ss2_load_frames_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "ss2_frames_to_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "ss2_stitch_stream_host_u8": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i64,
                                       ctypes.POINTER(_i), ctypes.POINTER(_i), _vp, _vp]),
    "ss2_stitch_stream_host_u8_async": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i64,
                                             ctypes.POINTER(_i), ctypes.POINTER(_i), _vp, _vp]),
    "ss2_stitch_stream_host_u8_prefetch": (_i, [_vp, _i, _vp, _vp, _i, _i, _i]),
    "ss2_stitch_stream_host_submit": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "ss2_stitch_stream_host_u8_submit": (_i, [_vp, _i, _vp, _vp, _i, _i, _i]),
    "ss2_stitch_stream_host_finish": (_i, [_vp, _i, _i, _i, _vp, _i64, ctypes.POINTER(_i), ctypes.POINTER(_i), _vp, _vp]),
}

_lib = None
_lock = threading.RLock()
_contexts = {}


class SS2Error(RuntimeError):
    pass


def load_library():
    """dlopen libss2.so and declare every prototype.  Works without a GPU (no CUDA call)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise SS2Error("libss2.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(stabstitch2_b200 has no CPU/PyTorch fallback)")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


class Context:
    """One ss2_ctx per CUDA device (per process)."""

    def __init__(self, device):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise SS2Error("stabstitch2_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = int(device)
        h = _vp()
        rc = self.lib.ss2_create(self.device, ctypes.byref(h))
        if rc != 0:
            raise SS2Error("ss2_create(device=%d) failed with %d" % (self.device, rc))
        self.handle = h
        self.synced = {}  # NET_ID -> signature of the module whose weights are loaded (NativeNet.sync_weights)

    def check(self, rc):
        if rc != 0:
            raise SS2Error("libss2 error %d: %s" % (rc, self.lib.ss2_last_error(self.handle).decode()))

    def launch_count(self, reset=False):
        return int(self.lib.ss2_launch_count(self.handle, 1 if reset else 0))

    def profile_enable(self, which, enable=True):
        self.check(self.lib.ss2_profile_enable(self.handle, which, 1 if enable else 0))

    def profile_read(self, which):
        ms, n, work = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
        self.check(self.lib.ss2_profile_read(self.handle, which, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(work)))
        return ms.value, n.value, work.value

    def load_state_dict(self, net_id, sd):
        for key, t in sd.items():
            if not torch.is_tensor(t) or not t.dtype.is_floating_point:
                continue  # num_batches_tracked etc.
            h = t.detach().to("cpu", torch.float32).contiguous()
            shape = (ctypes.c_int64 * max(h.dim(), 1))(*h.shape)
            self.check(self.lib.ss2_load_tensor(self.handle, net_id, key.encode(), h.data_ptr(), shape, h.dim()))
        self.check(self.lib.ss2_finalize_weights(self.handle, net_id))


def context(device=None):
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if isinstance(device, torch.device):
        device = device.index if device.index is not None else torch.cuda.current_device()
    with _lock:
        if device not in _contexts:
            _contexts[device] = Context(device)
        return _contexts[device]


def dev_f32(t, device=None):
    """contiguous fp32 CUDA view/copy of `t` (the reference calls .cuda() on everything)."""
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return t.to(device=device, dtype=torch.float32).contiguous()


def ptr(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def cur_stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def stem_pool(x, weight, bias=None, variant=2):
    """ss2_stem_pool: x [B,3,H,W] CUDA fp32 (NCHW), weight [64,3,7,7], bias [64] -> [B,Hp,Wp,64] (NHWC): conv 7x7 stride 2
    pad 3 + bias + ReLU + max-pool 3x3 stride 2 pad 1.  variant 2 = fused direct kernel, 1 = implicit GEMM + pool kernel,
    0 = SIMT fp32.  Test / reuse entry for the stem kernels."""
    ctx = context()
    x = dev_f32(x)
    B, _, H, W = x.shape
    hc, wc = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty(B, (hc - 1) // 2 + 1, (wc - 1) // 2 + 1, 64, device=x.device, dtype=torch.float32)
    wh = weight.detach().to("cpu", torch.float32).contiguous()
    bh = bias.detach().to("cpu", torch.float32).contiguous() if bias is not None else None
    ctx.check(ctx.lib.ss2_stem_pool(ctx.handle, ptr(x), B, H, W, _vp(wh.data_ptr()),
                                    _vp(bh.data_ptr()) if bh is not None else _vp(0), int(variant), ptr(out), cur_stream()))
    return out


def conv_nhwc(x, weight, bias=None, stride=1, pad=0, pad_d=0, relu=False, residual=None, use_tc=True):
    """ss2_conv_nhwc: x [B,(D,)H,W,Cin] CUDA fp32 channels-last, weight [Cout,Cin,(KD,)KH,KW] (any device),
    -> [B,(Do,)Ho,Wo,Cout].  Test / reuse entry for the convolution kernels."""
    ctx = context()
    x = dev_f32(x)
    three_d = weight.dim() == 5
    if not three_d:
        x5 = x.unsqueeze(1)
    else:
        x5 = x
    B, D, H, W, Cin = x5.shape
    wh = weight.detach().to("cpu", torch.float32).contiguous()
    bh = bias.detach().to("cpu", torch.float32).contiguous() if bias is not None else None
    Cout = wh.shape[0]
    KD = wh.shape[2] if three_d else 1
    KH, KW = wh.shape[-2], wh.shape[-1]
    Do = (D + 2 * pad_d - KD) + 1
    Ho = (H + 2 * pad - KH) // stride + 1
    Wo = (W + 2 * pad - KW) // stride + 1
    out = torch.empty(B, Do, Ho, Wo, Cout, device=x.device, dtype=torch.float32)
    res = dev_f32(residual) if residual is not None else None
    shape = (ctypes.c_int64 * wh.dim())(*wh.shape)
    ctx.check(ctx.lib.ss2_conv_nhwc(ctx.handle, ptr(x5), B, D, H, W, Cin, _vp(wh.data_ptr()), shape, wh.dim(),
                                    _vp(bh.data_ptr()) if bh is not None else _vp(0), stride, pad, pad_d,
                                    1 if relu else 0, ptr(res), 1 if use_tc else 0, ptr(out), cur_stream()))
    return out if three_d else out.squeeze(1)
