"""Import the UNMODIFIED reference (nie-lang/StabStitch2, Full_model_inference/Codes) on CPU.

Only usable in the build container (where /root/reference exists).  Used by
tests/golden/make_golden.py to generate the committed golden fixtures and by the
`not gpu` tests that pin the oracle against the live reference.  Nothing on the GPU
box imports this file's REF path.

Shims:
  1. torchvision resnet18(weights="DEFAULT") -> weights=None (no network)
     (spatial_network.py:268, temporal_network.py:113)
  2. Tensor.cuda / Module.cuda -> identity (unconditional .cuda() calls,
     spatial_network.py:84,88,304; temporal_network.py:28,131-134)
  3. stub modules imageio / skimage / matplotlib (imported, unused on the path)
  4. sys.path insert of the flat-import script directory
"""
import os
import sys
import types
import importlib

REF_DIR = "/root/reference/Full_model_inference/Codes"


def available():
    return os.path.isdir(REF_DIR)


_loaded = {}


def load():
    """Returns a dict of the reference modules, imported with the shims applied."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present: " + REF_DIR)
    import torch
    import torchvision

    # shim 1
    _orig_resnet18 = torchvision.models.resnet.resnet18

    def _resnet18_noweights(*a, **kw):
        kw.pop("weights", None)
        kw.pop("pretrained", None)
        return _orig_resnet18(weights=None)

    torchvision.models.resnet.resnet18 = _resnet18_noweights
    torchvision.models.resnet18 = _resnet18_noweights

    # shim 2 (only when there is no GPU)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **kw: self
        torch.nn.Module.cuda = lambda self, *a, **kw: self

    # shim 3
    for name in ("imageio", "skimage", "skimage.measure", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = types.ModuleType(name)
                if name == "matplotlib.pyplot":
                    m.rcParams = {}
                sys.modules[name] = m
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    # shim 4: import under private names so they never collide with our drop-in modules
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in
                  ("grid_res", "utils", "utils.torch_DLT", "utils.torch_homo_transform",
                   "utils.torch_tps_transform", "utils.torch_tps_transform_point",
                   "spatial_network", "temporal_network", "smooth_network", "test_online_tra", "test_metric_ssd")}
    for k in saved_mods:
        sys.modules.pop(k, None)
    sys.path.insert(0, REF_DIR)
    try:
        mods = {}
        for name in ("grid_res", "utils.torch_DLT", "utils.torch_homo_transform",
                     "utils.torch_tps_transform", "utils.torch_tps_transform_point",
                     "spatial_network", "temporal_network", "smooth_network"):
            mods[name] = importlib.import_module(name)
        # the driver script: import its functions without running __main__
        mods["test_online_tra"] = importlib.import_module("test_online_tra")
        mods["test_metric_ssd"] = importlib.import_module("test_metric_ssd")
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    _loaded.update(mods)
    return _loaded
