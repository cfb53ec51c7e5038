"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/ss2.h
declares (no compute calls), the drop-in modules keep the reference's state-dict contract,
the synthetic workload generators agree with the oracle's copies, and the temporal sharding
logic (plan + halo all-gather + canvas all-reduce) reproduces the single-process result under
a world_size-2 gloo group with the oracle doing the per-rank arithmetic."""
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

from oracle import stabstitch_oracle as O
from oracle import weights as Wt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ss2.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ss2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from stabstitch2_b200 import _lib
    lib = _lib.load_library()
    names = _header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libss2.so does not export " + n
        assert n in _lib.SIGNATURES, "no ctypes prototype for " + n
    assert sorted(_lib.SIGNATURES) == names
    assert lib.ss2_version().startswith(b"ss2 ")
    # host-only helper: canvas truncation like torch .int() (test_online_tra.py:140)
    from stabstitch2_b200.pipeline import canvas_size
    assert canvas_size([-10.5, 1733.9, -3.25, 724.0]) == (727, 1744)


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from stabstitch2_b200 import _lib
    from stabstitch2_b200.utils.torch_DLT import tensor_DLT
    with pytest.raises(_lib.SS2Error):
        tensor_DLT(torch.zeros(1, 4, 2), torch.zeros(1, 4, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "stabstitch2_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_state_dict_contract():
    """strict load of checkpoints with the reference's key names."""
    from stabstitch2_b200.smooth_network import SmoothNet
    from stabstitch2_b200.spatial_network import SpatialNet
    from stabstitch2_b200.temporal_network import TemporalNet
    for cls, sd in ((SpatialNet, Wt.spatial_state_dict()), (TemporalNet, Wt.temporal_state_dict()),
                    (SmoothNet, Wt.smooth_state_dict())):
        net = cls().eval()
        res = net.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        own = net.state_dict()
        assert set(own) == set(sd)
        for k in sd:
            assert tuple(own[k].shape) == tuple(sd[k].shape), k
    n = sum(v.numel() for k, v in SpatialNet().state_dict().items() if v.dtype.is_floating_point
            and "running" not in k)
    assert n == 11138756


def test_synthetic_generators_match_oracle_copies():
    from stabstitch2_b200 import synthetic as S
    for a, b in ((S.spatial_state_dict(mesh_scale=20.0), Wt.spatial_state_dict(mesh_scale=20.0)),
                 (S.temporal_state_dict(), Wt.temporal_state_dict()), (S.smooth_state_dict(), Wt.smooth_state_dict())):
        assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    x, y = S.synth_frame(5, 1, 90, 160), O.synth_frame(5, 1, 90, 160)
    assert torch.equal(x, y) and torch.equal(S.lowres(x), O.lowres(y))


def test_shard_plan():
    from stabstitch2_b200.pipeline import shard_plan
    with pytest.raises(ValueError):
        shard_plan(0, 2, 6)
    for world in (1, 2, 4, 8):
        for F in (7, 8, 16):
            seen = []
            for r in range(world):
                p = shard_plan(r, world, F)
                assert p["stop"] - p["start"] == F and p["ctx0"] == max(0, p["start"] - 6)
                assert p["with_head"] == (r == 0) and p["input_halo"] == (1 if r else 0)
                out = p["nwin"] + 6 if p["with_head"] else p["nwin"]
                assert out == F
                seen += list(range(p["start"], p["stop"]))
            assert seen == list(range(world * F))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sharded_worker(rank, world, port, F, raws, q):
    """Per-rank arithmetic with the ORACLE's ops, distributed logic from the product."""
    import torch.distributed as dist
    from stabstitch2_b200 import pipeline
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        plan = pipeline.shard_plan(rank, world, F)
        raw = raws[:, plan["start"]:plan["stop"]].clone()          # what this rank's nets produced
        allraw = pipeline.exchange_raw_meshes(raw)
        c0, stop = plan["ctx0"], plan["stop"]
        rig = O.rigid_mesh(1, 360, 480)
        nrig = O.norm_mesh(rig, 360, 480)
        smesh, ts = [[], []], [[], []]
        for v in range(2):
            for k in range(c0, stop):
                sm = rig + allraw[v, k][None]
                if k == 0:
                    t = sm * 0
                else:
                    moved = O.tps_point(O.norm_mesh(rig + allraw[2 + v, k][None], 360, 480), nrig,
                                        O.norm_mesh(rig + allraw[v, k - 1][None], 360, 480))
                    t = O.recover_mesh(moved, 360, 480) - sm
                smesh[v].append(sm)
                ts[v].append(t)
        sdm = Wt.smooth_state_dict()
        S = [[], []]
        for w in range(plan["nwin"]):
            a1, a2 = list(ts[0][w:w + 7]), list(ts[1][w:w + 7])
            a1[0], a2[0] = a1[0] * 0, a2[0] * 0
            o = O.smooth_forward(sdm, a1, a2, smesh[0][w:w + 7], smesh[1][w:w + 7])
            for v, key in enumerate(("smooth_mesh1", "smooth_mesh2")):
                if w == 0 and plan["with_head"]:
                    S[v] += [o[key][:, i] for i in range(7)]
                else:
                    S[v].append(o[key][:, 6])
        S1, S2 = torch.cat(S[0], 0), torch.cat(S[1], 0)
        assert S1.shape[0] == F
        m1, m2, wmin, hmin, ow, oh = O.canvas(S1[None], S2[None], 180, 320)
        mm = torch.stack([wmin, torch.maximum(m1[..., 0].max(), m2[..., 0].max()), hmin,
                          torch.maximum(m1[..., 1].max(), m2[..., 1].max())])
        g = pipeline.allreduce_canvas(mm)
        q.put((rank, S1.numpy(), S2.numpy(), g.numpy()))
    finally:
        dist.destroy_process_group()


def test_sharded_stream_matches_single_process_gloo():
    import torch.multiprocessing as mp
    world, F = 2, 7
    N = world * F
    g = torch.Generator().manual_seed(3)
    # raw network outputs of a 14-frame stream: smotion1/2 (view 2 shifted), tmotion1/2
    raws = torch.randn(4, N, 7, 9, 2, generator=g) * torch.tensor([3.0, 3.0, 1.0, 1.0])[:, None, None, None, None]
    raws[0, ..., 0] -= 80.0
    raws[1, ..., 0] += 80.0
    # single process: the oracle's own stream functions
    with torch.no_grad():
        sm1 = [raws[0, k][None] for k in range(N)]
        sm2 = [raws[1, k][None] for k in range(N)]
        tm1 = [raws[2, k][None] for k in range(N)]
        tm2 = [raws[3, k][None] for k in range(N)]
        smesh1, ts1 = O.tsmotion_prep(sm1, tm1)
        smesh2, ts2 = O.tsmotion_prep(sm2, tm2)
        S1, S2 = O.smooth_stream(Wt.smooth_state_dict(), smesh1, smesh2, ts1, ts2)
        m1, m2, wmin, hmin, ow, oh = O.canvas(S1, S2, 180, 320)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, F, raws, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, a, b, mm = q.get(timeout=300)
        res[r] = (a, b, mm)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got1 = np.concatenate([res[r][0] for r in range(world)], 0)
    got2 = np.concatenate([res[r][1] for r in range(world)], 0)
    assert np.abs(got1 - S1[0].numpy()).max() < 1e-4
    assert np.abs(got2 - S2[0].numpy()).max() < 1e-4
    for r in range(world):
        mm = res[r][2]
        assert abs(mm[0] - float(wmin)) < 1e-3 and abs(mm[2] - float(hmin)) < 1e-3
        assert abs((mm[1] - mm[0]) - float(ow)) < 1e-3 and abs((mm[3] - mm[2]) - float(oh)) < 1e-3


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys"""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--height", "96", "--width", "128"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0 and "workload" in line["config"]


def test_canvas_size_truncation_sweep():
    """ss2_canvas_size: extent = max - min in fp32, then truncation like torch's .int() (test_online_tra.py:119-120,140),
    swept across integer boundaries (host arithmetic only: no GPU needed)."""
    from stabstitch2_b200.pipeline import canvas_size
    rng = np.random.default_rng(3)
    for _ in range(4000):
        lo = np.float32(rng.uniform(-600, 600))
        k = np.float32(rng.integers(1, 4000))
        eps = np.float32(rng.choice([-1.0, 1.0]) * 10.0 ** rng.uniform(-6, -1))
        hi = np.float32(np.float32(lo + k) + eps)
        lo_y = np.float32(rng.uniform(-50, 50))
        hi_y = np.float32(lo_y + np.float32(rng.uniform(1, 2000)))
        ref_w = int((torch.tensor(hi) - torch.tensor(lo)).int())
        ref_h = int((torch.tensor(hi_y) - torch.tensor(lo_y)).int())
        assert canvas_size([float(lo), float(hi), float(lo_y), float(hi_y)]) == (ref_h, ref_w)


def test_dropin_flat_names_resolve_to_this_package():
    """with stabstitch2_b200/dropin first on sys.path the driver's import block
    (test_online_tra.py:7-9,14-15,21-23) resolves to this package, with every name the drivers use."""
    from tests import dropin_replay as R
    n = R.import_flat()
    for k in ("build_SpatialNet", "SpatialNet", "build_TemporalNet", "TemporalNet", "build_SmoothNet", "SmoothNet"):
        assert n[k].__module__.startswith("stabstitch2_b200."), (k, n[k].__module__)
    assert n["torch_tps_transform"].transformer.__module__ == "stabstitch2_b200.utils.torch_tps_transform"
    assert n["torch_tps_transform_point"].transformer.__module__ == "stabstitch2_b200.utils.torch_tps_transform_point"
    assert (n["grid_res"].GRID_H, n["grid_res"].GRID_W) == (6, 8)
    # the flat modules did not leak into sys.modules (a later import of the reference must not pick them up)
    assert "spatial_network" not in sys.modules or not getattr(sys.modules["spatial_network"], "__file__", "").startswith(R.DROPIN)


def test_reference_driver_imports_against_the_dropin():
    """The UNCHANGED reference driver module imports (and finds every name it needs) with the drop-in shims in place
    of its own model / utils modules.  Build container only (/root/reference is not on the GPU box)."""
    ref = "/root/reference/Full_model_inference/Codes"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    import subprocess
    code = r'''
import sys, types
for name in ("imageio", "skimage", "skimage.measure", "matplotlib", "matplotlib.pyplot"):
    try:
        __import__(name)
    except Exception:
        m = types.ModuleType(name)
        if name == "matplotlib.pyplot":
            m.rcParams = {}
        sys.modules[name] = m
sys.path.insert(0, "%s")                      # the driver's directory (for nothing but the driver itself)
sys.path.insert(0, "%s")                      # the repository root
sys.path.insert(0, "%s")                      # the flat-name shims FIRST
import test_online_tra as D
import inspect
assert D.SpatialNet.__module__ == "stabstitch2_b200.spatial_network", D.SpatialNet.__module__
assert D.build_TemporalNet.__module__ == "stabstitch2_b200.temporal_network"
assert D.build_SmoothNet.__module__ == "stabstitch2_b200.smooth_network"
assert D.torch_tps_transform.transformer.__module__ == "stabstitch2_b200.utils.torch_tps_transform"
assert D.torch_tps_transform_point.transformer.__module__ == "stabstitch2_b200.utils.torch_tps_transform_point"
assert "fusion_mode" in inspect.signature(D.get_stable_sqe).parameters
net = D.SpatialNet()
assert any(k.startswith("regressNet2_part1_ref") for k in net.state_dict())
print("ok")
''' % (ref, ROOT, os.path.join(ROOT, "stabstitch2_b200", "dropin"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_f16_split_planes_arithmetic():
    """The arithmetic the fp16 split planes rest on (csrc/common.cuh f16_split), restated in numpy:
    h = fp16(v), l = fp16((v - h) * 2^11) reproduces v to 2^-22 relative in fp16's normal range and to 2^-36 absolute below it,
    and three products of the halves, the two correction products accumulated with the factor 2^11 and folded in at the end
    (D1 + D2 / 2048), reproduce an fp32 dot product to ~2^-21 - the accuracy of the split-TF32 form."""
    rng = np.random.default_rng(7)
    mag = 10.0 ** rng.uniform(-9, 4.5, size=200000)
    v = (rng.standard_normal(200000) * mag).astype(np.float32)
    v = v[np.abs(v) <= 65504]

    def split(x):
        h = x.astype(np.float16)
        l = ((x - h.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)   # the subtraction is exact in fp32
        return h, l

    h, l = split(v)
    assert np.isfinite(h.astype(np.float32)).all() and np.isfinite(l.astype(np.float32)).all()
    rec = h.astype(np.float64) + l.astype(np.float64) / 2048.0
    err = np.abs(rec - v.astype(np.float64))
    normal = np.abs(v) >= 2.0 ** -14
    assert (err[normal] <= np.abs(v[normal]).astype(np.float64) * 2.0 ** -22).all()
    assert (err[~normal] <= 2.0 ** -36).all()
    # dot products of K = 576 terms (a 64-channel 3x3 layer), activations and weights of mixed magnitude
    K, M = 576, 2000
    a = (rng.standard_normal((M, K)) * 10.0 ** rng.uniform(-3, 1, size=(M, 1))).astype(np.float32)
    b = (rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    ah, al = split(a)
    bh, bl = split(b)
    f64 = lambda x: x.astype(np.float64)  # noqa: E731  (products of 11-bit halves are exact in fp32; the sums are taken in fp64 here)
    d1 = (f64(ah) * f64(bh)).sum(1)
    d2 = (f64(ah) * f64(bl)).sum(1) + (f64(al) * f64(bh)).sum(1)
    got = d1 + d2 / 2048.0
    ref = (f64(a) * f64(b)).sum(1)
    scale = (np.abs(f64(a)) * np.abs(f64(b))).sum(1)
    assert (np.abs(got - ref) <= scale * 2.0 ** -21).all()
    # without the correction products (plain fp16 operands) the same bound fails by orders of magnitude
    assert (np.abs(d1 - ref) > scale * 2.0 ** -21).mean() > 0.9
