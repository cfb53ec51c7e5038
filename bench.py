#!/usr/bin/env python
"""bench.py - stitched frames/s of the StabStitch++ inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one pass of the whole hot path (SpatialNet, TemporalNet x2, tsmotion, SmoothNet
windows, canvas, fused TPS resample + AVERAGE blend; test_online_tra.py:284-399 of the
reference) over one chunk of `--frames` consecutive synthetic 720p frame pairs per GPU.
N > 1 (under torchrun): the stream is sharded in time, rank r owns frames [r*F, (r+1)*F)
(weak scaling); KB-sized mesh halo all-gather + canvas min/max all-reduce over NCCL.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the
same through the C-ABI host calls with pinned HOST buffers, uint8 frames in and out (H2D + D2H
inside the timed region; `e2e_fp32_interface` = the fp32 tensor interface); `roofline` = the
resampler bracket's achieved algorithmic HBM GB/s against MEASURED_PEAKS.json, `roofline_tensor`
= the convolution kernels' algorithmic TFLOP/s; `cpu_baseline` = the CPU oracle port of the
reference timed on this box's host cores on a bounded sample, `gpu_eager_baseline` = the same
port run as eager PyTorch on the GPU; `shard_parity` (N > 1) = the sharded result against the
single-process one.  `--impl reference` times only the CPU port.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "frames/s"


def metric_name(H):
    """BASELINE.json's metric, named for the frame size actually run (720p is the quoted configuration)"""
    return "stitched frames/sec at %dp pair" % H


def make_config(H, W, F, world):
    """one definition of the workload for both arms (the driver compares the two `config` objects); results that
    depend on the data or the implementation (canvas size, field evaluation) are separate keys of the line"""
    return {"workload": "%dp synthetic pair stream, full Spatial+Temporal+Smooth inference + TPS "
                        "resample/AVERAGE blend" % H, "height": H, "width": W,
            "frames_per_step_per_gpu": F, "net_input": [360, 480], "window": 7,
            "l2": "inputs larger than L2: %.0f MB of frames per step per GPU" % (F * 2 * 3 * H * W * 4 / 1e6),
            "parallelism": "temporal shards x%d" % world}

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
# the one speed the reference publishes for this path (README.md:30,32: 28.3 fps end to end on
# one RTX 4090; the resolution is not stated, BASELINE.json labels it "720p pairs")
PUBLISHED_FPS = 28.3
BASELINE_NOTE = "value / 28.3 fps (reference README.md:30, 1x RTX 4090, resolution not stated; BASELINE.md section 1)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--frames", type=int, default=32, help="frame pairs per step per GPU")
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--tps", default="default", choices=["default", "exact", "lattice"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--views", type=int, default=2, choices=[2, 3, 4],
                    help="views per stitched frame: 2 = the pair pipeline (metric), 3 / 4 = the multi-view chain (config 5)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, in-process through NVML
    (nvidia_ml_py).  A polling `nvidia-smi -lms` child was measured to stall kernel launches for
    100-250 ms once per run on these hosts; NVML calls from a thread do not."""
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.first, self.stop_flag, self.thread, self.h = index, [], 0, False, None, None

    def mark(self):
        """samples before this point (warm-up) are not part of the timed region"""
        self.first = len(self.samples)

    def start(self):
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if visible:
                try:
                    idx = int(visible.split(",")[self.index])
                except (ValueError, IndexError):
                    idx = self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        masks = ((getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                 (getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                 (getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                 (getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap"))
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((mhz, [n for m, n in masks if bits & m]))
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "unavailable"}
        self.stop_flag = True
        self.thread.join(timeout=2)
        use = self.samples[self.first:] or self.samples[-1:]
        sm = sorted(x[0] for x in use)
        reasons = sorted({r for x in use for r in x[1]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": "nvml, 50 ms period"}


# ------------------------------------------------------------------------------------------
# CPU reference leg (the ONLY place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------
def cpu_reference_sample(H, W, warp_frames=2, n=7, want_canvas=False):
    """One bounded sample of the workload on the host cores: an n-frame stream (n-6 SmoothNet
    windows) through every network stage of the CPU oracle port, the resample+blend timed on
    `warp_frames` of the n frames and scaled to n.  Returns (frames/s, seconds spent)."""
    import torch
    from oracle import stabstitch_oracle as O
    from stabstitch2_b200 import synthetic
    sds = _cached("sds", lambda: synthetic.spatial_state_dict(mesh_scale=20.0))
    sdt = _cached("sdt", lambda: synthetic.temporal_state_dict(mesh_scale=10.0))
    sdm = _cached("sdm", lambda: synthetic.smooth_state_dict())
    hr = _cached("hr%dx%dx%d" % (n, H, W), lambda: [[synthetic.synth_frame(k, v, H, W) for k in range(n)] for v in range(2)])
    lr = _cached("lr%dx%dx%d" % (n, H, W), lambda: [[synthetic.lowres(x) for x in hr[v]] for v in range(2)])
    t0 = time.perf_counter()
    with torch.no_grad():
        sm1, sm2 = [], []
        for k in range(n):
            a, b = O.build_spatial(sds, lr[0][k], lr[1][k])
            sm1.append(a)
            sm2.append(b)
        tm1 = O.temporal_forward(sdt, lr[0])
        tm2 = O.temporal_forward(sdt, lr[1])
        smesh1, ts1 = O.tsmotion_prep(sm1, tm1)
        smesh2, ts2 = O.tsmotion_prep(sm2, tm2)
        S1, S2 = O.smooth_stream(sdm, smesh1, smesh2, ts1, ts2)
        m1, m2, wmin, hmin, ow, oh = O.canvas(S1, S2, H, W)
        t1 = time.perf_counter()
        for k in range(warp_frames):
            O.stable_frame(hr[0][k], hr[1][k], m1[:, k], m2[:, k], wmin, hmin, ow, oh)
        t2 = time.perf_counter()
    total = (t1 - t0) + (t2 - t1) * n / warp_frames
    if want_canvas:
        return n / total, t2 - t0, (int(oh.int()), int(ow.int()))
    return n / total, t2 - t0


_cache = {}


def _cached(key, fn):
    if key not in _cache:
        _cache[key] = fn()
    return _cache[key]


def run_reference(args):
    """--impl reference: the reference's own algorithm on the host cores.  The reference is
    pure Python/PyTorch without packaging (no setup.py) and /root/reference does not exist on
    the GPU box, so this arm times the CPU oracle port (oracle/stabstitch_oracle.py, pinned to
    the reference's outputs by tests/golden) with every host thread torch can use.  Same `config`
    as the native arm (the 32-frame-per-step stream); each timed step is a bounded sample of it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    H, W = args.height, args.width
    n_s, warp_s = 16, 4
    for _ in range(args.warmup):
        cpu_reference_sample(H, W, warp_frames=1, n=7)
    vals, t0 = [], time.perf_counter()
    canvas = None
    for _ in range(args.steps):
        fps, _, canvas = cpu_reference_sample(H, W, warp_frames=warp_s, n=n_s, want_canvas=True)
        vals.append(fps)
    wall = time.perf_counter() - t0
    # harmonic mean = total frames / total (scaled) time
    value = len(vals) / sum(1.0 / v for v in vals)
    sample = ("per step: %d consecutive frames of the %dx%d stream (%d SmoothNet windows) through all network stages; "
              "resample+blend timed on %d frames and scaled x%d; frames/s = %d / that time (the stages are per-frame, "
              "so the 32-frame step of `config` costs 2x this sample)" % (n_s, H, W, n_s - 6, warp_s, n_s // warp_s, n_s))
    line = {"impl": "reference", "metric": metric_name(H), "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": value / PUBLISHED_FPS if H == 720 else None,
            "baseline_note": BASELINE_NOTE, "dtype": "f32", "data": "synthetic",
            "config": make_config(H, W, args.frames, args.gpus),
            "canvas": list(canvas) if canvas else None, "tps_field": "exact (the reference's own 63-term evaluation)",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    from stabstitch2_b200 import _lib, pipeline, synthetic
    from stabstitch2_b200.smooth_network import SmoothNet
    from stabstitch2_b200.spatial_network import SpatialNet
    from stabstitch2_b200.temporal_network import TemporalNet
    from stabstitch2_b200.utils import torch_tps_transform as tt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tps = {"default": tt.DEFAULT_TPS, "exact": _lib.TPS_EXACT, "lattice": _lib.TPS_LATTICE}[args.tps]
    H, W, F = args.height, args.width, args.frames
    ctx = _lib.context()

    s, t, m = SpatialNet().cuda().eval(), TemporalNet().cuda().eval(), SmoothNet().cuda().eval()
    s.load_state_dict(synthetic.spatial_state_dict(mesh_scale=20.0), strict=True)
    t.load_state_dict(synthetic.temporal_state_dict(mesh_scale=10.0), strict=True)
    m.load_state_dict(synthetic.smooth_state_dict(), strict=True)

    # this rank's chunk of the stream (+ the one-frame input halo TemporalNet needs on rank > 0)
    f0 = rank * F
    halo = 1 if rank > 0 else 0
    hr1 = torch.cat([synthetic.synth_frame(k, 0, H, W) for k in range(f0 - halo, f0 + F)], 0)
    hr2 = torch.cat([synthetic.synth_frame(k, 1, H, W) for k in range(f0 - halo, f0 + F)], 0)
    lr1, lr2 = synthetic.lowres(hr1), synthetic.lowres(hr2)
    hr1, hr2 = hr1[halo:].contiguous(), hr2[halo:].contiguous()
    d_lr1, d_lr2, d_hr1, d_hr2 = lr1.cuda(), lr2.cuda(), hr1.cuda(), hr2.cuda()

    def step():
        if world == 1:
            return pipeline.stitch_stream(s, t, m, d_lr1, d_lr2, d_hr1, d_hr2, tps=tps)[0]
        return pipeline.stitch_stream_sharded(s, t, m, d_lr1, d_lr2, d_hr1, d_hr2, halo, tps=tps)[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi's start-up (NVML init) briefly contends with kernel launches: start it before the
    # warm-up so that only its periodic samples fall into the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        fused = step()
    barrier()
    Ho, Wo = int(fused.shape[2]), int(fused.shape[3])
    if rank == 0:
        sampler.mark()
    ctx.launch_count(reset=True)
    ctx.profile_enable(_lib.PROF_WARP, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    e0.record()
    for i in range(args.steps):
        fused = step()
        marks[i].record()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    per_step = sorted([(marks[i - 1] if i else e0).elapsed_time(marks[i]) for i in range(args.steps)])
    launches = ctx.launch_count()
    warp_ms, warp_n, warp_bytes = ctx.profile_read(_lib.PROF_WARP)
    ctx.profile_enable(_lib.PROF_WARP, False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tmax = torch.tensor([ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    value = world * F * args.steps / (ms / 1000.0)

    # ---- tensor-side roofline: two more (untimed) steps with the convolution kernels bracketed by events
    ctx.profile_enable(_lib.PROF_CONV, True)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    conv_ms, conv_n, conv_flops = ctx.profile_read(_lib.PROF_CONV)
    ctx.profile_enable(_lib.PROF_CONV, False)

    # ---- shard parity (N > 1, outside the timed region): rank 0 stitches the concatenated stream alone and every
    # rank's rows of the sharded result are compared with it
    shard_parity = None
    if world > 1:
        parts = [torch.empty_like(fused) for _ in range(world)]
        dist.all_gather(parts, fused.contiguous())
        lrs = [[torch.empty_like(d_lr1[halo:]) for _ in range(world)] for _ in range(2)]
        hrs = [[torch.empty_like(d_hr1) for _ in range(world)] for _ in range(2)]
        dist.all_gather(lrs[0], d_lr1[halo:].contiguous())
        dist.all_gather(lrs[1], d_lr2[halo:].contiguous())
        dist.all_gather(hrs[0], d_hr1)
        dist.all_gather(hrs[1], d_hr2)
        if rank == 0:
            ref = pipeline.stitch_stream(s, t, m, torch.cat(lrs[0], 0), torch.cat(lrs[1], 0), torch.cat(hrs[0], 0),
                                         torch.cat(hrs[1], 0), tps=tps)[0]
            got = torch.cat(parts, 0)
            same_shape = tuple(got.shape) == tuple(ref.shape)
            mx = float((got - ref).abs().max().item()) if same_shape else None
            per_rank = [bool(torch.equal(got[r * F:(r + 1) * F], ref[r * F:(r + 1) * F])) for r in range(world)] if same_shape else None
            shard_parity = {"max_abs": mx, "bit_identical": bool(same_shape and mx == 0.0), "per_rank_bit_identical": per_rank,
                            "frames_compared": int(ref.shape[0]), "canvas": [int(ref.shape[2]), int(ref.shape[3])],
                            "checksum_sharded": float(got.double().sum().item()), "checksum_single": float(ref.double().sum().item())}
            del ref, got
        del parts, lrs, hrs
        torch.cuda.empty_cache()

    # ---- e2e: HOST buffers through the C-ABI calls, H2D + D2H inside the timed region.  Primary: the uint8
    # interface (decoded BGR frames in, uint8 frames for the video writer out - the reference driver's own edges,
    # test_online_tra.py:252-264,152,414); also reported: the fp32 tensor interface of get_stable_sqe.
    def e2e_leg(u8):
        pins = outs = None
        ok = 1
        try:
            if u8:
                pins = [x.clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory() for x in (hr1, hr2)]
                outs = [torch.empty(F * 3 * (Ho + 64) * (Wo + 64), dtype=torch.uint8).pin_memory() for _ in range(3)]
            else:
                pins = [x.contiguous().pin_memory() for x in (lr1[halo:], lr2[halo:], hr1, hr2)]
                # every rank stitches ITS chunk as an independent stream here, so its canvas is the chunk's own (a few
                # pixels off the sharded run's global canvas): size the host buffers with a margin
                outs = [torch.empty(F * 3 * (Ho + 64) * (Wo + 64), dtype=torch.float32).pin_memory() for _ in range(3)]
        except RuntimeError as exc:
            ok = 0
            sys.stderr.write("rank %d: pinned host buffers unavailable (%s)\n" % (rank, str(exc).splitlines()[0]))
        if world > 1:  # every rank must take the same path through this leg (it contains barriers)
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if not ok:
            return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "pinned host buffers could not be allocated on every rank; leg skipped"}
        run = (lambda slot: pipeline.stitch_stream_host_u8_async(s, t, m, slot, *pins, outs[slot], tps=tps)) if u8 else \
              (lambda slot: pipeline.stitch_stream_host_async(s, t, m, slot, *pins, outs[slot], tps=tps))
        eh, ew = Ho, Wo
        for i in range(max(3, min(args.warmup, 6))):
            eh, ew = run(i % 3)
        for k in range(3):
            pipeline.stitch_stream_host_wait(k)
        barrier()
        t0 = time.perf_counter()
        # three slots rotating, two-phase calls: chunk i+1 is SUBMITTED (uploads + networks enqueued) before chunk i is
        # FINISHED (host waits for its data-dependent canvas, then enqueues resample + blend + downloads), so the GPU
        # never idles behind the canvas read or behind the downloads of chunk i-1; every chunk's inputs are copied from
        # pinned host memory and its frames land in pinned host memory inside the timed region
        pipeline.stitch_stream_host_submit(s, t, m, 0, *pins)
        for i in range(args.steps):
            if i + 1 < args.steps:
                pipeline.stitch_stream_host_submit(s, t, m, (i + 1) % 3, *pins)
            eh, ew = pipeline.stitch_stream_host_finish(i % 3, outs[i % 3], tps=tps)
        for k in range(3):
            pipeline.stitch_stream_host_wait(k)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tmax = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dt = float(tmax.item())
        esz = 1 if u8 else 4
        iface = ("uint8 BGR frames [n,H,W,3] in, uint8 frames [n,Ho,Wo,3] out (ss2_stitch_stream_host_u8_*; resize to "
                 "360x480, /127.5-1 and astype(uint8) on the device)") if u8 else \
                "fp32 tensors of the reference's get_stable_sqe interface (ss2_stitch_stream_host_*)"
        return {"value": world * F * args.steps / dt, "unit": UNIT,
                "h2d_bytes_per_step": int(sum(p.numel() for p in pins) * esz),
                "d2h_bytes_per_step": int(F * 3 * eh * ew * esz + 16), "interface": iface,
                "note": ("per GPU: every rank stitches its own chunk as an independent stream in this leg; " if world > 1 else "") +
                        "_submit/_finish/_wait through pinned host buffers: every chunk is copied H2D and its frames D2H "
                        "inside the timed region, three slots rotating"}

    e2e = e2e_fp32 = None
    if not args.no_e2e:
        e2e = e2e_leg(True)
        e2e_fp32 = e2e_leg(False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    # DRAM traffic of the same bracket from the committed `ncu --set full` capture of this round (a profiler figure
    # cannot be taken during a timed run); only reported when the capture was made on this exact configuration
    traffic = traffic_src = None
    try:
        import hashlib
        tj = json.load(open(os.path.join(ROOT, "profiles", "warp_kernel_traffic.json")))
        sha = hashlib.sha1(open(os.path.join(ROOT, "stabstitch2_b200", "csrc", "tps.cu"), "rb").read()).hexdigest()
        if (tj["height"], tj["width"], tj["frames_per_launch"], list(tj["canvas"])) != (H, W, F, [Ho, Wo]):
            traffic_src = "no ncu capture for this configuration"
        elif tj.get("tps_cu_sha1") != sha:
            traffic_src = "profiles/warp_kernel_traffic.json was captured from another version of csrc/tps.cu: dropped (re-run profiles/make_traffic_json.py)"
        else:
            traffic = float(tj["dram_bytes_read"] + tj["dram_bytes_write"])
            traffic_src = "profiles/warp_kernel_traffic.json: %s" % tj.get("source", "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch")
    except Exception:
        traffic = None
    achieved = (warp_bytes / warp_n) / (warp_ms / warp_n * 1e-3) / 1e9 if warp_n else None
    roofline = {"bound": "hbm", "kernel": "resampler bracket: stable_meshes + tps_solve + tps_nodes + tps_warp_lattice "
                                          "(fused TPS resample + AVERAGE blend), one bracket per 32-frame chunk",
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": warp_bytes / warp_n if warp_n else None,
                "avg_launch_ms": warp_ms / warp_n if warp_n else None, "launches_timed": warp_n,
                "share_of_step": warp_ms / ms if ms else None}
    tf_peak = tf_src = None
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        tf_peak, tf_src = float(d["bf16_tflops_sustained"]), "measured bf16 sustained (MEASURED_PEAKS.json)"
    except Exception:
        tf_peak, tf_src = 1389.4, "fallback"
    conv_tf = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms else None
    f16_planes = os.environ.get("SS2_F16", "3") != "0"
    roofline_tensor = {"bound": "tensor", "kernel": "all convolution / linear launches of a step (tcgen05 implicit GEMM + "
                                                    "direct 3x3, three exact split products per fp32 product: "
                                                    + ("fp16 split planes / kind::f16 MMAs behind the stem, split TF32 in "
                                                       "the stem)" if f16_planes else "split TF32 everywhere, SS2_F16=0)"),
                       "achieved": conv_tf, "peak": tf_peak, "peak_source": tf_src, "unit": "TFLOP/s",
                       "frac": conv_tf / tf_peak if conv_tf else None,
                       "note": "algorithmic fp32 FLOPs / summed kernel time; fp32-grade results cost 3 tensor-core products "
                               "per fp32 product: as fp16 MMAs the ceiling is 1/3 of this (fp16 = bf16 rate) peak, as TF32 "
                               "MMAs (half the rate) 1/6",
                       "frac_of_f16x3_ceiling": conv_tf / (tf_peak / 3.0) if conv_tf else None,
                       "frac_of_tf32x3_ceiling": conv_tf / (tf_peak / 6.0) if conv_tf else None,
                       "launches_timed": conv_n, "kernel_ms_per_step": conv_ms / 2.0, "flops_per_step": conv_flops / 2.0,
                       "traffic": None}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_reference_sample(H, W, warp_frames=1)  # warm the CPU caches / lazy inits
        fps, spent = cpu_reference_sample(H, W, warp_frames=16, n=48)
        cpu = {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": "48-frame %dx%d stream (42 SmoothNet windows) through all network stages of the CPU oracle "
                         "port; resample+blend timed on 16 frames and scaled x3 (%.1f s of CPU work)" % (H, W, spent)}
    gpu_eager = dropin = None
    if world == 1 and not args.no_gpu_eager:
        gpu_eager = gpu_eager_sample(H, W)
        dropin = dropin_replay_sample(H, W)
    line = {"metric": metric_name(H), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "step_ms": {"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1]},
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": value / PUBLISHED_FPS if H == 720 else None, "baseline_note": BASELINE_NOTE,
            "dtype": "f32", "data": "synthetic",
            "dtype_note": "fp32 tensors in and out, fp32-grade arithmetic: every fp32 product of a convolution is three exact "
                          "tensor-core products of 11-bit split operands (fp16 planes, TF32 in the stem) accumulated in fp32 - "
                          "meshes within 1e-3 px of the fp32 CPU reference (tests/, smoke); plain TF32 like the reference's "
                          "own GPU run fails that bound",
            "config": make_config(H, W, F, world), "canvas": [Ho, Wo],
            "tps_field": "exact" if tps == _lib.TPS_EXACT else "lattice",
            "clocks": clocks, "e2e": e2e, "e2e_fp32_interface": e2e_fp32, "gpu_launches": int(launches),
            "roofline": roofline, "roofline_tensor": roofline_tensor, "cpu_baseline": cpu,
            "gpu_eager_baseline": gpu_eager, "dropin_replay": dropin, "shard_parity": shard_parity}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def dropin_replay_sample(H, W, n=16):
    """The UNCHANGED driver's call sequence (tests/dropin_replay.py: batch-1 build_SpatialNet per pair, per-frame
    transformer + torch blend + .cpu() like test_online_tra.py:284-399) through the flat drop-in names: the frames/s
    a maintainer gets with zero edits, next to `value`, which batches a chunk per call."""
    import torch
    from stabstitch2_b200 import synthetic
    from tests import dropin_replay as R
    try:
        names = R.import_flat()
        s, t, m = names["SpatialNet"]().cuda().eval(), names["TemporalNet"]().cuda().eval(), names["SmoothNet"]().cuda().eval()
        s.load_state_dict(synthetic.spatial_state_dict(mesh_scale=20.0), strict=True)
        t.load_state_dict(synthetic.temporal_state_dict(mesh_scale=10.0), strict=True)
        m.load_state_dict(synthetic.smooth_state_dict(), strict=True)
        hr = [[synthetic.synth_frame(k, v, H, W) for k in range(n)] for v in range(2)]
        lr = [[synthetic.lowres(x) for x in hr[v]] for v in range(2)]
        R.replay(names, s, t, m, lr[0][:7], lr[1][:7], hr[0][:7], hr[1][:7])   # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        frames, _, _ = R.replay(names, s, t, m, lr[0], lr[1], hr[0], hr[1])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    except Exception as exc:
        return {"value": None, "error": "%s: %s" % (type(exc).__name__, str(exc).splitlines()[0][:200])}
    return {"value": n / dt, "unit": UNIT, "frames": n, "canvas": [int(frames[0].shape[0]), int(frames[0].shape[1])],
            "note": "per-frame Python loops of the reference driver, batch-1 calls, H2D of every frame and D2H of every fused "
                    "frame inside the timed region (wall clock)"}


def gpu_eager_sample(H, W, n=16, warp_frames=4):
    """The reference's algorithm as eager PyTorch ON THIS GPU ("the existing Blackwell kernel
    bar to beat"): the oracle port with every tensor on cuda, cuDNN / cuBLAS library kernels, TF32 defaults untouched
    (torch.backends.cudnn.allow_tf32 = True, matmul fp32), torch.cuda.synchronize() around each stage.  Same stream as
    the native arm; per-stage milliseconds for an n-frame chunk, resample+blend timed on `warp_frames` frames."""
    import torch
    from oracle import stabstitch_oracle as O
    from stabstitch2_b200 import synthetic
    dev = torch.device("cuda")
    to = lambda sd: {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sd.items()}  # noqa: E731
    sds, sdt, sdm = (to(synthetic.spatial_state_dict(mesh_scale=20.0)), to(synthetic.temporal_state_dict(mesh_scale=10.0)),
                     to(synthetic.smooth_state_dict()))
    hr = [[synthetic.synth_frame(k, v, H, W).to(dev) for k in range(n)] for v in range(2)]
    lr = [[synthetic.lowres(x) for x in hr[v]] for v in range(2)]
    stages = {}

    def timed(name, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        stages[name] = stages.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return r

    def once():
        stages.clear()
        with torch.no_grad(), torch.device(dev):
            sp = timed("spatial", lambda: [O.build_spatial(sds, lr[0][k], lr[1][k]) for k in range(n)])
            sm1, sm2 = [a for a, _ in sp], [b for _, b in sp]
            tm1 = timed("temporal", lambda: O.temporal_forward(sdt, lr[0]))
            tm2 = timed("temporal", lambda: O.temporal_forward(sdt, lr[1]))
            (smesh1, ts1), (smesh2, ts2) = timed("tsmotion", lambda: (O.tsmotion_prep(sm1, tm1), O.tsmotion_prep(sm2, tm2)))
            S1, S2 = timed("smooth", lambda: O.smooth_stream(sdm, smesh1, smesh2, ts1, ts2))
            m1, m2, wmin, hmin, ow, oh = timed("canvas", lambda: O.canvas(S1, S2, H, W))
            timed("warp_blend", lambda: [O.stable_frame(hr[0][k], hr[1][k], m1[:, k], m2[:, k], wmin, hmin, ow, oh)[0].cpu()
                                         for k in range(warp_frames)])
        stages["warp_blend"] *= n / warp_frames
        return sum(stages.values())

    try:
        once()  # warm-up: cuDNN autotune / lazy init
        total_ms = once()
    except Exception as exc:  # the port is CPU test infrastructure first; report rather than fail the bench
        return {"value": None, "error": "%s: %s" % (type(exc).__name__, str(exc).splitlines()[0][:200])}
    return {"value": n / (total_ms / 1e3), "unit": UNIT, "kind": "port (oracle/stabstitch_oracle.py) run as eager PyTorch on cuda:0",
            "frames": n, "stage_ms_per_chunk": {k: round(v, 2) for k, v in stages.items()},
            "note": "cuDNN/cuBLAS library kernels, TF32 convolutions (PyTorch default), one launch per ATen op, D2H of "
                    "each fused frame like the driver (test_online_tra.py:152); resample+blend timed on %d frames and "
                    "scaled x%d" % (warp_frames, n // warp_frames)}


def run_nview(args):
    """--views 3|4 (BASELINE.json config 5: the multi-video stitch of Full_model_inference on N GPUs): per step every rank
    runs the N-1 pair pipelines (SpatialNet, TemporalNet x2, tsmotion, SmoothNet) over its temporal shard of F frames,
    the middle-plane chain, and ONE fused N-image resample + AVERAGE blend.  Collectives per step: one mesh-halo
    all-gather per pair + two canvas all-reduces."""
    import torch
    import torch.distributed as dist
    from stabstitch2_b200 import _lib, pipeline, synthetic
    from stabstitch2_b200.smooth_network import SmoothNet
    from stabstitch2_b200.spatial_network import SpatialNet
    from stabstitch2_b200.temporal_network import TemporalNet
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, W, F, V = args.height, args.width, args.frames, args.views
    ctx = _lib.context()
    s, t, m = SpatialNet().cuda().eval(), TemporalNet().cuda().eval(), SmoothNet().cuda().eval()
    s.load_state_dict(synthetic.spatial_state_dict(mesh_scale=20.0), strict=True)
    t.load_state_dict(synthetic.temporal_state_dict(mesh_scale=10.0), strict=True)
    m.load_state_dict(synthetic.smooth_state_dict(), strict=True)
    f0, halo = rank * F, (1 if rank > 0 else 0)
    hr_all = [torch.cat([synthetic.synth_frame(k, v, H, W, chain=True) for k in range(f0 - halo, f0 + F)], 0) for v in range(V)]
    d_lr = [synthetic.lowres(x).cuda() for x in hr_all]
    hr = [x[halo:].contiguous() for x in hr_all]
    d_hr = [x.cuda() for x in hr]

    def step():
        return pipeline.stitch_nview_stream(s, t, m, d_lr, d_hr, halo)[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        fused = step()
    barrier()
    Ho, Wo = int(fused.shape[2]), int(fused.shape[3])
    if rank == 0:
        sampler.mark()
    ctx.launch_count(reset=True)
    ctx.profile_enable(_lib.PROF_WARP, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        fused = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count()
    warp_ms, warp_n, warp_bytes = ctx.profile_read(_lib.PROF_WARP)
    ctx.profile_enable(_lib.PROF_WARP, False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tmax = torch.tensor([ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    value = world * F * args.steps / (ms / 1000.0)
    # shard parity: rank 0 stitches the whole stream alone (outside the timed region)
    shard_parity = None
    if world > 1:
        parts = [torch.empty_like(fused) for _ in range(world)]
        dist.all_gather(parts, fused.contiguous())
        lrs = [[torch.empty_like(d_lr[v][halo:]) for _ in range(world)] for v in range(V)]
        hrs = [[torch.empty_like(d_hr[v]) for _ in range(world)] for v in range(V)]
        for v in range(V):
            dist.all_gather(lrs[v], d_lr[v][halo:].contiguous())
            dist.all_gather(hrs[v], d_hr[v])
        if rank == 0:
            # single-process call on the concatenated stream (no process group involved)
            pairs = [pipeline.stream_meshes(s, t, m, torch.cat(lrs[v], 0), torch.cat(lrs[v + 1], 0)) for v in range(V - 1)]
            hcat = [torch.cat(hrs[v], 0) for v in range(V)]
            sh, mi, mm1 = pipeline.nview_align(pairs, H, W)
            meshes, mm2 = pipeline.nview_remap(sh, mi, mm1.cpu().tolist())
            ref = pipeline.nview_frames(hcat, meshes, mm2.cpu().tolist())
            got = torch.cat(parts, 0)
            same = tuple(got.shape) == tuple(ref.shape)
            mx = float((got - ref).abs().max().item()) if same else None
            shard_parity = {"max_abs": mx, "bit_identical": bool(same and mx == 0.0), "frames_compared": int(ref.shape[0]),
                            "canvas": [int(ref.shape[2]), int(ref.shape[3])]}
            del ref, got
        del parts, lrs, hrs
        torch.cuda.empty_cache()
    # e2e: pinned host frames of all views in, fused frames out, through the public Python API
    e2e = None
    if not args.no_e2e:
        try:
            pin_lr = [x.cpu().pin_memory() for x in d_lr]
            pin_hr = [x.pin_memory() for x in hr]
            host_out = torch.empty(F, 3, Ho, Wo).pin_memory()   # the canvas of the (fixed) synthetic stream is known from warm-up
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                a = [x.cuda(non_blocking=True) for x in pin_lr]
                b = [x.cuda(non_blocking=True) for x in pin_hr]
                out = pipeline.stitch_nview_stream(s, t, m, a, b, halo)[0]
                host_out.copy_(out, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                tmax = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dt = float(tmax.item())
            e2e = {"value": world * F * args.steps / dt, "unit": UNIT,
                   "h2d_bytes_per_step": int(sum(x.numel() for x in pin_lr + pin_hr) * 4),
                   "d2h_bytes_per_step": int(F * 3 * Ho * Wo * 4),
                   "note": "fp32 tensors from pinned host memory through pipeline.stitch_nview_stream, fused frames back to "
                           "pinned host memory, copies inside the timed region (no chunk pipelining in this mode)"}
        except RuntimeError as exc:
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "note": "leg failed: " + str(exc).splitlines()[0][:160]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    achieved = (warp_bytes / warp_n) / (warp_ms / warp_n * 1e-3) / 1e9 if warp_n else None
    cfg = make_config(H, W, F, world)
    cfg["workload"] = "%d-view %dp synthetic multi-video stream: %d stitched pairs (Spatial+Temporal+Smooth) + middle-plane chain + " \
                      "fused %d-image TPS resample/AVERAGE blend" % (V, H, V - 1, V)
    cfg["views"] = V
    cfg["l2"] = "inputs larger than L2: %.0f MB of frames per step per GPU" % (F * V * 3 * H * W * 4 / 1e6)
    line = {"metric": "stitched frames/sec at %dp, %d views" % (H, V), "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "canvas": [Ho, Wo],
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "N-view resampler bracket: tps_solve + tps_nodes + tps_warp_lattice<V=%d> per 8-frame "
                                                    "chunk" % V, "achieved": achieved, "peak": peak, "peak_source": peak_src,
                         "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": None,
                         "algorithmic_bytes_per_launch": warp_bytes / warp_n if warp_n else None,
                         "avg_launch_ms": warp_ms / warp_n if warp_n else None, "launches_timed": warp_n,
                         "share_of_step": warp_ms / ms if ms else None},
            "cpu_baseline": None, "shard_parity": shard_parity}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.views > 2:
        run_nview(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
