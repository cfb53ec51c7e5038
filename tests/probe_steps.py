"""GPU probe (not a pytest file): wall/GPU time split of the pipeline stages."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stabstitch2_b200 import _lib, pipeline, synthetic
from stabstitch2_b200.smooth_network import SmoothNet
from stabstitch2_b200.spatial_network import SpatialNet, build_SpatialNet
from stabstitch2_b200.temporal_network import TemporalNet

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, W = 720, 1280
s, t, m = SpatialNet().cuda().eval(), TemporalNet().cuda().eval(), SmoothNet().cuda().eval()
s.load_state_dict(synthetic.spatial_state_dict(mesh_scale=20.0), strict=True)
t.load_state_dict(synthetic.temporal_state_dict(mesh_scale=10.0), strict=True)
m.load_state_dict(synthetic.smooth_state_dict(), strict=True)
hr1 = torch.cat([synthetic.synth_frame(k, 0, H, W) for k in range(N)], 0)
hr2 = torch.cat([synthetic.synth_frame(k, 1, H, W) for k in range(N)], 0)
lr1, lr2 = synthetic.lowres(hr1).cuda(), synthetic.lowres(hr2).cuda()
hr1, hr2 = hr1.cuda(), hr2.cuda()
ctx = _lib.context()
ctx.sync = torch.cuda.synchronize
t.sync_weights(ctx); m.sync_weights(ctx); s.sync_weights(ctx)


def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); t_issue = time.perf_counter() - t0
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    print("%-28s host-issue %7.2f ms  wall %7.2f ms  gpu %7.2f ms  launches %d" % (
        name, 1e3 * t_issue / reps, 1e3 * t_all / reps, e0.elapsed_time(e1) / reps, ctx.launch_count(reset=True) // (reps + 1)))
    return r


ctx.launch_count(reset=True)
sp = timed("build_SpatialNet", lambda: build_SpatialNet(s, lr1, lr2))
from stabstitch2_b200.temporal_network import build_TemporalNet
tm = torch.empty(N, 7, 9, 2, device="cuda")
timed("build_temporal (1 view)", lambda: ctx.check(ctx.lib.ss2_build_temporal(ctx.handle, _lib.ptr(lr1), N, _lib.ptr(tm), _lib.cur_stream())))
S = timed("stream_meshes", lambda: pipeline.stream_meshes(s, t, m, lr1, lr2))
mm = pipeline.canvas_minmax(S[0], S[1], H, W).cpu().tolist()
out = torch.empty(N, 3, *pipeline.canvas_size(mm), device="cuda")
timed("stable_frames", lambda: pipeline.stable_frames(hr1, hr2, S[0], S[1], mm, out=out))
timed("stitch_stream", lambda: pipeline.stitch_stream(s, t, m, lr1, lr2, hr1, hr2))
