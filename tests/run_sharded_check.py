"""Launched under torchrun by tests/test_gpu_multi.py (or by hand):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_sharded_check.py
Every rank stitches its temporal shard of a synthetic stream with stitch_stream_sharded (NCCL halo
all-gather + canvas all-reduce); rank 0 also runs the whole stream alone and compares: the sharded
rows must be BIT-IDENTICAL to the single-process result."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stabstitch2_b200 import pipeline, synthetic  # noqa: E402
from stabstitch2_b200.smooth_network import SmoothNet  # noqa: E402
from stabstitch2_b200.spatial_network import SpatialNet  # noqa: E402
from stabstitch2_b200.temporal_network import TemporalNet  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    F, H, W = 8, 360, 640
    s, t, m = SpatialNet().cuda().eval(), TemporalNet().cuda().eval(), SmoothNet().cuda().eval()
    s.load_state_dict(synthetic.spatial_state_dict(mesh_scale=20.0), strict=True)
    t.load_state_dict(synthetic.temporal_state_dict(mesh_scale=10.0), strict=True)
    m.load_state_dict(synthetic.smooth_state_dict(), strict=True)
    N = world * F
    hr = [torch.cat([synthetic.synth_frame(k, v, H, W) for k in range(N)], 0) for v in range(2)]
    lr = [synthetic.lowres(x) for x in hr]
    plan = pipeline.shard_plan(rank, world, F)
    a, b, h = plan["start"], plan["stop"], plan["input_halo"]
    fused, s1, s2 = pipeline.stitch_stream_sharded(s, t, m, lr[0][a - h:b].cuda(), lr[1][a - h:b].cuda(),
                                                   hr[0][a:b].cuda(), hr[1][a:b].cuda(), h)
    # gather everything on rank 0
    shape = torch.tensor(list(fused.shape), device="cuda")
    shapes = [torch.empty_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape)
    assert all(torch.equal(x, shapes[0]) for x in shapes), "ranks disagree on the canvas"
    parts = [torch.empty_like(fused) for _ in range(world)]
    dist.all_gather(parts, fused.contiguous())
    m1 = [torch.empty_like(s1) for _ in range(world)]
    dist.all_gather(m1, s1.contiguous())
    ok = True
    if rank == 0:
        ref, r1, r2 = pipeline.stitch_stream(s, t, m, lr[0].cuda(), lr[1].cuda(), hr[0].cuda(), hr[1].cuda())
        got = torch.cat(parts, 0)
        dm = (torch.cat(m1, 0) - r1).abs().max().item()
        df = (got - ref).abs().max().item()
        print("sharded x%d vs single process: mesh max |diff| %.3e, fused frames max |diff| %.3e, canvas %s"
              % (world, dm, df, tuple(ref.shape[2:])))
        ok = tuple(got.shape) == tuple(ref.shape) and dm == 0.0 and df == 0.0
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
