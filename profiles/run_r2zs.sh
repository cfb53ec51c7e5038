#!/bin/bash
# round-2 GPU session ZS (2 GPUs): sharded-stream parity under NCCL after ss2_build_spatial_temporal (both networks in one call, halo frame included), 2-GPU bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2zs_smi.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -20 ) > gpurun_out/r2zs_pytest_multi.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/run_sharded_check.py > gpurun_out/r2zs_sharded_check.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2zs_bench_2gpu.json 2> gpurun_out/r2zs_bench_2gpu.err
grep -E "passed|failed" gpurun_out/r2zs_pytest_multi.log; grep -E "sharded x|identical|max" gpurun_out/r2zs_sharded_check.log | tail -n 4; python - <<'PY'
import json
for f in ('gpurun_out/r2zs_bench_2gpu.json',):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, {k:d[k] for k in ['value','ms_per_step','n_gpus']}, 'e2e', d['e2e'] and d['e2e']['value'], 'shard_parity', d.get('shard_parity'))
    except Exception as e: print(f, 'failed', e)
PY
tail -n 3 gpurun_out/r2zs_bench_2gpu.err
