#!/bin/bash
# round-2 GPU session G (2 GPUs): sharded-stream parity under NCCL, 2-GPU bench lines (pair and 4-view)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2g_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -40 ) > gpurun_out/r2g_pytest.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/run_sharded_check.py > gpurun_out/r2g_sharded_check.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --views 4 --frames 16 --steps 5 > gpurun_out/r2g_bench_4view_2gpu.json 2> gpurun_out/r2g_bench_4view_2gpu.err
grep -E "passed|failed" gpurun_out/r2g_pytest.log; grep -E "sharded x" gpurun_out/r2g_sharded_check.log; python - <<'PY'
import json
for f in ('gpurun_out/r2g_bench_2gpu.json','gpurun_out/r2g_bench_4view_2gpu.json'):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, {k:d[k] for k in ['value','ms_per_step','n_gpus']}, 'e2e', d['e2e'] and d['e2e']['value'], 'shard_parity', d.get('shard_parity'))
    except Exception as e: print(f, 'failed', e)
PY
tail -n 3 gpurun_out/r2g_bench_2gpu.err; tail -n 3 gpurun_out/r2g_bench_4view_2gpu.err
