#!/bin/bash
# round-2 GPU session ZK (8 GPUs, end of the round): scaling + shard parity of the pair stream and the 4-view stream
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8 > gpurun_out/r2zk_smi.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 > gpurun_out/r2zk_bench_8gpu.json 2> gpurun_out/r2zk_bench_8gpu.err
timeout 600 $TR --nproc-per-node 4 --master-port 29512 bench.py --gpus 4 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2zk_bench_4gpu.json 2> gpurun_out/r2zk_bench_4gpu.err
timeout 600 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --views 4 > gpurun_out/r2zk_bench_4view_8gpu.json 2> gpurun_out/r2zk_bench_4view_8gpu.err
python - <<'PY'
import json
for f in ['r2zk_bench_8gpu','r2zk_bench_4gpu','r2zk_bench_4view_8gpu']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f ms %.3f e2e %s shard_parity %s'%(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('shard_parity')))
    except Exception as e:
        print(f, 'FAILED', e)
PY
tail -n 3 gpurun_out/r2zk_bench_8gpu.err
