#!/bin/bash
# round-2 GPU session W: two-view TemporalNet batch, 32-pair SpatialNet chunks: full validation
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 ) > gpurun_out/r2w_pytest.log 2>&1
tail -n 6 gpurun_out/r2w_pytest.log | head -n 3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 600 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-eager > gpurun_out/r2w_ncu_bench.log 2>&1
python profiles/launch_summary.py gpurun_out/r2w_launches.csv > gpurun_out/r2w_launches_summary.txt 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2w_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e_fp32_interface']['value'], 'frac', d['roofline']['frac'], 'tensor', d['roofline_tensor']['achieved'], 'launches', d['gpu_launches'])
PY
head -n 12 gpurun_out/r2w_launches_summary.txt
