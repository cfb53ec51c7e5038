#!/bin/bash
# round-2 GPU session E: LINEAR fusion, 8-warp solve, balanced temporal chunks
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -60 ) > gpurun_out/r2e_pytest.log 2>&1
timeout 120 python profiles/warp_bench.py --tag solve8 >> gpurun_out/r2e_sweep.jsonl 2>> gpurun_out/r2e_sweep.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
grep -E "passed|failed|LINEAR|linear" gpurun_out/r2e_pytest.log | tail -20; cat gpurun_out/r2e_sweep.jsonl; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e_fp32_interface']['value'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'])
PY
tail -3 gpurun_out/r2e_bench.err
