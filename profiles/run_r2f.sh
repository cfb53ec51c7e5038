#!/bin/bash
# round-2 GPU session F: N-view chain + fused 3/4-view resampler, branch-free solve, solve warp-count variants
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -70 ) > gpurun_out/r2f_pytest.log 2>&1
for v in default nw32 nw8; do
  if [ $v = default ]; then unset SS2_LIB; else export SS2_LIB=$PWD/profiles/exp/libss2_$v.so; fi
  timeout 120 python profiles/warp_bench.py --tag $v >> gpurun_out/r2f_sweep.jsonl 2>> gpurun_out/r2f_sweep.err
done
unset SS2_LIB
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 python bench.py --views 4 --frames 16 --steps 5 > gpurun_out/r2f_bench_4view.json 2> gpurun_out/r2f_bench_4view.err
grep -E "passed|failed|view|Error|error" gpurun_out/r2f_pytest.log | tail -20; cat gpurun_out/r2f_sweep.jsonl; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e_fp32_interface']['value'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'])
print('dropin', d.get('dropin_replay')); print('eager', d.get('gpu_eager_baseline',{}).get('value'))
try:
    d=json.loads(open('gpurun_out/r2f_bench_4view.json').read().strip().splitlines()[-1])
    print('4view', {k:d[k] for k in ['value','ms_per_step','canvas']}, d['e2e'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])
except Exception as e: print('4view failed', e)
PY
tail -3 gpurun_out/r2f_bench.err gpurun_out/r2f_bench_4view.err
