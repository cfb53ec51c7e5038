#!/bin/bash
# round-2 GPU session ZW: fp16 split planes through the implicit-GEMM kernel too (stride-2 entries, shortcuts, Conv3d, CCL) and
# the regressor stacks: unit tests, whole suite, smoke, A/B of SS2_F16 = 0 / 1 / 3
mkdir -p gpurun_out
( timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "f16 or ccl or conv_kernels" 2>&1 | tail -n 30 ) > gpurun_out/r2zw_pytest_f16.log 2>&1
tail -n 30 gpurun_out/r2zw_pytest_f16.log
( time timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 30 ) > gpurun_out/r2zw_pytest.log 2>&1
tail -n 12 gpurun_out/r2zw_pytest.log
( timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -n 5 ) > gpurun_out/r2zw_smoke.log 2>&1
tail -n 3 gpurun_out/r2zw_smoke.log
BQ="--no-cpu-baseline --no-gpu-eager"
run() { name=$1; shift; env "$@" timeout -s KILL 300 python bench.py $BQ > gpurun_out/r2zw_$name.json 2> gpurun_out/r2zw_$name.err; }
run tf32 SS2_F16=0
run f16_dc SS2_F16=1
run f16_all SS2_F16=3
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2zw_*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value %.1f ms %.3f e2e %.1f frac %.4f convms %.3f' % (d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('roofline_tensor') or {}).get('kernel_ms_per_step')), d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
tail -n 3 gpurun_out/r2zw_f16_all.err
