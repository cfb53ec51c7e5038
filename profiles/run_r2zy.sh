#!/bin/bash
# round-2 GPU session ZY: two input stages for the single-chunk kind::f16 layers (64 channels = one 128-byte row): plans + bench
mkdir -p gpurun_out
CONV_BENCH_ONLY=layer1 SS2_CONV_TEST_F16=3 timeout -s KILL 200 python profiles/conv_bench.py plans 5 > gpurun_out/r2zy_conv_f16_plans_layer1.jsonl 2> gpurun_out/r2zy_plans.err
CONV_BENCH_ONLY="64->64" SS2_CONV_TEST_F16=3 timeout -s KILL 200 python profiles/conv_bench.py dbg 5 > gpurun_out/r2zy_conv_f16_dbg.jsonl 2>> gpurun_out/r2zy_plans.err
cat gpurun_out/r2zy_conv_f16_plans_layer1.jsonl gpurun_out/r2zy_conv_f16_dbg.jsonl
( timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "f16 or conv_kernels or golden" 2>&1 | tail -n 5 ) > gpurun_out/r2zy_pytest.log 2>&1
tail -n 3 gpurun_out/r2zy_pytest.log
BQ="--no-cpu-baseline --no-gpu-eager"
run() { name=$1; shift; env "$@" timeout -s KILL 300 python bench.py $BQ > gpurun_out/r2zy_$name.json 2> gpurun_out/r2zy_$name.err; }
run f16_a SS2_F16=3
run f16_b SS2_F16=3
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2zy_*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value %.1f ms %.3f e2e %.1f frac %.4f convms %.3f' % (d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('roofline_tensor') or {}).get('kernel_ms_per_step')), d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
tail -n 3 gpurun_out/r2zy_f16_a.err gpurun_out/r2zy_plans.err
