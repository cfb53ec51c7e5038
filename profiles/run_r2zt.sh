#!/bin/bash
# round-2 GPU session ZT: fp16 split planes + kind::f16 MMAs in the direct 3x3 kernel (conv2 of the BasicBlocks): unit tests, then the networks with SS2_F16=1
mkdir -p gpurun_out
( timeout -s KILL 200 python -m pytest tests -m gpu -q -x -k "conv_dc_f16_planes" 2>&1 | tail -n 30 ) > gpurun_out/r2zt_pytest_f16.log 2>&1
tail -n 30 gpurun_out/r2zt_pytest_f16.log
( SS2_F16=1 timeout -s KILL 400 python -m pytest tests -m gpu -q -k "golden or smooth or stream or spatial or temporal" 2>&1 | tail -n 30 ) > gpurun_out/r2zt_pytest_nets_f16.log 2>&1
tail -n 30 gpurun_out/r2zt_pytest_nets_f16.log
BQ="--no-cpu-baseline --no-gpu-eager"
run() { name=$1; shift; env "$@" timeout -s KILL 300 python bench.py $BQ > gpurun_out/r2zt_$name.json 2> gpurun_out/r2zt_$name.err; }
run base SS2_F16=0
run f16 SS2_F16=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2zt_*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value %.1f ms %.3f e2e %.1f frac %.4f convms %.3f' % (d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('roofline_tensor') or {}).get('kernel_ms_per_step')), d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
tail -n 3 gpurun_out/r2zt_f16.err
