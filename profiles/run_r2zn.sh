#!/bin/bash
# round-2 GPU session ZN: SpatialNet and TemporalNet of a chunk on two streams (ss2_build_spatial_temporal): parity + A/B
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "one_call_bit_identical or stream_golden or temporal or test_stream_host_u8" 2>&1 | tail -n 6 ) > gpurun_out/r2zn_pytest.log 2>&1
tail -n 3 gpurun_out/r2zn_pytest.log
BQ="--no-cpu-baseline --no-gpu-eager"
SS2_NET_OVERLAP=0 timeout 600 python bench.py $BQ > gpurun_out/r2zn_bench_serial.json 2> gpurun_out/r2zn_bench_serial.err
timeout 600 python bench.py $BQ > gpurun_out/r2zn_bench_overlap.json 2> gpurun_out/r2zn_bench_overlap.err
python - <<'PY'
import json
for f in ['r2zn_bench_serial','r2zn_bench_overlap']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value', d.get('value'), 'ms', d.get('ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', (d.get('roofline') or {}).get('frac'), 'clocks', d.get('clocks'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
tail -n 2 gpurun_out/r2zn_bench_overlap.err
