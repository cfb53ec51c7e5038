#!/bin/bash
# round-2 GPU session ZZb: layer3 entry on fp16 planes too (layer2's last block writes them when stage 2 follows): tests + final default bench
mkdir -p gpurun_out
( timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "golden or smooth or stream or spatial or temporal or bit_identical" 2>&1 | tail -n 5 ) > gpurun_out/r2zzb_pytest.log 2>&1
tail -n 2 gpurun_out/r2zzb_pytest.log
( timeout -s KILL 200 python __graft_entry__.py smoke 2>&1 | tail -n 3 ) > gpurun_out/r2zzb_smoke.log 2>&1; tail -n 1 gpurun_out/r2zzb_smoke.log
timeout -s KILL 500 python bench.py > gpurun_out/r2zzb_bench.json 2> gpurun_out/r2zzb_bench.err
python - <<'PY'
import json
for f in ['r2zzb_bench']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value', d.get('value'), 'ms', d.get('ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', (d.get('roofline') or {}).get('frac'), 'tensor', (d.get('roofline_tensor') or {}).get('achieved'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'eager', (d.get('gpu_eager_baseline') or {}).get('value'), 'dropin', (d.get('dropin_replay') or {}).get('value'), 'clocks', d.get('clocks'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
tail -n 2 gpurun_out/r2zzb_bench.err
