#!/bin/bash
# round-2 GPU session ZO: programmatic dependent launches through the network launch chain (SS2_PDL): full GPU suite + A/B
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 8 ) > gpurun_out/r2zo_pytest.log 2>&1
tail -n 6 gpurun_out/r2zo_pytest.log
BQ="--no-cpu-baseline --no-gpu-eager"
for i in 1 2; do
SS2_PDL=0 timeout 600 python bench.py $BQ > gpurun_out/r2zo_bench_nopdl_$i.json 2> gpurun_out/r2zo_bench_nopdl_$i.err
timeout 600 python bench.py $BQ > gpurun_out/r2zo_bench_pdl_$i.json 2> gpurun_out/r2zo_bench_pdl_$i.err
done
python - <<'PY'
import json
for f in ['r2zo_bench_nopdl_1','r2zo_bench_pdl_1','r2zo_bench_nopdl_2','r2zo_bench_pdl_2']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.1f ms %.3f e2e %.1f frac %.4f convms %.3f' % (d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('roofline_tensor') or {}).get('kernel_ms_per_step')), d.get('clocks'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
tail -n 2 gpurun_out/r2zo_bench_pdl_2.err
