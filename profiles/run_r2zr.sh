#!/bin/bash
# round-2 GPU session ZR: CTA-pair direct 3x3 kernel: which kernel ran (ncu launch list of the tile-plan test), full suite, A/B
mkdir -p gpurun_out
SS2_DC_PAIR=1 timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_dc -c 60 --csv --log-file gpurun_out/r2zr_pair_launches.csv python -m pytest tests -m gpu -q -x -k "conv_dc_every_tile_plan" > gpurun_out/r2zr_ncu_pytest.log 2>&1
grep -c conv_dc_pair gpurun_out/r2zr_pair_launches.csv
( SS2_DC_PAIR=1 timeout -s KILL 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 8 ) > gpurun_out/r2zr_pytest_pair.log 2>&1
tail -n 4 gpurun_out/r2zr_pytest_pair.log
BQ="--no-cpu-baseline --no-gpu-eager"
run() { name=$1; shift; env "$@" timeout -s KILL 300 python bench.py $BQ > gpurun_out/r2zr_$name.json 2> gpurun_out/r2zr_$name.err; }
for i in 1 2; do
run base_$i SS2_DC_PAIR=0
run pair_$i SS2_DC_PAIR=1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2zr_*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value %.1f ms %.3f e2e %.1f frac %.4f convms %.3f' % (d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('roofline_tensor') or {}).get('kernel_ms_per_step')), d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
tail -n 3 gpurun_out/r2zr_pair_1.err
