#!/bin/bash
# round-2 GPU session R: resampler bracket capture (traffic json), 1080p and 4-view bench lines on one GPU
mkdir -p gpurun_out
BQ="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-eager"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tps_warp_lattice|tps_nodes|tps_solve|stable_meshes" -s 8 -c 5 -o gpurun_out/r2r_warp $BQ > gpurun_out/r2r_ncu_warp.log 2>&1
timeout 900 python bench.py --height 1080 --width 1920 --frames 16 > gpurun_out/r2r_bench_1080p.json 2> gpurun_out/r2r_bench_1080p.err
timeout 900 python bench.py --views 4 > gpurun_out/r2r_bench_4view.json 2> gpurun_out/r2r_bench_4view.err
timeout 600 python bench.py --impl reference > gpurun_out/r2r_bench_reference.json 2> gpurun_out/r2r_bench_reference.err
python - <<'PY'
import json
for f in ['r2r_bench_1080p','r2r_bench_4view','r2r_bench_reference']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value', d['value'], 'ms', d.get('ms_per_step'), 'e2e', d.get('e2e',{}).get('value'), 'frac', (d.get('roofline') or {}).get('frac'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
