#!/bin/bash
# round-2 GPU session X: final 1080p / 4-view / reference-arm lines, sanitizer over the kernels changed last, conv ncu capture
mkdir -p gpurun_out
timeout 900 python bench.py --height 1080 --width 1920 --frames 16 > gpurun_out/r2x_bench_1080p.json 2> gpurun_out/r2x_bench_1080p.err
timeout 900 python bench.py --views 4 > gpurun_out/r2x_bench_4view.json 2> gpurun_out/r2x_bench_4view.err
python - <<'PY'
import json
for f in ['r2x_bench_1080p','r2x_bench_4view']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value', d['value'], 'ms', d.get('ms_per_step'), 'e2e', d.get('e2e',{}).get('value'), 'frac', (d.get('roofline') or {}).get('frac'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
( timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cost_volume or temporal_pair or conv_dc_every_tile_plan or stem_pool_direct_vs_torch and not 33 or conv_kernels or ccl_c256 or spatial_forward" 2>&1 | tail -n 8 ) > gpurun_out/r2x_sanitizer_memcheck.log 2>&1
tail -n 3 gpurun_out/r2x_sanitizer_memcheck.log
( timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cost_volume or conv_dc_every_tile_plan or stem_pool_direct_vs_torch and 2-44 or conv_kernels" 2>&1 | tail -n 8 ) > gpurun_out/r2x_sanitizer_racecheck.log 2>&1
tail -n 3 gpurun_out/r2x_sanitizer_racecheck.log
BQ="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-eager"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_stem_pool|conv_dc_kernel|conv_tc_kernel|cost_volume_tiled" -c 16 -o gpurun_out/r2x_conv $BQ > gpurun_out/r2x_ncu_conv.log 2>&1
ls -la gpurun_out/r2x_conv.ncu-rep
