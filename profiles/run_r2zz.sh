#!/bin/bash
# round-2 GPU session ZZ (end of the round, one GPU, fp16 split planes as the default): full GPU test suite, smoke, default bench
# with every leg, launch list, ncu --set full of the convolution kernels, 1080p / 4-view lines, sanitizer over the kind::f16 kernels
mkdir -p gpurun_out
( time timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 ) > gpurun_out/r2zz_pytest.log 2>&1
tail -n 6 gpurun_out/r2zz_pytest.log | head -n 3
( timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -n 3 ) > gpurun_out/r2zz_smoke.log 2>&1; tail -n 1 gpurun_out/r2zz_smoke.log
timeout -s KILL 600 python bench.py > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err
BQ="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-eager"
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2zz_launches.csv $BQ > gpurun_out/r2zz_ncu_bench.log 2>&1
python profiles/launch_summary.py gpurun_out/r2zz_launches.csv > gpurun_out/r2zz_launches_summary.txt 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"conv_dc_kernel|conv_tc_kernel|conv_stem_pool" -c 22 -o gpurun_out/r2zz_conv $BQ > gpurun_out/r2zz_ncu_conv.log 2>&1
python profiles/ncu_conv_summary.py gpurun_out/r2zz_conv.ncu-rep > gpurun_out/r2zz_conv_ncu_summary.txt 2>&1
rm -f gpurun_out/r2zz_conv.ncu-rep
timeout -s KILL 400 python bench.py --height 1080 --width 1920 --frames 16 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2zz_bench_1080p.json 2> gpurun_out/r2zz_bench_1080p.err
timeout -s KILL 400 python bench.py --views 4 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2zz_bench_4view.json 2> gpurun_out/r2zz_bench_4view.err
python - <<'PY'
import json
for f in ['r2zz_bench','r2zz_bench_1080p','r2zz_bench_4view']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value', d.get('value'), 'ms', d.get('ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', (d.get('roofline') or {}).get('frac'), 'tensor', (d.get('roofline_tensor') or {}).get('achieved'), 'clocks', d.get('clocks'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
head -n 16 gpurun_out/r2zz_launches_summary.txt
( timeout -s KILL 240 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_conv_kernels_f16_planes_vs_torch and 3-case or test_conv_dc_f16_planes and 3-shape0" 2>&1 | tail -n 8 ) > gpurun_out/r2zz_sanitizer_memcheck.log 2>&1
tail -n 3 gpurun_out/r2zz_sanitizer_memcheck.log
( timeout -s KILL 240 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_conv_kernels_f16_planes_vs_torch and 3-case2 or test_conv_kernels_f16_planes_vs_torch and 3-case3 or test_conv_dc_f16_planes and 3-shape0" 2>&1 | tail -n 8 ) > gpurun_out/r2zz_sanitizer_racecheck.log 2>&1
tail -n 3 gpurun_out/r2zz_sanitizer_racecheck.log
