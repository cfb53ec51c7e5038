#!/bin/bash
# round-2 GPU session ZI (end of the round, one GPU): full GPU test suite, smoke, default bench, launch list, ncu --set full of
# the resampler bracket (-> warp_kernel_traffic.json), 1080p / 4-view / reference-arm lines, sanitizer over the changed kernels
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 ) > gpurun_out/r2zi_pytest.log 2>&1
tail -n 6 gpurun_out/r2zi_pytest.log | head -n 3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 600 python bench.py > gpurun_out/r2zi_bench.json 2> gpurun_out/r2zi_bench.err
BQ="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-eager"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2zi_launches.csv $BQ > gpurun_out/r2zi_ncu_bench.log 2>&1
python profiles/launch_summary.py gpurun_out/r2zi_launches.csv > gpurun_out/r2zi_launches_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tps_warp_lattice|tps_nodes|tps_solve|stable_meshes" -c 12 -o gpurun_out/r2zi_warp $BQ > gpurun_out/r2zi_ncu_warp.log 2>&1
ls -la gpurun_out/r2zi_warp.ncu-rep
timeout 900 python bench.py --height 1080 --width 1920 --frames 16 > gpurun_out/r2zi_bench_1080p.json 2> gpurun_out/r2zi_bench_1080p.err
timeout 900 python bench.py --views 4 > gpurun_out/r2zi_bench_4view.json 2> gpurun_out/r2zi_bench_4view.err
timeout 900 python bench.py --impl reference > gpurun_out/r2zi_bench_reference.json 2> gpurun_out/r2zi_bench_reference.err
python - <<'PY'
import json
for f in ['r2zi_bench','r2zi_bench_1080p','r2zi_bench_4view','r2zi_bench_reference']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value', d.get('value'), 'ms', d.get('ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), 'frac', (d.get('roofline') or {}).get('frac'), 'clocks', d.get('clocks'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
head -n 14 gpurun_out/r2zi_launches_summary.txt
( timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tps_point_golden or tps_warp_golden or u8_fused_store_bit_identical or fullsize_frame_vs_oracle_and_arbiter and lattice-720 or stream_golden or test_host_edges_u8_bit_exact and 96" 2>&1 | tail -n 8 ) > gpurun_out/r2zi_sanitizer_memcheck.log 2>&1
tail -n 3 gpurun_out/r2zi_sanitizer_memcheck.log
( timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tps_point_golden or u8_fused_store_bit_identical and NORMAL or fullsize_frame_vs_oracle_and_arbiter and lattice-720" 2>&1 | tail -n 8 ) > gpurun_out/r2zi_sanitizer_racecheck.log 2>&1
tail -n 3 gpurun_out/r2zi_sanitizer_racecheck.log
