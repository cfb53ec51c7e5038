#!/bin/bash
# round-2 GPU session P: stem with 25 k-steps; full tests; bench; ncu --set full of the conv / stem / resampler kernels;
# compute-sanitizer memcheck + racecheck of the tcgen05 and lattice kernels
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stem_pool or spatial_forward" 2>&1 | tail -n 6 ) > gpurun_out/r2p_stem_test.log 2>&1
tail -n 3 gpurun_out/r2p_stem_test.log
if ! grep -q "passed" gpurun_out/r2p_stem_test.log || grep -q "failed" gpurun_out/r2p_stem_test.log; then exit 1; fi
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 ) > gpurun_out/r2p_pytest.log 2>&1
tail -n 5 gpurun_out/r2p_pytest.log | head -n 2
timeout 600 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2p_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e_fp32_interface']['value'], 'frac', d['roofline']['frac'], 'tensor', d['roofline_tensor']['achieved'])
PY
BQ="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-eager"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_stem_pool|conv_dc_kernel|conv_tc_kernel|cost_volume_tiled" -c 14 -o gpurun_out/r2p_conv $BQ > gpurun_out/r2p_ncu_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tps_warp_lattice|tps_nodes|tps_solve" -c 4 -o gpurun_out/r2p_warp $BQ > gpurun_out/r2p_ncu_warp.log 2>&1
ls -la gpurun_out/*.ncu-rep
( timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stem_pool_direct_vs_torch and not 33 or conv_kernels or tps_warp_golden or ccl_c256 or cost_volume_c128" 2>&1 | tail -n 12 ) > gpurun_out/r2p_sanitizer_memcheck.log 2>&1
tail -n 4 gpurun_out/r2p_sanitizer_memcheck.log
( timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stem_pool_direct_vs_torch and 1-8 or stem_pool_direct_vs_torch and 2-44 or conv_kernels or tps_warp_golden" 2>&1 | tail -n 12 ) > gpurun_out/r2p_sanitizer_racecheck.log 2>&1
tail -n 4 gpurun_out/r2p_sanitizer_racecheck.log
