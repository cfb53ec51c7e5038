#!/bin/bash
# round-2 GPU session L: split TF32 as two MMAs per k-step (A_hi x [B_hi|B_lo], A_lo x B_hi) in conv_dc and the stem
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stem_pool or conv_kernels or spatial_forward" 2>&1 | tail -n 25 ) > gpurun_out/r2l_conv_test.log 2>&1
tail -n 6 gpurun_out/r2l_conv_test.log
if grep -q "passed" gpurun_out/r2l_conv_test.log && ! grep -q "failed" gpurun_out/r2l_conv_test.log; then
  timeout 600 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
  ( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 ) > gpurun_out/r2l_pytest.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-eager > gpurun_out/r2l_ncu_bench.log 2>&1
  python profiles/launch_summary.py gpurun_out/r2l_launches.csv > gpurun_out/r2l_launches_summary.txt 2>&1
  tail -n 4 gpurun_out/r2l_pytest.log; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2l_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e_fp32_interface']['value'], 'frac', d['roofline']['frac'], 'tensor', d['roofline_tensor']['achieved'], d['roofline_tensor']['kernel_ms_per_step'])
PY
  head -n 14 gpurun_out/r2l_launches_summary.txt
fi
