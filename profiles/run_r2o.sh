#!/bin/bash
# round-2 GPU session O: conv_tc two-MMA split + coalesced epilogue, stem coalesced epilogue
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stem_pool or conv_kernels or golden or ccl" 2>&1 | tail -n 8 ) > gpurun_out/r2o_conv_test.log 2>&1
tail -n 4 gpurun_out/r2o_conv_test.log
if grep -q "passed" gpurun_out/r2o_conv_test.log && ! grep -q "failed" gpurun_out/r2o_conv_test.log; then
  timeout 600 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
  ( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 ) > gpurun_out/r2o_pytest.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-eager > gpurun_out/r2o_ncu_bench.log 2>&1
  python profiles/launch_summary.py gpurun_out/r2o_launches.csv > gpurun_out/r2o_launches_summary.txt 2>&1
  tail -n 4 gpurun_out/r2o_pytest.log; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2o_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e_fp32_interface']['value'], 'frac', d['roofline']['frac'], 'tensor', d['roofline_tensor']['achieved'], d['roofline_tensor']['kernel_ms_per_step'])
PY
  head -n 16 gpurun_out/r2o_launches_summary.txt
fi
