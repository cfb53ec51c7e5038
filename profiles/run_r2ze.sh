#!/bin/bash
# round-2 GPU session ZE: uint8 store fused into the lattice resampler (ss2_stable_frames_u8): tests, default bench (e2e)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "u8 or fullsize or stream_golden or stable" 2>&1 | tail -n 8 ) > gpurun_out/r2ze_pytest.log 2>&1
tail -n 4 gpurun_out/r2ze_pytest.log
timeout 900 python bench.py > gpurun_out/r2ze_bench.json 2> gpurun_out/r2ze_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ze_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], d['e2e_fp32_interface']['value'], 'frac', d['roofline']['frac'], 'achieved', d['roofline']['achieved'], 'tensor', d['roofline_tensor']['achieved'], 'launches', d['gpu_launches'])
PY
tail -n 3 gpurun_out/r2ze_bench.err
