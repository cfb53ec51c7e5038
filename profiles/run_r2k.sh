#!/bin/bash
# round-2 GPU session K: conv_dc column tiles (two input stages for layer1) and 64-wide N tiles for small maps: sweep
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-gpu-eager --steps 10 --warmup 3"
run() { # tag, env...
  tag=$1; shift
  line=$(env "$@" timeout 300 $B 2>gpurun_out/r2k_$tag.err | grep '^{' | tail -n 1)
  echo "$line" > gpurun_out/r2k_$tag.json
  python - "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2k_%s.json'%sys.argv[1]).read())
    print(sys.argv[1], 'fps %.1f ms %.3f tensor %.1f conv_ms %.3f'%(d['value'], d['ms_per_step'], d['roofline_tensor']['achieved'], d['roofline_tensor']['kernel_ms_per_step']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
run default SS2_X=0
run tw1 SS2_DC_TW=1
run tw2 SS2_DC_TW=2
run tw3 SS2_DC_TW=3
run tw4 SS2_DC_TW=4
run small0 SS2_DC_SMALL=0
run small148 SS2_DC_SMALL=148
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 ) > gpurun_out/r2k_pytest.log 2>&1
tail -n 4 gpurun_out/r2k_pytest.log
