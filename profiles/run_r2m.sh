#!/bin/bash
# round-2 GPU session M: L2 prefetch of the next tile's input boxes, conv1 outputs without the plain plane
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "stem_pool or conv_kernels or golden" 2>&1 | grep -v "^$" | tail -n 40 ) > gpurun_out/r2m_conv_test.log 2>&1
tail -n 30 gpurun_out/r2m_conv_test.log
B="python bench.py --no-e2e --no-cpu-baseline --no-gpu-eager --steps 10 --warmup 3"
run() { tag=$1; shift
  line=$(env "$@" timeout 300 $B 2>gpurun_out/r2m_$tag.err | grep '^{' | tail -n 1)
  echo "$line" > gpurun_out/r2m_$tag.json
  python - "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2m_%s.json'%sys.argv[1]).read())
    print(sys.argv[1], 'fps %.1f ms %.3f tensor %.1f conv_ms %.3f'%(d['value'], d['ms_per_step'], d['roofline_tensor']['achieved'], d['roofline_tensor']['kernel_ms_per_step']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
run prefetch1 SS2_DC_PREFETCH=1
run prefetch0 SS2_DC_PREFETCH=0
