#!/bin/bash
# round-2 GPU session ZL: run-to-run determinism of the resampler bracket (final library), default bench line with the
# regenerated traffic figure
mkdir -p gpurun_out
python profiles/determinism_check.py --iters 400 > gpurun_out/r2zl_determinism.json 2> gpurun_out/r2zl_determinism.err
cat gpurun_out/r2zl_determinism.json
for i in 1 2 3; do python profiles/warp_bench.py --tag run$i --iters 5; done > gpurun_out/r2zl_warp_repeat.jsonl 2>/dev/null
python - <<'PY'
import json
print([ (json.loads(l)['checksum'], round(json.loads(l)['bracket_ms'],4)) for l in open('gpurun_out/r2zl_warp_repeat.jsonl')])
PY
timeout 600 python bench.py > gpurun_out/r2zl_bench.json 2> gpurun_out/r2zl_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2zl_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ['value','ms_per_step']}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'clocks', d['clocks'])
PY
