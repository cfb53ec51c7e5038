#!/bin/bash
# round-2 GPU session ZM: is the process-to-process checksum difference of warp_bench.py in the INPUTS (torch CPU synthetic frames)?
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python profiles/warp_bench.py --tag run$i --iters 3; done > gpurun_out/r2zm_warp_repeat.jsonl 2>/dev/null
python - <<'PY'
import json
for l in open('gpurun_out/r2zm_warp_repeat.jsonl'):
    d=json.loads(l); print(d['tag'], d['checksum'], d['input_checksum'], d['mesh_checksum'])
PY
