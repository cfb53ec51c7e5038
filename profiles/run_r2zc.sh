#!/bin/bash
# round-2 GPU session ZC: solve with the row-slot branch once per step, multipliers zeroed in the pivot row by the searcher
mkdir -p gpurun_out
: > gpurun_out/r2zc_sweep.jsonl
for tag in default nw16; do
  if [ $tag = default ]; then lib=$PWD/stabstitch2_b200/libss2.so; else lib=$PWD/profiles/exp/libss2_$tag.so; fi
  SS2_LIB=$lib python profiles/warp_bench.py --tag $tag >> gpurun_out/r2zc_sweep.jsonl 2>> gpurun_out/r2zc_sweep.err
  SS2_LIB=$lib timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tps_solve" -c 5 --csv --log-file gpurun_out/r2zc_launches_$tag.csv python profiles/warp_bench.py --iters 2 > gpurun_out/r2zc_ncu_$tag.log 2>&1
  echo $tag; python profiles/launch_summary.py gpurun_out/r2zc_launches_$tag.csv 2>&1 | head -n 2
done
python - <<'PY'
import json
for l in open('gpurun_out/r2zc_sweep.jsonl'):
    d=json.loads(l); print('%-8s bracket %.4f ms  %.0f GB/s  checksum %.6f' % (d['tag'], d['bracket_ms'], d['bracket_gbs'], d['checksum']))
PY
( timeout 900 python -m pytest tests -m gpu -q -x -k "tps or fullsize or stream_golden or stable or three_view or nview or linear or dropin" 2>&1 | tail -n 5 ) > gpurun_out/r2zc_pytest.log 2>&1
tail -n 3 gpurun_out/r2zc_pytest.log
