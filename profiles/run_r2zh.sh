#!/bin/bash
# round-2 GPU session ZH: solve with the bulk column updates after the barrier (8 / 16 / 4 warps); racecheck
mkdir -p gpurun_out
: > gpurun_out/r2zh_sweep.jsonl
for tag in default nw16 nw4; do
  if [ $tag = default ]; then lib=$PWD/stabstitch2_b200/libss2.so; else lib=$PWD/profiles/exp/libss2_$tag.so; fi
  SS2_LIB=$lib python profiles/warp_bench.py --tag $tag >> gpurun_out/r2zh_sweep.jsonl 2>> gpurun_out/r2zh_sweep.err
  SS2_LIB=$lib timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tps_solve" -c 5 --csv --log-file gpurun_out/r2zh_launches_$tag.csv python profiles/warp_bench.py --iters 2 > gpurun_out/r2zh_ncu_$tag.log 2>&1
  echo $tag; python profiles/launch_summary.py gpurun_out/r2zh_launches_$tag.csv 2>&1 | head -n 2 | tail -n 1
done
python - <<'PY'
import json
for l in open('gpurun_out/r2zh_sweep.jsonl'):
    d=json.loads(l); print('%-8s bracket %.4f ms  %.0f GB/s  checksum %.6f' % (d['tag'], d['bracket_ms'], d['bracket_gbs'], d['checksum']))
PY
( timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tps_point_golden or u8_fused_store_bit_identical and NORMAL" 2>&1 | tail -n 6 ) > gpurun_out/r2zh_racecheck.log 2>&1
tail -n 4 gpurun_out/r2zh_racecheck.log
