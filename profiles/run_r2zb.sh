#!/bin/bash
# round-2 GPU session ZB: solve with 4 / 8 / 16 warps per system (phased, barrier-free); lattice kernel with the coordinates of
# two rows computed together and the taps of one row in flight (LAT_RPI=2, LAT_TPI=1)
mkdir -p gpurun_out
: > gpurun_out/r2zb_sweep.jsonl
for tag in old default nw4 nw16 split7 split6; do
  if [ $tag = default ]; then lib=$PWD/stabstitch2_b200/libss2.so; else lib=$PWD/profiles/exp/libss2_$tag.so; fi
  SS2_LIB=$lib python profiles/warp_bench.py --tag $tag >> gpurun_out/r2zb_sweep.jsonl 2>> gpurun_out/r2zb_sweep.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2zb_sweep.jsonl'):
    d=json.loads(l); print('%-8s bracket %.4f ms  %.0f GB/s  checksum %.6f' % (d['tag'], d['bracket_ms'], d['bracket_gbs'], d['checksum']))
PY
for tag in default nw4 nw16; do
if [ $tag = default ]; then lib=$PWD/stabstitch2_b200/libss2.so; else lib=$PWD/profiles/exp/libss2_$tag.so; fi
SS2_LIB=$lib timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tps_solve" -c 5 --csv --log-file gpurun_out/r2zb_launches_$tag.csv python profiles/warp_bench.py --iters 2 > gpurun_out/r2zb_ncu_$tag.log 2>&1
echo $tag; python profiles/launch_summary.py gpurun_out/r2zb_launches_$tag.csv 2>&1 | head -n 3
done
( timeout 900 python -m pytest tests -m gpu -q -x -k "tps or fullsize or stream_golden or stable or three_view or nview or linear or dropin" 2>&1 | tail -n 5 ) > gpurun_out/r2zb_pytest.log 2>&1
tail -n 3 gpurun_out/r2zb_pytest.log
