#!/bin/bash
# round-2 GPU session Z: fp64 issue rate of the chip; solve with dead column slots skipped; nodes kernel with double-float
# accumulation (NODES_DF); lattice kernel with several cell rows per barrier pair (LAT_YC) - sweep through warp_bench
mkdir -p gpurun_out
profiles/exp/fp64_rate > gpurun_out/r2z_fp64_rate.txt 2>&1; cat gpurun_out/r2z_fp64_rate.txt
: > gpurun_out/r2z_sweep.jsonl
for tag in old default df yc2 yc4 yc4pf0 yc4nc8 yc2nc2; do
  if [ $tag = default ]; then lib=$PWD/stabstitch2_b200/libss2.so; else lib=$PWD/profiles/exp/libss2_$tag.so; fi
  SS2_LIB=$lib python profiles/warp_bench.py --tag $tag >> gpurun_out/r2z_sweep.jsonl 2>> gpurun_out/r2z_sweep.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2z_sweep.jsonl'):
    d=json.loads(l); print('%-8s bracket %.4f ms  %.0f GB/s  checksum %.6f' % (d['tag'], d['bracket_ms'], d['bracket_gbs'], d['checksum']))
PY
for tag in df yc4; do
SS2_LIB=$PWD/profiles/exp/libss2_$tag.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2z_launches_$tag.csv python profiles/warp_bench.py --iters 2 > gpurun_out/r2z_ncu_$tag.log 2>&1
python profiles/launch_summary.py gpurun_out/r2z_launches_$tag.csv 2>&1 | head -n 8 | grep -v "at::"
done
( SS2_LIB=$PWD/profiles/exp/libss2_yc4.so timeout 900 python -m pytest tests -m gpu -q -x -k "tps or fullsize or stream_golden or stable or three_view or nview or linear or smoke or dropin" 2>&1 | tail -n 5 ) > gpurun_out/r2z_pytest_yc4.log 2>&1
tail -n 3 gpurun_out/r2z_pytest_yc4.log
