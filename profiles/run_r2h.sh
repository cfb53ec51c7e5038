#!/bin/bash
# round-2 GPU session H: metric path tests, solve kernel timing, 4-view line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -50 ) > gpurun_out/r2h_pytest.log 2>&1
for v in default nw32; do
  if [ $v = default ]; then unset SS2_LIB; else export SS2_LIB=$PWD/profiles/exp/libss2_$v.so; fi
  timeout 120 python profiles/warp_bench.py --tag $v >> gpurun_out/r2h_sweep.jsonl 2>> gpurun_out/r2h_sweep.err
done
unset SS2_LIB
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"tps_solve|tps_nodes|tps_warp_lattice|stable_meshes" -c 8 --csv --log-file gpurun_out/r2h_solve_launches.csv python profiles/warp_bench.py --iters 1 > /dev/null 2>&1
timeout 600 python bench.py --views 4 --frames 16 --steps 5 > gpurun_out/r2h_bench_4view.json 2> gpurun_out/r2h_bench_4view.err
grep -E "passed|failed|psnr|Error" gpurun_out/r2h_pytest.log | tail; cat gpurun_out/r2h_sweep.jsonl | cut -c1-200; grep -E "tps_solve|tps_nodes|lattice" gpurun_out/r2h_solve_launches.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | tail -12
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2h_bench_4view.json').read().strip().splitlines() if l.startswith('{')][-1])
print('4view', {k:d[k] for k in ['value','ms_per_step','canvas']}, d['e2e']['value'], d['roofline']['frac'])
PY
