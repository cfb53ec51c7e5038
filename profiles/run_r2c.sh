#!/bin/bash
# round-2 GPU session C: view-per-warp resampler (lat4), register-resident solve
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r2c_pytest.log 2>&1
for v in default mb4 mb2 nc16 nc4 pf0; do
  if [ $v = default ]; then unset SS2_LIB; else export SS2_LIB=$PWD/profiles/exp/libss2_$v.so; fi
  timeout 120 python profiles/warp_bench.py --tag $v >> gpurun_out/r2c_sweep.jsonl 2>> gpurun_out/r2c_sweep.err
done
unset SS2_LIB
SS2_TPS_L3=0 timeout 120 python profiles/warp_bench.py --tag old_lattice >> gpurun_out/r2c_sweep.jsonl 2>> gpurun_out/r2c_sweep.err
SS2_TPS_L3=3 timeout 120 python profiles/warp_bench.py --tag lat3 >> gpurun_out/r2c_sweep.jsonl 2>> gpurun_out/r2c_sweep.err
timeout 120 python profiles/warp_bench.py --tag default_1080 --height 1080 --width 1920 --frames 16 >> gpurun_out/r2c_sweep.jsonl 2>> gpurun_out/r2c_sweep.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tps_warp_lat4|tps_nodes|tps_solve" -c 3 -o gpurun_out/r2c_lat4 python profiles/warp_bench.py --iters 1 > gpurun_out/r2c_ncu.log 2>&1
timeout 300 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -5 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_sweep.jsonl; tail -c 700 gpurun_out/r2c_bench.json
