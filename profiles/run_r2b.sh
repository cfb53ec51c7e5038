#!/bin/bash
# round-2 GPU session A: parity of the new resampler, variant sweep, ncu capture, bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2b_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2b_pytest.log 2>&1
for v in default pf0 pf1 pf3 nc16 mb3 pfl1; do
  if [ $v = default ]; then unset SS2_LIB; else export SS2_LIB=$PWD/profiles/exp/libss2_$v.so; fi
  timeout 120 python profiles/warp_bench.py --tag $v >> gpurun_out/r2b_sweep.jsonl 2>> gpurun_out/r2b_sweep.err
done
unset SS2_LIB
SS2_TPS_L3=0 timeout 120 python profiles/warp_bench.py --tag old_lattice >> gpurun_out/r2b_sweep.jsonl 2>> gpurun_out/r2b_sweep.err
timeout 120 python profiles/warp_bench.py --tag default_1080 --height 1080 --width 1920 --frames 16 >> gpurun_out/r2b_sweep.jsonl 2>> gpurun_out/r2b_sweep.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tps_warp_lat3|tps_nodes|tps_solve" -c 3 -o gpurun_out/r2b_lat3 python profiles/warp_bench.py --iters 1 > gpurun_out/r2b_ncu.log 2>&1
timeout 300 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -3 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_sweep.jsonl; tail -c 600 gpurun_out/r2b_bench.json
