#!/bin/bash
# round-2 GPU session Y: solve kernel with rows across lanes (one barrier per step), packed nodes kernel: A/B against the
# previous tps.cu (profiles/exp/libss2_old.so built from the parent commit), launch times, full GPU test suite
mkdir -p gpurun_out
SS2_LIB=$PWD/profiles/exp/libss2_old.so python profiles/warp_bench.py --tag old > gpurun_out/r2y_warp_old.json 2> gpurun_out/r2y_warp_old.err
python profiles/warp_bench.py --tag new > gpurun_out/r2y_warp_new.json 2> gpurun_out/r2y_warp_new.err
cat gpurun_out/r2y_warp_old.json gpurun_out/r2y_warp_new.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2y_warp_launches.csv python profiles/warp_bench.py --iters 2 > gpurun_out/r2y_ncu_warp.log 2>&1
python profiles/launch_summary.py gpurun_out/r2y_warp_launches.csv > gpurun_out/r2y_warp_launches_summary.txt 2>&1
head -n 12 gpurun_out/r2y_warp_launches_summary.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 15 ) > gpurun_out/r2y_pytest.log 2>&1
tail -n 8 gpurun_out/r2y_pytest.log
