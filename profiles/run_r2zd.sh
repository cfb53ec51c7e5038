#!/bin/bash
# round-2 GPU session ZD: two-level lattice nodes (coarse exact lattice + interpolation + difference terms): timing against the
# one-level evaluation (SS2_TPS_NODES1=1), difference of the fused frames, TPS tests (arbiter comparisons print their errors)
mkdir -p gpurun_out
python profiles/warp_bench.py --tag two_level > gpurun_out/r2zd_sweep.jsonl 2> gpurun_out/r2zd_sweep.err
SS2_TPS_NODES1=1 python profiles/warp_bench.py --tag one_level >> gpurun_out/r2zd_sweep.jsonl 2>> gpurun_out/r2zd_sweep.err
cat gpurun_out/r2zd_sweep.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2zd_launches.csv python profiles/warp_bench.py --iters 2 > gpurun_out/r2zd_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/r2zd_launches.csv 2>&1 | head -n 9 | grep -v "at::"
( timeout 900 python -m pytest tests -m gpu -q -x -s -k "tps or fullsize or stream_golden or stable or three_view or nview or linear or dropin" 2>&1 | tail -n 60 ) > gpurun_out/r2zd_pytest.log 2>&1
tail -n 3 gpurun_out/r2zd_pytest.log
grep -i "arbiter\|coord\|px" gpurun_out/r2zd_pytest.log | head -20
