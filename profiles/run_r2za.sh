#!/bin/bash
# round-2 GPU session ZA: barrier-free solve (step counter in shared memory), packed nodes kernel: timing, ncu --set full of
# solve + nodes, TPS-related tests
mkdir -p gpurun_out
SS2_LIB=$PWD/profiles/exp/libss2_old.so python profiles/warp_bench.py --tag old > gpurun_out/r2za_sweep.jsonl 2> gpurun_out/r2za_sweep.err
python profiles/warp_bench.py --tag new >> gpurun_out/r2za_sweep.jsonl 2>> gpurun_out/r2za_sweep.err
cat gpurun_out/r2za_sweep.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2za_launches.csv python profiles/warp_bench.py --iters 2 > gpurun_out/r2za_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/r2za_launches.csv 2>&1 | head -n 8 | grep -v "at::"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tps_solve_kernel|tps_nodes_kernel" -s 4 -c 2 -o gpurun_out/r2za_solve_nodes python profiles/warp_bench.py --iters 1 > gpurun_out/r2za_ncu_full.log 2>&1
ls -la gpurun_out/r2za_solve_nodes.ncu-rep
( timeout 900 python -m pytest tests -m gpu -q -x -k "tps or fullsize or stream_golden or stable or three_view or nview or linear or dropin" 2>&1 | tail -n 5 ) > gpurun_out/r2za_pytest.log 2>&1
tail -n 3 gpurun_out/r2za_pytest.log
