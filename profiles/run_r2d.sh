#!/bin/bash
# round-2 GPU session D: reverted resampler + register solve v2, uint8 host edges, new bench legs
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2d_pytest.log 2>&1
timeout 120 python profiles/warp_bench.py --tag final >> gpurun_out/r2d_sweep.jsonl 2>> gpurun_out/r2d_sweep.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tps_warp_lattice|tps_nodes|tps_solve|stable_meshes" -c 4 -o gpurun_out/r2d_warp python profiles/warp_bench.py --iters 1 > gpurun_out/r2d_ncu.log 2>&1
timeout 600 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "host_edges_u8_bit_exact and 96 or tps_point_golden or get_stable_sqe or stream_host_u8" > gpurun_out/r2d_sanitizer_memcheck.log 2>&1
tail -8 gpurun_out/r2d_pytest.log; cat gpurun_out/r2d_sweep.jsonl; tail -c 3000 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err; tail -5 gpurun_out/r2d_sanitizer_memcheck.log
