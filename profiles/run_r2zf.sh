#!/bin/bash
# round-2 GPU session ZF: compute-sanitizer memcheck + racecheck over the new solve / nodes / fused-u8 kernels
mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tps_point_golden or tps_warp_golden or u8_fused_store_bit_identical and NORMAL or fullsize_frame_vs_oracle_and_arbiter and lattice-720" 2>&1 | tail -n 12 ) > gpurun_out/r2zf_memcheck.log 2>&1
tail -n 4 gpurun_out/r2zf_memcheck.log
( timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tps_point_golden" 2>&1 | tail -n 40 ) > gpurun_out/r2zf_racecheck.log 2>&1
tail -n 25 gpurun_out/r2zf_racecheck.log
