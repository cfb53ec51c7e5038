#!/bin/bash
# round-2 GPU session ZQ: CTA-pair (cta_group::2) direct 3x3 kernel: first correctness run
mkdir -p gpurun_out
( SS2_DC_PAIR=1 timeout -s KILL 150 python -m pytest tests -m gpu -q -x -k "conv_dc_every_tile_plan or conv_kernels_vs_torch" 2>&1 | tail -n 25 ) > gpurun_out/r2zq_pytest_conv.log 2>&1
tail -n 25 gpurun_out/r2zq_pytest_conv.log
nvidia-smi --query-gpu=name,memory.used --format=csv
