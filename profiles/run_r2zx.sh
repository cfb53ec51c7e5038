#!/bin/bash
# round-2 GPU session ZX: per-layer timing of the kind::f16 direct kernel: decomposition (MMA only / TMA only / no stores) and tile plans
mkdir -p gpurun_out
SS2_CONV_TEST_F16=3 timeout -s KILL 200 python profiles/conv_bench.py dbg 5 > gpurun_out/r2zx_conv_f16_dbg.jsonl 2> gpurun_out/r2zx_dbg.err
SS2_CONV_TEST_F16=3 timeout -s KILL 300 python profiles/conv_bench.py plans 5 > gpurun_out/r2zx_conv_f16_plans.jsonl 2> gpurun_out/r2zx_plans.err
timeout -s KILL 200 python profiles/conv_bench.py dbg 5 > gpurun_out/r2zx_conv_tf32_dbg.jsonl 2>> gpurun_out/r2zx_dbg.err
cat gpurun_out/r2zx_conv_f16_dbg.jsonl | head -60
tail -3 gpurun_out/r2zx_dbg.err
