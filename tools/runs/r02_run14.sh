#!/bin/bash
# round 2, GPU call 14: 32-unit tiles of the fused step for the batch-sized recurrences -- kernel test, A/B
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -k "gru_step_tma" > gpurun_out/r02_run14_kernels.log 2>&1
echo "exit $?" >> gpurun_out/r02_run14_kernels.log
timeout -k 10 300 python tools/gru_step_bench.py > gpurun_out/r02_run14_stepbench.log 2>&1
timeout -k 10 600 python tools/step_ab.py FUSED_GRU_STEP_TMA_MIN_ROWS=4096 FUSED_GRU_STEP_TMA_MIN_ROWS=256 > gpurun_out/r02_run14_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run14_ab.log
tail -3 gpurun_out/r02_run14_kernels.log; cat gpurun_out/r02_run14_stepbench.log | tail -6; grep -v Warn gpurun_out/r02_run14_ab.log | tail -4
