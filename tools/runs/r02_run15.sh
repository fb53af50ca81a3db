#!/bin/bash
# round 2, GPU call 15: full suite after the packed-path restriction + zero-folding + coalesced grid_prepare; small-GEMM tile A/B
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run15_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run15_tests.log
timeout -k 10 900 python tools/step_ab.py "" SMALL_GEMM_CFG=6442 SMALL_GEMM_CFG=6433 SMALL_GEMM_CFG=12823 SMALL_GEMM_CFG=12832 FUSED_GRU_STEP_TMA_MIN_ROWS=4096 > gpurun_out/r02_run15_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run15_ab.log
tail -4 gpurun_out/r02_run15_tests.log; grep -v Warn gpurun_out/r02_run15_ab.log | tail -8
