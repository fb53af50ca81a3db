#!/bin/bash
# round 2, GPU call 36 (1 GPU): 64-unit tiles of the bf16 fused step for the concurrent short recurrences -- kernel test + step A/B
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "bf16_operands" -p no:cacheprovider > gpurun_out/r02_run36_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_run36_tests.log
timeout -k 10 600 python tools/step_ab.py "" "BF16_STEP_UNITS_SHORT=64" "BF16_STEP_UNITS_SHORT=64,BF16_STEP_SHORT_T=32" "" "BF16_STEP_UNITS_SHORT=64" > gpurun_out/r02_run36_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run36_ab.log
tail -3 gpurun_out/r02_run36_tests.log; cat gpurun_out/r02_run36_ab.log
