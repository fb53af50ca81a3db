#!/bin/bash
# round 2, GPU call 35 (1 GPU): clip folded into fused Adam (test + A/B), the largest stray aten kernels of the step
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "graphed_train_step" -p no:cacheprovider > gpurun_out/r02_run35_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_run35_tests.log
timeout -k 10 300 python tools/find_big_aten.py 15 > gpurun_out/r02_run35_aten.log 2>&1
echo "exit $?" >> gpurun_out/r02_run35_aten.log
timeout -k 10 300 python tools/step_ab.py "" "FOLD=0" > gpurun_out/r02_run35_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run35_ab.log
tail -3 gpurun_out/r02_run35_tests.log; cat gpurun_out/r02_run35_ab.log; head -70 gpurun_out/r02_run35_aten.log
