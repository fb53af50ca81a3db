#!/bin/bash
# round 2, GPU call 22: texture front end with saved arg-max; fused optimizer A/B
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -k "texture or batch512 or training_matches or graphed" > gpurun_out/r02_run22_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_run22_tests.log
timeout -k 10 300 python tools/step_ab.py "" > gpurun_out/r02_run22_ab.log 2>&1
PD_AB_FUSED_OPT=1 timeout -k 10 300 python tools/step_ab.py "" >> gpurun_out/r02_run22_ab.log 2>&1
tail -4 gpurun_out/r02_run22_tests.log; grep "ms/step" gpurun_out/r02_run22_ab.log
