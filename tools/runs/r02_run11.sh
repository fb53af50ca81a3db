#!/bin/bash
# round 2, GPU call 11: packed note level -- kernel tests, golden + batch-512 oracle parity, A/B timing
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_packed.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/r02_run11_kernels.log 2>&1
echo "exit $?" >> gpurun_out/r02_run11_kernels.log
timeout -k 10 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -s -k "packed or batch512" > gpurun_out/r02_run11_model.log 2>&1
echo "exit $?" >> gpurun_out/r02_run11_model.log
timeout -k 10 600 python tools/step_ab.py PACKED_NOTES=0 PACKED_NOTES=1 PACKED_NOTES=1,DEFER_WGRAD=0 > gpurun_out/r02_run11_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run11_ab.log
tail -5 gpurun_out/r02_run11_kernels.log; grep -E "passed|failed|parity|Error" gpurun_out/r02_run11_model.log | tail -8; grep -v Warn gpurun_out/r02_run11_ab.log | tail -5
