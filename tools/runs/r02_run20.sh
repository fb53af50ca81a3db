#!/bin/bash
# round 2, GPU call 20: bf16 operands for the batch-sized recurrences -- kernel tests, batch-512 parity vs the oracle, A/B
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -k "bf16 or gates_bwd" > gpurun_out/r02_run20_kernels.log 2>&1
echo "exit $?" >> gpurun_out/r02_run20_kernels.log
timeout -k 10 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -s -k "packed or batch512 or training_matches or graphed" > gpurun_out/r02_run20_model.log 2>&1
echo "exit $?" >> gpurun_out/r02_run20_model.log
timeout -k 10 300 python tools/step_ab.py BF16_RECURRENT=0 BF16_RECURRENT=1 > gpurun_out/r02_run20_ab.log 2>&1
tail -4 gpurun_out/r02_run20_kernels.log; grep -E "passed|failed|parity|Error" gpurun_out/r02_run20_model.log | tail -6; grep "ms/step" gpurun_out/r02_run20_ab.log
