#!/bin/bash
# round 2, GPU call 42 (1 GPU): parallel pack_order, length-aware colsum for the summariser -- kernel + model parity tests, step timing
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider \
    -k "pack_order or colsum or packed_loss_mode or baseline_batch512 or oracle_full_gradients" > gpurun_out/r02_run42_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_run42_tests.log
timeout -k 10 600 python tools/step_ab.py "" "" "" > gpurun_out/r02_run42_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run42_ab.log
tail -3 gpurun_out/r02_run42_tests.log; cat gpurun_out/r02_run42_ab.log
