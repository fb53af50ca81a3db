#!/bin/bash
# round 2, GPU call 50 (1 GPU): 32-unit tiles of the x-folded fused step for small row counts -- kernel test, sampling tests, timings
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_packed.py -m gpu -q -s -p no:cacheprovider \
   -k "tmax or greedy_pass_fused or batched_sampling or packed_loss_mode" > gpurun_out/r02_run50_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_run50_tests.log
timeout -k 10 600 python tools/ab_free_running.py "" "" > gpurun_out/r02_run50_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run50_ab.log
grep -h "token match\|passed\|failed\|exit\|rror" gpurun_out/r02_run50_tests.log | tail -8; cat gpurun_out/r02_run50_ab.log
