#!/bin/bash
# round 2, GPU call 46 (1 GPU): fused pick+embed tail of the greedy slot -- kernel test, decode / sampling parity tests, timings
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -s -p no:cacheprovider \
   -k "greedy_pick or greedy_tokens or large_batch_decode or graphed_decode or greedy_pass_fused or batched_sampling or persistent_decode" > gpurun_out/r02_run46_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_run46_tests.log
timeout -k 10 600 python tools/ab_free_running.py "" "" > gpurun_out/r02_run46_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run46_ab.log
grep -h "token match\|passed\|failed\|exit\|rror" gpurun_out/r02_run46_tests.log | tail -8; cat gpurun_out/r02_run46_ab.log
