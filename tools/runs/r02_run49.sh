#!/bin/bash
# round 2, GPU call 49 (1 GPU): four-warps-per-tile duration decoder for small inference calls -- kernel test, sampling / decode tests, timings
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -s -p no:cacheprovider \
   -k "dur_decoder or greedy_pass_fused or batched_sampling or packed_scheduled or training_matches_reference_golden" > gpurun_out/r02_run49_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_run49_tests.log
timeout -k 10 600 python tools/ab_free_running.py "" "" > gpurun_out/r02_run49_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run49_ab.log
grep -h "token match\|passed\|failed\|exit\|rror" gpurun_out/r02_run49_tests.log | tail -8; cat gpurun_out/r02_run49_ab.log
