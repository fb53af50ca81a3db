#!/bin/bash
# round 2, GPU call 43 (1 GPU): fused TF32 note-GRU slot in the training-time greedy pass -- test + free-running step timing
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s -p no:cacheprovider -k "greedy_pass_fused or batched_sampling or packed_scheduled" > gpurun_out/r02_run43_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_run43_tests.log
timeout -k 10 600 python tools/time_sched_sampling.py 512 > gpurun_out/r02_run43_ss.log 2>&1
echo "exit $?" >> gpurun_out/r02_run43_ss.log
grep -h "token match\|passed\|failed\|exit\|rror" gpurun_out/r02_run43_tests.log | tail -8; cat gpurun_out/r02_run43_ss.log | tail -5
