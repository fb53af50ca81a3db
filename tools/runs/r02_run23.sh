#!/bin/bash
# round 2, GPU call 23: fork_join side streams keyed by parent (allocator-safety fix) -- stress of the first-replay race, full suite, A/B
mkdir -p gpurun_out
timeout -k 10 300 python tools/debug_flaky.py "" "" "" > gpurun_out/r02_run23_flaky.log 2>&1
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run23_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run23_tests.log
timeout -k 10 300 python tools/step_ab.py "" > gpurun_out/r02_run23_ab.log 2>&1
PD_AB_FUSED_OPT=1 timeout -k 10 300 python tools/step_ab.py "" >> gpurun_out/r02_run23_ab.log 2>&1
grep "^\[" gpurun_out/r02_run23_flaky.log | cut -c1-250; tail -4 gpurun_out/r02_run23_tests.log; grep "ms/step" gpurun_out/r02_run23_ab.log
