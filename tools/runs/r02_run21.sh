#!/bin/bash
# round 2, GPU call 21: texture backward without shared atomics; full suite; bench; trace
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run21_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run21_tests.log
timeout -k 10 300 python tools/step_ab.py "" > gpurun_out/r02_run21_ab.log 2>&1
timeout -k 10 300 python tools/trace_step.py gpurun_out/trace_packed5.json > gpurun_out/r02_run21_trace.log 2>&1
tail -4 gpurun_out/r02_run21_tests.log; grep "ms/step" gpurun_out/r02_run21_ab.log; tail -1 gpurun_out/r02_run21_trace.log
