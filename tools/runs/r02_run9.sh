#!/bin/bash
# round 2, GPU call 9: deferred weight gradients (ops.defer) -- A/B timing of the captured step, parity at batch 512, full suite
mkdir -p gpurun_out
timeout -k 10 600 python tools/step_ab.py DEFER_WGRAD=0 DEFER_WGRAD=1 "DEFER_WGRAD=1,DEFER_MIN_ROWS=100000" > gpurun_out/r02_run9_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run9_ab.log
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run9_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run9_tests.log
timeout -k 10 600 python bench.py --no-cpu --no-decode-e2e --no-tfr0 > gpurun_out/r02_run9_bench.json 2> gpurun_out/r02_run9_bench.err
echo "exit $?" >> gpurun_out/r02_run9_bench.err
cat gpurun_out/r02_run9_ab.log; tail -5 gpurun_out/r02_run9_tests.log; head -c 400 gpurun_out/r02_run9_bench.json
