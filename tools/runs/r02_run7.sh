#!/bin/bash
# round 2, GPU call 7: x-projection folded into the fused training step (pd_gru_step_tmax) -- kernel tests, A/B timing, suite, bench
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 700 -s -p no:cacheprovider \
    -k "tmax or tma3 or batch512 or (training_matches_reference_golden and tf111)" > gpurun_out/r02_run7_tmax.log 2>&1
echo "exit $?" >> gpurun_out/r02_run7_tmax.log
timeout -k 10 300 python tools/gru_step_bench.py > gpurun_out/r02_run7_stepbench.log 2>&1
echo "exit $?" >> gpurun_out/r02_run7_stepbench.log
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run7_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run7_tests.log
timeout -k 10 900 python bench.py --no-cpu --no-decode-e2e > gpurun_out/r02_run7_bench.json 2> gpurun_out/r02_run7_bench.err
echo "exit $?" >> gpurun_out/r02_run7_bench.err
grep -E "passed|failed|parity" gpurun_out/r02_run7_tmax.log | tail -4; cat gpurun_out/r02_run7_stepbench.log; tail -3 gpurun_out/r02_run7_tests.log; head -c 700 gpurun_out/r02_run7_bench.json
