#!/bin/bash
# round 2, GPU call 3: persistent small-batch decode kernel (first run, own timeout), fused-step variants, suite, bench
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 300 -s -k "persistent or greedy_tokens or odd_batch" \
    -p no:cacheprovider > gpurun_out/r02_run3_persist.log 2>&1
echo "exit $?" >> gpurun_out/r02_run3_persist.log
timeout -k 10 300 python tools/gru_step_bench.py > gpurun_out/r02_run3_stepbench.log 2>&1
echo "exit $?" >> gpurun_out/r02_run3_stepbench.log
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run3_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run3_tests.log
timeout -k 10 900 python bench.py --no-cpu > gpurun_out/r02_run3_bench.json 2> gpurun_out/r02_run3_bench.err
echo "exit $?" >> gpurun_out/r02_run3_bench.err
tail -5 gpurun_out/r02_run3_persist.log; cat gpurun_out/r02_run3_stepbench.log; tail -3 gpurun_out/r02_run3_tests.log; tail -c 1500 gpurun_out/r02_run3_bench.json
