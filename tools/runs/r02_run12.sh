#!/bin/bash
# round 2, GPU call 12: packed note level -- kernel tests, full suite, bench line, launch list
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_packed.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r02_run12_kernels.log 2>&1
echo "exit $?" >> gpurun_out/r02_run12_kernels.log
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_packed.py > gpurun_out/r02_run12_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run12_tests.log
timeout -k 10 900 python bench.py > gpurun_out/r02_run12_bench.json 2> gpurun_out/r02_run12_bench.err
echo "exit $?" >> gpurun_out/r02_run12_bench.err
timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_v3.csv python tools/profile_step.py > gpurun_out/r02_run12_ncu1.log 2>&1
timeout -k 10 300 python tools/trace_step.py gpurun_out/trace_packed.json > gpurun_out/r02_run12_trace.log 2>&1
tail -5 gpurun_out/r02_run12_kernels.log; tail -4 gpurun_out/r02_run12_tests.log; head -c 500 gpurun_out/r02_run12_bench.json; tail -2 gpurun_out/r02_run12_trace.log
