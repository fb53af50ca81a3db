#!/bin/bash
# round 2, GPU call 26: final-state validation -- full suite, bench line, launch list v5, ncu --set full of the timed kernels, trace
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run26_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run26_tests.log
timeout -k 10 900 python bench.py > gpurun_out/r02_run26_bench.json 2> gpurun_out/r02_run26_bench.err
echo "exit $?" >> gpurun_out/r02_run26_bench.err
timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_v5.csv python tools/profile_step.py > gpurun_out/r02_run26_ncu1.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_kernels \
    python tools/profile_kernels_r02.py > gpurun_out/r02_run26_ncu2.log 2>&1
python tools/ncu_traffic.py gpurun_out/r02_kernels.ncu-rep gpurun_out/r02_kernel_traffic.json > gpurun_out/r02_ncu_full_kernels.md 2>&1
timeout -k 10 300 python tools/trace_step.py gpurun_out/trace_packed6.json > gpurun_out/r02_run26_trace.log 2>&1
timeout -k 10 300 python tools/replay_consistency.py "" > gpurun_out/r02_run26_consistency.log 2>&1
tail -4 gpurun_out/r02_run26_tests.log; head -c 300 gpurun_out/r02_run26_bench.json; tail -1 gpurun_out/r02_run26_trace.log; tail -3 gpurun_out/r02_run26_ncu2.log
