#!/bin/bash
# round 2, GPU call 2: full suite (nothing gated), fused-step A/B, full bench line, launch list, ncu --set full of the timed kernels
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run2_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run2_tests.log
timeout -k 10 300 python tools/gru_step_bench.py > gpurun_out/r02_run2_stepbench.log 2>&1
echo "exit $?" >> gpurun_out/r02_run2_stepbench.log
timeout -k 10 900 python bench.py > gpurun_out/r02_run2_bench.json 2> gpurun_out/r02_run2_bench.err
echo "exit $?" >> gpurun_out/r02_run2_bench.err
timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_v1.csv python tools/profile_step.py > gpurun_out/r02_run2_ncu1.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_kernels \
    python tools/profile_kernels_r02.py > gpurun_out/r02_run2_ncu2.log 2>&1
tail -3 gpurun_out/r02_run2_tests.log; cat gpurun_out/r02_run2_stepbench.log; tail -c 600 gpurun_out/r02_run2_bench.json
