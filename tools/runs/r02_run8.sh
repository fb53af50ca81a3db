#!/bin/bash
# round 2, GPU call 8 (first of the second session): HEAD validation -- full suite, bench line, launch list of the step,
# ncu --set full of the timed kernel variants (-> profiles/r02_kernel_traffic.json)
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run8_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run8_tests.log
timeout -k 10 900 python bench.py > gpurun_out/r02_run8_bench.json 2> gpurun_out/r02_run8_bench.err
echo "exit $?" >> gpurun_out/r02_run8_bench.err
timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_v2.csv python tools/profile_step.py > gpurun_out/r02_run8_ncu1.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_kernels \
    python tools/profile_kernels_r02.py > gpurun_out/r02_run8_ncu2.log 2>&1
python tools/ncu_traffic.py gpurun_out/r02_kernels.ncu-rep gpurun_out/r02_kernel_traffic.json > gpurun_out/r02_ncu_full_kernels.md 2>&1
tail -3 gpurun_out/r02_run8_tests.log; tail -c 1500 gpurun_out/r02_run8_bench.json; tail -3 gpurun_out/r02_run8_ncu2.log
