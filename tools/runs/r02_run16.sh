#!/bin/bash
# round 2, GPU call 16: packed scheduled sampling test, full bench line, launch list v4, ncu --set full of the timed kernels
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -k "packed or batched_sampling or graphed" > gpurun_out/r02_run16_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_run16_tests.log
timeout -k 10 900 python bench.py > gpurun_out/r02_run16_bench.json 2> gpurun_out/r02_run16_bench.err
echo "exit $?" >> gpurun_out/r02_run16_bench.err
timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_v4.csv python tools/profile_step.py > gpurun_out/r02_run16_ncu1.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_kernels \
    python tools/profile_kernels_r02.py > gpurun_out/r02_run16_ncu2.log 2>&1
python tools/ncu_traffic.py gpurun_out/r02_kernels.ncu-rep gpurun_out/r02_kernel_traffic.json > gpurun_out/r02_ncu_full_kernels.md 2>&1
tail -4 gpurun_out/r02_run16_tests.log; head -c 400 gpurun_out/r02_run16_bench.json; tail -2 gpurun_out/r02_run16_bench.err
