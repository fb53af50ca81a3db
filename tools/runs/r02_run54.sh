#!/bin/bash
# round 2, GPU call 54 (1 GPU): per-sequence loop of the length-aware colsum -- test + ncu time vs the dense kernel
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -k "colsum or packed_loss_mode" > gpurun_out/r02_run54_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_run54_tests.log
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --profile-from-start off -k regex:colsum --csv \
    --log-file gpurun_out/r02_run54_colsum.csv python tools/profile_kernels_r02b.py > gpurun_out/r02_run54_ncu.log 2>&1
tail -2 gpurun_out/r02_run54_tests.log; grep -v "^==" gpurun_out/r02_run54_colsum.csv | cut -d, -f5,13-15 | tail -5
