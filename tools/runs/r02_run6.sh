#!/bin/bash
# round 2, GPU call 6: split3 fused decode step (first run), suite, bench, final launch list + ncu --set full captures
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 500 -s -p no:cacheprovider \
    -k "tma3 or large_batch_decode or graphed_decode or greedy_tokens_match_oracle" > gpurun_out/r02_run6_tma3.log 2>&1
echo "exit $?" >> gpurun_out/r02_run6_tma3.log
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run6_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run6_tests.log
timeout -k 10 900 python bench.py --no-cpu > gpurun_out/r02_run6_bench.json 2> gpurun_out/r02_run6_bench.err
echo "exit $?" >> gpurun_out/r02_run6_bench.err
timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_decode_v2.csv python tools/profile_decode.py > gpurun_out/r02_run6_ncu_dec.log 2>&1
tail -4 gpurun_out/r02_run6_tma3.log; tail -3 gpurun_out/r02_run6_tests.log; tail -c 1600 gpurun_out/r02_run6_bench.json
