#!/bin/bash
# round 2, GPU call 5: decode launch list, compute-sanitizer passes over the round-2 kernels, final ncu captures
mkdir -p gpurun_out
timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_decode.csv python tools/profile_decode.py > gpurun_out/r02_run5_ncu_dec.log 2>&1
# memcheck: kernel tests (incl. fused TMA step, persistent GEMM routes, new misc kernels) + the persistent decode kernel
timeout -k 10 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_rows.py \
    -x -q -m gpu -p no:cacheprovider -k "not graphed and not persistent_routes" > gpurun_out/r02_memcheck_kernels.log 2>&1
echo "exit $?" >> gpurun_out/r02_memcheck_kernels.log
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_model.py \
    -x -q -m gpu -p no:cacheprovider -k "greedy_tokens_match_reference_golden or persistent_decode" > gpurun_out/r02_memcheck_persistent.log 2>&1
echo "exit $?" >> gpurun_out/r02_memcheck_persistent.log
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py \
    -x -q -m gpu -p no:cacheprovider -k "persistent_routes" > gpurun_out/r02_memcheck_gemm_routes.log 2>&1
echo "exit $?" >> gpurun_out/r02_memcheck_gemm_routes.log
timeout -k 10 900 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_kernels.py tests/test_gpu_rows.py -q -m gpu \
    -p no:cacheprovider -k "not graphed and not persistent_routes" > gpurun_out/r02_synccheck.log 2>&1
echo "exit $?" >> gpurun_out/r02_synccheck.log
timeout -k 10 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider \
    -k "greedy_tokens_match_reference_golden and w1 and tf32x3" > gpurun_out/r02_racecheck_persistent.log 2>&1
echo "exit $?" >> gpurun_out/r02_racecheck_persistent.log
for f in gpurun_out/r02_memcheck_kernels.log gpurun_out/r02_memcheck_persistent.log gpurun_out/r02_memcheck_gemm_routes.log gpurun_out/r02_synccheck.log gpurun_out/r02_racecheck_persistent.log; do echo "== $f"; grep -E "SUMMARY|passed|failed|exit" $f | tail -4; done
