#!/bin/bash
# round 2, GPU call 24: compute-sanitizer over the round's new kernels (memcheck / synccheck / racecheck)
mkdir -p gpurun_out
timeout -k 10 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_packed.py tests/test_gpu_kernels.py \
    -x -q -m gpu -p no:cacheprovider -k "packed or rows or gru_step_tma or bf16 or texture or gates_bwd or pack_order or grid_prepare or dur_decode or gemm_rows" > gpurun_out/r02b_memcheck_kernels.log 2>&1
echo "exit $?" >> gpurun_out/r02b_memcheck_kernels.log
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_model.py \
    -x -q -m gpu -p no:cacheprovider -k "packed_loss_mode_matches_oracle or packed_scheduled" > gpurun_out/r02b_memcheck_model.log 2>&1
echo "exit $?" >> gpurun_out/r02b_memcheck_model.log
timeout -k 10 900 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_packed.py -q -m gpu -p no:cacheprovider > gpurun_out/r02b_synccheck.log 2>&1
echo "exit $?" >> gpurun_out/r02b_synccheck.log
timeout -k 10 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_packed.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider \
    -k "pack_order or row_reductions or texture or dur_decode_rows" > gpurun_out/r02b_racecheck.log 2>&1
echo "exit $?" >> gpurun_out/r02b_racecheck.log
for f in gpurun_out/r02b_memcheck_kernels.log gpurun_out/r02b_memcheck_model.log gpurun_out/r02b_synccheck.log gpurun_out/r02b_racecheck.log; do echo "== $f"; grep -E "SUMMARY|passed|failed|exit" $f | tail -4; done
