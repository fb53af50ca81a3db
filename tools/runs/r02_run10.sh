#!/bin/bash
# round 2, GPU call 10: deferred weight gradients after the record_stream fix; background launch shapes A/B
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "graphed_train_step or batch512 or training_matches_reference_golden" > gpurun_out/r02_run10_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_run10_tests.log
timeout -k 10 900 python tools/step_ab.py DEFER_WGRAD=0 DEFER_WGRAD=1 DEFER_WGRAD=1,BG_GEMM_CFG=8000000 DEFER_WGRAD=1,BG_GEMM_CFG=8012832 \
   DEFER_WGRAD=1,BG_GEMM_CFG=16000000 DEFER_WGRAD=1,BG_GEMM_CFG=24012832 DEFER_WGRAD=1,BG_GEMM_CFG=8025622 DEFER_WGRAD=1,BG_GEMM_CFG=24000000 > gpurun_out/r02_run10_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run10_ab.log
tail -4 gpurun_out/r02_run10_tests.log; grep -v Warn gpurun_out/r02_run10_ab.log
