#!/bin/bash
# round 2, GPU call 19: summariser backward from an (R,H) gradient (no zero-filled (R,T,H) tensors) -- tests + A/B
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_packed.py tests/test_gpu_rows.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r02_run19_kernels.log 2>&1
echo "exit $?" >> gpurun_out/r02_run19_kernels.log
timeout -k 10 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -s -k "packed or batch512 or training_matches" > gpurun_out/r02_run19_model.log 2>&1
echo "exit $?" >> gpurun_out/r02_run19_model.log
timeout -k 10 300 python tools/step_ab.py "" PACKED_NOTES=0 > gpurun_out/r02_run19_ab.log 2>&1
timeout -k 10 300 python tools/trace_step.py gpurun_out/trace_packed4.json > gpurun_out/r02_run19_trace.log 2>&1
tail -3 gpurun_out/r02_run19_kernels.log; grep -E "passed|failed|parity" gpurun_out/r02_run19_model.log | tail -5; grep "ms/step" gpurun_out/r02_run19_ab.log; tail -1 gpurun_out/r02_run19_trace.log
