#!/bin/bash
# round 2, GPU call 13: packed path follow-ups (slot-major slab, balanced embedding backward, resident summary GRU on sorted rows)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_packed.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r02_run13_kernels.log 2>&1
echo "exit $?" >> gpurun_out/r02_run13_kernels.log
timeout -k 10 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -s -k "packed or batch512" > gpurun_out/r02_run13_model.log 2>&1
echo "exit $?" >> gpurun_out/r02_run13_model.log
timeout -k 10 600 python tools/step_ab.py RESIDENT_GRU128_SORTED=0 RESIDENT_GRU128_SORTED=1 > gpurun_out/r02_run13_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run13_ab.log
timeout -k 10 300 python tools/trace_step.py gpurun_out/trace_packed2.json > gpurun_out/r02_run13_trace.log 2>&1
tail -5 gpurun_out/r02_run13_kernels.log; grep -E "passed|failed|parity|Error" gpurun_out/r02_run13_model.log | tail -8; grep -v Warn gpurun_out/r02_run13_ab.log | tail -5; tail -1 gpurun_out/r02_run13_trace.log
