#!/bin/bash
# round 2, GPU call 18: summary-path row predicate -- kernel test, packed parity tests, A/B, bench, ncu traffic of the timed kernels
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_packed.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r02_run18_kernels.log 2>&1
echo "exit $?" >> gpurun_out/r02_run18_kernels.log
timeout -k 10 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -s -k "packed or batch512" > gpurun_out/r02_run18_model.log 2>&1
echo "exit $?" >> gpurun_out/r02_run18_model.log
timeout -k 10 300 python tools/step_ab.py "" > gpurun_out/r02_run18_ab.log 2>&1
timeout -k 10 900 python bench.py --no-decode-e2e > gpurun_out/r02_run18_bench.json 2> gpurun_out/r02_run18_bench.err
echo "exit $?" >> gpurun_out/r02_run18_bench.err
timeout -k 10 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_kernels \
    python tools/profile_kernels_r02.py > gpurun_out/r02_run18_ncu2.log 2>&1
python tools/ncu_traffic.py gpurun_out/r02_kernels.ncu-rep gpurun_out/r02_kernel_traffic.json > gpurun_out/r02_ncu_full_kernels.md 2>&1
timeout -k 10 300 python tools/trace_step.py gpurun_out/trace_packed3.json > gpurun_out/r02_run18_trace.log 2>&1
tail -3 gpurun_out/r02_run18_kernels.log; grep -E "passed|failed|parity" gpurun_out/r02_run18_model.log | tail -5; grep "ms/step" gpurun_out/r02_run18_ab.log; head -c 300 gpurun_out/r02_run18_bench.json
