#!/bin/bash
# round 2, GPU call 1: the full -m gpu suite incl. the tests that were gated in round 1 (each risky piece under its own timeout)
mkdir -p gpurun_out
export POLYDIS_TEST_EXPERIMENTAL=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_run1_smi.txt 2>&1
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -k "not gru_step_tma and not device_plan and not select_rows" \
    -p no:cacheprovider > gpurun_out/r02_run1_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run1_tests.log
timeout -k 10 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 150 -k "select_rows" -p no:cacheprovider > gpurun_out/r02_run1_select.log 2>&1
echo "exit $?" >> gpurun_out/r02_run1_select.log
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 200 -k "gru_step_tma" -p no:cacheprovider > gpurun_out/r02_run1_tma.log 2>&1
echo "exit $?" >> gpurun_out/r02_run1_tma.log
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 250 -k "device_plan" -p no:cacheprovider > gpurun_out/r02_run1_plan.log 2>&1
echo "exit $?" >> gpurun_out/r02_run1_plan.log
timeout -k 10 200 python tools/gru_step_bench.py > gpurun_out/r02_run1_stepbench.log 2>&1
echo "exit $?" >> gpurun_out/r02_run1_stepbench.log
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu --batched-sampling > gpurun_out/r02_run1_bench_bs.log 2>&1
echo "exit $?" >> gpurun_out/r02_run1_bench_bs.log
tail -3 gpurun_out/r02_run1_tests.log
