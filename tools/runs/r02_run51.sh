#!/bin/bash
# round 2, GPU call 51 (1 GPU): summary GRU of the greedy decode visits rows in sorted-length order -- tests, decode timing
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -s -p no:cacheprovider \
   -k "gru128 or greedy_tokens or large_batch_decode or graphed_decode or greedy_pass_fused or batched_sampling or persistent_decode" > gpurun_out/r02_run51_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_run51_tests.log
timeout -k 10 600 python bench.py --no-cpu --no-tfr0 --no-decode-e2e --steps 5 --warmup 3 > gpurun_out/r02_run51_bench.json 2> gpurun_out/r02_run51_bench.err
echo "exit $?" >> gpurun_out/r02_run51_bench.err
grep -h "token match\|passed\|failed\|exit\|rror" gpurun_out/r02_run51_tests.log | tail -8
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_run51_bench.json').read().strip().splitlines()[-1])
print({k: d['decode'][k] for k in ('value','ms_per_batch','tf32_value','fp32_ffma_value','latency_16_segments_ms')})
P
