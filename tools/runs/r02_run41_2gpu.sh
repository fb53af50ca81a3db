#!/bin/bash
# round 2, GPU call 41 (2 GPUs): compute-sanitizer memcheck / synccheck over the peer-memory exchange kernel test
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout -k 10 240 compute-sanitizer --tool $tool --target-processes all --print-limit 20 \
      python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k kernel_exact --timeout 200 -p no:cacheprovider > gpurun_out/r02_run41_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r02_run41_$tool.log
done
grep -h "ERROR SUMMARY\|passed\|failed\|exit\|not supported\|rror" gpurun_out/r02_run41_memcheck.log gpurun_out/r02_run41_synccheck.log | sort | uniq -c | head -20
