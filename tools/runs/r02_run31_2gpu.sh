#!/bin/bash
# round 2, GPU call 31 (2 GPUs): exchange kernel v2 (per-bucket flag slots, several exchange streams, fences by the signalling threads only)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
timeout -k 10 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k kernel_exact --timeout 280 -p no:cacheprovider > gpurun_out/r02_run31_kernel.log 2>&1
echo "exit $?" >> gpurun_out/r02_run31_kernel.log
timeout -k 10 300 $TR tools/ar_bench.py > gpurun_out/r02_run31_ar.log 2>&1
echo "exit $?" >> gpurun_out/r02_run31_ar.log
timeout -k 10 600 $TR tools/ddp_trace.py --trace gpurun_out/r02_run31_trace.json "impl=p2p,bucket_mb=8" > gpurun_out/r02_run31_trace.log 2>&1
echo "exit $?" >> gpurun_out/r02_run31_trace.log
timeout -k 10 600 $TR tools/ddp_trace.py "impl=p2p,bucket_mb=8,streams=1" "impl=p2p,bucket_mb=8,streams=2,blocks=64" "impl=p2p,bucket_mb=16" "impl=p2p,bucket_mb=4" "impl=p2p,bucket_mb=8,streams=8,blocks=16" > gpurun_out/r02_run31_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run31_ab.log
grep -h "world\|trace:\|exit\|rror\|passed\|failed\|p2p\|nccl" gpurun_out/r02_run31_kernel.log gpurun_out/r02_run31_ar.log gpurun_out/r02_run31_trace.log gpurun_out/r02_run31_ab.log | grep -v "exchange  start" | tail -60
