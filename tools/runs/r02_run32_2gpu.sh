#!/bin/bash
# round 2, GPU call 32 (2 GPUs): exchange kernel with 256-thread blocks (co-resident with the chain's CTAs); bucket / stream sweep
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516"
timeout -k 10 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k kernel_exact --timeout 280 -p no:cacheprovider > gpurun_out/r02_run32_kernel.log 2>&1
echo "exit $?" >> gpurun_out/r02_run32_kernel.log
timeout -k 10 600 $TR tools/ddp_trace.py "impl=p2p,bucket_mb=8" "impl=p2p,bucket_mb=16" "impl=p2p,bucket_mb=16,ready=1" "impl=p2p,bucket_mb=32" "impl=p2p,bucket_mb=16,blocks=16,streams=8" "impl=p2p,bucket_mb=16,prio=-1" "impl=p2p,bucket_mb=16,blocks=64,streams=2" > gpurun_out/r02_run32_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run32_ab.log
timeout -k 10 300 $TR tools/ar_bench.py > gpurun_out/r02_run32_ar.log 2>&1
echo "exit $?" >> gpurun_out/r02_run32_ar.log
grep -h "world\|trace:\|exit\|rror\|passed\|failed\|p2p\|nccl" gpurun_out/r02_run32_kernel.log gpurun_out/r02_run32_ab.log gpurun_out/r02_run32_ar.log | grep -v "exchange  start" | tail -60
