#!/bin/bash
# round 2, GPU call 29 (2 GPUs): the peer-memory all-reduce kernel -- A/B against NCCL, timeline, 2-rank parity test
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout -k 10 600 $TR tools/ddp_trace.py --trace gpurun_out/r02_run29_trace.json "impl=p2p,bucket_mb=8" > gpurun_out/r02_run29_trace.log 2>&1
echo "exit $?" >> gpurun_out/r02_run29_trace.log
timeout -k 10 600 $TR tools/ddp_trace.py "impl=p2p,bucket_mb=32" "impl=p2p,bucket_mb=16,ready=1" "impl=p2p,bucket_mb=4" "impl=nccl,bucket_mb=32" > gpurun_out/r02_run29_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run29_ab.log
timeout -k 10 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k p2p --timeout 800 -p no:cacheprovider > gpurun_out/r02_run29_multi.log 2>&1
echo "exit $?" >> gpurun_out/r02_run29_multi.log
grep -h "world\|trace:\|exit\|rror\|2-rank\|passed\|failed" gpurun_out/r02_run29_trace.log gpurun_out/r02_run29_ab.log gpurun_out/r02_run29_multi.log | tail -30
