#!/bin/bash
# round 2, GPU call 28 (2 GPUs): where the 2-GPU step loses 0.77 ms -- reducer A/B + rank-0 kernel timeline
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout -k 10 600 $TR tools/ddp_trace.py --trace gpurun_out/r02_run28_trace.json "bucket_mb=8" "bucket_mb=8,ready=1" > gpurun_out/r02_run28_trace.log 2>&1
echo "exit $?" >> gpurun_out/r02_run28_trace.log
timeout -k 10 600 $TR tools/ddp_trace.py "bucket_mb=8,prio=-1" "bucket_mb=32" "bucket_mb=2" "bucket_mb=8,ready=1,prio=-1" > gpurun_out/r02_run28_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run28_ab.log
timeout -k 10 300 python tools/ddp_trace.py "reducer=0" "bucket_mb=8" > gpurun_out/r02_run28_w1.log 2>&1
echo "exit $?" >> gpurun_out/r02_run28_w1.log
grep -h "world\|trace:\|exit\|rror" gpurun_out/r02_run28_trace.log gpurun_out/r02_run28_ab.log gpurun_out/r02_run28_w1.log | tail -30
