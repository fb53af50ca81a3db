#!/bin/bash
# round 2, GPU call 56 (2 GPUs): exchange blocks of 256 threads (co-resident with the chain's CTAs) at equal total threads
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
timeout -k 10 200 $TR tools/ddp_trace.py "impl=p2p,bucket_mb=16,ready=1" > gpurun_out/r02_run56_a.log 2>&1
PD_AR_THREADS=256 timeout -k 10 200 $TR tools/ddp_trace.py "impl=p2p,bucket_mb=16,ready=1,blocks=64,streams=2" "impl=p2p,bucket_mb=16,ready=1,blocks=48,streams=3" > gpurun_out/r02_run56_b.log 2>&1
grep -h "^world" gpurun_out/r02_run56_a.log gpurun_out/r02_run56_b.log
