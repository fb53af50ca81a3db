#!/bin/bash
# round 2, GPU call 30 (2 GPUs): micro-benchmark of the peer-memory exchange kernel (grid, access strength, unroll) vs NCCL
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514"
for v in 0 1 2; do
  PD_AR_VARIANT=$v timeout -k 10 300 $TR tools/ar_bench.py > gpurun_out/r02_run30_ar_v$v.log 2>&1
  echo "exit $?" >> gpurun_out/r02_run30_ar_v$v.log
done
nvidia-smi topo -m > gpurun_out/r02_run30_topo.txt 2>&1
grep -h "PD_AR\|p2p\|nccl\|exit\|rror" gpurun_out/r02_run30_ar_v*.log | tail -120
