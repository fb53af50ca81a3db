#!/bin/bash
# round 2, GPU call 40 (2 GPUs): final state -- all multi-GPU tests (NCCL + p2p + exchange kernel), 2-GPU bench line
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -s --timeout 1000 -p no:cacheprovider > gpurun_out/r02_run40_multi.log 2>&1
echo "exit $?" >> gpurun_out/r02_run40_multi.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 2 --steps 20 --warmup 3 --no-decode-e2e > gpurun_out/r02_run40_bench2.json 2> gpurun_out/r02_run40_bench2.err
echo "exit $?" >> gpurun_out/r02_run40_bench2.err
grep -E "2-rank|passed|failed|rror" gpurun_out/r02_run40_multi.log | tail -6; head -c 330 gpurun_out/r02_run40_bench2.json; echo; tail -2 gpurun_out/r02_run40_bench2.err
