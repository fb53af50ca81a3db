#!/bin/bash
# round 2, GPU call 4 (2 GPUs): NCCL gradient-exchange parity test + 2-GPU bench line (weak scaling + configs[3] point)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_run4_smi.txt 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s --timeout 800 -p no:cacheprovider > gpurun_out/r02_run4_multi.log 2>&1
echo "exit $?" >> gpurun_out/r02_run4_multi.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 --no-decode-e2e > gpurun_out/r02_run4_bench2.json 2> gpurun_out/r02_run4_bench2.err
echo "exit $?" >> gpurun_out/r02_run4_bench2.err
tail -15 gpurun_out/r02_run4_multi.log; tail -c 1200 gpurun_out/r02_run4_bench2.json; tail -5 gpurun_out/r02_run4_bench2.err
