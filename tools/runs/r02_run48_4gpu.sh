#!/bin/bash
# round 2, GPU call 48 (4 GPUs): the peer-memory exchange at W=4 -- exactness test, bench line
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout -k 10 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k kernel_exact --timeout 280 -p no:cacheprovider > gpurun_out/r02_run48_kernel.log 2>&1
echo "exit $? gpus $N" >> gpurun_out/r02_run48_kernel.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-decode-e2e --no-tfr0 > gpurun_out/r02_run48_bench4.json 2> gpurun_out/r02_run48_bench4.err
echo "exit $?" >> gpurun_out/r02_run48_bench4.err
tail -3 gpurun_out/r02_run48_kernel.log; head -c 330 gpurun_out/r02_run48_bench4.json; echo; tail -2 gpurun_out/r02_run48_bench4.err
