#!/bin/bash
# round 2, GPU call 34 (8 GPUs, or 4): the p2p exchange at W=8 -- exactness test, then the bench line (and NCCL for comparison)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout -k 10 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k kernel_exact --timeout 280 -p no:cacheprovider > gpurun_out/r02_run34_kernel.log 2>&1
echo "exit $? gpus $N" >> gpurun_out/r02_run34_kernel.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518"
timeout -k 10 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-decode-e2e --no-tfr0 --no-strong > gpurun_out/r02_run34_bench_p2p.json 2> gpurun_out/r02_run34_bench_p2p.err
echo "exit $?" >> gpurun_out/r02_run34_bench_p2p.err
timeout -k 10 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-decode-e2e --no-tfr0 --no-strong --exchange nccl --bucket-mb 32 > gpurun_out/r02_run34_bench_nccl.json 2> gpurun_out/r02_run34_bench_nccl.err
echo "exit $?" >> gpurun_out/r02_run34_bench_nccl.err
tail -3 gpurun_out/r02_run34_kernel.log; for f in p2p nccl; do head -c 330 gpurun_out/r02_run34_bench_$f.json; echo; tail -2 gpurun_out/r02_run34_bench_$f.err; done
