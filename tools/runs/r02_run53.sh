#!/bin/bash
# round 2, GPU call 53 (1 GPU): ncu --set full of the round's late kernels (greedy slot, pack_order, length-aware colsum)
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02b_kernels \
    python tools/profile_kernels_r02b.py > gpurun_out/r02_run53_ncu.log 2>&1
echo "exit $?" >> gpurun_out/r02_run53_ncu.log
python tools/ncu_brief.py gpurun_out/r02b_kernels.ncu-rep > gpurun_out/r02b_ncu_brief.txt 2>&1
tail -3 gpurun_out/r02_run53_ncu.log; grep -c "^==" gpurun_out/r02b_ncu_brief.txt
