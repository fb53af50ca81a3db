#!/bin/bash
# round 2, GPU call 37 (1 GPU): the prologue branch on a higher-priority stream -- step A/B, replay consistency
mkdir -p gpurun_out
timeout -k 10 600 python tools/step_ab.py "" "URGENT_PRIORITY=-1" "URGENT_PRIORITY=0" "" "URGENT_PRIORITY=-1" "URGENT_PRIORITY=-3" > gpurun_out/r02_run37_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run37_ab.log
timeout -k 10 300 python tools/replay_consistency.py "" > gpurun_out/r02_run37_consistency.log 2>&1
echo "exit $?" >> gpurun_out/r02_run37_consistency.log
python -c "import torch; print('priority range', torch.cuda.Stream.priority_range())" >> gpurun_out/r02_run37_ab.log 2>&1
cat gpurun_out/r02_run37_ab.log; tail -4 gpurun_out/r02_run37_consistency.log
