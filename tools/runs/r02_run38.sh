#!/bin/bash
# round 2, GPU call 38 (1 GPU): prologue issued before the encoders (graph nodes start in creation order) -- A/B + timeline
mkdir -p gpurun_out
timeout -k 10 600 python tools/step_ab.py "" "PROLOGUE_FIRST=0" "" "PROLOGUE_FIRST=0" > gpurun_out/r02_run38_ab.log 2>&1
echo "exit $?" >> gpurun_out/r02_run38_ab.log
timeout -k 10 300 python tools/trace_step.py gpurun_out/trace_packed7.json > gpurun_out/r02_run38_trace.log 2>&1
cat gpurun_out/r02_run38_ab.log; tail -2 gpurun_out/r02_run38_trace.log
