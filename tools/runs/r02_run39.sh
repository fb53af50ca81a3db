#!/bin/bash
# round 2, GPU call 39 (1 GPU): final-state validation -- full -m gpu suite, bench line, ncu launch list v6, timeline
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run39_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run39_tests.log
timeout -k 10 900 python bench.py > gpurun_out/r02_run39_bench.json 2> gpurun_out/r02_run39_bench.err
echo "exit $?" >> gpurun_out/r02_run39_bench.err
timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_v6.csv python tools/profile_step.py > gpurun_out/r02_run39_ncu1.log 2>&1
timeout -k 10 300 python tools/trace_step.py gpurun_out/trace_packed8.json > gpurun_out/r02_run39_trace.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_run39_smoke.log 2>&1
tail -4 gpurun_out/r02_run39_tests.log; head -c 400 gpurun_out/r02_run39_bench.json; tail -1 gpurun_out/r02_run39_trace.log; tail -2 gpurun_out/r02_run39_smoke.log
