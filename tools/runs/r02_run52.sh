#!/bin/bash
# round 2, GPU call 52 (1 GPU): final-state validation -- full -m gpu suite, smoke, bench line, launch list
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run52_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run52_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_run52_smoke.log 2>&1
timeout -k 10 900 python bench.py > gpurun_out/r02_run52_bench.json 2> gpurun_out/r02_run52_bench.err
echo "exit $?" >> gpurun_out/r02_run52_bench.err
timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_v7.csv python tools/profile_step.py > gpurun_out/r02_run52_ncu1.log 2>&1
tail -4 gpurun_out/r02_run52_tests.log; tail -2 gpurun_out/r02_run52_smoke.log; head -c 300 gpurun_out/r02_run52_bench.json
