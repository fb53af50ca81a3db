#!/bin/bash
# round 2, GPU call 47 (1 GPU): final-state validation after the last kernel changes -- full -m gpu suite, smoke, bench line
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_run47_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/r02_run47_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_run47_smoke.log 2>&1
timeout -k 10 900 python bench.py > gpurun_out/r02_run47_bench.json 2> gpurun_out/r02_run47_bench.err
echo "exit $?" >> gpurun_out/r02_run47_bench.err
tail -4 gpurun_out/r02_run47_tests.log; tail -2 gpurun_out/r02_run47_smoke.log; head -c 400 gpurun_out/r02_run47_bench.json
