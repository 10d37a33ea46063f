#!/bin/bash
# round 2, call 19: full -m gpu suite after the one-lead stem went back to the fp32-operand kernel; the driver's bench command
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c19_tests_all.log 2>&1; echo "rc=$?" >> gpurun_out/c19_tests_all.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err; echo "rc=$?" >> gpurun_out/c19_bench.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c19_smoke.log 2>&1
tail -n 6 gpurun_out/c19_tests_all.log | cut -c1-200; tail -c 600 gpurun_out/c19_bench.json; tail -n 2 gpurun_out/c19_smoke.log
