#!/bin/bash
# round 2, call 15: BN kernels templated on SYNC (lean N=1 instantiations): tests + bench at 200 and 20 steps
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "bn_ or syncbn or stem" > gpurun_out/c15_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c15_tests.log
B="python bench.py --warmup 10 --no-aug --no-large --no-cpu-baseline --no-library"
run() { name=$1; shift
  env "$@" timeout 300 $B 2> gpurun_out/c15_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f={x['family'][:10]:x['us_per_step'] for x in d['kernel_families']}; print('$name', d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], f)" >> gpurun_out/c15_ab.txt
}
B="$B --steps 200"; run s200
B="python bench.py --warmup 5 --no-aug --no-large --no-cpu-baseline --no-library --steps 20"; run s20; run s20_again
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c15_tests_all.log 2>&1; echo "rc=$?" >> gpurun_out/c15_tests_all.log
cat gpurun_out/c15_ab.txt; tail -n 4 gpurun_out/c15_tests.log; tail -n 6 gpurun_out/c15_tests_all.log | cut -c1-200
