#!/bin/bash
# round 2, call 14: RED epilogue with register-resident operands: tests, A/B, phase marks
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "dgrad_bnred" > gpurun_out/c14_tests_red.log 2>&1; echo "rc=$?" >> gpurun_out/c14_tests_red.log
B="python bench.py --steps 200 --warmup 10 --no-aug --no-large --no-cpu-baseline --no-library"
run() { name=$1; shift
  env "$@" timeout 300 $B 2> gpurun_out/c14_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f={x['family'][:10]:x['us_per_step'] for x in d['kernel_families']}; print('$name', d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], f)" >> gpurun_out/c14_ab.txt
}
run red0 SSB_FUSE_REDUCE=0
run red1 SSB_FUSE_REDUCE=1
SSB_LIB=$PWD/semi-seg-ecg_b200/lib/libsemiseg_b200_trace.so SSB_FUSE_REDUCE=1 timeout 300 python tools/trace_step.py --out gpurun_out/c14_trace_red1.md > gpurun_out/c14_trace.log 2>&1
cat gpurun_out/c14_ab.txt; tail -n 4 gpurun_out/c14_tests_red.log | cut -c1-300
