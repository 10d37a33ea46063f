#!/bin/bash
# round 2, call 4: cluster BN backward + range-pipelined AdamW: tests, A/B timing, timeline
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "bn_bwd_fused" > gpurun_out/c4_tests_bn.log 2>&1; echo "rc=$?" >> gpurun_out/c4_tests_bn.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_kernels_gpu.py::test_bn_bwd_fused > gpurun_out/c4_tests_all.log 2>&1; echo "rc=$?" >> gpurun_out/c4_tests_all.log
B="python bench.py --steps 200 --warmup 10 --no-aug --no-large --no-cpu-baseline --no-library"
run() { name=$1; shift
  env "$@" timeout 300 $B 2> gpurun_out/c4_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f={x['family'][:10]:x['us_per_step'] for x in d['kernel_families']}; print('$name', d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], f)" >> gpurun_out/c4_ab.txt
}
run base SSB_BN_CLUSTER=0 SSB_BUCKETS=0
run cluster16 SSB_BN_CLUSTER=16 SSB_BUCKETS=0
run cluster8 SSB_BN_CLUSTER=8 SSB_BUCKETS=0
run buckets SSB_BN_CLUSTER=0 SSB_BUCKETS=2
run both SSB_BN_CLUSTER=16 SSB_BUCKETS=2
run both_b3 SSB_BN_CLUSTER=16 SSB_BUCKETS=3
SSB_LIB=$PWD/semi-seg-ecg_b200/lib/libsemiseg_b200_trace.so timeout 300 python tools/trace_step.py --out gpurun_out/c4_trace.md > gpurun_out/c4_trace.log 2>&1
cat gpurun_out/c4_ab.txt; tail -n 5 gpurun_out/c4_tests_bn.log gpurun_out/c4_tests_all.log
