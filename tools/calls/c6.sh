#!/bin/bash
# round 2, call 6: batched loads in the fused BN backward (A/B slack, PDL level), full GPU suite, bench line + timeline
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "bn_bwd_fused" > gpurun_out/c6_tests_bn.log 2>&1; echo "rc=$?" >> gpurun_out/c6_tests_bn.log
B="python bench.py --steps 200 --warmup 10 --no-aug --no-large --no-cpu-baseline --no-library"
run() { name=$1; shift
  env "$@" timeout 300 $B 2> gpurun_out/c6_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f={x['family'][:10]:x['us_per_step'] for x in d['kernel_families']}; print('$name', d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], f)" >> gpurun_out/c6_ab.txt
}
run slack1 SSB_BWF_SLACK=1
run slack0 SSB_BWF_SLACK=0
run pdl1 SSB_PDL=1
run pdl1_slack0 SSB_PDL=1 SSB_BWF_SLACK=0
run buckets3 SSB_BUCKETS=3
SSB_LIB=$PWD/semi-seg-ecg_b200/lib/libsemiseg_b200_trace.so timeout 300 python tools/trace_step.py --out gpurun_out/c6_trace.md > gpurun_out/c6_trace.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_kernels_gpu.py::test_bn_bwd_fused > gpurun_out/c6_tests_all.log 2>&1; echo "rc=$?" >> gpurun_out/c6_tests_all.log
cat gpurun_out/c6_ab.txt; tail -n 4 gpurun_out/c6_tests_bn.log; tail -n 12 gpurun_out/c6_tests_all.log
