#!/bin/bash
# round 2, call 11: tensor-core stem (tests + width-128 step A/B), SyncBN exchange tests, driver-like bench A/B
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_bench_shapes_gpu.py -q -k "stem or syncbn or bn_bwd_fused" > gpurun_out/c11_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c11_tests.log
for W in fixmatch_resnet18w128_12x5000_b32+32 mean_teacher_resnet18_qtdb_2x2500_b16+16; do
 for S in 0 1; do
  SSB_STEM_TC=$S timeout 600 python bench.py --workload $W --steps 30 --warmup 5 --no-aug --no-large --no-cpu-baseline --no-library > gpurun_out/c11_${W}_tc$S.json 2> gpurun_out/c11_${W}_tc$S.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/c11_${W}_tc$S.json').read().strip().splitlines()[-1]); print('$W stem_tc=$S', d['ms_per_step'], d['value'], [(x['family'][:10],x['us_per_step'],x['tflops'],x['gbs']) for x in d['kernel_families']], d['parity_at_bench_shape'] if 'parity_at_bench_shape' in d else '')" >> gpurun_out/c11_ab.txt
 done
done
B="python bench.py --steps 20 --warmup 5 --no-aug --no-large --no-cpu-baseline --no-library"
run() { name=$1; shift
  env "$@" timeout 300 $B 2> gpurun_out/c11_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'])" >> gpurun_out/c11_ab.txt
}
run s20_buckets2 SSB_BUCKETS=2
run s20_buckets0 SSB_BUCKETS=0
run s20_buckets2_again SSB_BUCKETS=2
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c11_tests_all.log 2>&1; echo "rc=$?" >> gpurun_out/c11_tests_all.log
cat gpurun_out/c11_ab.txt; tail -n 5 gpurun_out/c11_tests.log; tail -n 6 gpurun_out/c11_tests_all.log
