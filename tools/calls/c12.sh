#!/bin/bash
# round 2, call 12: tensor-core stem weight gradient (tests + width-128 A/B), bench pre-roll at 20 steps, full suite
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_bench_shapes_gpu.py -q -k "stem" > gpurun_out/c12_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c12_tests.log
for W in fixmatch_resnet18w128_12x5000_b32+32 fixmatch_resnet18w128_12x5000_b64+64; do
 for S in 0 1; do
  SSB_STEM_TC=$S timeout 600 python bench.py --workload $W --steps 30 --warmup 5 --no-aug --no-large --no-library > gpurun_out/c12_${W}_tc$S.json 2> gpurun_out/c12_${W}_tc$S.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/c12_${W}_tc$S.json').read().strip().splitlines()[-1]); print('$W stem_tc=$S', d['ms_per_step'], d['value'], [(x['family'][:10],x['us_per_step'],x['tflops'],x['gbs']) for x in d['kernel_families']], d.get('parity_at_bench_shape'))" >> gpurun_out/c12_ab.txt
 done
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c12_bench20.json 2> gpurun_out/c12_bench20.err; echo "rc=$?" >> gpurun_out/c12_bench20.err
python -c "import json; d=json.loads(open('gpurun_out/c12_bench20.json').read().strip().splitlines()[-1]); print('bench20', d['value'], d['ms_per_step'], d['e2e'], d['roofline']['step_level'], d['parity_at_bench_shape'])" >> gpurun_out/c12_ab.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c12_tests_all.log 2>&1; echo "rc=$?" >> gpurun_out/c12_tests_all.log
cat gpurun_out/c12_ab.txt; tail -n 5 gpurun_out/c12_tests.log; tail -n 6 gpurun_out/c12_tests_all.log
