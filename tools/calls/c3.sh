#!/bin/bash
# round 2, call 3: updated parity tests; A/B of the wgrad split rule and of the staged H2D copies; width-128 step at two batches
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_bench_shapes_gpu.py -q -s -k "fp32 or mean_teacher or trained or bf16" > gpurun_out/c3_tests_new.log 2>&1; echo "rc=$?" >> gpurun_out/c3_tests_new.log
B="python bench.py --steps 200 --warmup 10 --no-aug --no-large --no-cpu-baseline --no-library"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $B 2> gpurun_out/c3_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f={x['family'][:10]:x['us_per_step'] for x in d['kernel_families']}; print('$name', d['ms_per_step'], d['e2e']['ms_per_step'], f)" >> gpurun_out/c3_ab.txt
}
run split0 SSB_WG_SPLIT=0
run split1 SSB_WG_SPLIT=1
run split1_f8 SSB_WG_SPLIT=1 SSB_WG_FIXED=8
run split1_f16 SSB_WG_SPLIT=1 SSB_WG_FIXED=16
run h2d0 SSB_STAGE_H2D=0
run h2d1 SSB_STAGE_H2D=1
for W in fixmatch_resnet18w128_12x5000_b32+32 fixmatch_resnet18w128_12x5000_b64+64; do
 for S in 0 1; do
  SSB_WG_SPLIT=$S timeout 600 python bench.py --workload $W --steps 20 --warmup 5 --no-aug --no-large --no-cpu-baseline --no-library > gpurun_out/c3_${W}_s$S.json 2> gpurun_out/c3_${W}_s$S.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/c3_${W}_s$S.json').read().strip().splitlines()[-1]); print('$W split$S', d['ms_per_step'], d['value'], [(x['family'][:10],x['us_per_step'],x['tflops']) for x in d['kernel_families']])" >> gpurun_out/c3_ab.txt
 done
done
cat gpurun_out/c3_ab.txt; grep -v "^  grad" gpurun_out/c3_tests_new.log | tail -n 12
