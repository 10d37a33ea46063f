#!/bin/bash
# round 2, call 17: stem conv + statistics in one launch, full suite, bench (100 and 20 steps), --set full of the layer-2/3/4 fprop at width 128
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c17_tests_all.log 2>&1; echo "rc=$?" >> gpurun_out/c17_tests_all.log
B="python bench.py --warmup 10 --no-aug --no-large --no-cpu-baseline --no-library"
run() { name=$1; shift
  timeout 300 $B "$@" 2> gpurun_out/c17_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f={x['family'][:10]:x['us_per_step'] for x in d['kernel_families']}; print('$name', d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], f)" >> gpurun_out/c17_ab.txt
}
run s200 --steps 200
run s20 --steps 20 --warmup 5
run mt_s200 --steps 200 --workload mean_teacher_resnet18_qtdb_2x2500_b16+16
W=fixmatch_resnet18w128_12x5000_b64+64
timeout 300 python bench.py --profile-mode --workload $W --steps 1 --warmup 3 > gpurun_out/c17_plain_w128.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tn3_kernel<\(int\)256, \(bool\)1, \(bool\)1, \(bool\)0" -s 27 -c 9 -o gpurun_out/c17_tn3_w128 python bench.py --profile-mode --workload $W --steps 1 --warmup 3 > gpurun_out/c17_ncu_tn3.log 2>&1
cat gpurun_out/c17_ab.txt; tail -n 6 gpurun_out/c17_tests_all.log | cut -c1-200; ls -la gpurun_out/c17_*rep
