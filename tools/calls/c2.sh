#!/bin/bash
# round 2, call 2: fixed tests, in-kernel timeline of the config-2 step, branch-skipping timing experiments, augmentation launch list
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_entrypoints_gpu.py -q -x > gpurun_out/c2_tests_entry.log 2>&1; echo "rc=$?" >> gpurun_out/c2_tests_entry.log
timeout 1200 python -m pytest tests/test_bench_shapes_gpu.py -q -s -k "fp32 or mean_teacher or trained or pseudo_dtype" > gpurun_out/c2_tests_new.log 2>&1; echo "rc=$?" >> gpurun_out/c2_tests_new.log
SSB_LIB=$PWD/semi-seg-ecg_b200/lib/libsemiseg_b200_trace.so timeout 300 python tools/trace_step.py --out gpurun_out/c2_trace.md > gpurun_out/c2_trace.log 2>&1
B="python bench.py --steps 200 --warmup 10 --no-aug --no-large --no-cpu-baseline --no-library"
for skip in none wgrad teacher wgrad,teacher; do
  SSB_DEBUG_SKIP=$skip timeout 300 $B 2> gpurun_out/c2_skip_$skip.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$skip', d['ms_per_step'], d['launches_per_step'], d['e2e']['ms_per_step'])" >> gpurun_out/c2_skip.txt
done
timeout 120 python tools/aug_profile.py > gpurun_out/c2_aug_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aug_ -s 12 -c 12 --csv --log-file gpurun_out/c2_aug.csv python tools/aug_profile.py 8 > gpurun_out/c2_aug_ncu.log 2>&1
cat gpurun_out/c2_skip.txt; tail -n 3 gpurun_out/c2_tests_entry.log; tail -n 8 gpurun_out/c2_tests_new.log
