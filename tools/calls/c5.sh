#!/bin/bash
# round 2, call 5: cluster-16 BN backward, FFT augmentation: tests, A/B, timeline with phase marks
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "bn_bwd_fused" > gpurun_out/c5_tests_bn.log 2>&1; echo "rc=$?" >> gpurun_out/c5_tests_bn.log
timeout 900 python -m pytest tests/test_augment_gpu.py -q > gpurun_out/c5_tests_aug.log 2>&1; echo "rc=$?" >> gpurun_out/c5_tests_aug.log
B="python bench.py --steps 200 --warmup 10 --no-aug --no-large --no-cpu-baseline --no-library"
run() { name=$1; shift
  env "$@" timeout 300 $B 2> gpurun_out/c5_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f={x['family'][:10]:x['us_per_step'] for x in d['kernel_families']}; print('$name', d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], f)" >> gpurun_out/c5_ab.txt
}
run cluster0 SSB_BN_CLUSTER=0
run cluster16 SSB_BN_CLUSTER=16
run cluster8 SSB_BN_CLUSTER=8
SSB_LIB=$PWD/semi-seg-ecg_b200/lib/libsemiseg_b200_trace.so SSB_BN_CLUSTER=16 timeout 300 python tools/trace_step.py --out gpurun_out/c5_trace_cl16.md > gpurun_out/c5_trace.log 2>&1
SSB_LIB=$PWD/semi-seg-ecg_b200/lib/libsemiseg_b200_trace.so SSB_BN_CLUSTER=0 timeout 300 python tools/trace_step.py --out gpurun_out/c5_trace_cl0.md >> gpurun_out/c5_trace.log 2>&1
SSB_AUG_FFT=0 timeout 120 python tools/aug_profile.py > gpurun_out/c5_aug_dense.log 2>&1
SSB_AUG_FFT=1 timeout 120 python tools/aug_profile.py > gpurun_out/c5_aug_fft.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aug_ -s 12 -c 12 --csv --log-file gpurun_out/c5_aug.csv python tools/aug_profile.py 8 > gpurun_out/c5_aug_ncu.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_kernels_gpu.py::test_bn_bwd_fused --deselect tests/test_augment_gpu.py > gpurun_out/c5_tests_all.log 2>&1; echo "rc=$?" >> gpurun_out/c5_tests_all.log
cat gpurun_out/c5_ab.txt; tail -n 4 gpurun_out/c5_tests_bn.log gpurun_out/c5_tests_aug.log gpurun_out/c5_tests_all.log gpurun_out/c5_aug_dense.log gpurun_out/c5_aug_fft.log
