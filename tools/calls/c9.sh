#!/bin/bash
# round 2, call 9: in-kernel SyncBN exchange (single-GPU protocol tests), augmentation, driver-like bench (20 steps), full suite
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "syncbn" > gpurun_out/c9_tests_sync.log 2>&1; echo "rc=$?" >> gpurun_out/c9_tests_sync.log
timeout 900 python -m pytest tests/test_augment_gpu.py -q > gpurun_out/c9_tests_aug.log 2>&1; echo "rc=$?" >> gpurun_out/c9_tests_aug.log
SSB_AUG_FFT=0 timeout 120 python tools/aug_profile.py > gpurun_out/c9_aug_dense.log 2>&1
SSB_AUG_FFT=1 timeout 120 python tools/aug_profile.py > gpurun_out/c9_aug_fft.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aug_ -s 12 -c 12 --csv --log-file gpurun_out/c9_aug.csv python tools/aug_profile.py 8 > gpurun_out/c9_aug_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c9_bench20.json 2> gpurun_out/c9_bench20.err; echo "rc=$?" >> gpurun_out/c9_bench20.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c9_ref20.json 2> gpurun_out/c9_ref20.err
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c9_tests_all.log 2>&1; echo "rc=$?" >> gpurun_out/c9_tests_all.log
python -c "
import json
d=json.loads(open('gpurun_out/c9_bench20.json').read().strip().splitlines()[-1])
print('bench20', d['value'], d['ms_per_step'], 'e2e', d['e2e'], d['gpu_augmentation'])"
tail -n 3 gpurun_out/c9_tests_sync.log gpurun_out/c9_tests_aug.log gpurun_out/c9_aug_dense.log gpurun_out/c9_aug_fft.log; tail -n 8 gpurun_out/c9_tests_all.log
