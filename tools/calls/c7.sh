#!/bin/bash
# round 2, call 7 (2 GPUs): DP == single process on hardware (pytest), N=2 bench with SyncBN on (default) and off
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c7_smi.txt
timeout 900 python -m pytest tests/test_dp_gpu.py -q -x > gpurun_out/c7_tests_dp.log 2>&1; echo "rc=$?" >> gpurun_out/c7_tests_dp.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/c7_bench_n2.json 2> gpurun_out/c7_bench_n2.err; echo "rc=$?" >> gpurun_out/c7_bench_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 200 --warmup 10 --workload mean_teacher_resnet18_qtdb_2x2500_b16+16 > gpurun_out/c7_bench_mt_n2.json 2> gpurun_out/c7_bench_mt_n2.err; echo "rc=$?" >> gpurun_out/c7_bench_mt_n2.err
timeout 300 python bench.py --steps 200 --warmup 10 --no-aug --no-large --no-cpu-baseline --no-library > gpurun_out/c7_bench_n1.json 2> gpurun_out/c7_bench_n1.err
for f in c7_bench_n1 c7_bench_n2 c7_bench_mt_n2; do python -c "
import json,sys
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['sync_bn'], d.get('sync_bn_off'), d.get('sync_bn_on'), d.get('replicas_equal'))"; done
tail -n 5 gpurun_out/c7_tests_dp.log; tail -n 3 gpurun_out/c7_bench_n2.err
