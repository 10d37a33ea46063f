#!/bin/bash
# round 2, call 10 (2 GPUs): in-kernel SyncBN exchange: DP == single process on hardware, N=2 bench A/B
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py -q -x > gpurun_out/c10_tests_dp.log 2>&1; echo "rc=$?" >> gpurun_out/c10_tests_dp.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
OPT="--gpus 2 --steps 200 --warmup 10 --no-aug --no-large --no-cpu-baseline --no-library"
timeout 600 $TR bench.py $OPT > gpurun_out/c10_fused.json 2> gpurun_out/c10_fused.err; echo "rc=$?" >> gpurun_out/c10_fused.err
SSB_SYNCBN_FUSED=0 timeout 600 $TR bench.py $OPT --no-syncbn-other > gpurun_out/c10_unfused.json 2> gpurun_out/c10_unfused.err; echo "rc=$?" >> gpurun_out/c10_unfused.err
timeout 600 $TR bench.py $OPT --workload fixmatch_resnet18w128_12x5000_b16+16 --steps 50 > gpurun_out/c10_w128.json 2> gpurun_out/c10_w128.err; echo "rc=$?" >> gpurun_out/c10_w128.err
for f in c10_fused c10_unfused c10_w128; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['sync_bn'], d.get('sync_bn_off'), d.get('replicas_equal'), d['run'])"; done
tail -n 5 gpurun_out/c10_tests_dp.log; tail -n 3 gpurun_out/c10_fused.err
