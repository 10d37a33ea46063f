#!/bin/bash
# round 2, call 18 (8 GPUs, final state: in-kernel SyncBN exchange, pipelined all-reduce + AdamW, tensor-core stem): BASELINE configs 3 / 4 / 5 and the default config through bench.py at N = 1..8
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c18_smi.txt
OPT="--steps 100 --warmup 5 --no-aug --no-large --no-cpu-baseline --no-library"
run() { # n workload tag
  n=$1; w=$2; tag=$3
  if [ "$n" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$((40+n))"; fi
  timeout 420 $L bench.py --gpus $n --workload $w $OPT > gpurun_out/c18_${tag}_n$n.json 2> gpurun_out/c18_${tag}_n$n.err; echo "rc=$?" >> gpurun_out/c18_${tag}_n$n.err
  python - <<PY >> gpurun_out/c18_summary.txt
import json
try:
    d=json.loads(open('gpurun_out/c18_${tag}_n$n.json').read().strip().splitlines()[-1])
    o=d.get('sync_bn_off') or d.get('sync_bn_on') or {}
    print('${tag}', 'N=$n', 'value', d['value'], 'ms', d['ms_per_step'], 'e2e_ms', d['e2e']['ms_per_step'], 'sync_bn', d['config']['sync_bn'], '| other:', o.get('sync_bn'), o.get('value'), o.get('ms_per_step'), '| replicas', (d.get('replicas_equal') or {}).get('ok'))
except Exception as e:
    print('${tag}', 'N=$n', 'FAILED', e)
PY
}
MT=mean_teacher_resnet18_qtdb_2x2500_b16+16
for n in 1 2 4 8; do run $n fixmatch_resnet18_ludb_1x2500_b16+16 cfg2; done
for n in 2 8; do run $n $MT mt; done
run 8 fixmatch_resnet18_merged_1x2500_b4+28 merged
for b in 16+16 64+64 256+256; do run 8 fixmatch_resnet18w128_12x5000_b$b w128_$b; done
cat gpurun_out/c18_summary.txt
