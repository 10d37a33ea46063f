#!/bin/bash
# round 2, call 16: profiles -- launch list + DRAM traffic of the config-2 step, --set full of the layer-2/3/4 fprop and the
# tap-fused wgrad kernels at width 128 (64+64), RED at width 128 A/B, the bench line
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; echo "rc=$?" >> gpurun_out/c16_bench.err
for S in 0 1; do
  SSB_FUSE_REDUCE=$S timeout 600 python bench.py --workload fixmatch_resnet18w128_12x5000_b32+32 --steps 30 --warmup 5 --no-aug --no-large --no-cpu-baseline --no-library > gpurun_out/c16_w128_red$S.json 2> gpurun_out/c16_w128_red$S.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/c16_w128_red$S.json').read().strip().splitlines()[-1]); print('w128 32+32 red=$S', d['ms_per_step'], [(x['family'][:10],x['us_per_step']) for x in d['kernel_families']])" >> gpurun_out/c16_ab.txt
done
timeout 300 python bench.py --profile-mode --steps 2 --warmup 3 > gpurun_out/c16_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --print-units base --clock-control none -s 411 -c 274 --csv --log-file gpurun_out/c16_traffic.csv python bench.py --profile-mode --steps 2 --warmup 3 > gpurun_out/c16_ncu.log 2>&1
W=fixmatch_resnet18w128_12x5000_b64+64
timeout 300 python bench.py --profile-mode --workload $W --steps 1 --warmup 3 > gpurun_out/c16_plain_w128.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tn3_kernel<256, 1, 1, 0, 0>" -s 27 -c 9 -o gpurun_out/c16_tn3_w128 python bench.py --profile-mode --workload $W --steps 1 --warmup 3 > gpurun_out/c16_ncu_tn3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad3_kernel -s 27 -c 9 -o gpurun_out/c16_wg3_w128 python bench.py --profile-mode --workload $W --steps 1 --warmup 3 > gpurun_out/c16_ncu_wg3.log 2>&1
cat gpurun_out/c16_ab.txt; tail -n 2 gpurun_out/c16_bench.err; ls -la gpurun_out/c16_*
