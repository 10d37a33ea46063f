#!/bin/bash
# round 2, call 1: full GPU test suite + new bench-shape parity tests + bench + ncu traffic pass
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_bench_shapes_gpu.py > gpurun_out/c1_tests_old.log 2>&1; echo "old tests rc=$?" >> gpurun_out/c1_tests_old.log
timeout 1200 python -m pytest tests/test_bench_shapes_gpu.py -q -s > gpurun_out/c1_tests_new.log 2>&1; echo "new tests rc=$?" >> gpurun_out/c1_tests_new.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?" >> gpurun_out/c1_bench.err
timeout 300 python bench.py --profile-mode --steps 2 --warmup 3 > gpurun_out/c1_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --print-units base --clock-control none -s 402 -c 268 --csv --log-file gpurun_out/c1_traffic.csv python bench.py --profile-mode --steps 2 --warmup 3 > gpurun_out/c1_ncu.log 2>&1
W=fixmatch_resnet18w128_12x5000_b32+32
timeout 300 python bench.py --profile-mode --workload $W --steps 1 --warmup 3 > gpurun_out/c1_plain_w128.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --print-units base --clock-control none -k regex:conv_ -s 240 -c 80 --csv --log-file gpurun_out/c1_conv_w128.csv python bench.py --profile-mode --workload $W --steps 1 --warmup 3 > gpurun_out/c1_ncu_w128.log 2>&1
tail -5 gpurun_out/c1_tests_old.log gpurun_out/c1_tests_new.log gpurun_out/c1_bench.err
