#!/bin/bash
# usage: tools/calls/run.sh <name> <timeout_s> [--gpus N]  -- runs tools/calls/<name>.sh on the GPU box, retrying while the pod is busy
name=$1; to=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" -- "bash tools/calls/$name.sh" > gpurun_out/${name}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)"; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3
