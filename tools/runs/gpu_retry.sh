#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit 3 / transient); usage: gpu_retry.sh <timeout> <script> [gpus]
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then out=$(gpurun --timeout $T -- "bash $S" 2>&1); else out=$(gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 90; continue; fi
  echo "$out" | tail -60; exit $rc
done
echo "gave up: pod busy"; exit 3
