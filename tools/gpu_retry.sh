#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> <script> [gpus]  -- retries while the pod answers busy (nothing is charged for those)
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient\|status=busy\|retry in a few minutes\|no box"; then sleep 100; continue; fi
  echo "$OUT" | grep -v "^+" | tail -${TAILN:-30}
  exit 0
done
echo "gave up after 40 tries"
