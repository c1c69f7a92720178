#!/bin/bash
# usage: run.sh <gpus> <timeout_s> <script> <logfile>   — retries while the pod has no free slot (nothing is charged for those)
gpus=$1; to=$2; script=$3; log=$4
for attempt in $(seq 1 40); do
  if [ "$gpus" = "1" ]; then gpurun --timeout $to -- "bash $script" > $log 2>&1; else gpurun --gpus $gpus --timeout $to -- "bash $script" > $log 2>&1; fi
  rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "done rc=$rc attempt=$attempt" >> $log
