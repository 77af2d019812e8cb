#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> <gpus> '<command>' <logfile>  — retries while the pod answers "busy" (exit 3)
T=$1; G=$2; CMD=$3; LOG=$4
for i in $(seq 1 40); do
  if [ "$G" -gt 1 ]; then /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" > "$LOG" 2>&1; else /usr/local/graft/bin/gpurun --timeout $T -- "$CMD" > "$LOG" 2>&1; fi
  rc=$?
  if ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 90
done
exit 3
