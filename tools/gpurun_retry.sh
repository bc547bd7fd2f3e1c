#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout_s> <logfile> <command...>   — retries while the pod answers "busy"
G=$1; T=$2; LOG=$3; shift 3
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" > "$LOG" 2>&1
  if grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null || grep -q "status=transient\|backing off" "$LOG"; then
    sleep 75
    continue
  fi
  break
done
tail -40 "$LOG"
