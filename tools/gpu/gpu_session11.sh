#!/bin/bash
# round-2 GPU session 11 (2 GPUs): which programmatic edge of the tp decode chain is the racy one
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=tools/tp_race_probe.py
: > gpurun_out/s11_all.log
run() { name=$1; shift; timeout 240 "$@" --tag $name 2> gpurun_out/s11_$name.err | grep "tp=" | tee -a gpurun_out/s11_all.log; }
for m in 1 2 4 8 16 6 14; do
  LLMLB_DEBUG_NO_PDL=$m run mask$m $TR --master-port $((29510 + m)) $P --max-seqs 64 --runs 3
done
