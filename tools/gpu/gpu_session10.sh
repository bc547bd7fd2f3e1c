#!/bin/bash
# round-2 GPU session 10 (2 GPUs): the tp run-to-run differences — PDL overlap inside a step, across steps, or graphs?
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=tools/tp_race_probe.py
: > gpurun_out/s10_all.log
run() { name=$1; shift; timeout 240 "$@" --tag $name 2> gpurun_out/s10_$name.err | grep "tp=" | tee -a gpurun_out/s10_all.log; }
run nographs $TR --master-port 29511 $P --max-seqs 64 --no-graphs
LLMLB_DEBUG_NO_PDL=1 run nopdl $TR --master-port 29512 $P --max-seqs 64
LLMLB_DEBUG_NO_PDL=1 run nopdl_nographs $TR --master-port 29513 $P --max-seqs 64 --no-graphs
CUDA_LAUNCH_BLOCKING=1 run blocking_graphs $TR --master-port 29514 $P --max-seqs 64
LLMLB_DEBUG_NO_PDL=1 run nopdl_flags $TR --master-port 29515 $P --max-seqs 64 --proto 1
timeout 300 python tools/determinism_probe.py > gpurun_out/s10_determinism_tp1.log 2>&1
tail -4 gpurun_out/s10_determinism_tp1.log
