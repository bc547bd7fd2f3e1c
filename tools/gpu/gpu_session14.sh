#!/bin/bash
# round-2 GPU session 14 (2 GPUs): LL reduce-scatter fails tp_check's 7-wide decode: data path or timing?
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
: > gpurun_out/s14_all.log
run() { name=$1; shift; timeout 200 "$@" 2> gpurun_out/s14_$name.err | grep "tp=" | sed "s/^/$name: /" | tee -a gpurun_out/s14_all.log; }
run default $TR --master-port 29511 tools/tp_check.py --geom tiny
LLMLB_DEBUG_NO_RSLL=1 run norsll $TR --master-port 29512 tools/tp_check.py --geom tiny
LLMLB_DEBUG_NO_PDL=16 run nopdl16 $TR --master-port 29513 tools/tp_check.py --geom tiny
LLMLB_DEBUG_NO_PDL=31 run nopdl31 $TR --master-port 29514 tools/tp_check.py --geom tiny
CUDA_LAUNCH_BLOCKING=1 run blocking $TR --master-port 29515 tools/tp_check.py --geom tiny
