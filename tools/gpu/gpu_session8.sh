#!/bin/bash
# round-2 GPU session 8 (2 GPUs): is the tp parity record sensitive to max_seqs / repeatable?
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --parity-only"
i=0
run() { i=$((i+1)); name=$1; shift; timeout 200 "$@" 2> gpurun_out/s8_$name.err | grep parity_only | cut -c1-400 | sed "s/^/$name /" | tee -a gpurun_out/s8_all.log; }
: > gpurun_out/s8_all.log
run s64_a $TR --master-port 29511 $B
run s64_b $TR --master-port 29512 $B
run s64_c $TR --master-port 29513 $B
run s0_a $TR --master-port 29514 $B --streams 0
run s16 $TR --master-port 29515 $B --streams 16
LLMLB_DEBUG_NO_AGWAIT=1 run s64_noag $TR --master-port 29516 $B
LLMLB_DEBUG_NO_KSPLIT=1 run s64_noks $TR --master-port 29517 $B
run s64_flags $TR --master-port 29518 $B --tp-proto 1
run s64_nographs $TR --master-port 29519 $B --no-graphs
