#!/bin/bash
# round-2 GPU session 9 (2 GPUs): what makes the pairs protocol's decode differ at max_seqs = 64
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=tools/tp_race_probe.py
: > gpurun_out/s9_all.log
run() { name=$1; shift; timeout 240 "$@" --tag $name 2> gpurun_out/s9_$name.err | grep "tp=" | tee -a gpurun_out/s9_all.log; }
run ms64 $TR --master-port 29511 $P --max-seqs 64
run ms4 $TR --master-port 29512 $P --max-seqs 4
run ms4_pool1024 $TR --master-port 29513 $P --max-seqs 4 --kv-pages 1024
run ms64_pool64 $TR --master-port 29514 $P --max-seqs 64 --kv-pages 64
run ms64_flags $TR --master-port 29515 $P --max-seqs 64 --proto 1
run ms64_look1 $TR --master-port 29516 $P --max-seqs 64 --lookahead 1
CUDA_LAUNCH_BLOCKING=1 run ms64_blocking $TR --master-port 29517 $P --max-seqs 64 --no-graphs
