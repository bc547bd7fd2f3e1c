#!/bin/bash
# round-2 GPU session 12 (2 GPUs): after removing __restrict__ from qkv in the decode attention kernel (q loads were hoisted
# above griddepcontrol.wait): determinism probe, tp parity, bench N=2 and N=1
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=tools/tp_race_probe.py
: > gpurun_out/s12_all.log
run() { name=$1; shift; timeout 240 "$@" --tag $name 2> gpurun_out/s12_$name.err | grep "tp=" | tee -a gpurun_out/s12_all.log; }
run pairs $TR --master-port 29511 $P --max-seqs 64
run flags $TR --master-port 29512 $P --max-seqs 64 --proto 1
run gather $TR --master-port 29513 $P --max-seqs 64 --proto 2
timeout 900 python -m pytest tests/test_tp_gpu.py -q -p no:cacheprovider > gpurun_out/s12_tp.log 2>&1
echo "tp rc=$?" >> gpurun_out/s12_tp.log
timeout 600 $TR --master-port 29520 bench.py --gpus 2 --steps 5 --warmup 3 --no-ref-shape > gpurun_out/s12_bench_n2.json 2> gpurun_out/s12_bench_n2.err
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-ref-shape > gpurun_out/s12_bench_n1.json 2> gpurun_out/s12_bench_n1.err
timeout 300 python tools/determinism_probe.py > gpurun_out/s12_determinism_tp1.log 2>&1
tail -3 gpurun_out/s12_tp.log; tail -3 gpurun_out/s12_determinism_tp1.log
