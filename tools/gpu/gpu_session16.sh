#!/bin/bash
# round-2 GPU session 16 (8 GPUs): tp=8 parity, timelines of both consumer variants, the bench line at N=8 (configs[3]: 128 streams),
# Llama-3-70B mixed batch (configs[4])
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/s16_topo.txt 2>&1
: > gpurun_out/s16_chk.log
for g in tiny 8b4; do for p in 0 2; do
  timeout 200 $TR --master-port $((29500 + RANDOM % 400)) tools/tp_check.py --geom $g --proto $p 2> gpurun_out/s16_chk_${g}_$p.err | grep "tp=" | tee -a gpurun_out/s16_chk.log
done; done
timeout 200 $TR --master-port 29901 tools/tp_timeline.py > gpurun_out/s16_timeline_tp8.log 2>&1
timeout 200 $TR --master-port 29902 tools/tp_timeline.py --proto 2 > gpurun_out/s16_timeline_tp8_gather.log 2>&1
timeout 200 $TR --master-port 29903 tools/tp_timeline.py --streams 128 --gen 3 > gpurun_out/s16_timeline_tp8_s128.log 2>&1
timeout 400 $TR --master-port 29904 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s16_bench_n8.json 2> gpurun_out/s16_bench_n8.err
timeout 300 $TR --master-port 29905 bench.py --gpus 8 --steps 5 --warmup 3 --tp-proto 2 --streams 0 --no-parity > gpurun_out/s16_bench_n8_gather.json 2> gpurun_out/s16_bench_n8_gather.err
timeout 400 $TR --master-port 29906 tools/bench_70b_mixed.py --rate 8 > gpurun_out/s16_70b_mixed.json 2> gpurun_out/s16_70b_mixed.err
timeout 300 $TR --master-port 29907 tools/bench_70b_mixed.py --kv-frac 0.35 --rate 0 > gpurun_out/s16_70b_mixed_oversub.json 2> gpurun_out/s16_70b_mixed_oversub.err
grep -h "consumers" gpurun_out/s16_timeline_tp8.log gpurun_out/s16_timeline_tp8_gather.log | cut -c1-200
tail -c 600 gpurun_out/s16_bench_n8.json; tail -c 400 gpurun_out/s16_70b_mixed.json
