#!/bin/bash
# round-2 GPU session 4 (8 GPUs): TP parity at 4 and 8, bench N=8 (+128 streams), timelines of a TP=8 decode and prefill step
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/s4_topo.txt 2>&1
timeout 900 python -m pytest tests/test_tp_gpu.py -q -p no:cacheprovider -k "8-tiny-pairs or 8-odd-pairs or 4-tiny-pairs or 8-tiny-flags" > gpurun_out/s4_tp.log 2>&1
echo "tp rc=$?" >> gpurun_out/s4_tp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/tp_timeline.py > gpurun_out/s4_timeline_tp8.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s4_bench_n8.json 2> gpurun_out/s4_bench_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 5 --warmup 3 --streams 0 > gpurun_out/s4_bench_n4.json 2> gpurun_out/s4_bench_n4.err
tail -3 gpurun_out/s4_tp.log; tail -c 1500 gpurun_out/s4_bench_n8.json
