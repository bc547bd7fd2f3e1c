#!/bin/bash
# round-2 GPU session 13 (2 GPUs): LL reduce-scatter for narrow steps
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_tp_gpu.py -q -p no:cacheprovider -k "pairs" > gpurun_out/s13_tp.log 2>&1
echo "tp rc=$?" >> gpurun_out/s13_tp.log
timeout 300 $TR --master-port 29519 tools/tp_timeline.py --streams 64 --gen 3 > gpurun_out/s13_timeline_tp2_s64.log 2>&1
timeout 600 $TR --master-port 29520 bench.py --gpus 2 --steps 3 --warmup 3 --no-ref-shape > gpurun_out/s13_bench_n2.json 2> gpurun_out/s13_bench_n2.err
LLMLB_DEBUG_NO_RSLL=1 timeout 600 $TR --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --no-ref-shape --no-parity > gpurun_out/s13_bench_n2_norsll.json 2> gpurun_out/s13_bench_n2_norsll.err
tail -3 gpurun_out/s13_tp.log
