#!/bin/bash
# round-2 GPU session 3 (2 GPUs): TP parity (both decode protocols), protocol B changes, timelines, bench N=2
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_gpu.py -q -p no:cacheprovider > gpurun_out/s3_tp.log 2>&1
echo "tp rc=$?" >> gpurun_out/s3_tp.log
timeout 600 python -m pytest tests/test_engine_gpu.py -q -p no:cacheprovider -k "70b or 8b_batched or oversubscription or preempted or queue_limit or many_short" > gpurun_out/s3_engine.log 2>&1
echo "engine rc=$?" >> gpurun_out/s3_engine.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/tp_timeline.py > gpurun_out/s3_timeline_tp2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/tp_timeline.py --streams 64 --gen 2 > gpurun_out/s3_timeline_tp2_s64.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s3_bench_n2.json 2> gpurun_out/s3_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 3 --warmup 3 --tp-proto 1 --streams 0 --no-parity > gpurun_out/s3_bench_n2_flags.json 2> gpurun_out/s3_bench_n2_flags.err
tail -3 gpurun_out/s3_tp.log; tail -3 gpurun_out/s3_engine.log
