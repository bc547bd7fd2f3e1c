#!/bin/bash
# round-2 GPU session 5 (2 GPUs): in-kernel K-split GEMMs, owner-fold consumer GEMV, all-gather flag wait in the consumer GEMMs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider -k "gemm" -x > gpurun_out/s5_gemm.log 2>&1
echo "gemm rc=$?" >> gpurun_out/s5_gemm.log
timeout 900 python -m pytest tests/test_tp_gpu.py -q -p no:cacheprovider > gpurun_out/s5_tp.log 2>&1
echo "tp rc=$?" >> gpurun_out/s5_tp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/tp_timeline.py > gpurun_out/s5_timeline_tp2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/tp_timeline.py --streams 64 --gen 3 > gpurun_out/s5_timeline_tp2_s64.log 2>&1
timeout 300 python tools/tp_timeline.py --streams 64 --gen 3 > gpurun_out/s5_timeline_tp1_s64.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --no-ref-shape > gpurun_out/s5_bench_n2.json 2> gpurun_out/s5_bench_n2.err
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-ref-shape > gpurun_out/s5_bench_n1.json 2> gpurun_out/s5_bench_n1.err
tail -3 gpurun_out/s5_gemm.log; tail -3 gpurun_out/s5_tp.log; tail -c 600 gpurun_out/s5_bench_n2.json
