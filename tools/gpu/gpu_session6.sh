#!/bin/bash
# round-2 GPU session 6 (2 GPUs): which change moved the tp=2 numerics at 8B widths (A/B by switch), reduce_norm with loads in flight,
# tensor-core decode attention for wide tp batches
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider -k "ksplit" -x > gpurun_out/s6_gemm.log 2>&1
echo "gemm rc=$?" >> gpurun_out/s6_gemm.log
timeout 200 $TR --master-port 29511 tools/tp_check.py --geom 8b4 --proto 0 > gpurun_out/s6_chk_default.log 2>&1; echo "rc=$?" >> gpurun_out/s6_chk_default.log
LLMLB_DEBUG_NO_KSPLIT=1 timeout 200 $TR --master-port 29512 tools/tp_check.py --geom 8b4 --proto 0 > gpurun_out/s6_chk_noksplit.log 2>&1; echo "rc=$?" >> gpurun_out/s6_chk_noksplit.log
LLMLB_DEBUG_NO_AGWAIT=1 timeout 200 $TR --master-port 29513 tools/tp_check.py --geom 8b4 --proto 0 > gpurun_out/s6_chk_noagwait.log 2>&1; echo "rc=$?" >> gpurun_out/s6_chk_noagwait.log
LLMLB_DEBUG_NO_AGWAIT=1 LLMLB_DEBUG_NO_KSPLIT=1 timeout 200 $TR --master-port 29514 tools/tp_check.py --geom 8b4 --proto 0 > gpurun_out/s6_chk_neither.log 2>&1; echo "rc=$?" >> gpurun_out/s6_chk_neither.log
timeout 200 $TR --master-port 29515 tools/tp_check.py --geom 8b4 --proto 2 > gpurun_out/s6_chk_gather.log 2>&1; echo "rc=$?" >> gpurun_out/s6_chk_gather.log
timeout 200 $TR --master-port 29516 tools/tp_check.py --geom 8b4 --proto 1 > gpurun_out/s6_chk_flags.log 2>&1; echo "rc=$?" >> gpurun_out/s6_chk_flags.log
timeout 300 $TR --master-port 29517 tools/tp_timeline.py > gpurun_out/s6_timeline_tp2.log 2>&1
timeout 300 $TR --master-port 29518 tools/tp_timeline.py --proto 2 > gpurun_out/s6_timeline_tp2_gather.log 2>&1
timeout 300 $TR --master-port 29519 tools/tp_timeline.py --streams 64 --gen 3 > gpurun_out/s6_timeline_tp2_s64.log 2>&1
timeout 600 $TR --master-port 29520 bench.py --gpus 2 --steps 5 --warmup 3 --no-ref-shape > gpurun_out/s6_bench_n2.json 2> gpurun_out/s6_bench_n2.err
grep -h "tp=2 geom" gpurun_out/s6_chk_*.log; tail -2 gpurun_out/s6_gemm.log; tail -c 400 gpurun_out/s6_bench_n2.json
