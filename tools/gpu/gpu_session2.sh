#!/bin/bash
# round-2 GPU session 2 (2 GPUs): tcgen05 prefill attention + PDL chain, determinism A/B, TP timelines
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "prefill_attention or sample_distribution or gemm_other" -p no:cacheprovider > gpurun_out/s2_attn.log 2>&1
echo "attn rc=$?" >> gpurun_out/s2_attn.log
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_server_gpu.py -q -p no:cacheprovider > gpurun_out/s2_engine.log 2>&1
echo "engine rc=$?" >> gpurun_out/s2_engine.log
timeout 300 python tools/determinism_probe.py > gpurun_out/s2_determinism_new.log 2>&1
PYTHONPATH=tools/_r1 timeout 300 python tools/determinism_probe.py > gpurun_out/s2_determinism_r1.log 2>&1
timeout 300 python tools/tp_timeline.py > gpurun_out/s2_timeline_tp1.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/tp_timeline.py > gpurun_out/s2_timeline_tp2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/tp_timeline.py --streams 64 > gpurun_out/s2_timeline_tp2_s64.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --streams 0 --no-ref-shape > gpurun_out/s2_bench_n1.json 2> gpurun_out/s2_bench_n1.err
tail -4 gpurun_out/s2_attn.log; tail -4 gpurun_out/s2_engine.log; cat gpurun_out/s2_determinism_new.log | tail -4; cat gpurun_out/s2_determinism_r1.log | tail -4
