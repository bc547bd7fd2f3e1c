#!/bin/bash
# round-2 GPU session 1 (2 GPUs): unit + TP parity, 8B parity (2 prompts), bench N=2 and N=1
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/s1_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_parity_8b_gpu.py -p no:cacheprovider > gpurun_out/s1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
LLMLB_PARITY_PROMPTS=2 timeout 600 python -m pytest tests/test_parity_8b_gpu.py -q -s -p no:cacheprovider > gpurun_out/s1_parity8b.log 2>&1
echo "parity rc=$?" >> gpurun_out/s1_parity8b.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s1_bench_n2.json 2> gpurun_out/s1_bench_n2.err
echo "bench n2 rc=$?" >> gpurun_out/s1_bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/s1_bench_n1.json 2> gpurun_out/s1_bench_n1.err
echo "bench n1 rc=$?" >> gpurun_out/s1_bench_n1.err
tail -3 gpurun_out/s1_pytest.log; tail -3 gpurun_out/s1_parity8b.log; tail -c 600 gpurun_out/s1_bench_n2.json; tail -2 gpurun_out/s1_bench_n2.err
