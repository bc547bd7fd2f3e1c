#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py > gpurun_out/tp_check.log 2>&1; echo "tp_check rc=$?" > gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_tp$N.json 2> gpurun_out/bench_tp$N.err; echo "bench tp rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 8 gpurun_out/tp_check.log; cut -c1-400 gpurun_out/bench_tp$N.json; tail -n 12 gpurun_out/bench_tp$N.err
