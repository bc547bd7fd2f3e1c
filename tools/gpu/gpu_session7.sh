#!/bin/bash
# round-2 GPU session 7 (2 GPUs): bench.py's tp parity record under each switch
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 1 --warmup 0 --streams 0 --no-ref-shape"
run() { name=$1; shift; timeout 200 "$@" > gpurun_out/s7_$name.json 2> gpurun_out/s7_$name.err; python - <<P
import json
for l in open("gpurun_out/s7_$name.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$name", d["value"], d["parity"])
P
}
run default $TR --master-port 29511 $B
LLMLB_DEBUG_NO_KSPLIT=1 run noksplit $TR --master-port 29512 $B
LLMLB_DEBUG_NO_AGWAIT=1 run noagwait $TR --master-port 29513 $B
LLMLB_DEBUG_NO_AGWAIT=1 LLMLB_DEBUG_NO_KSPLIT=1 run neither $TR --master-port 29514 $B
run nographs $TR --master-port 29515 $B --no-graphs
run flags $TR --master-port 29516 $B --tp-proto 1
