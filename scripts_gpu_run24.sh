#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py > gpurun_out/tp_check.log 2>&1; echo "tp_check rc=$?" > gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_tp$N.json 2> gpurun_out/bench_tp$N.err; echo "bench tp rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_tp$N.json 2> gpurun_out/bench_ref_tp$N.err; echo "bench ref tp rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep "tp=" gpurun_out/tp_check.log
python - <<'PY'
import json
for f in ['bench_tp2','bench_ref_tp2']:
    txt=open('gpurun_out/%s.json'%f).read().strip().splitlines()
    print(f,'stdout lines',len(txt))
    d=json.loads(txt[-1]); print('  value',d['value'],'prefill',d.get('prefill',{}).get('value'),'e2e',d['e2e']['value'],'n_gpus',d['n_gpus'])
PY
