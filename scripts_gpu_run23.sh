#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -q -m gpu --timeout 600 -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" > gpurun_out/summary.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 2 --warmup 2 --batch 64 --no-cpu-baseline --no-micro > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench b64 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 12 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ['bench','bench_b64']:
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f,'decode',round(d['value'],1),'prefill',round(d['prefill']['value']),'pf frac',round(d['prefill']['roofline']['frac'],3),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value'],1))
PY
