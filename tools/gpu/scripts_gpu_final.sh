#!/bin/bash
# what the driver runs at round end, in one go
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 600 -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" > gpurun_out/summary.txt
( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?" >> gpurun_out/summary.txt
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 8 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/smoke.log; tail -n 4 gpurun_out/bench_reference.err gpurun_out/bench_default.err
python - <<'PY'
import json
for f in ['bench_reference','bench_default']:
    try:
        txt=open('gpurun_out/%s.json'%f).read().strip().splitlines()
        print(f,'lines on stdout:',len(txt))
        d=json.loads(txt[-1])
        print('  value',d['value'],'prefill',d.get('prefill',{}).get('value'),'e2e',d['e2e']['value'],'cpu',d.get('cpu_baseline',{}).get('value'),'roof',d.get('roofline',{}).get('frac'),'launches',d.get('gpu_launches'),'clocks',d.get('clocks'))
    except Exception as e: print(f,'ERR',e)
PY
