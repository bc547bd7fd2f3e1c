#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python llmlb_b200/build.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest -q --timeout 300 -p no:cacheprovider tests/test_server_gpu.py > gpurun_out/t_server.log 2>&1; echo "server rc=$?" > gpurun_out/summary.txt
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full rc=$?" >> gpurun_out/summary.txt
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 2 --batch 64 --no-cpu-baseline --no-micro > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench b64 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 15 gpurun_out/t_server.log; tail -n 6 gpurun_out/bench_full.err gpurun_out/bench_ref.err gpurun_out/bench_b64.err
python - <<'PY'
import json
for f in ['bench_full','bench_ref','bench_b64']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, 'value',d['value'],'prefill',d.get('prefill',{}).get('value'),'e2e',d['e2e']['value'],'cpu',d.get('cpu_baseline'),'roof',d.get('roofline',{}).get('frac'))
    except Exception as e: print(f,'ERR',e)
PY
