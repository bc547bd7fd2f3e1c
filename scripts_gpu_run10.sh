#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest -q --timeout 600 -p no:cacheprovider tests/test_engine_gpu.py -k "8b" ) > gpurun_out/t_8b.log 2>&1; echo "8b tests rc=$?" > gpurun_out/summary.txt
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?" >> gpurun_out/summary.txt
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 12 gpurun_out/t_8b.log; tail -n 4 gpurun_out/bench_default.err gpurun_out/bench_ref.err
python - <<'PY'
import json
for f in ['bench_default','bench_ref']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'value',d['value'],'prefill',d.get('prefill',{}).get('value'),'cpu',d.get('cpu_baseline'))
    except Exception as e: print(f,'ERR',e)
PY
