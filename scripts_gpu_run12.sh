#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest -q -x --timeout 120 -p no:cacheprovider tests/test_engine_gpu.py -k "chain" > gpurun_out/t_chain.log 2>&1; echo "chain test rc=$?" > gpurun_out/summary.txt
timeout 600 python -m pytest -q --timeout 300 -p no:cacheprovider tests/test_engine_gpu.py -k "8b" > gpurun_out/t_8b.log 2>&1; echo "8b tests rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
LLMLB_DECODE_CHAIN=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench_nochain.json 2> gpurun_out/bench_nochain.err; echo "bench nochain rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 12 gpurun_out/t_chain.log; tail -n 8 gpurun_out/t_8b.log
python - <<'PY'
import json
for f in ['bench','bench_nochain']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'decode',round(d['value'],1),'prefill',round(d['prefill']['value']),'frac',round(d['roofline']['frac'],3),d['roofline']['what'][-50:],'launches',d['gpu_launches'])
    except Exception as e: print(f,'ERR',e)
PY
tail -n 3 gpurun_out/bench.err
