#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest -q -x --timeout 120 -p no:cacheprovider tests/test_ops_gpu.py -k "gemm" > gpurun_out/t_ops.log 2>&1; echo "gemm ops rc=$?" > gpurun_out/summary.txt
timeout 600 python -m pytest -q --timeout 200 -p no:cacheprovider tests/test_engine_gpu.py > gpurun_out/t_eng.log 2>&1; echo "engine rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 2 --warmup 2 --batch 64 --no-cpu-baseline --no-micro > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench b64 rc=$?" >> gpurun_out/summary.txt
LLMLB_GEMM_NO_SK=1 timeout 300 python bench.py --steps 2 --warmup 2 --batch 64 --no-cpu-baseline --no-micro > gpurun_out/bench_b64_nosk.json 2> gpurun_out/bench_b64_nosk.err; echo "bench b64 nosk rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 2 --warmup 2 --batch 16 --no-cpu-baseline --no-micro > gpurun_out/bench_b16.json 2> gpurun_out/bench_b16.err; echo "bench b16 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 6 gpurun_out/t_ops.log gpurun_out/t_eng.log
python - <<'PY'
import json
for f in ['bench_b64','bench_b64_nosk','bench_b16']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'decode',round(d['value'],1),'frac',round(d['roofline']['frac'],3),d['roofline']['what'][-40:],'prefill',round(d['prefill']['value']))
    except Exception as e: print(f,'ERR',e)
PY
