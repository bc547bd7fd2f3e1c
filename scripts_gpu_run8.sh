#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q --timeout 180 -p no:cacheprovider"
$T tests/test_ops_gpu.py -k "test_gemm and tc" > gpurun_out/t1_ops.log 2>&1; echo "gemm_tc rc=$?" > gpurun_out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 2 --batch 64 --no-cpu-baseline --no-micro > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench b64 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 6 gpurun_out/t1_ops.log
python - <<'PY'
import json
for f in ['bench','bench_b64']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'decode',round(d['value'],1),'prefill',round(d['prefill']['value']),'pf frac',round(d['prefill']['roofline']['frac'],3),'frac',round(d['roofline']['frac'],3),d['roofline']['what'][-60:])
    except Exception as e: print(f,'ERR',e)
PY
