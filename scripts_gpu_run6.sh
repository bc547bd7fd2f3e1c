#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q --timeout 180 -p no:cacheprovider"
$T tests/test_ops_gpu.py -k "gemv" > gpurun_out/t1_ops.log 2>&1; echo "ops rc=$?" > gpurun_out/summary.txt
$T tests/test_engine_gpu.py > gpurun_out/t2_engine.log 2>&1; echo "engine rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
LLMLB_GEMV_NO_BULK=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nobulk.json 2> gpurun_out/bench_nobulk.err; echo "bench nobulk rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/t1_ops.log gpurun_out/t2_engine.log; do echo "== $f"; tail -n 12 $f; done
python - <<'PY'
import json
for f in ['bench','bench_nobulk']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'decode',round(d['value'],1),'prefill',round(d['prefill']['value']),'frac',round(d['roofline']['frac'],3),d['roofline']['what'][-60:],'kernel',d['roofline']['kernel'] and round(d['roofline']['kernel']['frac'],3), d['clocks'])
    except Exception as e: print(f,'ERR',e)
PY
tail -n 5 gpurun_out/bench.err
