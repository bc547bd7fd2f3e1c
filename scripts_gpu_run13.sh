#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest -q -x --timeout 120 -p no:cacheprovider tests/test_ops_gpu.py -k "gemv" > gpurun_out/t_ops.log 2>&1; echo "gemv ops rc=$?" > gpurun_out/summary.txt
timeout 300 python -m pytest -q --timeout 120 -p no:cacheprovider tests/test_engine_gpu.py -k "not 8b" > gpurun_out/t_eng.log 2>&1; echo "engine rc=$?" >> gpurun_out/summary.txt
for mb in 0 24 48; do
LLMLB_PF_MB=$mb timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench_pf$mb.json 2> gpurun_out/bench_pf$mb.err; echo "bench pf$mb rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -n 4 gpurun_out/t_ops.log gpurun_out/t_eng.log
python - <<'PY'
import json
for f in ['bench_pf0','bench_pf24','bench_pf48']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'decode',round(d['value'],1),'frac',round(d['roofline']['frac'],3),d['roofline']['what'][-50:])
    except Exception as e: print(f,'ERR',e)
PY
