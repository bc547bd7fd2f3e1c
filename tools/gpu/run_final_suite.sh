#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest -q --timeout 300 -p no:cacheprovider tests -m gpu > gpurun_out/t_all_final.log 2>&1; echo "gpu suite rc=$?" > gpurun_out/summary.txt
timeout 120 python bench.py --steps 2 --warmup 2 --batch 64 --no-cpu-baseline --no-micro > gpurun_out/bench_b64_final.json 2> gpurun_out/bench_b64_final.err; echo "b64 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_all_final.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_b64_final.json'))
print('b64 decode',round(d['value'],1),'frac',round(d['roofline']['frac'],3),d['roofline']['what'][-40:])
PY
