#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest -q --timeout 120 -p no:cacheprovider tests/test_engine_gpu.py -k "not 8b and tcgen05" > gpurun_out/t_eng.log 2>&1; echo "engine rc=$?" > gpurun_out/summary.txt
for mb in 64 0; do
LLMLB_ATTN_PF_MB=$mb timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench_apf$mb.json 2> gpurun_out/bench_apf$mb.err; echo "bench apf$mb rc=$?" >> gpurun_out/summary.txt
done
timeout 300 python tools/decode_timeline.py > gpurun_out/timeline.txt 2> gpurun_out/timeline.err
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_eng.log
python - <<'PY'
import json
for f in ['bench_apf64','bench_apf0']:
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f,'decode',round(d['value'],1),'frac',round(d['roofline']['frac'],3),d['roofline']['what'][-50:])
PY
head -n 8 gpurun_out/timeline.txt; tail -n 9 gpurun_out/timeline.txt
