#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest -q --timeout 300 -p no:cacheprovider tests -m gpu > gpurun_out/t_all_pdl2.log 2>&1; echo "gpu suite (defaults) rc=$?" > gpurun_out/summary.txt
timeout 200 python bench.py --steps 2 --warmup 2 --batch 64 --no-cpu-baseline --no-micro > gpurun_out/bench_b64_all.json 2> gpurun_out/bench_b64_all.err; echo "b64 all rc=$?" >> gpurun_out/summary.txt
LLMLB_NORM_NO_PDL=1 timeout 200 python bench.py --steps 2 --warmup 2 --batch 64 --no-cpu-baseline --no-micro > gpurun_out/bench_b64_nonorm.json 2> gpurun_out/bench_b64_nonorm.err; echo "b64 no-norm-pdl rc=$?" >> gpurun_out/summary.txt
LLMLB_ATTN_NO_TRIGGER=1 timeout 200 python bench.py --steps 2 --warmup 2 --batch 64 --no-cpu-baseline --no-micro > gpurun_out/bench_b64_noattn.json 2> gpurun_out/bench_b64_noattn.err; echo "b64 no-attn-trigger rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_all_pdl2.log
python - <<'PY'
import json
for f in ['bench_b64_all','bench_b64_nonorm','bench_b64_noattn']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'decode',round(d['value'],1),'frac',round(d['roofline']['frac'],3),d['roofline']['what'][-40:],'prefill',round(d['prefill']['value']))
    except Exception as e: print(f,'ERR',e)
PY
