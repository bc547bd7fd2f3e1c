#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LLMLB_GEMM_PDL=1 timeout 600 python -m pytest -q --timeout 300 -p no:cacheprovider tests -m gpu > gpurun_out/t_all_pdl.log 2>&1; echo "gpu suite (PDL on) rc=$?" > gpurun_out/summary.txt
for m in off on; do
  if [ $m = on ]; then export LLMLB_GEMM_PDL=1; else unset LLMLB_GEMM_PDL; fi
  timeout 200 python bench.py --steps 2 --warmup 2 --batch 64 --no-cpu-baseline --no-micro > gpurun_out/bench_b64_pdl_$m.json 2> gpurun_out/bench_b64_pdl_$m.err; echo "b64 pdl $m rc=$?" >> gpurun_out/summary.txt
  timeout 200 python bench.py --steps 2 --warmup 2 --batch 16 --no-cpu-baseline --no-micro > gpurun_out/bench_b16_pdl_$m.json 2> gpurun_out/bench_b16_pdl_$m.err; echo "b16 pdl $m rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -n 4 gpurun_out/t_all_pdl.log
python - <<'PY'
import json
for f in ['bench_b64_pdl_off','bench_b64_pdl_on','bench_b16_pdl_off','bench_b16_pdl_on']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'decode',round(d['value'],1),'frac',round(d['roofline']['frac'],3),d['roofline']['what'][-40:],'prefill',round(d['prefill']['value']))
    except Exception as e: print(f,'ERR',e)
PY
