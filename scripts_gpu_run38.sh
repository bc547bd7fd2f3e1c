#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for pf in 0 16 32; do
echo "== LLMLB_GEMM2_PF=$pf" >> gpurun_out/gemm_prefill_pf.txt
LLMLB_GEMM2_PF=$pf timeout 300 python tools/gemm_prefill_bench.py 512 >> gpurun_out/gemm_prefill_pf.txt 2>&1
LLMLB_GEMM2_PF=$pf timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench_pf$pf.json 2> gpurun_out/bench_pf$pf.err
done
cat gpurun_out/gemm_prefill_pf.txt
python - <<'PY'
import json
for f in ['bench_pf0','bench_pf16','bench_pf32']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'decode',round(d['value'],1),'prefill',round(d['prefill']['value']),'frac',round(d['prefill']['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
