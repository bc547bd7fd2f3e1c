#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest -q --timeout 120 -p no:cacheprovider tests/test_engine_gpu.py -k "not 8b" > gpurun_out/t_eng.log 2>&1; echo "engine rc=$?" > gpurun_out/summary.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
LLMLB_GEMV_DYN=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench_static.json 2> gpurun_out/bench_static.err; echo "bench static rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/decode_timeline.py > gpurun_out/timeline.txt 2> gpurun_out/timeline.err
timeout 300 python tools/gemm_stalls.py > gpurun_out/gemm_stalls.txt 2> gpurun_out/gemm_stalls.err
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_eng.log
python - <<'PY'
import json
for f in ['bench','bench_static']:
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f,'decode',round(d['value'],1),'frac',round(d['roofline']['frac'],3),d['roofline']['what'][-50:])
PY
head -n 8 gpurun_out/timeline.txt; tail -n 9 gpurun_out/timeline.txt; cat gpurun_out/gemm_stalls.txt; tail -n 3 gpurun_out/gemm_stalls.err
