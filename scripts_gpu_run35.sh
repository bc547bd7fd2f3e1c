#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest -q -x --timeout 120 -p no:cacheprovider tests/test_ops_gpu.py -k "gemm" > gpurun_out/t_ops.log 2>&1; echo "gemm ops rc=$?" > gpurun_out/summary.txt
timeout 300 python tools/gemm_prefill_bench.py 512 2048 > gpurun_out/gemm_prefill_bench.txt 2>&1; echo "prefill microbench rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest -q --timeout 200 -p no:cacheprovider tests/test_engine_gpu.py > gpurun_out/t_eng.log 2>&1; echo "engine rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench_sk.json 2> gpurun_out/bench_sk.err; echo "bench rc=$?" >> gpurun_out/summary.txt
LLMLB_GEMM_NO_SK=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench_nosk.json 2> gpurun_out/bench_nosk.err; echo "bench nosk rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 5 gpurun_out/t_ops.log gpurun_out/t_eng.log; cat gpurun_out/gemm_prefill_bench.txt
python - <<'PY'
import json
for f in ['bench_sk','bench_nosk']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'decode',round(d['value'],1),'prefill',round(d['prefill']['value']),'frac',round(d['prefill']['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
