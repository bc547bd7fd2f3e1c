#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest -q -x --timeout 120 -p no:cacheprovider tests/test_ops_gpu.py -k "gemm" > gpurun_out/t_ops.log 2>&1; echo "gemm ops rc=$?" > gpurun_out/summary.txt
timeout 300 python tools/gemm_prefill_bench.py 512 > gpurun_out/gemm_prefill_bench.txt 2>&1
timeout 300 python tools/gemm_pair_timeline.py 512 > gpurun_out/pair_timeline.txt 2>&1
timeout 300 python tools/gemm_pair_timeline.py 512 6144 4096 > gpurun_out/pair_timeline_qkv.txt 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench_vec.json 2> gpurun_out/bench_vec.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_ops.log; cat gpurun_out/gemm_prefill_bench.txt; grep "epilogue warp\|span" gpurun_out/pair_timeline.txt gpurun_out/pair_timeline_qkv.txt
python - <<'PY'
import json
for f in ['bench_vec']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'decode',round(d['value'],1),'prefill',round(d['prefill']['value']),'frac',round(d['prefill']['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
