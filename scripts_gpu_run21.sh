#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest -q -x --timeout 60 -p no:cacheprovider tests/test_ops_gpu.py -k "test_gemm and tc" > gpurun_out/t_ops.log 2>&1; echo "gemm_tc rc=$?" > gpurun_out/summary.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/gemm_stalls.py > gpurun_out/gemm_stalls.txt 2> gpurun_out/gemm_stalls.err
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_ops.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('decode',round(d['value'],1),'prefill',round(d['prefill']['value']),'pf frac',round(d['prefill']['roofline']['frac'],3),'frac',round(d['roofline']['frac'],3))
PY
cat gpurun_out/gemm_stalls.txt
