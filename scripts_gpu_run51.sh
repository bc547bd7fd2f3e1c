#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest -q -x --timeout 120 -p no:cacheprovider tests/test_ops_gpu.py -k "decode_attention" > gpurun_out/t_ops.log 2>&1; echo "attn ops rc=$?" > gpurun_out/summary.txt
timeout 300 python -m pytest -q --timeout 200 -p no:cacheprovider tests/test_engine_gpu.py -k "continuous or golden or batched" > gpurun_out/t_eng.log 2>&1; echo "engine rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 2 --warmup 2 --batch 64 --no-cpu-baseline --no-micro > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench b64 rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_attention_mma -s 70 -c 2 -o gpurun_out/prof_r1_decode_attn_mma3 -f python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline --no-micro --no-graphs > gpurun_out/ncu_attn.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/t_ops.log gpurun_out/t_eng.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_b64.json'))
print('b64 decode',round(d['value'],1),'frac',round(d['roofline']['frac'],3),d['roofline']['what'][-40:])
PY
