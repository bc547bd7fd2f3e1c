#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q --timeout 180 -p no:cacheprovider"
$T tests/test_ops_gpu.py -k "gemv or decode_attention" > gpurun_out/t1_ops.log 2>&1; echo "ops rc=$?" > gpurun_out/summary.txt
$T tests/test_engine_gpu.py > gpurun_out/t2_engine.log 2>&1; echo "engine rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/t1_ops.log gpurun_out/t2_engine.log; do echo "== $f"; tail -n 12 $f; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(json.dumps({k:d[k] for k in ['value','ms_per_step','prefill','roofline','e2e','gpu_launches','clocks']},indent=1))
PY
tail -n 5 gpurun_out/bench.err
