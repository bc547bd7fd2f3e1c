#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
R=$(pwd)
mkdir -p gpurun_out
timeout 900 python -m pytest -q --timeout 300 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "gpu suite rc=$?" > gpurun_out/summary.txt
( cd $R/_ab_old && timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > $R/gpurun_out/ab_old1.json 2> $R/gpurun_out/ab_old1.err )
( cd $R && timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > $R/gpurun_out/ab_new1.json 2> $R/gpurun_out/ab_new1.err )
( cd $R && timeout 300 python bench.py --steps 2 --warmup 2 --batch 64 --no-cpu-baseline --no-micro > $R/gpurun_out/ab_new_b64.json 2> $R/gpurun_out/ab_new_b64.err )
cat gpurun_out/summary.txt; tail -n 4 gpurun_out/t_all.log
python - <<'PY'
import json
for f in ['ab_old1','ab_new1','ab_new_b64']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'decode',round(d['value'],1),'prefill',round(d['prefill']['value']),'frac',round(d['prefill']['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
