#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
R=$(pwd)
mkdir -p gpurun_out
for i in 1 2; do
( cd $R/_ab_old && timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > $R/gpurun_out/ab_old$i.json 2> $R/gpurun_out/ab_old$i.err )
( cd $R && timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-micro > $R/gpurun_out/ab_new$i.json 2> $R/gpurun_out/ab_new$i.err )
done
python - <<'PY'
import json
for f in ['ab_old1','ab_new1','ab_old2','ab_new2']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'decode',round(d['value'],1),'prefill',round(d['prefill']['value']),'frac',round(d['prefill']['roofline']['frac'],3), d['clocks'])
    except Exception as e: print(f,'ERR',e)
PY
