#!/bin/bash
# round-2 GPU session 19 (1 GPU): ragged-wave tail launch for the prefill projections: parity at 8B widths, A/B of the prefill rate
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
LLMLB_PARITY_PROMPTS=2 timeout 400 python -m pytest tests/test_parity_8b_gpu.py -q -p no:cacheprovider -x -k "batch1 or 8_concurrent" > gpurun_out/s19_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/s19_parity.log; tail -3 gpurun_out/s19_parity.log
timeout 200 python -m pytest tests/test_engine_gpu.py -q -p no:cacheprovider -x -k "8b or chunked or batching" > gpurun_out/s19_engine.log 2>&1
echo "engine rc=$?" >> gpurun_out/s19_engine.log; tail -2 gpurun_out/s19_engine.log
B="bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity --no-micro --no-ref-shape --stream-steps 1"
timeout 200 python $B > gpurun_out/s19_bench_tail.json 2> gpurun_out/s19_bench_tail.err
LLMLB_DEBUG_NO_WAVE_TAIL=1 timeout 200 python $B > gpurun_out/s19_bench_notail.json 2> gpurun_out/s19_bench_notail.err
python - <<'P'
import json
for n in ("tail", "notail"):
    for l in open("gpurun_out/s19_bench_%s.json" % n):
        if l.startswith("{"):
            d = json.loads(l); print(n, "decode", round(d["value"], 1), "prefill", round(d["prefill"]["value"]), "streams prefill", round(d["streams"]["prefill_tok_s"]), "streams decode", round(d["streams"]["decode_tok_s"]), d["clocks"]["sm_mhz"])
P
