#!/bin/bash
# round-2 GPU session 17 (1 GPU): the whole GPU suite, the default bench line, ncu launch list + full captures
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest -q -p no:cacheprovider tests -m gpu -x > gpurun_out/s17_suite.log 2>&1
echo "gpu suite rc=$?" >> gpurun_out/s17_suite.log
tail -4 gpurun_out/s17_suite.log
timeout 600 python bench.py > gpurun_out/s17_bench_n1.json 2> gpurun_out/s17_bench_n1.err
tail -c 700 gpurun_out/s17_bench_n1.json
timeout 300 python __graft_entry__.py --smoke > gpurun_out/s17_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s17_smoke.log
# ncu: launch list of a short run (one prefill + decode steps; under a profiler: shares only, never bench values)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/s17_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-micro --streams 0 --no-ref-shape > gpurun_out/s17_ncu_b.log 2>&1
python tools/launches.py gpurun_out/s17_launches.csv > gpurun_out/s17_launches_summary.txt 2>&1
# ncu --set full of the prefill attention (tcgen05) and the K-split GEMM of a 64-stream step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_attention_kernel_tc -s 40 -c 1 -o gpurun_out/s17_attn_tc \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-micro --streams 0 --no-ref-shape > gpurun_out/s17_ncu_attn.log 2>&1
ncu -i gpurun_out/s17_attn_tc.ncu-rep --page details > gpurun_out/s17_attn_tc_details.txt 2>&1
head -30 gpurun_out/s17_launches_summary.txt
