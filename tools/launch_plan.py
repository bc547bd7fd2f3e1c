"""The engine's launch plan for one step, read WITHOUT a GPU: runs the product library over tests/support/fake_cudart.cpp
(a test double of libcudart that logs every launch and executes nothing) and prints, per kernel, how often it is launched in
the step and with which grid / block / dynamic shared memory / cluster / PDL attribute.  Shapes only — no timing.

    python tools/launch_plan.py 8b 1 prefill 512          # model, tp, step kind, width (tokens or batch)
    python tools/launch_plan.py 8b 8 decode 128 --layers 32
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CHILD = r'''
import os, sys
sys.path.insert(0, %(root)r)
from llmlb_b200 import ffi
model = dict(ffi.LLAMA3_8B if %(model)r == "8b" else ffi.LLAMA3_70B, n_layers=%(layers)d)
tp, kind, width = %(tp)d, %(kind)r, %(width)d
e = ffi.Engine(model, tp_rank=0, tp_size=tp, max_seqs=max(8, width if kind == "decode" else 8), max_ctx=4096, max_step_tokens=2048)
others = [ffi.Engine(model, tp_rank=r, tp_size=tp, max_seqs=8, max_ctx=256) for r in range(1, tp)]
if tp > 1:
    hs = [x.tp_export() for x in [e] + others]
    for x in [e] + others:
        x.tp_import(hs)
def drain(rid):
    while True:
        ev = e.poll(rid, timeout_ms=-1)
        if ev and ev[-1]["finish_reason"]:
            break
    e.release(rid)
def mark(s):
    open(os.environ["FAKE_CUDART_LAUNCH_LOG"], "a").write("### " + s + "\n")
if kind == "prefill":
    mark("step")
    drain(e.submit([1] * width, 1, ignore_eos=True))
else:
    e.pause(True)
    rids = [e.submit([1] * 8, 6, ignore_eos=True) for _ in range(width)]
    e.pause(False)
    for r in rids:
        drain(r)
    # the width's graph is captured by now: one more batch, and the LAST capture-free decode step is what we want; the log
    # cannot see graph replays (one cudaGraphLaunch), so report the captured launches of this width instead
mark("end")
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model", choices=["8b", "70b"])
    ap.add_argument("tp", type=int)
    ap.add_argument("kind", choices=["prefill", "decode"])
    ap.add_argument("width", type=int)
    ap.add_argument("--layers", type=int, default=2, help="layers to instantiate (launch shapes do not depend on depth)")
    a = ap.parse_args()
    import test_engine_host_logic_cpu as HL
    with tempfile.TemporaryDirectory() as d:
        log = os.path.join(d, "launch.log")
        env = dict(os.environ, LD_PRELOAD=HL.build_fake(), FAKE_CUDART_LAUNCH_LOG=log)
        code = CHILD % dict(root=ROOT, model=a.model, layers=a.layers, tp=a.tp, kind=a.kind, width=a.width)
        subprocess.run([sys.executable, "-c", code], env=env, check=True)
        rows, on = [], a.kind == "decode"
        for line in open(log):
            if line.startswith("###"):
                on = line.strip() == "### step" or (a.kind == "decode" and not line.strip() == "### end")
                continue
            if on:
                rows.append(line.split())
    if a.kind == "decode":
        # decode steps run as CUDA graphs: what the log holds is the CAPTURE of each batch width (at warm-up or on first use);
        # pick the captured step whose embedding gather has `width` rows
        starts = [i for i, r in enumerate(rows) if "decode_prepare_kernel" in r[0]]
        ends = [i for i, r in enumerate(rows) if "step_finish_kernel" in r[0]]
        best = None
        for s in starts:
            e = next((x for x in ends if x > s), None)
            if e is not None and any("embed_kernel" in r[0] and int(r[1]) == a.width for r in rows[s:e]):
                best = (s, e)
        if best is None:
            sys.exit("no decode step of width %d was captured" % a.width)
        rows = rows[best[0]:best[1] + 1]
    cnt = collections.OrderedDict()
    for r in rows:
        name = subprocess.run(["c++filt", r[0]], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("llmlb::", "")
        key = (name, "x".join(r[1:4]), "x".join(r[4:7]), r[7], "x".join(r[8:11]), r[11])
        cnt[key] = cnt.get(key, 0) + 1
    print("launch plan: Llama-3-%s tp=%d (rank 0), %s step, width %d, %d layers — %d launches (shapes only; nothing was executed)" %
          (a.model.upper(), a.tp, a.kind, a.width, a.layers, len(rows)))
    print("%5s  %-58s %-12s %-10s %8s %-8s %s" % ("count", "kernel", "grid", "block", "smem", "cluster", "pdl"))
    for (name, g, b, smem, cl, pdl), n in cnt.items():
        print("%5d  %-58s %-12s %-10s %8s %-8s %s" % (n, name[:58], g, b, smem, cl, pdl))


if __name__ == "__main__":
    main()
