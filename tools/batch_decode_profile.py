"""Batched decode (64 concurrent streams): per-kernel CTA lifetimes from the debug trace hooks.
GEMM records carry clock64 cycles + role stalls, attention records globaltimer stamps."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmlb_b200 import ffi  # noqa: E402

NB = int(os.environ.get("NB", "64"))


def main():
    L = ffi.lib()
    L.llmlb_debug_trace_enable.argtypes = [C.c_uint32]
    L.llmlb_debug_trace_dump.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    model = ffi.LLAMA3_8B
    eng = ffi.Engine(model, max_seqs=NB, max_ctx=1024)
    rs = np.random.RandomState(0)
    prompts = [rs.randint(0, model["vocab"], 512).tolist() for _ in range(NB)]

    def run(gen):
        rids = [eng.submit(p, gen, ignore_eos=True) for p in prompts]
        for r in rids:
            while True:
                ev = eng.poll(r, timeout_ms=-1)
                if ev and ev[-1]["finish_reason"]:
                    break
            eng.release(r)
    run(4)
    h0 = eng.health()
    cap = 1500000
    L.llmlb_debug_trace_enable(cap)
    run(6)
    buf = np.zeros((cap, 6), dtype=np.uint64)
    n = C.c_uint32()
    L.llmlb_debug_trace_dump(buf.ctypes.data, cap, C.byref(n))
    L.llmlb_debug_trace_enable(0)
    h1 = eng.health()
    eng.close()
    dsteps = h1["steps_decode"] - h0["steps_decode"]
    print("decode steps %d, %.3f ms/step (CUDA events)" % (dsteps, (h1["gpu_ms_decode"] - h0["gpu_ms_decode"]) / dsteps))
    r = buf[: n.value]
    names = {0: "store_bf16", 1: "resid_f32", 2: "silu_mul", 3: "store_f32", 4: "partial_f32"}
    g = r[(r[:, 0] >> np.uint64(60)) == 2]
    print("%-40s %7s %9s %8s %9s %9s %9s" % ("gemm_tc (1-CTA) kernel", "ctas", "cycles", "~us", "prod_wait", "mma_wait", "epi_wait"))
    for tag in np.unique(g[:, 0]):
        rows = g[g[:, 0] == tag].astype(np.float64)
        t = int(tag)
        nm = "n_out=%d k=%d %s" % ((t >> 32) & 0xFFFFFFF, t & 0xFFFFFFF, names.get((t >> 28) & 0xF, "?"))
        tot = rows[:, 2]
        print("%-40s %7d %9.0f %8.1f %8.1f%% %8.1f%% %8.1f%%" % (nm, len(rows), tot.mean(), tot.mean() / 1.9e3, 100 * (rows[:, 3] / tot).mean(),
                                                               100 * (rows[:, 4] / tot).mean(), 100 * (rows[:, 5] / tot).mean()))
    k = r[(r[:, 0] >> np.uint64(60)) == 3]
    if len(k):
        print("%-40s %7s %8s %8s %8s %9s %9s %9s" % ("gemm_sk (stream-K) kernel", "ctas", "span_us", "life_us", "min_life", "prod_wait", "flag_wait", "epi_wait"))
    for tag in np.unique(k[:, 0]):
        rows = k[k[:, 0] == tag].astype(np.int64)
        t = int(tag)
        nm = "n_out=%d k=%d %s" % ((t >> 32) & 0xFFFFFFF, t & 0xFFFFFFF, names.get((t >> 28) & 0xF, "?"))
        # launches are separated in time: split the records of this tag into launches by start gaps
        order = np.argsort(rows[:, 2]); rows = rows[order]
        cuts = np.where(np.diff(rows[:, 2]) > 20000)[0] + 1
        spans, lifes, mins = [], [], []
        for grp in np.split(rows, cuts):
            spans.append((grp[:, 5].max() - grp[:, 2].min()) / 1e3)
            lifes.append((grp[:, 5] - grp[:, 2]).mean() / 1e3)
            mins.append((grp[:, 5] - grp[:, 2]).min() / 1e3)
        life_cyc = (rows[:, 5] - rows[:, 2]) * 1.9
        print("%-40s %7d %8.1f %8.1f %8.1f %8.1f%% %8.1f%% %8.1f%%" % (
            nm, len(rows) // max(1, len(spans)), np.mean(spans), np.mean(lifes), np.mean(mins),
            100 * ((rows[:, 3] >> 32) / life_cyc).mean(), 100 * ((rows[:, 3] & 0xFFFFFFFF) / life_cyc).mean(),
            100 * (rows[:, 4] / life_cyc).mean()))
    a = r[r[:, 0] == 1].astype(np.int64)
    if len(a):
        print("decode attention: %d CTAs, lifetime mean %.1f us max %.1f us; loop %.1f us; merge %.1f us" % (
            len(a), (a[:, 5] - a[:, 2]).mean() / 1e3, (a[:, 5] - a[:, 2]).max() / 1e3, (a[:, 4] - a[:, 3]).mean() / 1e3, (a[:, 5] - a[:, 4]).mean() / 1e3))


if __name__ == "__main__":
    main()
