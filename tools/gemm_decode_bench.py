"""Batched-decode projection GEMMs as weight streams: time llmlb_op_gemm (impl 0 = slab tiling,
impl 2 = stream-K) and llmlb_op_gemv on the Llama-3-8B layer shapes at small token counts, rotating
over > L2 worth of weight copies, and report GB/s of weight bytes against the measured HBM peak.

    python tools/gemm_decode_bench.py [T ...]         (default 16 64 128)
"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmlb_b200 import ffi  # noqa: E402

SHAPES = [("qkv", 6144, 4096, ffi.EPI_STORE_BF16), ("o", 4096, 4096, ffi.EPI_STORE_F32),
          ("gate_up", 28672, 4096, ffi.EPI_SILU_MUL), ("down", 4096, 14336, ffi.EPI_STORE_F32),
          ("lm_head", 128256, 4096, ffi.EPI_STORE_F32)]


def peak():
    try:
        return json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6501.2


def main():
    Ts = [int(a) for a in sys.argv[1:]] or [16, 64, 128]
    L = ffi.lib()
    st = None
    vp = lambda t: C.c_void_p(t.data_ptr())
    pk = peak()
    print("%-8s %4s %-8s %9s %9s %6s" % ("shape", "T", "impl", "us", "GB/s", "frac"))
    only = os.environ.get("SHAPES")
    for name, n, k, epi in SHAPES:
        if only and name not in only.split(","):
            continue
        copies = max(2, int(400e6 // (n * k * 2)) + 1)
        w = [torch.empty(n, k, dtype=torch.bfloat16, device="cuda").normal_(0, 0.02) for _ in range(copies)]
        for T in Ts:
            x = torch.randn(T, k, device="cuda").bfloat16()
            cols = n // 2 if epi == ffi.EPI_SILU_MUL else n
            out = torch.zeros(T, cols, dtype=torch.bfloat16 if epi in (ffi.EPI_STORE_BF16, ffi.EPI_SILU_MUL) else torch.float32, device="cuda")
            for impl in (0, 2):
                if impl == 2 and T > 128:
                    continue

                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

                def run(i):
                    ffi.check(L.llmlb_op_gemm(vp(w[i % copies]), vp(x), vp(out), T, n, k, epi, cols, impl, st))
                for i in range(4):
                    run(i)
                torch.cuda.synchronize()
                iters = 30
                # one CUDA graph of `iters` launches: no host time between kernels (the engine's decode
                # step is a graph too)
                side = torch.cuda.Stream()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    st = C.c_void_p(side.cuda_stream)
                    run(0)
                    side.synchronize()
                    with torch.cuda.graph(gr, stream=side):
                        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                        for i in range(iters):
                            run(i)
                    gr.replay()
                    side.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(side)
                    gr.replay()
                    b.record(side)
                    side.synchronize()
                us = a.elapsed_time(b) * 1e3 / iters
                gbs = n * k * 2 / us / 1e3
                print("%-8s %4d %-8s %9.2f %9.1f %6.3f" % (name, T, {0: "slab", 2: "streamk"}[impl], us, gbs, gbs / pk))
        del w
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
