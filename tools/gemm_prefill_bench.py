"""Prefill-sized projection GEMMs (T = 512 by default): llmlb_op_gemm impl 0 (tiles) vs impl 2
(stream-K) on the Llama-3-8B layer shapes, one CUDA graph of launches per point, TFLOP/s against
the measured cuBLAS bf16 peak.

    python tools/gemm_prefill_bench.py [T ...]
"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmlb_b200 import ffi  # noqa: E402

SHAPES = [("qkv", 6144, 4096, ffi.EPI_STORE_BF16), ("o", 4096, 4096, ffi.EPI_STORE_F32),
          ("gate_up", 28672, 4096, ffi.EPI_SILU_MUL), ("down", 4096, 14336, ffi.EPI_STORE_F32)]


def peak():
    try:
        return json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"]
    except Exception:
        return 1686.6


def main():
    Ts = [int(a) for a in sys.argv[1:]] or [512]
    L = ffi.lib()
    vp = lambda t: C.c_void_p(t.data_ptr())
    pk = peak()
    print("%-8s %5s %-8s %9s %9s %6s" % ("shape", "T", "impl", "us", "TFLOP/s", "frac"))
    tot = {}
    for name, n, k, epi in SHAPES:
        # rotate over > 2x L2 of weight copies: in a real prefill every layer's weights are cold
        copies = max(2, int(400e6 // (n * k * 2)) + 1)
        w = [torch.empty(n, k, dtype=torch.bfloat16, device="cuda").normal_(0, 0.02) for _ in range(copies)]
        for T in Ts:
            x = torch.randn(T, k, device="cuda").bfloat16()
            cols = n // 2 if epi == ffi.EPI_SILU_MUL else n
            out = torch.zeros(T, cols, dtype=torch.bfloat16 if epi in (ffi.EPI_STORE_BF16, ffi.EPI_SILU_MUL) else torch.float32, device="cuda")
            for impl in (0, 2):
                side = torch.cuda.Stream()
                with torch.cuda.stream(side):
                    st = C.c_void_p(side.cuda_stream)
                    run = lambda i=0: ffi.check(L.llmlb_op_gemm(vp(w[i % copies]), vp(x), vp(out), T, n, k, epi, cols, impl, st))
                    for i in range(3):
                        run(i)
                    side.synchronize()
                    gr = torch.cuda.CUDAGraph()
                    iters = 20
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
                tf = 2.0 * T * n * k / us / 1e6
                tot[(T, impl)] = tot.get((T, impl), 0.0) + us
                print("%-8s %5d %-8s %9.2f %9.1f %6.3f" % (name, T, {0: "tiles", 2: "streamk"}[impl], us, tf, tf / pk))
    for (T, impl), us in sorted(tot.items()):
        print("layer GEMMs T=%d %-8s %.1f us" % (T, {0: "tiles", 2: "streamk"}[impl], us))


if __name__ == "__main__":
    main()
