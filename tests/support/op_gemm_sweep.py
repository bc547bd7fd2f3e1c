"""llmlb_op_gemm (the kernel-level entry point of the tensor-core projections) over random shapes, on the fake CUDA runtime:
every call must be accepted and its launches logged (FAKE_CUDART_LAUNCH_LOG).  usage: op_gemm_sweep.py <calls> <seed>"""
import ctypes
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from llmlb_b200 import ffi  # noqa: E402

L = ffi.lib()
cu = ctypes.CDLL("libcudart.so.12")
cu.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]


def dmalloc(n):
    p = ctypes.c_void_p()
    assert cu.cudaMalloc(ctypes.byref(p), n) == 0
    return p


w, x, out = dmalloc(1 << 33), dmalloc(1 << 30), dmalloc(1 << 31)          # address space only: nothing is executed
rs = random.Random(int(sys.argv[2]))
LOG = os.environ["FAKE_CUDART_LAUNCH_LOG"]
bad = 0
for _ in range(int(sys.argv[1])):
    T = rs.choice([rs.randint(1, 8), rs.randint(1, 300), rs.randint(1, 2048)])
    n_out = rs.choice([128, 256, 384, 1024, 2048, 4096, 6144, 8192, 14336, 28672, 128256, 2 * rs.randint(1, 20000), 8 * rs.randint(1, 4000)])
    k = 8 * rs.choice([8, 16, 64, 128, 512, 1024, 1792, 3584, rs.randint(1, 2048)])
    epi = rs.randint(0, 3)
    if epi == 2 and n_out % 2:
        continue
    with open(LOG, "a") as f:
        f.write("### T=%d n_out=%d k=%d epi=%d\n" % (T, n_out, k, epi))
    rc = L.llmlb_op_gemm(w, x, out, T, n_out, k, epi, n_out if epi != 2 else n_out // 2, 0, None)
    if rc != 0:
        bad += 1
        print("rc", rc, T, n_out, k, epi, ffi.last_error())
sys.exit(1 if bad else 0)
