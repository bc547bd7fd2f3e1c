"""The tensor-parallel plan channel's transport on CPU (no GPU): a leader process writes a few
thousand records — many times the 8 MiB ring, so wrap-around and back-pressure happen — and a
follower process reads them back through POSIX shared memory; both sides must agree on every
byte (checksum over type, length and payload)."""
import ctypes as C
import multiprocessing as mp
import os

from llmlb_b200 import build


def _side(rank, name, n, seed, q):
    lib = C.CDLL(build.build(), mode=C.RTLD_GLOBAL)
    lib.llmlb_debug_plan_ring.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.llmlb_last_error.restype = C.c_char_p
    cs = C.c_uint64()
    rc = lib.llmlb_debug_plan_ring(name.encode(), rank, n, seed, C.byref(cs))
    q.put((rank, rc, cs.value, lib.llmlb_last_error().decode() if rc else ""))


def test_plan_ring_two_processes():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = "llmlb_ring_test_%d" % os.getpid()
    n, seed = 1500, 12345          # ~1500 x 28 KiB average = ~42 MiB through an 8 MiB ring
    procs = [ctx.Process(target=_side, args=(r, name, n, seed, q)) for r in (0, 1)]
    procs[0].start()                # the leader creates the ring; the follower waits for it to appear
    procs[1].start()
    got = dict()
    for _ in range(2):
        rank, rc, cs, err = q.get(timeout=120)
        got[rank] = (rc, cs, err)
    for p in procs:
        p.join(timeout=30)
    assert got[0][0] == 0 and got[1][0] == 0, got
    assert got[0][1] == got[1][1] and got[0][1] != 0
    assert not os.path.exists("/dev/shm/" + name)   # the leader unlinks it
