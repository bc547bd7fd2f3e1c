"""llmlb_engine_create with random configurations on the fake CUDA runtime: geometry, batch shape, page pool, tensor-parallel
rank / size, run-ahead, timeouts from {0, 1, odd, plausible, huge}.  Every call returns an engine (which then serves one
request and is destroyed) or a negative error code — never a crash, never an engine that cannot serve.  usage: <calls> <seed>"""
import ctypes as C
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from llmlb_b200 import ffi  # noqa: E402

if os.environ.get("LLMLB_HOST_LOGIC_LIB"):
    ffi.LIB_PATH = os.environ["LLMLB_HOST_LOGIC_LIB"]
L = ffi.lib()
rs = random.Random(int(sys.argv[2]))
ok = err = 0
for _ in range(int(sys.argv[1])):
    cfg = ffi.EngineConfig()
    cfg.abi_version = ffi.ABI_VERSION
    m = cfg.model
    m.hidden, m.n_layers, m.n_heads, m.n_kv_heads, m.head_dim, m.ffn, m.vocab, m.rope_theta, m.rms_eps = 512, 2, 8, 2, 128, 1024, 2048, 500000.0, 1e-5
    cfg.model_id, cfg.device, cfg.tp_size, cfg.tp_rank = b"m", 0, 1, 0
    cfg.max_seqs, cfg.max_ctx, cfg.kv_block_tokens, cfg.use_cuda_graphs = 8, 512, 64, 1
    odd = {
        "abi_version": [0, 99], "device": [1, -1, 99], "tp_size": [0, 2, 3, 4, 8, 16], "tp_rank": [1, 3, 9],
        "max_seqs": [0, 1, 2, 64, 128, 129, 1000, 100000], "max_ctx": [0, 1, 63, 64, 100, 1024, 8192, 1 << 20], "kv_block_tokens": [0, 16, 128],
        "kv_pages": [1, 2, 7, 8, 100, 1 << 20], "max_step_tokens": [1, 63, 64, 2048, 100000], "use_cuda_graphs": [0, 7], "gemm_impl": [1, 5],
        "lookahead": [1, 2, 7, 1000], "queue_max": [1, 100], "queue_timeout_ms": [1, 1000], "request_timeout_ms": [1, 1000], "attn_impl": [1, 9],
        "tp_proto": [1, 2, 3, 99], "synthetic_seed": [1 << 63],
    }
    odd_model = {"hidden": [0, 64, 128, 500, 1024, 4096], "n_layers": [0, 1, 3, 80], "n_heads": [0, 1, 4, 16, 32, 7], "n_kv_heads": [0, 1, 4, 8, 3],
                 "head_dim": [0, 64, 256], "ffn": [0, 256, 1000, 2048, 28672], "vocab": [0, 1, 1000, 3072, 128256],
                 "rope_theta": [0.0, 10000.0, float("nan")], "rms_eps": [0.0, -1.0]}
    for _ in range(rs.choice([0, 1, 1, 2, 3])):
        if rs.random() < 0.5:
            k = rs.choice(list(odd))
            setattr(cfg, k, rs.choice(odd[k]))
        else:
            k = rs.choice(list(odd_model))
            setattr(m, k, rs.choice(odd_model[k]))
    h = C.c_void_p()
    rc = L.llmlb_engine_create(C.byref(cfg), C.byref(h))
    assert rc <= 0, rc
    if rc != 0:
        err += 1
        assert not h.value
        continue
    ok += 1
    if cfg.tp_size == 1:                       # a created engine must be able to serve (tp > 1 needs its peers first)
        ids = (C.c_int32 * 3)(1 % m.vocab, 2 % m.vocab, 0)
        s = ffi.Sampling()
        s.max_tokens, s.ignore_eos = 2, 1
        rid = C.c_uint64()
        rc = L.llmlb_request_submit(h, ids, 3, C.byref(s), C.byref(rid))
        if rc == 0:
            ev = (ffi.TokenEvent * 8)()
            n = C.c_uint32()
            for _ in range(2000):
                L.llmlb_request_poll(h, rid.value, ev, 8, C.byref(n), 5)
                if n.value and ev[n.value - 1].finish_reason:
                    break
            else:
                raise AssertionError("engine created from %s never finished a request" % [(f[0], getattr(cfg, f[0])) for f in cfg._fields_ if f[0] != "model"])
    L.llmlb_engine_destroy(h)
print("creates", ok + err, "engines", ok, "refused", err)
