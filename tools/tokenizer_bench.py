"""Host-side tokenizer throughput: the native tokenizer (llmlb_b200/host/tokenizer.cpp) beside the
Hugging Face `tokenizers` library (Rust) on the same tokenizer.json and the same documents, one
thread each.  python tools/tokenizer_bench.py [n_docs]"""
import ctypes as C
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
from llmlb_b200 import build  # noqa: E402
import make_tokenizer_golden as mk  # noqa: E402


def main():
    n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    path = os.path.join(os.path.dirname(mk.__file__), "tokenizer_llama3_style.json")
    rng = random.Random(7)
    docs = [mk.synth_text(rng, rng.randint(200, 1200)) for _ in range(n_docs)]
    total = sum(len(d.encode("utf-8")) for d in docs)
    lib = C.CDLL(build.build_host())
    lib.llmlb_tok_create.restype = C.c_void_p
    lib.llmlb_tok_create.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint32]
    lib.llmlb_tok_encode.restype = C.c_int64
    lib.llmlb_tok_encode.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_uint64]
    lib.llmlb_tok_decode.restype = C.c_int64
    lib.llmlb_tok_decode.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_uint64, C.c_int, C.c_char_p, C.c_uint64]
    data = open(path, "rb").read()
    tok = lib.llmlb_tok_create(data, len(data), None, 0)
    raw = [d.encode("utf-8") for d in docs]
    cap = max(len(r) for r in raw) + 16
    out = (C.c_int32 * cap)()
    t0 = time.perf_counter()
    n_tok = 0
    ids_native = []
    for r in raw:
        n = lib.llmlb_tok_encode(tok, r, len(r), 0, 1, out, cap)
        n_tok += n
        ids_native.append(list(out[:n]))
    t_native = time.perf_counter() - t0
    from tokenizers import Tokenizer
    hf = Tokenizer.from_file(path)
    t0 = time.perf_counter()
    ids_hf = [hf.encode(d, add_special_tokens=False).ids for d in docs]
    t_hf = time.perf_counter() - t0
    assert ids_native == ids_hf
    buf = C.create_string_buffer(cap * 8)
    t0 = time.perf_counter()
    for ids in ids_native:
        arr = (C.c_int32 * len(ids))(*ids)
        lib.llmlb_tok_decode(tok, arr, len(ids), 0, buf, cap * 8)
    t_dec = time.perf_counter() - t0
    t0 = time.perf_counter()
    for ids in ids_hf:
        hf.decode(ids, skip_special_tokens=False)
    t_dec_hf = time.perf_counter() - t0
    print("%d documents, %.2f MB, %d tokens (identical ids)" % (n_docs, total / 1e6, n_tok))
    print("encode: native %.1f MB/s (%.2f M tok/s)   tokenizers %.1f MB/s (%.2f M tok/s)" % (
        total / 1e6 / t_native, n_tok / 1e6 / t_native, total / 1e6 / t_hf, n_tok / 1e6 / t_hf))
    print("decode: native %.2f M tok/s   tokenizers %.2f M tok/s" % (n_tok / 1e6 / t_dec, n_tok / 1e6 / t_dec_hf))


if __name__ == "__main__":
    main()
