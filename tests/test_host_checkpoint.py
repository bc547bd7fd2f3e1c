"""Native checkpoint readers (llmlb_b200/host/checkpoint.cpp) against the Python loaders
(llmlb_b200/gguf.py — pinned to llama.cpp's `gguf` package — and weights.py): every tensor of a
quantised .gguf and of a .safetensors file comes back with the same Hugging Face name, shape and
bf16 bits; random blocks of every quant type dequantise identically; geometry and the embedded
tokenizer are recovered.  CPU only."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from llmlb_b200 import build, gguf as G, weights

ref = pytest.importorskip("gguf")
from gguf import quants as RQ  # noqa: E402


@pytest.fixture(scope="module")
def K():
    lib = C.CDLL(build.build_host())
    lib.llmlb_ckpt_open.restype = C.c_void_p
    lib.llmlb_ckpt_open.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32]
    lib.llmlb_ckpt_close.argtypes = [C.c_void_p]
    lib.llmlb_ckpt_count.argtypes = [C.c_void_p]
    lib.llmlb_ckpt_count.restype = C.c_uint32
    lib.llmlb_ckpt_is_gguf.argtypes = [C.c_void_p]
    lib.llmlb_ckpt_geometry.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    lib.llmlb_ckpt_tensor_info.restype = C.c_int64
    lib.llmlb_ckpt_tensor_info.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.llmlb_ckpt_tensor_bf16.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint16), C.c_uint64]
    lib.llmlb_ckpt_tokenizer_json.restype = C.c_int64
    lib.llmlb_ckpt_tokenizer_json.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
    return lib


def read_all(K, path):
    err = C.create_string_buffer(256)
    h = K.llmlb_ckpt_open(str(path).encode(), err, 256)
    assert h, err.value
    out = {}
    name = C.create_string_buffer(256)
    rows, cols = C.c_uint64(), C.c_uint64()
    for i in range(K.llmlb_ckpt_count(h)):
        n = K.llmlb_ckpt_tensor_info(h, i, name, 256, C.byref(rows), C.byref(cols))
        buf = np.empty(n, dtype=np.uint16)
        assert K.llmlb_ckpt_tensor_bf16(h, i, buf.ctypes.data_as(C.POINTER(C.c_uint16)), n) == 0
        out[name.value.decode()] = buf.reshape(rows.value, cols.value)
    u7, f2 = (C.c_uint32 * 7)(), (C.c_float * 2)()
    known = K.llmlb_ckpt_geometry(h, u7, f2)
    geo = dict(zip(["hidden", "n_layers", "n_heads", "n_kv_heads", "head_dim", "ffn", "vocab"], list(u7)))
    geo.update(rope_theta=f2[0], rms_eps=f2[1], known=bool(known), gguf=bool(K.llmlb_ckpt_is_gguf(h)))
    n = K.llmlb_ckpt_tokenizer_json(h, None, 0)
    tj = C.create_string_buffer(max(1, n))
    K.llmlb_ckpt_tokenizer_json(h, tj, n)
    K.llmlb_ckpt_close(h)
    return out, geo, tj.raw[:n].decode("utf-8")


def _write_gguf(path, M, sd, quants):
    from gguf_util import _to_gguf_name
    Q = ref.GGMLQuantizationType
    w = ref.GGUFWriter(str(path), "llama")
    w.add_uint32("llama.block_count", M["n_layers"]); w.add_uint32("llama.embedding_length", M["hidden"])
    w.add_uint32("llama.feed_forward_length", M["ffn"]); w.add_uint32("llama.attention.head_count", M["n_heads"])
    w.add_uint32("llama.attention.head_count_kv", M["n_kv_heads"]); w.add_float32("llama.rope.freq_base", M["rope_theta"])
    w.add_float32("llama.attention.layer_norm_rms_epsilon", M["rms_eps"])
    k = 0
    for name, t in sd.items():
        t = np.asarray(t, dtype=np.float32)
        if name.endswith("q_proj.weight") or name.endswith("k_proj.weight"):
            nh = M["n_heads"] if "q_proj" in name else M["n_kv_heads"]
            t = np.ascontiguousarray(t.reshape(nh, 2, t.shape[0] // nh // 2, t.shape[1]).swapaxes(1, 2).reshape(t.shape))
        if t.ndim == 1 or "norm" in name:
            w.add_tensor(_to_gguf_name(name), t)
        else:
            qt = getattr(Q, quants[k % len(quants)])
            k += 1
            if qt in (Q.F16,):
                w.add_tensor(_to_gguf_name(name), t.astype(np.float16))
            else:
                q = RQ.quantize(t, qt)
                w.add_tensor(_to_gguf_name(name), q, raw_shape=q.shape, raw_dtype=qt)
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()


def test_gguf_tensors_match_the_python_loader(K, tmp_path):
    from llmlb_b200.ffi import LLAMA_TINY
    from oracle.synth import synth_state_dict
    sd = synth_state_dict(LLAMA_TINY, seed=4)
    p = tmp_path / "mixed.gguf"
    _write_gguf(p, LLAMA_TINY, sd, ["Q8_0", "Q4_0", "F16", "BF16"])
    want = dict(G.iter_hf_tensors(str(p)))
    got, geo, tj = read_all(K, p)
    assert set(got) == set(want)
    for name in want:
        assert got[name].shape == want[name].shape and np.array_equal(got[name], want[name]), name
    assert geo["known"] and geo["gguf"] and tj == ""
    meta, infos, mm = G.read_gguf(p)
    mm.close()
    pg = G.geometry(meta, infos)
    assert {k: geo[k] for k in ("hidden", "n_layers", "n_heads", "n_kv_heads", "head_dim", "ffn", "vocab")} == {k: pg[k] for k in ("hidden", "n_layers", "n_heads", "n_kv_heads", "head_dim", "ffn", "vocab")}
    assert abs(geo["rope_theta"] - pg["rope_theta"]) < 1 and abs(geo["rms_eps"] - pg["rms_eps"]) < 1e-9


@pytest.mark.parametrize("tname", ["Q8_0", "Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q4_K", "Q5_K", "Q6_K"])
def test_every_block_format_matches_llama_cpp(K, tmp_path, tname):
    """Random blocks written as a raw tensor of that type: the C++ dequantiser + bf16 rounding equals
    llama.cpp's dequantisation rounded the same way (K-quants cannot be produced by the Python
    package's quantiser, so the bytes are random with finite scales)."""
    from gguf_util import _random_blocks
    tid, per, blk = _random_blocks(tname, 64, seed=11)
    rows, cols = 8, blk.shape[0] * per // 8
    p = tmp_path / "blk.gguf"
    w = ref.GGUFWriter(str(p), "llama")
    w.add_uint32("llama.block_count", 1)
    qt = getattr(ref.GGMLQuantizationType, tname)
    w.add_tensor("blk.0.ffn_down.weight", blk.reshape(rows, -1), raw_shape=(rows, blk.size // rows), raw_dtype=qt)
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    got, _, _ = read_all(K, p)
    want = RQ.dequantize(blk, qt).astype(np.float32).reshape(rows, cols)
    assert np.array_equal(got["model.layers.0.mlp.down_proj.weight"], G.to_bf16_bits(want).reshape(rows, cols))


def test_safetensors_match_the_python_loader(K, tmp_path):
    from llmlb_b200.ffi import LLAMA_TINY
    from oracle.synth import f32_to_bf16_bits, synth_state_dict
    sd = synth_state_dict(LLAMA_TINY, seed=2)
    bits = {k: f32_to_bf16_bits(np.asarray(v, dtype=np.float32)) for k, v in sd.items()}
    p = tmp_path / "m.safetensors"
    weights.write_safetensors(p, bits)
    got, geo, _ = read_all(K, p)
    assert set(got) == set(bits)
    for k in bits:
        assert np.array_equal(got[k].reshape(-1), bits[k].reshape(-1)), k
    assert geo["known"] and not geo["gguf"]
    assert {k: geo[k] for k in ("hidden", "n_layers", "n_heads", "n_kv_heads", "ffn", "vocab")} == {k: LLAMA_TINY[k] for k in ("hidden", "n_layers", "n_heads", "n_kv_heads", "ffn", "vocab")}
    # fp32 / fp16 payloads are rounded to bf16 like the Python loader does
    f32 = np.random.RandomState(0).randn(4, 16).astype(np.float32)
    hdr = {"a.weight": {"dtype": "F32", "shape": [4, 16], "data_offsets": [0, 256]},
           "b.weight": {"dtype": "F16", "shape": [64], "data_offsets": [256, 384]}, "__metadata__": {"format": "pt"}}
    hj = json.dumps(hdr).encode()
    q = tmp_path / "f.safetensors"
    with open(q, "wb") as f:
        f.write(len(hj).to_bytes(8, "little")); f.write(hj); f.write(f32.tobytes()); f.write(f32.astype(np.float16).tobytes())
    got, _, _ = read_all(K, q)
    assert np.array_equal(got["a.weight"], weights._to_bf16_bits(f32.view(np.uint8).reshape(-1), "F32").reshape(4, 16))
    assert np.array_equal(got["b.weight"].reshape(-1), weights._to_bf16_bits(f32.astype(np.float16).view(np.uint8).reshape(-1), "F16"))


def test_embedded_tokenizer_and_errors(K, tmp_path):
    gold = os.path.join(os.path.dirname(__file__), "golden")
    tj = json.load(open(os.path.join(gold, "tokenizer_llama3_style.json"), encoding="utf-8"))
    n = max(max(tj["model"]["vocab"].values()), max(a["id"] for a in tj["added_tokens"])) + 1
    tokens, types = [""] * n, [1] * n
    for tok, i in tj["model"]["vocab"].items():
        tokens[i] = tok
    for a in tj["added_tokens"]:
        tokens[a["id"]], types[a["id"]] = a["content"], 3
    p = tmp_path / "tok.gguf"
    w = ref.GGUFWriter(str(p), "llama")
    w.add_tokenizer_model("gpt2"); w.add_tokenizer_pre("llama-bpe"); w.add_token_list(tokens); w.add_token_types(types)
    w.add_token_merges([m if isinstance(m, str) else " ".join(m) for m in tj["model"]["merges"]])
    w.add_bos_token_id(tokens.index("<|begin_of_text|>"))
    w.add_tensor("token_embd.weight", np.zeros((4, 32), dtype=np.float32))
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    got, geo, text = read_all(K, p)
    assert "lm_head.weight" in got and "model.embed_tokens.weight" in got          # tied head
    rebuilt = json.loads(text)
    meta, _, mm = G.read_gguf(p)
    mm.close()
    py = G.tokenizer_json_from_gguf(meta)
    assert rebuilt["model"]["vocab"] == py["model"]["vocab"] and rebuilt["model"]["merges"] == py["model"]["merges"]
    assert [(a["id"], a["content"], a["special"]) for a in rebuilt["added_tokens"]] == [(a["id"], a["content"], a["special"]) for a in py["added_tokens"]]
    # and the native tokenizer built from it reproduces the golden ids
    lib = K
    lib.llmlb_tok_create.restype = C.c_void_p
    lib.llmlb_tok_create.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint32]
    lib.llmlb_tok_encode.restype = C.c_int64
    lib.llmlb_tok_encode.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_uint64]
    raw = text.encode("utf-8")
    tok = lib.llmlb_tok_create(raw, len(raw), None, 0)
    assert tok
    out = (C.c_int32 * 8192)()
    for v in json.load(open(os.path.join(gold, "tokenizer_vectors.json"), encoding="utf-8"))["vectors"][:150]:
        b = v["text"].encode("utf-8")
        k = lib.llmlb_tok_encode(tok, b, len(b), 1, 1, out, 8192)
        assert list(out[:k]) == v["ids_bos"]
    # failure modes
    err = C.create_string_buffer(256)
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"GGUF" + (9).to_bytes(4, "little") + b"\0" * 32)
    assert not K.llmlb_ckpt_open(str(bad).encode(), err, 256) and b"version" in err.value
    bad.write_bytes((1 << 40).to_bytes(8, "little") + b"{}" + b"\0" * 16)
    assert not K.llmlb_ckpt_open(str(bad).encode(), err, 256)
    assert not K.llmlb_ckpt_open(str(tmp_path / "missing").encode(), err, 256)


def _gguf_bytes(kv, tensors, data=b"\0" * 4096):
    """A GGUF v3 image assembled by hand (the package's writer refuses to produce broken files).
    kv: [(key, type_id, packed_value)], tensors: [(name, dims innermost-first, ggml_type, offset)]."""
    s = lambda x: len(x).to_bytes(8, "little") + x
    out = b"GGUF" + (3).to_bytes(4, "little") + len(tensors).to_bytes(8, "little") + len(kv).to_bytes(8, "little")
    for k, t, v in kv:
        out += s(k.encode()) + t.to_bytes(4, "little") + v
    for name, dims, ty, off in tensors:
        out += s(name.encode()) + len(dims).to_bytes(4, "little") + b"".join(d.to_bytes(8, "little") for d in dims) + ty.to_bytes(4, "little") + off.to_bytes(8, "little")
    out += b"\0" * (-len(out) % 32)
    return out + data


def test_files_that_lie_about_their_sizes_are_refused(K, tmp_path):
    """Found by tools/fuzz (ASan/UBSan over mutated files): a zero `general.alignment` divided by zero, zero-sized dims
    handed memcpy a null pointer, 2^63-element shapes wrapped the size arithmetic, offsets near 2^64 wrapped the
    bounds check.  Every one of them is now an error from open(), with the tensor named."""
    import struct
    err = C.create_string_buffer(256)
    p = tmp_path / "m.gguf"
    u32, F32 = 4, 0
    ok_t = [("token_embd.weight", [32, 4], F32, 0)]

    def refused(kv, tensors, what):
        p.write_bytes(_gguf_bytes(kv, tensors))
        h = K.llmlb_ckpt_open(str(p).encode(), err, 256)
        assert not h and what in err.value, (what, err.value)

    p.write_bytes(_gguf_bytes([], ok_t))
    h = K.llmlb_ckpt_open(str(p).encode(), err, 256)
    assert h
    K.llmlb_ckpt_close(h)
    refused([("general.alignment", u32, (0).to_bytes(4, "little"))], ok_t, b"alignment")
    refused([("general.alignment", u32, (1 << 30).to_bytes(4, "little"))], ok_t, b"alignment")
    refused([], [("token_embd.weight", [32, 0], F32, 0)], b"shape of token_embd.weight")
    refused([], [("token_embd.weight", [1 << 62, 1 << 62], F32, 0)], b"shape of token_embd.weight")
    refused([], [("token_embd.weight", [1 << 62, 2], F32, 0)], b"past the end")               # n fits, n * 4 does not
    refused([], [("token_embd.weight", [32, 4], F32, (1 << 64) - 64)], b"past the end")        # base + off wraps
    refused([], [("token_embd.weight", [32, 4], F32, 4096 - 256)], b"past the end")            # 512 bytes do not fit in the last 256
    refused([("llama.block_count", u32, (1 << 31).to_bytes(4, "little"))], ok_t, b"block_count")
    # float metadata outside any integer range must not turn into garbage geometry
    p.write_bytes(_gguf_bytes([("llama.attention.head_count", 6, struct.pack("<f", float("inf"))), ("llama.embedding_length", 12, struct.pack("<d", -1e300))], ok_t))
    h = K.llmlb_ckpt_open(str(p).encode(), err, 256)
    assert h
    u7, f2 = (C.c_uint32 * 7)(), (C.c_float * 2)()
    assert K.llmlb_ckpt_geometry(h, u7, f2) == 0 and u7[0] == 0 and u7[2] == 0
    K.llmlb_ckpt_close(h)
    # safetensors: the same classes
    q = tmp_path / "m.safetensors"

    def st_refused(hdr, what):
        hj = json.dumps(hdr).encode()
        q.write_bytes(len(hj).to_bytes(8, "little") + hj + b"\0" * 1024)
        assert not K.llmlb_ckpt_open(str(q).encode(), err, 256) and what in err.value, (what, err.value)

    st_refused({"a.weight": {"dtype": "F32", "shape": [4, 16], "data_offsets": [0, (1 << 64) - 1]}}, b"offsets of a.weight")
    st_refused({"a.weight": {"dtype": "BF16", "shape": [1 << 59, 32], "data_offsets": [0, 0]}}, b"size of a.weight")
    st_refused({"model.layers.999999.mlp.up_proj.weight": {"dtype": "BF16", "shape": [2, 2], "data_offsets": [0, 8]}}, b"layer index")
