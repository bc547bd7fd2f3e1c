"""GGUF reader / dequantisers (llmlb_b200/gguf.py) pinned to llama.cpp's own
`gguf` Python package: block dequantisation bit for bit on random blocks, the file parser on
files written by gguf.GGUFWriter, the Q/K row permutation against the converter's formula, and
(GPU) an engine loaded from a quantised .gguf."""
import numpy as np
import pytest

from llmlb_b200 import gguf as G

ref = pytest.importorskip("gguf")
from gguf import quants as RQ  # noqa: E402
from gguf_util import _random_blocks, _to_gguf_name  # noqa: E402


@pytest.mark.parametrize("tname", ["Q8_0", "Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q4_K", "Q5_K", "Q6_K"])
def test_block_dequantisers_match_llama_cpp(tname):
    tid, per, blk = _random_blocks(tname, 257, seed=hash(tname) % 1000)
    want = RQ.dequantize(blk, getattr(ref.GGMLQuantizationType, tname)).reshape(-1)
    got = G.dequantize(blk.reshape(-1), tid, blk.shape[0] * per)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32))


def test_float_types_and_bf16_rounding():
    x = np.random.RandomState(1).randn(512).astype(np.float32)
    assert np.array_equal(G.dequantize(x.view(np.uint8), 0, 512), x)
    h = x.astype(np.float16)
    assert np.array_equal(G.dequantize(h.view(np.uint8), 1, 512), h.astype(np.float32))
    bf = RQ.quantize(x, ref.GGMLQuantizationType.BF16).reshape(-1)            # llama.cpp's fp32 -> bf16
    assert np.array_equal(G.to_bf16_bits(x), bf.view(np.uint16))
    assert np.array_equal(G.dequantize(bf, 30, 512).view(np.uint32) >> 16, bf.view(np.uint16).astype(np.uint32))
    with pytest.raises(G.GGUFError):
        G.dequantize(np.zeros(17, np.uint8), 2, 32)
    with pytest.raises(G.GGUFError):
        G.dequantize(np.zeros(18, np.uint8), 99, 32)


def _write(path, tensors, arch="llama", extra=None):
    w = ref.GGUFWriter(str(path), arch)
    w.add_uint32(arch + ".block_count", 2)
    w.add_uint32(arch + ".embedding_length", 512)
    w.add_uint32(arch + ".feed_forward_length", 1024)
    w.add_uint32(arch + ".attention.head_count", 4)
    w.add_uint32(arch + ".attention.head_count_kv", 2)
    w.add_float32(arch + ".rope.freq_base", 500000.0)
    w.add_float32(arch + ".attention.layer_norm_rms_epsilon", 1e-5)
    w.add_string("general.name", "tiny-llama ✓")
    w.add_array("tokenizer.ggml.tokens", ["a", "b", "ç"])
    w.add_array("tokenizer.ggml.scores", [0.5, -1.0, 2.0])
    for k, v in (extra or {}).items():
        w.add_uint32(k, v)
    for name, arr, qt in tensors:
        if qt is None:
            w.add_tensor(name, arr)
        else:
            q = RQ.quantize(arr, qt)
            w.add_tensor(name, q, raw_shape=q.shape, raw_dtype=qt)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()


def test_reader_matches_llama_cpp_reader(tmp_path):
    rs = np.random.RandomState(2)
    Q = ref.GGMLQuantizationType
    tensors = [("token_embd.weight", rs.randn(96, 64).astype(np.float32), None),
               ("blk.0.attn_q.weight", rs.randn(64, 64).astype(np.float32), Q.Q8_0),
               ("blk.0.ffn_down.weight", rs.randn(64, 128).astype(np.float32), Q.Q4_0),
               ("blk.0.attn_norm.weight", rs.rand(64).astype(np.float32), None),
               ("blk.1.ffn_up.weight", rs.randn(128, 64).astype(np.float16), None),
               ("rope_freqs.weight", rs.rand(16).astype(np.float32), None)]
    p = tmp_path / "t.gguf"
    _write(p, tensors)
    meta, infos, mm = G.read_gguf(p)
    try:
        r = ref.GGUFReader(str(p))
        assert meta["general.architecture"] == "llama" and meta["general.name"] == "tiny-llama ✓"
        assert meta["llama.attention.head_count_kv"] == 2 and abs(meta["llama.rope.freq_base"] - 500000.0) < 1e-3
        assert meta["tokenizer.ggml.tokens"] == ["a", "b", "ç"] and meta["tokenizer.ggml.scores"] == [0.5, -1.0, 2.0]
        assert [t["name"] for t in infos] == [t.name for t in r.tensors]
        for mine, theirs in zip(infos, r.tensors):
            assert mine["shape"] == [int(d) for d in reversed(theirs.shape)]
            assert mine["type"] == int(theirs.tensor_type) and mine["nbytes"] == int(theirs.n_bytes)
            raw = np.frombuffer(mm, dtype=np.uint8, count=mine["nbytes"], offset=mine["offset"])
            assert np.array_equal(raw, np.asarray(theirs.data).reshape(-1).view(np.uint8))
            n = int(np.prod(mine["shape"]))
            want = RQ.dequantize(np.asarray(theirs.data), theirs.tensor_type).reshape(-1).astype(np.float32)
            assert np.array_equal(G.dequantize(raw, mine["type"], n), want)
            del raw
        assert G.geometry(meta, infos) == {"hidden": 512, "n_layers": 2, "n_heads": 4, "n_kv_heads": 2, "head_dim": 128, "ffn": 1024,
                                           "vocab": 96, "rope_theta": 500000.0, "rms_eps": pytest.approx(1e-5)}
    finally:
        mm.close()
    with pytest.raises(G.GGUFError):
        bad = tmp_path / "bad.gguf"
        bad.write_bytes(b"GGML" + b"\0" * 64)
        G.read_gguf(bad)


def test_names_and_qk_permutation():
    assert G.hf_name("blk.17.attn_q.weight") == "model.layers.17.self_attn.q_proj.weight"
    assert G.hf_name("blk.0.ffn_gate.weight") == "model.layers.0.mlp.gate_proj.weight"
    assert G.hf_name("token_embd.weight") == "model.embed_tokens.weight" and G.hf_name("output.weight") == "lm_head.weight"
    assert G.hf_name("output_norm.weight") == "model.norm.weight"
    assert G.hf_name("rope_freqs.weight") is None and G.hf_name("blk.x.attn_q.weight") is None
    # convert_hf_to_gguf.py LlamaModel.permute: reshape(n_head, 2, rows/n_head/2, cols).swapaxes(1, 2).reshape
    w = np.arange(8 * 16 * 3, dtype=np.float32).reshape(8 * 16, 3)
    for n_head in (1, 2, 8):
        permuted = w.reshape(n_head, 2, w.shape[0] // n_head // 2, 3).swapaxes(1, 2).reshape(w.shape)
        assert np.array_equal(G.unpermute_qk(permuted, n_head), w)


def _tiny_state(seed):
    from llmlb_b200.ffi import LLAMA_TINY
    from oracle.synth import synth_state_dict
    return LLAMA_TINY, synth_state_dict(LLAMA_TINY, seed=seed)


def _write_tiny_gguf(path, seed, quant):
    M, sd = _tiny_state(seed)
    Q = ref.GGMLQuantizationType
    tensors = []
    for name, w in sd.items():
        w = np.asarray(w, dtype=np.float32)
        if name.endswith("q_proj.weight") or name.endswith("k_proj.weight"):     # what the converter does to Q/K
            nh = M["n_heads"] if "q_proj" in name else M["n_kv_heads"]
            w = w.reshape(nh, 2, w.shape[0] // nh // 2, w.shape[1]).swapaxes(1, 2).reshape(w.shape)
        qt = None if w.ndim == 1 or "norm" in name else quant
        tensors.append((_to_gguf_name(name), np.ascontiguousarray(w), qt))
    w = ref.GGUFWriter(str(path), "llama")
    w.add_uint32("llama.block_count", M["n_layers"])
    w.add_uint32("llama.embedding_length", M["hidden"])
    w.add_uint32("llama.feed_forward_length", M["ffn"])
    w.add_uint32("llama.attention.head_count", M["n_heads"])
    w.add_uint32("llama.attention.head_count_kv", M["n_kv_heads"])
    w.add_float32("llama.rope.freq_base", M["rope_theta"])
    w.add_float32("llama.attention.layer_norm_rms_epsilon", M["rms_eps"])
    for name, arr, qt in tensors:
        if qt is None:
            w.add_tensor(name, arr)
        else:
            q = RQ.quantize(arr, qt)
            w.add_tensor(name, q, raw_shape=q.shape, raw_dtype=qt)
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    return M, sd


@pytest.mark.parametrize("quant", ["Q8_0", "Q4_0"])
def test_iter_hf_tensors_restores_the_checkpoint(tmp_path, quant):
    """CPU: every tensor comes back under its Hugging Face name, Q/K rows in HF order, values equal to
    llama.cpp's own dequantisation of the same blocks rounded to bf16."""
    qt = getattr(ref.GGMLQuantizationType, quant)
    p = tmp_path / "tiny.gguf"
    M, sd = _write_tiny_gguf(p, 0, qt)
    got = dict(G.iter_hf_tensors(str(p)))
    assert set(got) == set(sd)
    for name, w in sd.items():
        w = np.asarray(w, dtype=np.float32)
        if w.ndim == 1 or "norm" in name:
            want = w
        else:
            src = w
            if name.endswith("q_proj.weight") or name.endswith("k_proj.weight"):
                nh = M["n_heads"] if "q_proj" in name else M["n_kv_heads"]
                src = w.reshape(nh, 2, w.shape[0] // nh // 2, w.shape[1]).swapaxes(1, 2).reshape(w.shape)
            want = RQ.dequantize(RQ.quantize(np.ascontiguousarray(src), qt), qt).astype(np.float32)
            if src is not w:
                want = G.unpermute_qk(want, nh)
        assert np.array_equal(got[name].reshape(-1), G.to_bf16_bits(want).reshape(-1)), name
        # and the quantisation error is what the format promises
        err = np.abs((got[name].astype(np.uint32) << 16).view(np.float32).reshape(w.shape if w.ndim == 2 else (1, -1)) - w.reshape(got[name].shape)).max()
        assert err < (0.02 if quant == "Q8_0" else 0.2) * max(1e-3, np.abs(w).max())


@pytest.mark.gpu
def test_engine_loaded_from_gguf(tmp_path, built_lib):
    from llmlb_b200 import ffi
    p = tmp_path / "tiny-q8.gguf"
    M, sd = _write_tiny_gguf(p, 0, ref.GGMLQuantizationType.Q8_0)
    want = dict(G.iter_hf_tensors(str(p)))
    prompt = list(range(7, 60))
    with ffi.Engine(M, max_seqs=4, max_ctx=256, seed=0) as exact:
        lg_exact = exact.debug_prefill_logits(prompt)
    with ffi.Engine(M, max_seqs=4, max_ctx=256, seed=99) as e:
        names = G.load_gguf(e, p)
        assert len(names) == 3 + 9 * M["n_layers"]
        for probe in ("model.layers.1.self_attn.q_proj.weight", "model.layers.0.mlp.down_proj.weight", "lm_head.weight"):
            assert np.array_equal(e.read_tensor(probe, 1 << 22).reshape(-1), want[probe].reshape(-1))
        lg = e.debug_prefill_logits(prompt)
    # Q8_0 weights: logits close to the unquantised model's, far from a different seed's
    assert np.abs(lg - lg_exact).max() < 0.15 * np.abs(lg_exact).max()
    assert np.corrcoef(lg.reshape(-1), lg_exact.reshape(-1))[0, 1] > 0.99


def test_tokenizer_embedded_in_gguf(tmp_path):
    """A .gguf alone is enough to serve text: its tokenizer.ggml.* arrays rebuild a tokenizer.json that
    the native tokenizer loads and that tokenises the golden vectors exactly like the original."""
    import ctypes as C
    import json
    import os
    from llmlb_b200 import build
    gold_dir = os.path.join(os.path.dirname(__file__), "golden")
    tj = json.load(open(os.path.join(gold_dir, "tokenizer_llama3_style.json"), encoding="utf-8"))
    vectors = json.load(open(os.path.join(gold_dir, "tokenizer_vectors.json"), encoding="utf-8"))["vectors"]
    n = max(max(tj["model"]["vocab"].values()), max(a["id"] for a in tj["added_tokens"])) + 1
    tokens, types = [""] * n, [1] * n
    for tok, i in tj["model"]["vocab"].items():
        tokens[i] = tok
    for a in tj["added_tokens"]:
        tokens[a["id"]], types[a["id"]] = a["content"], 3
    merges = [m if isinstance(m, str) else " ".join(m) for m in tj["model"]["merges"]]
    p = tmp_path / "tok.gguf"
    w = ref.GGUFWriter(str(p), "llama")
    w.add_tokenizer_model("gpt2")
    w.add_tokenizer_pre("llama-bpe")
    w.add_token_list(tokens)
    w.add_token_types(types)
    w.add_token_merges(merges)
    w.add_bos_token_id(tokens.index("<|begin_of_text|>"))
    w.add_tensor("token_embd.weight", np.zeros((4, 32), dtype=np.float32))
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    meta, _, mm = G.read_gguf(p)
    mm.close()
    rebuilt = json.dumps(G.tokenizer_json_from_gguf(meta), ensure_ascii=False).encode("utf-8")
    lib = C.CDLL(build.build_host())
    lib.llmlb_tok_create.restype = C.c_void_p
    lib.llmlb_tok_create.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint32]
    lib.llmlb_tok_encode.restype = C.c_int64
    lib.llmlb_tok_encode.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_uint64]
    lib.llmlb_tok_destroy.argtypes = [C.c_void_p]
    err = C.create_string_buffer(256)
    tok = lib.llmlb_tok_create(rebuilt, len(rebuilt), err, 256)
    assert tok, err.value
    out = (C.c_int32 * 8192)()
    for v in vectors:
        b = v["text"].encode("utf-8")
        k = lib.llmlb_tok_encode(tok, b, len(b), 0, 1, out, 8192)
        assert list(out[:k]) == v["ids"], repr(v["text"])
        k = lib.llmlb_tok_encode(tok, b, len(b), 1, 1, out, 8192)
        assert list(out[:k]) == v["ids_bos"], repr(v["text"])
    lib.llmlb_tok_destroy(tok)
    with pytest.raises(G.GGUFError):
        G.tokenizer_json_from_gguf({"tokenizer.ggml.model": "llama"})
