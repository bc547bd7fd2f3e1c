"""Helpers shared by the GGUF tests (Python loader and native reader)."""
import numpy as np

from llmlb_b200 import gguf as G


def _random_blocks(tname, n_blocks, seed):
    """Random bytes for `n_blocks` blocks of a type, with finite fp16 scale fields."""
    tid = [k for k, v in G.GGML_TYPES.items() if v[0] == tname][0]
    _, per, bsz = G.GGML_TYPES[tid]
    rs = np.random.RandomState(seed)
    blk = rs.randint(0, 256, size=(n_blocks, bsz)).astype(np.uint8)
    scale_fields = {"Q4_0": [0], "Q5_0": [0], "Q8_0": [0], "Q4_1": [0, 2], "Q5_1": [0, 2], "Q4_K": [0, 2], "Q5_K": [0, 2], "Q6_K": [208]}[tname]
    for off in scale_fields:
        vals = (rs.randn(n_blocks) * 0.05).astype(np.float16)
        blk[:, off:off + 2] = vals.view(np.uint8).reshape(n_blocks, 2)
    return tid, per, blk



def _to_gguf_name(hf):
    inv = {v: k for k, v in G._LAYER_MAP.items()}
    if hf == "model.embed_tokens.weight":
        return "token_embd.weight"
    if hf == "model.norm.weight":
        return "output_norm.weight"
    if hf == "lm_head.weight":
        return "output.weight"
    parts = hf.split(".")
    return "blk.%s.%s.weight" % (parts[2], inv[".".join(parts[3:-1])])




TINY_GGUF_MODEL = dict(hidden=512, n_layers=2, n_heads=8, n_kv_heads=2, head_dim=128, ffn=1024, vocab=3072, rope_theta=500000.0, rms_eps=1e-5)


def write_tiny_llama_gguf(path, seed=5, with_key_length=False):
    """A whole small Llama as ONE .gguf (Q8_0 matmul weights, F32 norms, llama-bpe tokenizer of the golden
    tokenizer.json), written with llama.cpp's own gguf.GGUFWriter.  head_dim (128) != hidden / heads (64)
    and `attention.key_length` is left out unless asked: readers must take the head width from the
    q projection's shape."""
    import json
    import os
    import gguf
    from gguf import quants as RQ
    from oracle.synth import synth_state_dict
    here = os.path.dirname(os.path.abspath(__file__))
    M = TINY_GGUF_MODEL
    sd = synth_state_dict(M, seed=seed)
    tj = json.load(open(os.path.join(here, "golden", "tokenizer_llama3_style.json"), encoding="utf-8"))
    tokens, types = ["<unused_%d>" % i for i in range(M["vocab"])], [5] * M["vocab"]
    for tok, i in tj["model"]["vocab"].items():
        tokens[i], types[i] = tok, 1
    for a in tj["added_tokens"]:
        tokens[a["id"]], types[a["id"]] = a["content"], 3
    w = gguf.GGUFWriter(str(path), "llama")
    w.add_uint32("llama.block_count", M["n_layers"]); w.add_uint32("llama.embedding_length", M["hidden"])
    w.add_uint32("llama.feed_forward_length", M["ffn"]); w.add_uint32("llama.attention.head_count", M["n_heads"])
    w.add_uint32("llama.attention.head_count_kv", M["n_kv_heads"]); w.add_float32("llama.rope.freq_base", M["rope_theta"])
    w.add_float32("llama.attention.layer_norm_rms_epsilon", M["rms_eps"])
    if with_key_length:
        w.add_uint32("llama.attention.key_length", M["head_dim"])
    w.add_tokenizer_model("gpt2"); w.add_tokenizer_pre("llama-bpe"); w.add_token_list(tokens); w.add_token_types(types)
    w.add_token_merges([m if isinstance(m, str) else " ".join(m) for m in tj["model"]["merges"]])
    w.add_bos_token_id(tokens.index("<|begin_of_text|>"))
    for name, t in sd.items():
        t = np.asarray(t, dtype=np.float32)
        if name.endswith("q_proj.weight") or name.endswith("k_proj.weight"):
            nh = M["n_heads"] if "q_proj" in name else M["n_kv_heads"]
            t = np.ascontiguousarray(t.reshape(nh, 2, t.shape[0] // nh // 2, t.shape[1]).swapaxes(1, 2).reshape(t.shape))
        if t.ndim == 1 or "norm" in name:
            w.add_tensor(_to_gguf_name(name), t)
        else:
            q = RQ.quantize(t, gguf.GGMLQuantizationType.Q8_0)
            w.add_tensor(_to_gguf_name(name), q, raw_shape=q.shape, raw_dtype=gguf.GGMLQuantizationType.Q8_0)
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    return M
