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


