"""Oracle restatement of the synthetic-weight generator (test infrastructure, not product).

Follows llmlb_b200/csrc/common.cuh `synth_value` bit for bit: a splitmix64-style integer hash of
(seed, tensor_id, global element index), the four 16-bit lanes summed (Irwin-Hall, n=4), centred,
scaled in fp32 and rounded to bf16 (RNE).  Distribution recipe from SURVEY.md §8d (N(0, 0.02^2)
matmul weights, unit norm gains); the reference repository holds no weights at all
(SURVEY.md §8c), so this generator is the shared source of truth for engine and oracle.
"""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
SYNTH_SUM_STD = np.float32(37837.22)


def _mix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def f32_to_bf16_bits(x):
    """fp32 -> bf16 bit pattern (uint16), round to nearest even (finite inputs)."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32)
    r = u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))
    return (r >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b):
    return (np.asarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def synth_bits(seed, tensor_id, rows, cols, std=0.02, row0=0, col0=0, ld=None):
    """bf16 bit patterns (uint16 [rows, cols]) of the slice at (row0, col0) of a tensor whose row
    pitch is `ld` elements."""
    ld = cols if ld is None else ld
    with np.errstate(over="ignore"):
        base = _mix64(np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + np.uint64(tensor_id))
        r = (np.arange(rows, dtype=np.uint64) + np.uint64(row0))[:, None]
        c = (np.arange(cols, dtype=np.uint64) + np.uint64(col0))[None, :]
        h = _mix64(base + (r * np.uint64(ld) + c))
    s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)) +
         ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + ((h >> np.uint64(48)) & np.uint64(0xFFFF)))
    s = s.astype(np.int64) - 131070
    scale = np.float32(std) / SYNTH_SUM_STD
    return f32_to_bf16_bits(s.astype(np.float32) * scale)


# tensor ids (llmlb_b200/csrc/engine.cu gen_weights): layer*16 + kind
KIND = {"self_attn.q_proj.weight": 0, "self_attn.k_proj.weight": 1, "self_attn.v_proj.weight": 2,
        "self_attn.o_proj.weight": 3, "mlp.gate_proj.weight": 4, "mlp.up_proj.weight": 5,
        "mlp.down_proj.weight": 6}
ID_EMBED, ID_LM_HEAD = 0xFFFF0000, 0xFFFF0002


def synth_state_dict(cfg, seed=0):
    """HF-named tensors (fp32 arrays holding bf16 values) for a Llama geometry dict."""
    H, nh, nkv, hd, F, V = (cfg["hidden"], cfg["n_heads"], cfg["n_kv_heads"], cfg["head_dim"],
                            cfg["ffn"], cfg["vocab"])
    sd = {}
    f = lambda tid, r, c: bf16_bits_to_f32(synth_bits(seed, tid, r, c))
    sd["model.embed_tokens.weight"] = f(ID_EMBED, V, H)
    sd["lm_head.weight"] = f(ID_LM_HEAD, V, H)
    sd["model.norm.weight"] = np.ones(H, dtype=np.float32)
    shapes = {"self_attn.q_proj.weight": (nh * hd, H), "self_attn.k_proj.weight": (nkv * hd, H),
              "self_attn.v_proj.weight": (nkv * hd, H), "self_attn.o_proj.weight": (H, nh * hd),
              "mlp.gate_proj.weight": (F, H), "mlp.up_proj.weight": (F, H),
              "mlp.down_proj.weight": (H, F)}
    for l in range(cfg["n_layers"]):
        for k, (r, c) in shapes.items():
            sd["model.layers.%d.%s" % (l, k)] = f(l * 16 + KIND[k], r, c)
        sd["model.layers.%d.input_layernorm.weight" % l] = np.ones(H, dtype=np.float32)
        sd["model.layers.%d.post_attention_layernorm.weight" % l] = np.ones(H, dtype=np.float32)
    return sd
