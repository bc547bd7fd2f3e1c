"""ctypes wrapper of oracle/synth_c.c (test infrastructure): full-size synthetic weights on the
host, as torch bf16 tensors, for the CPU baseline."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_synth.so")


def build():
    src = os.path.join(HERE, "synth_c.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", src, "-o", LIB])
    return LIB


_lib = None


def _get():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_synth_bf16.restype = None
        _lib.oracle_synth_bf16.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                           C.c_uint64, C.c_uint64, C.c_uint32, C.c_float]
    return _lib


def synth_bits(seed, tensor_id, rows, cols, std=0.02, row0=0, col0=0, ld=None):
    out = np.empty((rows, cols), dtype=np.uint16)
    _get().oracle_synth_bf16(out.ctypes.data, rows, cols, row0, col0, cols if ld is None else ld,
                             seed, tensor_id, std)
    return out


def synth_state_dict_bf16(cfg, seed=0):
    """HF-named torch.bfloat16 tensors for a Llama geometry (8B: ~16 GB, a few seconds)."""
    import torch
    from .synth import ID_EMBED, ID_LM_HEAD, KIND
    H, nh, nkv, hd, F, V = (cfg["hidden"], cfg["n_heads"], cfg["n_kv_heads"], cfg["head_dim"],
                            cfg["ffn"], cfg["vocab"])
    t = lambda tid, r, c: torch.from_numpy(synth_bits(seed, tid, r, c).view(np.int16)).view(torch.bfloat16)
    sd = {"model.embed_tokens.weight": t(ID_EMBED, V, H), "lm_head.weight": t(ID_LM_HEAD, V, H),
          "model.norm.weight": torch.ones(H, dtype=torch.bfloat16)}
    shapes = {"self_attn.q_proj.weight": (nh * hd, H), "self_attn.k_proj.weight": (nkv * hd, H),
              "self_attn.v_proj.weight": (nkv * hd, H), "self_attn.o_proj.weight": (H, nh * hd),
              "mlp.gate_proj.weight": (F, H), "mlp.up_proj.weight": (F, H),
              "mlp.down_proj.weight": (H, F)}
    for l in range(cfg["n_layers"]):
        for k, (r, c) in shapes.items():
            sd["model.layers.%d.%s" % (l, k)] = t(l * 16 + KIND[k], r, c)
        sd["model.layers.%d.input_layernorm.weight" % l] = torch.ones(H, dtype=torch.bfloat16)
        sd["model.layers.%d.post_attention_layernorm.weight" % l] = torch.ones(H, dtype=torch.bfloat16)
    return sd
