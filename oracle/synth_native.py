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


CPU_LIB = os.path.join(HERE, "liboracle_cpu.so")
_cpu = None


def build_cpu():
    src = os.path.join(HERE, "llama_cpu.c")
    if not os.path.exists(CPU_LIB) or os.path.getmtime(CPU_LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", src, "-o", CPU_LIB])
    return CPU_LIB


def effective_cpus():
    """Cores this process may actually use: affinity mask and cgroup CPU quota (a container can
    report 128 CPUs while being throttled to a fraction; spinning OpenMP threads then thrash)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, n)


def set_threads(n):
    global _cpu
    if _cpu is None:
        linear_bf16(np.zeros((1, 8), dtype=np.uint16), np.zeros((1, 8), dtype=np.float32))
    _cpu.oracle_set_threads(int(n))


def linear_bf16(w_bits, x):
    """y = x @ W^T with W as bf16 bit patterns (numpy uint16 [n, k]) and x fp32 [T, k]."""
    global _cpu
    if _cpu is None:
        _cpu = C.CDLL(build_cpu())
        _cpu.oracle_linear_bf16.restype = None
        _cpu.oracle_linear_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]
        _cpu.oracle_set_threads.restype = None
        _cpu.oracle_set_threads.argtypes = [C.c_int]
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, k = w_bits.shape
    y = np.empty((x.shape[0], n), dtype=np.float32)
    _cpu.oracle_linear_bf16(w_bits.ctypes.data, x.ctypes.data, y.ctypes.data, n, k, x.shape[0])
    return y


def _cpu_lib():
    global _cpu
    if _cpu is None:
        linear_bf16(np.zeros((1, 32), dtype=np.uint16), np.zeros((1, 32), dtype=np.float32))   # loads + prototypes the bf16 entry
    if not getattr(_cpu, "_q4_ready", False):
        _cpu.oracle_quantize_q4_0.restype = None
        _cpu.oracle_quantize_q4_0.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        _cpu.oracle_linear_q4_0.restype = None
        _cpu.oracle_linear_q4_0.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]
        _cpu.oracle_quantize_q8_0.restype = None
        _cpu.oracle_quantize_q8_0.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        _cpu.oracle_linear_q4_0_q8_0.restype = None
        _cpu.oracle_linear_q4_0_q8_0.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]
        _cpu._q4_ready = True
    return _cpu


def quantize_q4_0(w_bits):
    """bf16 bit pattern [n, k] -> ggml Q4_0 blocks, uint8 [n, k // 32 * 18] (oracle/llama_cpu.c)."""
    n, k = w_bits.shape
    assert k % 32 == 0
    out = np.empty((n, k // 32 * 18), dtype=np.uint8)
    _cpu_lib().oracle_quantize_q4_0(np.ascontiguousarray(w_bits).ctypes.data, out.ctypes.data, n, k)
    return out


def linear_q4_0(wq, k, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = wq.shape[0]
    y = np.empty((x.shape[0], n), dtype=np.float32)
    _cpu_lib().oracle_linear_q4_0(wq.ctypes.data, x.ctypes.data, y.ctypes.data, n, k, x.shape[0])
    return y


def quantize_q8_0(x):
    """fp32 [T, k] -> ggml Q8_0 blocks, uint8 [T, k // 32 * 34] (llama.cpp quantises the ACTIVATIONS this way)"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    t, k = x.shape
    assert k % 32 == 0
    out = np.empty((t, k // 32 * 34), dtype=np.uint8)
    _cpu_lib().oracle_quantize_q8_0(x.ctypes.data, out.ctypes.data, t, k)
    return out


def linear_q4_0_q8_0(wq, k, x):
    """y = W(Q4_0) . x with x quantised to Q8_0 first and integer block dot products: llama.cpp's CPU scheme"""
    xq = quantize_q8_0(x)
    n = wq.shape[0]
    y = np.empty((xq.shape[0], n), dtype=np.float32)
    _cpu_lib().oracle_linear_q4_0_q8_0(wq.ctypes.data, xq.ctypes.data, y.ctypes.data, n, k, xq.shape[0])
    return y


class Q4Weight:
    """ggml Q4_0 blocks of a [n, k] weight; llama_ref._mm dispatches to the C kernel for it."""

    def __init__(self, blocks, k, int_dot=True):
        self.blocks, self.k = blocks, k
        self.int_dot = int_dot          # True: Q8_0 activations + integer dot (llama.cpp's CPU scheme); False: fp32 activations
        self.dtype = "q4_0"


def q4_from_bits(sd_bits, int_dot=True):
    """A state dict of Bf16Weight (synth_state_dict_bits) -> the same model with every matmul weight (lm_head included) as
    ggml Q4_0 blocks; embedding rows and norm gains unchanged.  4.5 bits per weight: 4.2 GB for Llama-3-8B, ~3 s on 8 cores."""
    sd = dict(sd_bits)
    for name in list(sd):
        w = sd[name]
        if getattr(w, "dtype", None) == "bf16_bits" and name != "model.embed_tokens.weight":
            sd[name] = Q4Weight(quantize_q4_0(w.bits), w.bits.shape[1], int_dot)
    return sd


def synth_state_dict_q4(cfg, seed=0, int_dot=True):
    return q4_from_bits(synth_state_dict_bits(cfg, seed), int_dot)


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


class Bf16Weight:
    """bf16 weight kept as its bit pattern; llama_ref._mm dispatches to the C kernel for it."""

    def __init__(self, bits):
        self.bits = bits
        self.dtype = "bf16_bits"


def synth_state_dict_bits(cfg, seed=0):
    """HF-named weights for the CPU baseline: matmul weights as Bf16Weight, the rest fp32."""
    import torch
    from .synth import ID_EMBED, ID_LM_HEAD, KIND, bf16_bits_to_f32
    H, nh, nkv, hd, F, V = (cfg["hidden"], cfg["n_heads"], cfg["n_kv_heads"], cfg["head_dim"],
                            cfg["ffn"], cfg["vocab"])
    sd = {"model.embed_tokens.weight": Bf16Weight(synth_bits(seed, ID_EMBED, V, H)),
          "lm_head.weight": Bf16Weight(synth_bits(seed, ID_LM_HEAD, V, H)),
          "model.norm.weight": torch.ones(H)}
    shapes = {"self_attn.q_proj.weight": (nh * hd, H), "self_attn.k_proj.weight": (nkv * hd, H),
              "self_attn.v_proj.weight": (nkv * hd, H), "self_attn.o_proj.weight": (H, nh * hd),
              "mlp.gate_proj.weight": (F, H), "mlp.up_proj.weight": (F, H),
              "mlp.down_proj.weight": (H, F)}
    for l in range(cfg["n_layers"]):
        for k, (r, c) in shapes.items():
            sd["model.layers.%d.%s" % (l, k)] = Bf16Weight(synth_bits(seed, l * 16 + KIND[k], r, c))
        sd["model.layers.%d.input_layernorm.weight" % l] = torch.ones(H)
        sd["model.layers.%d.post_attention_layernorm.weight" % l] = torch.ones(H)
    return sd


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
