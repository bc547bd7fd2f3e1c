"""The CPU baseline's Q4_0 leg (oracle/llama_cpu.c: ggml's block format and reference quantiser restated in C) pinned to
llama.cpp's own `gguf` Python package — the third-party format of the reference path's CPU endpoint in BASELINE.json
configs[0] ("llama.cpp ... Llama-3-8B q4"); llama.cpp itself is external to the reference tree."""
import numpy as np
import pytest
import torch

gguf = pytest.importorskip("gguf")
from gguf import quants  # noqa: E402

from oracle import synth_native as S  # noqa: E402
from oracle.llama_ref import LlamaRef  # noqa: E402
from oracle.synth import bf16_bits_to_f32  # noqa: E402

Q4 = gguf.GGMLQuantizationType.Q4_0


def _bf16(f):
    b = torch.from_numpy(np.ascontiguousarray(f, dtype=np.float32)).to(torch.bfloat16)
    return b.view(torch.int16).numpy().view(np.uint16), b.float().numpy()


def test_quantiser_is_bit_exact_against_gguf():
    S.build_cpu()
    rs = np.random.RandomState(2)
    cases = [bf16_bits_to_f32(S.synth_bits(0, 3, 96, 512)),                                      # the synthetic weights themselves
             rs.randn(256, 1024) * rs.choice([1e-7, 1e-6, 1e-3, 1.0, 300.0, 6e4], size=(256, 1)),  # subnormal fp16 scales ... large values
             np.zeros((4, 64)), -np.ones((4, 64)), np.tile(np.linspace(-8, 7, 32), (3, 2))]         # all-zero blocks, ties at the rounding points
    for f in cases:
        bits, fr = _bf16(f)
        ref = quants.quantize(fr, Q4)
        got = S.quantize_q4_0(bits)
        assert got.shape == (fr.shape[0], fr.shape[1] // 32 * 18) and np.array_equal(ref.reshape(got.shape), got)


def test_linear_equals_dequantised_matmul():
    rs = np.random.RandomState(3)
    for n, k, t in ((64, 256, 1), (130, 4096, 3), (33, 512, 8)):
        bits, _ = _bf16(rs.randn(n, k) * 0.02)
        q = S.quantize_q4_0(bits)
        w = quants.dequantize(q.reshape(-1), Q4).reshape(n, k) if q.ndim == 1 else quants.dequantize(q, Q4)
        x = rs.randn(t, k).astype(np.float32)
        y = S.linear_q4_0(q, k, x)
        ref = x.astype(np.float64) @ w.astype(np.float64).T
        assert np.abs(y - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


def test_q8_activation_quantiser_and_integer_dot():
    """llama.cpp multiplies a Q4_0 weight by activations quantised to Q8_0, on integers: the activation quantiser is bit-exact
    against gguf-py and the AVX2 dot equals a float64 evaluation of the very same integers and scales."""
    Q8 = gguf.GGMLQuantizationType.Q8_0
    rs = np.random.RandomState(4)
    x = (rs.randn(6, 1024) * rs.choice([1e-4, 1.0, 40.0], size=(6, 1))).astype(np.float32)
    x[2] = 0
    x[3, :32] = np.linspace(-1, 1, 32) * 127 / 127.0      # ties of roundf
    q8 = S.quantize_q8_0(x)
    assert np.array_equal(quants.quantize(x, Q8).reshape(q8.shape), q8)
    bits, _ = _bf16(rs.randn(200, 1024) * 0.02)
    q4 = S.quantize_q4_0(bits)
    ref = quants.dequantize(q8, Q8).astype(np.float64) @ quants.dequantize(q4, Q4).astype(np.float64).T
    y = S.linear_q4_0_q8_0(q4, 1024, x)
    assert np.abs(y - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def test_q4_model_forward_equals_the_dequantised_fp32_model():
    from llmlb_b200 import ffi
    cfg = ffi.LLAMA_TINY
    sd_q = S.synth_state_dict_q4(cfg, seed=0, int_dot=False)
    sd_f = {}
    for name, w in sd_q.items():
        if getattr(w, "dtype", None) == "q4_0":
            sd_f[name] = torch.from_numpy(quants.dequantize(w.blocks, Q4).astype(np.float32))
        elif getattr(w, "dtype", None) == "bf16_bits":
            sd_f[name] = torch.from_numpy(bf16_bits_to_f32(w.bits).copy())
        else:
            sd_f[name] = w
    prompt = np.random.RandomState(0).randint(0, cfg["vocab"], 24).tolist()
    a = LlamaRef(cfg, sd_q).forward(prompt).numpy()
    b = LlamaRef(cfg, sd_f).forward(prompt).numpy()
    assert np.abs(a - b).max() < 1e-3
    # 4.5 bits per weight: 18 bytes per 32 weights
    n_w = sum(w.blocks.shape[0] * w.k for w in sd_q.values() if getattr(w, "dtype", None) == "q4_0")
    n_b = sum(w.blocks.nbytes for w in sd_q.values() if getattr(w, "dtype", None) == "q4_0")
    assert abs(n_b * 8 / n_w - 4.5) < 1e-9


def test_integer_path_stays_close_to_the_fp32_activation_path():
    """Same Q4_0 weights, activations through Q8_0 (llama.cpp's scheme) vs fp32: logits move by the activation rounding only."""
    from llmlb_b200 import ffi
    cfg = ffi.LLAMA_TINY
    prompt = np.random.RandomState(1).randint(0, cfg["vocab"], 24).tolist()
    a = LlamaRef(cfg, S.synth_state_dict_q4(cfg, seed=0, int_dot=True)).forward(prompt).numpy()
    b = LlamaRef(cfg, S.synth_state_dict_q4(cfg, seed=0, int_dot=False)).forward(prompt).numpy()
    assert np.abs(a - b).max() < 0.05 * b.std() + 0.01 and (a.argmax(-1) == b.argmax(-1)).mean() > 0.8
