"""Pins the model-side oracle (oracle/llama_ref.py) against golden vectors produced by
HuggingFace transformers' LlamaForCausalLM (tests/golden/make_llama_golden.py)."""
import os

import numpy as np
import pytest

from llmlb_b200.ffi import LLAMA_TINY
from oracle.llama_ref import LlamaRef
from oracle.synth import bf16_bits_to_f32, f32_to_bf16_bits, synth_bits, synth_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden", "llama_tiny_golden.npz")


@pytest.fixture(scope="module")
def ref():
    return LlamaRef(LLAMA_TINY, synth_state_dict(LLAMA_TINY, seed=0))


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_logits_match_transformers(ref, case):
    g = np.load(GOLD)
    ref.reset()
    lg = ref.forward(g["prompt_" + case]).numpy()
    want = g["logits_" + case]
    assert np.abs(lg[-want.shape[0]:] - want).max() < 2e-5  # fp32 vs fp32, different op order


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_greedy_tokens_match_transformers(ref, case):
    g = np.load(GOLD)
    toks, _ = ref.greedy(g["prompt_" + case], len(g["greedy_" + case]))
    assert toks == g["greedy_" + case].tolist()


def test_incremental_equals_full(ref):
    rng = np.random.RandomState(5)
    p = rng.randint(0, LLAMA_TINY["vocab"], size=70)
    ref.reset()
    full = ref.forward(p).numpy()
    ref.reset()
    a = ref.forward(p[:33]).numpy()
    b = ref.forward(p[33:]).numpy()
    assert np.abs(np.concatenate([a, b]) - full).max() < 1e-4


def test_bf16_emulation_is_close_to_fp32():
    sd = synth_state_dict(LLAMA_TINY, seed=0)
    p = np.random.RandomState(6).randint(0, LLAMA_TINY["vocab"], size=48)
    a = LlamaRef(LLAMA_TINY, sd).forward(p).numpy()
    b = LlamaRef(LLAMA_TINY, sd, emulate_bf16=True).forward(p).numpy()
    assert np.abs(a - b).max() < 0.05 * a.std() + 0.02


def test_synth_generator_properties():
    bits = synth_bits(0, 7, 64, 512)
    vals = bf16_bits_to_f32(bits)
    assert abs(float(vals.std()) - 0.02) < 0.002 and abs(float(vals.mean())) < 0.002
    # a slice of a larger tensor equals the same region generated in one go (TP sharding relies on it)
    whole = synth_bits(3, 5, 32, 96)
    part = synth_bits(3, 5, 8, 32, row0=16, col0=64, ld=96)
    assert np.array_equal(whole[16:24, 64:96], part)
    # different tensors / seeds decorrelate
    assert not np.array_equal(synth_bits(0, 1, 4, 64), synth_bits(0, 2, 4, 64))
    assert not np.array_equal(synth_bits(0, 1, 4, 64), synth_bits(1, 1, 4, 64))


def test_bf16_rounding_is_rne():
    import torch
    x = np.random.RandomState(0).randn(4096).astype(np.float32)
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(f32_to_bf16_bits(x), want)


# ---- more geometries (tests/golden/make_llama_variants_golden.py): MHA / MQA / GQA, head_dim 64 / 128 ----
@pytest.mark.parametrize("name", ["mha64", "mqa128", "gqa4"])
def test_variant_geometries_match_transformers(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(os.path.dirname(__file__), "golden", "make_llama_variants_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    M = mk.VARIANTS[name]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "llama_variants_golden.npz"))
    r = LlamaRef(M, synth_state_dict(M, seed=7))
    lg = r.forward(g["prompt_" + name]).numpy()
    assert np.abs(lg[-6:] - g["logits_" + name]).max() < 3e-5
    toks, _ = r.greedy(g["prompt_" + name], len(g["greedy_" + name]))
    assert toks == g["greedy_" + name].tolist()
