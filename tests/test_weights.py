"""safetensors reader (CPU) and loading real tensors through llmlb_engine_load_tensor (GPU)."""
import os

import numpy as np
import pytest

from llmlb_b200 import weights
from llmlb_b200.ffi import LLAMA_TINY
from oracle.synth import f32_to_bf16_bits, synth_bits, synth_state_dict


def _tiny_bits(seed):
    sd = synth_state_dict(LLAMA_TINY, seed=seed)
    return {k: f32_to_bf16_bits(v) for k, v in sd.items()}


def test_safetensors_roundtrip(tmp_path):
    t = {"a.weight": synth_bits(1, 2, 5, 16), "b.weight": synth_bits(1, 3, 1, 8).reshape(8)}
    p = tmp_path / "x.safetensors"
    weights.write_safetensors(p, t)
    header, base = weights.read_safetensors_header(p)
    assert header["a.weight"]["shape"] == [5, 16] and header["b.weight"]["dtype"] == "BF16"
    raw = np.fromfile(p, dtype=np.uint8)
    a, b = header["a.weight"]["data_offsets"]
    assert np.array_equal(raw[base + a: base + b].view(np.uint16).reshape(5, 16), t["a.weight"])


def test_fp32_and_fp16_are_rounded_to_bf16():
    x = np.random.RandomState(0).randn(64).astype(np.float32)
    assert np.array_equal(weights._to_bf16_bits(x.view(np.uint8), "F32"), f32_to_bf16_bits(x))
    h = x.astype(np.float16)
    assert np.array_equal(weights._to_bf16_bits(h.view(np.uint8), "F16"), f32_to_bf16_bits(h.astype(np.float32)))


@pytest.mark.gpu
def test_engine_loaded_from_safetensors_equals_generated(tmp_path, built_lib):
    """An engine created with seed 99 and then overwritten from a seed-0 checkpoint must behave
    exactly like an engine generated with seed 0 (every tensor really is replaced, including the
    fused QKV rows and the interleaved gate/up rows)."""
    from llmlb_b200 import ffi
    p = tmp_path / "tiny.safetensors"
    weights.write_safetensors(p, _tiny_bits(0))
    prompt = list(range(7, 60))
    with ffi.Engine(LLAMA_TINY, max_seqs=4, max_ctx=256, seed=0) as ref:
        want_lg = ref.debug_prefill_logits(prompt)
        ref.debug_reset()
        want_tok, _ = ref.generate(prompt, 12, ignore_eos=True)
    with ffi.Engine(LLAMA_TINY, max_seqs=4, max_ctx=256, seed=99) as e:
        before = e.debug_prefill_logits(prompt)
        e.debug_reset()
        assert np.abs(before - want_lg).max() > 0.1          # different weights -> different logits
        names = weights.load_safetensors(e, p)
        assert len(names) == 3 + 9 * LLAMA_TINY["n_layers"]
        got = e.read_tensor("model.layers.1.mlp.up_proj.weight", 1 << 20)
        assert np.array_equal(got, _tiny_bits(0)["model.layers.1.mlp.up_proj.weight"])
        lg = e.debug_prefill_logits(prompt)
        e.debug_reset()
        tok, _ = e.generate(prompt, 12, ignore_eos=True)
    assert np.array_equal(lg, want_lg) and tok == want_tok
