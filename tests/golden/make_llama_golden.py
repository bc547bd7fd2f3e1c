"""Generates tests/golden/llama_tiny_golden.npz by running HuggingFace transformers'
LlamaForCausalLM (fp32, CPU) on the synthetic tiny-geometry weights.  Run in the build container
(transformers 5.5 is importable there); the GPU box only reads the committed .npz.

    python tests/golden/make_llama_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.synth import synth_state_dict  # noqa: E402

TINY = dict(hidden=512, n_layers=2, n_heads=8, n_kv_heads=2, head_dim=128, ffn=1024, vocab=2048,
            rope_theta=500000.0, rms_eps=1e-5)


def main():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=TINY["hidden"], num_hidden_layers=TINY["n_layers"],
                      num_attention_heads=TINY["n_heads"], num_key_value_heads=TINY["n_kv_heads"],
                      head_dim=TINY["head_dim"], intermediate_size=TINY["ffn"],
                      vocab_size=TINY["vocab"], rope_theta=TINY["rope_theta"],
                      rms_norm_eps=TINY["rms_eps"], max_position_embeddings=8192,
                      tie_word_embeddings=False, attention_bias=False, mlp_bias=False,
                      hidden_act="silu")
    cfg._attn_implementation = "eager"
    model = LlamaForCausalLM(cfg).to(torch.float32).eval()
    sd = {k: torch.from_numpy(v.copy()) for k, v in synth_state_dict(TINY, seed=0).items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in m or "inv_freq" in m for m in missing), missing
    out = {}
    rng = np.random.RandomState(1000)
    for name, n_prompt, n_new in [("a", 37, 8), ("b", 130, 6), ("c", 1, 4)]:
        prompt = rng.randint(0, TINY["vocab"], size=n_prompt).astype(np.int64)
        ids = torch.from_numpy(prompt)[None]
        with torch.no_grad():
            full = model(ids).logits[0].float().numpy()
            gen = model.generate(ids, max_new_tokens=n_new, do_sample=False, use_cache=True,
                                 pad_token_id=0, eos_token_id=None)[0, n_prompt:].numpy()
        out["prompt_" + name] = prompt.astype(np.int32)
        out["logits_" + name] = full[-4:].astype(np.float32) if n_prompt >= 4 else full.astype(np.float32)
        out["greedy_" + name] = gen.astype(np.int32)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "llama_tiny_golden.npz"), **out)
    print("wrote", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
