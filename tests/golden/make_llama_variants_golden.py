"""More geometries for the model oracle's pin against Hugging Face transformers' LlamaForCausalLM
(fp32, CPU, eager attention): multi-head, multi-query and 4:1 grouped attention, head_dim 64 and
128, RoPE theta 1e4 / 5e5, RMSNorm eps 1e-6 / 1e-5, 1-3 layers.  Writes
tests/golden/llama_variants_golden.npz; run in the build container:

    python tests/golden/make_llama_variants_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.synth import synth_state_dict  # noqa: E402

VARIANTS = {
    "mha64": dict(hidden=256, n_layers=3, n_heads=4, n_kv_heads=4, head_dim=64, ffn=512, vocab=640, rope_theta=10000.0, rms_eps=1e-6),
    "mqa128": dict(hidden=512, n_layers=1, n_heads=4, n_kv_heads=1, head_dim=128, ffn=768, vocab=1024, rope_theta=500000.0, rms_eps=1e-5),
    "gqa4": dict(hidden=1024, n_layers=2, n_heads=8, n_kv_heads=2, head_dim=128, ffn=1536, vocab=512, rope_theta=500000.0, rms_eps=1e-5),
}


def main():
    from transformers import LlamaConfig, LlamaForCausalLM
    out = {}
    for name, M in VARIANTS.items():
        cfg = LlamaConfig(hidden_size=M["hidden"], num_hidden_layers=M["n_layers"], num_attention_heads=M["n_heads"],
                          num_key_value_heads=M["n_kv_heads"], head_dim=M["head_dim"], intermediate_size=M["ffn"],
                          vocab_size=M["vocab"], rope_theta=M["rope_theta"], rms_norm_eps=M["rms_eps"],
                          max_position_embeddings=4096, tie_word_embeddings=False, attention_bias=False, mlp_bias=False,
                          hidden_act="silu")
        cfg._attn_implementation = "eager"
        model = LlamaForCausalLM(cfg).to(torch.float32).eval()
        sd = {k: torch.from_numpy(v.copy()) for k, v in synth_state_dict(M, seed=7).items()}
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
        rng = np.random.RandomState(hash(name) % 1000)
        prompt = rng.randint(0, M["vocab"], size=77).astype(np.int64)
        ids = torch.from_numpy(prompt)[None]
        with torch.no_grad():
            full = model(ids).logits[0].float().numpy()
            gen = model.generate(ids, max_new_tokens=10, do_sample=False, use_cache=True, pad_token_id=0, eos_token_id=None)[0, len(prompt):].numpy()
        out["prompt_" + name] = prompt.astype(np.int32)
        out["logits_" + name] = full[-6:].astype(np.float32)
        out["greedy_" + name] = gen.astype(np.int32)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "llama_variants_golden.npz"), **out)
    print("wrote", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
