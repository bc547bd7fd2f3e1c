"""Oracle: CPU restatement of the Llama-3 decoder forward (test infrastructure, not product).

PARITY ANCHOR.  The reference repository (akiojin/llmlb @ a61826a) contains no model arithmetic:
the path behind its HTTP hop (llmlb/src/api/openai.rs:995-1005, llmlb/src/api/proxy.rs:372-401)
runs in external, un-vendored engines (xLLM / llama.cpp / Ollama / vLLM), none of
which is pinned in Cargo.lock.  This file therefore restates the PUBLISHED Llama-3 algorithm
(Llama 3 model card; identical to transformers `LlamaForCausalLM`, v5.5 here) and is pinned
against transformers itself: tests/golden/llama_tiny_golden.npz is produced by
tests/golden/make_llama_golden.py importing transformers in the build container, and
tests/test_oracle_llama.py checks this restatement against it.

Algorithm (per layer): h = x + Wo·Attn(RoPE(Wq·n1(x)), RoPE(Wk·n1(x)), Wv·n1(x));
x' = h + Wd·(silu(Wg·n2(h)) * (Wu·n2(h))); n(x) = x * rsqrt(mean(x^2)+eps) * g;
RoPE rotate-half pairing (i, i+64), theta 5e5; GQA causal softmax(QK^T/sqrt(128)); untied lm_head.

`emulate_bf16=True` additionally rounds activations to bf16 at the points where the CUDA path
stores bf16 (normed input, q/k/v, attention output, SwiGLU output) — a tight checker; the default
fp32 mode is the numerically "true" answer used to state the tolerance.
"""
import math

import numpy as np
import torch


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


class LlamaRef:
    def __init__(self, cfg, state_dict, emulate_bf16=False, weight_dtype=torch.float32,
                 threads=None):
        self.cfg = dict(cfg)
        self.emulate = emulate_bf16
        self.wd = weight_dtype
        if threads:
            torch.set_num_threads(threads)
        self.w = {}
        for k, v in state_dict.items():
            if getattr(v, "dtype", None) in ("bf16_bits", "q4_0"):   # oracle.synth_native.Bf16Weight / Q4Weight: C kernel path
                self.w[k] = v
            else:
                t = torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v
                self.w[k] = t.to(weight_dtype)
        hd = cfg["head_dim"]
        self.inv_freq = 1.0 / (cfg["rope_theta"] ** (torch.arange(0, hd, 2, dtype=torch.float64) / hd))
        self.reset()

    def reset(self):
        L = self.cfg["n_layers"]
        self.k_cache = [None] * L
        self.v_cache = [None] * L
        self.pos = 0

    # ---- pieces ----
    def _norm(self, x, g):
        var = x.pow(2).mean(-1, keepdim=True)
        y = x * torch.rsqrt(var + self.cfg["rms_eps"]) * g.to(torch.float32)
        return _bf16_round(y) if self.emulate else y

    def _mm(self, x, w):
        if getattr(w, "dtype", None) == "bf16_bits":  # CPU-baseline mode: bf16 weights, C/OpenMP GEMV
            from .synth_native import linear_bf16
            return torch.from_numpy(linear_bf16(w.bits, x.numpy()))
        if getattr(w, "dtype", None) == "q4_0":       # CPU baseline at the reference configuration's weight width (ggml Q4_0)
            from .synth_native import linear_q4_0, linear_q4_0_q8_0
            # llama.cpp's scheme by default: activations quantised to Q8_0, integer block dot products
            fn = linear_q4_0_q8_0 if getattr(w, "int_dot", True) else linear_q4_0
            return torch.from_numpy(fn(w.blocks, w.k, x.numpy()))
        if w.dtype == torch.float32:
            return x @ w.t()
        return (x.to(w.dtype) @ w.t()).to(torch.float32)

    def _rope(self, x, positions):
        # x [T, heads, head_dim]; rotate-half: (x[..., :hd/2], x[..., hd/2:]) pairs
        ang = positions.to(torch.float64)[:, None] * self.inv_freq[None, :]
        cos = torch.cos(ang).to(torch.float32)[:, None, :]
        sin = torch.sin(ang).to(torch.float32)[:, None, :]
        half = x.shape[-1] // 2
        a, b = x[..., :half], x[..., half:]
        return torch.cat([a * cos - b * sin, b * cos + a * sin], dim=-1)

    @torch.no_grad()
    def forward(self, ids):
        """Feeds len(ids) new tokens after the cached context; returns fp32 logits [T, vocab]."""
        c = self.cfg
        nh, nkv, hd = c["n_heads"], c["n_kv_heads"], c["head_dim"]
        ids = torch.as_tensor(ids, dtype=torch.long)
        T = ids.numel()
        positions = torch.arange(self.pos, self.pos + T)
        emb = self.w["model.embed_tokens.weight"]
        if getattr(emb, "dtype", None) == "bf16_bits":
            from .synth import bf16_bits_to_f32
            x = torch.from_numpy(bf16_bits_to_f32(emb.bits[ids.numpy()]).copy())
        else:
            x = emb[ids].to(torch.float32)
        for l in range(c["n_layers"]):
            p = "model.layers.%d." % l
            y = self._norm(x, self.w[p + "input_layernorm.weight"])
            q = self._mm(y, self.w[p + "self_attn.q_proj.weight"]).view(T, nh, hd)
            k = self._mm(y, self.w[p + "self_attn.k_proj.weight"]).view(T, nkv, hd)
            v = self._mm(y, self.w[p + "self_attn.v_proj.weight"]).view(T, nkv, hd)
            if self.emulate:
                q, k, v = _bf16_round(q), _bf16_round(k), _bf16_round(v)
            q, k = self._rope(q, positions), self._rope(k, positions)
            if self.emulate:
                q, k = _bf16_round(q), _bf16_round(k)
            self.k_cache[l] = k if self.k_cache[l] is None else torch.cat([self.k_cache[l], k], 0)
            self.v_cache[l] = v if self.v_cache[l] is None else torch.cat([self.v_cache[l], v], 0)
            K, V = self.k_cache[l], self.v_cache[l]
            S = K.shape[0]
            g = nh // nkv
            Kh = K.repeat_interleave(g, dim=1)  # [S, nh, hd]
            Vh = V.repeat_interleave(g, dim=1)
            scores = torch.einsum("thd,shd->hts", q, Kh) / math.sqrt(hd)
            kv_pos = torch.arange(S)
            mask = kv_pos[None, :] > positions[:, None]
            scores = scores.masked_fill(mask[None], float("-inf"))
            pr = torch.softmax(scores, dim=-1)
            a = torch.einsum("hts,shd->thd", pr, Vh).reshape(T, nh * hd)
            if self.emulate:
                a = _bf16_round(a)
            x = x + self._mm(a, self.w[p + "self_attn.o_proj.weight"])
            y = self._norm(x, self.w[p + "post_attention_layernorm.weight"])
            gate = self._mm(y, self.w[p + "mlp.gate_proj.weight"])
            up = self._mm(y, self.w[p + "mlp.up_proj.weight"])
            hmid = torch.nn.functional.silu(gate) * up
            if self.emulate:
                hmid = _bf16_round(hmid)
            x = x + self._mm(hmid, self.w[p + "mlp.down_proj.weight"])
        self.pos += T
        y = self._norm(x, self.w["model.norm.weight"])
        return self._mm(y, self.w["lm_head.weight"])

    @torch.no_grad()
    def greedy(self, prompt, n_new):
        """Greedy continuation; returns (tokens, logits_per_step [n_new, vocab])."""
        self.reset()
        lg = self.forward(prompt)[-1]
        toks, all_lg = [], []
        for _ in range(n_new):
            all_lg.append(lg.clone())
            t = int(torch.argmax(lg))
            toks.append(t)
            if len(toks) == n_new:
                break
            lg = self.forward([t])[-1]
        return toks, torch.stack(all_lg)
