"""Build tools/fuzz/fuzz_host.cpp with ASan + UBSan and run every mode over freshly generated seeds.

    python tools/fuzz/run.py [--iters 20000] [--seed 1] [--keep]

Seeds: a one-layer Llama as GGUF with EVERY supported block format and the embedded llama-bpe tokenizer (written by
llama.cpp's own gguf.GGUFWriter), the same model as safetensors, the golden tokenizer.json, an OpenAI chat
completion body and an SSE transcript (chat chunks + a Responses event stream).  Exit status 0 = no sanitizer
report in any mode."""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
HOST = os.path.join(ROOT, "llmlb_b200", "host")
MODEL = dict(hidden=256, n_layers=1, n_heads=2, n_kv_heads=1, head_dim=128, ffn=256, vocab=3072, rope_theta=500000.0, rms_eps=1e-5)
QUANTS = ["Q8_0", "Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q4_K", "Q5_K", "Q6_K", "F16"]


def make_seeds(d):
    import gguf
    from gguf import quants as RQ
    from gguf_util import _random_blocks, _to_gguf_name
    from llmlb_b200 import weights
    from oracle.synth import f32_to_bf16_bits, synth_state_dict
    sd = synth_state_dict(MODEL, seed=3)
    tj_path = os.path.join(ROOT, "tests", "golden", "tokenizer_llama3_style.json")
    tj = json.load(open(tj_path, encoding="utf-8"))
    tokens, types = ["<unused_%d>" % i for i in range(MODEL["vocab"])], [5] * MODEL["vocab"]
    for tok, i in tj["model"]["vocab"].items():
        tokens[i], types[i] = tok, 1
    for a in tj["added_tokens"]:
        tokens[a["id"]], types[a["id"]] = a["content"], 3
    Q = gguf.GGMLQuantizationType

    def write_gguf(p, with_tokenizer):
        w = gguf.GGUFWriter(p, "llama")
        w.add_uint32("llama.block_count", MODEL["n_layers"]); w.add_uint32("llama.embedding_length", MODEL["hidden"])
        w.add_uint32("llama.feed_forward_length", MODEL["ffn"]); w.add_uint32("llama.attention.head_count", MODEL["n_heads"])
        w.add_uint32("llama.attention.head_count_kv", MODEL["n_kv_heads"]); w.add_float32("llama.rope.freq_base", MODEL["rope_theta"])
        w.add_float32("llama.attention.layer_norm_rms_epsilon", MODEL["rms_eps"]); w.add_uint32("llama.attention.key_length", MODEL["head_dim"])
        if with_tokenizer:
            w.add_tokenizer_model("gpt2"); w.add_tokenizer_pre("llama-bpe"); w.add_token_list(tokens); w.add_token_types(types)
            w.add_token_merges([m if isinstance(m, str) else " ".join(m) for m in tj["model"]["merges"]])
            w.add_bos_token_id(tokens.index("<|begin_of_text|>"))
        k = 0
        for name, t in sd.items():
            t = np.asarray(t, dtype=np.float32)
            if t.ndim == 1 or "norm" in name:
                w.add_tensor(_to_gguf_name(name), t)
                continue
            if name in ("model.embed_tokens.weight", "lm_head.weight"):
                qt = Q.Q8_0
            else:
                qt = getattr(Q, QUANTS[k % len(QUANTS)]); k += 1
            if qt == Q.F16:
                w.add_tensor(_to_gguf_name(name), t.astype(np.float16))
            elif qt.name.endswith("_K"):                          # the gguf package cannot quantise K blocks: random valid ones
                _, per, blk = _random_blocks(qt.name, t.size // 256, 17 + k)
                w.add_tensor(_to_gguf_name(name), blk.reshape(t.shape[0], -1), raw_shape=(t.shape[0], blk.size // t.shape[0]), raw_dtype=qt)
            else:
                q = RQ.quantize(t, qt)
                w.add_tensor(_to_gguf_name(name), q, raw_shape=q.shape, raw_dtype=qt)
        w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
        return p

    p = write_gguf(os.path.join(d, "seed.gguf"), True)
    # without the tokenizer arrays the tensor directory (names, dims, types, offsets) sits in the first 2 KB, where
    # most mutations land
    p2 = write_gguf(os.path.join(d, "seed_notok.gguf"), False)
    st = os.path.join(d, "seed.safetensors")
    small = {k: v for k, v in sd.items() if "embed" not in k and "lm_head" not in k}
    weights.write_safetensors(st, {k: f32_to_bf16_bits(np.asarray(v, dtype=np.float32)) for k, v in small.items()})
    body = {"id": "chatcmpl-1", "object": "chat.completion", "created": 1, "model": "m", "choices": [
        {"index": 0, "message": {"role": "assistant", "content": "héllo 😀 \"q\" \\ \n"}, "finish_reason": "stop"}],
        "usage": {"prompt_tokens": 12, "completion_tokens": 7, "total_tokens": 19, "nested": [1.5e3, -0.0, None, True, {"a": []}]},
        "response": {"usage": {"input_tokens": 3, "output_tokens": 4}}}
    js = os.path.join(d, "seed.json")
    open(js, "w").write(json.dumps(body))
    sse = os.path.join(d, "seed.sse")
    with open(sse, "w") as f:
        for piece in ("Hel", "lo", " wörld"):
            f.write("data: " + json.dumps({"id": "c", "object": "chat.completion.chunk", "choices": [{"index": 0, "delta": {"content": piece}}]}) + "\n\n")
        f.write("data: " + json.dumps({"choices": [], "usage": {"prompt_tokens": 5, "completion_tokens": 3}}) + "\n\n")
        f.write("event: response.output_text.delta\ndata: " + json.dumps({"type": "response.output_text.delta", "delta": "abc"}) + "\n\n")
        f.write("event: response.completed\ndata: " + json.dumps({"type": "response.completed", "response": {"usage": {"input_tokens": 1, "output_tokens": 2}}}) + "\n\n")
        f.write(": keep-alive\r\n\r\ndata: [DONE]\n\n")
    return {"gguf": p, "gguf_notok": p2, "safetensors": st, "tok": tj_path, "json": js, "sse": sse}


ODD = [None, True, False, -1, 0, 2 ** 31, 2 ** 40, 10 ** 30, 1.5, -0.0, "", " ", "\u0000", "a b c", "Ġ", [], {}, [[]], {"id": -5}, "<|begin_of_text|>"]


def structured_tokenizer_variants(tj_path, d, n, rs):
    """Type- and value-level damage that byte flips almost never produce while keeping the file valid JSON: ids that are
    negative / huge / fractional / strings, merges of the wrong arity, empty added tokens, missing sections."""
    base = json.load(open(tj_path, encoding="utf-8"))
    out = []
    for i in range(n):
        t = json.loads(json.dumps(base))
        for _ in range(rs.randint(1, 4)):
            kind = rs.randint(0, 9)
            try:
                v, m, a = t["model"]["vocab"], t["model"]["merges"], t["added_tokens"]
                keys = list(v)
                assert isinstance(v, dict) and isinstance(m, list) and isinstance(a, list) and keys and m and a
            except (TypeError, KeyError, AssertionError):
                break                                               # an earlier mutation removed a whole section: keep that damage
            if kind == 0:
                v[keys[rs.randint(len(keys))]] = ODD[rs.randint(len(ODD))]
            elif kind == 1:
                m[rs.randint(len(m))] = ODD[rs.randint(len(ODD))]
            elif kind == 2:
                a[rs.randint(len(a))][["id", "content", "special"][rs.randint(3)]] = ODD[rs.randint(len(ODD))]
            elif kind == 3:
                v[["", " ", "\u0000", "\ud7ff", "ĠĠĠĠ"][rs.randint(5)]] = int(rs.randint(0, 4000))
            elif kind == 4:
                sect = [("model",), ("model", "vocab"), ("model", "merges"), ("added_tokens",), ("pre_tokenizer",), ("post_processor",), ("decoder",)][rs.randint(7)]
                node = t
                for k in sect[:-1]:
                    node = node[k]
                if rs.randint(2):
                    node.pop(sect[-1], None)
                else:
                    node[sect[-1]] = ODD[rs.randint(len(ODD))]
            elif kind == 5:
                a.append({"id": int(rs.randint(0, 5000)), "content": ["", "<", "<|", "a", "<|eot_id|>"][rs.randint(5)], "special": bool(rs.randint(2))})
            elif kind == 6:
                for k in keys[:: max(1, rs.randint(1, 50))]:
                    v.pop(k, None)                                  # holes: merges now point at missing tokens, bytes are missing
            elif kind == 7:
                t["model"]["ignore_merges"] = ODD[rs.randint(len(ODD))]
            else:
                t["post_processor"] = {"type": "TemplateProcessing", "single": ODD[rs.randint(len(ODD))]}
        p = os.path.join(d, "tokvar_%04d.json" % i)
        with open(p, "w", encoding="utf-8") as f:
            json.dump(t, f, ensure_ascii=bool(rs.randint(2)))
        out.append(p)
    return out


def build(out):
    srcs = [os.path.join(ROOT, "tools", "fuzz", "fuzz_host.cpp")] + [os.path.join(HOST, f) for f in ("checkpoint.cpp", "gateway.cpp", "tokenizer.cpp", "anthropic.cpp")]
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-fno-omit-frame-pointer", "-pthread", *srcs, "-o", out])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--bin", default="")
    ap.add_argument("--only-structured", action="store_true")
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        exe = args.bin or os.path.join(d, "fuzz_host")
        if not args.bin:
            build(exe)
        seeds = make_seeds(d)
        env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0:allocator_may_return_null=1", UBSAN_OPTIONS="print_stacktrace=1")
        bad = 0
        for mode, seed, n in () if args.only_structured else (("ckpt", seeds["gguf"], args.iters), ("ckpt", seeds["gguf_notok"], args.iters), ("ckpt", seeds["safetensors"], args.iters),
                              ("tok", seeds["tok"], max(1, args.iters // 20)), ("json", seeds["json"], args.iters * 5), ("sse", seeds["sse"], args.iters * 5)):
            r = subprocess.run([exe, mode, seed, str(n), str(args.seed), os.path.join(d, "scratch.bin")], env=env, capture_output=True, text=True)
            print("%-4s %-28s rc=%d %s" % (mode, os.path.basename(seed), r.returncode, r.stdout.strip()))
            if r.returncode != 0:
                bad += 1
                print(r.stderr[-6000:])
        rs = np.random.RandomState(args.seed)
        variants = structured_tokenizer_variants(seeds["tok"], d, max(20, args.iters // 100), rs)
        loaded = 0
        for v in variants:
            r = subprocess.run([exe, "tokfile", v, "1", "0"], env=env, capture_output=True, text=True)
            loaded += "1 loaded" in r.stdout
            if r.returncode != 0:
                bad += 1
                print("tokfile %s rc=%d\n%s" % (os.path.basename(v), r.returncode, r.stderr[-4000:]))
                import shutil
                shutil.copy(v, "/tmp/" + os.path.basename(v))
        print("tokfile: %d structured variants, %d loaded, %s" % (len(variants), loaded, "clean" if not bad else "FINDINGS"))
        sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
