"""8B logits vs the CPU oracle for a few engine configurations / prompt lengths (debug)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmlb_b200 import ffi
from oracle import synth_native
from oracle.llama_ref import LlamaRef

cfg = ffi.LLAMA3_8B
synth_native.set_threads(synth_native.effective_cpus())
sd = synth_native.synth_state_dict_bits(cfg, seed=0)
prompts = {n: np.random.RandomState(5 + n).randint(0, cfg["vocab"], n).tolist() for n in (3, 21, 40)}
refs = {}
for n, p in prompts.items():
    t = time.time()
    r = LlamaRef(cfg, sd, emulate_bf16=True)
    refs[n] = (r.forward(p).numpy()[-1], r.forward([12345]).numpy()[-1])
    print("oracle n=%d %.1fs" % (n, time.time() - t), flush=True)
for impl in (0, 1):
    with ffi.Engine(cfg, max_seqs=4, max_ctx=1024, seed=0, gemm_impl=impl) as e:
        for n, p in prompts.items():
            e.debug_reset()
            lg = e.debug_prefill_logits(p)
            d1 = e.debug_decode_logits(12345)
            e.debug_reset()
            print("impl=%d n=%d  prefill mean|d|=%.4f max=%.3f   decode mean|d|=%.4f max=%.3f" % (
                impl, n, np.abs(lg - refs[n][0]).mean(), np.abs(lg - refs[n][0]).max(),
                np.abs(d1 - refs[n][1]).mean(), np.abs(d1 - refs[n][1]).max()), flush=True)
