"""Is a single greedy request bit-reproducible?  Runs the 8B engine on one GPU, generates the same
512-token prompt N times (interleaved with other requests so pages and slots get recycled) and
reports how many distinct token sequences came out, per configuration:
  graphs on / off, lookahead 2 / 1.
Run with PYTHONPATH=tools/_r1 to probe the round-1 library on the same box (A/B)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if not any(p.endswith("_r1") for p in sys.path):
    sys.path.insert(0, ROOT)
from llmlb_b200 import ffi  # noqa: E402


def main():
    cfg = ffi.LLAMA3_8B
    rs = np.random.RandomState(9)
    prompts = [rs.randint(0, cfg["vocab"], n).tolist() for n in (512, 77, 300)]
    print("library:", ffi.LIB_PATH)
    for graphs, look in ((True, 0), (False, 0), (True, 1)):
        with ffi.Engine(cfg, max_seqs=8, max_ctx=1024, seed=0, use_cuda_graphs=graphs, lookahead=look) as e:
            seqs = []
            for it in range(8):
                seqs.append(tuple(e.generate(prompts[0], 12, ignore_eos=True)[0]))
                if it % 2 == 0:
                    e.generate(prompts[1 + (it // 2) % 2], 5, ignore_eos=True)
            distinct = sorted(set(seqs), key=seqs.index)
            first_div = [next((i for i, (a, b) in enumerate(zip(s, seqs[0])) if a != b), None) for s in seqs]
            # how close was the call at the first divergence?  teacher-force the parity hooks along run 0
            note = ""
            if len(distinct) > 1:
                d = min(x for x in first_div if x is not None)
                lg = e.debug_prefill_logits(prompts[0])
                for t in seqs[0][:d]:
                    lg = e.debug_decode_logits(t)
                e.debug_reset()
                cands = sorted({s[d] for s in seqs})
                note = " | logits of the candidates at step %d: %s (max %.4f)" % (d, ["%d: %.4f" % (c, lg[c]) for c in cands], float(lg.max()))
            print("graphs=%s lookahead=%d: %d distinct sequences in %d runs, first divergence per run %s%s"
                  % (graphs, look or 2, len(distinct), len(seqs), first_div, note))


if __name__ == "__main__":
    main()
