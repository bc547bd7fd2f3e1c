"""Tensor-parallel parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/tp_check.py
Every rank builds the tiny geometry sharded tp=world, wires the peer-memory exchange through the
C ABI, and rank 0 compares prefill logits + greedy tokens with the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llmlb_b200 import ffi  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = dict(ffi.LLAMA_TINY)
    cfg["n_kv_heads"] = 8 if world > 2 else 2   # kv heads must divide by tp
    cfg["n_heads"] = 32 if world > 2 else 8
    cfg["hidden"] = 1024 if world > 2 else 512
    eng = ffi.Engine(cfg, device=local, tp_rank=rank, tp_size=world, max_seqs=8, max_ctx=512, seed=0)
    handles = [None] * world
    dist.all_gather_object(handles, eng.tp_export())
    eng.tp_import(handles)
    dist.barrier()
    prompt = np.random.RandomState(3).randint(0, cfg["vocab"], 150).tolist()
    lg = eng.debug_prefill_logits(prompt)
    step = [eng.debug_decode_logits(7), eng.debug_decode_logits(99)]
    eng.debug_reset()
    dist.barrier()
    # identical queues on every rank, then go
    eng.pause(True)
    rids = [eng.submit(prompt[: 40 + 30 * i], 24, ignore_eos=True) for i in range(3)]
    dist.barrier()
    eng.pause(False)
    toks = []
    for r in rids:
        out = []
        while True:
            ev = eng.poll(r, timeout_ms=-1)
            out += [e["token_id"] for e in ev if e["token_id"] >= 0]
            if ev and ev[-1]["finish_reason"]:
                break
        toks.append(out)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lg[:64].tolist(), toks))
    ok = True
    if rank == 0:
        from oracle.llama_ref import LlamaRef
        from oracle.synth import synth_state_dict
        ref = LlamaRef(cfg, synth_state_dict(cfg, 0))
        rl = ref.forward(prompt).numpy()[-1]
        e0 = float(np.abs(lg - rl).max())
        e1 = float(np.abs(step[0] - ref.forward([7]).numpy()[-1]).max())
        e2 = float(np.abs(step[1] - ref.forward([99]).numpy()[-1]).max())
        same = all(g == gathered[0] for g in gathered)
        # teacher-forced: every engine token must be a (near-)arg-max of the oracle's logits
        agree, near = 0, True
        sd = synth_state_dict(cfg, 0)
        for i, t in enumerate(toks):
            r2 = LlamaRef(cfg, sd)
            cur = r2.forward(prompt[: 40 + 30 * i]).numpy()[-1]
            for tok in t:
                near &= bool(cur[tok] >= cur.max() - 0.06)
                agree += int(tok == int(np.argmax(cur)))
                cur = r2.forward([tok]).numpy()[-1]
        print("tp=%d max|dlogit| prefill %.4g decode %.4g %.4g; ranks identical: %s; top-1 agreement %d/72, all near-argmax: %s"
              % (world, e0, e1, e2, same, agree, near))
        ok = e0 < 0.03 and e1 < 0.03 and e2 < 0.03 and same and near and agree >= 66
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
