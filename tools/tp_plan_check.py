"""Tensor-parallel serving through the plan channel (torchrun, one rank per GPU): rank 0 alone
submits — staggered arrivals, different lengths, a cancellation — and the follower ranks replay
its scheduler log.  Checks: no deadlock, identical step / token counters on every rank, the same
request alone is bit-identical to the lock-step (pause + barrier) path, and repeatable."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmlb_b200 import build, ffi  # noqa: E402


def make_engine(model, rank, world, local, cpu):
    eng = ffi.Engine(model, model_id="tp-plan", device=local, tp_rank=rank, tp_size=world, max_seqs=8, max_ctx=1024, seed=0)
    handles = [None] * world
    dist.all_gather_object(handles, eng.tp_export(), group=cpu)
    eng.tp_import(handles)
    dist.barrier(group=cpu)
    return eng


def drain(eng, rid):
    toks = []
    while True:
        ev = eng.poll(rid, cap=256, timeout_ms=-1)
        toks += [e["token_id"] for e in ev if e["token_id"] >= 0]
        if ev and ev[-1]["finish_reason"]:
            return toks, ev[-1]["finish_reason"]


def main():
    build.build()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # control-plane syncs go over gloo: a NCCL barrier is a GPU kernel that would sit on the
    # follower's device, spinning, for as long as rank 0 serves
    cpu = dist.new_group(backend="gloo")
    barrier = lambda: dist.barrier(group=cpu)
    model = ffi.LLAMA_MID if hasattr(ffi, "LLAMA_MID") else ffi.LLAMA_TINY
    rs = np.random.RandomState(3)
    prompts = [rs.randint(0, model["vocab"], n).tolist() for n in (64, 200, 17, 333, 90, 41)]
    gens = [24, 40, 12, 30, 900, 20]        # request 4 is long: it gets cancelled mid-flight
    name = "llmlb_plan_%s" % os.environ.get("MASTER_PORT", "0")

    # ---- A: plan channel ----
    eng = make_engine(model, rank, world, local, cpu)
    if rank == 0:
        eng.tp_plan_channel(name)
    barrier()
    if rank != 0:
        eng.tp_plan_channel(name)
    barrier()
    results = {}
    if rank == 0:
        t0 = time.time()
        solo, fr = drain(eng, eng.submit(prompts[0], gens[0], ignore_eos=True))     # alone: batch 1
        assert len(solo) == gens[0]
        solo2, _ = drain(eng, eng.submit(prompts[0], gens[0], ignore_eos=True))
        rids = []
        for i, (p, g) in enumerate(zip(prompts, gens)):                             # staggered arrivals
            rids.append(eng.submit(p, g, ignore_eos=True))
            time.sleep(0.004 * (i % 3))
        time.sleep(0.005)
        eng.cancel(rids[4])                                                          # mid-flight cancellation
        outs = [drain(eng, r) for r in rids]
        for r in rids:
            eng.release(r)
        results = {"solo": solo, "solo2": solo2, "lens": [len(t) for t, _ in outs], "reasons": [f for _, f in outs], "secs": time.time() - t0}
    else:
        try:
            eng.submit(prompts[0], 4)
            raise SystemExit("follower accepted a submit")
        except ffi.LlmlbError:
            pass
    barrier()          # followers keep replaying until rank 0 is done
    h = eng.health()
    counters = [None] * world
    dist.all_gather_object(counters, {k: h[k] for k in ("steps_decode", "tokens_decode", "tokens_prefill", "kernel_launches")}, group=cpu)
    eng.close()
    barrier()

    # ---- B: lock-step reference for the solo request ----
    eng = make_engine(model, rank, world, local, cpu)
    eng.pause(True)
    rid = eng.submit(prompts[0], gens[0], ignore_eos=True)
    barrier()
    eng.pause(False)
    lock, _ = drain(eng, rid)
    barrier()
    eng.close()

    if rank == 0:
        same = all(c == counters[0] for c in counters)
        ok = (same and results["solo"] == lock and results["solo"] == results["solo2"]
              and results["reasons"][4] == 3 and results["lens"][4] < gens[4]
              and all(results["lens"][i] == gens[i] and results["reasons"][i] == 2 for i in (0, 1, 2, 3, 5)))
        print("plan channel tp=%d: counters equal on all ranks: %s %s" % (world, same, counters[0]))
        print("solo == lock-step: %s, repeatable: %s, lens %s, reasons %s, %.2f s" % (
            results["solo"] == lock, results["solo"] == results["solo2"], results["lens"], results["reasons"], results["secs"]))
        print("PLAN_CHANNEL_OK" if ok else "PLAN_CHANNEL_FAILED")
        if not ok:
            sys.exit(1)
    barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
