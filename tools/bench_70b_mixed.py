"""BASELINE.json configs[4]: Llama-3-70B bf16, tensor parallel over all ranks, paged KV, 8192-token
contexts, mixed prefill + decode continuous batch.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_70b_mixed.py

Workload: R requests (default 48) with prompt lengths drawn U[512, 8064] (seed 7), 128 greedy tokens
each, EOS disabled; Poisson arrivals (seed 7, --rate req/s; 0 = all at t = 0) so that chunked
prefills of late arrivals interleave with the decode steps of earlier ones.  KV pages are
over-subscribed on purpose with --kv-frac < 1 (pages for that fraction of sum(prompt + out)): the
scheduler then has to allocate mid-flight and preempt / recompute.

Every rank submits the identical queue at identical (virtual) times: arrivals are released in
lock-step rounds (pause -> submit the round's arrivals -> barrier -> resume), the same way
bench.py drives tp > 1.  Rank 0 prints one JSON line.

--layers N truncates the stack (debug runs on fewer GPUs; NOT the benchmark config)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llmlb_b200 import ffi  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--requests", type=int, default=48)
    ap.add_argument("--gen", type=int, default=128)
    ap.add_argument("--min-prompt", type=int, default=512)
    ap.add_argument("--max-prompt", type=int, default=8064)
    ap.add_argument("--rate", type=float, default=4.0, help="Poisson arrival rate in requests/s (0: all at t=0)")
    ap.add_argument("--max-seqs", type=int, default=32)
    ap.add_argument("--max-step-tokens", type=int, default=2048)
    ap.add_argument("--kv-frac", type=float, default=0.0, help=">0: KV pool = this fraction of the whole job's tokens (forces eviction)")
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--model", default="70b", choices=["70b", "8b"])
    args = ap.parse_args()
    import torch
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # arrivals are announced over a CPU-side group: a NCCL broadcast is a kernel that sits on an SM of every
        # rank until rank 0 reaches it, and the engine's narrow GEMMs need all SMs co-resident (the first 8-GPU run
        # of this tool broadcast over NCCL while serving and tripped the 2 s spin guard of the in-kernel K-split)
        cpu_group = dist.new_group(backend="gloo")
    model = dict(ffi.LLAMA3_70B if args.model == "70b" else ffi.LLAMA3_8B)
    if args.layers:
        model["n_layers"] = args.layers
    rs = np.random.RandomState(7)
    lens = rs.randint(args.min_prompt, args.max_prompt + 1, args.requests)
    gaps = rs.exponential(1.0 / args.rate, args.requests) if args.rate > 0 else np.zeros(args.requests)
    arrive = np.cumsum(gaps) - gaps[0]
    prompts = [np.random.RandomState(9000 + i).randint(0, model["vocab"], int(n)).astype("int32").tolist() for i, n in enumerate(lens)]
    max_ctx = 8192
    kv_pages = 0
    if args.kv_frac > 0:
        kv_pages = int(args.kv_frac * sum((int(n) + args.gen + 63) // 64 for n in lens))
    eng = ffi.Engine(model, model_id="llama-3-70b-synthetic", device=local, tp_rank=rank, tp_size=world, max_seqs=args.max_seqs,
                     max_ctx=max_ctx, kv_pages=kv_pages, max_step_tokens=args.max_step_tokens, seed=1)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, eng.tp_export())
        eng.tp_import(handles)
        dist.barrier()

    def barrier():
        if world > 1:
            dist.barrier()

    def drive(idx, clocked):
        """idx: request indices in arrival order.  Lock-step release rounds: rank 0's clock decides which arrivals
        are due, the count is broadcast, every rank submits them while paused."""
        rids, evs, t_sub = {}, {}, {}
        nxt, t0 = 0, time.perf_counter()
        pending = set()
        while nxt < len(idx) or pending:
            due = nxt
            if nxt < len(idx):
                now = time.perf_counter() - t0
                while due < len(idx) and (not clocked or arrive[idx[due]] - arrive[idx[0]] <= now):
                    due += 1
                if world > 1:
                    t = torch.tensor([due])
                    dist.broadcast(t, 0, group=cpu_group)
                    due = int(t.item())
            if due > nxt:
                if world > 1:
                    eng.pause(True)
                for j in range(nxt, due):
                    i = idx[j]
                    t_sub[i] = time.perf_counter()
                    rids[i] = eng.submit(prompts[i], args.gen, ignore_eos=True)
                    evs[i] = []
                    pending.add(i)
                if world > 1:
                    dist.barrier(group=cpu_group)
                    eng.pause(False)
                nxt = due
            for i in list(pending):
                got = eng.poll(rids[i], cap=256, timeout_ms=1)
                evs[i].extend(got)
                if got and got[-1]["finish_reason"]:
                    pending.discard(i)
                    eng.release(rids[i])
            if not pending and nxt < len(idx) and clocked:
                time.sleep(0.002)
        return evs, time.perf_counter() - t0

    # warm-up: two short requests (graph capture for small widths happens on demand anyway)
    warm = [0, 1]
    save = [prompts[0], prompts[1]]
    prompts[0], prompts[1] = prompts[0][:600], prompts[1][:700]
    drive(warm, False)
    prompts[0], prompts[1] = save
    torch.cuda.synchronize()
    barrier()
    h0 = eng.health()
    evs, wall = drive(list(range(args.requests)), args.rate > 0)
    torch.cuda.synchronize()
    h1 = eng.health()
    d = lambda k: h1[k] - h0[k]
    vals = torch.tensor([d("gpu_ms_decode"), d("gpu_ms_prefill"), wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    dec_ms, pre_ms, wall_ms = [float(v) for v in vals]
    toks = {i: [e for e in v if e["token_id"] >= 0] for i, v in evs.items()}
    all_ids = [[e["token_id"] for e in toks[i]] for i in range(args.requests)]
    same = True
    if world > 1:
        g = [None] * world
        dist.all_gather_object(g, all_ids)
        same = all(x == g[0] for x in g)
    if rank == 0:
        hbm_peak, tf_burst, _, src = bench.measured_peaks()
        ttft = [toks[i][0]["t_ms"] for i in toks]
        tps = [len(toks[i]) / (toks[i][-1]["t_ms"] / 1e3) for i in toks]
        steps_dec = max(1, d("steps_decode"))
        avg_batch = d("tokens_decode") / steps_dec
        avg_ctx = float(np.mean(lens)) + args.gen / 2
        bytes_step = bench.algorithmic_bytes_per_decode_step(model, avg_batch, avg_ctx, world)
        step_s = dec_ms / 1e3 / steps_dec
        fl = sum(bench.prefill_flops(model, int(n)) for n in lens) / world
        line = {"workload": "BASELINE.json configs[4]: Llama-3-%s bf16, tp%d, paged KV (64-token pages), prompts U[%d,%d] + %d out, %d requests, %s, chunked prefill %d tokens/step, max %d running"
                            % (args.model.upper(), world, args.min_prompt, args.max_prompt, args.gen, args.requests,
                               ("Poisson arrivals %.1f req/s (seed 7)" % args.rate) if args.rate > 0 else "all submitted at t=0", args.max_step_tokens, args.max_seqs),
                "model_geometry": model, "n_gpus": world, "requests": args.requests, "prompt_tokens_total": int(lens.sum()),
                "generated_tokens_total": int(sum(len(v) for v in toks.values())),
                "all_requests_complete": all(len(toks[i]) == args.gen for i in toks), "ranks_identical": same,
                "decode_tok_s": d("tokens_decode") / (dec_ms / 1e3), "decode_steps": int(steps_dec), "avg_decode_batch": avg_batch,
                "decode_ms_per_step": step_s * 1e3,
                "decode_roofline": {"bound": "hbm", "achieved": bytes_step / step_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": bytes_step / step_s / 1e9 / hbm_peak, "peak_source": src,
                                    "what": "%.2f GB algorithmic per step per GPU (weights/tp + avg batch %.1f x avg ctx %.0f KV/tp)" % (bytes_step / 1e9, avg_batch, avg_ctx)},
                "prefill_tok_s": d("tokens_prefill") / (pre_ms / 1e3), "prefill_tokens": int(d("tokens_prefill")),
                "prefill_roofline": {"bound": "tensor", "achieved": fl * (d("tokens_prefill") / float(lens.sum())) / (pre_ms / 1e3) / 1e12, "peak": tf_burst,
                                     "unit": "TFLOP/s per GPU", "frac": fl * (d("tokens_prefill") / float(lens.sum())) / (pre_ms / 1e3) / 1e12 / tf_burst},
                "recomputed_prompt_tokens": int(d("tokens_prefill") - lens.sum()),
                "whole_job_tok_s": (int(lens.sum()) + args.requests * args.gen) / (wall_ms / 1e3),
                "whole_job_generated_tok_s": args.requests * args.gen / (wall_ms / 1e3), "wall_s": wall_ms / 1e3,
                "gpu_busy_frac": (dec_ms + pre_ms) / wall_ms,
                "ttft_ms_p50": bench.pct(ttft, 0.5), "ttft_ms_p95": bench.pct(ttft, 0.95),
                "request_tps_mean": float(np.mean(tps)), "request_tps_p50": bench.pct(tps, 0.5),
                "kv_pages": kv_pages or "default (max_seqs x max_ctx)", "gpu_launches": int(d("kernel_launches")),
                "preemptions": int(d("preemptions")), "total_kv_pages": int(h1["total_kv_pages"])}
        print(json.dumps(line))
        sys.stdout.flush()
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
