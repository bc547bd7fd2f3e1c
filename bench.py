#!/usr/bin/env python
"""bench.py — decode tok/s + prefill tok/s for Llama-3-8B bf16, 512-in/128-out (BASELINE.json).

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  One "step" = one whole request through the engine: a 512-token prompt is
prefilled and 128 tokens are decoded greedily (EOS disabled).

  value        decode tok/s from CUDA-event time of the decode steps (inputs resident in HBM)
  prefill      prefill tok/s from CUDA-event time of the prefill steps (+ tensor roofline)
  e2e          the same decode metric measured by a client through the C ABI with HOST buffers:
               llmlb_request_submit(host prompt ids) -> llmlb_request_poll(token events)
  roofline     HBM roofline of the decode step (GEMV-dominated) and of the dominant GEMV kernel
  cpu_baseline the oracle (CPU restatement) timed on the host cores on a bounded sample

N > 1 (torchrun, one rank per GPU): the model is sharded tensor-parallel over the N ranks
(peer-memory all-reduce after the O and down projections) and the SAME workload is timed
("scaling": "strong").

`--impl reference` times the reference path's CPU stand-in: the oracle port of the model on the
host cores (the reference repository proxies to an external CPU llama.cpp/Ollama server that is
not in the tree and cannot be installed here — SURVEY.md §8c/8d).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout and would precede the JSON line

PROMPT, GEN = 512, 128


def algorithmic_bytes_per_decode_step(m, batch, ctx, tp=1):
    """all matmul weights + lm_head once per step, plus the KV of every sequence."""
    H, L, F, V = m["hidden"], m["n_layers"], m["ffn"], m["vocab"]
    nq, nkv, hd = m["n_heads"], m["n_kv_heads"], m["head_dim"]
    per_layer = H * (nq + 2 * nkv) * hd + nq * hd * H + 3 * H * F
    w = (per_layer * L + V * H) * 2
    kv = 2 * L * nkv * hd * 2 * ctx * batch
    return (w + kv) / tp


def prefill_flops(m, n):
    H, L, F, V = m["hidden"], m["n_layers"], m["ffn"], m["vocab"]
    nq, nkv, hd = m["n_heads"], m["n_kv_heads"], m["head_dim"]
    per_layer = H * (nq + 2 * nkv) * hd + nq * hd * H + 3 * H * F
    attn = L * 4 * n * n * nq * hd / 2
    return 2.0 * per_layer * L * n + 2.0 * V * H + attn


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (recipe: B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_prompt(i, vocab, n=PROMPT):
    import numpy as np
    return np.random.RandomState(1000 + i).randint(0, vocab, n).astype("int32").tolist()


# ------------------------------------------------------------------------------ CPU arms ----
def cpu_reference_run(model, steps, warmup, sample_prompt=32, sample_gen=4, verbose=False):
    """Oracle port on the host cores: synthetic bf16 weights (C generator), torch CPU matmuls.
    Bounded sample of the 512/128 workload: `sample_prompt` prompt tokens + `sample_gen` decoded
    tokens per step, full 8B geometry (nothing skipped)."""
    import torch
    from oracle import synth_native
    from oracle.llama_ref import LlamaRef
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    cores = synth_native.effective_cpus()
    torch.set_num_threads(cores)
    synth_native.set_threads(cores)
    t0 = time.time()
    sd = synth_native.synth_state_dict_bits(model, seed=0)
    ref = LlamaRef(model, sd)
    t_load = time.time() - t0
    # probe one lm_head-sized GEMV and shrink the sample if this host is slow, so the run stays
    # inside its time box (~30 s of CPU work per step) whatever the core quota is
    import numpy as np
    probe_w = sd["lm_head.weight"].bits
    px = np.zeros((1, probe_w.shape[1]), dtype=np.float32)
    synth_native.linear_bf16(probe_w, px)
    tp0 = time.time(); synth_native.linear_bf16(probe_w, px); probe = time.time() - tp0
    est_token_s = probe * (algorithmic_bytes_per_decode_step(model, 0, 0) / (probe_w.size * 2.0))
    if est_token_s * (sample_prompt / 4.0 + sample_gen) > 30.0:
        sample_gen = max(1, min(sample_gen, int(10.0 / max(est_token_s, 1e-3))))
        sample_prompt = max(4, min(sample_prompt, int(4 * 15.0 / max(est_token_s, 1e-3))))
    dec_tok = dec_s = pre_tok = pre_s = 0.0
    for it in range(warmup + steps):
        prompt = make_prompt(it, model["vocab"], sample_prompt)
        ref.reset()
        a = time.time()
        lg = ref.forward(prompt)[-1]
        b = time.time()
        tok = int(torch.argmax(lg))
        for _ in range(sample_gen):
            lg = ref.forward([tok])[-1]
            tok = int(torch.argmax(lg))
        c = time.time()
        if it >= warmup:
            pre_tok += sample_prompt; pre_s += b - a
            dec_tok += sample_gen; dec_s += c - b
    return {"decode_tok_s": dec_tok / dec_s, "prefill_tok_s": pre_tok / pre_s, "cores": cores,
            "weights_s": t_load, "ms_per_step": 1e3 * (pre_s + dec_s) / max(1, steps),
            "sample": "%d prompt + %d decoded tokens per step, full Llama-3-8B geometry, bf16 weights, C/OpenMP fp32-accumulate linears (oracle/llama_cpu.c), %d threads" % (sample_prompt, sample_gen, cores)}


def run_reference(args):
    from llmlb_b200.ffi import LLAMA3_8B, LLAMA_TINY
    model = LLAMA_TINY if args.model == "tiny" else LLAMA3_8B
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(model, args.steps, args.warmup, args.cpu_prompt, args.cpu_gen)
    line = {"impl": "reference", "metric": "decode_tok_s", "value": r["decode_tok_s"], "unit": "tok/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, model, 1),
            "prefill": {"value": r["prefill_tok_s"], "unit": "tok/s"},
            "cpu_baseline": {"value": r["decode_tok_s"], "unit": "tok/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["decode_tok_s"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, model, tp):
    return {"workload": "Llama-3-8B bf16, batch=1, %d-in/%d-out greedy decode (BASELINE.json configs[1])" % (PROMPT, GEN)
            if args.model == "8b" else "tiny-geometry debug run (NOT the benchmark config)",
            "model_geometry": model, "prompt_tokens": PROMPT, "generated_tokens": GEN, "batch": args.batch,
            "parallelism": "tp%d" % tp, "weights": "synthetic seed 0, N(0,0.02^2) via integer hash",
            "l2": "inputs larger than L2: 15 GB of weights are streamed per decode step (126 MB L2)"}


# ------------------------------------------------------------------------------ GPU arm -----
def gemv_microbench(model, hbm_peak):
    """Dominant decode kernel (gate/up GEMV, fused RMSNorm + SiLU*mul): CUDA events on the launch
    stream, rotating over > L2 worth of distinct weight copies."""
    import torch
    from llmlb_b200 import ffi
    L = ffi.lib()
    H, F = model["hidden"], model["ffn"]
    n_out = 2 * F
    copies = max(2, int(400e6 // (n_out * H * 2)) + 1)
    w = [torch.empty(n_out, H, dtype=torch.bfloat16, device="cuda").normal_(0, 0.02) for _ in range(copies)]
    x = torch.randn(1, H, device="cuda")
    g = torch.ones(H, dtype=torch.bfloat16, device="cuda")
    out = torch.empty(1, F, dtype=torch.bfloat16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: C.c_void_p(t.data_ptr())

    def run(i):
        ffi.check(L.llmlb_op_gemv(vp(w[i % copies]), vp(x), vp(g), 1e-5, vp(out), 1, n_out, H, ffi.EPI_SILU_MUL, F, st))
    for i in range(6):
        run(i)
    torch.cuda.synchronize()
    iters = 40
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        run(i)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / iters
    bytes_ = n_out * H * 2
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this kernel (one launch) from the
    # committed `ncu --set full` capture profiles/r1_final_gemv_ks_ncu_full.txt (241.20 + 3.24 MB)
    traffic = 244.44e6 if (n_out, H) == (28672, 4096) else None
    return {"kernel": "gemv_ks_kernel<1,SILU_MUL,NORM,1> %dx%d (gate/up, fused RMSNorm + SiLU*up)" % (n_out, H),
            "bytes_per_launch": bytes_, "us_per_launch": us, "achieved": bytes_ / us / 1e3, "unit": "GB/s",
            "frac": bytes_ / us / 1e3 / hbm_peak, "traffic": traffic}


def run_ours(args):
    import torch
    from llmlb_b200 import build, ffi
    build.build()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    tp = world
    dist = None
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation: keep fd 1 clean for
        # the ONE JSON line by pointing it at stderr until the wiring collectives are done
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = ffi.LLAMA_TINY if args.model == "tiny" else ffi.LLAMA3_8B
    hbm_peak, tf_burst, tf_sust, peak_src = measured_peaks()
    eng = ffi.Engine(model, model_id="llama-3-8b-synthetic", device=local, tp_rank=rank, tp_size=tp,
                     max_seqs=max(4, args.batch), max_ctx=1024, seed=0, gemm_impl=args.gemm_impl,
                     use_cuda_graphs=not args.no_graphs)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, eng.tp_export())
        eng.tp_import(handles)
        dist.barrier()
        t = torch.zeros(1, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()  # create the NCCL communicator now
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)

    def one_step(i):
        """One request per stream in the batch; returns per-request event lists."""
        if world > 1:
            eng.pause(True)
        rids = [eng.submit(make_prompt(i * args.batch + j, model["vocab"]), GEN, ignore_eos=True) for j in range(args.batch)]
        if world > 1:
            dist.barrier()
            eng.pause(False)
        evs = {r: [] for r in rids}
        pending = set(rids)
        while pending:
            for r in list(pending):
                got = eng.poll(r, cap=256, timeout_ms=-1 if len(pending) == 1 else 2)
                evs[r].extend(got)
                if got and got[-1]["finish_reason"]:
                    pending.discard(r)
        for r in rids:
            eng.release(r)
        return [evs[r] for r in rids]

    for i in range(args.warmup):
        one_step(10_000 + i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    h0 = eng.health()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    first_to_last_ms, n_dec_tokens, req_tps = 0.0, 0, []
    for i in range(args.steps):
        lo, hi = [], []
        for ev in one_step(i):
            toks = [e for e in ev if e["token_id"] >= 0]
            assert len(toks) == GEN, "request produced %d tokens" % len(toks)
            lo.append(toks[0]["t_ms"]); hi.append(toks[-1]["t_ms"])
            n_dec_tokens += len(toks) - 1
            req_tps.append(len(toks) / (toks[-1]["t_ms"] / 1e3))
        # all streams of a step are submitted together: decode phase = first first-token .. last last-token
        first_to_last_ms += max(hi) - min(lo)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    h1 = eng.health()
    d = lambda k: h1[k] - h0[k]
    gpu_ms_dec, gpu_ms_pre = d("gpu_ms_decode"), d("gpu_ms_prefill")
    vals = torch.tensor([gpu_ms_dec, gpu_ms_pre, wall * 1e3, first_to_last_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.barrier()
    gpu_ms_dec, gpu_ms_pre, wall_ms, ftl_ms = [float(v) for v in vals]
    dec_tokens, pre_tokens = d("tokens_decode"), d("tokens_prefill")
    dec_steps = d("steps_decode")
    decode_tok_s = dec_tokens / (gpu_ms_dec / 1e3)
    prefill_tok_s = pre_tokens / (gpu_ms_pre / 1e3)
    e2e_decode_tok_s = n_dec_tokens / (ftl_ms / 1e3) * 1.0
    if rank == 0:
        ctx = PROMPT + GEN / 2
        bytes_step = algorithmic_bytes_per_decode_step(model, args.batch, ctx, tp)
        step_s = gpu_ms_dec / 1e3 / max(1, dec_steps)
        achieved = bytes_step / step_s / 1e9
        kern = gemv_microbench(model, hbm_peak) if (world == 1 and not args.no_micro) else None
        fl = prefill_flops(model, PROMPT) * (pre_tokens / PROMPT) / tp
        pre_tf = fl / (gpu_ms_pre / 1e3) / 1e12
        line = {
            "metric": "decode_tok_s", "value": decode_tok_s, "unit": "tok/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args, model, tp),
            "prefill": {"value": prefill_tok_s, "unit": "tok/s",
                        "roofline": {"bound": "tensor", "achieved": pre_tf, "peak": tf_burst, "unit": "TFLOP/s",
                                     "frac": pre_tf / tf_burst, "peak_source": peak_src + " (burst cuBLAS bf16)"}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         # per decode step: 32 x (QKV 50.39 + O 33.60 + gate/up 244.44 + down 120.83 MB) of DRAM traffic in the
                         # ncu captures under profiles/ (lm_head not captured: its 1050.7 MB algorithmic) vs 15.08 GB algorithmic
                         "traffic": (32 * (50.39e6 + 33.60e6 + 244.44e6 + 120.83e6) + 1050.7e6) if (args.model == "8b" and tp == 1 and args.batch == 1) else None,
                         "peak_source": peak_src,
                         "what": "whole decode step (%.2f GB algorithmic per step per GPU / %.3f ms CUDA-event step time)" % (bytes_step / 1e9, step_s * 1e3),
                         "kernel": kern},
            "e2e": {"value": e2e_decode_tok_s, "unit": "tok/s", "h2d_bytes_per_step": PROMPT * 4 * args.batch,
                    "d2h_bytes_per_step": GEN * 4 * args.batch,
                    "what": "client view through llmlb_request_submit/poll with host buffers: decoded tokens / (first token event -> last token event)",
                    "request_tps_reference_formula": sum(req_tps) / len(req_tps)},
            "gpu_launches": int(d("kernel_launches")), "clocks": clocks,
        }
        if args.cpu_baseline and world == 1:
            eng.close()
            r = cpu_reference_run(model, 1, 0, args.cpu_prompt, args.cpu_gen)
            line["cpu_baseline"] = {"value": r["decode_tok_s"], "unit": "tok/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"], "prefill_tok_s": r["prefill_tok_s"]}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="8b", choices=["8b", "tiny"])
    ap.add_argument("--batch", type=int, default=1, help="concurrent streams per step (1 = BASELINE configs[1])")
    ap.add_argument("--gemm-impl", type=int, default=0)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--cpu-prompt", type=int, default=32)
    ap.add_argument("--cpu-gen", type=int, default=8)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
