"""N>1 path on CPU (world_size 2, gloo): the tensor-parallel sharding the engine uses
(csrc/engine.cu gen_weights / resolve_tensor) restated on the oracle — q/k/v/gate/up by output
rows, o/down by input columns, lm_head by vocab, every shard generated from GLOBAL indices — and
checked across two real processes: partial products all-reduced over gloo equal the unsharded
layer, and vocab shards all-gathered equal the full logits."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.synth import KIND, bf16_bits_to_f32, synth_bits

CFG = dict(hidden=256, n_heads=4, n_kv_heads=2, head_dim=128, ffn=512, vocab=1024)


def _shard(seed, kind, rows, cols, rank, tp, by):
    """The slice rank `rank` holds: by='rows' -> row block, by='cols' -> column block."""
    if by == "rows":
        r = rows // tp
        return bf16_bits_to_f32(synth_bits(seed, kind, r, cols, row0=rank * r, col0=0, ld=cols))
    c = cols // tp
    return bf16_bits_to_f32(synth_bits(seed, kind, rows, c, row0=0, col0=rank * c, ld=cols))


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, F, V = CFG["hidden"], CFG["ffn"], CFG["vocab"]
    nq, hd = CFG["n_heads"], CFG["head_dim"]
    rs = np.random.RandomState(0)
    y = rs.randn(3, H).astype(np.float32)          # normalised activations (replicated)
    # column-parallel: gate/up rows of this rank -> h slice; row-parallel: down columns -> partial
    wg = _shard(0, KIND["mlp.gate_proj.weight"], F, H, rank, world, "rows")
    wu = _shard(0, KIND["mlp.up_proj.weight"], F, H, rank, world, "rows")
    g, u = y @ wg.T, y @ wu.T
    h = (g / (1 + np.exp(-g))) * u                 # [3, F/tp]
    wd = _shard(0, KIND["mlp.down_proj.weight"], H, F, rank, world, "cols")
    part = torch.from_numpy((h @ wd.T).astype(np.float32))
    dist.all_reduce(part)                           # the exchange step (sum over ranks)
    # attention output projection: columns of o_proj follow the head shard
    a = rs.randn(3, nq * hd).astype(np.float32)
    wo = _shard(0, KIND["self_attn.o_proj.weight"], H, nq * hd, rank, world, "cols")
    c = nq * hd // world
    po = torch.from_numpy((a[:, rank * c:(rank + 1) * c] @ wo.T).astype(np.float32))
    dist.all_reduce(po)
    # vocab-sharded logits, gathered
    wl = _shard(0, 0xFFFF0002, V, H, rank, world, "rows")
    lg = torch.from_numpy((y @ wl.T).astype(np.float32))
    parts = [torch.empty_like(lg) for _ in range(world)]
    dist.all_gather(parts, lg)
    if rank == 0:
        out_q.put((part.numpy(), po.numpy(), torch.cat(parts, dim=1).numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_unsharded():
    H, F, V = CFG["hidden"], CFG["ffn"], CFG["vocab"]
    nq, hd = CFG["n_heads"], CFG["head_dim"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    part, po, logits = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    rs = np.random.RandomState(0)
    y = rs.randn(3, H).astype(np.float32)
    full = lambda kind, r, c: bf16_bits_to_f32(synth_bits(0, kind, r, c))
    g = y @ full(KIND["mlp.gate_proj.weight"], F, H).T
    u = y @ full(KIND["mlp.up_proj.weight"], F, H).T
    want = ((g / (1 + np.exp(-g))) * u) @ full(KIND["mlp.down_proj.weight"], H, F).T
    assert np.allclose(part, want, rtol=1e-4, atol=1e-5)
    a = rs.randn(3, nq * hd).astype(np.float32)
    assert np.allclose(po, a @ full(KIND["self_attn.o_proj.weight"], H, nq * hd).T, rtol=1e-4, atol=1e-5)
    assert np.allclose(logits, y @ full(0xFFFF0002, V, H).T, rtol=1e-4, atol=1e-5)
