"""Tensor-parallel partition: host logic on CPU, including a world_size-2 gloo run that checks the
sharded o_proj / down_proj arithmetic (partials summed in rank order) against the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import binding as orc
from oracle import np_forward as npf
from qwen3_rs_b200.tp_plan import prefill_exchange_slices, shard_plan


def test_plan_covers_everything_once():
    cfg = dict(n_heads=32, n_kv_heads=8, head_dim=128, hidden_dim=12288, vocab_size=151936, group_size=64)
    for tp in (1, 2, 4, 8):
        plans = [shard_plan(cfg, r, tp) for r in range(tp)]
        for attr, total in (("q_rows", 4096), ("kv_rows", 1024), ("hidden_rows", 12288), ("vocab_rows", 151936)):
            got = sorted(i for p in plans for i in getattr(p, attr))
            assert got == list(range(total))
        for p in plans:  # GQA groups intact, quantisation groups never straddle a shard
            assert len(p.q_rows) // len(p.kv_rows) == 4
            assert p.hidden_rows.start % 64 == 0 and p.attn_cols.start % 128 == 0


def test_prefill_exchange_slices_tile_the_block_and_give_the_direct_sum():
    """Reduce-scatter + all-gather form of the prefill exchange: the slices tile the block, and summing slice r on rank r in rank
    order then gathering gives bit for bit what every rank would get by adding all partial blocks itself (the tp <= 2 form)."""
    rng = np.random.default_rng(0)
    for tp, n4 in ((3, 10), (4, 37 * 640 // 4), (8, 2048 * 4096 // 4 // 64), (8, 7)):
        sl = prefill_exchange_slices(n4, tp)
        assert sorted(i for s_ in sl for i in s_) == list(range(n4))
        parts = rng.standard_normal((tp, n4, 4)).astype(np.float32)
        direct = parts[0].copy()
        for r in range(1, tp):
            direct = (direct + parts[r]).astype(np.float32)
        gathered = np.empty_like(direct)
        for r, s_ in enumerate(sl):  # rank r reduces its slice ...
            acc = parts[0][s_.start:s_.stop].copy()
            for q in range(1, tp):
                acc = (acc + parts[q][s_.start:s_.stop]).astype(np.float32)
            gathered[s_.start:s_.stop] = acc  # ... and everybody gathers it
        assert np.array_equal(gathered, direct)


def test_plan_rejects_bad_splits():
    cfg = dict(n_heads=32, n_kv_heads=8, head_dim=128, hidden_dim=9728, vocab_size=151936, group_size=128)
    shard_plan(cfg, 0, 4)
    with pytest.raises(ValueError):
        shard_plan(cfg, 0, 8)  # 4B gs128: 9728/8 = 1216 is not a multiple of 128
    with pytest.raises(ValueError):
        shard_plan(dict(cfg, n_kv_heads=8), 0, 3)
    with pytest.raises(ValueError):
        shard_plan(cfg, 4, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = npf.NpModel(path)
        cfg = dict(n_heads=m.n_heads, n_kv_heads=m.n_kv, head_dim=m.hd, hidden_dim=m.hidden, vocab_size=m.vocab, group_size=m.gs)
        plan = shard_plan(cfg, rank, world)
        gs, dim = m.gs, m.dim
        rng = np.random.default_rng(0)  # same activations on every rank (x is replicated under TP)
        att = rng.standard_normal(m.AH).astype(np.float32)
        hb = rng.standard_normal(m.hidden).astype(np.float32)
        aq, as_ = orc.quantize(att, gs)
        hq, hs = orc.quantize(hb, gs)
        results = {}
        for name, (wq, ws), n, xq, xs, cols in (("wo", m.wo[0], m.AH, aq, as_, plan.attn_cols), ("w2", m.w2[0], m.hidden, hq, hs, plan.hidden_rows)):
            w2d, s2d = wq.reshape(dim, n), ws.reshape(dim, n // gs)
            sl, gsl = slice(cols.start, cols.stop), slice(cols.start // gs, cols.stop // gs)
            wl = np.ascontiguousarray(w2d[:, sl]).reshape(-1)
            sl_s = np.ascontiguousarray(s2d[:, gsl]).reshape(-1)
            nl = len(cols)
            part = orc.matmul(xq[sl], xs[gsl], wl, sl_s, nl, dim, gs)                 # this rank's partial sum
            dots = orc.group_dots(xq[sl], wl, nl, dim, gs)                            # its int32 group dots
            parts = [torch.zeros(dim) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(part))
            total = np.zeros(dim, np.float32)
            for p in parts:                                                         # rank order, like the kernel
                total = (total + p.numpy()).astype(np.float32)
            full = orc.matmul(xq, xs, wq, ws, n, dim, gs)
            full_dots = orc.group_dots(xq, wq, n, dim, gs)
            results[name] = (bool(np.array_equal(dots, full_dots[:, gsl])), float(np.abs(total - full).max()),
                             float(np.abs(full).max()))
        # vocab-sharded argmax: per-rank best (value, index) -> global, last max wins on ties
        logits = rng.standard_normal(m.vocab).astype(np.float32)
        logits[[3, m.vocab - 2]] = 9.0
        lo, hi = plan.vocab_rows.start, plan.vocab_rows.stop
        local = lo + orc.argmax(logits[lo:hi])
        cands = [None] * world
        dist.all_gather_object(cands, (float(logits[local]), int(local)))
        results["argmax"] = (max(cands)[1], orc.argmax(logits))
        if rank == 0:
            q.put(results)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharded_projections_match_unsharded(ckpt):
    path = ckpt("tiny-untied", 64, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for name in ("wo", "w2"):
        dots_equal, err, scale = res[name]
        assert dots_equal, f"{name}: sharded int32 group dots differ from the unsharded run"
        assert err <= 1e-5 * max(1.0, scale)
    assert res["argmax"][0] == res["argmax"][1]
