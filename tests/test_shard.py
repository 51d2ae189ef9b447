"""Multi-GPU partition logic (idash2019_2_b200/shard.py) on CPU: properties of the partition, and a world-size-2 gloo
run in which every rank evaluates its shard (with the oracle standing in for the GPU evaluator -- this test is about
the host-side sharding, gathering and ordering) and the gathered result equals the unsharded evaluation."""
import os
import socket

import numpy as np
import pytest

from idash2019_2_b200 import shard, synth
from oracle import pyoracle as po

ALPHA2 = 2.0 ** -50


def _case(S, T, G, n, seed):
    geo = synth.Geometry(S, T, G)
    tag, tgt = synth.make_positions(T, G, seed)
    model = synth.make_model(tag, tgt, n, seed)
    cts = synth.random_ciphertexts(geo.n_in_ct_used, seed)
    return geo, model, cts


@pytest.mark.parametrize("G,world", [(10, 1), (10, 3), (7, 8), (80882, 8), (5, 2), (0, 2)])
def test_target_ranges_partition(G, world):
    r = [shard.target_range(G, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == G
    assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
    sizes = [b - a for a, b in r]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.target_range(G, world, world)


@pytest.mark.parametrize("S,n,world", [(1004, 5, 2), (1004, 20, 4), (335, 5, 3), (16, 5, 2), (1004, 5, 8)])
def test_shards_cover_rows_and_slabs_cover_inputs(S, n, world):
    geo, model, cts = _case(S, 120, 300, n, seed=5)
    shards = [shard.make_shard(model, geo.NR, 300, k, world) for k in range(world)]
    assert sum(s.model.n_out for s in shards) == model.n_out
    assert np.array_equal(np.concatenate([s.model.out_bidx for s in shards]), model.out_bidx)
    for s in shards:
        real = s.model.col != shard.CONSTANT_BIDX
        assert (s.model.col[real] // geo.NR < s.n_ct).all()           # every entry lies inside the slab
        g = model.rows(s.row_lo, s.row_hi)
        assert np.array_equal(s.model.col[real] + np.uint32(s.ct_min * geo.NR), g.col[real])   # regions are kept
        assert np.array_equal(s.model.coef, g.coef)
    # slabs are banded: they advance with the rank and overlap only by a halo bounded by the window width
    assert all(a.ct_min <= b.ct_min for a, b in zip(shards, shards[1:]))
    assert max(shard.halo(shards)) <= (3 * n) // geo.NR + 3
    # per-GPU input bytes scale as 1/P plus the halo
    assert max(s.n_ct for s in shards) <= geo.n_in_ct_used // world + 3 * n + 3 * ((120 // world) // 4 + 4)


def test_shard_of_constant_only_rows():
    m = synth.CsrModel(np.arange(6, dtype=np.uint32), np.arange(7, dtype=np.uint64), np.full(6, 0xFFFFFFFF, np.uint32),
                       np.arange(6, dtype=np.int32))
    s = shard.make_shard(m, 1, 2, 1, 2)
    assert s.n_ct == 0 and s.model.n_out == 3 and (s.model.col == 0xFFFFFFFF).all()
    with pytest.raises(ValueError):
        shard.make_shard(m, 1, 3, 0, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, S, T, G, n, seed, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        geo, model, cts = _case(S, T, G, n, seed)
        sh = shard.make_shard(model, geo.NR, G, rank, world)
        slab = np.ascontiguousarray(sh.slab(cts))
        var = np.full(len(slab), ALPHA2)
        out, ovar = po.cloud_port(S, geo.NR, geo.RS, np.arange(len(slab), dtype=np.uint32), slab, var, sh.model.row_ptr,
                                  sh.model.col, sh.model.coef)
        # gather (the only communication of the path): ragged row counts -> pad to the largest shard
        n_max = max(shard.target_range(G, k, world)[1] - shard.target_range(G, k, world)[0] for k in range(world)) * 3
        pad = np.zeros((n_max, 2048), np.int64)
        pad[:len(out)] = out
        bufs = [torch.zeros((n_max, 2048), dtype=torch.int64) for _ in range(world)]
        dist.all_gather(bufs, torch.from_numpy(pad))
        t = torch.tensor([float(len(out))], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                      # the timing reduction of bench.py, on gloo
        if rank == 0:
            rows = [shard.target_range(G, k, world) for k in range(world)]
            full = np.concatenate([bufs[k].numpy()[:3 * (hi - lo)] for k, (lo, hi) in enumerate(rows)]).astype(np.uint32)
            ref, _ = po.cloud_port(S, geo.NR, geo.RS, np.arange(len(cts), dtype=np.uint32), cts, np.full(len(cts), ALPHA2),
                                   model.row_ptr, model.col, model.coef)
            q.put((bool(np.array_equal(full, ref)), int(t.item()), model.n_out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("S,T,G,n", [(1004, 40, 90, 5), (335, 30, 61, 4)])
def test_two_rank_gloo_sharded_evaluation_equals_unsharded(S, T, G, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, T, G, n, 9, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, total, n_out = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and total == n_out
