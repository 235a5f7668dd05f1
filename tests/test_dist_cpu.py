"""N > 1 path on CPU: two gloo ranks shard a batch, compute the SS2D parameter gradients of their shard (CPU oracle as the
compute stand-in -- tests only), reduce them with the product's bucket all-reduce, and must reproduce the gradients of the
whole batch.  Also checks shard_range balance/coverage."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from conftest import rel_err
from xfmamba_b200.dp import FlatBucket, allreduce_param_grads, shard_range


def test_shard_range_partitions_the_batch():
    for gb in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            parts = [shard_range(gb, r, world) for r in range(world)]
            assert [i for p in parts for i in p] == list(range(gb))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _inputs():
    rng = np.random.default_rng(0)
    Bsz, D, N, H, W = 6, 3, 2, 5, 4
    L = H * W
    f = lambda *s: rng.standard_normal(s, dtype=np.float32)
    return dict(x=f(Bsz, D, H, W), delta=0.5 * rng.random((Bsz, 4 * D, L), dtype=np.float32), A=-0.5 * rng.random((4 * D, N), dtype=np.float32),
                Bs=f(Bsz, 4, N, L), Cs=f(Bsz, 4, N, L), Ds=f(4 * D), delta_bias=0.5 * rng.random(4 * D, dtype=np.float32), dy=f(Bsz, D, L))


def _param_grads(c, rows):
    sl = slice(rows.start, rows.stop)
    g = oracle.ss2d_bwd(c["x"][sl], c["delta"][sl], c["A"], c["Bs"][sl], c["Cs"][sl], c["Ds"], c["delta_bias"], c["dy"][sl], True, "f64")
    return [torch.from_numpy(np.ascontiguousarray(g[i])) for i in (2, 5, 6)]      # dA, dDs, ddelta_bias


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = _inputs()
    grads = _param_grads(c, shard_range(c["x"].shape[0], rank, world))
    allreduce_param_grads(grads)                      # one flat all-reduce (sum)
    b = FlatBucket(grads)                             # mean variant through the bucket API
    b.pack([torch.ones_like(t) * (rank + 1) for t in grads])
    means = [v.clone() for v in b.allreduce(average=True)]
    if rank == 0:
        q.put(([t.numpy() for t in grads], [float(m.mean()) for m in means]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_full_batch():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, means = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = _inputs()
    full = _param_grads(c, range(c["x"].shape[0]))
    for a, b in zip(got, full):
        assert rel_err(a, b.numpy()) < 1e-12
    assert means == [1.5, 1.5, 1.5]
