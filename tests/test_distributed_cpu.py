"""N>1 path on CPU: two gloo ranks shard independent backtests contiguously and all-gather their metric rows —
the only collective of the path.  The per-rank 'backtests' here are the oracle's (the CUDA kernels
need a GPU); what is under test is the sharding + gather logic of engine.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _metrics_of(b):
    """deterministic stand-in for one backtest's five metrics"""
    rng = np.random.default_rng(1000 + b)
    return rng.standard_normal(5)


def _worker(rank, world, port, n_total, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from koopman_mpc_portfolio_rebalancing_b200.engine import gather_metrics, shard_range
    lo, hi = shard_range(n_total, rank, world)
    local = torch.from_numpy(np.stack([_metrics_of(b) for b in range(lo, hi)]).reshape(-1, 5)) if hi > lo else torch.zeros((0, 5), dtype=torch.float64)
    full = gather_metrics(local, n_total, rank, world)
    out_q.put((rank, lo, hi, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 8, 3])
def test_two_rank_shard_and_gather(n_total):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([_metrics_of(b) for b in range(n_total)])
    covered = []
    for rank, lo, hi, full in results:
        assert full.shape == (n_total, 5)
        assert np.array_equal(full, want)
        covered += list(range(lo, hi))
    assert sorted(covered) == list(range(n_total))
