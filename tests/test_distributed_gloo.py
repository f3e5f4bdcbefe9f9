"""CPU tests (gloo, world_size 2) of the multi-GPU host logic in flowconductor_b200/distributed.py: row
sharding, the log-likelihood reduction and the flat-bucket gradient all-reduce.  The per-rank compute is the
oracle (CPU), because the product has no CPU path; what is under test is the partitioning / reduction logic."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flowconductor_b200 import distributed as fdist
from flowconductor_b200 import workloads
from tests.helpers import ROOT, golden_state, load_golden


def test_shard_bounds_cover_rows_exactly():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [fdist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        fdist.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from oracle import restated

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gold = load_golden("cfg2_small")
        wl = workloads.get_workload("cfg2_small")
        specs = workloads.oracle_specs(wl)
        state = golden_state(gold)
        x = gold["x"]
        # (1) sharded log-likelihood == single-process log-likelihood
        mine = fdist.shard_rows(x)
        with torch.no_grad():
            lp = restated.flow_log_prob(state, specs, mine)
        total, count = fdist.reduce_log_likelihood(lp)
        # (2) gradient all-reduce: average of per-rank gradients, identical on every rank
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Tanh(), torch.nn.Linear(3, 1))
        fdist.broadcast_parameters(net)
        data = torch.arange(32, dtype=torch.float32).reshape(8, 4) / 10
        shard = fdist.shard_rows(data)
        net(shard).pow(2).mean().backward()
        local = [p.grad.clone() for p in net.parameters()]
        n = fdist.allreduce_gradients(net)
        reduced = [p.grad.clone() for p in net.parameters()]
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        if rank == 0:
            torch.save({"total": total, "count": count, "reduced": reduced, "locals": gathered, "n": n}, out)
    finally:
        dist.destroy_process_group()


def test_sharded_log_prob_and_gradient_allreduce(tmp_path):
    from oracle import restated

    world = 2
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = torch.load(out, weights_only=False)
    gold = load_golden("cfg2_small")
    wl = workloads.get_workload("cfg2_small")
    with torch.no_grad():
        full = restated.flow_log_prob(golden_state(gold), workloads.oracle_specs(wl), gold["x"])
    assert res["count"] == gold["x"].shape[0]
    assert abs(float(res["total"]) - float(full.double().sum())) < 1e-6 * abs(float(full.double().sum()))
    for i, red in enumerate(res["reduced"]):
        expect = sum(loc[i] for loc in res["locals"]) / world
        assert torch.allclose(red, expect, atol=1e-7)
    assert res["n"] == sum(t.numel() for t in res["reduced"])


def test_single_process_helpers_without_process_group():
    lp = torch.tensor([1.0, 2.0, 3.5])
    total, count = fdist.reduce_log_likelihood(lp)
    assert float(total) == 6.5 and count == 3
    net = torch.nn.Linear(2, 2)
    net(torch.ones(1, 2)).sum().backward()
    assert fdist.allreduce_gradients(net) == 6
    fdist.broadcast_parameters(net)  # no-op without a process group


def test_chunked_log_prob_matches_unchunked():
    class Fake:
        def log_prob(self, inputs, context=None):
            return inputs.sum(1) + (0 if context is None else context.sum(1))

    x = torch.randn(10, 3)
    c = torch.randn(10, 2)
    a = fdist.sharded_log_prob(Fake(), x, c)
    b = fdist.sharded_log_prob(Fake(), x, c, chunk_rows=4)
    assert torch.equal(a, b)
