"""World-size-2 gloo test of the data-parallel plumbing: the path shards by image,
the only collective is the flat gradient all-reduce (mean) of dg.allreduce_gradients, issued with
parameter subsets that differ by training mode but are identical across ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dgod_b200.dg import allreduce_gradients
        torch.manual_seed(0)                                   # identical replicas
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
        params = list(net.parameters())
        g = torch.Generator().manual_seed(100 + rank)          # each rank owns different images
        x = torch.randn(4, 6, generator=g)
        # "mode" step: the last layer is unused -> its parameters have no gradient on any rank
        loss = net[2](net[1](net[0](x))).square().mean()
        loss.backward()
        assert params[-1].grad is None
        local = [p.grad.clone() if p.grad is not None else None for p in params]
        allreduce_gradients(params, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        for i, p in enumerate(params):
            if p.grad is None:
                assert all(gl[i] is None for gl in gathered)
                continue
            mean = sum(gl[i] for gl in gathered) / world
            torch.testing.assert_close(p.grad, mean, rtol=1e-6, atol=1e-7)
        # whole-job throughput accounting of bench.py: max over ranks of the per-rank time
        t = torch.tensor([10.0 + rank])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == 10.0 + world - 1
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() == "ok"


def _sync_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dgod_b200.ddp import GradSync
        torch.manual_seed(0)                                   # identical replicas
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2),
                                  torch.nn.Linear(3, 2))
        params = list(net.parameters())
        sync = GradSync(params, world, bucket_bytes=64)        # tiny buckets: several all-reduces per step
        assert len(sync.bucket_ranges) >= 3 and sum(len(m) for m in sync.members) == len(params)
        g = torch.Generator().manual_seed(100 + rank)          # each rank owns different images

        def local_grads(loss_fn, x):
            ref = [torch.nn.Parameter(p.detach().clone()) for p in params]
            w = dict(zip(["0.weight", "0.bias", "2.weight", "2.bias", "3.weight", "3.bias", "4.weight", "4.bias"], ref))
            loss_fn(lambda t, i: torch.nn.functional.linear(t, w[f"{i}.weight"], w[f"{i}.bias"]), x).backward()
            return [p.grad for p in ref]

        def fwd_a(lin, x):     # "mode A": head 4 unused on every rank
            return lin(lin(lin(x, 0).relu(), 2), 3).square().mean()

        def fwd_b(lin, x):     # "mode B": which head is used depends on the rank's data (per-image loops of modes 2-4)
            h = lin(lin(x, 0).relu(), 2)
            return lin(h, 3 if rank == 0 else 4).square().mean()

        module_lin = lambda t, i: net[i](t)
        for step, (key, fn, consistent) in enumerate([("A", fwd_a, True), ("A", fwd_a, True), ("B", fwd_b, False), ("A", fwd_a, True)]):
            x = torch.randn(4, 6, generator=g)
            want_local = local_grads(fn, x)
            loss = fn(module_lin, x)
            sync.begin(key)
            loss.backward()
            sync.finish(consistent_across_ranks=consistent)
            gathered = [None] * world
            dist.all_gather_object(gathered, want_local)
            for i, p in enumerate(params):
                touched_anywhere = any(gl[i] is not None for gl in gathered)
                if not touched_anywhere:
                    assert p.grad is None, (step, i)           # the optimizer skips it, as for the reference
                    continue
                mean = sum(gl[i] if gl[i] is not None else torch.zeros_like(p) for gl in gathered) / world
                torch.testing.assert_close(p.grad, mean, rtol=1e-6, atol=1e-7)
            assert all(p.grad is None or p.grad is v for p, v in zip(params, sync.views))      # no copy back: .grad IS the view
            torch.optim.SGD(params, lr=0.1, weight_decay=0.1).step()
        assert sync._expected["A"] == [c for c in sync._expected["A"]] and sum(sync._expected["A"]) == 6
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_grad_sync_buckets_world_size_2():
    """dgod_b200.ddp.GradSync: gradients as views of one flat buffer, bucketed all-reduce from backward hooks, parameters
    untouched on every rank hidden from the optimizer, rank-dependent touched sets reduced correctly."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get() == "ok"


def test_single_rank_is_a_no_op():
    from dgod_b200.dg import allreduce_gradients
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    allreduce_gradients([p], 1)
    assert torch.equal(p.grad, torch.full((3,), 2.0))
