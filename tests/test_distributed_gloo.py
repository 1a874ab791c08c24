"""world_size-2 gloo tests (CPU) of the N>1 host logic: shard placement of the single latent
all-gather, the (lse,pos) exchange and the global-loss assembly the ranks derive from it."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clip_oracle as O


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from x_clip_b200 import distributed as D
    assert D.world() == (rank, world)
    b, dim = 3, 8
    g = torch.Generator().manual_seed(0)
    full = [torch.nn.functional.normalize(torch.randn(world * b, dim, generator=g), dim=-1) for _ in range(4)]
    shards = [f[rank * b:(rank + 1) * b].clone() for f in full]
    got = D.gather_rows(shards)
    ok = all(torch.equal(a, f) for a, f in zip(got, full))

    # each rank evaluates only its row block; the gathered (lse,pos) give the reference's global loss
    temp = torch.tensor(1.0).exp()
    zt, zi = got[0], got[1]
    s_rows = temp * shards[0] @ zi.t()                       # local texts vs all images
    s_cols = temp * shards[1] @ zt.t()                       # local images vs all texts
    idx = torch.arange(b) + rank * b
    stats = torch.stack([torch.logsumexp(s_rows, -1), s_rows[torch.arange(b), idx],
                         torch.logsumexp(s_cols, -1), s_cols[torch.arange(b), idx]])
    gs = D.gather_stats(stats)
    loss = ((gs[0] - gs[1]).sum() + (gs[2] - gs[3]).sum()) / (2 * world * b)
    cfg = O.ClipConfig()
    ref = O.contrastive_loss(full[0], full[1], full[0], full[1], torch.tensor(1.0), cfg)
    t = torch.ones(1) * (rank + 1)
    D.all_reduce_sum_(t)
    q.put((rank, ok, abs(loss.item() - ref.item()), t.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_loss_assembly():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29733, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    import queue as _q, time as _t
    t0 = _t.time()
    while len(res) < world and _t.time() - t0 < 240:
        try:
            res.append(q.get(timeout=2))
        except _q.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=30)
        if p.is_alive():
            p.terminate()
    assert len(res) == world, "a rank died (see its traceback above)"
    res.sort()
    for rank, ok, err, red in res:
        assert ok, f"rank {rank}: gathered rows are not in rank order"
        assert err < 1e-6, f"rank {rank}: assembled loss differs from the full-batch loss by {err}"
        assert red == 3.0


def test_single_process_is_passthrough():
    from x_clip_b200 import distributed as D
    assert D.world() == (0, 1)
    a = torch.randn(4, 8)
    assert torch.equal(D.gather_rows([a])[0], a)
    assert D.gather_stats(a) is a


def _rs_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from x_clip_b200 import distributed as D
    t = torch.arange(world * 3 * 4, dtype=torch.float32).view(world * 3, 4) * (rank + 1)
    out = D.reduce_scatter_rows(t)
    full = torch.arange(world * 3 * 4, dtype=torch.float32).view(world * 3, 4) * sum(range(1, world + 1))
    q.put((rank, torch.equal(out, full[rank * 3:(rank + 1) * 3])))
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_scatter_rows_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rs_worker, args=(r, world, 29735, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(ok for _, ok in res)


def _gs_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from x_clip_b200 import distributed as D
    torch.manual_seed(0)                                      # identical weights on every rank
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.GELU(), torch.nn.Linear(16, 16),
                              torch.nn.GELU(), torch.nn.Linear(16, 3))
    frozen = torch.nn.Linear(3, 3)                            # never receives a gradient
    mod = torch.nn.ModuleList([net, frozen])
    import copy
    ref_net = copy.deepcopy(net)                              # hook-free twin for the expected values
    sync = D.GradSync(mod, bucket_bytes=600)                  # several small buckets
    nb = len(sync._buckets)
    errs = []
    for step in range(2):                                     # re-arming across steps
        data = [torch.randn(5, 6, generator=torch.Generator().manual_seed(10 * step + r))
                for r in range(world)]
        mod.zero_grad(set_to_none=True)
        net(data[rank]).square().sum().backward()
        sync.finish()
        got = [p.grad.clone() for p in net.parameters()]
        want = [torch.zeros_like(p) for p in net.parameters()]
        for r in range(world):                                # single-process reference: mean over ranks
            ref_net.zero_grad(set_to_none=True)
            ref_net(data[r]).square().sum().backward()
            for w, p in zip(want, ref_net.parameters()):
                w += p.grad / world
        errs.append(max((g - w).abs().max().item() for g, w in zip(got, want)))
    unused_ok = all(p.grad is None for p in frozen.parameters())
    # gradient accumulation: the first pass stays local (no_sync), the second one reduces the sum
    mod.zero_grad(set_to_none=True)
    d0 = [torch.randn(5, 6, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    d1 = [torch.randn(5, 6, generator=torch.Generator().manual_seed(200 + r)) for r in range(world)]
    with sync.no_sync():
        net(d0[rank]).square().sum().backward()
    net(d1[rank]).square().sum().backward()
    sync.finish()
    got = [p.grad.clone() for p in net.parameters()]
    want = [torch.zeros_like(p) for p in net.parameters()]
    for r in range(world):
        ref_net.zero_grad(set_to_none=True)
        (ref_net(d0[r]).square().sum() + ref_net(d1[r]).square().sum()).backward()
        for w, p in zip(want, ref_net.parameters()):
            w += p.grad / world
    errs.append(max((g - w).abs().max().item() for g, w in zip(got, want)))

    # micro-batched backward (engine.ChunkedClipLossFn pattern): one autograd pass per chunk INSIDE
    # a backward; every pass but the last defers the hooks, so each bucket is reduced once, with
    # the gradient accumulated over all chunks (first-chunk-only reduction)
    class Chunked(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, defer):
            ctx.x, ctx.defer = x, defer
            with torch.no_grad():
                return sum(net(c).square().sum() for c in x.chunk(3))

        @staticmethod
        def backward(ctx, g):
            chunks = ctx.x.chunk(3)
            for k, c in enumerate(chunks):
                with torch.enable_grad(), D.defer_grad_sync(ctx.defer and k != len(chunks) - 1):
                    torch.autograd.backward(net(c).square().sum(), g)
            return None, None

    dch = [torch.randn(6, 6, generator=torch.Generator().manual_seed(300 + r)) for r in range(world)]
    mod.zero_grad(set_to_none=True)
    Chunked.apply(dch[rank].requires_grad_(True), True).backward()
    sync.finish()
    got = [p.grad.clone() for p in net.parameters()]
    want = [torch.zeros_like(p) for p in net.parameters()]
    for r in range(world):
        ref_net.zero_grad(set_to_none=True)
        ref_net(dch[r].detach()).square().sum().backward()
        for w, p in zip(want, ref_net.parameters()):
            w += p.grad / world
    errs.append(max((g - w).abs().max().item() for g, w in zip(got, want)))
    # the same step with in-place gradient accumulation (engine._accumulate_into_grad): the middle
    # chunk adds its parameter gradients straight into `.grad` and hands autograd nothing, the hooks
    # fire only in the last chunk - and must then see (and reduce) the sum over ALL chunks
    class ChunkedInPlace(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            ctx.x = x
            with torch.no_grad():
                return sum(net(c).square().sum() for c in x.chunk(3))

        @staticmethod
        def backward(ctx, g):
            chunks = ctx.x.chunk(3)
            params = [p for p in net.parameters()]
            for k, c in enumerate(chunks):
                with torch.enable_grad(), D.defer_grad_sync(k != len(chunks) - 1):
                    loss = net(c).square().sum()
                    if k == 1:                                # "kernels accumulate into .grad, return None"
                        gs = torch.autograd.grad(loss, params, g)
                        for p_, g_ in zip(params, gs):
                            p_.grad.add_(g_)
                    else:
                        torch.autograd.backward(loss, g)
            return None

    mod.zero_grad(set_to_none=True)
    ChunkedInPlace.apply(dch[rank].requires_grad_(True)).backward()
    sync.finish()
    got = [p.grad.clone() for p in net.parameters()]
    errs.append(max((g - w).abs().max().item() for g, w in zip(got, want)))
    # without the deferral the second arrival of a parameter is an error, not a silent wrong result
    mod.zero_grad(set_to_none=True)
    raised = False
    try:
        Chunked.apply(dch[rank].requires_grad_(True), False).backward()
    except RuntimeError as e:
        raised = "second gradient" in str(e)
    for b in sync._buckets:                       # drain whatever the failed pass launched
        if b["work"] is not None:
            b["work"].wait()
        b["arrived"], b["work"], b["flat"], b["members"] = 0, None, None, None
    # unequal local batches are reported, not hung on
    uneq = False
    try:
        D.assert_equal_local_batch(3 + rank, torch.device("cpu"))
    except RuntimeError:
        uneq = True
    D.assert_equal_local_batch(3, torch.device("cpu"))
    q.put((rank, nb, max(errs), unused_ok and raised and uneq))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_sync_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gs_worker, args=(r, world, 29735, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    import queue as _q, time as _t
    t0 = _t.time()
    while len(res) < world and _t.time() - t0 < 240:
        try:
            res.append(q.get(timeout=2))
        except _q.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=30)
        if p.is_alive():
            p.terminate()
    assert len(res) == world, "a rank died (see its traceback above)"
    for rank, nb, err, unused_ok in res:
        assert nb >= 3, "the test is meant to exercise several buckets"
        assert err < 1e-6, f"rank {rank}: synced gradients differ from the mean of the ranks' gradients by {err}"
        assert unused_ok


def test_grad_sync_single_rank_is_noop():
    from x_clip_b200 import distributed as D
    lin = torch.nn.Linear(4, 4)
    sync = D.GradSync(lin)
    lin(torch.randn(2, 4)).sum().backward()
    g = lin.weight.grad.clone()
    sync.finish()
    assert torch.equal(lin.weight.grad, g)
