"""World-size-2 gloo test of the rank plumbing used by bench.py --gpus N: the path shards by volume batch
(one volume per rank, no data-path collective); timings are reduced with MAX over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(1234 + rank)                      # same seeding rule as bench.py
    x = torch.rand(4)
    t = torch.tensor([10.0 + rank, 20.0 - rank], dtype=torch.float64)   # per-rank timings
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [torch.empty(4) for _ in range(world)]
    dist.all_gather(gathered, x)
    # the NMF buffers must be identical on all ranks (DDP broadcasts them from rank 0)
    u0 = torch.rand(8, 1) if rank == 0 else torch.zeros(8, 1)
    dist.broadcast(u0, src=0)
    if rank == 0:
        out.put((t.tolist(), [g.tolist() for g in gathered], u0.flatten().tolist()))
    else:
        out.put(("u0", u0.flatten().tolist()))
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_reduce():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    main = next(r for r in res if r[0] != "u0")
    other = next(r for r in res if r[0] == "u0")
    assert main[0] == [11.0, 20.0]                      # MAX over ranks
    assert main[1][0] != main[1][1]                     # different volumes per rank
    assert main[2] == other[1]                          # broadcast buffers agree


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from factorizer_b200.distributed import BucketedGradAllReduce
    torch.manual_seed(7)                                # same weights on both ranks
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.GELU(), torch.nn.Linear(16, 16), torch.nn.GELU(),
                              torch.nn.Linear(16, 3))
    unused = torch.nn.Parameter(torch.ones(5))         # never reached by backward: its bucket is launched by wait()
    params = list(net.parameters()) + [unused]
    torch.manual_seed(100 + rank)                      # different volumes per rank
    x = torch.randn(9, 6)
    # the plain answer: local gradients averaged over the ranks
    net(x).square().sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    want = []
    for g in local:
        parts = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(parts, g)
        want.append(sum(parts) / world)
    for p in params:
        p.grad = None
    red = BucketedGradAllReduce(params, bucket_bytes=600)          # several buckets (the 16 x 16 weight alone exceeds one)
    nb = len(red.buckets)
    errs = []
    for _ in range(2):                                 # two steps: the state resets
        red.zero_grad()
        net(x).square().sum().backward()
        red.wait()
        errs.append(max(float((p.grad - w).abs().max()) for p, w in zip(net.parameters(), want)))
    views = all(p.grad.untyped_storage().data_ptr() == red.buckets[red._bucket_of[p]].untyped_storage().data_ptr() for p in params)
    unused_zero = bool((unused.grad == 0).all())
    red.remove()
    out.put((rank, nb, errs, views, unused_zero, all(p.grad is None for p in params)))
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_two_ranks():
    """factorizer_b200.distributed.BucketedGradAllReduce (the training config's only collective): gradients averaged over
    two ranks equal the plain all-gather average, .grad stay views of the flat buckets, buckets of unused parameters
    are still reduced (every rank launches the same collectives), remove() restores free-standing gradients."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, nb, errs, views, unused_zero, cleared in res:
        assert nb >= 3
        assert max(errs) < 1e-6
        assert views and unused_zero and cleared
