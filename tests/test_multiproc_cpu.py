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
