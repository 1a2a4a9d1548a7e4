"""world_size-2 gloo test of the N>1 host logic (sharding + the single flat gradient all-reduce)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from neural_ode_ion_channels_b200 import parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    B = 11
    y0 = torch.arange(B * 2, dtype=torch.float64).reshape(B, 2)
    g = torch.arange(B, dtype=torch.float64)
    mine = parallel.shard_batch({'y0': y0, 'g': g, 'E': None})
    lo, hi = parallel.shard_bounds(B, world, rank)
    assert mine['E'] is None and torch.equal(mine['y0'], y0[lo:hi]) and torch.equal(mine['g'], g[lo:hi])
    # per-rank "gradient" = a known function of the shard; the all-reduce must give the full-batch sum
    grads = [mine['g'].sum() * torch.ones(3, 2, dtype=torch.float64),
             mine['y0'].sum() * torch.ones(5, dtype=torch.float64)]
    loss = mine['g'].pow(2).sum()
    out, tot = parallel.allreduce_gradients(grads, loss)
    assert torch.allclose(out[0], g.sum() * torch.ones(3, 2, dtype=torch.float64))
    assert torch.allclose(out[1], y0.sum() * torch.ones(5, dtype=torch.float64))
    assert torch.allclose(tot, g.pow(2).sum())
    dist.barrier()
    dist.destroy_process_group()
    ret[rank] = True


def test_shard_and_flat_allreduce_gloo_world2():
    world = 2
    port = 29500 + os.getpid() % 2000
    ctx = mp.get_context('spawn')
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert all(ret.get(r) for r in range(world))


def test_shard_bounds_and_lpt():
    sys.path.insert(0, ROOT)
    from neural_ode_ion_channels_b200 import parallel
    for n, w in ((65536, 8), (11, 4), (3, 8), (0, 2)):
        spans = [parallel.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    # s00..s11 MACs: LPT puts the three 500-wide nets on different ranks
    macs = [200600, 40600, 400600, 530, 130, 1030, 1251500, 251500, 2501500, 50300, 10300, 100300]
    assign = parallel.lpt_assign(macs, 8)
    assert sorted(i for a in assign for i in a) == list(range(12))
    big = [next(r for r, a in enumerate(assign) if i in a) for i in (8, 6, 2)]
    assert len(set(big)) == 3


def _flat_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from neural_ode_ion_channels_b200 import parallel
    flat = torch.arange(10, dtype=torch.float64) * (rank + 1)
    loss = torch.tensor(float(rank + 1), dtype=torch.float64)
    out, tot = parallel.allreduce_flat(flat, loss)
    # the architecture sweep of bench.py: 24 fits, longest first, every fit on exactly one rank
    costs = [float((i * 7919) % 97 + 1) for i in range(24)]
    assign = parallel.lpt_assign(costs, world)
    q.put((rank, out.tolist(), float(tot), assign))
    dist.destroy_process_group()


def test_flat_allreduce_and_sweep_assignment_world2():
    """bench.py's multi-GPU legs on the CPU: ONE flat all-reduce carries the gradient and the loss
    (train / train1m legs), and the LPT assignment of the 24 sweep fits is a partition that every
    rank computes identically (no collective on the sweep's data path)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    want = [3.0 * i for i in range(10)]
    for rank, out, tot, assign in res:
        assert out == want and tot == 3.0
        assert sorted(i for part in assign for i in part) == list(range(24))
    assert res[0][3] == res[1][3]
    loads = [sum(float((i * 7919) % 97 + 1) for i in part) for part in res[0][3]]
    total = sum(loads)
    assert max(loads) <= total / 2 + 97            # LPT: within one largest item of the ideal split
