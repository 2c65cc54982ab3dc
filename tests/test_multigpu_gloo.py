"""Host logic of the multi-GPU path on CPU: world_size-2 gloo processes (no data-path collective;
the only exchange is the output gather)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spfy = ge.load_package()
        mg = spfy.multigpu
        # --- N sharding on image boundaries + gather of the column slabs ---
        spatial, batch, m = 7, 5, 6           # N = 35 columns, uneven 3/2 image split
        full = torch.arange(m * spatial * batch, dtype=torch.float32).reshape(m, spatial * batch)
        counts = []
        for r in range(world):
            c0, c1 = mg.shard_columns(spatial, batch, world, r)
            counts.append(c1 - c0)
        c0, c1 = mg.shard_columns(spatial, batch, world, rank)
        local = full[:, c0:c1].contiguous()
        got = mg.gather_columns(local, counts)
        ok1 = torch.equal(got, full)
        # --- layer sharding by LPT + broadcast gather ---
        costs = [5.0, 1.0, 4.0, 2.0, 2.0, 3.0]
        owned = mg.partition_layers_lpt(costs, world)
        owners = {i: r for r, idx in enumerate(owned) for i in idx}
        outs = {i: (torch.full((3,), float(i)) if owners[i] == rank else torch.zeros(3)) for i in range(len(costs))}
        mg.gather_layers(outs, owners)
        ok2 = all(torch.equal(outs[i], torch.full((3,), float(i))) for i in outs)
        q.put((rank, ok1, ok2, owned))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok1, ok2, owned in res:
        assert ok1 and ok2
        assert sorted(i for o in owned for i in o) == list(range(6))


def test_partition_and_shard_helpers(spfy):
    mg = spfy.multigpu
    owned = mg.partition_layers_lpt([5, 1, 4, 2, 2, 3], 2)
    loads = [sum([5, 1, 4, 2, 2, 3][i] for i in o) for o in owned]
    assert abs(loads[0] - loads[1]) <= 1
    assert mg.partition_layers_lpt([], 4) == [[], [], [], []]
    for world in (1, 2, 4, 8):
        spans = [mg.shard_batch(32, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 32
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    assert [mg.shard_batch(5, 2, r) for r in range(2)] == [(0, 3), (3, 5)]
    assert mg.shard_columns(196, 256, 8, 3) == (3 * 32 * 196, 4 * 32 * 196)
