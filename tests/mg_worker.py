"""Worker of tests/test_gpu_multi.py (launched by torchrun, one process per GPU): the output gather on real NCCL.

N sharding: every rank prunes the same weights, multiplies them into ITS images' columns, writing D_r straight into
slab r of the gather arena; one in-place spfy_mg_allgather fills in the others;
then the same through the FUSED gather (PeerArena + replicated plan: the epilogue stores into every peer's memory) -- the
two arenas must be bit-identical.  Every rank can regenerate every other
rank's inputs (seeded by rank), so it checks all slabs against a torch fp32 matmul of the pruned weights.
Layer sharding: spfy_mg_broadcast_many from each layer's owner."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def inputs(rank, dev, M, K, N):
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    w = (torch.rand(M, K, device=dev, generator=g) * 2 - 1).half()
    g.manual_seed(99 + rank)
    b = (torch.rand(K, N, device=dev, generator=g) * 2 - 1).half()
    return w, b


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    spfy = ge.load_package()
    gather = spfy.multigpu.OutputGather()
    assert (gather.rank, gather.world) == (rank, world)
    shapes = [(256, 576, 3136 * 2), (64, 152, 1000), (512, 1024, 392)]
    # ---- N sharding: [g][M][N/g] arena, GEMM output in place, one all-gather ----
    elems = sum(M * N for M, K, N in shapes)
    arena = torch.zeros(world, elems, dtype=torch.float16, device=dev)
    off, probs, views = 0, [], []
    for M, K, N in shapes:
        w, b = inputs(rank, dev, M, K, N)
        comp = spfy.prune24(w)
        d = arena[rank, off: off + M * N].view(M, N)
        probs.append(dict(comp=comp, b=b, out=d))
        views.append((off, M, K, N))
        off += M * N
    plan = spfy.SpmmaPlan(probs)
    plan.run()
    received = gather.allgather(arena)
    torch.cuda.synchronize()
    assert received == elems * 2 * (world - 1)
    for r in range(world):
        for off, M, K, N in views:
            w, b = inputs(r, dev, M, K, N)
            pruned = torch.empty_like(w)
            spfy.prune24(w, out_dense=pruned, compress=False)
            want = pruned.float() @ b.float()
            got = arena[r, off: off + M * N].view(M, N).float()
            scale = torch.clamp(want.abs(), min=1e-2 * float(want.abs().max()))
            err = float(((got - want).abs() / scale).max())
            assert err <= 1e-2, (rank, r, M, K, N, err)
    # ---- the fused gather: the GEMM epilogue stores every tile into every peer's arena (no collective) ----
    slab = -(-elems * 2 // 128) * 128
    pa = spfy.multigpu.PeerArena(slab)
    pa.local.zero_()
    torch.cuda.synchronize()
    dist.barrier()  # nobody stores into an arena its owner is still clearing
    mine = pa.local[rank].view(torch.float16)
    probs2 = []
    for q, (off, M, K, N) in zip(probs, views):
        probs2.append(dict(comp=q["comp"], b=q["b"], out=mine[off: off + M * N].view(M, N),
                           replicas=pa.replica_addresses(off * 2)))
    plan2 = spfy.SpmmaPlan(probs2)
    assert plan2.replicas == world - 1
    plan2.run()
    pa.barrier()
    torch.cuda.synchronize()
    fused = pa.local.view(torch.float16)[:, :elems]
    assert torch.equal(fused, arena), (rank, "fused gather differs from GEMM + ncclAllGather")
    plan2.close()
    pa.close()
    # ---- layer sharding: every layer's output travels from its owner ----
    costs = [M * K + K * N + M * N for M, K, N in shapes]
    owned = spfy.multigpu.partition_layers_lpt(costs, world)
    owner = [next(r for r in range(world) if i in owned[r]) for i in range(len(shapes))]
    outs = []
    for i, (M, K, N) in enumerate(shapes):
        t = torch.full((M, N), float(owner[i] + 1) if owner[i] == rank else -1.0, dtype=torch.float32, device=dev)
        outs.append(t)
    gather.broadcast_many(outs, owner)
    torch.cuda.synchronize()
    for i, t in enumerate(outs):
        assert bool((t == float(owner[i] + 1)).all()), (rank, i)
    gather.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MG_WORKER_OK")


if __name__ == "__main__":
    main()
