"""Multi-GPU sharding of the per-layer problems (one process per GPU, torch.distributed).

The reference has no multi-GPU code.  The per-layer GEMMs of datasets/*.csv
are independent and, inside one GEMM, so are the N = batch x spatial columns, so the path
shards with NO data-path collective (weak scaling); the only exchange is an optional gather of
outputs.  Two plans:

  layer sharding (BASELINE config 4): layers -> ranks by LPT on algorithmic bytes
      (the kernels are HBM-bound, so bytes, not FLOPs, balance time);
  N sharding (BASELINE config 5): rank r owns whole images [r*b/g, (r+1)*b/g) of every layer,
      i.e. columns [r*N/g, (r+1)*N/g) of B and D; weights are replicated and pruned
      redundantly (<= 4.7 MB, cheaper than a broadcast).

Everything here is host logic and is covered by world_size-2 gloo tests on CPU.
"""
import torch
import torch.distributed as dist


def partition_layers_lpt(costs, world):
    """Longest-processing-time-first assignment.  Returns a list (per rank) of layer indices,
    each in ascending order.  Deterministic: ties go to the lower rank / lower layer index."""
    loads = [0.0] * world
    owned = [[] for _ in range(world)]
    for i in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
        r = min(range(world), key=lambda r: (loads[r], r))
        loads[r] += costs[i]
        owned[r].append(i)
    return [sorted(o) for o in owned]


def shard_batch(batch, world, rank):
    """Images [lo, hi) of the batch owned by `rank` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_columns(spatial, batch, world, rank):
    """Column range [c0, c1) of the N = spatial*batch dimension owned by `rank`, split on image
    boundaries so that a shard equals data-parallel inference."""
    lo, hi = shard_batch(batch, world, rank)
    return lo * spatial, hi * spatial


def gather_columns(local, counts, group=None):
    """Gather per-rank column slabs [M, n_r] into the full [M, sum n_r] row-major matrix on
    every rank.  Uses all_gather on [n_r, M]-transposed slabs padded to the largest shard so the
    collective has equal counts (NCCL all-gather over NVLink on GPUs, gloo on CPU)."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    m = local.shape[0]
    nmax = max(counts)
    slab = torch.zeros(nmax, m, dtype=local.dtype, device=local.device)
    slab[: local.shape[1]] = local.t()
    out = torch.empty(world * nmax, m, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, slab, group=group)
    parts = [out[r * nmax: r * nmax + counts[r]] for r in range(world)]
    return torch.cat(parts, 0).t().contiguous()


def gather_layers(local_outputs, owners, group=None):
    """All ranks obtain every layer's output.  `owners[i]` is the rank that computed layer i;
    `local_outputs` maps layer index -> tensor on the owner.  Shapes must be known to every rank:
    pass tensors of the right shape (uninitialised) for layers a rank does not own."""
    world = dist.get_world_size(group)
    if world == 1:
        return local_outputs
    for i, t in sorted(local_outputs.items()):
        dist.broadcast(t, src=owners[i], group=group)
    return local_outputs
