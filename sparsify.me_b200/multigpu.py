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

The partitioning helpers are host logic covered by world_size-2 gloo tests on CPU.  `OutputGather` is the data path
on GPUs: the C ABI's spfy_mg_* (NCCL loaded at run time) on the ranks' own communicator, bootstrapped through
torch.distributed -- ncclAllGather straight into the documented [g][M][N/g] layout, in place, no transpose.
"""
import ctypes
import os

import torch
import torch.distributed as dist


def nccl_library_path():
    """libnccl.so.2 as shipped with PyTorch (nvidia-nccl wheel), else whatever the loader finds"""
    try:
        import nvidia.nccl as n
        base = os.path.dirname(n.__file__) if getattr(n, "__file__", None) else list(n.__path__)[0]
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    except Exception:  # noqa: BLE001
        pass
    return "libnccl.so.2"


class OutputGather:
    """spfy_mg_* on one communicator per process (include/spfy_b200.h, "Multi-GPU").

    N sharding: every rank allocates `slab_bytes * world` for a layer (or a group of layers), lets its GEMMs write
    D_r into slab `rank` (`slab(buf)`), and `allgather(buf)` fills in the other ranks' slabs in place.
    Layer sharding: `broadcast_many(tensors, owners)` sends each layer's output from its owner to everybody as one
    NCCL group."""

    def __init__(self, group=None):
        from . import capi
        self._capi = capi
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        os.environ.setdefault("SPFY_NCCL_LIB", nccl_library_path())
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = ctypes.create_string_buffer(128)
            capi.spfy_mg_unique_id(ctypes.cast(buf, ctypes.c_void_p))
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        dev = torch.device("cuda", torch.cuda.current_device())
        backend = dist.get_backend(group)
        t = uid.to(dev) if backend == "nccl" else uid
        dist.broadcast(t, src=0, group=group)  # 128 bytes over whatever the process group runs on
        raw = bytes(t.cpu().numpy().tobytes())
        self._h = ctypes.c_void_p()
        capi.spfy_mg_create(self.rank, self.world, ctypes.c_char_p(raw), ctypes.byref(self._h))

    def slab(self, buf, rank=None):
        """view of rank's slab of a gather buffer whose leading dimension is the world size"""
        return buf[self.rank if rank is None else rank]

    def allgather(self, buf, stream=None):
        """buf: [world, ...] contiguous; slab `rank` holds this rank's data, the others are filled in"""
        assert buf.is_contiguous() and buf.shape[0] == self.world
        per = buf[0].numel() * buf.element_size()
        s = stream if stream is not None else torch.cuda.current_stream()
        send = buf.data_ptr() + self.rank * per
        self._capi.spfy_mg_allgather(self._h, ctypes.c_void_p(send), ctypes.c_void_p(buf.data_ptr()), per,
                                     ctypes.c_void_p(s.cuda_stream))
        return per * (self.world - 1)  # bytes this rank received

    def broadcast_many(self, tensors, owners, stream=None):
        n = len(tensors)
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])
        sizes = (ctypes.c_size_t * n)(*[t.numel() * t.element_size() for t in tensors])
        roots = (ctypes.c_int * n)(*owners)
        s = stream if stream is not None else torch.cuda.current_stream()
        self._capi.spfy_mg_broadcast_many(self._h, ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(sizes, ctypes.c_void_p),
                                          ctypes.cast(roots, ctypes.c_void_p), n, ctypes.c_void_p(s.cuda_stream))
        return sum(t.numel() * t.element_size() for t, o in zip(tensors, owners) if o != self.rank)

    def close(self):
        if self._h:
            self._capi.spfy_mg_destroy(self._h)
            self._h = None


class _DeviceBuffer:
    """raw device memory as seen by torch.as_tensor (CUDA array interface v2)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerArena:
    """A gather arena every rank of the box can store into: the fused form of the output gather.

    Every rank allocates `world * slab_bytes` of peer-mappable device memory (spfy_peer_alloc), the 64-byte handles
    travel through torch.distributed, and every rank maps every other rank's arena (spfy_peer_open).  Rank r's GEMMs
    write D_r into slab r of its own arena and, through `SpmmaPlan(..., replicas=...)`, into slab r of every peer's
    arena in the same TMA stores: when all ranks' launches have completed (`barrier()`), every arena holds
    [g][...] exactly as `OutputGather.allgather` would have left it -- without a collective.
    `local` is this rank's arena as a uint8 tensor [world, slab_bytes]."""

    def __init__(self, slab_bytes, group=None):
        from . import capi
        self._capi = capi
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.group = group
        self.slab_bytes = int(slab_bytes)
        assert self.slab_bytes % 128 == 0, "slab size must be a multiple of 128 bytes"
        self.dev = torch.device("cuda", torch.cuda.current_device())
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        capi.spfy_peer_alloc(self.world * self.slab_bytes, ctypes.byref(ptr), ctypes.cast(handle, ctypes.c_void_p))
        self.base = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.peer_base = [None] * self.world  # base address of rank p's arena as mapped HERE
        for p_, h in enumerate(handles):
            if p_ == self.rank:
                self.peer_base[p_] = self.base
                continue
            q = ctypes.c_void_p()
            capi.spfy_peer_open(ctypes.c_char_p(h), ctypes.byref(q))
            self.peer_base[p_] = q.value
        self.local = torch.as_tensor(_DeviceBuffer(self.base, self.world * self.slab_bytes), device=self.dev).view(
            self.world, self.slab_bytes)
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.dev)

    def replica_addresses(self, byte_offset):
        """where a matrix at `byte_offset` of MY slab lives in every PEER's arena (addresses valid in this process)"""
        off = self.rank * self.slab_bytes + int(byte_offset)
        return [self.peer_base[p_] + off for p_ in range(self.world) if p_ != self.rank]

    def barrier(self):
        """stream-ordered rendezvous: returns (on the stream) once every rank's earlier work on its current stream has
        completed, i.e. once every slab of every arena is written"""
        dist.all_reduce(self._flag, group=self.group)

    def close(self):
        if self.base is None:
            return
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # nobody unmaps or frees while a peer may still be storing
        for p_, q in enumerate(self.peer_base):
            if p_ != self.rank and q:
                self._capi.spfy_peer_close(ctypes.c_void_p(q))
        dist.barrier(group=self.group)
        self.local = None
        self._capi.spfy_peer_free(ctypes.c_void_p(self.base))
        self.base = None


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
