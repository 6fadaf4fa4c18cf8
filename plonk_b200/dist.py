"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The prover path shards by independent proofs (no data-path collective).  The
only exchange step is the one a single large MSM needs when its points are partitioned across GPUs
(section 8e-ii, BASELINE.json configs[3]): every rank reduces its contiguous slice of the commit key
to one G1 point, the 96-byte results are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU
tests) and each rank adds them locally - EC addition is not an NCCL reduction op."""
from __future__ import annotations

import ctypes
from typing import List, Tuple

from ._lib import check, lib

G1_RAW_BYTES = 96


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [first, first+count) slice of n items for `rank`; sizes differ by at most one."""
    base, extra = divmod(n, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def shard_proofs(total: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of independent proofs to ranks (BASELINE.json configs[4])."""
    return list(range(rank, total, world))


def g1_sum(points_raw: List[bytes]) -> bytes:
    """Sum of affine points in the 96-byte raw layout (host-side; a few dozen additions at most)."""
    acc = bytes(G1_RAW_BYTES)
    out = ctypes.create_string_buffer(G1_RAW_BYTES)
    for p in points_raw:
        check(lib().pb200_g1_add_affine(acc, p, out))
        acc = out.raw
    return acc


def allgather_g1_sum(partial_raw: bytes, device=None) -> bytes:
    """All-gather every rank's partial MSM result and add them (same value on all ranks)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return partial_raw
    t = torch.frombuffer(bytearray(partial_raw), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return g1_sum([bytes(o.cpu().numpy().tobytes()) for o in outs])


def max_over_ranks(ms: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_msm(key, scalars: bytes, device=None) -> bytes:
    """One MSM over the whole commit key with points partitioned across ranks.  `key` is a
    plonk_b200.CommitKey holding the full key on every rank (bases are static and pre-placed);
    `scalars` is the full scalar vector, each rank only touches its slice."""
    import torch.distributed as dist

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    n = len(scalars) // 32
    first, count = shard_range(n, rank, world)
    out = ctypes.create_string_buffer(G1_RAW_BYTES)
    sl = scalars[first * 32 : (first + count) * 32]
    check(lib().pb200_msm_g1_range(key._h, first, sl if count else None, count, out))
    return allgather_g1_sum(out.raw, device)


class NcclComm:
    """An ncclComm_t over the ranks of the default torch.distributed group, for the C ABI's
    pb200_msm_g1_allgather (a host without torch creates one with ncclCommInitRank the same way).
    The unique id travels through torch.distributed's own broadcast."""

    class _UniqueId(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_char * 128)]

    def __init__(self):
        import torch.distributed as dist

        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self._nccl = ctypes.CDLL("libnccl.so.2")  # the copy torch already loaded, if any
        uid = self._UniqueId()
        if self.rank == 0:
            self._check(self._nccl.ncclGetUniqueId(ctypes.byref(uid)))
        # string_at, not the field: ctypes hands a c_char array field back truncated at its first NUL
        box = [ctypes.string_at(ctypes.addressof(uid), 128) if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        assert len(box[0]) == 128
        ctypes.memmove(ctypes.addressof(uid), box[0], 128)
        self.handle = ctypes.c_void_p()
        self._nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, self._UniqueId, ctypes.c_int]
        self._check(self._nccl.ncclCommInitRank(ctypes.byref(self.handle), self.world, uid, self.rank))

    def _check(self, rc: int):
        if rc != 0:
            self._nccl.ncclGetErrorString.restype = ctypes.c_char_p
            raise RuntimeError("NCCL: " + self._nccl.ncclGetErrorString(rc).decode())

    def destroy(self):
        if self.handle:
            self._nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
            self._nccl.ncclCommDestroy(self.handle)
            self.handle = None


class ShardedCommitKey:
    """A commit key whose points are partitioned across the ranks of `comm`: rank r
    holds points [first, first + count) of the n_points-long key, every slice uploaded with ONE window width
    (pb200_msm_window_for of the largest slice) so the ranks' digit sums can be added.  `slice_raw` is this
    rank's 96-byte raw points.

    MSMs shorter than `threshold` scalars do not pay for an exchange: when `replica_raw` (the first `threshold`
    points of the key, the same bytes on every rank) is given, they run on that replica on every rank with no
    collective - a 2^16-point MSM is 1.7 ms on one GPU and was 2.8 ms across two in round 1."""

    def __init__(self, slice_raw: bytes, n_points: int, comm, threshold: int = 1 << 18, replica_raw: bytes | None = None):
        self.comm, self.n_points, self.threshold = comm, n_points, threshold
        world, rank = (comm.world, comm.rank) if comm is not None else (1, 0)
        self.first, self.count = shard_range(n_points, rank, world)
        assert len(slice_raw) == self.count * G1_RAW_BYTES
        L = lib()
        self.window = L.pb200_msm_window_for(shard_range(n_points, 0, world)[1])
        self._h = ctypes.c_void_p()
        if self.count:
            check(L.pb200_srs_upload_window(slice_raw, self.count, self.window, ctypes.byref(self._h)))
        else:  # more ranks than points: a one-point placeholder that is never read (n_scalars = 0)
            check(L.pb200_srs_upload_window(bytes(G1_RAW_BYTES), 1, self.window, ctypes.byref(self._h)))
        self._replica = ctypes.c_void_p()
        self.replica_points = 0
        if replica_raw:
            self.replica_points = len(replica_raw) // G1_RAW_BYTES
            check(L.pb200_srs_upload(replica_raw, self.replica_points, ctypes.byref(self._replica)))

    def slice_of(self, n_scalars: int):
        """(first, count) of this rank's share of an MSM over the first n_scalars points."""
        lo = min(self.first, n_scalars)
        return lo, max(0, min(n_scalars, self.first + self.count) - lo)

    def uses_collective(self, n_scalars: int) -> bool:
        world = self.comm.world if self.comm is not None else 1
        return world > 1 and not (self._replica and n_scalars <= min(self.threshold, self.replica_points))

    def commit_dev(self, d_scalars: int, n_scalars: int, stream=None) -> bytes:
        """One commitment from a device-resident scalar vector (full length on every rank; each rank reads
        only its slice).  Returns the 96-byte affine sum, the same on every rank."""
        L = lib()
        out = ctypes.create_string_buffer(G1_RAW_BYTES)
        if not self.uses_collective(n_scalars):
            # the replica below the threshold; with a single rank the "slice" is the whole key
            h = self._replica if (self._replica and n_scalars <= self.replica_points) else self._h
            check(L.pb200_msm_g1_dev(h, d_scalars, n_scalars, 1, max(n_scalars, 1), out, stream))
            return out.raw
        lo, cnt = self.slice_of(n_scalars)
        ptr = ctypes.c_void_p(d_scalars + 32 * lo) if cnt else None
        check(L.pb200_msm_g1_allgather_dev(self._h, ptr, cnt, 1, max(cnt, 1), self.comm.handle, self.comm.world, out, stream))
        return out.raw

    def free(self):
        for h in (self._h, self._replica):
            if h:
                lib().pb200_srs_free(h)
        self._h = self._replica = ctypes.c_void_p()


def sharded_msm_nccl(key_slice, scalars_slice: bytes, comm: NcclComm) -> bytes:
    """pb200_msm_g1_allgather: `key_slice` is a CommitKey holding this rank's points only,
    `scalars_slice` the matching coefficients.  Returns the 96-byte sum (same on every rank)."""
    out = ctypes.create_string_buffer(G1_RAW_BYTES)
    n = len(scalars_slice) // 32
    check(lib().pb200_msm_g1_allgather(key_slice._h, scalars_slice if n else None, n, 1, max(n, 1), comm.handle, comm.world, out))
    return out.raw
