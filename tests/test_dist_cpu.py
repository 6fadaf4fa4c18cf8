"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, G1 all-gather + local sum,
max-over-ranks timing.  The EC additions run through the library's host-side helper, which needs
no GPU."""
import os
import random
import socket

import torch.multiprocessing as mp

from oracle import pyref as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, partials, expect, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from plonk_b200 import dist as pd

    total = pd.allgather_g1_sum(partials[rank])
    mx = pd.max_over_ranks(10.0 + rank)
    q.put((rank, total == expect, mx, pd.shard_range(10, rank, world), pd.shard_proofs(7, rank, world)))
    dist.destroy_process_group()


def test_allgather_g1_sum_and_sharding_world2():
    rng = random.Random(2)
    pts = [R.g1_mul(R.G1_GEN, rng.randrange(1, R.R_MOD)) for _ in range(2)]
    partials = [R.g1_to_raw_bytes(p) for p in pts]
    expect = R.g1_to_raw_bytes(R.g1_add(pts[0], pts[1]))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, partials, expect, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [11.0, 11.0]
    assert [r[3] for r in res] == [(0, 5), (5, 5)]
    assert [r[4] for r in res] == [[0, 2, 4, 6], [1, 3, 5]]


def test_host_g1_helpers_match_oracle():
    from plonk_b200 import dist as pd
    import ctypes
    from plonk_b200._lib import check, lib

    rng = random.Random(3)
    pts = [R.g1_mul(R.G1_GEN, rng.randrange(1, R.R_MOD)) for _ in range(5)] + [None]
    pts.append(R.g1_neg(pts[0]))
    want = None
    for p in pts:
        want = R.g1_add(want, p)
    assert R.g1_from_raw_bytes(pd.g1_sum([R.g1_to_raw_bytes(p) for p in pts])) == want
    assert R.g1_from_raw_bytes(pd.g1_sum([R.g1_to_raw_bytes(pts[1]), R.g1_to_raw_bytes(pts[1])])) == R.g1_add(pts[1], pts[1])
    out = ctypes.create_string_buffer(48)
    for p in pts:
        check(lib().pb200_g1_compress(R.g1_to_raw_bytes(p), out))
        assert out.raw == R.g1_compress(p)


def test_msm_combine_parts_matches_the_bucket_reduction_formula():
    """Host tail of the point-sharded MSM (pb200_msm_combine_parts, the finish of pb200_msm_g1_allgather*): the
    per-digit sums of several ranks are added digit by digit, then R = sum A + g * sum_j 2^shift_j D_j."""
    import ctypes

    from plonk_b200._lib import check, lib

    rng = random.Random(11)
    one = ((1 << 384) % R.P_MOD).to_bytes(48, "little")

    def xyzz(p):
        return bytes(192) if p is None else R.g1_to_raw_bytes(p) + one + one

    for c, n_parts, batch in ((13, 3, 2), (16, 2, 1), (4, 1, 3), (20, 8, 1)):
        nb = 1 << (c - 1)
        g = min(8, nb)
        n_groups = nb // g
        total_bits = max(0, (n_groups - 1).bit_length())
        ndig = (total_bits + 3) // 4
        shifts, sh = [], 0
        for j in range(ndig):
            shifts.append(sh)
            sh += (total_bits - sh) // (ndig - j)
        wpe = ctypes.c_size_t()
        check(lib().pb200_msm_combine_parts(None, 0, c, batch, None, ctypes.byref(wpe)))
        assert wpe.value == (ndig + 1) * 48
        pts = [[[R.g1_mul(R.G1_GEN, rng.randrange(1, R.R_MOD)) if rng.random() < 0.8 else None for _ in range(ndig + 1)]
                for _ in range(batch)] for _ in range(n_parts)]
        blob = b"".join(xyzz(p) for part in pts for entry in part for p in entry)
        out = ctypes.create_string_buffer(96 * batch)
        check(lib().pb200_msm_combine_parts(blob, n_parts, c, batch, out, ctypes.byref(wpe)))
        for b in range(batch):
            want = None
            for part in pts:
                for j in range(ndig):
                    if part[b][j] is not None:
                        want = R.g1_add(want, R.g1_mul(part[b][j], g << shifts[j]))
                want = R.g1_add(want, part[b][ndig])
            assert R.g1_from_raw_bytes(out.raw[96 * b : 96 * b + 96]) == want, (c, b)


def test_sharded_key_slices_and_threshold_logic():
    """plonk_b200.dist.ShardedCommitKey's host-side bookkeeping (which points a rank reads, when the exchange
    is skipped) without touching a GPU."""
    from plonk_b200 import dist as pd

    class FakeComm:
        def __init__(self, rank, world):
            self.rank, self.world = rank, world

    k = pd.ShardedCommitKey.__new__(pd.ShardedCommitKey)
    k.comm, k.n_points, k.threshold = FakeComm(2, 4), 1000, 256
    k.first, k.count = pd.shard_range(1000, 2, 4)
    k._replica, k.replica_points = object(), 256
    assert (k.first, k.count) == (500, 250)
    assert k.slice_of(1000) == (500, 250) and k.slice_of(600) == (500, 100) and k.slice_of(500) == (500, 0) and k.slice_of(10) == (10, 0)
    assert not k.uses_collective(256) and k.uses_collective(257)
    k._replica = None
    assert k.uses_collective(10)
    k.comm = FakeComm(0, 1)
    assert not k.uses_collective(10**6)
