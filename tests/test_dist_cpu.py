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
