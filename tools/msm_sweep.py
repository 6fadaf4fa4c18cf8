"""BASELINE.json configs[3]: standalone G1 MSM sweep with the points of one MSM partitioned across
the GPUs of a box; per-rank partial results are all-gathered over NCCL and added locally
(plonk_b200/dist.py).

  python tools/msm_sweep.py --sizes 16,18,20                       # one GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/msm_sweep.py --sizes 16,20,22,24

Every rank generates the same seeded commit key on its own GPU ([x^i] g, pb200_srs_setup_from_secret),
uploads only its slice, and checks the reduced result against [p(x)] g computed by rank 0 on the GPU
as a 1-point MSM."""
import argparse
import ctypes
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonk_b200 import dist as pd  # noqa: E402
from plonk_b200._lib import check, lib  # noqa: E402

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
mont = lambda v: ((v << 256) % R_MOD).to_bytes(32, "little")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="16,18,20")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--via", default="torch", choices=["torch", "capi"],
                    help="torch: device-resident scalars + torch.distributed all_gather; capi: pb200_msm_g1_allgather "
                         "(host scalars, ncclAllGather inside the C ABI on a communicator created here)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = lib()
    check(L.pb200_init(local))
    x, gs = 0x1234567, 0x7654321
    comm = pd.NcclComm() if (args.via == "capi" and world > 1) else None
    if args.via == "capi" and comm is None:
        raise SystemExit("--via capi needs torchrun with at least 2 ranks")
    for log_n in (int(v) for v in args.sizes.split(",")):
        n = 1 << log_n
        first, count = pd.shard_range(n, rank, world)
        # this rank's slice of the key: [x^(first+i)] g = [x^i] ([x^first] g)
        slice_raw = ctypes.create_string_buffer(96 * count)
        check(L.pb200_srs_setup_from_secret(mont(x), mont(gs * pow(x, first, R_MOD) % R_MOD), count, slice_raw))
        h = ctypes.c_void_p()
        check(L.pb200_srs_upload(slice_raw, count, ctypes.byref(h)))
        gen = torch.Generator(device="cuda").manual_seed(1000 + log_n)
        sc = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda", generator=gen)  # same on every rank
        mine = sc[first : first + count].contiguous()
        out = ctypes.create_string_buffer(96)
        dev = torch.device("cuda", local)
        times = []
        for it in range(args.iters + 1):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            if comm is not None:
                if it == 0:
                    mine_host = mine.cpu().numpy().tobytes()
                check(L.pb200_msm_g1_allgather(h, mine_host, count, 1, count, comm.handle, world, out))
                total = out.raw
            else:
                check(L.pb200_msm_g1_dev(h, mine.data_ptr(), count, 1, count, out, None))
                total = pd.allgather_g1_sum(out.raw, dev)
            torch.cuda.synchronize()
            ms = pd.max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
            if it:
                times.append(ms)
        # check: sum_i s_i x^i evaluated on the host (canonical scalars from the Montgomery tensor)
        if rank == 0 and log_n <= 20:
            rinv = pow(1 << 256, -1, R_MOD)
            vals = sc.cpu().numpy().astype("uint64")
            acc, p = 0, 1
            for row in vals:
                v = (int(row[0]) | int(row[1]) << 64 | int(row[2]) << 128 | int(row[3]) << 192) * rinv % R_MOD
                acc = (acc + v * p) % R_MOD
                p = p * x % R_MOD
            one_pt = ctypes.create_string_buffer(96)
            check(L.pb200_srs_setup_from_secret(mont(1), mont(gs * acc % R_MOD), 1, one_pt))
            assert one_pt.raw == total, "sharded MSM result differs from [p(x)] g"
        L.pb200_srs_free(h)
        if rank == 0:
            best = min(times)
            print(json.dumps({"log_n": log_n, "gpus": world, "ms": best, "points_per_s": n / best * 1e3,
                              "checked": log_n <= 20, "via": args.via}), flush=True)
    if comm is not None:
        comm.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
