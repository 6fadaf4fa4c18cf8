"""Kernel-level timing of the NTT and MSM through the device-pointer C ABI (development aid).

Inputs are resident in HBM (torch tensors used only as device buffers); timing with CUDA events on
the launching stream.  Prints butterflies/s and G1 adds/s next to the measured IMAD.WIDE peak."""
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonk_b200._lib import check, lib  # noqa: E402

L = lib()
check(L.pb200_init(0))
_s = torch.cuda.Stream()
torch.cuda.set_stream(_s)
stream = _s.cuda_stream
assert stream != 0


def time_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rand_fr_dev(n):
    # random 255-bit values below r: clear the top 2 bits of the top limb (values < 2^254 < r)
    t = torch.randint(0, 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda")
    t[:, 3] &= (1 << 61) - 1
    return t


res = {}
d = ctypes.c_double()
check(L.pb200_imad_peak(ctypes.byref(d)))
res["imad_wide_per_s"] = d.value
for log_n, batch in [(12, 4), (16, 1), (16, 4), (19, 1), (19, 5), (20, 1), (23, 1)]:
    n = 1 << log_n
    x = rand_fr_dev(n * batch)
    y = torch.empty_like(x)
    for inv, coset in [(0, 0), (1, 1)]:
        f = lambda: check(L.pb200_ntt_dev(x.data_ptr(), n, y.data_ptr(), log_n, inv, coset, batch, n, n, stream))
        ms = time_ms(f)
        bf = batch * (n // 2) * log_n
        res[f"ntt_2^{log_n}_b{batch}_inv{inv}_coset{coset}"] = dict(ms=ms, gbutterflies_per_s=bf / ms / 1e6, gbytes_per_s=64 * n * batch / ms / 1e6)
        print(f"ntt 2^{log_n} batch {batch} inv={inv} coset={coset}: {ms:.4f} ms  {bf/ms/1e6:.2f} G butterflies/s  {64*n*batch/ms/1e6:.1f} GB/s algorithmic", flush=True)

# MSM: a true SRS shape (powers of a secret times a generator multiple), made on the device by the library itself
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
mont = lambda v: ((v << 256) % R_MOD).to_bytes(32, "little")
for log_n in (12, 16, 18):
    n = (1 << log_n) + 7
    srs_raw = ctypes.create_string_buffer(n * 96)
    t0 = time.time()
    check(L.pb200_srs_setup_from_secret(mont(12345), mont(6789), n, srs_raw))
    print(f"device srs setup 2^{log_n}: {time.time()-t0:.2f}s", flush=True)
    h = ctypes.c_void_p()
    t0 = time.time()
    check(L.pb200_srs_upload(srs_raw, n, ctypes.byref(h)))
    print(f"srs upload+precompute 2^{log_n}: {time.time()-t0:.3f}s", flush=True)
    for batch in (1, 4):
        s = rand_fr_dev(n * batch)
        out = ctypes.create_string_buffer(96 * batch)
        f = lambda: check(L.pb200_msm_g1_dev(h, s.data_ptr(), n, batch, n, out, stream))
        ms = time_ms(f, iters=5, warm=2)
        L.pb200_profile_enable(1)
        f(); f()
        acc_ms, adds, nl, pts = ctypes.c_double(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        L.pb200_profile_read(ctypes.byref(acc_ms), ctypes.byref(adds), ctypes.byref(nl), ctypes.byref(pts))
        L.pb200_profile_enable(0)
        print(f"   accumulate kernel: {acc_ms.value/2:.3f} ms/launch, {adds.value/acc_ms.value/1e6:.3f} G adds/s", flush=True)
        res[f"msm_2^{log_n}_b{batch}"] = dict(ms=ms, mpoints_per_s=n * batch / ms / 1e3)
        print(f"msm 2^{log_n} batch {batch}: {ms:.3f} ms  ({n*batch/ms/1e3:.2f} M points/s)", flush=True)
    L.pb200_srs_free(h)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/kbench.json", "w"), indent=1)
