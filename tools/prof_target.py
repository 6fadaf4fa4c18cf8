"""Small GPU workload for ncu captures: one batch-4 MSM over 2^16+7 points and one batch-5 coset NTT of 2^19."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonk_b200._lib import check, lib
L = lib(); check(L.pb200_init(0))
n = (1 << 16) + 7
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
mont = lambda v: ((v << 256) % R_MOD).to_bytes(32, "little")
srs = ctypes.create_string_buffer(n * 96)
check(L.pb200_srs_setup_from_secret(mont(0x1234567), mont(0x7654321), n, srs))
h = ctypes.c_void_p(); check(L.pb200_srs_upload(srs, n, ctypes.byref(h)))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    x = torch.randint(0, 2**62, (4 * n, 4), dtype=torch.int64, device="cuda")
    out = ctypes.create_string_buffer(96 * 4)
    for _ in range(2):
        check(L.pb200_msm_g1_dev(h, x.data_ptr(), n, 4, n, out, s.cuda_stream))
    N = 1 << 19
    a = torch.randint(0, 2**62, (5 * N, 4), dtype=torch.int64, device="cuda"); b = torch.empty_like(a)
    for _ in range(2):
        check(L.pb200_ntt_dev(a.data_ptr(), N // 8 + 3, b.data_ptr(), 19, 0, 1, 5, N, N, s.cuda_stream))
    s.synchronize()
print("done")
