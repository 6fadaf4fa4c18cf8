"""Single-stream proof latency for a given circuit size (development aid): python tools/prove_time.py 16 18 20"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonk_b200 import Prover
from plonk_b200._lib import check, lib
from plonk_b200.composer import synthetic_circuit
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
mont = lambda v: ((v << 256) % R_MOD).to_bytes(32, "little")
L = lib(); check(L.pb200_init(0))
for lg in (int(a) for a in sys.argv[1:] or ["16"]):
    n_srs = (1 << lg) + 7
    srs = ctypes.create_string_buffer(n_srs * 96)
    check(L.pb200_srs_setup_from_secret(mont(0x1234567), mont(0x7654321), n_srs, srs))
    t0 = time.time(); arrays = synthetic_circuit((1 << lg) - 6, seed=lg).arrays(); t_circ = time.time() - t0
    t0 = time.time(); p = Prover(b"t", arrays.constraints, arrays.selectors, arrays.wires, arrays.n_witnesses, srs.raw); t_new = time.time() - t0
    bl = b"".join(mont(i + 1) for i in range(14))
    for _ in range(2): p.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, bl)
    k = 5; t0 = time.time()
    for _ in range(k): p.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, bl)
    print(f"2^{lg} gates: prove {1e3*(time.time()-t0)/k:.2f} ms/proof (single stream, host witnesses), preprocessing {t_new:.2f} s, circuit build {t_circ:.1f} s", flush=True)
    del p
