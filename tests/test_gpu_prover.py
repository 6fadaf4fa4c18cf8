"""GPU prover parity: Proof bytes from the device-resident prover vs the reference's golden digest
and vs the CPU oracles on seeded circuits."""
import hashlib
import random

import pytest

from oracle import cref
from oracle import pyref as R
from tests.util import bases_to_abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import plonk_b200
    from plonk_b200._lib import check, lib

    check(lib().pb200_init(0))
    return plonk_b200


def _gpu_prover(pb, label, arrays, srs_raw):
    return pb.Prover(label, arrays.constraints, arrays.selectors, arrays.wires, arrays.n_witnesses, srs_raw)


def test_gpu_prover_reproduces_reference_golden_digest(pb):
    # reference src/compiler/prover.rs:1132-1162
    pp = R.srs_setup(1 << 10, R.StdRng.seed_from_u64(0x9235E700), keep=64)
    comp = R.Composer.initialized()
    R.minimal_circuit(comp)
    arrays = cref.CircuitArrays(comp)
    prover = _gpu_prover(pb, b"proof-compatibility", arrays, bases_to_abi(pp))
    pd = R.compile_circuit(pp, b"proof-compatibility", comp)
    assert prover.commitments() == [R.g1_compress(pd.comms[k]) for k in R.POLY_NAMES]
    blinders = cref.draw_blinders(R.StdRng.seed_from_u64(0x9235E701))
    proof = prover.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, blinders)
    assert hashlib.blake2b(proof).digest() == R.KAT_DIGEST


@pytest.mark.parametrize("log_gates,n_public", [(5, 2), (7, 0), (10, 3), (12, 2)])
def test_gpu_prover_matches_cpu_oracle(pb, log_gates, n_public):
    rng = random.Random(log_gates)
    n_gates = (1 << log_gates) - 6 if log_gates != 7 else (1 << 7)  # also a circuit that fills its domain exactly
    srs_raw = cref.srs_from_secret((1 << (log_gates + 1)) + 7, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, n_gates, seed=1000 + log_gates, n_public=n_public)
    arrays = cref.CircuitArrays(comp)
    label = b"synthetic-%d" % log_gates
    cpu = cref.CrefProver(label, arrays, srs_raw)
    gpu = _gpu_prover(pb, label, arrays, srs_raw)
    assert gpu.commitments() == cpu.commitments()
    for seed in (1, 2):
        blinders = cref.draw_blinders(R.StdRng.seed_from_u64(seed))
        assert gpu.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, blinders) == cpu.prove(blinders)
    # a wrong witness must be rejected like the reference does (Error::CircuitUnsatisfied)
    bad = bytearray(arrays.witnesses)
    bad[32 * 9] ^= 1
    with pytest.raises(pb.CircuitUnsatisfied):
        gpu.prove(bytes(bad), arrays.pi_idx, arrays.pi_vals, cref.draw_blinders(R.StdRng.seed_from_u64(3)))


@pytest.mark.parametrize("log_gates,widgets", [(8, 6), (11, 40)])
def test_gpu_prover_all_gate_families(pb, log_gates, widgets):
    """Circuits with satisfied range / logic / fixed-base / curve-addition rows: every widget of the
    quotient kernel and of the linearisation runs with a non-zero selector polynomial."""
    rng = random.Random(100 + log_gates)
    n_gates = (1 << log_gates) - 6
    srs_raw = cref.srs_from_secret((1 << log_gates) + 7, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, n_gates, seed=500 + log_gates, n_public=2, widgets=widgets)
    arrays = cref.CircuitArrays(comp)
    cpu = cref.CrefProver(b"widgets", arrays, srs_raw)
    gpu = _gpu_prover(pb, b"widgets", arrays, srs_raw)
    assert gpu.commitments() == cpu.commitments()
    assert all(c[0] & 0x40 == 0 for c in gpu.commitments()[:11])  # no selector commits to the identity
    blinders = cref.draw_blinders(R.StdRng.seed_from_u64(77))
    assert gpu.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, blinders) == cpu.prove(blinders)
    bad = bytearray(arrays.witnesses)
    bad[32 * 10] ^= 1  # breaks a range row
    with pytest.raises(pb.CircuitUnsatisfied):
        gpu.prove(bytes(bad), arrays.pi_idx, arrays.pi_vals, blinders)


@pytest.mark.parametrize("degree", [1 << 12, 1 << 13, 1 << 16])  # 2^12: BASELINE.json configs[0]
def test_gpu_prover_reference_bench_circuit(pb, degree):
    """benches/plonk.rs BenchCircuit<DEGREE> (the circuit BASELINE.json's metric is quoted on): built
    by the product's native composer and proved on the GPU vs built by the oracle's composer and
    proved by the CPU restatement.  2^16 is the headline configuration (64129 gates, n = 2^16)."""
    from oracle import gadgets as G
    from plonk_b200 import gadgets as N

    comp = G.GadgetComposer.initialized()
    G.bench_circuit(comp, degree)
    oracle_arrays = cref.CircuitArrays(comp)
    arrays = N.bench_circuit(degree).arrays()
    assert arrays.constraints == oracle_arrays.constraints and arrays.witnesses == oracle_arrays.witnesses
    n = 1 << (arrays.constraints - 1).bit_length()
    rng = random.Random(degree)
    srs_raw = cref.srs_from_secret(n + 7, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    cpu = cref.CrefProver(b"dusk-network", oracle_arrays, srs_raw)
    gpu = _gpu_prover(pb, b"dusk-network", arrays, srs_raw)
    assert gpu.commitments() == cpu.commitments()
    blinders = cref.draw_blinders(R.StdRng.seed_from_u64(degree))
    assert gpu.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, blinders) == cpu.prove(blinders)
    bad = bytearray(arrays.witnesses)
    bad[32 * 8] ^= 2  # w_a: breaks the logic / range / decomposition rows
    with pytest.raises(pb.CircuitUnsatisfied):
        gpu.prove(bytes(bad), arrays.pi_idx, arrays.pi_vals, blinders)


def test_gpu_prover_2_16_gates_matches_cpu_oracle(pb):
    """BASELINE.json configs[1]: 2^16-gate circuit, Proof bytes == CPU restatement."""
    n_gates = (1 << 16) - 6
    srs_raw = cref.srs_from_secret((1 << 16) + 7, 0x1234567, 0x7654321)
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, n_gates, seed=16)
    arrays = cref.CircuitArrays(comp)
    cpu = cref.CrefProver(b"bench-2^16", arrays, srs_raw)
    gpu = _gpu_prover(pb, b"bench-2^16", arrays, srs_raw)
    assert gpu.commitments() == cpu.commitments()
    blinders = cref.draw_blinders(R.StdRng.seed_from_u64(16))
    assert gpu.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, blinders) == cpu.prove(blinders)


def test_gpu_prover_2_18_gates_matches_cpu_oracle(pb):
    """A larger domain (n = 2^18, quotient domain 2^21: three-pass NTT plan, bigger MSM tables)."""
    log_gates = 18
    n_gates = (1 << log_gates) - 6
    from plonk_b200._lib import check, lib
    import ctypes

    n_srs = (1 << log_gates) + 7
    raw = ctypes.create_string_buffer(96 * n_srs)
    check(lib().pb200_srs_setup_from_secret(R.fr_to_mont_bytes(0xABCDEF), R.fr_to_mont_bytes(0x13579), n_srs, raw))
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, n_gates, seed=18, widgets=3)
    arrays = cref.CircuitArrays(comp)
    cpu = cref.CrefProver(b"bench-2^18", arrays, raw.raw)
    gpu = _gpu_prover(pb, b"bench-2^18", arrays, raw.raw)
    assert gpu.commitments() == cpu.commitments()
    blinders = cref.draw_blinders(R.StdRng.seed_from_u64(18))
    assert gpu.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, blinders) == cpu.prove(blinders)


def test_gpu_prover_concurrent_threads_are_deterministic(pb):
    """The C ABI is called concurrently from several host threads (as rayon workers would): every
    thread must get the same bytes as a lone call."""
    from concurrent.futures import ThreadPoolExecutor

    rng = random.Random(4)
    srs_raw = cref.srs_from_secret((1 << 11) + 7, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, (1 << 10) - 6, seed=41, widgets=4)
    arrays = cref.CircuitArrays(comp)
    gpu = _gpu_prover(pb, b"threads", arrays, srs_raw)
    blinders = [cref.draw_blinders(R.StdRng.seed_from_u64(s)) for s in range(6)]
    alone = [gpu.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, b) for b in blinders]
    with ThreadPoolExecutor(6) as ex:
        for _ in range(3):
            got = list(ex.map(lambda b: gpu.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, b), blinders))
            assert got == alone
    assert len(set(alone)) == 6  # different blinders, different proofs


def test_gpu_prover_rejects_bad_arguments(pb):
    rng = random.Random(5)
    srs_raw = cref.srs_from_secret(64 + 7, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, 58, seed=5)
    arrays = cref.CircuitArrays(comp)
    with pytest.raises(pb.Pb200Error):  # commit key too small for the circuit (TruncatedDegreeTooLarge)
        pb.Prover(b"x", arrays.constraints, arrays.selectors, arrays.wires, arrays.n_witnesses, srs_raw[: 96 * 40])
    gpu = _gpu_prover(pb, b"x", arrays, srs_raw)
    from plonk_b200._lib import lib
    import ctypes

    out = ctypes.create_string_buffer(1008)
    bl = cref.draw_blinders(R.StdRng.seed_from_u64(1))
    # wrong witness count
    assert lib().pb200_prove(gpu._h, arrays.witnesses, arrays.n_witnesses - 1, arrays.pi_idx, arrays.pi_vals, arrays.n_pi, bl, out) == -4
    # public input position outside the circuit
    bad_idx = (10**6).to_bytes(8, "little") + arrays.pi_idx[8:]
    assert lib().pb200_prove(gpu._h, arrays.witnesses, arrays.n_witnesses, bad_idx, arrays.pi_vals, arrays.n_pi, bl, out) == -4
    # public inputs announced but not given; positions not strictly increasing (the reference keeps them in a BTreeMap)
    assert arrays.n_pi >= 2
    assert lib().pb200_prove(gpu._h, arrays.witnesses, arrays.n_witnesses, None, arrays.pi_vals, arrays.n_pi, bl, out) == -4
    assert lib().pb200_prove(gpu._h, arrays.witnesses, arrays.n_witnesses, arrays.pi_idx, None, arrays.n_pi, bl, out) == -4
    dup_idx = arrays.pi_idx[:8] + arrays.pi_idx[:8] + arrays.pi_idx[16:]
    assert lib().pb200_prove(gpu._h, arrays.witnesses, arrays.n_witnesses, dup_idx, arrays.pi_vals, arrays.n_pi, bl, out) == -4
    swapped = arrays.pi_idx[8:16] + arrays.pi_idx[:8] + arrays.pi_idx[16:]
    assert lib().pb200_prove(gpu._h, arrays.witnesses, arrays.n_witnesses, swapped, arrays.pi_vals, arrays.n_pi, bl, out) == -4
    # device-resident witnesses: the table length is checked too
    import torch

    d_wit = torch.frombuffer(bytearray(arrays.witnesses), dtype=torch.uint8).cuda()
    assert lib().pb200_prove_dev(gpu._h, d_wit.data_ptr(), arrays.n_witnesses + 1, arrays.pi_idx, arrays.pi_vals, arrays.n_pi, bl, out, None) == -4
    assert lib().pb200_prove_dev(gpu._h, d_wit.data_ptr(), arrays.n_witnesses, arrays.pi_idx, arrays.pi_vals, arrays.n_pi, bl, out, None) == 0
    assert out.raw == gpu.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, bl)


def test_gpu_prover_2_20_gates_matches_cpu_oracle(pb):
    """BASELINE.json configs[2]: 2^20-gate circuit (quotient domain 2^23 in the reference's schedule, 2^22 on
    the 4n coset; 2^20-point commit key).  About 1.5 minutes, nearly all of it the CPU oracle's proof."""
    from plonk_b200._lib import check, lib
    from plonk_b200.composer import synthetic_circuit
    import ctypes

    log_gates = 20
    n_srs = (1 << log_gates) + 7
    raw = ctypes.create_string_buffer(96 * n_srs)
    check(lib().pb200_srs_setup_from_secret(R.fr_to_mont_bytes(0xABCDEF), R.fr_to_mont_bytes(0x13579), n_srs, raw))
    arrays = synthetic_circuit((1 << log_gates) - 6, seed=20, widgets=2).arrays()
    cpu = cref.CrefProver(b"bench-2^20", arrays, raw.raw)
    gpu = _gpu_prover(pb, b"bench-2^20", arrays, raw.raw)
    assert gpu.commitments() == cpu.commitments()
    blinders = cref.draw_blinders(R.StdRng.seed_from_u64(20))
    assert gpu.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, blinders) == cpu.prove(blinders)


def test_cpp_mirror_produces_the_same_proof(pb, tmp_path):
    """include/plonk_b200.hpp (EvaluationDomain / CommitKey / Prover in C++) end to end."""
    import struct
    import subprocess

    from tests.test_host_logic import _build_api_check

    rng = random.Random(6)
    srs_raw = cref.srs_from_secret(256 + 7, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, 250, seed=61, n_public=2, widgets=5)
    a = cref.CircuitArrays(comp)
    bl = cref.draw_blinders(R.StdRng.seed_from_u64(6))
    label = b"cpp-mirror"
    blob = struct.pack("<5Q", len(srs_raw) // 96, a.constraints, a.n_witnesses, a.n_pi, len(label))
    blob += srs_raw + a.selectors + a.wires + a.witnesses + a.pi_idx + a.pi_vals + bl + label
    f = tmp_path / "case.bin"
    f.write_bytes(blob)
    out = subprocess.run([_build_api_check(), str(f)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip() == cref.CrefProver(label, a, srs_raw).prove(bl).hex()


def test_cpp_mirror_proves_the_reference_bench_circuit(pb, tmp_path):
    """benches/plonk.rs through the C++ mirror only (Composer -> Prover::prove in C++): the proof equals
    the CPU restatement's on the oracle-built circuit."""
    import struct
    import subprocess

    from oracle import gadgets as G
    from tests.test_host_logic import _build_cpp

    comp = G.GadgetComposer.initialized()
    G.bench_circuit(comp, 32)
    a = cref.CircuitArrays(comp)
    rng = random.Random(8)
    srs_raw = cref.srs_from_secret(4096 + 7, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    bl = cref.draw_blinders(R.StdRng.seed_from_u64(8))
    f = tmp_path / "case.bin"
    f.write_bytes(struct.pack("<Q", len(srs_raw) // 96) + srs_raw + bl)
    out = subprocess.run([_build_cpp("bench_circuit"), "prove", "32", str(f)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip() == cref.CrefProver(b"dusk-network", a, srs_raw).prove(bl).hex()
