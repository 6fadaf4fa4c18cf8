"""Proofs against the reference Verifier's equation (oracle/verify.py: Proof::verify with the final
pairing replaced by the equivalent G1 identity under the known test-SRS secret).  CPU only: the proofs
come from the restated CPU provers, whose bytes the GPU prover is tested to reproduce; the GPU suite
repeats the check on GPU-made proofs (tests/test_gpu_gadget_circuits.py)."""
import random

import pytest

from oracle import cref
from oracle import gadgets as G
from oracle import pyref as R
from oracle import verify as V
from tests.test_gpu_gadget_circuits import CASES


def test_reference_golden_proof_satisfies_the_verifier_equation():
    # the KAT of src/compiler/prover.rs:1132-1162: its SRS secret is the first draw of the seeded RNG (srs.rs:74-77)
    x = R.random_nonzero_bls_scalar(R.StdRng.seed_from_u64(0x9235E700))
    pp = R.srs_setup(1 << 10, R.StdRng.seed_from_u64(0x9235E700), keep=64)
    comp = R.Composer.initialized()
    R.minimal_circuit(comp)
    pd = R.compile_circuit(pp, b"proof-compatibility", comp)
    proof = R.kat_proof()
    args = (b"proof-compatibility", len(comp.constraints), pd.comms, comp.public_input_indexes(), comp.public_inputs_vec(), pp[0])
    assert V.verify_with_secret(proof, *args, x)
    assert not V.verify_with_secret(proof, *args, x + 1)
    for pos in (5, 48 * 4 + 7, 48 * 9 + 1, 528 + 3, 528 + 32 * 7, 1007):  # a commitment, z, W_z, evaluations
        bad = bytearray(proof)
        bad[pos] ^= 1
        try:
            ok = V.verify_with_secret(bytes(bad), *args, x)
        except AssertionError:  # the flipped bit left the curve or the canonical range: rejected at decoding
            ok = False
        assert not ok, pos


def _cref_case(label, comp, x=0x1234567, gs=0x7654321):
    arr = cref.CircuitArrays(comp)
    n = 1 << (arr.constraints + 6 - 1).bit_length()
    srs = cref.srs_from_secret(n + 7, x, gs)
    prover = cref.CrefProver(label, arr, srs)
    comms = {k: R.g1_decompress(c) for k, c in zip(R.POLY_NAMES, prover.commitments())}
    return arr, prover, comms, R.g1_from_raw_bytes(srs[:96]), x


def test_synthetic_circuit_with_public_inputs_and_all_widgets_verifies():
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, 500, seed=77, n_public=3, widgets=9)
    arr, prover, comms, g, x = _cref_case(b"verify-synthetic", comp)
    proof = prover.prove(cref.draw_blinders(R.StdRng.seed_from_u64(1)))
    idx, vals = comp.public_input_indexes(), comp.public_inputs_vec()
    assert V.verify_with_secret(proof, b"verify-synthetic", arr.constraints, comms, idx, vals, g, x)
    assert not V.verify_with_secret(proof, b"verify-synthetic", arr.constraints, comms, idx, [vals[0] + 1] + vals[1:], g, x)
    assert not V.verify_with_secret(proof, b"other-label", arr.constraints, comms, idx, vals, g, x)


def test_reference_bench_circuit_proof_verifies():
    comp = G.GadgetComposer.initialized()
    G.bench_circuit(comp, 1 << 12)
    arr, prover, comms, g, x = _cref_case(b"dusk-network", comp)
    proof = prover.prove(cref.draw_blinders(R.StdRng.seed_from_u64(2)))
    assert V.verify_with_secret(proof, b"dusk-network", arr.constraints, comms, [], [], g, x)


@pytest.mark.parametrize("name,build,default,satisfied,unsatisfied", CASES, ids=[c[0] for c in CASES])
def test_gadget_circuit_proofs_verify(name, build, default, satisfied, unsatisfied):
    """Every satisfying assignment of the gadget table verifies against the key compiled from the default one."""
    comp = G.GadgetComposer.initialized()
    build(comp, *default)
    arr, prover, comms, g, x = _cref_case(name.encode(), comp, x=0x5EED + len(name), gs=0xACE)
    for k, vals in enumerate([default] + satisfied):
        c = G.GadgetComposer.initialized()
        build(c, *vals)
        proof = prover.prove(cref.draw_blinders(R.StdRng.seed_from_u64(300 + k)), cref.CircuitArrays(c))
        assert V.verify_with_secret(proof, name.encode(), arr.constraints, comms, c.public_input_indexes(), c.public_inputs_vec(), g, x), vals
