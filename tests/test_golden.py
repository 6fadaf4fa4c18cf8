"""Committed golden vectors (tests/golden/vectors.json): consistency with the oracle on the CPU,
and the CUDA path against them on the GPU."""
import hashlib
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = json.load(open(os.path.join(HERE, "golden", "vectors.json")))


def test_golden_file_is_what_the_oracle_produces():
    from tests.golden.make_golden import build

    assert build() == VEC


def test_golden_kat_proof_matches_reference_digest():
    # the digest literal of reference src/compiler/prover.rs:1151-1158
    assert hashlib.blake2b(bytes.fromhex(VEC["kat_proof_hex"])).hexdigest() == (
        "e8564ec22d8cc0ba603626025da3755077aaf0323261908dab68d694736fc273"
        "d31e256cbd3a6a21e7ade63191ac5c9d44a113ac4989a52e4be3abeb1d333237"
    )


@pytest.mark.gpu
def test_gpu_against_golden_vectors():
    import plonk_b200
    from plonk_b200._lib import check, lib

    check(lib().pb200_init(0))
    n = VEC["ntt"]
    dom = plonk_b200.EvaluationDomain(1 << n["log_n"])
    x = bytes.fromhex(n["input"])
    assert dom.fft(x).hex() == n["fft"]
    assert dom.ifft(x).hex() == n["ifft"]
    assert dom.coset_fft(x).hex() == n["coset_fft"]
    assert dom.coset_ifft(x).hex() == n["coset_ifft"]
    m = VEC["msm"]
    key = plonk_b200.CommitKey(bytes.fromhex(m["bases"]))
    c = key.commit(bytes.fromhex(m["scalars"]))
    assert c.raw.hex() == m["result_raw"] and c.to_bytes().hex() == m["result_compressed"]
