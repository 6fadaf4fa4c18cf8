"""Regenerates tests/golden/vectors.json from the pinned Python oracle (oracle/pyref.py).

The reference is Rust and cannot run here, so these are not outputs of the reference binary; they are
outputs of the restatement that reproduces the reference's golden proof digest
(src/compiler/prover.rs:1151-1158).  The first entry *is* pinned by the reference: the 1008 proof
bytes whose blake2b-512 equals that digest.

  python tests/golden/make_golden.py        # rewrites vectors.json
"""
import hashlib
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref as R  # noqa: E402


def build():
    rng = random.Random(0xB200)
    out = {}
    proof = R.kat_proof()
    assert hashlib.blake2b(proof).digest() == R.KAT_DIGEST
    out["kat_proof_hex"] = proof.hex()
    out["kat_digest_hex"] = R.KAT_DIGEST.hex()
    # NTT: n = 16, input of 11 coefficients (zero padded), all four directions; ABI byte layout
    x = [rng.randrange(R.R_MOD) for _ in range(11)]
    d = R.EvaluationDomain(16)
    out["ntt"] = {
        "log_n": 4,
        "input": R.fr_vec_to_mont_bytes(x).hex(),
        "fft": R.fr_vec_to_mont_bytes(d.fft(x)).hex(),
        "ifft": R.fr_vec_to_mont_bytes(d.ifft(x)).hex(),
        "coset_fft": R.fr_vec_to_mont_bytes(d.coset_fft(x)).hex(),
        "coset_ifft": R.fr_vec_to_mont_bytes(d.coset_ifft(x)).hex(),
    }
    # MSM: 12 bases [s^i] g (SRS shape), scalars incl. 0, 1, r-1; affine result and its compression
    pts = R.srs_from_secret(12, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    sc = [0, 1, R.R_MOD - 1] + [rng.randrange(R.R_MOD) for _ in range(9)]
    res = R.jac_to_affine(R.msm_naive(pts, sc))
    out["msm"] = {
        "bases": b"".join(R.g1_to_raw_bytes(p) for p in pts).hex(),
        "scalars": R.fr_vec_to_mont_bytes(sc).hex(),
        "result_raw": R.g1_to_raw_bytes(res).hex(),
        "result_compressed": R.g1_compress(res).hex(),
    }
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json")
    json.dump(build(), open(path, "w"), indent=1)
    print("wrote", path)
