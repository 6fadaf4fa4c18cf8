"""Level-1 drop-in in situ: the oracle's prover (a line-by-line restatement of Prover::prove_inner)
runs with ONLY the five reference functions of SURVEY.md section 8b swapped for the CUDA backend -
EvaluationDomain::{fft, ifft, coset_fft, coset_ifft} and CommitKey::commit - and must still produce
the reference's golden proof digest and the same bytes as the unpatched oracle."""
import hashlib
import random

import pytest

from oracle import pyref as R
from tests.util import bases_to_abi, from_abi, to_abi

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gpu_backend(monkeypatch):
    import plonk_b200
    from plonk_b200._lib import check, lib

    check(lib().pb200_init(0))
    calls = {"ntt": 0, "commit": 0}

    class GpuDomain(R.EvaluationDomain):
        def _gpu(self):
            return plonk_b200.EvaluationDomain(self.size)

        def fft(self, c):
            calls["ntt"] += 1
            return from_abi(self._gpu().fft(to_abi(list(c))))

        def ifft(self, e):
            calls["ntt"] += 1
            return from_abi(self._gpu().ifft(to_abi(list(e))))

        def coset_fft(self, c):
            calls["ntt"] += 1
            return from_abi(self._gpu().coset_fft(to_abi(list(c))))

        def coset_ifft(self, e):
            calls["ntt"] += 1
            return from_abi(self._gpu().coset_ifft(to_abi(list(e))))

    keys = {}

    def gpu_commit(powers_of_g, poly):
        calls["commit"] += 1
        k = id(powers_of_g)
        if k not in keys:
            keys[k] = plonk_b200.CommitKey(bases_to_abi(powers_of_g))
        try:
            return R.g1_from_raw_bytes(keys[k].commit(to_abi(list(poly))).raw)
        except plonk_b200.PolynomialDegreeTooLarge:
            raise ValueError("PolynomialDegreeTooLarge")

    monkeypatch.setattr(R, "EvaluationDomain", GpuDomain)
    monkeypatch.setattr(R, "commit", gpu_commit)
    return calls


def test_reference_kat_with_gpu_ntt_and_commit(gpu_backend):
    proof = R.kat_proof()
    assert hashlib.blake2b(proof).digest() == R.KAT_DIGEST
    # preprocessing: 15 ifft + 16 coset_fft + 4 fft, 15 commits; proving: 6 ifft + 6 coset_fft + 1 coset_ifft, 11 commits
    assert gpu_backend["ntt"] == 15 + 16 + 4 + 13 and gpu_backend["commit"] == 15 + 11


def test_synthetic_circuit_with_gpu_ntt_and_commit(gpu_backend, monkeypatch):
    rng = random.Random(12)
    pp = R.srs_from_secret(512 + 7, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, 400, seed=12, n_public=2, widgets=4)
    got = R.prove(R.compile_circuit(pp, b"level1", comp), R.StdRng.seed_from_u64(9), comp)
    monkeypatch.undo()  # back to the pure oracle
    want = R.prove(R.compile_circuit(pp, b"level1", comp), R.StdRng.seed_from_u64(9), comp)
    assert got == want
