"""Full-vector NTT parity at the sizes the Python oracle cannot reach (2^16, 2^19, 2^20, 2^23 = the quotient domain
of a 2^20-gate circuit in the reference's schedule): every output element of the CUDA transform against the C++
restatement of the reference's own schedule (oracle/cref.cpp: bit-reversal + DIT stages with running twiddles,
src/fft/domain.rs:383-463), which equals the Python oracle on every size that one reaches (tests/test_cref.py).
All four directions; a zero-padded input (the coset transform of n/8 + 3 coefficients, as round 3 of the prover
issues it) and complete inputs.  (tests/test_gpu_kernels.py::test_ntt_large_properties checks the same sizes through
round trips and single outputs against the definition of the DFT.)"""
import random

import pytest

from oracle import cref
from tests.util import rand_fr, to_abi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_n", [16, 19, 20, 23])
def test_ntt_full_vectors_match_the_restated_reference(log_n):
    import plonk_b200
    from plonk_b200._lib import check, lib

    check(lib().pb200_init(0))
    rng = random.Random(100 + log_n)
    n = 1 << log_n
    dom = plonk_b200.EvaluationDomain(n)
    xb = to_abi(rand_fr(rng, n // 8 + 3))
    ev = dom.coset_fft(xb)
    assert ev == cref.ntt(xb, log_n, 0, 1)
    assert dom.coset_ifft(ev) == cref.ntt(ev, log_n, 1, 1)
    assert dom.fft(ev) == cref.ntt(ev, log_n, 0, 0)
    assert dom.ifft(ev) == cref.ntt(ev, log_n, 1, 0)
    assert dom.coset_fft(ev) == cref.ntt(ev, log_n, 0, 1)
