"""CPU tests of the host-side product logic: composer mirror, ABI surface, domain sizing."""
import os
import re

from oracle import cref
from oracle import pyref as R


def test_product_composer_matches_oracle_composer():
    from plonk_b200.composer import synthetic_circuit

    ours = synthetic_circuit(300, seed=5, n_public=3).arrays()
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, 300, seed=5, n_public=3)
    ref = cref.CircuitArrays(comp)
    for f in ("constraints", "selectors", "wires", "witnesses", "pi_idx", "pi_vals", "n_witnesses", "n_pi"):
        assert getattr(ours, f) == getattr(ref, f), f
    # with rows of every gate family
    ours = synthetic_circuit(200, seed=9, n_public=1, widgets=6).arrays()
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, 200, seed=9, n_public=1, widgets=6)
    ref = cref.CircuitArrays(comp)
    for f in ("constraints", "selectors", "wires", "witnesses", "pi_idx", "pi_vals"):
        assert getattr(ours, f) == getattr(ref, f), f


def test_library_exports_every_declared_symbol():
    import ctypes

    from plonk_b200._lib import EXPORTS, LIB_PATH

    header = open(os.path.join(os.path.dirname(LIB_PATH), "..", "include", "plonk_b200.h")).read()
    declared = set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(EXPORTS)
    L = ctypes.CDLL(LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), sym


def test_domain_sizes_and_errors():
    from plonk_b200.domain import EvaluationDomain, InvalidEvalDomainSize

    assert [EvaluationDomain(k).size for k in (0, 1, 2, 3, 5, 8, 9, 4097)] == [1, 1, 2, 4, 8, 8, 16, 8192]
    import pytest

    with pytest.raises(InvalidEvalDomainSize):
        EvaluationDomain(1 << 32)  # domain.rs:132-137


def test_no_cuda_device_fails_loudly():
    """No CPU fallback: without a GPU the C ABI reports PB200_ERR_CUDA instead of computing."""
    import ctypes

    import torch

    if torch.cuda.is_available():
        return
    from plonk_b200._lib import lib

    out = ctypes.create_string_buffer(64)
    assert lib().pb200_ntt(bytes(64), 2, out, 1, 0, 0, 1, 2, 2) == -1


def _build_api_check():
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    exe = os.path.join(here, "cpp", "api_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(here, "cpp", "api_check.cpp"),
                           "-L" + os.path.join(root, "plonk_b200"), "-lplonk_b200",
                           "-Wl,-rpath," + os.path.join(root, "plonk_b200")])
    return exe


def test_cpp_mirror_header_compiles_and_links():
    """include/plonk_b200.hpp (the compiled-language mirror of the reference interface) against the .so."""
    assert os.path.exists(_build_api_check())
