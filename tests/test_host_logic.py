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

    from plonk_b200._lib import COMPOSER_EXPORTS, EXPORTS, LIB_PATH

    L = ctypes.CDLL(LIB_PATH)
    for name, exports in (("plonk_b200.h", EXPORTS), ("plonk_b200_composer.h", COMPOSER_EXPORTS)):
        header = open(os.path.join(os.path.dirname(LIB_PATH), "..", "include", name)).read()
        header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)  # prototypes only, not the prose
        declared = set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", header))
        assert declared == set(exports), name
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


def _build_cpp(name: str):
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    exe = os.path.join(here, "cpp", name)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, os.path.join(here, "cpp", name + ".cpp"),
                           "-L" + os.path.join(root, "plonk_b200"), "-lplonk_b200",
                           "-Wl,-rpath," + os.path.join(root, "plonk_b200")])
    return exe


def _build_api_check():
    return _build_cpp("api_check")


def test_cpp_mirror_composer_builds_the_reference_bench_circuit():
    """benches/plonk.rs written against include/plonk_b200.hpp's Composer, gadget by gadget: the
    exported arrays equal the oracle composer's (no GPU involved)."""
    import subprocess

    from oracle import cref, gadgets

    comp = gadgets.GadgetComposer.initialized()
    gadgets.bench_circuit(comp, 1 << 13)
    a = cref.CircuitArrays(comp)
    h = 0xCBF29CE484222325
    for blob in (a.selectors, a.wires, a.witnesses):
        for byte in blob:
            h = ((h ^ byte) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    out = subprocess.run([_build_cpp("bench_circuit"), "export", str(1 << 13)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split() == [str(a.constraints), str(a.n_witnesses), "%016x" % h]


def test_cpp_mirror_header_compiles_and_links():
    """include/plonk_b200.hpp (the compiled-language mirror of the reference interface) against the .so."""
    assert os.path.exists(_build_api_check())
