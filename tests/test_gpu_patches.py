"""GPU parity for the switches that tools/patches/*.patch add (PB200_QUOT4N: round 3 on the 4n coset;
PB200_LAGRANGE: wire commitments in the Lagrange basis).

Each case is skipped until its patch is applied and the library rebuilt: the variable's name is then a
string in the library.  The switches are read once per process, so the prover parity tests (golden
digest, Proof bytes == CPU oracle, every gate family, the reference's BenchCircuit, CircuitUnsatisfied on
a bad witness) are re-run in a child process with the variable set."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("var,patch", [("PB200_QUOT4N", "quot4n_prover.patch"), ("PB200_LAGRANGE", "lagrange_wires_prover.patch")])
def test_prover_parity_with_switch(var, patch):
    from plonk_b200._lib import LIB_PATH

    if var.encode() not in open(LIB_PATH, "rb").read():
        pytest.skip(f"tools/patches/{patch} is not applied in this build")
    env = dict(os.environ, **{var: "1"})
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_prover.py", "tests/test_gpu_gadget_circuits.py", "-m", "gpu", "-x", "-q",
                        "-k", "not 2_18 and not 2_20 and not cpp_mirror"], cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
