"""The prover's two schedule switches, both ON by default since round 2 (validated on a B200 in
profiles/ab_patches_r02/): PB200_QUOT4N (round 3 on the 4n coset instead of the reference's 8n one,
quotient_poly.rs:50-137) and PB200_LAGRANGE (wire commitments through the wire values and the Lagrange form
of the commit key instead of CommitKey::commit of the interpolated polynomials, prover.rs:187-210).

The default -m gpu suite therefore exercises the new paths everywhere; here the reference's own schedules
are kept under test: the switches are read once per process, so the prover parity tests (golden digest,
Proof bytes == CPU oracle, every gate family, the reference's BenchCircuit, CircuitUnsatisfied on a bad
witness) are re-run in a child process with a switch set to 0.  Proof bytes must not depend on either."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"PB200_QUOT4N": "0"}, {"PB200_LAGRANGE": "0"}, {"PB200_QUOT4N": "0", "PB200_LAGRANGE": "0"}],
                         ids=["quotient_8n", "monomial_wire_commitments", "both_reference_schedules"])
def test_prover_parity_with_reference_schedule(env):
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_prover.py", "tests/test_gpu_gadget_circuits.py", "-m", "gpu", "-x", "-q",
                        "-k", "not 2_18 and not 2_20 and not cpp_mirror"], cwd=ROOT, env=dict(os.environ, **env), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
