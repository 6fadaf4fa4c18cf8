"""The opt-in batched-affine bucket accumulation (PB200_MSM_AFFINE=1: k_msm_affine_fwd / k_fp_batch_inverse /
k_msm_affine_back in csrc/msm.cu) ships in the library although the XYZZ kernels are the default (they measure
faster), so the default -m gpu suite keeps it under test: the switch is read once per
process, hence the MSM parity tests - oracle comparisons for small sizes, identity / repeated / negated bases,
all-equal and 0/1 scalars, skewed distributions, 2^16- and 2^20-point keys against a known discrete logarithm -
and the prover's golden-digest and Proof-bytes tests are re-run in a child process with it set."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("args", [
    ["tests/test_gpu_kernels.py", "-k", "msm and not large_points and not repeated_scalar_block"],
    ["tests/test_gpu_prover.py", "-k", "golden_digest or matches_cpu_oracle and not 2_16 and not 2_18 and not 2_20"],
], ids=["msm_parity", "prover_parity"])
def test_parity_with_batched_affine_accumulation(args):
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q"] + args, cwd=ROOT,
                       env=dict(os.environ, PB200_MSM_AFFINE="1"), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
