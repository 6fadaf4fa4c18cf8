"""tools/patches/*.patch (changes awaiting GPU validation) must keep applying to the tree and compiling."""
import glob
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCHES = sorted(glob.glob(os.path.join(ROOT, "tools", "patches", "*.patch")))


@pytest.mark.parametrize("patch", PATCHES, ids=[os.path.basename(p) for p in PATCHES])
def test_patch_applies_and_compiles(patch, tmp_path):
    if not shutil.which("nvcc") or not shutil.which("patch"):
        pytest.skip("needs nvcc and patch")
    dst = tmp_path / "plonk_b200" / "csrc"
    shutil.copytree(os.path.join(ROOT, "plonk_b200", "csrc"), dst)
    shutil.copytree(os.path.join(ROOT, "include"), tmp_path / "include")
    subprocess.check_call(["patch", "-p1", "--quiet", "-i", patch], cwd=tmp_path)
    touched = [l.split()[1][2:] for l in open(patch) if l.startswith("+++ b/")]
    for f in touched:
        if f.endswith(".cu"):
            subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                                   "-c", "-o", str(tmp_path / "out.o"), str(tmp_path / f)])
