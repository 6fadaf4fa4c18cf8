"""The reference's serialized key formats read by the product's loaders:
CommitKey::to_raw_var_bytes / PublicParameters::to_raw_var_bytes (key.rs:215-298, srs.rs:114-146) and
Prover::to_bytes (prover.rs:236-350).  The byte strings are produced by oracle/serialize.py from the oracle's
own keys (the reference holds no golden bytes for these formats)."""
import ctypes
import random

import pytest

from oracle import cref
from oracle import pyref as R
from oracle import serialize as S
from tests.util import bases_to_abi, progression_bases


def _parse(blob, checked):
    from plonk_b200._lib import lib

    n = ctypes.c_size_t()
    rc = lib().pb200_raw_commit_key_points(blob, len(blob), checked, ctypes.byref(n))
    return rc, n.value


def test_raw_commit_key_parsing_without_a_gpu():
    """from_slice_unchecked semantics need no device: record conversion, identity flag, truncated input."""
    from plonk_b200._lib import check, lib

    pts = progression_bases(9, 3, 5)
    pts[4] = None
    blob = S.commit_key_to_raw_var_bytes(pts)
    assert len(blob) == 8 + 9 * 97
    assert _parse(blob, 0) == (0, 9) and _parse(blob, 1) == (0, 9)
    out = ctypes.create_string_buffer(96 * 9)
    check(lib().pb200_commit_key_from_raw_var_bytes(blob, len(blob), 0, out))
    assert out.raw == bases_to_abi(pts)  # the identity record (x = 0, y = 1, flag 1) becomes zeros
    # unchecked: as many whole records as there are, at most the announced count (chunks_exact + zip)
    assert _parse(blob[:-1], 0) == (0, 8) and _parse(blob + bytes(97), 0) == (0, 9)
    assert _parse((3).to_bytes(8, "little") + blob[8:], 0) == (0, 3)
    # checked (from_raw_var_bytes): exact length, non-empty
    assert _parse(blob[:-1], 1)[0] == -4 and _parse(blob + b"\0", 1)[0] == -4 and _parse(blob[:7], 1)[0] == -4
    assert _parse((0).to_bytes(8, "little"), 1)[0] == -10
    # PublicParameters::to_raw_var_bytes = 240 bytes of opening key in front
    pp_blob = bytes(S.OPENING_KEY_BYTES) + blob
    assert _parse(pp_blob[240:], 1) == (0, 9)


@pytest.mark.gpu
def test_checked_raw_commit_key_validates_points_on_the_gpu():
    from plonk_b200 import kzg
    from plonk_b200._lib import check, lib

    check(lib().pb200_init(0))
    pts = progression_bases(40, 7, 11) + [None]
    blob = S.commit_key_to_raw_var_bytes(pts)
    key = kzg.CommitKey.from_raw_var_bytes(blob)
    rng = random.Random(1)
    coeffs = [rng.randrange(R.R_MOD) for _ in range(41)]
    assert R.g1_from_raw_bytes(key.commit(R.fr_vec_to_mont_bytes(coeffs)).raw) == R.jac_to_affine(R.msm_naive(pts, coeffs))
    # a point off the curve, and one on the curve outside the prime-order subgroup
    off_curve = (pts[3][0], (pts[3][1] + 1) % R.P_MOD)
    x = 2
    while True:
        y2 = (x ** 3 + 4) % R.P_MOD
        y = pow(y2, (R.P_MOD + 1) // 4, R.P_MOD)
        if y * y % R.P_MOD == y2 and R.jac_to_affine(R.jac_mul(R.jac_from_affine((x, y)), R.R_MOD)) is not None:
            break
        x += 1
    for idx, bad in ((3, off_curve), (17, (x, y))):
        rec = R.g1_to_raw_bytes(bad) + b"\0"
        broken = blob[: 8 + 97 * idx] + rec + blob[8 + 97 * (idx + 1) :]
        with pytest.raises(kzg.PointMalformed) as e:
            kzg.CommitKey.from_raw_var_bytes(broken)
        assert f"point {idx}" in str(e.value)
        assert kzg.CommitKey.from_slice_unchecked(broken).n_points == 41  # the unchecked loader does not look


@pytest.mark.gpu
def test_prover_from_bytes_proves_like_the_compiled_prover():
    """Prover::to_bytes -> pb200_prover_from_bytes: same commitments, same Proof bytes as the prover compiled
    from the circuit (and as the CPU oracle), with rows of every gate family."""
    import plonk_b200
    from plonk_b200._lib import check, lib

    check(lib().pb200_init(0))
    rng = random.Random(9)
    pp = R.srs_from_secret(512 + 7, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, 300, seed=77, n_public=2, widgets=5)
    pd = R.compile_circuit(pp, b"serialized", comp)
    blob = S.prover_to_bytes(pd)
    arrays = cref.CircuitArrays(comp)
    loaded = plonk_b200.Prover.from_bytes(blob, arrays.wires, arrays.n_witnesses)
    compiled = plonk_b200.Prover(b"serialized", arrays.constraints, arrays.selectors, arrays.wires, arrays.n_witnesses, bases_to_abi(pp))
    assert loaded.commitments() == compiled.commitments()
    bl = cref.draw_blinders(R.StdRng.seed_from_u64(3))
    proof = loaded.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, bl)
    assert proof == compiled.prove(arrays.witnesses, arrays.pi_idx, arrays.pi_vals, bl)
    assert proof == cref.CrefProver(b"serialized", arrays, bases_to_abi(pp)).prove(bl)
    # malformed inputs, as Prover::try_from_bytes reports them
    L = lib()
    h = ctypes.c_void_p()

    def rc(b):
        return L.pb200_prover_from_bytes(b, len(b), arrays.wires, arrays.n_witnesses, ctypes.byref(h))

    assert rc(blob[:40]) == -4 and rc(blob[:-1000]) == -4                       # NotEnoughBytes
    assert rc(blob[:32] + (pd.size * 2).to_bytes(8, "big") + blob[40:]) == -10  # size != next_pow2(constraints)
    head = 48 + len(pd.label) + 16 + 8
    non_canonical = blob[:head] + (R.R_MOD + 1).to_bytes(32, "little") + blob[head + 32 :]
    assert rc(non_canonical) == -10                                             # a coefficient >= r: InvalidData
