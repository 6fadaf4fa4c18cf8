"""GPU parity tests (run on a B200 via gpurun): CUDA kernels behind the C ABI vs the Python oracle.

Bit-exact everywhere: all arithmetic is integer/modular."""
import ctypes
import random

import pytest

from oracle import pyref as R
from tests.util import bases_to_abi, from_abi, progression_bases, rand_fr, to_abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import plonk_b200
    from plonk_b200._lib import check, lib

    check(lib().pb200_init(0))
    return plonk_b200


def _edge(mod):
    return [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, (1 << 255) % mod]


@pytest.mark.parametrize("name,mod,nbytes", [("fr", R.R_MOD, 32), ("fp", R.P_MOD, 48)])
def test_device_montgomery_mul(pb, name, mod, nbytes):
    """The PTX carry chains as compiled by ptxas (the host test only covers the algorithm)."""
    from plonk_b200._lib import check, lib

    rng = random.Random(7)
    vals = _edge(mod) + [rng.randrange(mod) for _ in range(4000)]
    a = vals
    b = [rng.choice(vals) for _ in a]
    A = b"".join(x.to_bytes(nbytes, "little") for x in a)
    B = b"".join(x.to_bytes(nbytes, "little") for x in b)
    out = ctypes.create_string_buffer(len(A))
    check(getattr(lib(), f"pb200_selftest_{name}_mul")(A, B, out, len(a)))
    rinv = pow(1 << (8 * nbytes), -1, mod)
    got = [int.from_bytes(out.raw[i * nbytes : (i + 1) * nbytes], "little") for i in range(len(a))]
    assert got == [x * y * rinv % mod for x, y in zip(a, b)]


def test_device_fp_product_forms(pb):
    """a*b, a^2 and a*b - c*d as ptxas compiles them for the G1 formulas (two-pipe Fp product: integer
    product + DFMA.RZ Montgomery reduction on 48-bit limbs), incl. operands with saturated limbs."""
    from plonk_b200._lib import check, lib

    mod, nb = R.P_MOD, 48
    rng = random.Random(11)
    top = mod >> (8 * nb - 32)
    edge = _edge(mod) + [mod - 1 - k for k in range(20)] + [((top - k) << (8 * nb - 32)) | ((1 << (8 * nb - 32)) - 1) for k in (1, 2)]
    edge += [(1 << (48 * k)) % mod for k in range(1, 8)] + [((1 << (48 * k)) - 1) % mod for k in range(1, 9)]
    vals = edge + [rng.randrange(mod) for _ in range(6000)]
    a = vals + [rng.choice(edge) for _ in range(1000)]
    b, c, d = ([rng.choice(vals) for _ in a] for _ in range(3))
    b[-1000:] = [rng.choice(edge) for _ in range(1000)]
    pack = lambda xs: b"".join(x.to_bytes(nb, "little") for x in xs)
    n = len(a)
    out = ctypes.create_string_buffer(3 * n * nb)
    check(lib().pb200_selftest_fp_ops(pack(a), pack(b), pack(c), pack(d), out, ctypes.c_size_t(n)))
    got = [int.from_bytes(out.raw[i * nb : (i + 1) * nb], "little") for i in range(3 * n)]
    rinv = pow(1 << 384, -1, mod)
    assert got[:n] == [x * y * rinv % mod for x, y in zip(a, b)]
    assert got[n : 2 * n] == [x * x * rinv % mod for x in a]
    assert got[2 * n :] == [(x * y - z * w) * rinv % mod for x, y, z, w in zip(a, b, c, d)]


def test_lagrange_key_matches_direct_sum(pb):
    """Inverse NTT over group elements (csrc/ecntt.cu) against [L_j(x)]G = (1/n) sum_i w^(-ij) [x^i]G, and a
    commitment through evaluations against CommitKey::commit of the interpolated polynomial."""
    import importlib.util
    import os

    from plonk_b200._lib import check, lib

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("lagrange_commit_model", os.path.join(here, "models", "lagrange_commit_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    rng = random.Random(5)
    for log_n in (0, 1, 3, 5):
        n = 1 << log_n
        powers = R.srs_setup(n + 8, R.StdRng.seed_from_u64(40 + log_n), keep=n)
        out = ctypes.create_string_buffer(96 * n)
        check(lib().pb200_g1_lagrange_key(bases_to_abi(powers), ctypes.c_size_t(n), out))
        got = [R.g1_from_raw_bytes(out.raw[96 * j : 96 * j + 96]) for j in range(n)]
        if log_n <= 3:
            assert got == m.lagrange_key(powers, n)
        vals = [rng.choice([0, 1, 3, rng.randrange(R.R_MOD)]) for _ in range(n)]
        want = R.commit(powers, R.EvaluationDomain(n).ifft(vals))
        assert R.jac_to_affine(R.msm_pippenger(got, vals)) == want
    assert lib().pb200_g1_lagrange_key(bases_to_abi(powers[:3]), ctypes.c_size_t(3), out) == -2  # not a power of two


@pytest.mark.parametrize("log_n", list(range(0, 14)))
def test_ntt_matches_oracle_all_directions(pb, log_n):
    rng = random.Random(1000 + log_n)
    n = 1 << log_n
    dom = pb.EvaluationDomain(n)
    ref = R.EvaluationDomain(n)
    assert dom.size == ref.size == n
    for in_len in sorted({n, max(n // 8 + 3, 1) if n >= 8 else n, 0, n + 5}):
        x = rand_fr(rng, in_len)
        xb = to_abi(x)
        assert from_abi(dom.fft(xb)) == ref.fft(x), (log_n, in_len, "fft")
        assert from_abi(dom.ifft(xb)) == ref.ifft(x), (log_n, in_len, "ifft")
        assert from_abi(dom.coset_fft(xb)) == ref.coset_fft(x), (log_n, in_len, "coset_fft")
        assert from_abi(dom.coset_ifft(xb)) == ref.coset_ifft(x), (log_n, in_len, "coset_ifft")


def test_ntt_edge_vectors_and_closed_forms(pb):
    # reference src/fft/domain.rs:570-651
    n = 1 << 12
    dom, ref = pb.EvaluationDomain(n), R.EvaluationDomain(n)
    ramp = [i + 1 for i in range(n)]  # domain.rs:574-576
    ev = dom.fft(to_abi(ramp))
    assert from_abi(ev) == ref.fft(ramp)
    assert from_abi(dom.ifft(ev)) == ramp
    for vec in ([0] * n, [1] + [0] * (n - 1), [R.R_MOD - 1] * n):
        assert from_abi(dom.fft(to_abi(vec))) == ref.fft(vec)
    d8 = pb.EvaluationDomain(1 << 8)
    r8 = R.EvaluationDomain(1 << 8)
    lin = from_abi(d8.coset_fft(to_abi([0, 1])))  # linear_coset_evaluations_match_closed_form
    assert lin == [R.GENERATOR * pow(r8.group_gen, i, R.R_MOD) % R.R_MOD for i in range(1 << 8)]


@pytest.mark.parametrize("log_n", [16, 19, 20, 23])  # 2^23: the quotient domain of a 2^20-gate circuit
def test_ntt_large_properties(pb, log_n):
    """Sizes the Python oracle cannot reach in seconds: size-independent properties + spot checks
    of single outputs against the definition of the DFT (Horner evaluation at w^k)."""
    rng = random.Random(log_n)
    n = 1 << log_n
    dom = pb.EvaluationDomain(n)
    ref = R.EvaluationDomain(n)
    x = rand_fr(rng, n // 8 + 3)
    xb = to_abi(x)
    ev = dom.coset_fft(xb)
    back = dom.coset_ifft(ev)
    assert back[: len(xb)] == xb and back[len(xb) :] == bytes(len(back) - len(xb))
    evals = from_abi(ev[: 32 * 4]) + from_abi(ev[32 * (n - 1) :])
    for k, got in zip([0, 1, 2, 3, n - 1], evals):
        assert got == R.poly_eval(x, R.GENERATOR * pow(ref.group_gen, k, R.R_MOD) % R.R_MOD)
    full = rand_fr(rng, 64)
    fb = to_abi(full)
    assert dom.ifft(dom.fft(fb))[: len(fb)] == fb
    # batched call == individual calls
    v2 = to_abi(rand_fr(rng, 40))
    outs = dom.batch([xb, v2], 0, 1)
    assert outs[0] == ev and outs[1] == dom.coset_fft(v2)


def _check_msm(pb, pts, scalars_list):
    key = pb.CommitKey(bases_to_abi(pts))
    got = key.commit_batch([to_abi(s) for s in scalars_list])
    for g, s in zip(got, scalars_list):
        want = R.jac_to_affine(R.msm_pippenger(pts, R.poly_trim(s)))
        assert R.g1_from_raw_bytes(g.raw) == want
        assert g.to_bytes() == R.g1_compress(want)


@pytest.mark.parametrize("n", [1, 2, 23, 100, 700])
def test_msm_matches_oracle(pb, n):
    rng = random.Random(n)
    pts = progression_bases(n, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    _check_msm(pb, pts, [rand_fr(rng, n), rand_fr(rng, max(1, n // 2)), [rng.randrange(1 << 16) for _ in range(n)]])


def test_msm_edge_cases(pb):
    rng = random.Random(5)
    n = 64
    pts = progression_bases(n, 3, 5)
    pts[7] = None  # identity base
    pts[9] = pts[8]  # repeated point -> doubling inside a bucket
    pts[11] = R.g1_neg(pts[10])  # P and -P in the same bucket -> cancellation
    cases = [
        [0] * n,
        [1] * n,
        [R.R_MOD - 1] * n,
        [5] * n,  # all-equal scalars: one bucket per window
        [rng.randrange(2) for _ in range(n)],
        rand_fr(rng, n)[:-1] + [0],
        [0] * (n - 1) + [7],
    ]
    _check_msm(pb, pts, cases)
    key = pb.CommitKey(bases_to_abi(pts[:8]))
    with pytest.raises(pb.PolynomialDegreeTooLarge):  # key.rs:816-824
        key.commit(to_abi([1] * 9))
    assert key.commit(to_abi([3] * 8 + [0])).raw  # trailing zeros are trimmed first


def test_msm_large_known_discrete_log(pb):
    """2^16 points: bases [p0 + i*b] G so the answer is a single scalar multiplication."""
    rng = random.Random(16)
    n = (1 << 16) + 7
    p0, step = rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD)
    pts = progression_bases(n, p0, step)
    key = pb.CommitKey(bases_to_abi(pts))
    polys = [rand_fr(rng, n - 5), rand_fr(rng, n), rand_fr(rng, n - 6), rand_fr(rng, 1 << 15)]
    got = key.commit_batch([to_abi(p) for p in polys])
    for g, s in zip(got, polys):
        k = sum(si * (p0 + i * step) for i, si in enumerate(s)) % R.R_MOD
        assert R.g1_from_raw_bytes(g.raw) == R.g1_mul(R.G1_GEN, k)


def test_srs_setup_matches_oracle(pb):
    """PublicParameters::setup restated on the device (srs.rs:61-100) vs the oracle."""
    from plonk_b200._lib import check, lib

    x, gs = 0x1234567, 0x7654321
    n = 40
    out = ctypes.create_string_buffer(96 * n)
    check(lib().pb200_srs_setup_from_secret(R.fr_to_mont_bytes(x), R.fr_to_mont_bytes(gs), n, out))
    want = R.srs_from_secret(n, x, gs)
    assert [R.g1_from_raw_bytes(out.raw[96 * i : 96 * i + 96]) for i in range(n)] == want


def test_msm_2_20_points_against_known_secret(pb):
    """BASELINE.json configs[2]/[3] size: 2^20-point commit key generated on the device as [x^i] g;
    the commitment must equal [p(x)] g (SURVEY.md Appendix D note) - one scalar multiplication."""
    from plonk_b200._lib import check, lib

    rng = random.Random(20)
    n = (1 << 20) + 7
    x, gs = rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD)
    raw = ctypes.create_string_buffer(96 * n)
    check(lib().pb200_srs_setup_from_secret(R.fr_to_mont_bytes(x), R.fr_to_mont_bytes(gs), n, raw))
    key = pb.CommitKey(raw.raw)
    poly = rand_fr(rng, n - 3)
    got = key.commit(to_abi(poly))
    assert R.g1_from_raw_bytes(got.raw) == R.g1_mul(R.g1_mul(R.G1_GEN, gs), R.poly_eval(poly, x))
    # partial MSMs over point ranges add up to the full one (the multi-GPU sharding of section 8e-ii)
    from plonk_b200 import dist as pd

    parts = []
    for r in range(3):
        first, count = pd.shard_range(len(poly), r, 3)
        out = ctypes.create_string_buffer(96)
        check(lib().pb200_msm_g1_range(key._h, first, to_abi(poly[first : first + count]), count, out))
        parts.append(out.raw)
    assert pd.g1_sum(parts) == got.raw


def _structured_msm_case(log_n, seed):
    """n = 2^log_n distinct full-range scalars s_i = u[i mod 4096] + v[i div 4096] (mod r) whose polynomial has a
    closed form, so that [p(x)]g is one scalar multiplication: p(x) = U(x) (x^n - 1)/(x^m - 1) + V(x^m) (x^m - 1)/(x - 1)."""
    import torch

    import bench

    rng = random.Random(seed)
    n, m = 1 << log_n, 4096
    x, gs = rng.randrange(2, R.R_MOD), rng.randrange(1, R.R_MOD)
    u, v = rand_fr(rng, m), rand_fr(rng, n // m)
    scalars = bench.fr_outer_sum(torch, u, v).cpu().numpy().tobytes()
    xm = pow(x, m, R.R_MOD)
    geo_n = (pow(x, n, R.R_MOD) - 1) * pow(xm - 1, -1, R.R_MOD) % R.R_MOD
    geo_m = (xm - 1) * pow(x - 1, -1, R.R_MOD) % R.R_MOD
    px = (R.poly_eval(u, x) * geo_n + R.poly_eval(v, xm) * geo_m) % R.R_MOD
    assert from_abi(scalars[: 32 * 3]) == [(a + v[0]) % R.R_MOD for a in u[:3]]
    assert from_abi(scalars[32 * (m + 1) : 32 * (m + 2)]) == [(u[1] + v[1]) % R.R_MOD]
    return n, x, gs, scalars, px


@pytest.mark.parametrize("log_n", [22, 24])
def test_msm_large_points_against_known_secret(pb, log_n):
    """BASELINE.json configs[3] upper range on one GPU: 2^22 and 2^24 points, 20-bit windows, distinct
    full-range scalars; the result must equal [p(x)] g."""
    from plonk_b200._lib import check, lib

    n, x, gs, scalars, px = _structured_msm_case(log_n, log_n)
    assert len(scalars) == 32 * n
    raw = ctypes.create_string_buffer(96 * n)
    check(lib().pb200_srs_setup_from_secret(R.fr_to_mont_bytes(x), R.fr_to_mont_bytes(gs), n, raw))
    key = pb.CommitKey(raw.raw)
    del raw
    got = key.commit(scalars)
    assert R.g1_from_raw_bytes(got.raw) == R.g1_mul(R.g1_mul(R.G1_GEN, gs), px)


def test_msm_repeated_scalar_block_is_skewed_but_exact(pb):
    """A coefficient vector repeating one 4096-element block puts 256 entries into each of 53 000 of the 2^19
    buckets and none into the rest: p(x) = B(x) * sum_j x^(4096 j)."""
    from plonk_b200._lib import check, lib

    rng = random.Random(22)
    n = 1 << 20
    x, gs = rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD)
    raw = ctypes.create_string_buffer(96 * n)
    check(lib().pb200_srs_setup_from_secret(R.fr_to_mont_bytes(x), R.fr_to_mont_bytes(gs), n, raw))
    key = pb.CommitKey(raw.raw)
    del raw
    block = rand_fr(rng, 1 << 12)
    got = key.commit(to_abi(block) * (n >> 12))
    geo = (pow(x, n, R.R_MOD) - 1) * pow(pow(x, 1 << 12, R.R_MOD) - 1, -1, R.R_MOD) % R.R_MOD
    assert R.g1_from_raw_bytes(got.raw) == R.g1_mul(R.g1_mul(R.G1_GEN, gs), R.poly_eval(block, x) * geo % R.R_MOD)


def test_commit_key_window_and_sharded_key_on_one_rank(pb):
    """pb200_srs_upload_window (the ranks of a point-sharded MSM must agree on one window width) gives the same
    commitments whatever the width, and plonk_b200.dist.ShardedCommitKey degenerates to a plain commit on one rank
    (device-resident scalars; the replica below the threshold, the slice above it)."""
    import torch

    from plonk_b200 import dist as pd
    from plonk_b200._lib import check, lib

    L = lib()
    rng = random.Random(44)
    n = 3000
    pts = progression_bases(n, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD))
    raw = bases_to_abi(pts)
    scalars = rand_fr(rng, n)
    want = R.g1_to_raw_bytes(R.jac_to_affine(R.msm_pippenger(pts, scalars)))
    assert L.pb200_msm_window_for(n) == 11 and L.pb200_msm_window_for(1 << 20) == 20
    for c in (0, 7, 11, 13):
        h = ctypes.c_void_p()
        check(L.pb200_srs_upload_window(raw, n, c, ctypes.byref(h)))
        assert L.pb200_srs_window(h) == (c or 11)
        out = ctypes.create_string_buffer(96)
        check(L.pb200_msm_g1(h, to_abi(scalars), n, 1, n, out))
        assert out.raw == want, c
        L.pb200_srs_free(h)
    h = ctypes.c_void_p()
    assert L.pb200_srs_upload_window(raw, n, 21, ctypes.byref(h)) == -4 and L.pb200_srs_upload_window(raw, n, 1, ctypes.byref(h)) == -4
    d_sc = torch.frombuffer(bytearray(to_abi(scalars)), dtype=torch.uint8).cuda()
    for threshold in (1 << 18, 1000):  # replica path, slice path
        key = pd.ShardedCommitKey(raw, n, None, threshold=threshold, replica_raw=raw[: 96 * min(n, threshold)])
        assert not key.uses_collective(n)
        assert key.commit_dev(d_sc.data_ptr(), n) == want
        assert key.commit_dev(d_sc.data_ptr(), 500) == R.g1_to_raw_bytes(R.jac_to_affine(R.msm_pippenger(pts[:500], scalars[:500])))
        key.free()


def test_msm_skewed_scalars_heavy_buckets(pb):
    """Scalar distributions that put thousands of points into one bucket (the reference's rayon path
    has no such cliff; ours routes buckets longer than 512 entries to whole CTAs)."""
    import time

    rng = random.Random(33)
    n = 1 << 14
    p0, step = rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD)
    pts = progression_bases(n, p0, step)
    key = pb.CommitKey(bases_to_abi(pts))
    big = rng.randrange(R.R_MOD)
    cases = [[big] * n, [rng.randrange(2) for _ in range(n)], [rng.choice([3, big, R.R_MOD - 1]) for _ in range(n)],
             [big] * (n // 2) + rand_fr(rng, n // 2)]
    t0 = time.time()
    got = key.commit_batch([to_abi(s) for s in cases])
    dt = time.time() - t0
    for g, s in zip(got, cases):
        k = sum(si * (p0 + i * step) for i, si in enumerate(s)) % R.R_MOD
        assert R.g1_from_raw_bytes(g.raw) == R.g1_mul(R.G1_GEN, k)
    assert dt < 5.0, f"skewed MSM took {dt:.2f}s"


def test_g1_decompress_matches_oracle_and_rejects_malformed_points(pb):
    """CommitKey::from_slice (key.rs:319-326): compressed commit key -> raw points on the GPU."""
    from plonk_b200.kzg import CommitKey, PointMalformed, g1_compress, g1_decompress

    rng = random.Random(21)
    pts = progression_bases(300, rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD)) + [None, R.G1_GEN, R.g1_neg(R.G1_GEN)]
    comp = b"".join(R.g1_compress(p) for p in pts)
    raw = g1_decompress(comp)
    assert raw == bases_to_abi(pts)
    assert g1_compress(raw) == comp  # and back through the product's own encoder
    assert g1_decompress(b"") == b""
    key = CommitKey.from_slice(comp[: 48 * 64])
    coeffs = rand_fr(rng, 64)
    assert R.g1_to_raw_bytes(R.jac_to_affine(R.msm_naive(pts[:64], coeffs))) == key.commit(to_abi(coeffs)).raw

    def bad(i, enc, **kw):
        blob = comp[: 48 * i] + enc + comp[48 * (i + 1) :]
        with pytest.raises(PointMalformed) as e:
            g1_decompress(blob, **kw)
        assert f"point {i}" in str(e.value)

    good = bytearray(R.g1_compress(pts[5]))
    bad(5, bytes([good[0] & 0x7F]) + bytes(good[1:]))                      # compression flag missing
    bad(7, bytes([0xC0]) + bytes(46) + b"\x01")                            # infinity with a non-zero x
    bad(7, bytes([0xE0]) + bytes(47))                                      # infinity with the sort flag
    bad(9, bytes([0x80 | 0x1A]) + (R.P_MOD.to_bytes(48, "big"))[1:])       # x = p: not canonical
    x = 1
    while pow((x ** 3 + 4) % R.P_MOD, (R.P_MOD - 1) // 2, R.P_MOD) == 1:
        x += 1
    bad(11, bytes([0x80]) + x.to_bytes(48, "big")[1:])                     # x^3 + 4 is not a square
    # on the curve but outside the prime-order subgroup (the curve's cofactor is ~2^126)
    x = 2
    while True:
        y2 = (x ** 3 + 4) % R.P_MOD
        y = pow(y2, (R.P_MOD + 1) // 4, R.P_MOD)
        if y * y % R.P_MOD == y2 and R.jac_to_affine(R.jac_mul(R.jac_from_affine((x, y)), R.R_MOD)) is not None:
            break
        x += 1
    enc = R.g1_compress((x, y))
    bad(299, enc)
    blob = comp[: 48 * 299] + enc + comp[48 * 300 :]
    assert g1_decompress(blob, check_subgroup=False)[96 * 299 : 96 * 300] == R.g1_to_raw_bytes((x, y))
    # the first malformed point is the one reported
    with pytest.raises(PointMalformed) as e:
        g1_decompress(enc + comp[48:96] + bytes([0x00]) * 48)
    assert "point 0" in str(e.value)


def test_kzg_open_and_check_with_gpu_commitments(pb):
    """key.rs:826-1018 (test_basic_commit, test_aggregate_witness) on the GPU MSM: commitments and
    opening witnesses are committed with CommitKey.commit; the reference's pairing check
    e(C - [v]g, H) == e(W, [x - z]H) is the G1 identity C - [v]g == [x - z]W under the known secret."""
    rng = random.Random(826)
    x, gs = rng.randrange(1, R.R_MOD), rng.randrange(1, R.R_MOD)
    pts = R.srs_from_secret(28 + 1, x, gs)
    key = pb.CommitKey(bases_to_abi(pts))
    g = pts[0]
    z = 10

    def commit(poly):
        return R.g1_from_raw_bytes(key.commit(to_abi(poly)).raw)

    def check(c, v, w):  # OpeningKey::check (key.rs:489-507)
        lhs = R.g1_add(c, R.g1_neg(R.g1_mul(g, v))) if v else c
        return lhs == R.g1_mul(w, (x - z) % R.R_MOD)

    # open_single (key.rs:743-760): witness = (p - v) / (X - z)
    poly = rand_fr(rng, 26)
    v = R.poly_eval(poly, z)
    w = commit(R.ruffini(poly, z))  # ruffini drops the remainder, which is p(z)
    assert check(commit(poly), v, w)
    assert not check(commit(poly), (v + 1) % R.R_MOD, w)
    # open_multiple / compute_aggregate_witness (key.rs:394-417, 762-794): sum_k v^k p_k, one witness
    polys = [rand_fr(rng, 26), rand_fr(rng, 28), rand_fr(rng, 28)]
    vch = rng.randrange(R.R_MOD)
    agg = [0] * 28
    power = 1
    for p in polys:
        for i, c in enumerate(p):
            agg[i] = (agg[i] + c * power) % R.R_MOD
        power = power * vch % R.R_MOD
    w = commit(R.ruffini(agg, z))
    flat_c, flat_v, power = None, 0, 1
    for p in polys:  # AggregateProof::flatten (commitment_scheme/kzg10/proof.rs)
        flat_c = R.g1_mul(commit(p), power) if flat_c is None else R.g1_add(flat_c, R.g1_mul(commit(p), power))
        flat_v = (flat_v + R.poly_eval(p, z) * power) % R.R_MOD
        power = power * vch % R.R_MOD
    assert check(flat_c, flat_v, w)
