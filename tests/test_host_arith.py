"""Validates the multi-limb Montgomery / G1 algorithms of plonk_b200/csrc/{bigint,g1}.cuh on the CPU.

The headers are compiled for the host by g++ (PTX carry chains emulated) into tests/hosttest and
compared with the Python oracle.  This checks the *algorithms* that the CUDA kernels instantiate;
the GPU parity tests (-m gpu) check the compiled kernels themselves."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import pyref as R

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ht():
    so = os.path.join(HERE, "hosttest", "libhosttest.so")
    src = os.path.join(HERE, "hosttest", "hosttest.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-frounding-math", "-mfma", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def _edge(mod):
    return [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (1 << 32) - 1, 1 << 32, (1 << 255) % mod, (1 << 64) - 1]


def _run(fn, op, a, b, nbytes):
    n = len(a)
    A = b"".join(x.to_bytes(nbytes, "little") for x in a)
    B = b"".join(x.to_bytes(nbytes, "little") for x in b) if b is not None else None
    out = ctypes.create_string_buffer(n * nbytes)
    assert fn(op, A, B, out, ctypes.c_size_t(n)) == 0
    return [int.from_bytes(out.raw[i * nbytes : (i + 1) * nbytes], "little") for i in range(n)]


@pytest.mark.parametrize("name,mod,nbytes", [("fr", R.R_MOD, 32), ("fp", R.P_MOD, 48)])
def test_field_ops(ht, name, mod, nbytes):
    fn = getattr(ht, f"ht_{name}_op")
    rng = random.Random(1234)
    vals = _edge(mod) + [rng.randrange(mod) for _ in range(300)]
    a = [x for x in vals for _ in range(3)]
    b = [rng.choice(vals) for _ in a]
    Rm = 1 << (8 * nbytes)
    Rinv = pow(Rm, -1, mod)
    assert _run(fn, 0, a, b, nbytes) == [x * y * Rinv % mod for x, y in zip(a, b)]
    assert _run(fn, 1, a, b, nbytes) == [(x + y) % mod for x, y in zip(a, b)]
    assert _run(fn, 2, a, b, nbytes) == [(x - y) % mod for x, y in zip(a, b)]
    assert _run(fn, 4, a, None, nbytes) == [x * Rm % mod for x in a]
    assert _run(fn, 5, a, None, nbytes) == [x * Rinv % mod for x in a]
    sq = vals + [mod - 1 - k for k in range(40)] + [(1 << (8 * nbytes - 1)) % mod, ((1 << 31) - 1) * ((1 << (8 * nbytes)) // ((1 << 32) - 1)) % mod]
    top = mod >> (8 * nbytes - 32)  # top limb of the modulus: largest operands with every lower limb saturated
    sq += [((top - k) << (8 * nbytes - 32)) | ((1 << (8 * nbytes - 32)) - 1) for k in (1, 2)] + [rng.randrange(mod) for _ in range(2000)]
    assert _run(fn, 6, sq, None, nbytes) == [x * x * Rinv % mod for x in sq]  # dedicated squaring (half the partial products)
    small = vals[:40]
    # inverse in Montgomery form: inv(aR) = a^-1 R
    got = _run(fn, 3, [x * Rm % mod for x in small], None, nbytes)
    assert got == [(pow(x, -1, mod) * Rm % mod) if x else 0 for x in small]


def test_fp_two_products_one_reduction(ht):
    """Field::mul2 / mul_sub (a*b +- c*d with a single Montgomery reduction), including operands at
    the top of the range, where the running sum is largest."""
    mod, nb = R.P_MOD, 48
    rng = random.Random(4321)
    vals = _edge(mod) + [rng.randrange(mod) for _ in range(200)]
    quads = [(mod - 1,) * 4, (mod - 1, mod - 1, 0, 5), (0, 0, mod - 1, mod - 1), (mod - 1, mod - 2, mod - 1, 1)]
    quads += [tuple(rng.choice(vals) for _ in range(4)) for _ in range(2000)]
    Rinv = pow(1 << 384, -1, mod)
    cols = [b"".join(q[k].to_bytes(nb, "little") for q in quads) for k in range(4)]
    for sub in (0, 1):
        out = ctypes.create_string_buffer(len(quads) * nb)
        assert ht.ht_fp_mul2(sub, *cols, out, ctypes.c_size_t(len(quads))) == 0
        got = [int.from_bytes(out.raw[i * nb : (i + 1) * nb], "little") for i in range(len(quads))]
        sign = -1 if sub else 1
        assert got == [(a * b + sign * c * d) * Rinv % mod for a, b, c, d in quads]


def test_g1_sum_and_mul(ht):
    rng = random.Random(99)
    g = R.G1_GEN
    pts = [R.g1_mul(g, rng.randrange(1, R.R_MOD)) for _ in range(12)]
    # special cases: repeated point (doubling), P then -P (cancellation), identity entries
    seq = pts + [pts[0], pts[0], None, pts[3], pts[3]]
    neg = [rng.randrange(2) for _ in pts] + [0, 0, 0, 0, 1]
    raw = b"".join(R.g1_to_raw_bytes(p) for p in seq)
    out = ctypes.create_string_buffer(96)
    assert ht.ht_g1_sum(raw, bytes(neg), ctypes.c_size_t(len(seq)), out) == 0
    want = None
    for p, s in zip(seq, neg):
        want = R.g1_add(want, R.g1_neg(p) if s else p)
    assert R.g1_from_raw_bytes(out.raw) == want
    # everything cancels -> identity
    seq2 = [pts[1], pts[1]]
    assert ht.ht_g1_sum(b"".join(R.g1_to_raw_bytes(p) for p in seq2), bytes([0, 1]), ctypes.c_size_t(2), out) == 0
    assert R.g1_from_raw_bytes(out.raw) is None
    for k in (0, 1, 2, 3, 0xDEADBEEFCAFEF00D):
        assert ht.ht_g1_mul_small(R.g1_to_raw_bytes(pts[2]), ctypes.c_uint64(k), out) == 0
        assert R.g1_from_raw_bytes(out.raw) == R.g1_mul(pts[2], k)


def test_fp_two_pipe_product(ht):
    """Field::mul_hybrid / sqr_hybrid / mul2_hybrid (integer product + double-precision Montgomery
    reduction on 48-bit limbs, csrc/bigint.cuh) against big-integer arithmetic."""
    mod, nb = R.P_MOD, 48
    rng = random.Random(4321)
    top = mod >> (8 * nb - 32)
    edge = _edge(mod) + [mod - 1 - k for k in range(20)] + [((top - k) << (8 * nb - 32)) | ((1 << (8 * nb - 32)) - 1) for k in (1, 2)]
    edge += [(1 << (48 * k)) % mod for k in range(1, 8)] + [((1 << (48 * k)) - 1) % mod for k in range(1, 9)]
    vals = edge + [rng.randrange(mod) for _ in range(3000)]
    a = vals + [rng.choice(edge) for _ in range(500)]
    b = [rng.choice(vals) for _ in a[: len(vals)]] + [rng.choice(edge) for _ in range(500)]
    c = [rng.choice(vals) for _ in a]
    d = [rng.choice(vals) for _ in a]
    Rinv = pow(1 << 384, -1, mod)
    pack = lambda xs: b"".join(x.to_bytes(nb, "little") for x in xs)
    out = ctypes.create_string_buffer(len(a) * nb)
    unpack = lambda: [int.from_bytes(out.raw[i * nb : (i + 1) * nb], "little") for i in range(len(a))]
    assert ht.ht_fp_hybrid(0, pack(a), pack(b), None, None, out, ctypes.c_size_t(len(a))) == 0
    assert unpack() == [x * y * Rinv % mod for x, y in zip(a, b)]
    assert ht.ht_fp_hybrid(1, pack(a), None, None, None, out, ctypes.c_size_t(len(a))) == 0
    assert unpack() == [x * x * Rinv % mod for x in a]
    assert ht.ht_fp_hybrid(2, pack(a), pack(b), pack(c), pack(d), out, ctypes.c_size_t(len(a))) == 0
    assert unpack() == [(x * y + z * w) * Rinv % mod for x, y, z, w in zip(a, b, c, d)]


def test_fr_two_pipe_product(ht):
    """Fr::mul_hybrid: 256 bits are five 48-bit reduction steps and one of 16 bits."""
    mod, nb = R.R_MOD, 32
    rng = random.Random(987)
    edge = _edge(mod) + [mod - 1 - k for k in range(20)] + [(1 << (48 * k)) % mod for k in range(1, 6)] + [((1 << (48 * k)) - 1) % mod for k in range(1, 6)]
    edge += [(1 << 240) - 1, ((1 << 255) - 1) % mod, (0x73ED << 240) - 1]
    vals = edge + [rng.randrange(mod) for _ in range(4000)]
    a = vals + [rng.choice(edge) for _ in range(500)]
    b = [rng.choice(vals) for _ in vals] + [rng.choice(edge) for _ in range(500)]
    pack = lambda xs: b"".join(x.to_bytes(nb, "little") for x in xs)
    out = ctypes.create_string_buffer(len(a) * nb)
    assert ht.ht_fr_hybrid(pack(a), pack(b), out, ctypes.c_size_t(len(a))) == 0
    Rinv = pow(1 << 256, -1, mod)
    got = [int.from_bytes(out.raw[i * nb : (i + 1) * nb], "little") for i in range(len(a))]
    assert got == [x * y * Rinv % mod for x, y in zip(a, b)]


def test_group_element_inverse_ntt_gives_the_lagrange_commit_key(ht):
    """csrc/ecntt.cuh (butterfly, twiddle scalar multiplication, bit-reversed read-out) on the host against
    the direct sum [L_j(x)]G = (1/n) sum_i w^(-ij) [x^i]G of tests/models/lagrange_commit_model.py."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("lagrange_commit_model", os.path.join(HERE, "models", "lagrange_commit_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    for log_n, seed in ((0, 1), (1, 2), (3, 3), (4, 4)):
        n = 1 << log_n
        powers = R.srs_setup(n + 8, R.StdRng.seed_from_u64(seed), keep=n)
        want = m.lagrange_key(powers, n)
        out = ctypes.create_string_buffer(96 * n)
        assert ht.ht_ec_intt(b"".join(R.g1_to_raw_bytes(p) for p in powers[:n]), log_n, out) == 0
        assert [R.g1_from_raw_bytes(out.raw[96 * j : 96 * j + 96]) for j in range(n)] == want


@pytest.fixture(scope="module")
def hp():
    so = os.path.join(HERE, "hosttest", "libhostproduct.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "hosttest", "hostproduct.cpp")])
    return ctypes.CDLL(so)


def test_product_transcript_matches_oracle(hp):
    """plonk_b200/csrc/transcript.h (merlin/STROBE/Keccak + TranscriptProtocol) vs oracle/pyref.py."""
    out = (ctypes.c_uint64 * 4)()
    hp.hp_transcript_scalar(b"test protocol", b"some label", b"some data", ctypes.c_size_t(9), b"challenge", out)
    t = R.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert R.fr_from_mont_bytes(bytes(out)) == t.challenge_scalar(b"challenge")
    rng = random.Random(3)
    for n in (5, 65530, 1 << 40):
        s = rng.randrange(R.R_MOD)
        hp.hp_transcript_mix(R.fr_to_mont_bytes(s), ctypes.c_uint64(n), out)
        t = R.Transcript(b"mix")
        t.circuit_domain_sep(n)
        t.append_scalar(b"s", s)
        t.append_commitment(b"c", None)
        a = t.challenge_scalar(b"a")
        t.append_scalar(b"a", a)
        b = t.challenge_scalar(b"b")
        assert R.fr_from_mont_bytes(bytes(out)) == (a * b + s) % R.R_MOD


def test_product_host_field_and_compression(hp):
    rng = random.Random(4)
    Rm = 1 << 384
    for _ in range(20):
        x = rng.randrange(1, R.P_MOD)
        out = (ctypes.c_uint64 * 6)()
        hp.hp_fp_inv((x * Rm % R.P_MOD).to_bytes(48, "little"), out)
        assert int.from_bytes(bytes(out), "little") == pow(x, -1, R.P_MOD) * Rm % R.P_MOD
    for p in [None] + [R.g1_mul(R.G1_GEN, rng.randrange(1, R.R_MOD)) for _ in range(10)]:
        out = ctypes.create_string_buffer(48)
        hp.hp_g1_compress(R.g1_to_raw_bytes(p), out)
        assert out.raw == R.g1_compress(p)
    # binary-GCD inverses: edge values (1, 2, p - 1, powers of two, values with long runs of zero bits) and random ones
    for mod, nb, fn in ((R.P_MOD, 48, hp.hp_fp_inv_bingcd), (R.R_MOD, 32, hp.hp_fr_inv_bingcd)):
        mont = 1 << (8 * nb)
        vals = [1, 2, 3, mod - 1, mod - 2, (mod - 1) // 2, 1 << 64, 1 << 200, (1 << 250) - 1, 0] + [rng.randrange(1, mod) for _ in range(300)]
        for x in vals:
            out = (ctypes.c_uint64 * (nb // 8))()
            fn((x * mont % mod).to_bytes(nb, "little"), out)
            want = pow(x, -1, mod) * mont % mod if x else 0
            assert int.from_bytes(bytes(out), "little") == want, (nb, x)
