"""CPU oracle (pure Python big-int) for the dusk-plonk hot path and the prover glue around it.

TEST INFRASTRUCTURE ONLY.  Nothing under ``plonk_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg use it, and only as the
checker.  Parity status: PINNED - ``prove()`` below reproduces the reference's own known-answer test
``deterministic_v3_proof_matches_base_digest`` (reference ``src/compiler/prover.rs:1132-1162``)
bit-for-bit (tests/test_oracle_kat.py).

Every function cites the reference file:line it restates (paths relative to /root/reference).
Field/curve arithmetic, StdRng and merlin live in un-vendored dependencies of the reference
(dusk-bls12_381 0.14, rand 0.8 / rand_chacha 0.3, merlin 3.0); they are restated from their public
specifications and validated through the KAT above.

All Fr / Fp values in this module are canonical Python ints (NOT Montgomery form); the helpers
``fr_to_mont_bytes`` / ``fr_from_mont_bytes`` translate to the 4xu64 little-endian Montgomery layout
the C ABI uses.
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------------
# Fields (dusk-bls12_381: BlsScalar = Fr, Fp).  Constants cross-checked in SURVEY.md Appendix C
# against src/composer.rs:334-339 (Montgomery -1) and src/commitment_scheme/kzg10/key.rs:1033-1040.
# --------------------------------------------------------------------------------------------
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
P_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
FR_MONT_R = (1 << 256) % R_MOD
FP_MONT_R = (1 << 384) % P_MOD
GENERATOR = 7  # dusk_bls12_381::GENERATOR (src/fft/domain.rs:115)
TWO_ADACITY = 32
ROOT_OF_UNITY = pow(GENERATOR, (R_MOD - 1) >> TWO_ADACITY, R_MOD)
K1, K2, K3 = 7, 13, 17  # src/composer/permutation/constants.rs:14-16
# dusk_jubjub::EDWARDS_D = -(10240/10241) mod r
EDWARDS_D = (-10240 * pow(10241, R_MOD - 2, R_MOD)) % R_MOD

G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)


def fr_inv(a: int) -> int:
    return pow(a, R_MOD - 2, R_MOD)


def fp_inv(a: int) -> int:
    return pow(a, P_MOD - 2, P_MOD)


def fr_to_mont_bytes(a: int) -> bytes:
    """BlsScalar in-memory layout: 4 x u64 LE limbs of a*2^256 mod r."""
    return ((a * FR_MONT_R) % R_MOD).to_bytes(32, "little")


def fr_from_mont_bytes(b: bytes) -> int:
    return (int.from_bytes(b, "little") * fr_inv(FR_MONT_R)) % R_MOD


def fr_vec_to_mont_bytes(v: Sequence[int]) -> bytes:
    return b"".join(fr_to_mont_bytes(x) for x in v)


def fr_vec_from_mont_bytes(b: bytes) -> List[int]:
    rinv = fr_inv(FR_MONT_R)
    return [(int.from_bytes(b[i : i + 32], "little") * rinv) % R_MOD for i in range(0, len(b), 32)]


def fr_to_bytes(a: int) -> bytes:
    """BlsScalar::to_bytes: 32-byte little-endian canonical."""
    return a.to_bytes(32, "little")


def fr_from_bytes_wide(b: bytes) -> int:
    """BlsScalar::from_bytes_wide: 512-bit little-endian integer reduced mod r."""
    assert len(b) == 64
    return int.from_bytes(b, "little") % R_MOD


# --------------------------------------------------------------------------------------------
# G1 (y^2 = x^3 + 4 over Fp).  Affine points are (x, y) tuples, identity is None.
# --------------------------------------------------------------------------------------------
Affine = Optional[Tuple[int, int]]
Jac = Tuple[int, int, int]  # (X, Y, Z), Z == 0 is the identity
JAC_ID: Jac = (1, 1, 0)


def g1_is_on_curve(p: Affine) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - 4) % P_MOD == 0


def jac_from_affine(p: Affine) -> Jac:
    return JAC_ID if p is None else (p[0], p[1], 1)


def jac_to_affine(p: Jac) -> Affine:
    """Commitment::from(G1Projective) (src/commitment_scheme/kzg10/commitment.rs:89-93)."""
    X, Y, Z = p
    if Z == 0:
        return None
    zi = fp_inv(Z)
    zi2 = zi * zi % P_MOD
    return (X * zi2 % P_MOD, Y * zi2 * zi % P_MOD)


def jac_double(p: Jac) -> Jac:
    X, Y, Z = p
    if Z == 0 or Y == 0:
        return JAC_ID
    A = X * X % P_MOD
    B = Y * Y % P_MOD
    C = B * B % P_MOD
    D = 2 * ((X + B) * (X + B) - A - C) % P_MOD
    E = 3 * A % P_MOD
    F = E * E % P_MOD
    X3 = (F - 2 * D) % P_MOD
    Y3 = (E * (D - X3) - 8 * C) % P_MOD
    Z3 = 2 * Y * Z % P_MOD
    return (X3, Y3, Z3)


def jac_add(p: Jac, q: Jac) -> Jac:
    if p[2] == 0:
        return q
    if q[2] == 0:
        return p
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    Z1Z1 = Z1 * Z1 % P_MOD
    Z2Z2 = Z2 * Z2 % P_MOD
    U1 = X1 * Z2Z2 % P_MOD
    U2 = X2 * Z1Z1 % P_MOD
    S1 = Y1 * Z2 * Z2Z2 % P_MOD
    S2 = Y2 * Z1 * Z1Z1 % P_MOD
    if U1 == U2:
        if S1 == S2:
            return jac_double(p)
        return JAC_ID
    H = (U2 - U1) % P_MOD
    I = (2 * H) * (2 * H) % P_MOD
    J = H * I % P_MOD
    r = 2 * (S2 - S1) % P_MOD
    V = U1 * I % P_MOD
    X3 = (r * r - J - 2 * V) % P_MOD
    Y3 = (r * (V - X3) - 2 * S1 * J) % P_MOD
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % P_MOD
    return (X3, Y3, Z3)


def jac_add_affine(p: Jac, q: Affine) -> Jac:
    return jac_add(p, jac_from_affine(q))


def jac_neg(p: Jac) -> Jac:
    return (p[0], (-p[1]) % P_MOD, p[2])


def jac_mul(p: Jac, k: int) -> Jac:
    acc = JAC_ID
    for bit in bin(k)[2:] if k else "":
        acc = jac_double(acc)
        if bit == "1":
            acc = jac_add(acc, p)
    return acc


def g1_mul(p: Affine, k: int) -> Affine:
    return jac_to_affine(jac_mul(jac_from_affine(p), k % R_MOD))


def g1_add(p: Affine, q: Affine) -> Affine:
    return jac_to_affine(jac_add(jac_from_affine(p), jac_from_affine(q)))


def g1_neg(p: Affine) -> Affine:
    return None if p is None else (p[0], (-p[1]) % P_MOD)


def g1_compress(p: Affine) -> bytes:
    """G1Affine::to_bytes: 48-byte zcash-style compressed encoding (commitment.rs:95-101)."""
    if p is None:
        return bytes([0xC0]) + bytes(47)
    x, y = p
    out = bytearray(x.to_bytes(48, "big"))
    out[0] |= 0x80
    if y > (P_MOD - y):
        out[0] |= 0x20
    return bytes(out)


def g1_decompress(b: bytes) -> Affine:
    assert len(b) == 48 and b[0] & 0x80
    if b[0] & 0x40:
        return None
    sign = bool(b[0] & 0x20)
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    y = pow((x * x * x + 4) % P_MOD, (P_MOD + 1) // 4, P_MOD)
    if (y > P_MOD - y) != sign:
        y = P_MOD - y
    assert g1_is_on_curve((x, y))
    return (x, y)


def g1_to_raw_bytes(p: Affine) -> bytes:
    """ABI layout of one base point: x, y as 6 x u64 LE Montgomery limbs (R = 2^384) = 96 bytes;
    the identity is encoded as x = y = 0 (not on the curve, so unambiguous)."""
    if p is None:
        return bytes(96)
    x, y = p
    return ((x * FP_MONT_R) % P_MOD).to_bytes(48, "little") + ((y * FP_MONT_R) % P_MOD).to_bytes(48, "little")


def g1_from_raw_bytes(b: bytes) -> Affine:
    rinv = fp_inv(FP_MONT_R)
    x = int.from_bytes(b[:48], "little") * rinv % P_MOD
    y = int.from_bytes(b[48:96], "little") * rinv % P_MOD
    if x == 0 and y == 0:
        return None
    return (x, y)


def jac_from_raw_bytes(b: bytes) -> Jac:
    """ABI layout of an MSM result: X, Y, Z as 6 x u64 LE Montgomery limbs = 144 bytes."""
    rinv = fp_inv(FP_MONT_R)
    return tuple(int.from_bytes(b[i * 48 : (i + 1) * 48], "little") * rinv % P_MOD for i in range(3))  # type: ignore


def batch_normalize(points: Sequence[Jac]) -> List[Affine]:
    """G1Projective::batch_normalize (used by srs.rs:87)."""
    acc = 1
    prods = []
    for X, Y, Z in points:
        prods.append(acc)
        if Z != 0:
            acc = acc * Z % P_MOD
    inv = fp_inv(acc)
    out: List[Affine] = [None] * len(points)
    for i in range(len(points) - 1, -1, -1):
        X, Y, Z = points[i]
        if Z == 0:
            continue
        zi = inv * prods[i] % P_MOD
        inv = inv * Z % P_MOD
        zi2 = zi * zi % P_MOD
        out[i] = (X * zi2 % P_MOD, Y * zi2 * zi % P_MOD)
    return out


# --------------------------------------------------------------------------------------------
# MSM (CommitKey::commit -> dusk_bls12_381::multiscalar_mul::msm_variable_base, key.rs:376-388)
# --------------------------------------------------------------------------------------------
def msm_naive(bases: Sequence[Affine], scalars: Sequence[int]) -> Jac:
    """Definition: sum_i scalars[i] * bases[i]; zip semantics (shorter of the two)."""
    acc = JAC_ID
    for b, s in zip(bases, scalars):
        acc = jac_add(acc, jac_mul(jac_from_affine(b), s % R_MOD))
    return acc


def msm_pippenger(bases: Sequence[Affine], scalars: Sequence[int]) -> Jac:
    """Pippenger bucket method as published for msm_variable_base in dusk-bls12_381 0.14 (zexe
    lineage): c = 3 if n < 32 else floor(ln n) + 2; per window 2^c - 1 buckets filled with mixed
    adds, running-sum bucket reduction, then Horner over windows with c doublings each.  The
    result is canonical after affine normalisation, so only its value (not its schedule) is pinned."""
    n = min(len(bases), len(scalars))
    if n == 0:
        return JAC_ID
    c = 3 if n < 32 else int(math.log(n)) + 2
    num_bits = 255
    window_sums = []
    for w_start in range(0, num_bits, c):
        buckets = [JAC_ID] * ((1 << c) - 1)
        res = JAC_ID
        for i in range(n):
            s = scalars[i] % R_MOD
            if s == 0:
                continue
            if s == 1:
                if w_start == 0:
                    res = jac_add_affine(res, bases[i])
                continue
            d = (s >> w_start) & ((1 << c) - 1)
            if d:
                buckets[d - 1] = jac_add_affine(buckets[d - 1], bases[i])
        running = JAC_ID
        for b in reversed(buckets):
            running = jac_add(running, b)
            res = jac_add(res, running)
        window_sums.append(res)
    total = window_sums[-1]
    for ws in reversed(window_sums[:-1]):
        for _ in range(c):
            total = jac_double(total)
        total = jac_add(total, ws)
    return total


def commit(powers_of_g: Sequence[Affine], poly: Sequence[int]) -> Affine:
    """CommitKey::commit (key.rs:376-388).  Raises ValueError for PolynomialDegreeTooLarge."""
    poly = poly_trim(poly)
    degree = max(len(poly) - 1, 0)
    if degree > len(powers_of_g) - 1:
        raise ValueError("PolynomialDegreeTooLarge")
    return jac_to_affine(msm_pippenger(powers_of_g, poly))


# --------------------------------------------------------------------------------------------
# Polynomial helpers (src/fft/polynomial.rs)
# --------------------------------------------------------------------------------------------
def poly_trim(c: Sequence[int]) -> List[int]:
    """Polynomial::from_coefficients_vec trims trailing zeros (polynomial.rs:79-93)."""
    c = list(c)
    while c and c[-1] == 0:
        c.pop()
    return c


def poly_eval(c: Sequence[int], x: int) -> int:
    """Polynomial::evaluate (polynomial.rs:120-137)."""
    acc = 0
    for coeff in reversed(c):
        acc = (acc * x + coeff) % R_MOD
    return acc


def poly_add(a: Sequence[int], b: Sequence[int]) -> List[int]:
    n = max(len(a), len(b))
    out = [0] * n
    for i, v in enumerate(a):
        out[i] = v
    for i, v in enumerate(b):
        out[i] = (out[i] + v) % R_MOD
    return poly_trim(out)


def poly_scale(a: Sequence[int], k: int) -> List[int]:
    return poly_trim([(v * k) % R_MOD for v in a])


def ruffini(c: Sequence[int], z: int) -> List[int]:
    """Polynomial::ruffini (polynomial.rs:345-367): quotient of division by (X - z)."""
    quotient = []
    k = 0
    for coeff in reversed(c):
        t = (coeff + k) % R_MOD
        quotient.append(t)
        k = z * t % R_MOD
    if quotient:
        quotient.pop()
    quotient.reverse()
    return poly_trim(quotient)


def batch_inversion(v: List[int]) -> None:
    """util::batch_inversion (src/util.rs:87-118); zeros are left untouched."""
    for i, x in enumerate(v):
        if x:
            v[i] = fr_inv(x)


# --------------------------------------------------------------------------------------------
# EvaluationDomain (src/fft/domain.rs)
# --------------------------------------------------------------------------------------------
def bitreverse(n: int, l: int) -> int:
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def serial_fft(a: List[int], omega: int, log_n: int) -> None:
    """serial_fft / best_fft (domain.rs:383-463): bit-reverse, then log_n DIT stages whose
    twiddles are a running product (domain.rs:472-489)."""
    n = len(a)
    assert n == 1 << log_n
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    m = 1
    for _ in range(log_n):
        w_m = pow(omega, n // (2 * m), R_MOD)
        for start in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = a[start + j + m] * w % R_MOD
                a[start + j + m] = (a[start + j] - t) % R_MOD
                a[start + j] = (a[start + j] + t) % R_MOD
                w = w * w_m % R_MOD
        m *= 2


def dft_naive(a: Sequence[int], omega: int) -> List[int]:
    """Definition of the transform: out[k] = sum_j a[j] * omega^(j k)."""
    n = len(a)
    return [sum(a[j] * pow(omega, j * k, R_MOD) for j in range(n)) % R_MOD for k in range(n)]


class EvaluationDomain:
    """EvaluationDomain::new (domain.rs:122-158)."""

    def __init__(self, num_coeffs: int):
        size = 1 if num_coeffs <= 1 else 1 << (num_coeffs - 1).bit_length()
        log = size.bit_length() - 1
        if log >= TWO_ADACITY:
            raise ValueError("InvalidEvalDomainSize")
        self.size = size
        self.log_size_of_group = log
        g = ROOT_OF_UNITY
        for _ in range(log, TWO_ADACITY):
            g = g * g % R_MOD
        self.group_gen = g
        self.group_gen_inv = fr_inv(g)
        self.size_inv = fr_inv(size % R_MOD)
        self.generator_inv = fr_inv(GENERATOR)

    def _resize(self, v: Sequence[int]) -> List[int]:
        v = list(v)[: self.size]
        return v + [0] * (self.size - len(v))

    def fft(self, coeffs: Sequence[int]) -> List[int]:  # domain.rs:166-176
        a = self._resize(coeffs)
        serial_fft(a, self.group_gen, self.log_size_of_group)
        return a

    def ifft(self, evals: Sequence[int]) -> List[int]:  # domain.rs:179-196
        a = self._resize(evals)
        serial_fft(a, self.group_gen_inv, self.log_size_of_group)
        return [x * self.size_inv % R_MOD for x in a]

    def coset_fft(self, coeffs: Sequence[int]) -> List[int]:  # domain.rs:198-218
        out, p = [], 1
        for c in coeffs:
            out.append(c * p % R_MOD)
            p = p * GENERATOR % R_MOD
        return self.fft(out)

    def coset_ifft(self, evals: Sequence[int]) -> List[int]:  # domain.rs:221-232
        a = self.ifft(evals)
        out, p = [], 1
        for c in a:
            out.append(c * p % R_MOD)
            p = p * self.generator_inv % R_MOD
        return out

    def elements(self) -> List[int]:  # domain.rs:520-539
        out, w = [], 1
        for _ in range(self.size):
            out.append(w)
            w = w * self.group_gen % R_MOD
        return out

    def evaluate_vanishing_polynomial(self, tau: int) -> int:  # domain.rs:286-291
        return (pow(tau, self.size, R_MOD) - 1) % R_MOD

    def first_lagrange_coefficient(self, tau: int) -> int:
        """evaluate_all_lagrange_coefficients(tau)[0] (domain.rs:237-284)."""
        t_size = pow(tau, self.size, R_MOD)
        if t_size == 1:
            return 1 if tau == 1 else 0
        return (t_size - 1) * self.size_inv % R_MOD * fr_inv((tau - 1) % R_MOD) % R_MOD

    def vanishing_poly_over_coset(self, poly_degree: int) -> List[int]:  # domain.rs:340-351
        point = pow(GENERATOR, poly_degree, R_MOD)
        step = pow(self.group_gen, poly_degree, R_MOD)
        out = []
        for _ in range(self.size):
            out.append((point - 1) % R_MOD)
            point = point * step % R_MOD
        return out


# --------------------------------------------------------------------------------------------
# rand 0.8 StdRng (= ChaCha12Rng) with seed_from_u64 (PCG32 expansion); BlsScalar::random
# --------------------------------------------------------------------------------------------
def _rotl32(x: int, n: int) -> int:
    return ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF


class StdRng:
    def __init__(self, key: bytes):
        assert len(key) == 32
        self.key = [int.from_bytes(key[i : i + 4], "little") for i in range(0, 32, 4)]
        self.counter = 0
        self.buf = b""

    @classmethod
    def seed_from_u64(cls, state: int) -> "StdRng":
        MUL, INC = 6364136223846793005, 11634580027462260723
        seed = bytearray()
        for _ in range(8):
            state = (state * MUL + INC) & 0xFFFFFFFFFFFFFFFF
            xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
            rot = state >> 59
            x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF
            seed += x.to_bytes(4, "little")
        return cls(bytes(seed))

    def _block(self) -> bytes:
        c = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574]
        st = c + self.key + [self.counter & 0xFFFFFFFF, (self.counter >> 32) & 0xFFFFFFFF, 0, 0]
        x = list(st)

        def qr(a, b, c_, d):
            x[a] = (x[a] + x[b]) & 0xFFFFFFFF
            x[d] = _rotl32(x[d] ^ x[a], 16)
            x[c_] = (x[c_] + x[d]) & 0xFFFFFFFF
            x[b] = _rotl32(x[b] ^ x[c_], 12)
            x[a] = (x[a] + x[b]) & 0xFFFFFFFF
            x[d] = _rotl32(x[d] ^ x[a], 8)
            x[c_] = (x[c_] + x[d]) & 0xFFFFFFFF
            x[b] = _rotl32(x[b] ^ x[c_], 7)

        for _ in range(6):  # 12 rounds
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        self.counter += 1
        return b"".join(((x[i] + st[i]) & 0xFFFFFFFF).to_bytes(4, "little") for i in range(16))

    def fill_bytes(self, n: int) -> bytes:
        assert n % 4 == 0
        while len(self.buf) < n:
            self.buf += self._block()
        out, self.buf = self.buf[:n], self.buf[n:]
        return out


def fr_random(rng: StdRng) -> int:
    """BlsScalar::random: 64 RNG bytes -> from_bytes_wide (src/util.rs:124-163)."""
    return fr_from_bytes_wide(rng.fill_bytes(64))


def random_nonzero_bls_scalar(rng: StdRng) -> int:  # src/util.rs:59-68
    while True:
        s = fr_random(rng)
        if s:
            return s


# --------------------------------------------------------------------------------------------
# merlin 3.0 Transcript over STROBE-128 / Keccak-f[1600]; TranscriptProtocol (src/transcript.rs)
# --------------------------------------------------------------------------------------------
_KECCAK_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
    0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
    0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
    0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_KECCAK_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def keccak_f1600(state: bytearray) -> None:
    A = [[int.from_bytes(state[8 * (x + 5 * y) : 8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    rol = lambda v, n: ((v << n) | (v >> (64 - n))) & _M64 if n else v
    for rc in _KECCAK_RC:
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        D = [C[(x - 1) % 5] ^ rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = rol(A[x][y], _KECCAK_ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= rc
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y) : 8 * (x + 5 * y) + 8] = A[x][y].to_bytes(8, "little")


class Strobe128:
    RATE = 166
    FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32

    def __init__(self, protocol_label: bytes):
        st = bytearray(200)
        st[0:6] = bytes([1, self.RATE + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600(st)
        self.state, self.pos, self.pos_begin, self.cur_flags = st, 0, 0, 0
        self.meta_ad(protocol_label, False)

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[self.RATE + 1] ^= 0x80
        keccak_f1600(self.state)
        self.pos = self.pos_begin = 0

    def _absorb(self, data: bytes):
        for b in data:
            self.state[self.pos] ^= b
            self.pos += 1
            if self.pos == self.RATE:
                self._run_f()

    def _squeeze(self, n: int) -> bytes:
        out = bytearray()
        for _ in range(n):
            out.append(self.state[self.pos])
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == self.RATE:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags: int, more: bool):
        if more:
            assert self.cur_flags == flags
            return
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if flags & (self.FLAG_C | self.FLAG_K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data: bytes, more: bool):
        self._begin_op(self.FLAG_M | self.FLAG_A, more)
        self._absorb(data)

    def ad(self, data: bytes, more: bool):
        self._begin_op(self.FLAG_A, more)
        self._absorb(data)

    def prf(self, n: int, more: bool) -> bytes:
        self._begin_op(self.FLAG_I | self.FLAG_A | self.FLAG_C, more)
        return self._squeeze(n)


class Transcript:
    def __init__(self, label: bytes):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label: bytes, message: bytes):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(len(message).to_bytes(4, "little"), True)
        self.strobe.ad(message, False)

    def append_u64(self, label: bytes, x: int):
        self.append_message(label, x.to_bytes(8, "little"))

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(n.to_bytes(4, "little"), True)
        return self.strobe.prf(n, False)

    # TranscriptProtocol (src/transcript.rs:89-108)
    def append_commitment(self, label: bytes, comm: Affine):
        self.append_message(label, g1_compress(comm))

    def append_scalar(self, label: bytes, s: int):
        self.append_message(label, fr_to_bytes(s))

    def challenge_scalar(self, label: bytes) -> int:
        return fr_from_bytes_wide(self.challenge_bytes(label, 64))

    def circuit_domain_sep(self, n: int):
        self.append_message(b"dom-sep", b"circuit_size")
        self.append_u64(b"n", n)


# --------------------------------------------------------------------------------------------
# PublicParameters::setup / trim (src/commitment_scheme/kzg10/srs.rs:61-100, 188-196)
# --------------------------------------------------------------------------------------------
ADDED_BLINDING_DEGREE = 6


def srs_setup(max_degree: int, rng: StdRng, keep: Optional[int] = None) -> List[Affine]:
    """Returns commit_key.powers_of_g.  The G2 draw is performed to keep the RNG stream aligned.
    ``keep`` (optional) computes only the first ``keep`` powers - value-identical to computing all
    max_degree+7 and trimming, which is all the prover ever does with them (srs.rs:188-196)."""
    assert max_degree >= 1
    max_degree += ADDED_BLINDING_DEGREE
    x = random_nonzero_bls_scalar(rng)
    g = jac_mul(jac_from_affine(G1_GEN), random_nonzero_bls_scalar(rng))
    random_nonzero_bls_scalar(rng)  # random_g2_point (srs.rs:91)
    pts, p = [], 1
    for _ in range(min(max_degree + 1, keep or (max_degree + 1))):
        pts.append(jac_mul(g, p))
        p = p * x % R_MOD
    return batch_normalize(pts)


def srs_from_secret(n_points: int, x: int, g_scalar: int) -> List[Affine]:
    """Same shape as srs_setup but from explicit secrets and with incremental point arithmetic
    (used to build larger synthetic SRSs quickly; value-identical to [x^i] g)."""
    g = jac_mul(jac_from_affine(G1_GEN), g_scalar)
    pts, p = [], 1
    for _ in range(n_points):
        pts.append(jac_mul(g, p))
        p = p * x % R_MOD
    return batch_normalize(pts)


def srs_trim(powers_of_g: Sequence[Affine], truncated_degree: int) -> List[Affine]:
    """PublicParameters::trim + CommitKey::truncate (srs.rs:188-196, key.rs:336-355)."""
    d = truncated_degree + ADDED_BLINDING_DEGREE
    if d == 0:
        raise ValueError("TruncatedDegreeIsZero")
    if d > len(powers_of_g) - 1:
        raise ValueError("TruncatedDegreeTooLarge")
    if d == 1:
        d += 1
    return list(powers_of_g[: d + 1])


# --------------------------------------------------------------------------------------------
# Composer subset (src/composer.rs): witnesses, width-4 gates with 11 selectors, permutation map
# --------------------------------------------------------------------------------------------
SELECTORS = ["q_m", "q_l", "q_r", "q_o", "q_f", "q_c", "q_arith", "q_range", "q_logic", "q_fixed_group_add", "q_variable_group_add"]


@dataclass
class Gate:
    sel: Dict[str, int]
    a: int = 0
    b: int = 0
    c: int = 0
    d: int = 0


@dataclass
class Composer:
    constraints: List[Gate] = field(default_factory=list)
    public_inputs: Dict[int, int] = field(default_factory=dict)
    witnesses: List[int] = field(default_factory=list)
    witness_map: List[List[Tuple[int, int]]] = field(default_factory=list)  # witness -> [(column, gate)]

    ZERO = 0
    ONE = 1

    @classmethod
    def initialized(cls) -> "Composer":
        """Composer::initialized + append_dummy_gates (composer.rs:177-240)."""
        s = cls()
        zero = s.append_witness(0)
        one = s.append_witness(1)
        s.assert_equal_constant(zero, 0)
        s.assert_equal_constant(one, 1)
        six = s.append_witness(6)
        one_ = s.append_witness(1)
        seven = s.append_witness(7)
        min_twenty = s.append_witness((-20) % R_MOD)
        s.append_gate(dict(q_m=1, q_l=2, q_r=3, q_f=1, q_c=4, q_o=4), a=six, b=seven, d=one_, c=min_twenty)
        s.append_gate(dict(q_m=1, q_l=1, q_r=1, q_c=127, q_o=1), a=min_twenty, b=six, c=seven)
        return s

    def append_witness(self, v: int) -> int:
        self.witnesses.append(v % R_MOD)
        self.witness_map.append([])
        return len(self.witnesses) - 1

    def append_custom_gate(self, sel: Dict[str, int], a=0, b=0, c=0, d=0, public: Optional[int] = None):
        """append_custom_gate_internal (composer.rs:113-165)."""
        n = len(self.constraints)
        full = {k: 0 for k in SELECTORS}
        for k, v in sel.items():
            full[k] = v % R_MOD
        self.constraints.append(Gate(full, a, b, c, d))
        if public is not None:
            self.public_inputs[n] = public % R_MOD
        for col, w in enumerate((a, b, c, d)):
            self.witness_map[w].append((col, n))

    def append_gate(self, sel: Dict[str, int], a=0, b=0, c=0, d=0, public: Optional[int] = None):
        """append_gate = Constraint::arithmetic (q_arith = 1) (composer.rs:268-276)."""
        sel = dict(sel)
        sel["q_arith"] = 1
        self.append_custom_gate(sel, a, b, c, d, public)

    def assert_equal_constant(self, a: int, constant: int, public: Optional[int] = None):
        self.append_gate(dict(q_l=-1, q_c=constant), a=a, public=public)  # composer.rs:385-400

    def assert_equal(self, a: int, b: int):
        self.append_gate(dict(q_l=1, q_r=-1), a=a, b=b)  # composer.rs:373-379

    def append_public(self, public: int) -> int:
        w = self.append_witness(public)
        self.append_gate(dict(q_l=-1), a=w, public=public)  # composer.rs:358-370
        return w

    def gate_evaluated(self, sel: Dict[str, int], a=0, b=0, d=0, public: Optional[int] = None) -> int:
        """gate_add / gate_mul: q_o = -1 and c solved from the inputs (composer.rs:298-352, 402-417)."""
        W = self.witnesses
        g = lambda k: sel.get(k, 0) % R_MOD
        x = (g("q_m") * W[a] * W[b] + g("q_l") * W[a] + g("q_r") * W[b] + g("q_f") * W[d] + g("q_c") + (public or 0)) % R_MOD
        c = self.append_witness(x)
        s = dict(sel)
        s["q_o"] = -1
        self.append_gate(s, a=a, b=b, c=c, d=d, public=public)
        return c

    def public_input_indexes(self) -> List[int]:
        return sorted(self.public_inputs)

    def public_inputs_vec(self) -> List[int]:
        return [self.public_inputs[i] for i in self.public_input_indexes()]


def compute_sigma_permutations(comp: Composer, n: int) -> List[List[Tuple[int, int]]]:
    """Permutation::compute_sigma_permutations (src/composer/permutation.rs:106-141)."""
    sigmas = [[(col, i) for i in range(n)] for col in range(4)]
    for wires in comp.witness_map:
        for idx, (col, gate) in enumerate(wires):
            sigmas[col][gate] = wires[(idx + 1) % len(wires)]
    return sigmas


# --------------------------------------------------------------------------------------------
# Compiler::preprocess (src/compiler.rs:116-461) and Prover::new (src/compiler/prover.rs:53-115)
# --------------------------------------------------------------------------------------------
CIRCUIT_SIZE_PADDING = 6


@dataclass
class ProverData:
    label: bytes
    constraints: int
    size: int
    commit_key: List[Affine]
    polys: Dict[str, List[int]]  # 11 selectors + s_sigma_1..4, coefficient form (trimmed)
    evals_8n: Dict[str, List[int]]  # the same 15 + "linear" + "v_h", coset evaluations over 8n
    comms: Dict[str, Affine]
    sigma_evals: List[List[int]]  # fft of sigma polys over n
    vanishing_coset_inverses: List[int]


POLY_NAMES = SELECTORS + ["s_sigma_1", "s_sigma_2", "s_sigma_3", "s_sigma_4"]


def compile_circuit(pp: Sequence[Affine], label: bytes, comp: Composer) -> ProverData:
    constraints = len(comp.constraints)
    n_trim = 1 << (constraints + CIRCUIT_SIZE_PADDING - 1).bit_length()
    commit_key = srs_trim(pp, n_trim)
    size = 1 << (constraints - 1).bit_length() if constraints > 1 else 1
    domain = EvaluationDomain(size)
    cols: Dict[str, List[int]] = {k: [0] * size for k in SELECTORS}
    for i, g in enumerate(comp.constraints):
        for k in SELECTORS:
            cols[k][i] = g.sel[k]
    polys = {k: poly_trim(domain.ifft(cols[k])) for k in SELECTORS}
    roots = domain.elements()
    kk = [1, K1, K2, K3]
    sigmas = compute_sigma_permutations(comp, size)
    for j in range(4):
        lag = [kk[col] * roots[idx] % R_MOD for (col, idx) in sigmas[j]]
        polys[f"s_sigma_{j + 1}"] = poly_trim(domain.ifft(lag))
    comms = {}
    for k in SELECTORS:
        try:
            comms[k] = commit(commit_key, polys[k])
        except ValueError:
            comms[k] = None  # unwrap_or_default (compiler.rs:213-227)
    for j in range(4):
        comms[f"s_sigma_{j + 1}"] = commit(commit_key, polys[f"s_sigma_{j + 1}"])
    domain_8n = EvaluationDomain(8 * size)
    evals = {k: domain_8n.coset_fft(polys[k]) for k in POLY_NAMES}
    evals["linear"] = domain_8n.coset_fft([0, 1])
    evals["v_h"] = domain_8n.vanishing_poly_over_coset(size)
    inv = list(evals["v_h"][:8])
    batch_inversion(inv)
    sigma_evals = [domain.fft(polys[f"s_sigma_{j + 1}"]) for j in range(4)]
    return ProverData(label, constraints, size, commit_key, polys, evals, comms, sigma_evals, inv)


def base_transcript_v3(pd: ProverData) -> Transcript:
    """Transcript::base_v3 + VerifierKey::seed_transcript (transcript.rs:131-145, widget.rs:218-257)."""
    t = Transcript(pd.label)
    t.circuit_domain_sep(pd.constraints)
    for lab, key in [
        (b"q_m", "q_m"), (b"q_l", "q_l"), (b"q_r", "q_r"), (b"q_o", "q_o"), (b"q_c", "q_c"), (b"q_f", "q_f"),
        (b"q_arith", "q_arith"), (b"q_range", "q_range"), (b"q_logic", "q_logic"),
        (b"q_variable_group_add", "q_variable_group_add"), (b"q_fixed_group_add", "q_fixed_group_add"),
        (b"s_sigma_1", "s_sigma_1"), (b"s_sigma_2", "s_sigma_2"), (b"s_sigma_3", "s_sigma_3"), (b"s_sigma_4", "s_sigma_4"),
    ]:
        t.append_commitment(lab, pd.comms[key])
    t.circuit_domain_sep(pd.constraints)
    return t


# --------------------------------------------------------------------------------------------
# Gate widgets (src/proof_system/widget/**/proverkey.rs): quotient terms and linearisation scalars
# --------------------------------------------------------------------------------------------
def _delta(f: int) -> int:
    return f * (f - 1) * (f - 2) * (f - 3) % R_MOD


def _delta_xor_and(a: int, b: int, w: int, c: int, q_c: int) -> int:  # logic/proverkey.rs:120-144
    F = w * (w * (4 * w - 18 * (a + b) + 81) + 18 * (a * a + b * b) - 81 * (a + b) + 83) % R_MOD
    E = (3 * (a + b + c) - 2 * F) % R_MOD
    B = q_c * (9 * c - 3 * (a + b)) % R_MOD
    return (B + E) % R_MOD


def widget_arith(q, a, b, c, d) -> int:  # arithmetic/proverkey.rs:44-69
    return (a * b * q["q_m"] + a * q["q_l"] + b * q["q_r"] + c * q["q_o"] + d * q["q_f"] + q["q_c"]) * q["q_arith"] % R_MOD


def widget_range_scalar(ch, a, b, c, d, d_w) -> int:  # range/proverkey.rs:32-57 without the selector
    kappa = ch * ch % R_MOD
    k2 = kappa * kappa % R_MOD
    k3 = k2 * kappa % R_MOD
    return (_delta(c - 4 * d) + _delta(b - 4 * c) * kappa + _delta(a - 4 * b) * k2 + _delta(d_w - 4 * a) * k3) % R_MOD * ch % R_MOD


def widget_logic_scalar(ch, q_c, a, a_w, b, b_w, c, d, d_w) -> int:  # logic/proverkey.rs:34-71
    kappa = ch * ch % R_MOD
    k2 = kappa * kappa % R_MOD
    k3 = k2 * kappa % R_MOD
    k4 = k3 * kappa % R_MOD
    A = (a_w - 4 * a) % R_MOD
    B = (b_w - 4 * b) % R_MOD
    D = (d_w - 4 * d) % R_MOD
    c0 = _delta(A)
    c1 = _delta(B) * kappa
    c2 = _delta(D) * k2
    c3 = (c - A * B) * k3
    c4 = _delta_xor_and(A, B, c, D, q_c) * k4
    return (c0 + c1 + c2 + c3 + c4) % R_MOD * ch % R_MOD


def widget_fixed_base_scalar(ch, q_l, q_r, q_c, a, a_w, b, b_w, c, d, d_w) -> int:  # fixed_base/proverkey.rs:39-103
    kappa = ch * ch % R_MOD
    k2 = kappa * kappa % R_MOD
    k3 = k2 * kappa % R_MOD
    bit = (d_w - d - d) % R_MOD
    bit_consistency = bit * (bit - 1) * (bit + 1) % R_MOD
    y_alpha = (bit * bit * (q_r - 1) + 1) % R_MOD
    x_alpha = bit * q_l % R_MOD
    xy_consistency = (bit * q_c - c) * kappa % R_MOD
    x_acc = ((a_w + a_w * c * a * b * EDWARDS_D) - (a * y_alpha + b * x_alpha)) * k2 % R_MOD
    y_acc = ((b_w - b_w * c * a * b * EDWARDS_D) - (b * y_alpha + a * x_alpha)) * k3 % R_MOD
    return (bit_consistency + x_acc + y_acc + xy_consistency) % R_MOD * ch % R_MOD


def widget_curve_add_scalar(ch, a, a_w, b, b_w, c, d, d_w) -> int:  # curve_addition/proverkey.rs:33-79
    kappa = ch * ch % R_MOD
    x1, x3, y1, y3, x2, y2, x1y2 = a, a_w, b, b_w, c, d, d_w
    xy = (x1 * y2 - x1y2) % R_MOD
    y1x2, y1y2, x1x2 = y1 * x2 % R_MOD, y1 * y2 % R_MOD, x1 * x2 % R_MOD
    x3c = ((x1y2 + y1x2) - (x3 + x3 * EDWARDS_D * x1y2 * y1x2)) * kappa % R_MOD
    y3c = ((y1y2 + x1x2) - (y3 - y3 * EDWARDS_D * x1y2 * y1x2)) * kappa * kappa % R_MOD
    return (xy + x3c + y3c) % R_MOD * ch % R_MOD


def quotient_numerator_i(q: Dict[str, int], ch: Dict[str, int], a, b, c, d, a_w, b_w, d_w, z, z_w, pi, l1_alpha_sq) -> int:
    """One point of t_1 + t_2 (src/proof_system/quotient_poly.rs:160-310).  ``q`` holds the 17
    prover-key evaluations at this point (15 polys + 'linear')."""
    t = widget_arith(q, a, b, c, d)
    t += widget_range_scalar(ch["range"], a, b, c, d, d_w) * q["q_range"]
    t += widget_logic_scalar(ch["logic"], q["q_c"], a, a_w, b, b_w, c, d, d_w) * q["q_logic"]
    t += widget_fixed_base_scalar(ch["fixed"], q["q_l"], q["q_r"], q["q_c"], a, a_w, b, b_w, c, d, d_w) * q["q_fixed_group_add"]
    t += widget_curve_add_scalar(ch["var"], a, a_w, b, b_w, c, d, d_w) * q["q_variable_group_add"]
    t += pi
    alpha, beta, gamma = ch["alpha"], ch["beta"], ch["gamma"]
    x = q["linear"]
    ident = (a + beta * x + gamma) * (b + beta * K1 * x + gamma) % R_MOD * (c + beta * K2 * x + gamma) % R_MOD * (d + beta * K3 * x + gamma) % R_MOD * z % R_MOD * alpha
    copy = (a + beta * q["s_sigma_1"] + gamma) * (b + beta * q["s_sigma_2"] + gamma) % R_MOD * (c + beta * q["s_sigma_3"] + gamma) % R_MOD * (d + beta * q["s_sigma_4"] + gamma) % R_MOD * z_w % R_MOD * alpha
    t += ident - copy + (z - 1) * l1_alpha_sq
    return t % R_MOD


def compute_permutation_vec(domain: EvaluationDomain, wires, beta, gamma, sigma_evals) -> List[int]:
    """Permutation::compute_permutation_vec (src/composer/permutation.rs:213-294)."""
    n = domain.size
    roots = domain.elements()
    out, product = [], 1
    for i in range(n):
        out.append(product)
        if i + 1 < n:
            br = beta * roots[i] % R_MOD
            num = (wires[0][i] + br + gamma) * (wires[1][i] + br * K1 + gamma) % R_MOD * (wires[2][i] + br * K2 + gamma) % R_MOD * (wires[3][i] + br * K3 + gamma) % R_MOD
            den = 1
            for j in range(4):
                den = den * (wires[j][i] + beta * sigma_evals[j][i] + gamma) % R_MOD
            assert den != 0, "permutation denominator must be nonzero"
            product = product * num % R_MOD * fr_inv(den) % R_MOD
    return out


def blind_poly(domain: EvaluationDomain, witnesses: Sequence[int], blinders: Sequence[int]) -> List[int]:
    """Prover::blind_poly_with_blinders (src/compiler/prover.rs:139-152)."""
    coeffs = domain.ifft(witnesses)
    for i, b in enumerate(blinders):
        coeffs[i] = (coeffs[i] - b) % R_MOD
        coeffs.append(b)
    return poly_trim(coeffs)


def compute_barycentric_eval(evaluations: Sequence[int], point: int, domain: EvaluationDomain) -> int:
    """proof::alloc::compute_barycentric_eval (src/proof_system/proof.rs:1042-1094)."""
    numerator = (pow(point, domain.size, R_MOD) - 1) * domain.size_inv % R_MOD
    total = 0
    for i, ev in enumerate(evaluations):
        if ev:
            den = (pow(domain.group_gen_inv, i, R_MOD) * point - 1) % R_MOD
            total += fr_inv(den) * ev
    return total % R_MOD * numerator % R_MOD


@dataclass
class ProofTrace:
    """Everything the GPU prover is compared against, stage by stage."""
    proof_bytes: bytes = b""
    values: Dict[str, object] = field(default_factory=dict)


def prove(pd: ProverData, rng: StdRng, comp: Composer, trace: Optional[ProofTrace] = None) -> bytes:
    """Prover::prove_inner, PlonkVersion::V3 (src/compiler/prover.rs:415-761).  Returns
    Proof::to_bytes (src/proof_system/proof.rs:137-162): 11 x 48 B commitments + 15 x 32 B evals."""
    if len(comp.constraints) != pd.constraints:
        raise ValueError("InvalidCircuitSize")
    tr = trace.values if trace is not None else {}
    size = pd.size
    domain = EvaluationDomain(size)
    domain_8n = EvaluationDomain(8 * size)
    ck = pd.commit_key
    transcript = base_transcript_v3(pd)
    public_inputs = comp.public_inputs_vec()
    pi_idx = comp.public_input_indexes()
    dense_pi = [0] * size
    for i, v in zip(pi_idx, public_inputs):
        dense_pi[i] = v
    for pi in public_inputs:
        transcript.append_scalar(b"pi", pi)

    # round 1
    wires = [[0] * size for _ in range(4)]
    for i, g in enumerate(comp.constraints):
        wires[0][i], wires[1][i], wires[2][i], wires[3][i] = (comp.witnesses[w] for w in (g.a, g.b, g.c, g.d))
    blinders = [[fr_random(rng), fr_random(rng)] for _ in range(4)]
    w_polys = [blind_poly(domain, wires[j], blinders[j]) for j in range(4)]
    w_comms = [commit(ck, p) for p in w_polys]
    for lab, cm in zip((b"a_comm", b"b_comm", b"c_comm", b"d_comm"), w_comms):
        transcript.append_commitment(lab, cm)
    tr["wire_polys"] = w_polys
    tr["wire_comms"] = w_comms

    # round 2
    beta = transcript.challenge_scalar(b"beta")
    transcript.append_scalar(b"beta", beta)
    gamma = transcript.challenge_scalar(b"gamma")
    perm = compute_permutation_vec(domain, wires, beta, gamma, pd.sigma_evals)
    z_poly = blind_poly(domain, perm, [fr_random(rng) for _ in range(3)])
    z_comm = commit(ck, z_poly)
    transcript.append_commitment(b"z_comm", z_comm)
    tr.update(beta=beta, gamma=gamma, perm=perm, z_poly=z_poly, z_comm=z_comm)

    # round 3
    alpha = transcript.challenge_scalar(b"alpha")
    ch = dict(alpha=alpha, beta=beta, gamma=gamma)
    ch["range"] = transcript.challenge_scalar(b"range separation challenge")
    ch["logic"] = transcript.challenge_scalar(b"logic separation challenge")
    ch["fixed"] = transcript.challenge_scalar(b"fixed base separation challenge")
    ch["var"] = transcript.challenge_scalar(b"variable base separation challenge")
    tr.update(alpha=alpha, ch=ch)  # recorded before the quotient so that a rejected circuit still leaves them in the trace
    pi_poly = poly_trim(domain.ifft(dense_pi))
    # quotient_poly::compute (src/proof_system/quotient_poly.rs:20-137)
    n8 = 8 * size
    z8 = domain_8n.coset_fft(z_poly)
    a8, b8, c8, d8 = (domain_8n.coset_fft(p) for p in w_polys)
    pi8 = domain_8n.coset_fft(pi_poly)
    ev = pd.evals_8n
    l1_alpha_sq = alpha * alpha % R_MOD
    proving_domain_size_inv = domain_8n.size_inv * 8 % R_MOD
    quotient = []
    for i in range(n8):
        iw = (i + 8) % n8
        q = {k: ev[k][i] for k in POLY_NAMES}
        q["linear"] = ev["linear"][i]
        l1 = fr_inv((ev["linear"][i] - 1) % R_MOD) * ev["v_h"][i] % R_MOD * proving_domain_size_inv % R_MOD
        t = quotient_numerator_i(q, ch, a8[i], b8[i], c8[i], d8[i], a8[iw], b8[iw], d8[iw], z8[i], z8[iw], pi8[i], l1 * l1_alpha_sq % R_MOD)
        quotient.append(t * pd.vanishing_coset_inverses[i & 7] % R_MOD)
    t_poly = poly_trim(domain_8n.coset_ifft(quotient))
    if len(t_poly) > 7 * size:
        raise ValueError("CircuitUnsatisfied")
    tr["t_poly"] = t_poly
    t_full = t_poly + [0] * max(0, 3 * size - len(t_poly))
    t_low, t_mid, t_high, t_fourth = t_full[:size], t_full[size : 2 * size], t_full[2 * size : 3 * size], t_full[3 * size :]
    t_low += [0] * (size - len(t_low)); t_mid += [0] * (size - len(t_mid)); t_high += [0] * (size - len(t_high))
    b12, b13, b14 = fr_random(rng), fr_random(rng), fr_random(rng)
    t_low.append(b12)
    t_mid[0] = (t_mid[0] - b12) % R_MOD
    t_mid.append(b13)
    t_high[0] = (t_high[0] - b13) % R_MOD
    t_high.append(b14)
    if not t_fourth:
        raise IndexError("t_fourth_vec[0]")  # the reference would panic here (prover.rs:569)
    t_fourth = list(t_fourth)
    t_fourth[0] = (t_fourth[0] - b14) % R_MOD
    t_polys = [poly_trim(t_low), poly_trim(t_mid), poly_trim(t_high), poly_trim(t_fourth)]
    t_comms = [commit(ck, p) for p in t_polys]
    for lab, cm in zip((b"t_low_comm", b"t_mid_comm", b"t_high_comm", b"t_fourth_comm"), t_comms):
        transcript.append_commitment(lab, cm)
    tr.update(alpha=alpha, ch=ch, t_polys=t_polys, t_comms=t_comms)

    # round 4
    z_ch = transcript.challenge_scalar(b"z_challenge")
    zw = z_ch * domain.group_gen % R_MOD
    P = pd.polys
    e = {}
    e["a"], e["b"], e["c"], e["d"] = (poly_eval(p, z_ch) for p in w_polys)
    e["s1"], e["s2"], e["s3"] = (poly_eval(P[f"s_sigma_{j}"], z_ch) for j in (1, 2, 3))
    e["z"] = poly_eval(z_poly, zw)
    for lab, k in ((b"a_eval", "a"), (b"b_eval", "b"), (b"c_eval", "c"), (b"d_eval", "d"), (b"s_sigma_1_eval", "s1"),
                   (b"s_sigma_2_eval", "s2"), (b"s_sigma_3_eval", "s3"), (b"z_eval", "z")):
        transcript.append_scalar(lab, e[k])
    e["a_w"], e["b_w"], e["d_w"] = poly_eval(w_polys[0], zw), poly_eval(w_polys[1], zw), poly_eval(w_polys[3], zw)
    e["q_arith"], e["q_c"], e["q_l"], e["q_r"] = (poly_eval(P[k], z_ch) for k in ("q_arith", "q_c", "q_l", "q_r"))
    for lab, k in ((b"a_w_eval", "a_w"), (b"b_w_eval", "b_w"), (b"d_w_eval", "d_w"), (b"q_arith_eval", "q_arith"),
                   (b"q_c_eval", "q_c"), (b"q_l_eval", "q_l"), (b"q_r_eval", "q_r")):
        transcript.append_scalar(lab, e[k])
    tr.update(z_challenge=z_ch, evals=dict(e))

    # round 5: linearization_poly::compute (src/proof_system/linearization_poly.rs:168-264)
    v_ch = transcript.challenge_scalar(b"v_challenge")
    r = poly_scale(P["q_m"], e["a"] * e["b"] % R_MOD)
    r = poly_add(r, poly_scale(P["q_l"], e["a"]))
    r = poly_add(r, poly_scale(P["q_r"], e["b"]))
    r = poly_add(r, poly_scale(P["q_o"], e["c"]))
    r = poly_add(r, poly_scale(P["q_f"], e["d"]))
    r = poly_add(r, P["q_c"])
    r = poly_scale(r, e["q_arith"])
    r = poly_add(r, poly_scale(P["q_range"], widget_range_scalar(ch["range"], e["a"], e["b"], e["c"], e["d"], e["d_w"])))
    r = poly_add(r, poly_scale(P["q_logic"], widget_logic_scalar(ch["logic"], e["q_c"], e["a"], e["a_w"], e["b"], e["b_w"], e["c"], e["d"], e["d_w"])))
    r = poly_add(r, poly_scale(P["q_fixed_group_add"], widget_fixed_base_scalar(ch["fixed"], e["q_l"], e["q_r"], e["q_c"], e["a"], e["a_w"], e["b"], e["b_w"], e["c"], e["d"], e["d_w"])))
    r = poly_add(r, poly_scale(P["q_variable_group_add"], widget_curve_add_scalar(ch["var"], e["a"], e["a_w"], e["b"], e["b_w"], e["c"], e["d"], e["d_w"])))
    pi_eval = compute_barycentric_eval(public_inputs, z_ch, domain)
    r = poly_add(r, [pi_eval])
    # permutation/proverkey.rs:128-270
    bz = beta * z_ch % R_MOD
    s_ident = (e["a"] + bz + gamma) * (e["b"] + K1 * bz + gamma) % R_MOD * (e["c"] + K2 * bz + gamma) % R_MOD * (e["d"] + K3 * bz + gamma) % R_MOD * alpha % R_MOD
    s_copy = (e["a"] + beta * e["s1"] + gamma) * (e["b"] + beta * e["s2"] + gamma) % R_MOD * (e["c"] + beta * e["s3"] + gamma) % R_MOD * (beta * e["z"] % R_MOD) % R_MOD * alpha % R_MOD
    l1_z = EvaluationDomain(len(z_poly) - 1 - 2).first_lagrange_coefficient(z_ch)
    r = poly_add(r, poly_scale(z_poly, s_ident))
    r = poly_add(r, poly_scale(P["s_sigma_4"], (-s_copy) % R_MOD))
    r = poly_add(r, poly_scale(z_poly, l1_z * alpha % R_MOD * alpha % R_MOD))
    z_n = pow(z_ch, size, R_MOD)
    quot = poly_add(t_polys[0], poly_scale(t_polys[1], z_n))
    quot = poly_add(quot, poly_scale(t_polys[2], z_n * z_n % R_MOD))
    quot = poly_add(quot, poly_scale(t_polys[3], z_n * z_n * z_n % R_MOD))
    z_h_eval = (-domain.evaluate_vanishing_polynomial(z_ch)) % R_MOD
    r_poly = poly_add(r, poly_scale(quot, z_h_eval))
    tr["r_poly"] = r_poly

    def aggregate_witness(polys, point, v):  # key.rs:394-417
        max_len = max(len(p) for p in polys)
        coeffs = [0] * max_len
        power = 1
        for p in polys:
            for i, t in enumerate(p):
                coeffs[i] = (coeffs[i] + t * power) % R_MOD
            power = power * v % R_MOD
        return ruffini(poly_trim(coeffs), point)

    w_z = aggregate_witness([r_poly, *w_polys, P["s_sigma_1"], P["s_sigma_2"], P["s_sigma_3"], P["q_arith"], P["q_c"], P["q_l"], P["q_r"]], z_ch, v_ch)
    w_z_comm = commit(ck, w_z)
    v_w = transcript.challenge_scalar(b"v_w_challenge")
    w_zw = aggregate_witness([z_poly, w_polys[0], w_polys[1], w_polys[3]], zw, v_w)
    w_zw_comm = commit(ck, w_zw)
    tr.update(v_challenge=v_ch, v_w_challenge=v_w, w_z=w_z, w_zw=w_zw, w_z_comm=w_z_comm, w_zw_comm=w_zw_comm)

    out = b"".join(g1_compress(c) for c in (*w_comms, z_comm, *t_comms, w_z_comm, w_zw_comm))
    for k in ("a", "b", "c", "d", "a_w", "b_w", "d_w", "q_arith", "q_c", "q_l", "q_r", "s1", "s2", "s3", "z"):
        out += fr_to_bytes(e[k])
    assert len(out) == 1008
    if trace is not None:
        trace.proof_bytes = out
    return out


# --------------------------------------------------------------------------------------------
# Circuits used by the tests
# --------------------------------------------------------------------------------------------
def minimal_circuit(comp: Composer) -> None:
    """MinimalCircuit of the reference KAT (src/compiler/prover.rs:776-785)."""
    w = comp.append_witness(7)
    comp.assert_equal_constant(w, 7)


def synthetic_arith_circuit(comp: Composer, n_gates: int, seed: int, n_public: int = 2, widgets: int = 0) -> None:
    """a satisfied arithmetic-gate circuit with random witnesses, random
    copy constraints (witness re-use) and a few public inputs, grown until the composer holds
    exactly ``n_gates`` constraints.  Deterministic in ``seed`` (SplitMix64)."""
    state = [seed & 0xFFFFFFFFFFFFFFFF]

    def nxt() -> int:
        state[0] = (state[0] + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = state[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def rfr() -> int:
        return (nxt() | (nxt() << 64) | (nxt() << 128) | (nxt() << 192)) % R_MOD

    if widgets:  # `widgets` rows of every non-arithmetic gate family first
        synthetic_widget_rows(comp, nxt, widgets, widgets, widgets, widgets)
    assert n_gates >= len(comp.constraints) + n_public + 1
    pool = [comp.append_witness(rfr()) for _ in range(4)]
    for _ in range(n_public):
        pool.append(comp.append_public(rfr()))
    while len(comp.constraints) < n_gates:
        a, b, d = (pool[nxt() % len(pool)] for _ in range(3))
        kind = nxt() % 3
        if kind == 0:
            c = comp.gate_evaluated(dict(q_m=rfr(), q_c=rfr()), a=a, b=b, d=d)
        elif kind == 1:
            c = comp.gate_evaluated(dict(q_l=rfr(), q_r=rfr(), q_f=rfr()), a=a, b=b, d=d)
        else:
            c = comp.gate_evaluated(dict(q_m=1, q_l=rfr(), q_f=1, q_c=nxt()), a=a, b=b, d=d)
        pool.append(c)
        if len(pool) > 64:
            pool.pop(nxt() % 32)


# ---- JubJub (twisted Edwards -x^2 + y^2 = 1 + d x^2 y^2 over Fr), only what the ECC gate widgets need
def fr_sqrt(a: int) -> Optional[int]:
    """Tonelli-Shanks over Fr (2-adicity 32)."""
    a %= R_MOD
    if a == 0:
        return 0
    if pow(a, (R_MOD - 1) // 2, R_MOD) != 1:
        return None
    q, s = R_MOD - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = ROOT_OF_UNITY  # generator of the 2^32 subgroup = GENERATOR^q
    m, c, t, r = s, z, pow(a, q, R_MOD), pow(a, (q + 1) // 2, R_MOD)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % R_MOD
            i += 1
        b = pow(c, 1 << (m - i - 1), R_MOD)
        m, c = i, b * b % R_MOD
        t, r = t * c % R_MOD, r * b % R_MOD
    return r


def jubjub_point_from_y(y: int) -> Optional[Tuple[int, int]]:
    num = (y * y - 1) % R_MOD
    den = (1 + EDWARDS_D * y * y) % R_MOD
    x = fr_sqrt(num * fr_inv(den) % R_MOD)
    return None if x is None else (x, y % R_MOD)


def jubjub_add(p: Tuple[int, int], q: Tuple[int, int]) -> Tuple[int, int]:
    x1, y1 = p
    x2, y2 = q
    t = EDWARDS_D * x1 * x2 % R_MOD * y1 % R_MOD * y2 % R_MOD
    x3 = (x1 * y2 + y1 * x2) * fr_inv((1 + t) % R_MOD) % R_MOD
    y3 = (y1 * y2 + x1 * x2) * fr_inv((1 - t) % R_MOD) % R_MOD
    return (x3, y3)


def synthetic_widget_rows(comp: Composer, nxt, n_range: int, n_logic: int, n_fixed: int, n_var: int) -> None:
    """Satisfied rows for every non-arithmetic gate family, written straight from the identities
    their widgets enforce (src/proof_system/widget/{range,logic,ecc/**}/proverkey.rs), so that
    the quotient and linearisation code of all widgets is exercised by a provable circuit.
    `nxt()` is the caller's deterministic 64-bit generator."""
    link = lambda **w: comp.append_custom_gate({}, **w)  # a row with all selectors zero: constrains nothing
    # range: d, c, b, a, d_next form a base-4 accumulator chain (range/proverkey.rs:32-57)
    if n_range:
        acc = nxt() % 4
        d = comp.append_witness(acc)
        for _ in range(n_range):
            ws = []
            for _ in range(3):
                acc = (4 * acc + nxt() % 4) % R_MOD
                ws.append(comp.append_witness(acc))
            comp.append_custom_gate(dict(q_range=1), a=ws[2], b=ws[1], c=ws[0], d=d)
            acc = (4 * acc + nxt() % 4) % R_MOD
            d = comp.append_witness(acc)
        link(d=d)
    # logic: a, b, d accumulate base-4 digits of the operands and of AND / XOR; c holds the product of
    # the two digits (logic/proverkey.rs:34-71)
    if n_logic:
        for q_c in (1, -1):  # AND then XOR (Constraint::logic / logic_xor)
            A = B = D = 0
            wa, wb, wd = (comp.append_witness(0) for _ in range(3))
            for _ in range(n_logic):
                qa, qb = nxt() % 4, nxt() % 4
                wc = comp.append_witness(qa * qb)
                comp.append_custom_gate(dict(q_logic=q_c, q_c=q_c), a=wa, b=wb, c=wc, d=wd)
                A, B = 4 * A + qa, 4 * B + qb
                D = 4 * D + ((qa & qb) if q_c == 1 else (qa ^ qb))
                wa, wb, wd = comp.append_witness(A), comp.append_witness(B), comp.append_witness(D)
            link(a=wa, b=wb, d=wd)
    # JubJub points for the two ECC families
    def point():
        while True:
            pt = jubjub_point_from_y(nxt() | (nxt() << 64) | (nxt() << 128) | ((nxt() >> 3) << 192))
            if pt is not None and pt[0] != 0:
                return pt
    # fixed-base scalar mul rows (ecc/scalar_mul/fixed_base/proverkey.rs:39-103)
    if n_fixed:
        acc_pt, scalar = point(), nxt() % 1000
        wx, wy, wd = comp.append_witness(acc_pt[0]), comp.append_witness(acc_pt[1]), comp.append_witness(scalar)
        for _ in range(n_fixed):
            beta = point()
            bit = (nxt() % 3) - 1
            alpha = (0, 1) if bit == 0 else (beta if bit == 1 else ((-beta[0]) % R_MOD, beta[1]))
            wc = comp.append_witness(alpha[0] * alpha[1])
            comp.append_custom_gate(dict(q_fixed_group_add=1, q_l=beta[0], q_r=beta[1], q_c=beta[0] * beta[1]), a=wx, b=wy, c=wc, d=wd)
            acc_pt = jubjub_add(acc_pt, alpha)
            scalar = (2 * scalar + bit) % R_MOD
            wx, wy, wd = comp.append_witness(acc_pt[0]), comp.append_witness(acc_pt[1]), comp.append_witness(scalar)
        link(a=wx, b=wy, d=wd)
    # variable-base curve addition rows (ecc/curve_addition/proverkey.rs:33-79)
    for _ in range(n_var):
        p1, p2 = point(), point()
        p3 = jubjub_add(p1, p2)
        comp.append_custom_gate(dict(q_variable_group_add=1), a=comp.append_witness(p1[0]), b=comp.append_witness(p1[1]),
                                c=comp.append_witness(p2[0]), d=comp.append_witness(p2[1]))
        link(a=comp.append_witness(p3[0]), b=comp.append_witness(p3[1]), d=comp.append_witness(p1[0] * p2[1]))


KAT_DIGEST = bytes.fromhex(
    "e8564ec22d8cc0ba603626025da3755077aaf0323261908dab68d694736fc273"
    "d31e256cbd3a6a21e7ade63191ac5c9d44a113ac4989a52e4be3abeb1d333237"
)  # src/compiler/prover.rs:1151-1158


def kat_proof(trace: Optional[ProofTrace] = None) -> bytes:
    """deterministic_v3_proof_matches_base_digest (src/compiler/prover.rs:1132-1162)."""
    pp = srs_setup(1 << 10, StdRng.seed_from_u64(0x9235E700), keep=64)
    comp = Composer.initialized()
    minimal_circuit(comp)
    pd = compile_circuit(pp, b"proof-compatibility", comp)
    comp2 = Composer.initialized()
    minimal_circuit(comp2)
    return prove(pd, StdRng.seed_from_u64(0x9235E701), comp2, trace)


if __name__ == "__main__":
    import time

    t0 = time.time()
    proof = kat_proof()
    digest = hashlib.blake2b(proof).digest()
    print("digest", digest.hex())
    print("match ", digest == KAT_DIGEST, f"({time.time() - t0:.2f}s)")
