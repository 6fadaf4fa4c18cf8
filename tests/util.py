"""Shared helpers for the parity tests: seeded inputs and layout conversion (oracle ints <-> ABI bytes)."""
import random

from oracle import pyref as R


def rand_fr(rng: random.Random, n: int):
    return [rng.randrange(R.R_MOD) for _ in range(n)]


def to_abi(v):
    return R.fr_vec_to_mont_bytes(v)


def from_abi(b):
    return R.fr_vec_from_mont_bytes(b)


def progression_bases(n: int, p0: int, step: int):
    """P_i = [p0 + i*step] G built with one group addition per point (cheap for large n)."""
    P = R.jac_mul(R.jac_from_affine(R.G1_GEN), p0)
    B = R.jac_mul(R.jac_from_affine(R.G1_GEN), step)
    pts = []
    for _ in range(n):
        pts.append(P)
        P = R.jac_add(P, B)
    return R.batch_normalize(pts)


def bases_to_abi(pts):
    return b"".join(R.g1_to_raw_bytes(p) for p in pts)
