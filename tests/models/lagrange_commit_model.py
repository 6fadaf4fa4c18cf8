"""Model of wire commitments in the Lagrange basis (TEST INFRASTRUCTURE; imports oracle/ as the checker).

Prover::prove commits to the blinded wire polynomials in the monomial basis (reference
src/compiler/prover.rs:139-152, 187-210): the scalars are interpolated coefficients, full-range whatever
the witness.  The same group element is

    sum_i w_i [L_i(x)]G  +  b0 ([x^n]G - [1]G)  +  b1 ([x^(n+1)]G - [x]G)

because blinding adds b_k X^k (X^n - 1).  [L_j(x)]G = (1/n) sum_i w^(-ij) [x^i]G is an inverse NTT over
group elements of the first n commit-key points, needed once per prover key.  The scalars of that MSM are
the witness values themselves: on the reference's BenchCircuit<2^16> 71 % of the wire slots hold zero and
the rest average one non-zero 16-bit window digit instead of sixteen (tools/witness_digit_stats.py).

Run: python tests/models/lagrange_commit_model.py"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyref as R  # noqa: E402

P = R.R_MOD


def lagrange_key(powers, n):
    """[L_j(x)]G, j < n, from the first n monomial commit-key points (group-element inverse DFT)."""
    dom = R.EvaluationDomain(n)
    w_inv, n_inv = dom.group_gen_inv, dom.size_inv
    out = []
    for j in range(n):
        acc = None
        for i in range(n):
            term = R.jac_mul(R.jac_from_affine(powers[i]), pow(w_inv, i * j, P) * n_inv % P)
            acc = term if acc is None else R.jac_add(acc, term)
        out.append(R.jac_to_affine(acc))
    return out


def commit_lagrange(powers, lag, values, blinders):
    n = len(lag)
    acc = None
    for v, pt in zip(values, lag):
        if v:  # zero witnesses cost nothing
            term = R.jac_mul(R.jac_from_affine(pt), v)
            acc = term if acc is None else R.jac_add(acc, term)
    for k, b in enumerate(blinders):
        for idx, s in ((n + k, b), (k, (P - b) % P)):
            term = R.jac_mul(R.jac_from_affine(powers[idx]), s)
            acc = term if acc is None else R.jac_add(acc, term)
    return R.jac_to_affine(acc)


def check(n=16, seed=3):
    rng = random.Random(seed)
    powers = R.srs_setup(n + 8, R.StdRng.seed_from_u64(seed), keep=n + 7)
    lag = lagrange_key(powers, n)
    dom = R.EvaluationDomain(n)
    for trial in range(3):
        values = [rng.choice([0, 0, 0, 1, 2, 3, rng.randrange(P)]) for _ in range(n)]
        blinders = [rng.randrange(P) for _ in range(2 + trial % 2)]  # wires have two blinders, z has three
        want = R.commit(powers, R.blind_poly(dom, values, blinders))
        assert commit_lagrange(powers, lag, values, blinders) == want
    print(f"n = {n}: Lagrange-basis commitments equal CommitKey::commit of the blinded polynomials")


if __name__ == "__main__":
    check()
    print("ok")
