"""Python model of a cheaper round 3: the quotient polynomial from a 4n coset instead of the 8n one.

TEST INFRASTRUCTURE (like ntt_model.py): validates, against the oracle's own quotient polynomial, the
algebra a CUDA implementation would use.  Not on the product path; imports oracle/ as the checker.

The reference evaluates the numerator on the 8n coset g*H_8n, divides by Z_H pointwise and
interpolates (src/proof_system/quotient_poly.rs:20-137).  t(X) has at most 4n+7 coefficients
(four parts of n plus the blinding overflow, src/compiler/prover.rs:540-575), so a 4n coset
misses only the top seven:

  1. On g*H_4n (the even points of the 8n coset, so the prover key's 8n tables serve unchanged at
     index 2i) compute N/Z_H and interpolate: u(X) = t(X) mod (X^4n - g^4n) = t_lo + g^4n t_hi,
     where t = t_lo + X^4n t_hi, deg t_hi <= 6.
  2. At eight further points x_k = h w8^k, h = g*w_8n (odd points of the 8n coset) evaluate every
     polynomial directly (Horner) and form t(x_k) = N(x_k)/Z_H(x_k).  Then
        t_hi(x_k) = (t(x_k) - u(x_k)) / (x_k^4n - g^4n),
     and since deg t_hi < 8 the 8-point inverse DFT on the coset h*H_8 returns its coefficients; the
     eighth one must vanish - that is the divisibility test which replaces the reference's
     `len > 7n` check (a numerator that Z_H does not divide leaves it non-zero).
  3. t_lo = u - g^4n t_hi.

Cost at n = 2^16: 7 transforms of 2^18 instead of 2^19 and 4n instead of 8n quotient points
(-46 % of the Fr products of a proof) against 6 x 16 + 8 extra Horner evaluations (+10 %).

Run: python tests/models/quotient_4n_model.py   (checks several circuits incl. every gate family and a
corrupted witness; a few seconds of pure Python)."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyref as R  # noqa: E402

P = R.R_MOD


def quotient_via_4n(pd, comp, tr):
    """t(X) (trimmed coefficient list) from the traced round-1/2 polynomials and challenges."""
    n = pd.size
    dom, d4, d8 = R.EvaluationDomain(n), R.EvaluationDomain(4 * n), R.EvaluationDomain(8 * n)
    a, b, c, d = tr["wire_polys"]
    z = tr["z_poly"]
    ch, alpha = tr["ch"], tr["alpha"]
    dense_pi = [0] * n
    for i, v in zip(comp.public_input_indexes(), comp.public_inputs_vec()):
        dense_pi[i] = v
    pi = R.poly_trim(dom.ifft(dense_pi))
    g = R.GENERATOR
    l1a2 = alpha * alpha % P
    n_inv = R.fr_inv(n)

    def point(q, x, av, bv, cv, dv, aw, bw, dw, zv, zw, piv):
        zh = (pow(x, n, P) - 1) % P
        l1 = zh * n_inv % P * R.fr_inv((x - 1) % P) % P
        q = dict(q)
        q["linear"] = x
        return R.quotient_numerator_i(q, ch, av, bv, cv, dv, aw, bw, dw, zv, zw, piv, l1 * l1a2 % P) * R.fr_inv(zh) % P

    # 1. the 4n coset: prover-key tables are the even entries of the 8n tables
    n4 = 4 * n
    A, B, C, D, Z, PI = (d4.coset_fft(p) for p in (a, b, c, d, z, pi))
    xs = d4.coset_fft([0, 1])
    quot = []
    for i in range(n4):
        iw = (i + 4) % n4
        q = {k: pd.evals_8n[k][2 * i] for k in R.POLY_NAMES}
        assert xs[i] == pd.evals_8n["linear"][2 * i]
        quot.append(point(q, xs[i], A[i], B[i], C[i], D[i], A[iw], B[iw], D[iw], Z[i], Z[iw], PI[i]))
    u = d4.coset_ifft(quot)

    # 2. eight odd points of the 8n coset, everything by Horner
    w8n, wn, w8 = d8.group_gen, dom.group_gen, pow(d8.group_gen, n, P)
    h = g * w8n % P
    g4n = pow(g, n4, P)
    e = []
    for k in range(8):
        x = h * pow(w8, k, P) % P
        xw = x * wn % P
        assert x == pd.evals_8n["linear"][1 + n * k]
        q = {name: R.poly_eval(pd.polys[name], x) for name in R.POLY_NAMES}
        assert all(q[name] == pd.evals_8n[name][1 + n * k] for name in R.POLY_NAMES)  # or read them from the 8n tables
        ev = lambda p, at: R.poly_eval(p, at)
        t_x = point(q, x, ev(a, x), ev(b, x), ev(c, x), ev(d, x), ev(a, xw), ev(b, xw), ev(d, xw), ev(z, x), ev(z, xw), ev(pi, x))
        e.append((t_x - R.poly_eval(u, x)) * R.fr_inv((pow(x, n4, P) - g4n) % P) % P)
    inv8, w8_inv, h_inv = R.fr_inv(8), R.fr_inv(w8), R.fr_inv(h)
    t_hi = [sum(e[k] * pow(w8_inv, j * k, P) for k in range(8)) % P * inv8 % P * pow(h_inv, j, P) % P for j in range(8)]
    if t_hi[7] != 0:
        raise ValueError("CircuitUnsatisfied")

    # 3. assemble
    t = list(u) + t_hi[:7]
    for j in range(7):
        t[j] = (t[j] - g4n * t_hi[j]) % P
    return R.poly_trim(t)


def check(n_gates, seed, widgets, n_public=2, corrupt=False):
    rng = random.Random(seed)
    comp = R.Composer.initialized()
    R.synthetic_arith_circuit(comp, n_gates, seed=seed, n_public=n_public, widgets=widgets)
    n_trim = 1 << (len(comp.constraints) + R.CIRCUIT_SIZE_PADDING - 1).bit_length()
    pp = R.srs_setup(1 << 12, R.StdRng.seed_from_u64(rng.randrange(1 << 32)), keep=n_trim + 7)
    pd = R.compile_circuit(pp, b"quotient-4n", comp)
    if corrupt:
        comp.witnesses[len(comp.witnesses) // 2] = (comp.witnesses[len(comp.witnesses) // 2] + 1) % P
    trace = R.ProofTrace()
    try:
        R.prove(pd, R.StdRng.seed_from_u64(seed), comp, trace)
        ref_unsat = False
    except ValueError as ex:
        assert "CircuitUnsatisfied" in str(ex)
        ref_unsat = True
    tr = trace.values
    try:
        t = quotient_via_4n(pd, comp, tr)
        ours_unsat = False
    except ValueError:
        ours_unsat = True
    assert ours_unsat == ref_unsat == corrupt, (ours_unsat, ref_unsat, corrupt)
    if not corrupt:
        assert t == tr["t_poly"], "quotient polynomial differs"
        print(f"n = {pd.size:4d} ({len(comp.constraints)} gates, widgets={widgets}): t(X) from the 4n coset + 8 points == reference "
              f"({len(t)} coefficients, 4n = {4 * pd.size})")
    else:
        print(f"n = {pd.size:4d} corrupted witness: both reject (CircuitUnsatisfied)")


if __name__ == "__main__":
    check(20, 1, 0)
    check(60, 2, 7)
    check(120, 3, 7, n_public=0)
    check(60, 4, 7, corrupt=True)
    print("ok")
