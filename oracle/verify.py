"""TEST INFRASTRUCTURE - CPU oracle, not product code.

The reference Verifier's equation (`Proof::verify`, src/proof_system/proof.rs:216-516, with the widget
verifier keys' `compute_linearization_commitment`, src/proof_system/widget/**/verifierkey.rs) restated
on top of pyref, for proofs made over a *test* SRS whose secret is known.

The reference ends with the pairing check  e(-(W_z + u W_zw), [x]H) * e(R, H) == 1  where
R = z W_z + u z w W_zw + [F] - [E]  (proof.rs:452-512).  H generates a group of prime order, so the
check is the G1 identity  R == [x](W_z + u W_zw); with the trusted-setup secret x in hand (every SRS in
this repo's tests comes from `srs_from_secret` or a replayable seeded RNG) it needs no pairing.  What it
adds to the byte-parity tests: the prover-side linearisation, evaluations and opening witnesses are
checked against the verifier-side formulas - an error common to the GPU prover and the restated CPU
prover would make proofs that are byte-identical and invalid; this is the test that would see it.

Only tests/ may import this file."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

try:
    from . import pyref as P
except ImportError:
    import pyref as P

R_MOD = P.R_MOD
Affine = Optional[tuple]

EVAL_ORDER = ("a", "b", "c", "d", "a_w", "b_w", "d_w", "q_arith", "q_c", "q_l", "q_r", "s1", "s2", "s3", "z")
COMM_ORDER = ("a", "b", "c", "d", "z", "t_low", "t_mid", "t_high", "t_fourth", "w_z", "w_zw")


def parse_proof(proof: bytes):
    """Proof::from_bytes layout (proof.rs:137-162, linearization_poly.rs:98-124): 11 compressed
    commitments then 15 scalars."""
    assert len(proof) == 1008
    comms = {k: P.g1_decompress(proof[48 * i : 48 * (i + 1)]) for i, k in enumerate(COMM_ORDER)}
    evals = {}
    for i, k in enumerate(EVAL_ORDER):
        v = int.from_bytes(proof[528 + 32 * i : 528 + 32 * (i + 1)], "little")
        assert v < R_MOD, "non-canonical scalar"
        evals[k] = v
    return comms, evals


def _msm(points: Sequence[Affine], scalars: Sequence[int]):
    acc = None
    for p, s in zip(points, scalars):
        if p is None or s % R_MOD == 0:
            continue
        t = P.jac_mul(P.jac_from_affine(p), s % R_MOD)
        acc = t if acc is None else P.jac_add(acc, t)
    return None if acc is None else P.jac_to_affine(acc)


def verify_with_secret(proof: bytes, label: bytes, constraints: int, key_comms: Dict[str, Affine], pi_idx: Sequence[int],
                       pi_vals: Sequence[int], g: Affine, x: int) -> bool:
    """`key_comms`: the 15 verifier-key commitments by pyref.POLY_NAMES name; `g` = opening_key.g =
    powers_of_g[0]; `x` the SRS secret."""
    comm, e = parse_proof(proof)
    n = 1 << (constraints - 1).bit_length() if constraints > 1 else 1
    domain = P.EvaluationDomain(n)

    class _PD:  # what base_transcript_v3 reads
        pass

    pd = _PD()
    pd.label, pd.constraints, pd.comms = label, constraints, key_comms
    t = P.base_transcript_v3(pd)
    for pi in pi_vals:  # Verifier::verify_with_version (compiler/verifier.rs:228-232)
        t.append_scalar(b"pi", pi % R_MOD)
    for k in ("a", "b", "c", "d"):
        t.append_commitment(k.encode() + b"_comm", comm[k])
    beta = t.challenge_scalar(b"beta")
    t.append_scalar(b"beta", beta)
    gamma = t.challenge_scalar(b"gamma")
    t.append_commitment(b"z_comm", comm["z"])
    alpha = t.challenge_scalar(b"alpha")
    ch_range = t.challenge_scalar(b"range separation challenge")
    ch_logic = t.challenge_scalar(b"logic separation challenge")
    ch_fixed = t.challenge_scalar(b"fixed base separation challenge")
    ch_var = t.challenge_scalar(b"variable base separation challenge")
    for k in ("t_low", "t_mid", "t_high", "t_fourth"):
        t.append_commitment(k.encode() + b"_comm", comm[k])
    z_ch = t.challenge_scalar(b"z_challenge")
    for lab, k in ((b"a_eval", "a"), (b"b_eval", "b"), (b"c_eval", "c"), (b"d_eval", "d"), (b"s_sigma_1_eval", "s1"),
                   (b"s_sigma_2_eval", "s2"), (b"s_sigma_3_eval", "s3"), (b"z_eval", "z"), (b"a_w_eval", "a_w"),
                   (b"b_w_eval", "b_w"), (b"d_w_eval", "d_w"), (b"q_arith_eval", "q_arith"), (b"q_c_eval", "q_c"),
                   (b"q_l_eval", "q_l"), (b"q_r_eval", "q_r")):
        t.append_scalar(lab, e[k])
    v = t.challenge_scalar(b"v_challenge")
    v_w = t.challenge_scalar(b"v_w_challenge")
    t.append_commitment(b"w_z_chall_comm", comm["w_z"])
    t.append_commitment(b"w_z_chall_w_comm", comm["w_zw"])
    u = t.challenge_scalar(b"u_challenge")

    z_h = domain.evaluate_vanishing_polynomial(z_ch)
    # compute_lagrange_and_barycentric_evaluations (proof.rs:997-1040)
    if (z_ch - 1) % R_MOD == 0:
        return False
    l1 = z_h * P.fr_inv(n * (z_ch - 1) % R_MOD) % R_MOD
    w_inv = P.fr_inv(domain.group_gen)
    pi_eval = 0
    for idx, val in zip(pi_idx, pi_vals):
        if val % R_MOD == 0:
            continue
        den = (pow(w_inv, idx, R_MOD) * z_ch - 1) % R_MOD
        if den == 0:
            return False
        pi_eval = (pi_eval + val * P.fr_inv(den)) % R_MOD
    pi_eval = pi_eval * z_h % R_MOD * P.fr_inv(n) % R_MOD

    r0 = (pi_eval - l1 * alpha * alpha
          - alpha * (e["a"] + beta * e["s1"] + gamma) * (e["b"] + beta * e["s2"] + gamma) % R_MOD
          * (e["c"] + beta * e["s3"] + gamma) % R_MOD * (e["d"] + gamma) % R_MOD * e["z"]) % R_MOD

    V = 11  # V_MAX_DEGREE
    vc = [v]
    for _ in range(1, V):
        vc.append(vc[-1] * v % R_MOD)
    vc.append(v_w * u % R_MOD)
    vc.append(vc[V] * v_w % R_MOD)
    vc.append(vc[V + 1] * v_w % R_MOD)
    e_list = [e[k] for k in ("a", "b", "c", "d", "s1", "s2", "s3", "q_arith", "q_c", "q_l", "q_r", "a_w", "b_w", "d_w")]
    E = (sum(a * b for a, b in zip(e_list, vc)) - r0 + u * e["z"]) % R_MOD

    scalars: List[int] = []
    points: List[Affine] = []

    def term(s, p):
        scalars.append(s % R_MOD)
        points.append(p)

    K = key_comms
    # arithmetic / range / logic / fixed-base / curve-addition verifier keys
    term(e["a"] * e["b"] * e["q_arith"], K["q_m"])
    term(e["a"] * e["q_arith"], K["q_l"])
    term(e["b"] * e["q_arith"], K["q_r"])
    term(e["c"] * e["q_arith"], K["q_o"])
    term(e["d"] * e["q_arith"], K["q_f"])
    term(e["q_arith"], K["q_c"])
    term(P.widget_range_scalar(ch_range, e["a"], e["b"], e["c"], e["d"], e["d_w"]), K["q_range"])
    term(P.widget_logic_scalar(ch_logic, e["q_c"], e["a"], e["a_w"], e["b"], e["b_w"], e["c"], e["d"], e["d_w"]), K["q_logic"])
    term(P.widget_fixed_base_scalar(ch_fixed, e["q_l"], e["q_r"], e["q_c"], e["a"], e["a_w"], e["b"], e["b_w"], e["c"], e["d"], e["d_w"]),
         K["q_fixed_group_add"])
    term(P.widget_curve_add_scalar(ch_var, e["a"], e["a_w"], e["b"], e["b_w"], e["c"], e["d"], e["d_w"]), K["q_variable_group_add"])
    # permutation verifier key (permutation/verifierkey.rs)
    bz = beta * z_ch % R_MOD
    xs = (e["a"] + bz + gamma) * (e["b"] + P.K1 * bz + gamma) % R_MOD * (e["c"] + P.K2 * bz + gamma) % R_MOD \
        * ((e["d"] + P.K3 * bz + gamma) * alpha % R_MOD) % R_MOD
    term(xs + l1 * alpha * alpha + u, comm["z"])
    ys = (e["a"] + beta * e["s1"] + gamma) * (e["b"] + beta * e["s2"] + gamma) % R_MOD * (e["c"] + beta * e["s3"] + gamma) % R_MOD \
        * (beta * e["z"] % R_MOD * alpha % R_MOD) % R_MOD
    term(-ys, K["s_sigma_4"])
    # quotient chunks (proof.rs:872-888)
    z_pow_n = (z_h + 1) % R_MOD
    term(-z_h, comm["t_low"])
    term(z_pow_n * -z_h, comm["t_mid"])
    term(z_pow_n * z_pow_n * -z_h, comm["t_high"])
    term(z_pow_n * z_pow_n * z_pow_n * -z_h, comm["t_fourth"])
    # [F]: openings at z and, grouped with them, the shifted ones (proof.rs:402-433)
    f = vc[:V]
    f[0] = (f[0] + vc[V]) % R_MOD
    f[1] = (f[1] + vc[V + 1]) % R_MOD
    f[3] = (f[3] + vc[V + 2]) % R_MOD
    for s, p in zip(f, (comm["a"], comm["b"], comm["c"], comm["d"], K["s_sigma_1"], K["s_sigma_2"], K["s_sigma_3"],
                        K["q_arith"], K["q_c"], K["q_l"], K["q_r"])):
        term(s, p)
    term(-E, g)
    term(z_ch, comm["w_z"])
    term(u * z_ch % R_MOD * domain.group_gen, comm["w_zw"])

    right = _msm(points, scalars)
    left = _msm([comm["w_z"], comm["w_zw"]], [1, u])
    want = None if left is None else P.jac_to_affine(P.jac_mul(P.jac_from_affine(left), x % R_MOD))
    return right == want
