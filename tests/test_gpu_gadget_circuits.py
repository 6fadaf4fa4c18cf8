"""Gadget circuits through the GPU prover, modelled on the reference's integration tests
(tests/{boolean,decomposition,range,logic,select_bls,select_point,ecc,truncate,gate_add_mul,
assert_scalar,assert_point}.rs): a circuit is compiled once from its default values
(`Compiler::compile`), then proved with other witnesses.  A satisfying assignment must give the proof
the CPU restatement gives (`check_satisfied_circuit`; the reference then verifies by pairing, which is
out of scope here - byte equality with the oracle replaces it); an unsatisfying one must be refused
with Error::CircuitUnsatisfied by both (`check_unsatisfied_circuit`).

The product side builds every circuit with the native composer (csrc/composer.cpp), the oracle side
with oracle/gadgets.py; the row-by-row identity checker of the oracle says which assignments satisfy."""
import random

import pytest

from oracle import cref
from oracle import gadgets as G
from oracle import pyref as R
from oracle import verify as V

pytestmark = pytest.mark.gpu

R_MOD = R.R_MOD
GEN = G.JUBJUB_GENERATOR
rng = random.Random(0xB200)
RANDOM_FR = rng.randrange(1 << 200, R_MOD)
P7 = G.jj_mul(GEN, 7)
P9 = G.jj_mul(GEN, 9)


def c_boolean(c, bit):  # tests/boolean.rs
    c.component_boolean(c.append_witness(bit))


def c_decomposition(n):  # tests/decomposition.rs: the bits are compared with expected ones
    def build(c, a, expected):
        bits = c.component_decomposition(c.append_witness(a), n)
        for w, e in zip(bits, [(expected >> i) & 1 for i in range(n)]):
            c.assert_equal(w, c.append_witness(e))
    return build


def c_range(bits):  # tests/range.rs
    return lambda c, a: c.component_range_bits(c.append_witness(a), bits)


def c_logic(pairs, xor):  # tests/logic.rs: result compared with an expected witness
    def build(c, a, b, result):
        wa, wb = c.append_witness(a), c.append_witness(b)
        out = (c.append_logic_xor if xor else c.append_logic_and)(wa, wb, pairs)
        c.assert_equal(out, c.append_witness(result))
    return build


def c_select(c, bit, a, b, result):  # tests/select_bls.rs
    out = c.component_select(c.append_witness(bit), c.append_witness(a), c.append_witness(b))
    c.assert_equal(out, c.append_witness(result))


def c_select_one(c, bit, v, result):
    c.assert_equal(c.component_select_one(c.append_witness(bit), c.append_witness(v)), c.append_witness(result))


def c_select_zero(c, bit, v, result):
    c.assert_equal(c.component_select_zero(c.append_witness(bit), c.append_witness(v)), c.append_witness(result))


def c_select_point(c, bit, a, b, result):  # tests/select_point.rs
    out = c.component_select_point(c.append_witness(bit), c.append_point(a), c.append_point(b))
    c.assert_equal_point(out, c.append_point(result))


def c_select_identity(c, bit, a, result):
    out = c.component_select_identity(c.append_witness(bit), c.append_point(a))
    c.assert_equal_point(out, c.append_point(result))


def c_mul_generator(c, k, expected):  # tests/ecc.rs mul_generator: the product is a public input
    out = c.component_mul_generator(c.append_witness(k), GEN)
    c.assert_equal_public_point(out, expected)


def c_mul_point(c, k, p, expected):  # tests/ecc.rs mul_point
    out = c.component_mul_point(c.append_witness(k), c.append_point(p))
    c.assert_equal_public_point(out, expected)


def c_add_point(c, a, b, expected):  # tests/ecc.rs add_point
    out = c.component_add_point(c.append_point(a), c.append_point(b))
    c.assert_equal_public_point(out, expected)


def c_torsion_free(c, p):  # src/composer/tests/soundness/point.rs torsion tests
    c.assert_torsion_free_point(c.append_point(p))


def c_truncate(n):  # tests/truncate.rs
    def build(c, a, expected):
        c.assert_equal(c.component_truncate(c.append_witness(a), n), c.append_witness(expected))
    return build


def c_gate_add_mul(c, a, b, d, pi, result):  # tests/gate_add_mul.rs
    wa, wb, wd = c.append_witness(a), c.append_witness(b), c.append_witness(d)
    s = c.gate_add(dict(q_l=1, q_r=1, q_f=1, q_c=5), a=wa, b=wb, d=wd, public=pi)
    m = c.gate_mul(dict(q_m=3, q_f=2), a=s, b=wb, d=wd)
    c.assert_equal(m, c.append_witness(result))


def c_assert_scalar(c, a, pi):  # tests/assert_scalar.rs: a == 2 + pi, the constant being part of the circuit
    w = c.append_witness(a)
    c.assert_equal_constant(w, 2, public=pi)
    c.assert_equal(w, c.append_public(a))


def gam(a, b, d, pi):
    s = (a + b + d + 5 + pi) % R_MOD
    return (3 * s * b + 2 * d) % R_MOD


T2 = (0, R_MOD - 1)  # the order-2 point
M = (1 << 64) - 1
CASES = [
    # name, builder, default values, satisfying value sets, unsatisfying value sets
    ("boolean", c_boolean, (0,), [(1,), (0,)], [(R_MOD - 1,), (RANDOM_FR,), (2,)]),
    ("decomposition_1", c_decomposition(1), (0, 0), [(1, 1)], [(1, 0), (2, 0)]),
    ("decomposition_64", c_decomposition(64), (0, 0), [(M, M), (0xDEADBEEF, 0xDEADBEEF)], [(M, M - 1), (M + 1, 0)]),
    ("decomposition_252", c_decomposition(252), (0, 0), [((1 << 252) - 1, (1 << 252) - 1)], [(1 << 252, 0)]),
    ("range_0", c_range(0), (0,), [(0,)], [(1,), (RANDOM_FR,)]),
    ("range_2", c_range(2), (0,), [(1,), (3,)], [(4,)]),
    ("range_7", c_range(7), (0,), [(127,)], [(128,)]),
    ("range_74", c_range(74), (0,), [(1 << 73,), ((1 << 74) - 1,)], [(1 << 74,), (RANDOM_FR,)]),
    ("range_256", c_range(256), (0,), [(RANDOM_FR,), (R_MOD - 1,)], []),
    ("logic_and_1", c_logic(1, False), (0, 0, 0), [(3, 2, 2), (7, 5, 1)], [(3, 2, 3)]),
    ("logic_and_32", c_logic(32, False), (0, 0, 0), [(0xFFFF0000FFFF, 0x0F0F0F0F0F0F, 0x0F0F00000F0F), (RANDOM_FR, M, RANDOM_FR & M)],
     [(M, M, M - 1)]),
    ("logic_xor_32", c_logic(32, True), (0, 0, 0), [(0xFFFF0000FFFF, 0x0F0F0F0F0F0F, 0xF0F00F0FF0F0), (RANDOM_FR, 0, RANDOM_FR & M)],
     [(1, 1, 1)]),
    ("logic_xor_127", c_logic(127, True), (0, 0, 0), [(RANDOM_FR, 5, (RANDOM_FR ^ 5) & ((1 << 254) - 1))], [(RANDOM_FR, 5, RANDOM_FR)]),
    ("select", c_select, (0, 0, 0, 0), [(1, 11, 22, 11), (0, 11, 22, 22)], [(1, 11, 22, 22), (0, 11, 22, 11)]),
    ("select_one", c_select_one, (0, 0, 1), [(1, 9, 9), (0, 9, 1)], [(0, 9, 9)]),
    ("select_zero", c_select_zero, (0, 0, 0), [(1, 9, 9), (0, 9, 0)], [(0, 9, 9)]),
    ("select_point", c_select_point, (0, GEN, GEN, GEN), [(1, P7, P9, P7), (0, P7, P9, P9)], [(1, P7, P9, P9)]),
    ("select_identity", c_select_identity, (1, GEN, GEN), [(1, P7, P7), (0, P7, (0, 1))], [(0, P7, P7), (2, P7, P7)]),
    ("mul_generator", c_mul_generator, (1, GEN), [(7, P7), (G.JUBJUB_ORDER - 2, G.jj_mul(GEN, G.JUBJUB_ORDER - 2))], [(7, P9), (9, P7)]),
    ("mul_point", c_mul_point, (1, GEN, GEN), [(9, P7, G.jj_mul(GEN, 63)), (0, P7, (0, 1))], [(9, P7, P9)]),
    ("add_point", c_add_point, (GEN, GEN, G.jj_add(GEN, GEN)), [(P7, P9, G.jj_mul(GEN, 16)), (P7, G.jj_neg(P7), (0, 1))], [(P7, P9, P9)]),
    ("torsion_free", c_torsion_free, (GEN,), [(P9,), ((0, 1),)], [(T2,), (G.jj_add(P7, T2),)]),
    ("truncate_100", c_truncate(100), (0, 0), [(RANDOM_FR, RANDOM_FR & ((1 << 100) - 1)), (R_MOD - 1, (R_MOD - 1) & ((1 << 100) - 1))],
     [(RANDOM_FR, (RANDOM_FR & ((1 << 100) - 1)) ^ 1), (RANDOM_FR, RANDOM_FR)]),
    ("gate_add_mul", c_gate_add_mul, (0, 0, 0, 0, 0), [(3, 4, 5, 6, gam(3, 4, 5, 6)), (RANDOM_FR, 1, 2, R_MOD - 1, gam(RANDOM_FR, 1, 2, R_MOD - 1))],
     [(3, 4, 5, 6, 1)]),
    ("assert_scalar", c_assert_scalar, (5, 3), [(10, 8), (2, 0), (1, R_MOD - 1)], [(5, 4), (6, 3)]),
]


@pytest.fixture(scope="module")
def pb():
    import plonk_b200
    from plonk_b200._lib import check, lib

    check(lib().pb200_init(0))
    return plonk_b200


@pytest.mark.parametrize("name,build,default,satisfied,unsatisfied", CASES, ids=[c[0] for c in CASES])
def test_gadget_circuit(pb, name, build, default, satisfied, unsatisfied):
    from plonk_b200 import gadgets as N

    def both(vals):
        o, n = G.GadgetComposer.initialized(), N.Composer.initialized()
        build(o, *vals)
        build(n, *vals)
        return o, cref.CircuitArrays(o), n.arrays()

    o, oarr, narr = both(default)
    assert G.unsatisfied_rows(o) == [], "the default circuit must be satisfied"
    assert (narr.selectors, narr.wires, narr.witnesses) == (oarr.selectors, oarr.wires, oarr.witnesses)
    n = 1 << (oarr.constraints + 6 - 1).bit_length()  # pp.trim(next_pow2(constraints + 6)), compiler.rs:121-124
    secret = 0x5EED + len(name)
    srs_raw = cref.srs_from_secret(n + 7, secret, 0xACE)
    label = name.encode()
    cpu = cref.CrefProver(label, oarr, srs_raw)
    gpu = pb.Prover(label, narr.constraints, narr.selectors, narr.wires, narr.n_witnesses, srs_raw)
    assert gpu.commitments() == cpu.commitments()
    key_comms = {k: R.g1_decompress(c) for k, c in zip(R.POLY_NAMES, gpu.commitments())}
    g = R.g1_from_raw_bytes(srs_raw[:96])
    for k, vals in enumerate([default] + satisfied):
        o, oa, na = both(vals)
        assert (na.selectors, na.wires) == (narr.selectors, narr.wires), "gate layout must not depend on the witness"
        assert G.unsatisfied_rows(o) == [], vals
        blinders = cref.draw_blinders(R.StdRng.seed_from_u64(100 + k))
        proof = gpu.prove(na.witnesses, na.pi_idx, na.pi_vals, blinders)
        assert proof == cpu.prove(blinders, oa), vals
        # and the GPU-made proof satisfies the reference Verifier's equation (oracle/verify.py)
        assert V.verify_with_secret(proof, label, oarr.constraints, key_comms, o.public_input_indexes(), o.public_inputs_vec(), g, secret), vals
    for k, vals in enumerate(unsatisfied):
        o, oa, na = both(vals)
        assert (na.selectors, na.wires) == (narr.selectors, narr.wires)
        assert G.unsatisfied_rows(o) != [], vals
        blinders = cref.draw_blinders(R.StdRng.seed_from_u64(200 + k))
        with pytest.raises(pb.CircuitUnsatisfied):
            gpu.prove(na.witnesses, na.pi_idx, na.pi_vals, blinders)
        with pytest.raises(ValueError, match="-5"):
            cpu.prove(blinders, oa)
