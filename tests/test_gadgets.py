"""Circuit gadget library: oracle and product front ends against the
reference's own gate-layout goldens, and against each other on witness values.

The nine `gate_digest` goldens are copied from the reference's soundness suite (file:line on each);
`gate_digest` itself is src/composer/tests/soundness/support.rs:93-135.  The fixed-base golden folds
the coordinates of 2^i * dusk_jubjub::GENERATOR into the digest, which pins the JubJub constants
that live in the un-vendored dusk-jubjub crate.  CPU only: circuit construction involves no GPU."""
import random

import pytest

from oracle import gadgets as G
from oracle import pyref as R
from plonk_b200 import gadgets as N
from plonk_b200._lib import Pb200Error

R_MOD = R.R_MOD

GOLDEN = {
    # src/composer/tests/soundness/range.rs:152-166 (component_range::<16|32|128>(-1))
    "range16": [77, 31, 113, 140, 168, 100, 230, 119, 141, 149, 133, 230, 149, 247, 247, 146, 198, 131, 151, 72, 86, 226, 37, 227, 151, 105, 226, 40, 34, 107, 152, 55],
    "range32": [211, 119, 147, 191, 124, 192, 26, 156, 231, 67, 118, 215, 252, 91, 144, 70, 167, 7, 86, 187, 217, 252, 99, 197, 167, 153, 185, 163, 50, 167, 5, 33],
    "range128": [61, 254, 139, 94, 245, 111, 49, 233, 147, 232, 116, 107, 73, 148, 236, 197, 128, 124, 52, 202, 152, 56, 66, 82, 119, 96, 65, 141, 195, 208, 155, 98],
    # src/composer/tests/soundness/logic.rs:962-971 (append_logic_{xor,and}::<125>(-1, ZERO))
    "xor": [32, 161, 124, 154, 171, 19, 190, 198, 173, 78, 161, 187, 34, 114, 227, 147, 171, 13, 126, 114, 18, 25, 221, 142, 1, 219, 8, 111, 203, 199, 216, 57],
    "and": [155, 86, 135, 38, 200, 186, 164, 136, 179, 4, 230, 133, 119, 57, 60, 185, 191, 172, 130, 234, 223, 96, 62, 50, 62, 203, 224, 5, 39, 207, 143, 32],
    # src/composer/tests/soundness/fixed_base.rs:1016-1019 (component_mul_generator(r_jubjub - 2, GENERATOR))
    "fixed": [12, 248, 174, 80, 4, 183, 76, 71, 51, 243, 231, 56, 142, 43, 223, 49, 71, 66, 186, 118, 187, 39, 149, 99, 2, 10, 183, 18, 145, 85, 227, 83],
    # src/composer/tests/soundness/point.rs:351-354 (assert_torsion_free_point(0xdeadbeef * GENERATOR))
    "torsion": [29, 237, 27, 86, 26, 113, 3, 36, 200, 203, 232, 100, 142, 46, 26, 186, 229, 225, 226, 228, 94, 68, 79, 22, 245, 233, 57, 1, 14, 37, 206, 53],
    # src/composer/tests/soundness/point.rs:1078-1082 (component_mul_point(17, 0xdeadbeef * GENERATOR))
    "mulpoint": [250, 132, 56, 170, 228, 252, 166, 13, 108, 124, 132, 6, 89, 188, 88, 247, 231, 121, 77, 144, 115, 248, 63, 117, 196, 123, 96, 37, 146, 174, 156, 68],
    # src/composer/tests/soundness/point.rs:1606-1609 (component_add_point(GENERATOR, 2 * GENERATOR))
    "addpoint": [228, 59, 231, 43, 120, 95, 179, 34, 228, 43, 10, 248, 22, 142, 41, 174, 127, 155, 191, 155, 9, 56, 184, 82, 223, 173, 215, 132, 79, 23, 42, 4],
}

GEN = G.JUBJUB_GENERATOR


def prime_order_point():
    return G.jj_mul(GEN, 0xDEADBEEF)


def golden_circuits():
    """name -> function building the golden's circuit on any composer with the shared method names."""
    def rng(bp):
        return lambda c: c.component_range(c.append_witness(R_MOD - 1), bp)

    def mulpoint(c):
        s = c.append_witness(17)
        c.component_mul_point(s, c.append_point(prime_order_point()))

    def addpoint(c):
        a = c.append_point(GEN)
        b = c.append_point(G.jj_add(GEN, GEN))
        c.component_add_point(a, b)

    return {
        "range16": rng(16), "range32": rng(32), "range128": rng(128),
        "xor": lambda c: c.append_logic_xor(c.append_witness(R_MOD - 1), 0, 125),
        "and": lambda c: c.append_logic_and(c.append_witness(R_MOD - 1), 0, 125),
        "fixed": lambda c: c.component_mul_generator(c.append_witness(G.JUBJUB_ORDER - 2), GEN),
        "torsion": lambda c: c.assert_torsion_free_point(c.append_point(prime_order_point())),
        "mulpoint": mulpoint,
        "addpoint": addpoint,
    }


def oracle_arrays(c):
    sel = b"".join(R.fr_vec_to_mont_bytes([g.sel[k] for g in c.constraints]) for k in R.SELECTORS)
    wires = b"".join(int(getattr(g, col)).to_bytes(4, "little") for col in "abcd" for g in c.constraints)
    idx = c.public_input_indexes()
    return (len(c.constraints), sel, wires, R.fr_vec_to_mont_bytes(c.witnesses),
            b"".join(i.to_bytes(8, "little") for i in idx), R.fr_vec_to_mont_bytes(c.public_inputs_vec()))


def native_arrays(c):
    a = c.arrays()
    return (a.constraints, a.selectors, a.wires, a.witnesses, a.pi_idx, a.pi_vals)


def digest_of_arrays(arr) -> bytes:
    """gate_digest recomputed from the flat export (column-major selectors / wires)."""
    n, sel, wires = arr[0], R.fr_vec_from_mont_bytes(arr[1]), arr[2]
    acc = 0
    for i in range(n):
        for s in range(11):
            acc = (acc * 1_000_003 + sel[s * n + i]) % R_MOD
        for k in range(4):
            acc = (acc * 1_000_003 + int.from_bytes(wires[4 * (k * n + i) : 4 * (k * n + i) + 4], "little")) % R_MOD
    return acc.to_bytes(32, "little")


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_oracle_gadgets_match_reference_layout_goldens(name):
    c = G.GadgetComposer.initialized()
    golden_circuits()[name](c)
    assert G.gate_digest(c) == bytes(GOLDEN[name])
    if not name.startswith("range"):  # the range goldens check -1 against 32..256 bits: layout only
        assert G.unsatisfied_rows(c) == []


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_native_gadgets_match_reference_layout_goldens(name):
    c = N.Composer.initialized()
    golden_circuits()[name](c)
    assert digest_of_arrays(native_arrays(c)) == bytes(GOLDEN[name])


def test_jubjub_constants():
    assert N.jubjub_generator() == GEN
    assert G.jj_is_on_curve(GEN) and G.jj_is_prime_order(GEN)
    assert 8 * G.EIGHT_INV % G.JUBJUB_ORDER == 1  # point.rs:17-22
    for k in (1, 2, 7, 0xDEADBEEF, G.JUBJUB_ORDER - 1, G.JUBJUB_ORDER):
        assert N.jubjub_mul(GEN, k) == G.jj_mul(GEN, k)
    # signed digits recompose and are non-adjacent (JubJubScalar::compute_windowed_naf(2))
    rng = random.Random(5)
    for k in [0, 1, 3, G.JUBJUB_ORDER - 1] + [rng.randrange(G.JUBJUB_ORDER) for _ in range(20)]:
        d = G.compute_windowed_naf2(k)
        assert len(d) == 256 and sum(x << i for i, x in enumerate(d)) == k
        assert all(d[i] == 0 or d[i + 1] == 0 for i in range(255))


def build_mixed(c, rng: random.Random):
    """Every public gadget once, on seeded values; same calls on either composer."""
    rf = lambda: rng.randrange(R_MOD)
    a, b = c.append_witness(rf()), c.append_witness(rf())
    small = c.append_witness(rng.randrange(1 << 61))
    bit = c.append_witness(1)
    nobit = c.append_witness(0)
    c.component_boolean(bit)
    c.component_decomposition(small, 61)
    c.component_decomposition(a, 255)
    c.component_range_bits(small, 61)   # odd width
    c.component_range_bits(small, 64)
    c.component_range_bits(c.ZERO, 0)
    c.component_range(small, 32)
    c.append_logic_and(a, b, 127)
    c.append_logic_xor(a, b, 3)
    c.append_logic_xor(a, b, 0)
    c.component_truncate(a, 7)
    c.component_truncate(a, 254)
    c.component_truncate(c.append_witness(R_MOD - 1), 100)  # high part at its maximum: the canonical guard is active
    c.component_select(bit, a, b)
    c.component_select(nobit, a, b)
    c.component_select_one(nobit, a)
    c.component_select_zero(bit, a)
    c.append_constant(rf())
    c.append_public(rf())
    c.assert_equal_constant(c.append_witness(5), 2, public=3)
    p = G.jj_mul(GEN, rng.randrange(G.JUBJUB_ORDER))
    q = G.jj_mul(GEN, rng.randrange(G.JUBJUB_ORDER))
    wp, wq = c.append_point(p), c.append_public_point(q)
    c.append_constant_point(p)
    c.assert_equal_public_point(wq, q)
    c.assert_torsion_free_point(wp)
    s = c.component_add_point(wp, wq)
    d = c.component_sub_point(s, wq)
    c.assert_equal_point(d, wp)
    c.component_neg_point(wp)
    c.component_add_point(wp, c.component_neg_point(wp))  # lands on the identity
    c.component_select_identity(nobit, wp)
    c.component_select_point(bit, wp, wq)
    k = c.append_witness(rng.randrange(G.JUBJUB_ORDER))
    c.component_mul_generator(k, GEN)
    c.component_mul_generator(k, q)
    c.component_mul_point(k, wq)
    ev = dict(q_m=rf(), q_l=rf(), q_r=rf(), q_f=rf(), q_c=rf())
    c.gate_add(ev, a=a, b=b, d=small, public=rf())
    return c


@pytest.mark.parametrize("seed", [1, 2])
def test_native_composer_equals_oracle_on_every_gadget(seed):
    o = build_mixed(G.GadgetComposer.initialized(), random.Random(seed))
    n = build_mixed(N.Composer.initialized(), random.Random(seed))
    assert G.unsatisfied_rows(o) == []
    assert native_arrays(n) == oracle_arrays(o)


def test_append_evaluated_output_solves_for_any_output_selector():
    rng = random.Random(9)
    for q_o in (1, R_MOD - 1, 0, rng.randrange(R_MOD)):
        c = N.Composer.initialized()
        a, b = c.append_witness(rng.randrange(R_MOD)), c.append_witness(rng.randrange(R_MOD))
        sel = dict(q_m=3, q_l=5, q_r=7, q_c=11, q_o=q_o)
        before = c.constraints()
        w = c.append_evaluated_output(sel, a=a, b=b)
        assert c.constraints() == before + 1  # composer.rs:298-352: always exactly one gate
        if q_o == 0:
            assert w is None
        else:
            x = (3 * c[a] * c[b] + 5 * c[a] + 7 * c[b] + 11) % R_MOD
            assert (x + q_o * c[w]) % R_MOD == 0


@pytest.mark.parametrize("degree", [1 << 5, 1 << 13])
def test_bench_circuit_native_equals_oracle(degree):
    """benches/plonk.rs BenchCircuit<DEGREE>: the native front end's export is byte-identical to the
    oracle's, and every row satisfies its gate identity."""
    o = G.GadgetComposer.initialized()
    G.bench_circuit(o, degree)
    assert G.unsatisfied_rows(o, limit=1) == []
    n = N.bench_circuit(degree)
    assert native_arrays(n) == oracle_arrays(o)
    assert n.constraints() == 4 + 3375 * (1 if degree == 32 else 2)


def test_gadget_errors_mirror_the_reference():
    c = N.Composer.initialized()
    k = c.append_witness(G.JUBJUB_ORDER)  # not a canonical JubJub scalar
    with pytest.raises(Pb200Error) as e:
        c.component_mul_generator(k, GEN)
    assert e.value.code == N.PB200_ERR_JUBJUB_SCALAR  # Error::JubJubScalarMalformed, fixed_base.rs:57-61
    ok = c.append_witness(5)
    torsion2 = (0, R_MOD - 1)  # the order-2 point (point.rs tests: torsion_order_2)
    for bad in (torsion2, (0, 1), (1, 1)):  # small order, identity, off-curve
        with pytest.raises(Pb200Error) as e:
            c.component_mul_generator(ok, bad)
        assert e.value.code == N.PB200_ERR_JUBJUB_GENERATOR  # Error::JubJubGeneratorNotPrimeOrder, fixed_base.rs:50-55
    with pytest.raises(Pb200Error) as e:
        c.append_constant_point(torsion2)
    assert e.value.code == N.PB200_ERR_JUBJUB_POINT  # Error::JubJubPointNotTorsionFree, point.rs:77-80
    with pytest.raises(Pb200Error):
        c.assert_equal(0, 10 ** 6)  # unallocated witness
    with pytest.raises(Pb200Error):
        c.append_logic_and(0, 1, 128)  # BIT_PAIRS <= 127, logic.rs:50-55
    o = G.GadgetComposer.initialized()
    with pytest.raises(ValueError):
        o.component_mul_generator(o.append_witness(G.JUBJUB_ORDER), GEN)


def test_witness_only_composer_reproduces_the_witness_table():
    """The composer of Prover::prove (prover.rs:425) re-runs the circuit for its witnesses only: same witness
    table, same gate count and public inputs as the full composer, no gate layout kept."""
    import ctypes

    from plonk_b200._lib import check, lib

    L = lib()
    full, wo = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.pb200_composer_new(ctypes.byref(full)))
    check(L.pb200_composer_new(ctypes.byref(wo)))
    check(L.pb200_composer_set_witness_only(wo, 1))
    for h in (full, wo):
        check(L.pb200_composer_bench_circuit(h, 1 << 12))
        pub = (5).to_bytes(32, "little")
        w = ctypes.c_uint32()
        check(L.pb200_composer_append_public(h, pub, ctypes.byref(w)))
    n_w, n_g = L.pb200_composer_witnesses(full), L.pb200_composer_constraints(full)
    assert (L.pb200_composer_witnesses(wo), L.pb200_composer_constraints(wo), L.pb200_composer_public_inputs(wo)) == (n_w, n_g, 1)
    a, b = ctypes.create_string_buffer(32 * n_w), ctypes.create_string_buffer(32 * n_w)
    ia, ib = ctypes.create_string_buffer(8), ctypes.create_string_buffer(8)
    va, vb = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
    check(L.pb200_composer_export(full, None, None, a, ia, va))
    check(L.pb200_composer_export(wo, None, None, b, ib, vb))
    assert a.raw == b.raw and ia.raw == ib.raw and va.raw == vb.raw
    sel = ctypes.create_string_buffer(11 * 32 * n_g)
    assert L.pb200_composer_export(wo, sel, None, None, None, None) == -4  # no gate layout in this mode
    assert L.pb200_composer_set_witness_only(wo, 0) == -4
    L.pb200_composer_free(full)
    L.pb200_composer_free(wo)
