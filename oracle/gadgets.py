"""TEST INFRASTRUCTURE - CPU oracle, not product code.

Pure-Python restatement of the reference's circuit gadget library (the part of the turbo Composer
that `benches/plonk.rs::BenchCircuit` drives), on top of `pyref.Composer`:

    src/composer/bits.rs        component_boolean, component_decomposition
    src/composer/range.rs       component_range_bits / component_range, range_check(_even)
    src/composer/logic.rs       append_logic_and / append_logic_xor (+ truncation binding)
    src/composer/truncate.rs    bind_truncation_split, component_truncate, assert_canonical_truncation
    src/composer/select.rs      component_select, component_select_one, component_select_zero
    src/composer/point.rs       append_point, append_constant_point, assert_equal_point,
                                component_add_point, component_mul_point, component_select_identity,
                                component_select_point, assert_torsion_free_point, ...
    src/composer/fixed_base.rs  component_mul_generator

JubJub itself lives in the un-vendored dusk-jubjub 0.15; what is used here is restated from its
published definition (twisted Edwards -u^2 + v^2 = 1 + d u^2 v^2 over BLS12-381's Fr,
d = -(10240/10241), prime subgroup order below, generator v = 18) and PINNED by the reference's own
gate-layout goldens: `component_mul_generator_layout_matches_golden`
(src/composer/tests/soundness/fixed_base.rs:1013-1033) folds the affine coordinates of
2^i * GENERATOR into the digest, so a wrong generator, curve constant or doubling formula cannot
reproduce it (tests/test_gadgets.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file."""
from __future__ import annotations

from typing import List, Optional, Tuple

try:
    from . import pyref as P
except ImportError:  # imported as a top-level module (sys.path contains oracle/)
    import pyref as P

R_MOD = P.R_MOD
EDWARDS_D = P.EDWARDS_D
# dusk_jubjub scalar field modulus (order of the prime subgroup)
JUBJUB_ORDER = 0x0E7DB4EA6533AFA906673B0101343B00A6682093CCC81082D0970E5ED6F72CB7
# dusk_jubjub::GENERATOR (affine u, v)
JUBJUB_GENERATOR = (0x3FD2814C43AC65A6F1FBF02D0FD6CCE62E3EBB21FD6C54ED4DF7B7FFEC7BEACA, 18)
JUBJUB_IDENTITY = (0, 1)
# src/composer/point.rs:17-22 (JubJubScalar::from_raw => canonical limbs, little-endian)
EIGHT_INV = 0x01CFB69D4CA675F520CCE7602026876014CD0412799902105A12E1CBDADEE597

Point = Tuple[int, int]


# ---- JubJub (affine; the reference computes in extended coordinates and normalises - same values)
def jj_add(p: Point, q: Point) -> Point:
    """Complete twisted-Edwards addition.  A vanishing denominator (only possible off the prime
    subgroup) is the reference's `sum.get_z() == 0` case and maps to the identity (point.rs:226-231)."""
    x1, y1 = p
    x2, y2 = q
    t = EDWARDS_D * x1 % R_MOD * x2 % R_MOD * y1 % R_MOD * y2 % R_MOD
    dx, dy = (1 + t) % R_MOD, (1 - t) % R_MOD
    if dx == 0 or dy == 0:
        return JUBJUB_IDENTITY
    return ((x1 * y2 + y1 * x2) * P.fr_inv(dx) % R_MOD, (y1 * y2 + x1 * x2) * P.fr_inv(dy) % R_MOD)


def jj_neg(p: Point) -> Point:
    return ((-p[0]) % R_MOD, p[1])


def jj_mul(p: Point, k: int) -> Point:
    acc = JUBJUB_IDENTITY
    while k:
        if k & 1:
            acc = jj_add(acc, p)
        p = jj_add(p, p)
        k >>= 1
    return acc


def jj_is_on_curve(p: Point) -> bool:
    x2, y2 = p[0] * p[0] % R_MOD, p[1] * p[1] % R_MOD
    return (y2 - x2) % R_MOD == (1 + EDWARDS_D * x2 % R_MOD * y2) % R_MOD


def jj_is_torsion_free(p: Point) -> bool:
    return jj_mul(p, JUBJUB_ORDER) == JUBJUB_IDENTITY


def jj_is_prime_order(p: Point) -> bool:
    return jj_is_torsion_free(p) and p != JUBJUB_IDENTITY


def compute_windowed_naf2(k: int) -> List[int]:
    """JubJubScalar::compute_windowed_naf(2): 256 digits in {-1, 0, 1}, least significant first,
    no two adjacent non-zero (dusk-jubjub; call site fixed_base.rs:62-63)."""
    res = [0] * 256
    i = 0
    while k >= 1:
        if k & 1:
            ki = k & 3
            if ki >= 2:
                ki -= 4
            res[i] = ki
            k -= ki
        k >>= 1
        i += 1
    return res


def to_bits(v: int) -> List[int]:
    """BlsScalar::to_bits: 256 bits, least significant first."""
    return [(v >> i) & 1 for i in range(256)]


def recompose_bits(bits: List[int], start: int, end: int) -> int:  # bits.rs:11-26
    v = 0
    for i in range(end - 1, start - 1, -1):
        v = (2 * v + bits[i]) % R_MOD
    return v


class GadgetComposer(P.Composer):
    """pyref.Composer + the gadget library.  Witnesses are plain indices; a point is an (x, y) pair
    of witness indices."""

    IDENTITY = (0, 1)  # (Composer::ZERO, Composer::ONE), composer.rs:76-80

    def constraints_len(self) -> int:
        return len(self.constraints)

    def val(self, w: int) -> int:
        return self.witnesses[w]

    def gate_add(self, sel=None, public: Optional[int] = None, **kw) -> int:  # composer.rs:402-409
        """Selectors either as a dict or as keywords (q_l=..., q_r=...); wires a, b, d as keywords."""
        wires = {k: kw.pop(k) for k in ("a", "b", "d") if k in kw}
        return self.gate_evaluated(dict(sel or {}, **kw), public=public, **wires)

    gate_mul = gate_add  # composer.rs:411-417: same evaluation, q_o = -1

    def append_constant(self, constant: int) -> int:  # composer.rs:342-352
        w = self.append_witness(constant)
        self.assert_equal_constant(w, constant)
        return w

    # ---- bits.rs
    def component_boolean(self, a: int):
        self.append_gate(dict(q_m=1, q_o=-1), a=a, b=a, c=a, d=0)  # bits.rs:36-47

    def component_decomposition(self, scalar: int, n: int) -> List[int]:  # bits.rs:60-98
        assert 0 < n <= 256
        bits = to_bits(self.val(scalar))
        acc = self.ZERO
        out = []
        for i in range(n):
            w_bit = self.append_witness(bits[i])
            self.component_boolean(w_bit)
            acc = self.gate_add(q_l=pow(2, i, R_MOD), q_r=1, a=w_bit, b=acc)
            out.append(w_bit)
        self.assert_equal(acc, scalar)
        return out

    # ---- range.rs
    def component_range_bits(self, witness: int, bits: int):
        assert bits <= 256
        self.range_check(witness, bits)

    def component_range(self, witness: int, bit_pairs: int):  # deprecated entry point, range.rs:52-57
        self.range_check_even(witness, min(bit_pairs * 2, 256))

    def range_check(self, value: int, num_bits: int):  # range.rs:59-85
        if num_bits % 2 == 0:
            self.range_check_even(value, num_bits)
            return
        top = num_bits - 1
        vbits = to_bits(self.val(value))
        lower = self.append_witness(recompose_bits(vbits, 0, top))
        self.range_check_even(lower, top)
        top_bit = self.append_witness(vbits[top])
        self.component_boolean(top_bit)
        recomposed = self.gate_add(q_l=1, q_r=pow(2, top, R_MOD), a=lower, b=top_bit)
        self.assert_equal(recomposed, value)

    def range_check_even(self, witness: int, num_bits: int):  # range.rs:87-169
        if num_bits == 0:
            self.append_gate(dict(q_l=1), a=witness)
            return
        bits = to_bits(self.val(witness))
        num_gates = (num_bits >> 3) + (1 if num_bits % 8 else 0)
        num_quads = num_gates * 4
        pad = 1 + (((num_quads << 1) - num_bits) >> 1)
        used_gates = num_gates + 1
        rows = [dict(sel=dict(q_range=1), w=[0, 0, 0, 0]) for _ in range(used_gates)]
        accumulators = []
        acc = 0
        for i in range(pad, num_quads + 1):
            bit_index = (num_quads - i) << 1
            quad = bits[bit_index] + 2 * bits[bit_index + 1]
            acc = (4 * acc + quad) % R_MOD
            w = self.append_witness(acc)
            accumulators.append(w)
            rows[i // 4]["w"][(3, 2, 1, 0)[i % 4]] = w  # D, C, B, A
        rows[-1] = dict(sel={}, w=[0, 0, 0, 0])
        if accumulators:
            rows[-1]["w"][3] = accumulators[-1]
        for r in rows:
            a, b, c, d = r["w"]
            self.append_custom_gate(r["sel"], a=a, b=b, c=c, d=d)
        if accumulators:
            self.assert_equal(accumulators[-1], witness)

    # ---- logic.rs
    def append_logic_component(self, a: int, b: int, bit_pairs: int, is_xor: bool) -> int:  # logic.rs:44-153
        assert bit_pairs <= 127
        num_bits = bit_pairs * 2
        a_bits = [(self.val(a) >> i) & 1 for i in range(num_bits - 1, -1, -1)]  # most significant first
        b_bits = [(self.val(b) >> i) & 1 for i in range(num_bits - 1, -1, -1)]
        sel = dict(q_c=-1, q_logic=-1) if is_xor else dict(q_c=1, q_logic=1)
        wa = wb = wc = wd = 0
        left_acc = right_acc = out_acc = 0
        for i in range(bit_pairs):
            lq = (a_bits[2 * i] << 1) + a_bits[2 * i + 1]
            rq = (b_bits[2 * i] << 1) + b_bits[2 * i + 1]
            oq = (lq ^ rq) if is_xor else (lq & rq)
            left_acc = (left_acc * 4 + lq) % R_MOD
            right_acc = (right_acc * 4 + rq) % R_MOD
            out_acc = (out_acc * 4 + oq) % R_MOD
            wit_a = self.append_witness(left_acc)
            wit_b = self.append_witness(right_acc)
            wit_c = self.append_witness(lq * rq)
            wit_d = self.append_witness(out_acc)
            wc = wit_c
            self.append_custom_gate(sel, a=wa, b=wb, c=wc, d=wd)
            wa, wb, wd = wit_a, wit_b, wit_d
        self.append_custom_gate({}, a=wa, b=wb, d=wd)
        if bit_pairs:  # bind_logic_accumulators, logic.rs:155-170
            self.bind_truncation_split(a, wa, num_bits)
            self.bind_truncation_split(b, wb, num_bits)
        return wd

    def append_logic_and(self, a: int, b: int, bit_pairs: int) -> int:
        return self.append_logic_component(a, b, bit_pairs, False)

    def append_logic_xor(self, a: int, b: int, bit_pairs: int) -> int:
        return self.append_logic_component(a, b, bit_pairs, True)

    # ---- truncate.rs
    def bind_truncation_split(self, inp: int, low: int, num_bits: int):  # truncate.rs:24-44
        high_bits = 255 - num_bits
        high = self.append_witness(recompose_bits(to_bits(self.val(inp)), num_bits, 256))
        self.range_check(high, high_bits)
        recomposed = self.gate_add(q_l=pow(2, num_bits, R_MOD), q_r=1, a=high, b=low)
        self.assert_equal(recomposed, inp)
        self.assert_canonical_truncation(high, low, num_bits)

    def component_truncate(self, witness: int, n: int) -> int:  # truncate.rs:46-63
        assert n <= 254
        low = self.append_witness(recompose_bits(to_bits(self.val(witness)), 0, n))
        self.range_check(low, n)
        self.bind_truncation_split(witness, low, n)
        return low

    def assert_canonical_truncation(self, high: int, low: int, num_bits: int):  # truncate.rs:65-107
        high_bits = 255 - num_bits
        mbits = to_bits(R_MOD - 1)
        r_low = recompose_bits(mbits, 0, num_bits)
        r_high = recompose_bits(mbits, num_bits, 256)
        diff = self.gate_add(q_l=-1, q_c=r_high, a=high)
        self.range_check(diff, high_bits)
        dv = self.val(diff)
        inverse = self.append_witness(P.fr_inv(dv) if dv else 0)
        product = self.gate_mul(q_m=1, a=diff, b=inverse)
        is_top = self.gate_add(q_l=-1, q_c=1, a=product)
        self.append_gate(dict(q_m=1), a=diff, b=is_top)
        r_low_minus_low = self.gate_add(q_l=-1, q_c=r_low, a=low)
        guard = self.gate_mul(q_m=1, a=is_top, b=r_low_minus_low)
        self.range_check(guard, num_bits)

    # ---- select.rs
    def component_select(self, bit: int, a: int, b: int) -> int:  # select.rs:20-44
        bit_times_a = self.gate_mul(q_m=1, a=bit, b=a)
        one_min_bit = self.gate_add(q_l=-1, q_c=1, a=bit)
        one_min_bit_b = self.gate_mul(q_m=1, a=one_min_bit, b=b)
        return self.gate_add(q_l=1, q_r=1, a=one_min_bit_b, b=bit_times_a)

    def component_select_one(self, bit: int, value: int) -> int:  # select.rs:52-76
        b, v = self.val(bit), self.val(value)
        f_x = self.append_witness((1 - b + b * v) % R_MOD)
        self.append_gate(dict(q_m=1, q_l=-1, q_o=-1, q_c=1), a=bit, b=value, c=f_x)
        return f_x

    def component_select_zero(self, bit: int, value: int) -> int:  # select.rs:84-92
        return self.gate_mul(q_m=1, a=bit, b=value)

    # ---- point.rs
    def append_point(self, point: Point) -> Tuple[int, int]:  # point.rs:40-59
        return (self.append_witness(point[0]), self.append_witness(point[1]))

    def append_constant_point(self, point: Point) -> Tuple[int, int]:  # point.rs:70-89
        if not (jj_is_on_curve(point) and jj_is_torsion_free(point)):
            raise ValueError("JubJubPointNotTorsionFree")
        return (self.append_constant(point[0]), self.append_constant(point[1]))

    def append_public_point(self, point: Point) -> Tuple[int, int]:  # point.rs:97-121
        w = self.append_point(point)
        self.assert_equal_constant(w[0], 0, public=point[0])
        self.assert_equal_constant(w[1], 0, public=point[1])
        return w

    def assert_equal_point(self, a, b):  # point.rs:124-127
        self.assert_equal(a[0], b[0])
        self.assert_equal(a[1], b[1])

    def assert_equal_public_point(self, point, public: Point):  # point.rs:134-158
        self.assert_equal_constant(point[0], 0, public=public[0])
        self.assert_equal_constant(point[1], 0, public=public[1])

    def assert_torsion_free_point(self, point):  # point.rs:171-190
        pv = (self.val(point[0]), self.val(point[1]))
        q = jj_mul(pv, EIGHT_INV) if jj_is_on_curve(pv) else JUBJUB_IDENTITY
        self.assert_torsion_free_gates(point, q)
        return point

    def assert_torsion_free_gates(self, point, q: Point):  # point.rs:192-221
        qw = self.append_point(q)
        qu, qv = qw
        u2 = self.gate_mul(q_m=1, a=qu, b=qu)
        v2 = self.gate_mul(q_m=1, a=qv, b=qv)
        u2v2 = self.gate_mul(q_m=1, a=u2, b=v2)
        self.append_gate(dict(q_l=-1, q_r=1, q_o=-EDWARDS_D, q_c=-1), a=u2, b=v2, c=u2v2)
        q2 = self.add_point_gates(qw, qw)
        q4 = self.add_point_gates(q2, q2)
        q8 = self.add_point_gates(q4, q4)
        self.assert_equal_point(point, q8)

    def component_neg_point(self, p):  # point.rs:224-235
        return (self.gate_mul(q_l=-1, a=p[0]), p[1])

    def component_sub_point(self, a, b):  # point.rs:238-246
        return self.component_add_point(a, self.component_neg_point(b))

    def component_add_point(self, a, b):  # point.rs:256-264
        return self.add_point_gates(a, b)

    def add_point_gates(self, a, b):  # point.rs:266-312
        x_1, y_1 = a
        x_2, y_2 = b
        p1 = (self.val(x_1), self.val(y_1))
        p2 = (self.val(x_2), self.val(y_2))
        s = jj_add(p1, p2)
        x_1_y_2 = self.append_witness(p1[0] * p2[1])
        x_3 = self.append_witness(s[0])
        y_3 = self.append_witness(s[1])
        self.append_custom_gate(dict(q_variable_group_add=1), a=x_1, b=y_1, c=x_2, d=y_2)
        self.append_custom_gate({}, a=x_3, b=y_3, d=x_1_y_2)
        return (x_3, y_3)

    def component_select_identity(self, bit: int, a):  # point.rs:322-332
        self.component_boolean(bit)
        return self.select_identity_gates(bit, a)

    def select_identity_gates(self, bit: int, a):  # point.rs:334-343
        x = self.component_select_zero(bit, a[0])
        y = self.component_select_one(bit, a[1])
        return (x, y)

    def component_mul_point(self, jubjub: int, point):  # point.rs:361-378
        scalar_bits = self.component_decomposition(jubjub, 252)
        result = self.IDENTITY
        for bit in reversed(scalar_bits):
            result = self.add_point_gates(result, result)
            point_to_add = self.select_identity_gates(bit, point)
            result = self.add_point_gates(result, point_to_add)
        return result

    def component_select_point(self, bit: int, a, b):  # point.rs:387-397
        x = self.component_select(bit, a[0], b[0])
        y = self.component_select(bit, a[1], b[1])
        return (x, y)

    # ---- fixed_base.rs
    JUBJUB_SCALAR_BITS = 252
    FIXED_BASE_ROUNDS = 256
    FIXED_BASE_LEADING_ZERO_ROUNDS = 256 - (252 + 1)

    def component_mul_generator(self, jubjub: int, generator: Point):  # fixed_base.rs:47-77
        if not jj_is_on_curve(generator) or not jj_is_prime_order(generator):
            raise ValueError("JubJubGeneratorNotPrimeOrder")
        scalar = self.val(jubjub)
        if scalar >= JUBJUB_ORDER:
            raise ValueError("JubJubScalarMalformed")
        return self.append_fixed_base_signed_digits(jubjub, generator, compute_windowed_naf2(scalar))

    def append_fixed_base_signed_digits(self, jubjub: int, generator: Point, digits: List[int]):  # fixed_base.rs:79-226
        n = self.FIXED_BASE_ROUNDS
        self.assert_canonical_jubjub_scalar(jubjub)
        multiples = [generator]
        for _ in range(1, n):
            multiples.append(jj_add(multiples[-1], multiples[-1]))
        multiples.reverse()
        scalar_acc = [0]
        point_acc = [JUBJUB_IDENTITY]
        xy_alphas = []
        for i, entry in enumerate(reversed(digits)):
            if entry == 0:
                s_add, p_add = 0, JUBJUB_IDENTITY
            elif entry == -1:
                s_add, p_add = -1, jj_neg(multiples[i])
            elif entry == 1:
                s_add, p_add = 1, multiples[i]
            else:
                raise ValueError("UnsupportedWNAF2k")
            scalar_acc.append((2 * scalar_acc[i] + s_add) % R_MOD)
            point_acc.append(jj_add(point_acc[i], p_add))
            xy_alphas.append(p_add[0] * p_add[1] % R_MOD)
        leading = self.ZERO
        for i in range(n):
            acc_x = self.append_witness(point_acc[i][0])
            acc_y = self.append_witness(point_acc[i][1])
            accumulated_bit = self.append_witness(scalar_acc[i])
            if i == self.FIXED_BASE_LEADING_ZERO_ROUNDS:
                leading = accumulated_bit
            if i == 0:
                self.assert_equal_constant(acc_x, 0)
                self.assert_equal_constant(acc_y, 1)
                self.assert_equal_constant(accumulated_bit, 0)
            x_beta, y_beta = multiples[i]
            xy_alpha = self.append_witness(xy_alphas[i])
            self.append_custom_gate(dict(q_fixed_group_add=1, q_l=x_beta, q_r=y_beta, q_c=x_beta * y_beta),
                                    a=acc_x, b=acc_y, c=xy_alpha, d=accumulated_bit)
        acc_x = self.append_witness(point_acc[n][0])
        acc_y = self.append_witness(point_acc[n][1])
        last = self.append_witness(scalar_acc[n])
        self.append_gate({}, a=acc_x, b=acc_y, d=last)
        self.assert_equal_constant(leading, 0)
        self.assert_equal(last, jubjub)
        return (acc_x, acc_y)

    def assert_canonical_jubjub_scalar(self, scalar: int):  # fixed_base.rs:228-240
        self.range_check(scalar, self.JUBJUB_SCALAR_BITS)
        distance = self.gate_add(q_l=-1, q_c=JUBJUB_ORDER - 1, a=scalar)
        self.range_check(distance, self.JUBJUB_SCALAR_BITS)


def gate_digest(comp: P.Composer) -> bytes:
    """src/composer/tests/soundness/support.rs:93-135: fold of every selector and wire index."""
    mult, acc = 1_000_003, 0
    for g in comp.constraints:
        for k in P.SELECTORS:
            acc = (acc * mult + g.sel[k]) % R_MOD
        for w in (g.a, g.b, g.c, g.d):
            acc = (acc * mult + w) % R_MOD
    return acc.to_bytes(32, "little")


def bench_circuit(comp: GadgetComposer, degree: int) -> None:
    """benches/plonk.rs:12-82 BenchCircuit<DEGREE>::circuit with the Default values (a=2, b=3, x=6,
    y=7, z = 7 * GENERATOR)."""
    z = jj_mul(JUBJUB_GENERATOR, 7)
    w_a = comp.append_witness(2)
    w_b = comp.append_witness(3)
    w_x = comp.append_witness(6)
    w_y = comp.append_witness(7)
    w_z = comp.append_point(z)
    diff = 0
    prev = comp.constraints_len()
    while prev + diff < degree:
        r_w = comp.gate_mul(q_m=1, a=w_a, b=w_b)
        comp.append_constant(15)
        comp.append_constant_point(z)
        comp.assert_equal(w_x, r_w)
        comp.assert_equal_point(w_z, w_z)
        comp.gate_add(q_l=1, q_r=1, a=w_a, b=w_b)
        comp.component_add_point(w_z, w_z)
        comp.append_logic_and(w_a, w_b, 127)
        comp.append_logic_xor(w_a, w_b, 127)
        comp.component_boolean(comp.ONE)
        comp.component_decomposition(w_a, 254)
        comp.component_mul_generator(w_y, JUBJUB_GENERATOR)
        comp.component_mul_point(w_y, w_z)
        comp.component_range_bits(w_a, 256)
        comp.component_select(comp.ONE, w_a, w_b)
        comp.component_select_identity(comp.ONE, w_z)
        comp.component_select_one(comp.ONE, w_a)
        comp.component_select_point(comp.ONE, w_z, w_z)
        comp.component_select_zero(comp.ONE, w_a)
        diff = comp.constraints_len() - prev
        prev = comp.constraints_len()


def unsatisfied_rows(comp: P.Composer, n: Optional[int] = None, limit: int = 8) -> List[Tuple[int, str]]:
    """Row-by-row check of every gate identity on the witness table (the widgets' quotient terms
    evaluated on the domain instead of the coset; `*_w` are the next row's wires, the table being
    zero-padded to the domain size n and cyclic).  Returns up to `limit` (row, family) failures."""
    rows = comp.constraints
    m = len(rows)
    if n is None:
        n = 1
        while n < m:
            n <<= 1
    W = comp.witnesses
    ch = 0x1234567  # any non-trivial separation challenge
    zero_row = (0, 0, 0, 0)

    def wires(i: int):
        i %= n
        if i >= m:
            return zero_row
        g = rows[i]
        return (W[g.a], W[g.b], W[g.c], W[g.d])

    bad = []
    for i, g in enumerate(rows):
        a, b, c, d = wires(i)
        a_w, b_w, _, d_w = wires(i + 1)
        q = g.sel
        checks = (
            ("arith", (P.widget_arith(q, a, b, c, d) + comp.public_inputs.get(i, 0)) % R_MOD if q["q_arith"] else 0),
            ("range", P.widget_range_scalar(ch, a, b, c, d, d_w) if q["q_range"] else 0),
            ("logic", P.widget_logic_scalar(ch, q["q_c"], a, a_w, b, b_w, c, d, d_w) if q["q_logic"] else 0),
            ("fixed", P.widget_fixed_base_scalar(ch, q["q_l"], q["q_r"], q["q_c"], a, a_w, b, b_w, c, d, d_w) if q["q_fixed_group_add"] else 0),
            ("var", P.widget_curve_add_scalar(ch, a, a_w, b, b_w, c, d, d_w) if q["q_variable_group_add"] else 0),
        )
        for name, v in checks:
            if v % R_MOD:
                bad.append((i, name))
                if len(bad) >= limit:
                    return bad
    return bad
