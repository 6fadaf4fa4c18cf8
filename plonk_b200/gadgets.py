"""Gadget-level circuit front end: the reference's turbo Composer with its gadget library
(src/composer.rs, src/composer/{bits,range,logic,truncate,select,point,fixed_base}.rs), as built
into libplonk_b200 (csrc/composer.cpp, C ABI in include/plonk_b200_composer.h).

Host-side: circuit construction is CPU work in the reference too.  Values are canonical ints mod r
here and Montgomery limbs across the ABI; a witness is its index, a witness point an (x, y) pair of
indices, a JubJub point an (u, v) pair of ints."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

from . import _lib
from .composer import R_MOD, SELECTORS, CircuitArrays

_MONT_R = (1 << 256) % R_MOD
_MONT_R_INV = pow(_MONT_R, R_MOD - 2, R_MOD)

PB200_ERR_JUBJUB_POINT = -7
PB200_ERR_JUBJUB_GENERATOR = -8
PB200_ERR_JUBJUB_SCALAR = -9

Point = Tuple[int, int]
_U64x4 = C.c_uint64 * 4
_U64x8 = C.c_uint64 * 8
_U32x2 = C.c_uint32 * 2
_U32x4 = C.c_uint32 * 4


def _fr(v: int):
    m = (v % R_MOD) * _MONT_R % R_MOD
    return _U64x4(*[(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)])


def _fr_out(limbs) -> int:
    return sum(int(limbs[i]) << (64 * i) for i in range(4)) * _MONT_R_INV % R_MOD


def _pt(p: Point):
    a, b = _fr(p[0]), _fr(p[1])
    return _U64x8(*(list(a) + list(b)))


def _xy(p) -> "_U32x2":
    return _U32x2(int(p[0]), int(p[1]))


def _selectors(sel: dict):
    arr = (C.c_uint64 * (4 * len(SELECTORS)))()
    for k, v in sel.items():
        i = SELECTORS.index(k)
        arr[4 * i : 4 * i + 4] = list(_fr(v))
    return arr


def jubjub_generator() -> Point:
    """dusk_jubjub::GENERATOR."""
    out = _U64x8()
    _lib.check(_lib.lib().pb200_jubjub_generator(out))
    return (_fr_out(out[0:4]), _fr_out(out[4:8]))


def jubjub_mul(p: Point, k: int) -> Point:
    out = _U64x8()
    ks = _U64x4(*[(k >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)])
    _lib.check(_lib.lib().pb200_jubjub_mul(_pt(p), ks, out))
    return (_fr_out(out[0:4]), _fr_out(out[4:8]))


class Composer:
    """Composer::initialized() plus gadgets.  Method names and argument order follow the reference;
    const-generic widths (`::<N>`) are trailing arguments."""

    ZERO = 0
    ONE = 1
    IDENTITY = (0, 1)

    def __init__(self):
        self._L = _lib.lib()
        h = C.c_void_p()
        _lib.check(self._L.pb200_composer_new(C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.pb200_composer_free(h)

    @classmethod
    def initialized(cls) -> "Composer":
        return cls()

    # ---- inspection
    def constraints(self) -> int:
        return self._L.pb200_composer_constraints(self._h)

    def n_witnesses(self) -> int:
        return self._L.pb200_composer_witnesses(self._h)

    def __getitem__(self, w: int) -> int:
        out = _U64x4()
        _lib.check(self._L.pb200_composer_witness_value(self._h, w, out))
        return _fr_out(out)

    # ---- core
    def _w(self, fn, *args) -> int:
        out = C.c_uint32()
        _lib.check(fn(self._h, *args, C.byref(out)))
        return out.value

    def _p(self, fn, *args) -> Tuple[int, int]:
        out = _U32x2()
        _lib.check(fn(self._h, *args, out))
        return (out[0], out[1])

    def append_witness(self, v: int) -> int:
        return self._w(self._L.pb200_composer_append_witness, _fr(v))

    def append_gate(self, sel: dict, a=0, b=0, c=0, d=0, public: Optional[int] = None):
        _lib.check(self._L.pb200_composer_append_gate(self._h, _selectors(sel), _U32x4(a, b, c, d), None if public is None else _fr(public), 0))

    def append_custom_gate(self, sel: dict, a=0, b=0, c=0, d=0, public: Optional[int] = None):
        _lib.check(self._L.pb200_composer_append_gate(self._h, _selectors(sel), _U32x4(a, b, c, d), None if public is None else _fr(public), 1))

    def append_evaluated_output(self, sel: dict, a=0, b=0, d=0, public: Optional[int] = None) -> Optional[int]:
        out, solved = C.c_uint32(), C.c_int()
        _lib.check(self._L.pb200_composer_append_evaluated_output(self._h, _selectors(sel), _U32x4(a, b, 0, d),
                                                                  None if public is None else _fr(public), C.byref(out), C.byref(solved)))
        return out.value if solved.value else None

    def gate_add(self, sel: dict, a=0, b=0, d=0, public: Optional[int] = None) -> int:
        return self._w(self._L.pb200_composer_gate_add, _selectors(sel), _U32x4(a, b, 0, d), None if public is None else _fr(public))

    gate_mul = gate_add

    def append_constant(self, v: int) -> int:
        return self._w(self._L.pb200_composer_append_constant, _fr(v))

    def append_public(self, v: int) -> int:
        return self._w(self._L.pb200_composer_append_public, _fr(v))

    def assert_equal(self, a: int, b: int):
        _lib.check(self._L.pb200_composer_assert_equal(self._h, a, b))

    def assert_equal_constant(self, a: int, constant: int, public: Optional[int] = None):
        _lib.check(self._L.pb200_composer_assert_equal_constant(self._h, a, _fr(constant), None if public is None else _fr(public)))

    # ---- bits / range / logic / truncate / select
    def component_boolean(self, a: int):
        _lib.check(self._L.pb200_composer_component_boolean(self._h, a))

    def component_decomposition(self, scalar: int, n: int) -> List[int]:
        out = (C.c_uint32 * max(n, 1))()
        _lib.check(self._L.pb200_composer_component_decomposition(self._h, scalar, n, out))
        return list(out[:n])

    def component_range_bits(self, w: int, bits: int):
        _lib.check(self._L.pb200_composer_component_range_bits(self._h, w, bits))

    def component_range(self, w: int, bit_pairs: int):
        _lib.check(self._L.pb200_composer_component_range(self._h, w, bit_pairs))

    def append_logic_and(self, a: int, b: int, bit_pairs: int) -> int:
        return self._w(self._L.pb200_composer_append_logic, a, b, bit_pairs, 0)

    def append_logic_xor(self, a: int, b: int, bit_pairs: int) -> int:
        return self._w(self._L.pb200_composer_append_logic, a, b, bit_pairs, 1)

    def component_truncate(self, w: int, n: int) -> int:
        return self._w(self._L.pb200_composer_component_truncate, w, n)

    def component_select(self, bit: int, a: int, b: int) -> int:
        return self._w(self._L.pb200_composer_component_select, bit, a, b)

    def component_select_one(self, bit: int, value: int) -> int:
        return self._w(self._L.pb200_composer_component_select_one, bit, value)

    def component_select_zero(self, bit: int, value: int) -> int:
        return self._w(self._L.pb200_composer_component_select_zero, bit, value)

    # ---- points
    def append_point(self, p: Point):
        return self._p(self._L.pb200_composer_append_point, _pt(p), 0)

    def append_constant_point(self, p: Point):
        return self._p(self._L.pb200_composer_append_point, _pt(p), 1)

    def append_public_point(self, p: Point):
        return self._p(self._L.pb200_composer_append_point, _pt(p), 2)

    def assert_equal_point(self, a, b):
        _lib.check(self._L.pb200_composer_assert_equal_point(self._h, _xy(a), _xy(b)))

    def assert_equal_public_point(self, p, public: Point):
        _lib.check(self._L.pb200_composer_assert_equal_public_point(self._h, _xy(p), _pt(public)))

    def assert_torsion_free_point(self, p):
        _lib.check(self._L.pb200_composer_assert_torsion_free_point(self._h, _xy(p)))
        return p

    def component_add_point(self, a, b):
        return self._p(self._L.pb200_composer_point_op, 0, _xy(a), _xy(b))

    def component_sub_point(self, a, b):
        return self._p(self._L.pb200_composer_point_op, 1, _xy(a), _xy(b))

    def component_neg_point(self, a):
        return self._p(self._L.pb200_composer_point_op, 2, _xy(a), None)

    def component_select_identity(self, bit: int, a):
        return self._p(self._L.pb200_composer_component_select_identity, bit, _xy(a))

    def component_select_point(self, bit: int, a, b):
        return self._p(self._L.pb200_composer_component_select_point, bit, _xy(a), _xy(b))

    def component_mul_point(self, jubjub: int, p):
        return self._p(self._L.pb200_composer_component_mul_point, jubjub, _xy(p))

    def component_mul_generator(self, jubjub: int, generator: Optional[Point] = None):
        return self._p(self._L.pb200_composer_component_mul_generator, jubjub, None if generator is None else _pt(generator))

    # ---- whole circuits / export
    def bench_circuit(self, degree: int):
        """BenchCircuit<DEGREE>::circuit with its Default values (benches/plonk.rs:12-82)."""
        _lib.check(self._L.pb200_composer_bench_circuit(self._h, degree))

    def arrays(self) -> CircuitArrays:
        n, nw, npi = self.constraints(), self.n_witnesses(), self._L.pb200_composer_public_inputs(self._h)
        sel = C.create_string_buffer(len(SELECTORS) * n * 32)
        wires = C.create_string_buffer(4 * n * 4)
        wit = C.create_string_buffer(nw * 32)
        idx = C.create_string_buffer(max(npi, 1) * 8)
        vals = C.create_string_buffer(max(npi, 1) * 32)
        _lib.check(self._L.pb200_composer_export(self._h, sel, wires, wit, idx, vals))
        return CircuitArrays(n, sel.raw, wires.raw, wit.raw, idx.raw[: npi * 8], vals.raw[: npi * 32])


def bench_circuit(degree: int) -> Composer:
    """The reference's benchmark circuit (benches/plonk.rs) for `run::<DEGREE>`."""
    comp = Composer.initialized()
    comp.bench_circuit(degree)
    return comp
