"""Circuit front-end mirror: the subset of the reference's turbo Composer (src/composer.rs) needed
to describe arithmetic circuits for the GPU prover, producing the flat arrays the C ABI takes.

Host-side Python (the reference's composer is CPU, O(gates) and out of the GPU hot path - SURVEY.md
section 2 row 13).  Values are canonical ints mod r; `arrays()` converts to the reference's
in-memory Montgomery layout."""
from __future__ import annotations

from typing import Dict, List, Optional

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_MONT_R = (1 << 256) % R_MOD
SELECTORS = ["q_m", "q_l", "q_r", "q_o", "q_f", "q_c", "q_arith", "q_range", "q_logic", "q_fixed_group_add", "q_variable_group_add"]


def _mont_bytes(vals) -> bytes:
    return b"".join(((v * _MONT_R) % R_MOD).to_bytes(32, "little") for v in vals)


class CircuitArrays:
    """n_constraints, selectors (11 columns), wires (4 columns of u32), witnesses, public inputs."""

    def __init__(self, constraints, selectors, wires, witnesses, pi_idx, pi_vals):
        self.constraints = constraints
        self.selectors = selectors
        self.wires = wires
        self.witnesses = witnesses
        self.n_witnesses = len(witnesses) // 32
        self.pi_idx = pi_idx
        self.pi_vals = pi_vals
        self.n_pi = len(pi_idx) // 8


class Composer:
    ZERO = 0  # Composer::ZERO / ONE (composer.rs:84-95)
    ONE = 1

    def __init__(self):
        self.gates: List[tuple] = []  # (selector dict, a, b, c, d)
        self.public_inputs: Dict[int, int] = {}
        self.witnesses: List[int] = []

    @classmethod
    def initialized(cls) -> "Composer":
        """Composer::initialized: constants 0 and 1 plus the two dummy gates (composer.rs:177-240)."""
        s = cls()
        zero, one = s.append_witness(0), s.append_witness(1)
        s.assert_equal_constant(zero, 0)
        s.assert_equal_constant(one, 1)
        six, one_, seven, m20 = s.append_witness(6), s.append_witness(1), s.append_witness(7), s.append_witness(-20)
        s.append_gate(dict(q_m=1, q_l=2, q_r=3, q_f=1, q_c=4, q_o=4), a=six, b=seven, d=one_, c=m20)
        s.append_gate(dict(q_m=1, q_l=1, q_r=1, q_c=127, q_o=1), a=m20, b=six, c=seven)
        return s

    def constraints(self) -> int:
        return len(self.gates)

    def append_witness(self, v: int) -> int:
        self.witnesses.append(v % R_MOD)
        return len(self.witnesses) - 1

    def append_custom_gate(self, sel, a=0, b=0, c=0, d=0, public: Optional[int] = None):
        if public is not None:
            self.public_inputs[len(self.gates)] = public % R_MOD
        self.gates.append(({k: v % R_MOD for k, v in sel.items()}, a, b, c, d))

    def append_gate(self, sel, a=0, b=0, c=0, d=0, public: Optional[int] = None):
        s = dict(sel)
        s["q_arith"] = 1  # Constraint::arithmetic (constraint.rs:203-205)
        self.append_custom_gate(s, a, b, c, d, public)

    def assert_equal_constant(self, a: int, constant: int, public: Optional[int] = None):
        self.append_gate(dict(q_l=-1, q_c=constant), a=a, public=public)

    def assert_equal(self, a: int, b: int):
        self.append_gate(dict(q_l=1, q_r=-1), a=a, b=b)

    def append_constant(self, constant: int) -> int:
        w = self.append_witness(constant)
        self.assert_equal_constant(w, constant)
        return w

    def append_public(self, public: int) -> int:
        w = self.append_witness(public)
        self.append_gate(dict(q_l=-1), a=w, public=public)
        return w

    def _evaluated(self, sel, a, b, d, public):
        W = self.witnesses
        g = lambda k: sel.get(k, 0) % R_MOD
        x = (g("q_m") * W[a] * W[b] + g("q_l") * W[a] + g("q_r") * W[b] + g("q_f") * W[d] + g("q_c") + (public or 0)) % R_MOD
        c = self.append_witness(x)
        s = dict(sel)
        s["q_o"] = -1
        self.append_gate(s, a=a, b=b, c=c, d=d, public=public)
        return c

    def gate_add(self, sel, a=0, b=0, d=0, public: Optional[int] = None) -> int:
        """gate_add (composer.rs:402-409): c := q_l a + q_r b + q_f d + q_c + PI."""
        return self._evaluated(sel, a, b, d, public)

    def gate_mul(self, sel, a=0, b=0, d=0, public: Optional[int] = None) -> int:
        """gate_mul (composer.rs:411-417): c := q_m a b + q_f d + q_c + PI."""
        return self._evaluated(sel, a, b, d, public)

    def arrays(self) -> CircuitArrays:
        n = len(self.gates)
        selectors = b"".join(_mont_bytes([g[0].get(k, 0) for g in self.gates]) for k in SELECTORS)
        wires = b"".join(int(g[1 + col]).to_bytes(4, "little") for col in range(4) for g in self.gates)
        idx = sorted(self.public_inputs)
        return CircuitArrays(n, selectors, wires, _mont_bytes(self.witnesses), b"".join(i.to_bytes(8, "little") for i in idx),
                             _mont_bytes([self.public_inputs[i] for i in idx]))


def synthetic_circuit(n_gates: int, seed: int, n_public: int = 2) -> Composer:
    """The bench workload: a satisfied arithmetic-gate circuit with random
    witnesses, random copy constraints and a few public inputs, exactly `n_gates` constraints."""
    comp = Composer.initialized()
    state = [seed & 0xFFFFFFFFFFFFFFFF]

    def nxt() -> int:
        state[0] = (state[0] + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = state[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def rfr() -> int:
        return (nxt() | (nxt() << 64) | (nxt() << 128) | (nxt() << 192)) % R_MOD

    pool = [comp.append_witness(rfr()) for _ in range(4)]
    for _ in range(n_public):
        pool.append(comp.append_public(rfr()))
    while comp.constraints() < n_gates:
        a, b, d = (pool[nxt() % len(pool)] for _ in range(3))
        kind = nxt() % 3
        if kind == 0:
            c = comp.gate_mul(dict(q_m=rfr(), q_c=rfr()), a=a, b=b, d=d)
        elif kind == 1:
            c = comp.gate_add(dict(q_l=rfr(), q_r=rfr(), q_f=rfr()), a=a, b=b, d=d)
        else:
            c = comp.gate_mul(dict(q_m=1, q_l=rfr(), q_f=1, q_c=nxt()), a=a, b=b, d=d)
        pool.append(c)
        if len(pool) > 64:
            pool.pop(nxt() % 32)
    return comp
