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


EDWARDS_D = (-10240 * pow(10241, R_MOD - 2, R_MOD)) % R_MOD  # dusk_jubjub::EDWARDS_D
_TWO_ADIC_ROOT = pow(7, (R_MOD - 1) >> 32, R_MOD)


def _fr_sqrt(a: int):
    """Tonelli-Shanks over Fr (2-adicity 32); None for non-residues."""
    a %= R_MOD
    if a == 0:
        return 0
    if pow(a, (R_MOD - 1) // 2, R_MOD) != 1:
        return None
    q, s = R_MOD - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    m, c, t, r = s, _TWO_ADIC_ROOT, pow(a, q, R_MOD), pow(a, (q + 1) // 2, R_MOD)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % R_MOD
            i += 1
        b = pow(c, 1 << (m - i - 1), R_MOD)
        m, c = i, b * b % R_MOD
        t, r = t * c % R_MOD, r * b % R_MOD
    return r


def _jubjub_from_y(y: int):
    x = _fr_sqrt((y * y - 1) * pow((1 + EDWARDS_D * y * y) % R_MOD, R_MOD - 2, R_MOD) % R_MOD)
    return None if x is None else (x, y % R_MOD)


def _jubjub_add(p, q):
    (x1, y1), (x2, y2) = p, q
    t = EDWARDS_D * x1 * x2 % R_MOD * y1 % R_MOD * y2 % R_MOD
    inv = lambda v: pow(v % R_MOD, R_MOD - 2, R_MOD)
    return ((x1 * y2 + y1 * x2) * inv(1 + t) % R_MOD, (y1 * y2 + x1 * x2) * inv(1 - t) % R_MOD)


def widget_rows(comp: "Composer", nxt, n_range: int, n_logic: int, n_fixed: int, n_var: int) -> None:
    """Satisfied rows for the range, logic, fixed-base and curve-addition gate families, written from
    the identities their widgets enforce (reference src/proof_system/widget/{range,logic,ecc/**}/
    proverkey.rs).  The reference builds such rows with its gadget library
    (src/composer/{range,logic,fixed_base,point}.rs); the prover backend only sees the rows."""
    link = lambda **w: comp.append_custom_gate({}, **w)
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
    if n_logic:
        for q_c in (1, -1):
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

    def point():
        while True:
            pt = _jubjub_from_y(nxt() | (nxt() << 64) | (nxt() << 128) | ((nxt() >> 3) << 192))
            if pt is not None and pt[0] != 0:
                return pt

    if n_fixed:
        acc_pt, scalar = point(), nxt() % 1000
        wx, wy, wd = comp.append_witness(acc_pt[0]), comp.append_witness(acc_pt[1]), comp.append_witness(scalar)
        for _ in range(n_fixed):
            beta = point()
            bit = (nxt() % 3) - 1
            alpha = (0, 1) if bit == 0 else (beta if bit == 1 else ((-beta[0]) % R_MOD, beta[1]))
            wc = comp.append_witness(alpha[0] * alpha[1])
            comp.append_custom_gate(dict(q_fixed_group_add=1, q_l=beta[0], q_r=beta[1], q_c=beta[0] * beta[1]), a=wx, b=wy, c=wc, d=wd)
            acc_pt = _jubjub_add(acc_pt, alpha)
            scalar = (2 * scalar + bit) % R_MOD
            wx, wy, wd = comp.append_witness(acc_pt[0]), comp.append_witness(acc_pt[1]), comp.append_witness(scalar)
        link(a=wx, b=wy, d=wd)
    for _ in range(n_var):
        p1, p2 = point(), point()
        p3 = _jubjub_add(p1, p2)
        comp.append_custom_gate(dict(q_variable_group_add=1), a=comp.append_witness(p1[0]), b=comp.append_witness(p1[1]),
                                c=comp.append_witness(p2[0]), d=comp.append_witness(p2[1]))
        link(a=comp.append_witness(p3[0]), b=comp.append_witness(p3[1]), d=comp.append_witness(p1[0] * p2[1]))


def synthetic_circuit(n_gates: int, seed: int, n_public: int = 2, widgets: int = 0) -> Composer:
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

    if widgets:
        widget_rows(comp, nxt, widgets, widgets, widgets, widgets)
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
