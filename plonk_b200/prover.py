"""Prover mirror (reference src/compiler/prover.rs: Prover::new :53-115, Prover::prove :352-362)
over the device-resident CUDA prover.

A circuit crosses the boundary as flat arrays, i.e. what Compiler::preprocess reads out of the
Composer (reference src/compiler.rs:132-170): 11 selector columns, 4 wire columns, the witness
table and the sparse public inputs."""
from __future__ import annotations

import ctypes

from ._lib import PB200_ERR_UNSATISFIED, Pb200Error, check, lib

PROOF_BYTES = 1008


class CircuitUnsatisfied(ValueError):
    """Error::CircuitUnsatisfied (reference src/proof_system/quotient_poly.rs:132-134)."""


class Prover:
    def __init__(self, label: bytes, n_constraints: int, selectors: bytes, wires: bytes, n_witnesses: int, srs_raw: bytes):
        assert len(selectors) == 11 * n_constraints * 32 and len(wires) == 4 * n_constraints * 4
        h = ctypes.c_void_p()
        check(lib().pb200_prover_new(label, len(label), n_constraints, selectors, wires, n_witnesses, srs_raw,
                                     len(srs_raw) // 96, ctypes.byref(h)))
        self._h = h
        self.n_constraints = n_constraints
        self.n_witnesses = n_witnesses

    @classmethod
    def from_bytes(cls, prover_bytes: bytes, wires: bytes, n_witnesses: int) -> "Prover":
        """Prover::try_from_bytes (prover.rs:265-350) for the output of Prover::to_bytes; the circuit's wiring
        (4 x constraints u32) is not part of that format and comes alongside."""
        self = cls.__new__(cls)
        h = ctypes.c_void_p()
        check(lib().pb200_prover_from_bytes(prover_bytes, len(prover_bytes), wires, n_witnesses, ctypes.byref(h)))
        self._h = h
        self.n_constraints = len(wires) // 16
        self.n_witnesses = n_witnesses
        return self

    def commitments(self):
        out = ctypes.create_string_buffer(15 * 48)
        check(lib().pb200_prover_commitments(self._h, out))
        return [out.raw[48 * i : 48 * i + 48] for i in range(15)]

    def prove(self, witnesses: bytes, pi_idx: bytes, pi_vals: bytes, blinders: bytes) -> bytes:
        """witnesses: n_witnesses x 32 B; pi_idx: u64 LE positions; pi_vals: 32 B each; blinders: 14 x 32 B."""
        assert len(blinders) == 14 * 32 and len(witnesses) == self.n_witnesses * 32
        n_pi = len(pi_idx) // 8
        out = ctypes.create_string_buffer(PROOF_BYTES)
        try:
            check(lib().pb200_prove(self._h, witnesses, self.n_witnesses, pi_idx or None, pi_vals or None, n_pi, blinders, out))
        except Pb200Error as e:
            if e.code == PB200_ERR_UNSATISFIED:
                raise CircuitUnsatisfied() from e
            raise
        return out.raw

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().pb200_prover_free(self._h)
                self._h = None
        except Exception:
            pass
