"""How many non-zero signed 16-bit window digits do the wire VALUES of a circuit have?

A commitment to a wire polynomial in the Lagrange basis (sum_i w_i [L_i(x)]G + the blinder terms) is an
MSM whose scalars are the witness values themselves, so every zero digit is a bucket addition that does
not happen; in the monomial basis (what CommitKey::commit gets today, reference
src/compiler/prover.rs:187-210) the scalars are interpolated coefficients, full-range whatever the witness.
Host only (the native composer of libplonk_b200; no GPU).

    python tools/witness_digit_stats.py [log2 gates, default 16]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonk_b200.gadgets import bench_circuit  # noqa: E402

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
R_INV = pow(1 << 256, -1, R_MOD)


def digits(v: int) -> int:
    v = min(v, R_MOD - v)  # the negated point is free
    n = carry = 0
    for w in range(16):
        d = ((v >> (16 * w)) & 0xFFFF) + carry
        carry = 1 if d > 0x8000 else 0
        n += 1 if (d & 0xFFFF if carry else d) else 0
    return n


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    arr = bench_circuit(1 << lg).arrays()
    vals = [int.from_bytes(arr.witnesses[32 * i : 32 * i + 32], "little") * R_INV % R_MOD for i in range(arr.n_witnesses)]
    wires = np.frombuffer(arr.wires, dtype=np.uint32).reshape(4, -1)
    nd = [digits(v) for v in vals]
    hist = [0] * 17
    for col in range(4):
        for i in range(arr.constraints):
            hist[nd[wires[col][i]]] += 1
    slots = 4 * arr.constraints
    avg = sum(k * h for k, h in enumerate(hist)) / slots
    print(f"BenchCircuit<2^{lg}>: {arr.constraints} gates, {slots} wire slots")
    print(f"  zero values {hist[0]} ({100 * hist[0] / slots:.1f} %), one digit {hist[1]}, all sixteen {hist[16]}")
    print(f"  average non-zero digits per wire value: {avg:.2f} of 16 (monomial-basis coefficients: ~16)")
    print(f"  histogram by digit count: {hist}")


if __name__ == "__main__":
    main()
