// Inverse NTT over G1 points: the Lagrange form of a commit key.
//
// [L_j(x)]G = (1/n) sum_i w^(-ij) [x^i]G for the first n = 2^k powers of the key (reference: the
// polynomials behind CommitKey::commit, src/commitment_scheme/kzg10/key.rs:376-388, are then committed
// through their evaluations instead of their coefficients - DESIGN.md, next steps).  Radix-2
// decimation in frequency on XYZZ points, natural order in, bit-reversed order out; the twiddle
// multiplication is a double-and-add scalar multiplication.  The butterfly is host-compilable so that
// tests/hosttest can run the whole transform on the CPU against the Python model.
#pragma once
#include "g1.cuh"

namespace pb {

// (a, b) <- (a + b, (a - b) * tw), tw = w_n^(-e) in Montgomery form (tw_is_one: e == 0)
PB_HD void ec_butterfly(G1Xyzz& a, G1Xyzz& b, const Fr& tw, bool tw_is_one) {
  G1Xyzz s = a, d = a;
  xyzz_add(s, b);
  xyzz_add(d, b.neg());
  a = s;
  if (tw_is_one) {
    b = d;
  } else {
    const Fr c = tw.from_mont();
    b = xyzz_mul(d, c.v, 8);
  }
}

PB_HD unsigned ec_bitrev(unsigned i, int log_n) {
  unsigned r = 0;
  for (int b = 0; b < log_n; b++) r |= ((i >> b) & 1u) << (log_n - 1 - b);
  return r;
}

}  // namespace pb
