// EXPERIMENT (not on the product path): an Fp Montgomery product entirely on the FP64 pipe.
// Operands are eight 52-bit limbs held as doubles (exact integers), R = 2^416.  Every limb product
// is split exactly into two 52-bit halves by two DFMA.RZ and one DADD,
//   hi = fma_rz(a, b, 2^104)                  bits(hi) = bits(2^104) + floor(ab / 2^52)
//   lo = fma_rz(a, b, (2^104 + 2^52) - hi)    bits(lo) = bits(2^52)  + (ab mod 2^52)
// and the raw patterns are summed into 64-bit integer columns by IADD3 (offsets are multiples of
// 2^52, so the low 52 bits of a column are always right).  CIOS by rows of b.  Stands in for
// dusk-bls12_381's Fp product behind msm_variable_base (reference
// src/commitment_scheme/kzg10/key.rs:384); measured against pb::Fp's IMAD product by mulbench.
#pragma once
#include <stdint.h>
#include "../../plonk_b200/csrc/bigint.cuh"

namespace pb52 {
using pb::bits_dbl;
using pb::dbl_bits;
using pb::fma_rz;

struct F52 {
  double l[8];
};

PB_HD constexpr uint64_t p52(int j) {
  constexpr uint64_t T[8] = {0xeffffffffaaabull, 0xfeb153ffffb9full, 0x6b0f6241eabffull, 0x12bf6730d2a0full,
                             0x764774b84f385ull, 0x1ba7b6434bacdull, 0x1ea397fe69a4bull, 0x1a011ull};
  return T[j];
}
constexpr uint32_t PINV52_LO = 0xfffcfffdu, PINV52_HI = 0x3fffcu;  // -p^-1 mod 2^52
constexpr uint64_t B52 = 0x4330000000000000ull, B104 = 0x4670000000000000ull, M52 = 0xfffffffffffffull;

// a * b * 2^-416 mod p, limbs normalised to 52 bits, value fully reduced (inputs < p).
PB_HD F52 mul52(const F52& a, const F52& b) {
  uint64_t c[9], off[9];  // off[] tracks the pattern offsets; it is data independent and folds away
#pragma unroll
  for (int k = 0; k < 9; k++) c[k] = off[k] = 0;
  const double C1 = 0x1p104, C2 = 0x1p104 + 0x1p52;
#pragma unroll
  for (int i = 0; i < 8; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const double hi = fma_rz(a.l[j], b.l[i], C1);
      const double lo = fma_rz(a.l[j], b.l[i], C2 - hi);
      c[j] += dbl_bits(lo);
      c[j + 1] += dbl_bits(hi);
      off[j] += B52;
      off[j + 1] += B104;
    }
    // m = c[0] * (-p^-1) mod 2^52 on the (otherwise idle) integer multiply pipe
    const uint32_t qlo = (uint32_t)c[0], qhi = (uint32_t)(c[0] >> 32);
    const uint64_t r = (uint64_t)qlo * PINV52_LO;
    const uint32_t rhi = ((uint32_t)(r >> 32) + qlo * PINV52_HI + qhi * PINV52_LO) & 0xfffffu;
    const double m = bits_dbl(0x43300000u | rhi, (uint32_t)r) - 0x1p52;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const double pj = (double)p52(j);
      const double hi = fma_rz(m, pj, C1);
      const double lo = fma_rz(m, pj, C2 - hi);
      c[j] += dbl_bits(lo);
      c[j + 1] += dbl_bits(hi);
      off[j] += B52;
      off[j + 1] += B104;
    }
    // column 0 is now 0 mod 2^52: shift the window down by one limb
    const uint64_t carry = (c[0] - off[0]) >> 52;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      c[k] = c[k + 1];
      off[k] = off[k + 1];
    }
    c[8] = 0;
    off[8] = 0;
    c[0] += carry;
  }
  uint64_t t[8];
#pragma unroll
  for (int k = 0; k < 8; k++) t[k] = c[k] - off[k];
#pragma unroll
  for (int k = 0; k < 7; k++) {
    t[k + 1] += t[k] >> 52;
    t[k] &= M52;
  }
  // conditional subtraction of p (the result is below 2p)
  uint64_t s[8], borrow = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const uint64_t d = t[k] - p52(k) - borrow;
    borrow = d >> 63;
    s[k] = d & M52;
  }
  F52 res;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const uint64_t v = borrow ? t[k] : s[k];
    res.l[k] = bits_dbl(0x43300000u | (uint32_t)(v >> 32), (uint32_t)v) - 0x1p52;
  }
  return res;
}

// conversions for the host check
inline F52 from_words(const uint32_t* w) {  // 384-bit little-endian -> 52-bit limbs
  F52 r;
  for (int k = 0; k < 8; k++) {
    uint64_t v = 0;
    for (int bit = 0; bit < 52; bit++) {
      const int pos = 52 * k + bit;
      if (pos < 384 && ((w[pos / 32] >> (pos % 32)) & 1u)) v |= 1ull << bit;
    }
    r.l[k] = (double)v;
  }
  return r;
}
inline void to_words(const F52& x, uint32_t* w) {
  for (int i = 0; i < 12; i++) w[i] = 0;
  for (int k = 0; k < 8; k++) {
    const uint64_t v = (uint64_t)x.l[k];
    for (int bit = 0; bit < 52; bit++) {
      const int pos = 52 * k + bit;
      if (pos < 384 && ((v >> bit) & 1ull)) w[pos / 32] |= 1u << (pos % 32);
    }
  }
}

}  // namespace pb52
