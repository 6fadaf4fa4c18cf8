// BLS12-381 G1 point arithmetic for the bucket MSM (y^2 = x^3 + 4 over Fp, a = 0).
//
// Replaces the group arithmetic inside dusk_bls12_381::multiscalar_mul::msm_variable_base, the
// callee of CommitKey::commit (reference src/commitment_scheme/kzg10/key.rs:376-388).  The MSM
// result is representation independent once normalised to affine (Commitment::from,
// reference src/commitment_scheme/kzg10/commitment.rs:89-93), so we are free to pick the cheapest
// coordinates: buckets are accumulated in XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), where a mixed
// addition of an affine base costs 8M + 2S and needs no field inversion.
#pragma once
#include "bigint.cuh"

namespace pb {

// Affine base point as stored in HBM: x, y Montgomery limbs; the identity is x = y = 0 (not on the
// curve, so the encoding is unambiguous).
struct G1Affine {
  Fp x, y;
  PB_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
};

struct G1Xyzz {
  Fp x, y, zz, zzz;
  static PB_HD G1Xyzz identity() {
    G1Xyzz r;
    r.x = Fp::zero();
    r.y = Fp::zero();
    r.zz = Fp::zero();
    r.zzz = Fp::zero();
    return r;
  }
  PB_HD bool is_inf() const { return zz.is_zero(); }
  static PB_HD G1Xyzz from_affine(const G1Affine& p) {
    G1Xyzz r;
    if (p.is_inf()) return identity();
    r.x = p.x;
    r.y = p.y;
    r.zz = Fp::one();
    r.zzz = Fp::one();
    return r;
  }
  PB_HD G1Xyzz neg() const {
    G1Xyzz r = *this;
    r.y = y.neg();
    return r;
  }
};

// 2*P for affine P (mdbl-2008-s-1).  P must not be the identity.
PB_HD G1Xyzz xyzz_dbl_affine(const Fp& x1, const Fp& y1) {
  G1Xyzz r;
  if (y1.is_zero()) return G1Xyzz::identity();  // 2-torsion: cannot happen in the prime-order group
  Fp u = y1.dbl();
  Fp v = u.sqr();
  Fp w = u * v;
  Fp s = x1 * v;
  Fp xx = x1.sqr();
  Fp m = xx.dbl() + xx;
  r.x = m.sqr() - s.dbl();
  r.y = Fp::mul_sub(m, s - r.x, w, y1);  // two products, one reduction
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2*P in XYZZ (dbl-2008-s-1).
PB_HD G1Xyzz xyzz_dbl(const G1Xyzz& p) {
  if (p.is_inf()) return p;
  if (p.y.is_zero()) return G1Xyzz::identity();
  G1Xyzz r;
  Fp u = p.y.dbl();
  Fp v = u.sqr();
  Fp w = u * v;
  Fp s = p.x * v;
  Fp xx = p.x.sqr();
  Fp m = xx.dbl() + xx;
  r.x = m.sqr() - s.dbl();
  r.y = Fp::mul_sub(m, s - r.x, w, p.y);
  r.zz = v * p.zz;
  r.zzz = w * p.zzz;
  return r;
}

// acc += (x2, +-y2) for an affine, non-identity point (madd-2008-s), all special cases handled:
// acc = identity, acc == P (doubling), acc == -P (result is the identity).
PB_HD void xyzz_madd(G1Xyzz& acc, const Fp& x2, const Fp& y2) {
  if (acc.is_inf()) {
    acc.x = x2;
    acc.y = y2;
    acc.zz = Fp::one();
    acc.zzz = Fp::one();
    return;
  }
  Fp u2 = x2 * acc.zz;
  Fp s2 = y2 * acc.zzz;
  Fp p = u2 - acc.x;
  Fp r = s2 - acc.y;
  if (p.is_zero()) {
    if (r.is_zero())
      acc = xyzz_dbl_affine(x2, y2);
    else
      acc = G1Xyzz::identity();
    return;
  }
  Fp pp = p.sqr();
  Fp ppp = p * pp;
  Fp q = acc.x * pp;
  Fp x3 = r.sqr() - ppp - q.dbl();
  Fp y3 = Fp::mul_sub(r, q - x3, acc.y, ppp);
  acc.x = x3;
  acc.y = y3;
  acc.zz = acc.zz * pp;
  acc.zzz = acc.zzz * ppp;
}

// acc += o, both XYZZ (add-2008-s), all special cases handled.
PB_HD void xyzz_add(G1Xyzz& acc, const G1Xyzz& o) {
  if (o.is_inf()) return;
  if (acc.is_inf()) {
    acc = o;
    return;
  }
  Fp u1 = acc.x * o.zz;
  Fp u2 = o.x * acc.zz;
  Fp s1 = acc.y * o.zzz;
  Fp s2 = o.y * acc.zzz;
  Fp p = u2 - u1;
  Fp r = s2 - s1;
  if (p.is_zero()) {
    if (r.is_zero())
      acc = xyzz_dbl(acc);
    else
      acc = G1Xyzz::identity();
    return;
  }
  Fp pp = p.sqr();
  Fp ppp = p * pp;
  Fp q = u1 * pp;
  Fp x3 = r.sqr() - ppp - q.dbl();
  Fp y3 = Fp::mul_sub(r, q - x3, s1, ppp);
  acc.x = x3;
  acc.y = y3;
  acc.zz = acc.zz * o.zz * pp;
  acc.zzz = acc.zzz * o.zzz * ppp;
}

// k * P for a canonical (non-Montgomery) little-endian integer k of `words` 32-bit words: plain
// double-and-add, variable time (public data only: twiddle factors of the group-element NTT).
PB_HD G1Xyzz xyzz_mul(const G1Xyzz& p, const uint32_t* k, int words) {
  G1Xyzz acc = G1Xyzz::identity();
  bool started = false;
  for (int w = words - 1; w >= 0; w--) {
    for (int bit = 31; bit >= 0; bit--) {
      if (started) acc = xyzz_dbl(acc);
      if ((k[w] >> bit) & 1u) {
        if (started) {
          xyzz_add(acc, p);
        } else {
          acc = p;
          started = true;
        }
      }
    }
  }
  return acc;
}

// Affine normalisation (one inversion): x = X/ZZ, y = Y/ZZZ.
PB_HD G1Affine xyzz_to_affine(const G1Xyzz& p) {
  G1Affine r;
  if (p.is_inf()) {
    r.x = Fp::zero();
    r.y = Fp::zero();
    return r;
  }
  Fp i = (p.zz * p.zzz).inv();
  Fp izz = i * p.zzz;
  Fp izzz = i * p.zz;
  r.x = p.x * izz;
  r.y = p.y * izzz;
  return r;
}

}  // namespace pb
