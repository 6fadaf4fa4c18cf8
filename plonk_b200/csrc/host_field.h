// Host-side (CPU) 64-bit-limb Montgomery arithmetic for the few serial tails the GPU path leaves
// on the host: the last Horner step of the MSM bucket reduction, projective->affine normalisation
// (Commitment::from, reference src/commitment_scheme/kzg10/commitment.rs:89-93), G1 compression
// (commitment.rs:95-101) and the Fiat-Shamir scalars of the prover.  Product code (part of
// libplonk_b200), independent of oracle/.
#pragma once
#include <stdint.h>
#include <string.h>

namespace pbh {

typedef unsigned __int128 u128;

template <int N>
struct Mod64 {
  uint64_t p[N];
  uint64_t inv;    // -p^-1 mod 2^64
  uint64_t r1[N];  // 2^(64N) mod p
  uint64_t r2[N];  // 2^(128N) mod p
};

template <int N, const Mod64<N>& M>
struct HField {
  uint64_t v[N];

  static HField zero() { HField r; memset(r.v, 0, sizeof r.v); return r; }
  static HField one() { HField r; memcpy(r.v, M.r1, sizeof r.v); return r; }
  bool is_zero() const { uint64_t x = 0; for (int i = 0; i < N; i++) x |= v[i]; return x == 0; }
  bool operator==(const HField& o) const { return memcmp(v, o.v, sizeof v) == 0; }
  bool operator!=(const HField& o) const { return !(*this == o); }

  static bool geq_p(const uint64_t* t) {
    for (int i = N - 1; i >= 0; i--) {
      if (t[i] > M.p[i]) return true;
      if (t[i] < M.p[i]) return false;
    }
    return true;
  }
  static void sub_p(uint64_t* t) {
    u128 borrow = 0;
    for (int i = 0; i < N; i++) {
      u128 d = (u128)t[i] - M.p[i] - borrow;
      t[i] = (uint64_t)d;
      borrow = (d >> 64) & 1;
    }
  }
  friend HField operator+(const HField& a, const HField& b) {
    HField r;
    u128 c = 0;
    for (int i = 0; i < N; i++) {
      c += (u128)a.v[i] + b.v[i];
      r.v[i] = (uint64_t)c;
      c >>= 64;
    }
    if (c || geq_p(r.v)) sub_p(r.v);
    return r;
  }
  friend HField operator-(const HField& a, const HField& b) {
    HField r;
    u128 borrow = 0;
    for (int i = 0; i < N; i++) {
      u128 d = (u128)a.v[i] - b.v[i] - borrow;
      r.v[i] = (uint64_t)d;
      borrow = (d >> 64) & 1;
    }
    if (borrow) {
      u128 c = 0;
      for (int i = 0; i < N; i++) {
        c += (u128)r.v[i] + M.p[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
      }
    }
    return r;
  }
  HField neg() const { return zero() - *this; }
  HField dbl() const { return *this + *this; }
  friend HField operator*(const HField& a, const HField& b) {  // CIOS
    uint64_t t[N + 2];
    memset(t, 0, sizeof t);
    for (int i = 0; i < N; i++) {
      u128 c = 0;
      for (int j = 0; j < N; j++) {
        c += (u128)a.v[j] * b.v[i] + t[j];
        t[j] = (uint64_t)c;
        c >>= 64;
      }
      c += t[N];
      t[N] = (uint64_t)c;
      t[N + 1] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * M.inv;
      c = (u128)m * M.p[0] + t[0];
      c >>= 64;
      for (int j = 1; j < N; j++) {
        c += (u128)m * M.p[j] + t[j];
        t[j - 1] = (uint64_t)c;
        c >>= 64;
      }
      c += t[N];
      t[N - 1] = (uint64_t)c;
      t[N] = t[N + 1] + (uint64_t)(c >> 64);
    }
    HField r;
    memcpy(r.v, t, sizeof r.v);
    if (t[N] || geq_p(r.v)) sub_p(r.v);
    return r;
  }
  HField sqr() const { return (*this) * (*this); }
  HField pow(const uint64_t* e, int words) const {
    HField acc = one();
    for (int w = words - 1; w >= 0; w--)
      for (int bit = 63; bit >= 0; bit--) {
        acc = acc.sqr();
        if ((e[w] >> bit) & 1) acc = acc * (*this);
      }
    return acc;
  }
  HField inv() const {  // Fermat; 0 -> 0
    uint64_t e[N];
    memcpy(e, M.p, sizeof e);
    e[0] -= 2;  // p is odd and p[0] >= 2 for both moduli
    return pow(e, N);
  }
  HField to_mont() const { HField r2; memcpy(r2.v, M.r2, sizeof r2.v); return (*this) * r2; }
  HField from_mont() const { HField o = zero(); o.v[0] = 1; return (*this) * o; }
  static HField from_u64(uint64_t x) { HField r = zero(); r.v[0] = x; return r.to_mont(); }
  // canonical integer comparison helper: is (this, canonical form) > (p-1)/2 ?
  bool canonical_gt_half() const {
    HField c = from_mont();
    // compare 2*c with p: c > (p-1)/2  <=>  2c > p - 1  <=>  2c >= p + 1 > p (p odd)
    uint64_t t[N + 1];
    uint64_t carry = 0;
    for (int i = 0; i < N; i++) {
      t[i] = (c.v[i] << 1) | carry;
      carry = c.v[i] >> 63;
    }
    if (carry) return true;
    for (int i = N - 1; i >= 0; i--) {
      if (t[i] > M.p[i]) return true;
      if (t[i] < M.p[i]) return false;
    }
    return false;
  }
};

extern const Mod64<6> kFpMod;
extern const Mod64<4> kFrMod;
typedef HField<6, kFpMod> HFp;
typedef HField<4, kFrMod> HFr;

// G1 in XYZZ coordinates on the host (same formulas as csrc/g1.cuh).
struct HXyzz {
  HFp x, y, zz, zzz;
  static HXyzz identity() { HXyzz r; r.x = r.y = r.zz = r.zzz = HFp::zero(); return r; }
  bool is_inf() const { return zz.is_zero(); }
};
HXyzz hxyzz_dbl(const HXyzz& p);
void hxyzz_add(HXyzz& acc, const HXyzz& o);
// returns false for the identity; out x,y in Montgomery form
bool hxyzz_to_affine(const HXyzz& p, HFp* x, HFp* y);
void g1_compress_raw(const uint64_t* affine_raw /*12 limbs*/, uint8_t out[48]);

}  // namespace pbh
