// Multi-limb (32-bit) Montgomery arithmetic for sm_100a, generic over the limb count.
//
// The GPU arithmetic underneath the two hot kernels of the dusk-plonk prover: BlsScalar (Fr, 8 limbs)
// for EvaluationDomain::{fft,ifft,coset_fft,coset_ifft} (reference src/fft/domain.rs:166-232) and
// Fp (12 limbs) for the G1 MSM behind CommitKey::commit (reference
// src/commitment_scheme/kzg10/key.rs:376-388).  The reference keeps both in dusk-bls12_381 as
// 4x/6x u64 Montgomery limbs; the in-memory little-endian layout is identical, we just address it
// as 32-bit limbs because the Blackwell integer pipe is IMAD (32x32+64).
//
// Every carry chain is written with PTX add.cc/addc/mad.lo.cc/madc.hi.cc so ptxas can fuse
// lo/hi pairs into IMAD.WIDE.U32 with predicate carries.  When compiled for the host (g++, used
// only by tests/hosttest to validate the algorithms without a GPU) the same primitives are
// emulated with an explicit carry flag.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#define PB_D __device__ __forceinline__
#else
#define PB_HD inline
#define PB_D inline
#endif

#ifndef PB_SPLIT
#define PB_SPLIT 0
#endif

namespace pb {

#if defined(__CUDA_ARCH__)
#define PB_ASM2(name, ins)                                                         \
  PB_D uint32_t name(uint32_t a, uint32_t b) {                                     \
    uint32_t r;                                                                    \
    asm volatile(ins " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                   \
    return r;                                                                      \
  }
#define PB_ASM3(name, ins)                                                         \
  PB_D uint32_t name(uint32_t a, uint32_t b, uint32_t c) {                         \
    uint32_t r;                                                                    \
    asm volatile(ins " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));       \
    return r;                                                                      \
  }
PB_ASM2(add_cc, "add.cc.u32")
PB_ASM2(addc_cc, "addc.cc.u32")
PB_ASM2(addc, "addc.u32")
PB_ASM2(sub_cc, "sub.cc.u32")
PB_ASM2(subc_cc, "subc.cc.u32")
PB_ASM2(subc, "subc.u32")
PB_ASM2(mul_lo, "mul.lo.u32")
PB_ASM2(mul_hi, "mul.hi.u32")
PB_ASM3(mad_lo_cc, "mad.lo.cc.u32")
PB_ASM3(madc_lo_cc, "madc.lo.cc.u32")
PB_ASM3(mad_hi_cc, "mad.hi.cc.u32")
PB_ASM3(madc_hi_cc, "madc.hi.cc.u32")
PB_ASM3(madc_hi, "madc.hi.u32")
PB_ASM3(madc_lo, "madc.lo.u32")
#undef PB_ASM2
#undef PB_ASM3
// (lo,hi) pair forms: both halves of one 32x32 product in a single asm block, which is the shape
// ptxas fuses into one IMAD.WIDE.U32[.X] with predicate carry-in/out.
//   CIN: consume the carry flag; COUT: leave the carry flag set for the next pair.
template <bool CIN, bool COUT>
PB_D void mad_pair(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  if (CIN && COUT)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
  else if (CIN && !COUT)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
  else if (!CIN && COUT)
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
  else
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
PB_D void mul_pair(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// Same contract as mad_pair, but the 64-bit product is formed by a carry-free IMAD.WIDE and added
// with two carry-chained IADD3 on the ALU pipe.  IMAD.WIDE.U32.X (carry in/out) issues at half the
// rate of the carry-free form (tools/mulbench), so moving part of the links of a chain to this
// shape balances the multiply pipe against the otherwise idle ALU pipe.
template <bool CIN, bool COUT>
PB_D void mad_pair_split(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  uint32_t tl, th;
  asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(tl), "=r"(th) : "r"(a), "r"(b));
  if (CIN && COUT)
    asm volatile("addc.cc.u32 %0, %2, %3; addc.cc.u32 %1, %4, %5;" : "=r"(lo), "=r"(hi) : "r"(clo), "r"(tl), "r"(chi), "r"(th));
  else if (CIN && !COUT)
    asm volatile("addc.cc.u32 %0, %2, %3; addc.u32 %1, %4, %5;" : "=r"(lo), "=r"(hi) : "r"(clo), "r"(tl), "r"(chi), "r"(th));
  else if (!CIN && COUT)
    asm volatile("add.cc.u32 %0, %2, %3; addc.cc.u32 %1, %4, %5;" : "=r"(lo), "=r"(hi) : "r"(clo), "r"(tl), "r"(chi), "r"(th));
  else
    asm volatile("add.cc.u32 %0, %2, %3; addc.u32 %1, %4, %5;" : "=r"(lo), "=r"(hi) : "r"(clo), "r"(tl), "r"(chi), "r"(th));
}
#else
// Host emulation of the PTX condition-code register (one flag, as in PTX).
static thread_local uint32_t pb_cf = 0;
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; pb_cf = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + pb_cf; pb_cf = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + pb_cf; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b; pb_cf = (uint32_t)(s >> 63); return (uint32_t)s; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - pb_cf; pb_cf = (uint32_t)(s >> 63); return (uint32_t)s; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - pb_cf; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(a * b, c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(a * b, c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return addc(a * b, c); }
template <bool CIN, bool COUT>
inline void mad_pair(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  uint32_t l = CIN ? madc_lo_cc(a, b, clo) : mad_lo_cc(a, b, clo);
  uint32_t h = COUT ? madc_hi_cc(a, b, chi) : madc_hi(a, b, chi);
  lo = l;
  hi = h;
}
inline void mul_pair(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { lo = mul_lo(a, b); hi = mul_hi(a, b); }
template <bool CIN, bool COUT>
inline void mad_pair_split(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  mad_pair<CIN, COUT>(lo, hi, a, b, clo, chi);
}
#endif

// ---------------------------------------------------------------------------------------------
// Field<P>: P supplies N (even), MOD(i), inv() (= -MOD^-1 mod 2^32), R1 (=2^(32N) mod p), R2.
// Values are always kept fully reduced in [0, p), Montgomery form unless stated otherwise.
// ---------------------------------------------------------------------------------------------
template <class P>
struct Field {
  static constexpr int N = P::N;
  uint32_t v[N];

  static PB_HD Field zero() {
    Field r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
    return r;
  }
  static PB_HD Field one() {  // Montgomery form of 1
    Field r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::R1(i);
    return r;
  }
  static PB_HD Field r2() {
    Field r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::R2(i);
    return r;
  }
  PB_HD bool is_zero() const {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < N; i++) x |= v[i];
    return x == 0;
  }
  PB_HD bool operator==(const Field& o) const {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < N; i++) x |= v[i] ^ o.v[i];
    return x == 0;
  }
  PB_HD bool operator!=(const Field& o) const { return !(*this == o); }

  // r = (t >= p) ? t - p : t, for t < 2p given with an extra top word `hi` (0 or 1).
  static PB_HD void final_sub(uint32_t* r, const uint32_t* t, uint32_t hi) {
    uint32_t s[N];
    s[0] = sub_cc(t[0], P::MOD(0));
#pragma unroll
    for (int i = 1; i < N; i++) s[i] = subc_cc(t[i], P::MOD(i));
    uint32_t borrow = subc(hi, 0u);  // 0 if t >= p, 0xffffffff (or hi-1) otherwise
    bool ge = (borrow == 0u);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = ge ? s[i] : t[i];
  }

  friend PB_HD Field operator+(const Field& a, const Field& b) {
    uint32_t t[N];
    t[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = addc_cc(a.v[i], b.v[i]);
    uint32_t hi = addc(0u, 0u);
    Field r;
    final_sub(r.v, t, hi);
    return r;
  }
  friend PB_HD Field operator-(const Field& a, const Field& b) {
    uint32_t t[N];
    t[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = subc_cc(a.v[i], b.v[i]);
    uint32_t borrow = subc(0u, 0u);  // 0xffffffff when a < b
    Field r;
    uint32_t m = borrow;  // add p back under mask
    r.v[0] = add_cc(t[0], P::MOD(0) & m);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(t[i], P::MOD(i) & m);
    r.v[N - 1] = addc(t[N - 1], P::MOD(N - 1) & m);
    return r;
  }
  PB_HD Field neg() const { return zero() - *this; }
  PB_HD Field dbl() const { return *this + *this; }

  // Montgomery product a*b*2^(-32N) mod p (CIOS by rows of b).
  //
  // The running sum T is kept as two N-limb accumulators, T = E + O*2^32: products of the even
  // limbs of the multiplicand land in E as (lo,hi) pairs at even indices, products of the odd limbs
  // land in O the same way, so every pair is an aligned register pair and a whole row is two carry
  // chains of mad.lo.cc/madc.hi.cc (IMAD.WIDE.U32 with predicate carry in SASS).  After the
  // Montgomery step E[0] == 0 and T/2^32 = O + (E >> 32): the accumulators swap roles (new E = O,
  // new O = E >> 64) and the one left-over limb E[1] is folded into the next row's carry chain.
  template <bool SPLIT, bool CIN, bool COUT>
  static PB_HD void link(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
    if (SPLIT)
      mad_pair_split<CIN, COUT>(lo, hi, a, b, clo, chi);
    else
      mad_pair<CIN, COUT>(lo, hi, a, b, clo, chi);
  }
  // One row of the product: T += a * bi.  On entry X is the old E (its low limb is zero after the
  // previous Montgomery step, so E >> 64 is what remains) and Y the old O; on exit X is the new O and
  // Y the new E.
  static PB_HD void mul_row(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t bi) {
    constexpr bool S0 = (PB_SPLIT & 1) != 0, S1 = (PB_SPLIT & 2) != 0;
    Y[0] = add_cc(Y[0], X[1]);
#pragma unroll
    for (int k = 0; k < N; k += 2) {
      if (k + 2 < N)
        link<S0, true, true>(X[k], X[k + 1], a[k + 1], bi, X[k + 2], X[k + 3]);
      else
        link<S0, true, false>(X[k], X[k + 1], a[k + 1], bi, 0u, 0u);
    }
    link<S1, false, true>(Y[0], Y[1], a[0], bi, Y[0], Y[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) link<S1, true, true>(Y[j], Y[j + 1], a[j], bi, Y[j], Y[j + 1]);
    X[N - 1] = addc(X[N - 1], 0u);
  }
  // A further product in the same row: T += c * di, with X = O and Y = E already in their new roles.
  static PB_HD void acc_row(uint32_t* X, uint32_t* Y, const uint32_t* c, uint32_t di) {
    constexpr bool S0 = (PB_SPLIT & 1) != 0, S1 = (PB_SPLIT & 2) != 0;
    link<S0, false, true>(X[0], X[1], c[1], di, X[0], X[1]);
#pragma unroll
    for (int k = 2; k < N; k += 2) {
      if (k + 2 < N)
        link<S0, true, true>(X[k], X[k + 1], c[k + 1], di, X[k], X[k + 1]);
      else
        link<S0, true, false>(X[k], X[k + 1], c[k + 1], di, X[k], X[k + 1]);
    }
    link<S1, false, true>(Y[0], Y[1], c[0], di, Y[0], Y[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) link<S1, true, true>(Y[j], Y[j + 1], c[j], di, Y[j], Y[j + 1]);
    X[N - 1] = addc(X[N - 1], 0u);
  }
  // The Montgomery step of a row: T += m * p with m chosen so that the low limb (Y[0]) vanishes.
  static PB_HD void red_row(uint32_t* X /* O */, uint32_t* Y /* E */) {
    constexpr bool S2 = (PB_SPLIT & 4) != 0, S3 = (PB_SPLIT & 8) != 0;
    const uint32_t m = mul_lo(Y[0], P::inv());
#pragma unroll
    for (int k = 0; k < N; k += 2) {
      if (k == 0)
        link<S2, false, true>(X[k], X[k + 1], P::MOD(k + 1), m, X[k], X[k + 1]);
      else if (k + 2 < N)
        link<S2, true, true>(X[k], X[k + 1], P::MOD(k + 1), m, X[k], X[k + 1]);
      else
        link<S2, true, false>(X[k], X[k + 1], P::MOD(k + 1), m, X[k], X[k + 1]);
    }
    link<S3, false, true>(Y[0], Y[1], P::MOD(0), m, Y[0], Y[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) link<S3, true, true>(Y[j], Y[j + 1], P::MOD(j), m, Y[j], Y[j + 1]);
    X[N - 1] = addc(X[N - 1], 0u);
  }
  template <bool FIRST>
  static PB_HD void mont_row(uint32_t* X /* old E -> new O */, uint32_t* Y /* old O -> new E */,
                             const uint32_t* a, uint32_t bi) {
    if (!FIRST) mul_row(X, Y, a, bi);
    red_row(X, Y);
  }

  friend PB_HD Field operator*(const Field& a, const Field& b) {
    uint32_t A[N], B[N];
    {
      const uint32_t bi = b.v[0];
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        mul_pair(A[j], A[j + 1], a.v[j], bi);
        mul_pair(B[j], B[j + 1], a.v[j + 1], bi);
      }
      mont_row<true>(B, A, a.v, bi);  // E = A, O = B
    }
#pragma unroll
    for (int i = 1; i < N; i += 2) {
      mont_row<false>(A, B, a.v, b.v[i]);                     // E = B, O = A
      if (i + 1 < N) mont_row<false>(B, A, a.v, b.v[i + 1]);  // E = A, O = B
    }
    // N is even: E = B (B[0] == 0), O = A.  T/2^32 = A + (B >> 32).
    uint32_t t[N];
    t[0] = add_cc(A[0], B[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) t[k] = addc_cc(A[k], B[k + 1]);
    t[N - 1] = addc(A[N - 1], 0u);
    Field r;
    final_sub(r.v, t, 0u);
    return r;
  }
  // Squaring with the symmetric partial products taken once: a^2 = sum_i a_i * B_i with
  // B_i = a_i 2^(32i) + 2 sum_{k>i} a_k 2^(32k), so row i multiplies a_i only by the limbs k >= i of
  // the doubled operand (N(N+1)/2 multiply-adds instead of N^2; the Montgomery steps are unchanged).
  // The skipped links of the shifting accumulator become plain carry-propagating adds.
  // Row 0 adds a_0 * 2a at once, twice what a row of the ordinary product adds, so the running sum
  // needs 3p * 2^32 < 2^(32(N+1)): two spare bits in the top limb.  Fp has three; Fr (one) keeps the
  // ordinary product.
  template <int I>
  static PB_HD uint32_t sq_limb(const uint32_t* a, const uint32_t* a2, int k) {
    return k == I ? a[k] : (k == I + 1 ? (a[k] << 1) : a2[k]);
  }
  template <int I>
  static PB_HD void sqr_row(uint32_t* X, uint32_t* Y, const uint32_t* a, const uint32_t* a2) {
    constexpr bool S0 = (PB_SPLIT & 1) != 0, S1 = (PB_SPLIT & 2) != 0;
    const uint32_t bi = a[I];
    Y[0] = add_cc(Y[0], X[1]);
#pragma unroll
    for (int k = 0; k < N; k += 2) {
      if (k + 1 < I) {  // no product at this limb: shift and propagate the carry
        X[k] = addc_cc(X[k + 2], 0u);
        X[k + 1] = addc_cc(X[k + 3], 0u);
      } else if (k + 2 < N) {
        link<S0, true, true>(X[k], X[k + 1], sq_limb<I>(a, a2, k + 1), bi, X[k + 2], X[k + 3]);
      } else {
        link<S0, true, false>(X[k], X[k + 1], sq_limb<I>(a, a2, k + 1), bi, 0u, 0u);
      }
    }
    constexpr int J0 = (I + 1) & ~1;  // first even limb >= I
    if (J0 < N) {
      link<S1, false, true>(Y[J0], Y[J0 + 1], sq_limb<I>(a, a2, J0), bi, Y[J0], Y[J0 + 1]);
#pragma unroll
      for (int j = J0 + 2; j < N; j += 2) link<S1, true, true>(Y[j], Y[j + 1], sq_limb<I>(a, a2, j), bi, Y[j], Y[j + 1]);
      X[N - 1] = addc(X[N - 1], 0u);
    }
  }
  template <int I>
  static PB_HD void sqr_rows(uint32_t* A, uint32_t* B, const uint32_t* a, const uint32_t* a2) {
    if constexpr (I < N) {
      sqr_row<I>(A, B, a, a2);  // E = B, O = A
      red_row(A, B);
      if constexpr (I + 1 < N) {
        sqr_row<I + 1>(B, A, a, a2);  // E = A, O = B
        red_row(B, A);
      }
      sqr_rows<I + 2>(A, B, a, a2);
    }
  }
  PB_HD Field sqr() const {
    if constexpr ((P::MOD(N - 1) >> 30) != 0) {
      return (*this) * (*this);
    } else {
      return sqr_half();
    }
  }
  PB_HD Field sqr_half() const {
    static_assert((P::MOD(N - 1) >> 30) == 0, "needs two spare bits in the top limb of the modulus");
    uint32_t a2[N];
    a2[0] = v[0] << 1;
#pragma unroll
    for (int k = 1; k < N; k++) a2[k] = (v[k] << 1) | (v[k - 1] >> 31);
    uint32_t A[N], B[N];
    {
      const uint32_t bi = v[0];
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        mul_pair(A[j], A[j + 1], sq_limb<0>(v, a2, j), bi);
        mul_pair(B[j], B[j + 1], sq_limb<0>(v, a2, j + 1), bi);
      }
      red_row(B, A);  // E = A, O = B
    }
    sqr_rows<1>(A, B, v, a2);
    uint32_t t[N];
    t[0] = add_cc(A[0], B[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) t[k] = addc_cc(A[k], B[k + 1]);
    t[N - 1] = addc(A[N - 1], 0u);
    Field r;
    final_sub(r.v, t, 0u);
    return r;
  }

  // a*b + c*d with ONE Montgomery reduction: each row accumulates both partial products before its
  // Montgomery step (3N^2 multiply-adds instead of 4N^2).  Only for moduli with at least two spare
  // bits in the top limb (Fp: 381 of 384 bits), where the running sum keeps fitting the two
  // accumulators and the result stays below 2p; Fr (255 of 256 bits) must not use it.
  static PB_HD Field mul2(const Field& a, const Field& b, const Field& c, const Field& d) {
    static_assert((P::MOD(N - 1) >> 30) == 0, "mul2 needs two spare bits in the top limb of the modulus");
    uint32_t A[N], B[N];
    {
      const uint32_t bi = b.v[0];
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        mul_pair(A[j], A[j + 1], a.v[j], bi);
        mul_pair(B[j], B[j + 1], a.v[j + 1], bi);
      }
      acc_row(B, A, c.v, d.v[0]);
      red_row(B, A);  // E = A, O = B
    }
#pragma unroll
    for (int i = 1; i < N; i += 2) {
      mul_row(A, B, a.v, b.v[i]);
      acc_row(A, B, c.v, d.v[i]);
      red_row(A, B);  // E = B, O = A
      if (i + 1 < N) {
        mul_row(B, A, a.v, b.v[i + 1]);
        acc_row(B, A, c.v, d.v[i + 1]);
        red_row(B, A);  // E = A, O = B
      }
    }
    uint32_t t[N];
    t[0] = add_cc(A[0], B[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) t[k] = addc_cc(A[k], B[k + 1]);
    t[N - 1] = addc(A[N - 1], 0u);
    Field r;
    final_sub(r.v, t, 0u);
    return r;
  }
  // a*b - c*d
  static PB_HD Field mul_sub(const Field& a, const Field& b, const Field& c, const Field& d) { return mul2(a, b, c.neg(), d); }

  // (Interleaving several independent products in program order was tried for instruction-level
  // parallelism and gives nothing: ptxas keeps at most ~6 carry chains in flight - there are only 7
  // predicate registers - and one product already uses 4.  See DESIGN.md section 4.)

  // out of / into Montgomery form
  PB_HD Field from_mont() const {
    Field o = zero();
    o.v[0] = 1;
    return (*this) * o;
  }
  PB_HD Field to_mont() const { return (*this) * r2(); }

  // this^e for a little-endian multi-word exponent (square and multiply, variable time).
  PB_HD Field pow(const uint32_t* e, int words) const {
    Field acc = one();
    bool started = false;  // skip the leading zero bits of the exponent
    for (int w = words - 1; w >= 0; w--) {
      for (int bit = 31; bit >= 0; bit--) {
        if (started) acc = acc.sqr();
        if ((e[w] >> bit) & 1u) {
          acc = started ? acc * (*this) : *this;
          started = true;
        }
      }
    }
    return acc;
  }
  PB_HD Field pow_u64(uint64_t e) const {
    uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
    return pow(w, 2);
  }
  // Fermat inverse (0 -> 0).
  PB_HD Field inv() const {
    uint32_t e[N];
    e[0] = sub_cc(P::MOD(0), 2u);
#pragma unroll
    for (int i = 1; i < N; i++) e[i] = subc_cc(P::MOD(i), 0u);
    return pow(e, N);
  }
};

// Constants are exposed through constexpr accessor functions with function-local tables so that
// they are usable from device code and fold to immediates once the limb loops are unrolled.
#define PB_LIMB_TABLE(name, n, ...)                     \
  static PB_HD constexpr uint32_t name(int i) {         \
    constexpr uint32_t T[n] = {__VA_ARGS__};            \
    return T[i];                                        \
  }

// Fr's Montgomery constant is -1 mod 2^32.  If ptxas can see that, it rewrites m = -t0 and splits
// every IMAD.WIDE of the reduction rows into IMAD.X + IMAD.HI.X (2x the issue slots, measured in
// SASS).  Reading the constant from __constant__ memory keeps it opaque to the optimiser.
#if defined(__CUDACC__)
static __constant__ uint32_t c_fr_inv = 0xffffffffu;
#endif

struct FrParams {
  static constexpr int N = 8;
  static PB_HD uint32_t inv() {
#if defined(__CUDA_ARCH__)
    return c_fr_inv;
#else
    return 0xffffffffu;
#endif
  }
  PB_LIMB_TABLE(MOD, 8, 0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u,
                0x299d7d48u, 0x73eda753u)
  PB_LIMB_TABLE(R1, 8, 0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu,
                0xacc5056fu, 0x1824b159u)
  PB_LIMB_TABLE(R2, 8, 0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u,
                0x9f59ff11u, 0x0748d9d9u)
};

struct FpParams {
  static constexpr int N = 12;
  static PB_HD uint32_t inv() { return 0xfffcfffdu; }
  PB_LIMB_TABLE(MOD, 12, 0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau)
  PB_LIMB_TABLE(R1, 12, 0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u)
  PB_LIMB_TABLE(R2, 12, 0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u)
};

typedef Field<FrParams> Fr;
typedef Field<FpParams> Fp;

}  // namespace pb
