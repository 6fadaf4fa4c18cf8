// Host build (g++) of the device arithmetic headers, for CPU-side validation of the multi-limb
// algorithms against the Python oracle.  TEST INFRASTRUCTURE: the PTX carry chains are emulated by
// bigint.cuh's host primitives; nothing here is reachable from the product library.
#include <fenv.h>
#include <stddef.h>
#include <string.h>

#include "../../plonk_b200/csrc/ecntt.cuh"

using namespace pb;

extern "C" {

// op: 0 mul, 1 add, 2 sub, 3 inv(a), 4 to_mont(a), 5 from_mont(a)
int ht_fr_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fr x, y, r;
    memcpy(x.v, a + 8 * i, 32);
    if (b) memcpy(y.v, b + 8 * i, 32);
    switch (op) {
      case 0: r = x * y; break;
      case 1: r = x + y; break;
      case 2: r = x - y; break;
      case 3: r = x.inv(); break;
      case 4: r = x.to_mont(); break;
      case 5: r = x.from_mont(); break;
      case 6: r = x.sqr(); break;
      default: return -1;
    }
    memcpy(out + 8 * i, r.v, 32);
  }
  return 0;
}

int ht_fp_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fp x, y, r;
    memcpy(x.v, a + 12 * i, 48);
    if (b) memcpy(y.v, b + 12 * i, 48);
    switch (op) {
      case 0: r = x * y; break;
      case 1: r = x + y; break;
      case 2: r = x - y; break;
      case 3: r = x.inv(); break;
      case 4: r = x.to_mont(); break;
      case 5: r = x.from_mont(); break;
      case 6: r = x.sqr(); break;
      default: return -1;
    }
    memcpy(out + 12 * i, r.v, 48);
  }
  return 0;
}

// out = a*b + c*d (sub = 0) or a*b - c*d (sub = 1), one Montgomery reduction (Field::mul2)
int ht_fp_mul2(int sub, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fp x, y, z, w;
    memcpy(x.v, a + 12 * i, 48);
    memcpy(y.v, b + 12 * i, 48);
    memcpy(z.v, c + 12 * i, 48);
    memcpy(w.v, d + 12 * i, 48);
    const Fp r = sub ? Fp::mul_sub(x, y, z, w) : Fp::mul2(x, y, z, w);
    memcpy(out + 12 * i, r.v, 48);
  }
  return 0;
}

// The two-pipe Fp product (Field::mul_hybrid / sqr_hybrid / mul2_hybrid: integer product + DFMA
// Montgomery reduction).  The device rounds toward zero per instruction (DFMA.RZ); the host emulation
// needs the rounding mode set around the calls (and -frounding-math -mfma when compiling).
// op: 0 a*b, 1 a^2, 2 a*b + c*d
int ht_fp_hybrid(int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* out, size_t n) {
  const int old = fegetround();
  fesetround(FE_TOWARDZERO);
  for (size_t i = 0; i < n; i++) {
    Fp x, y, z, w, r;
    memcpy(x.v, a + 12 * i, 48);
    if (b) memcpy(y.v, b + 12 * i, 48);
    if (c) memcpy(z.v, c + 12 * i, 48);
    if (d) memcpy(w.v, d + 12 * i, 48);
    r = op == 0 ? Fp::mul_hybrid(x, y) : (op == 1 ? x.sqr_hybrid() : Fp::mul2_hybrid(x, y, z, w));
    memcpy(out + 12 * i, r.v, 48);
  }
  fesetround(old);
  return 0;
}

// The same for Fr (a*b only: 256 = 5 x 48 + 16, the last reduction step is 16 bits wide).
int ht_fr_hybrid(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  const int old = fegetround();
  fesetround(FE_TOWARDZERO);
  for (size_t i = 0; i < n; i++) {
    Fr x, y;
    memcpy(x.v, a + 8 * i, 32);
    memcpy(y.v, b + 8 * i, 32);
    const Fr r = Fr::mul_hybrid(x, y);
    memcpy(out + 8 * i, r.v, 32);
  }
  fesetround(old);
  return 0;
}

// Sum of n affine points (raw 96-byte layout) through xyzz_madd, with signs (1 = negate);
// then chain-adds the partial sums pairwise through xyzz_add to also exercise the full adder.
int ht_g1_sum(const uint32_t* pts, const uint8_t* neg, size_t n, uint32_t* out_affine) {
  G1Xyzz acc0 = G1Xyzz::identity(), acc1 = G1Xyzz::identity();
  for (size_t i = 0; i < n; i++) {
    G1Affine p;
    memcpy(&p, pts + 24 * i, 96);
    if (p.is_inf()) continue;
    Fp y = (neg && neg[i]) ? p.y.neg() : p.y;
    xyzz_madd((i & 1) ? acc1 : acc0, p.x, y);
  }
  xyzz_add(acc0, acc1);
  G1Affine r = xyzz_to_affine(acc0);
  memcpy(out_affine, &r, 96);
  return 0;
}

// k * P by double-and-add over XYZZ (exercises xyzz_dbl and xyzz_add).
int ht_g1_mul_small(const uint32_t* pt, uint64_t k, uint32_t* out_affine) {
  G1Affine p;
  memcpy(&p, pt, 96);
  G1Xyzz base = G1Xyzz::from_affine(p), acc = G1Xyzz::identity();
  for (int bit = 63; bit >= 0; bit--) {
    acc = xyzz_dbl(acc);
    if ((k >> bit) & 1) xyzz_add(acc, base);
  }
  G1Affine r = xyzz_to_affine(acc);
  memcpy(out_affine, &r, 96);
  return 0;
}

// Inverse NTT over n = 2^log_n affine G1 points (csrc/ecntt.cuh): out[j] = (1/n) sum_i w^(-ij) pts[i].
// The stages, twiddle exponents and the bit-reversed read-out are the ones of the CUDA driver
// (csrc/ecntt.cu); the twiddles are recomputed here from GENERATOR = 7.
int ht_ec_intt(const uint32_t* pts, int log_n, uint32_t* out) {
  const size_t n = (size_t)1 << log_n;
  G1Xyzz* A = new G1Xyzz[n];
  for (size_t i = 0; i < n; i++) {
    G1Affine p;
    memcpy(&p, pts + 24 * i, 96);
    A[i] = G1Xyzz::from_affine(p);
  }
  // ROOT_OF_UNITY = 7^((r-1)/2^32); w_n = ROOT^(2^(32-log_n))
  Fr seven = Fr::zero();
  seven.v[0] = 7;
  seven = seven.to_mont();
  uint32_t e[8];
  for (int i = 0; i < 8; i++) e[i] = FrParams::MOD(i);
  e[0] -= 1;
  uint32_t es[7];
  for (int i = 0; i < 7; i++) es[i] = e[i + 1];  // (r - 1) >> 32
  Fr w = seven.pow(es, 7);
  for (int i = 0; i < 32 - log_n; i++) w = w.sqr();
  const Fr w_inv = w.inv();
  Fr* tw = new Fr[n / 2 + 1];
  tw[0] = Fr::one();
  for (size_t k = 1; k < n / 2; k++) tw[k] = tw[k - 1] * w_inv;
  for (int s = 0; s < log_n; s++) {
    const size_t len = n >> s, half = len >> 1;
    for (size_t t = 0; t < n / 2; t++) {
      const size_t blk = t / half, j = t % half, i0 = blk * len + j, ex = j << s;
      ec_butterfly(A[i0], A[i0 + half], tw[ex], ex == 0);
    }
  }
  Fr nn = Fr::zero();
  nn.v[0] = (uint32_t)n;
  const Fr n_inv = nn.to_mont().inv().from_mont();
  for (size_t i = 0; i < n; i++) {
    const G1Affine r = xyzz_to_affine(xyzz_mul(A[i], n_inv.v, 8));
    memcpy(out + 24 * ec_bitrev((unsigned)i, log_n), &r, 96);
  }
  delete[] A;
  delete[] tw;
  return 0;
}
}
