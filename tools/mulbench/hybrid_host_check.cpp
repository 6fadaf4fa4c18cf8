// Host validation of the two-pipe Fp product of csrc/bigint.cuh (Field::mul_hybrid / sqr_hybrid /
// mul2_hybrid: IMAD product + DFMA Montgomery reduction) against the all-integer forms, which
// tests/hosttest pins against the big-integer oracle.
//   g++ -std=c++17 -O1 -frounding-math -mfma -o hybrid_host_check hybrid_host_check.cpp && ./hybrid_host_check
#include <cfenv>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../plonk_b200/csrc/bigint.cuh"
#include "fp_dfma52.cuh"
using pb::Fp;

static Fp reduce_once(Fp x) {  // bring an arbitrary 384-bit pattern below p (patterns < 2^384 < 10p)
  for (int k = 0; k < 10; k++) {
    uint32_t t[12];
    for (int i = 0; i < 12; i++) t[i] = x.v[i];
    Fp::final_sub(x.v, t, 0u);
  }
  return x;
}

int main(int argc, char** argv) {
  std::fesetround(FE_TOWARDZERO);
  long n = argc > 1 ? atol(argv[1]) : 200000;
  std::mt19937_64 g(0xB200);
  long bad = 0;
  auto check = [&](const Fp& a, const Fp& b) {
    Fp r0 = Fp::mul_imad(a, b), r1 = Fp::mul_hybrid(a, b);
    if (a == b) r1 = a.sqr_hybrid(), r0 = a.sqr_half();
    Fp m0 = Fp::mul2_imad(a, b, b, r0), m1 = Fp::mul2_hybrid(a, b, b, r0);
    // all-FP64 product on 52-bit limbs (R = 2^416): a b 2^-416 = (a b 2^-384) 2^352 2^-384
    Fp k352 = Fp::zero();
    k352.v[11] = 1;
    Fp e52 = Fp::mul_imad(Fp::mul_imad(a, b), k352), g52;
    pb52::to_words(pb52::mul52(pb52::from_words(a.v), pb52::from_words(b.v)), g52.v);
    if (e52 != g52) r1 = Fp::zero() - Fp::one();  // force a report below
    if (r0 != r1 || m0 != m1) {
      if (bad < 5) {
        printf("MISMATCH\n a=");
        for (int i = 11; i >= 0; i--) printf("%08x", a.v[i]);
        printf("\n b=");
        for (int i = 11; i >= 0; i--) printf("%08x", b.v[i]);
        printf("\n want=");
        for (int i = 11; i >= 0; i--) printf("%08x", r0.v[i]);
        printf("\n got =");
        for (int i = 11; i >= 0; i--) printf("%08x", r1.v[i]);
        printf("\n");
      }
      bad++;
    }
  };
  Fp pm1 = Fp::zero() - Fp::one().from_mont();  // p - 1
  Fp specials[6] = {Fp::zero(), Fp::one(), Fp::one().from_mont(), pm1, Fp::r2(), pm1 - Fp::one().from_mont()};
  for (auto& a : specials)
    for (auto& b : specials) check(a, b);
  for (long it = 0; it < n; it++) {
    Fp a, b;
    for (int i = 0; i < 12; i++) { a.v[i] = (uint32_t)g(); b.v[i] = (uint32_t)g(); }
    int mode = it & 7;
    if (mode == 1) for (int i = 0; i < 12; i++) a.v[i] = 0xffffffffu;      // saturated limbs
    if (mode == 2) for (int i = 0; i < 12; i++) b.v[i] = (g() & 1) ? 0xffffffffu : 0u;
    if (mode == 3) for (int i = 0; i < 12; i++) a.v[i] &= (uint32_t)g() & (uint32_t)g();  // sparse
    a = reduce_once(a); b = reduce_once(b);
    check(a, b);
    check(a, a);
  }
  printf("%ld products checked, %ld mismatches\n", 2 * n + 36, bad);
  return bad != 0;
}
