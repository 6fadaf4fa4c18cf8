// CPU restatement (C++17 + OpenMP) of the dusk-plonk prover hot path and the glue around it.
//
// TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Nothing under plonk_b200/ links, loads or executes this
// file: it is built into oracle/_build/libcref.so and used by tests/ (as the checker at sizes the
// pure-Python oracle cannot reach), by __graft_entry__.smoke() and by bench.py's cpu_baseline /
// --impl reference legs.  Parity status: PINNED through oracle/pyref.py - this restatement must
// reproduce the reference's golden digest (src/compiler/prover.rs:1151-1158) itself
// (tests/test_cref.py) and equal pyref on shared vectors.
//
// It follows the reference's own schedules so that timing it is a fair stand-in for the rayon
// prover (which cannot be built here - no Rust toolchain):
//   * NTT: best_fft / serial_fft (src/fft/domain.rs:383-463): bit-reversal, log n DIT stages with
//     running-product twiddles, chunks in parallel when there are >= 4 of them, otherwise the
//     chunk split into per-thread ranges seeded by pow (parallel_butterfly_chunk :492-516).
//   * MSM: window-parallel Pippenger as published for dusk-bls12_381 0.14 msm_variable_base
//     (c = ln n + 2, 2^c - 1 buckets per window, running-sum reduction, Horner with doublings).
//   * Prover: Prover::prove_inner (src/compiler/prover.rs:415-761) with the same 4-way / 5-way
//     outer concurrency (rayon::join sites :163-210, quotient_poly.rs:139-157).
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------------------
// Montgomery fields, 64-bit limbs
// ---------------------------------------------------------------------------------------------
template <int N>
struct Ctx {
  uint64_t p[N];
  uint64_t inv;
  uint64_t r1[N];
  uint64_t r2[N];
};

static const Ctx<4> FR = {
    {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull},
    0xfffffffeffffffffull,
    {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full},
    {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull}};
static const Ctx<6> FP = {
    {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull,
     0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull},
    0x89f3fffcfffcfffdull,
    {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull,
     0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull},
    {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull, 0x67eb88a9939d83c0ull,
     0x9a793e85b519952dull, 0x11988fe592cae3aaull}};

template <int N, const Ctx<N>& C>
struct F {
  uint64_t l[N];
  static F zero() { F r; memset(r.l, 0, sizeof r.l); return r; }
  static F one() { F r; memcpy(r.l, C.r1, sizeof r.l); return r; }
  bool is_zero() const { uint64_t x = 0; for (int i = 0; i < N; i++) x |= l[i]; return !x; }
  bool operator==(const F& o) const { return !memcmp(l, o.l, sizeof l); }
  bool operator!=(const F& o) const { return !(*this == o); }
  static bool ge_p(const uint64_t* t) {
    for (int i = N - 1; i >= 0; i--) {
      if (t[i] != C.p[i]) return t[i] > C.p[i];
    }
    return true;
  }
  static void sub_p(uint64_t* t) {
    u128 b = 0;
    for (int i = 0; i < N; i++) {
      u128 d = (u128)t[i] - C.p[i] - b;
      t[i] = (uint64_t)d;
      b = (d >> 64) & 1;
    }
  }
  F operator+(const F& o) const {
    F r;
    u128 c = 0;
    for (int i = 0; i < N; i++) { c += (u128)l[i] + o.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (c || ge_p(r.l)) sub_p(r.l);
    return r;
  }
  F operator-(const F& o) const {
    F r;
    u128 b = 0;
    for (int i = 0; i < N; i++) { u128 d = (u128)l[i] - o.l[i] - b; r.l[i] = (uint64_t)d; b = (d >> 64) & 1; }
    if (b) {
      u128 c = 0;
      for (int i = 0; i < N; i++) { c += (u128)r.l[i] + C.p[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    }
    return r;
  }
  F neg() const { return zero() - *this; }
  F dbl() const { return *this + *this; }
  F operator*(const F& o) const {
    uint64_t t[N + 2];
    memset(t, 0, sizeof t);
    for (int i = 0; i < N; i++) {
      u128 c = 0;
      for (int j = 0; j < N; j++) { c += (u128)l[j] * o.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
      c += t[N]; t[N] = (uint64_t)c; t[N + 1] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * C.inv;
      c = ((u128)m * C.p[0] + t[0]) >> 64;
      for (int j = 1; j < N; j++) { c += (u128)m * C.p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
      c += t[N]; t[N - 1] = (uint64_t)c; t[N] = t[N + 1] + (uint64_t)(c >> 64);
    }
    F r;
    memcpy(r.l, t, sizeof r.l);
    if (t[N] || ge_p(r.l)) sub_p(r.l);
    return r;
  }
  F sqr() const { return *this * *this; }
  F pow(const uint64_t* e, int words) const {
    F acc = one();
    for (int w = words - 1; w >= 0; w--)
      for (int b = 63; b >= 0; b--) { acc = acc.sqr(); if ((e[w] >> b) & 1) acc = acc * *this; }
    return acc;
  }
  F pow64(uint64_t e) const { return pow(&e, 1); }
  F inv() const { uint64_t e[N]; memcpy(e, C.p, sizeof e); e[0] -= 2; return pow(e, N); }
  static F from_u64(uint64_t x) { F r = zero(); r.l[0] = x; F r2; memcpy(r2.l, C.r2, sizeof r2.l); return r * r2; }
  F from_mont() const { F o = zero(); o.l[0] = 1; return *this * o; }
  F to_mont() const { F r2; memcpy(r2.l, C.r2, sizeof r2.l); return *this * r2; }
};
typedef F<4, FR> Fr;
typedef F<6, FP> Fp;

// ---------------------------------------------------------------------------------------------
// G1, Jacobian coordinates (a = 0)
// ---------------------------------------------------------------------------------------------
struct Aff { Fp x, y; bool inf() const { return x.is_zero() && y.is_zero(); } };
struct Jac { Fp x, y, z; bool inf() const { return z.is_zero(); } };
static Jac jac_id() { Jac r; r.x = Fp::one(); r.y = Fp::one(); r.z = Fp::zero(); return r; }

static Jac jac_dbl(const Jac& p) {
  if (p.inf() || p.y.is_zero()) return jac_id();
  Fp a = p.x.sqr(), b = p.y.sqr(), c = b.sqr();
  Fp d = ((p.x + b).sqr() - a - c).dbl();
  Fp e = a.dbl() + a, f = e.sqr();
  Jac r;
  r.x = f - d.dbl();
  r.y = e * (d - r.x) - c.dbl().dbl().dbl();
  r.z = (p.y * p.z).dbl();
  return r;
}
static Jac jac_add_mixed(const Jac& p, const Aff& q) {
  if (q.inf()) return p;
  if (p.inf()) { Jac r; r.x = q.x; r.y = q.y; r.z = Fp::one(); return r; }
  Fp z1z1 = p.z.sqr(), u2 = q.x * z1z1, s2 = q.y * p.z * z1z1;
  if (u2 == p.x) {
    if (s2 == p.y) return jac_dbl(p);
    return jac_id();
  }
  Fp h = u2 - p.x, hh = h.sqr(), i = hh.dbl().dbl(), j = h * i, rr = (s2 - p.y).dbl(), v = p.x * i;
  Jac r;
  r.x = rr.sqr() - j - v.dbl();
  r.y = rr * (v - r.x) - (p.y * j).dbl();
  r.z = (p.z + h).sqr() - z1z1 - hh;
  return r;
}
static Jac jac_add(const Jac& p, const Jac& q) {
  if (p.inf()) return q;
  if (q.inf()) return p;
  Fp z1z1 = p.z.sqr(), z2z2 = q.z.sqr();
  Fp u1 = p.x * z2z2, u2 = q.x * z1z1, s1 = p.y * q.z * z2z2, s2 = q.y * p.z * z1z1;
  if (u1 == u2) {
    if (s1 == s2) return jac_dbl(p);
    return jac_id();
  }
  Fp h = u2 - u1, i = h.dbl().sqr(), j = h * i, rr = (s2 - s1).dbl(), v = u1 * i;
  Jac r;
  r.x = rr.sqr() - j - v.dbl();
  r.y = rr * (v - r.x) - (s1 * j).dbl();
  r.z = ((p.z + q.z).sqr() - z1z1 - z2z2) * h;
  return r;
}
static Aff jac_to_aff(const Jac& p) {
  Aff a;
  if (p.inf()) { a.x = Fp::zero(); a.y = Fp::zero(); return a; }
  Fp zi = p.z.inv(), zi2 = zi.sqr();
  a.x = p.x * zi2;
  a.y = p.y * zi2 * zi;
  return a;
}
static Jac jac_mul(const Jac& p, const Fr& k_mont) {
  Fr k = k_mont.from_mont();
  Jac acc = jac_id();
  for (int w = 3; w >= 0; w--)
    for (int b = 63; b >= 0; b--) { acc = jac_dbl(acc); if ((k.l[w] >> b) & 1) acc = jac_add(acc, p); }
  return acc;
}
static void g1_compress(const Aff& a, uint8_t out[48]) {
  if (a.inf()) { memset(out, 0, 48); out[0] = 0xC0; return; }
  Fp xc = a.x.from_mont(), yc = a.y.from_mont(), ny = a.y.neg().from_mont();
  for (int i = 0; i < 6; i++) for (int b = 0; b < 8; b++) out[47 - (8 * i + b)] = (uint8_t)(xc.l[i] >> (8 * b));
  out[0] |= 0x80;
  bool gt = false;  // y > p - y ?
  for (int i = 5; i >= 0; i--) if (yc.l[i] != ny.l[i]) { gt = yc.l[i] > ny.l[i]; break; }
  if (gt) out[0] |= 0x20;
}

// ---------------------------------------------------------------------------------------------
// EvaluationDomain (src/fft/domain.rs)
// ---------------------------------------------------------------------------------------------
static const uint64_t ROOT_OF_UNITY_CANON[4] = {0x3829971f439f0d2bull, 0xb63683508c2280b9ull, 0xd09b681922c813b4ull, 0x16a2a19edfe81f20ull};

struct Domain {
  uint64_t size; uint32_t log; Fr gen, gen_inv, size_inv, coset_gen, coset_gen_inv;
  explicit Domain(uint64_t num_coeffs) {
    size = 1; log = 0;
    while (size < num_coeffs) { size <<= 1; log++; }
    Fr g; memcpy(g.l, ROOT_OF_UNITY_CANON, 32); g = g.to_mont();
    for (uint32_t i = log; i < 32; i++) g = g.sqr();
    gen = g; gen_inv = g.inv(); size_inv = Fr::from_u64(size).inv();
    coset_gen = Fr::from_u64(7); coset_gen_inv = coset_gen.inv();
  }
};

static inline uint32_t bitrev32(uint32_t n, uint32_t l) { uint32_t r = 0; for (uint32_t i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; } return r; }

static void butterfly_range(Fr* left, Fr* right, size_t len, const Fr& w_m, Fr w) {  // domain.rs:472-489
  for (size_t i = 0; i < len; i++) {
    Fr t = right[i] * w;
    right[i] = left[i] - t;
    left[i] = left[i] + t;
    w = w * w_m;
  }
}

// best_fft (domain.rs:383-422) with `threads` workers.
static void best_fft(Fr* a, size_t n, const Fr& omega, uint32_t log_n, int threads) {
  for (uint32_t k = 0; k < n; k++) { uint32_t rk = bitrev32(k, log_n); if (k < rk) std::swap(a[k], a[rk]); }
  const bool par = n >= (1u << 12) && threads > 1;
  size_t m = 1;
  for (uint32_t s = 0; s < log_n; s++) {
    Fr w_m = omega.pow64(n / (2 * m));
    size_t chunk_len = 2 * m, chunk_count = n / chunk_len;
    if (!par) {
      for (size_t c = 0; c < chunk_count; c++) butterfly_range(a + c * chunk_len, a + c * chunk_len + m, m, w_m, Fr::one());
    } else if (chunk_count >= 4) {
#pragma omp parallel for num_threads(threads) schedule(static)
      for (size_t c = 0; c < chunk_count; c++) butterfly_range(a + c * chunk_len, a + c * chunk_len + m, m, w_m, Fr::one());
    } else {
      for (size_t c = 0; c < chunk_count; c++) {  // parallel_butterfly_chunk (domain.rs:492-516)
        Fr* left = a + c * chunk_len; Fr* right = left + m;
        size_t range_len = (m + threads - 1) / threads, range_count = (m + range_len - 1) / range_len;
        Fr seed_step = w_m.pow64(range_len);
        std::vector<Fr> seeds(range_count);
        Fr sd = Fr::one();
        for (size_t r = 0; r < range_count; r++) { seeds[r] = sd; sd = sd * seed_step; }
#pragma omp parallel for num_threads(threads) schedule(static)
        for (size_t r = 0; r < range_count; r++) {
          size_t lo = r * range_len, len = std::min(range_len, m - lo);
          butterfly_range(left + lo, right + lo, len, w_m, seeds[r]);
        }
      }
    }
    m *= 2;
  }
}

static std::vector<Fr> dom_fft(const Domain& d, const Fr* in, size_t in_len, bool inverse, bool coset, int threads) {
  std::vector<Fr> a(d.size, Fr::zero());
  size_t use = std::min<size_t>(in_len, d.size);
  if (coset && !inverse) {  // distribute_powers before the transform (domain.rs:198-218); serial, as in the reference
    Fr p = Fr::one();
    for (size_t i = 0; i < in_len; i++) { if (i < use) a[i] = in[i] * p; p = p * d.coset_gen; }
  } else {
    for (size_t i = 0; i < use; i++) a[i] = in[i];
  }
  best_fft(a.data(), d.size, inverse ? d.gen_inv : d.gen, d.log, threads);
  if (inverse) {
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < d.size; i++) a[i] = a[i] * d.size_inv;
    if (coset) { Fr p = Fr::one(); for (size_t i = 0; i < d.size; i++) { a[i] = a[i] * p; p = p * d.coset_gen_inv; } }
  }
  return a;
}

// ---------------------------------------------------------------------------------------------
// MSM (msm_variable_base) and commit
// ---------------------------------------------------------------------------------------------
static Jac msm(const Aff* bases, const Fr* scalars_mont, size_t n, int threads) {
  if (n == 0) return jac_id();
  int c = n < 32 ? 3 : (int)log((double)n) + 2;
  std::vector<Fr> sc(n);
#pragma omp parallel for num_threads(threads) schedule(static)
  for (size_t i = 0; i < n; i++) sc[i] = scalars_mont[i].from_mont();
  int n_win = (255 + c - 1) / c;
  std::vector<Jac> wsum(n_win);
  Fr one_c = Fr::zero(); one_c.l[0] = 1;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int w = 0; w < n_win; w++) {
    int w_start = w * c;
    std::vector<Jac> buckets((size_t(1) << c) - 1, jac_id());
    Jac res = jac_id();
    for (size_t i = 0; i < n; i++) {
      const Fr& s = sc[i];
      if (s.is_zero()) continue;
      if (s == one_c) { if (w_start == 0) res = jac_add_mixed(res, bases[i]); continue; }
      int word = w_start >> 6, off = w_start & 63;
      uint64_t d = s.l[word] >> off;
      if (off && word + 1 < 4) d |= s.l[word + 1] << (64 - off);
      d &= (1ull << c) - 1;
      if (d) buckets[d - 1] = jac_add_mixed(buckets[d - 1], bases[i]);
    }
    Jac running = jac_id();
    for (size_t b = buckets.size(); b-- > 0;) { running = jac_add(running, buckets[b]); res = jac_add(res, running); }
    wsum[w] = res;
  }
  Jac total = wsum[n_win - 1];
  for (int w = n_win - 2; w >= 0; w--) { for (int k = 0; k < c; k++) total = jac_dbl(total); total = jac_add(total, wsum[w]); }
  return total;
}

static size_t trimmed_len(const Fr* p, size_t n) { while (n && p[n - 1].is_zero()) n--; return n; }

// CommitKey::commit (key.rs:376-388); returns false for PolynomialDegreeTooLarge
static bool commit(const std::vector<Aff>& ck, const Fr* poly, size_t len, Aff* out, int threads) {
  len = trimmed_len(poly, len);
  size_t degree = len ? len - 1 : 0;
  if (degree > ck.size() - 1) return false;
  *out = jac_to_aff(msm(ck.data(), poly, len, threads));
  return true;
}

// ---------------------------------------------------------------------------------------------
// merlin transcript (STROBE-128 over Keccak-f[1600])
// ---------------------------------------------------------------------------------------------
static void keccak_f(uint64_t st[25]) {
  static const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull, 0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  for (int rnd = 0; rnd < 24; rnd++) {
    uint64_t C[5], D[5], B[25];
    for (int x = 0; x < 5; x++) C[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
    for (int x = 0; x < 5; x++) { uint64_t c1 = C[(x + 1) % 5]; D[x] = C[(x + 4) % 5] ^ ((c1 << 1) | (c1 >> 63)); }
    for (int i = 0; i < 25; i++) st[i] ^= D[i % 5];
    for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) {
      int i = x + 5 * y; uint64_t v = st[i]; int r = ROT[i];
      B[y + 5 * ((2 * x + 3 * y) % 5)] = r ? ((v << r) | (v >> (64 - r))) : v;
    }
    for (int y = 0; y < 5; y++) for (int x = 0; x < 5; x++) st[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
    st[0] ^= RC[rnd];
  }
}
struct Strobe {
  uint8_t st[200]; int pos, pos_begin, cur_flags;
  enum { R = 166, FI = 1, FA = 2, FC = 4, FT = 8, FM = 16, FK = 32 };
  void run_f() { st[pos] ^= (uint8_t)pos_begin; st[pos + 1] ^= 0x04; st[R + 1] ^= 0x80; uint64_t w[25]; memcpy(w, st, 200); keccak_f(w); memcpy(st, w, 200); pos = pos_begin = 0; }
  void absorb(const uint8_t* d, size_t n) { for (size_t i = 0; i < n; i++) { st[pos++] ^= d[i]; if (pos == R) run_f(); } }
  void squeeze(uint8_t* d, size_t n) { for (size_t i = 0; i < n; i++) { d[i] = st[pos]; st[pos++] = 0; if (pos == R) run_f(); } }
  void begin_op(int flags, bool more) {
    if (more) return;
    uint8_t hdr[2] = {(uint8_t)pos_begin, (uint8_t)flags};
    pos_begin = pos + 1; cur_flags = flags;
    absorb(hdr, 2);
    if ((flags & (FC | FK)) && pos != 0) run_f();
  }
  void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(FM | FA, more); absorb(d, n); }
  void ad(const uint8_t* d, size_t n, bool more) { begin_op(FA, more); absorb(d, n); }
  void prf(uint8_t* d, size_t n) { begin_op(FI | FA | FC, false); squeeze(d, n); }
  explicit Strobe(const char* proto) {
    memset(st, 0, 200);
    const uint8_t hdr[6] = {1, R + 2, 1, 0, 1, 96};
    memcpy(st, hdr, 6); memcpy(st + 6, "STROBEv1.0.2", 12);
    uint64_t w[25]; memcpy(w, st, 200); keccak_f(w); memcpy(st, w, 200);
    pos = pos_begin = cur_flags = 0;
    meta_ad((const uint8_t*)proto, strlen(proto), false);
  }
};
struct Transcript {
  Strobe s;
  Transcript(const uint8_t* label, size_t n) : s("Merlin v1.0") { append("dom-sep", label, n); }
  void append(const char* label, const uint8_t* m, size_t n) {
    uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    s.meta_ad((const uint8_t*)label, strlen(label), false); s.meta_ad(len, 4, true); s.ad(m, n, false);
  }
  void append_u64(const char* label, uint64_t x) { uint8_t b[8]; for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i)); append(label, b, 8); }
  void append_commitment(const char* label, const Aff& a) { uint8_t b[48]; g1_compress(a, b); append(label, b, 48); }
  void append_scalar(const char* label, const Fr& x) { Fr c = x.from_mont(); append(label, (const uint8_t*)c.l, 32); }
  Fr challenge_scalar(const char* label) {
    uint8_t len[4] = {64, 0, 0, 0}, buf[64];
    s.meta_ad((const uint8_t*)label, strlen(label), false); s.meta_ad(len, 4, true); s.prf(buf, 64);
    // from_bytes_wide: (lo + hi * 2^256) mod r, via Montgomery: lo*R2*R^-1... use d0*R2 + d1*R3
    Fr lo, hi; memcpy(lo.l, buf, 32); memcpy(hi.l, buf + 32, 32);
    Fr r2; memcpy(r2.l, FR.r2, 32);
    Fr r3 = r2 * r2;            // R^3 in Montgomery-of-integers sense: mont(R2,R2) = R^3
    return lo * r2 + hi * r3;   // mont(lo, R^2) = lo*R ; mont(hi, R^3) = hi*R^2 = (hi*2^256)*R
  }
  void circuit_domain_sep(uint64_t n) { append("dom-sep", (const uint8_t*)"circuit_size", 12); append_u64("n", n); }
};

// ---------------------------------------------------------------------------------------------
// Prover (compile + prove)
// ---------------------------------------------------------------------------------------------
enum { Q_M, Q_L, Q_R, Q_O, Q_F, Q_C, Q_ARITH, Q_RANGE, Q_LOGIC, Q_FIXED, Q_VAR, S1, S2, S3, S4, N_POLY };
static const Fr& K(int i) { static Fr k[4] = {Fr::from_u64(1), Fr::from_u64(7), Fr::from_u64(13), Fr::from_u64(17)}; return k[i]; }

struct Prover {
  std::vector<uint8_t> label;
  size_t constraints, size;
  std::vector<Aff> ck;
  std::vector<Fr> polys[N_POLY];     // trimmed coefficient form
  std::vector<Fr> evals8[N_POLY];    // coset evaluations over 8n
  std::vector<Fr> linear8, vh8;
  Aff comms[N_POLY];
  std::vector<Fr> sigma_evals[4];
  Fr vh_inv[8];
  std::vector<uint32_t> wires[4];
  int threads;
};

static void trim(std::vector<Fr>& v) { while (!v.empty() && v.back().is_zero()) v.pop_back(); }
static Fr poly_eval(const std::vector<Fr>& p, const Fr& x) { Fr acc = Fr::zero(); for (size_t i = p.size(); i-- > 0;) acc = acc * x + p[i]; return acc; }
static Fr fr_small(int64_t v) { return v >= 0 ? Fr::from_u64((uint64_t)v) : Fr::from_u64((uint64_t)(-v)).neg(); }
static Fr EDWARDS_D() { static Fr d = (Fr::from_u64(10240) * Fr::from_u64(10241).inv()).neg(); return d; }

static Fr delta(const Fr& f) { Fr o = Fr::one(); Fr f1 = f - o, f2 = f1 - o, f3 = f2 - o; return f * f1 * f2 * f3; }
static Fr delta_xor_and(const Fr& a, const Fr& b, const Fr& w, const Fr& c, const Fr& q_c) {
  Fr nine = fr_small(9), two = fr_small(2), three = fr_small(3), four = fr_small(4), e18 = fr_small(18), e81 = fr_small(81), e83 = fr_small(83);
  Fr F_ = w * (w * (four * w - e18 * (a + b) + e81) + e18 * (a.sqr() + b.sqr()) - e81 * (a + b) + e83);
  Fr E = three * (a + b + c) - two * F_;
  Fr B = q_c * (nine * c - three * (a + b));
  return B + E;
}
struct Wire8 { Fr a, b, c, d, a_w, b_w, d_w; };
static Fr w_range(const Fr& ch, const Wire8& v) {
  Fr four = fr_small(4), k = ch.sqr(), k2 = k.sqr(), k3 = k2 * k;
  return (delta(v.c - four * v.d) + delta(v.b - four * v.c) * k + delta(v.a - four * v.b) * k2 + delta(v.d_w - four * v.a) * k3) * ch;
}
static Fr w_logic(const Fr& ch, const Fr& q_c, const Wire8& v) {
  Fr four = fr_small(4), k = ch.sqr(), k2 = k.sqr(), k3 = k2 * k, k4 = k3 * k;
  Fr A = v.a_w - four * v.a, B = v.b_w - four * v.b, D = v.d_w - four * v.d;
  return (delta(A) + delta(B) * k + delta(D) * k2 + (v.c - A * B) * k3 + delta_xor_and(A, B, v.c, D, q_c) * k4) * ch;
}
static Fr w_fixed(const Fr& ch, const Fr& q_l, const Fr& q_r, const Fr& q_c, const Wire8& v) {
  Fr one = Fr::one(), k = ch.sqr(), k2 = k.sqr(), k3 = k2 * k;
  Fr bit = v.d_w - v.d - v.d;
  Fr bit_c = bit * (bit - one) * (bit + one);
  Fr y_alpha = bit.sqr() * (q_r - one) + one, x_alpha = bit * q_l;
  Fr xy = (bit * q_c - v.c) * k;
  Fr t = v.c * v.a * v.b * EDWARDS_D();
  Fr xa = ((v.a_w + v.a_w * t) - (v.a * y_alpha + v.b * x_alpha)) * k2;
  Fr ya = ((v.b_w - v.b_w * t) - (v.b * y_alpha + v.a * x_alpha)) * k3;
  return (bit_c + xa + ya + xy) * ch;
}
static Fr w_var(const Fr& ch, const Wire8& v) {
  Fr k = ch.sqr();
  const Fr &x1 = v.a, &x3 = v.a_w, &y1 = v.b, &y3 = v.b_w, &x2 = v.c, &y2 = v.d, &x1y2 = v.d_w;
  Fr xy = x1 * y2 - x1y2, y1x2 = y1 * x2, y1y2 = y1 * y2, x1x2 = x1 * x2;
  Fr t = EDWARDS_D() * x1y2 * y1x2;
  Fr x3c = ((x1y2 + y1x2) - (x3 + x3 * t)) * k;
  Fr y3c = ((y1y2 + x1x2) - (y3 - y3 * t)) * k.sqr();
  return (xy + x3c + y3c) * ch;
}

static void fr_store(uint8_t* out, const Fr& x) { Fr c = x.from_mont(); memcpy(out, c.l, 32); }

extern "C" {

int cref_threads() { return omp_get_max_threads(); }

int cref_ntt(const uint64_t* in, size_t in_len, uint64_t* out, uint32_t log_n, int inverse, int coset, int threads) {
  Domain d((uint64_t)1 << log_n);
  std::vector<Fr> r = dom_fft(d, (const Fr*)in, in_len, inverse, coset, threads);
  memcpy(out, r.data(), r.size() * 32);
  return 0;
}

int cref_msm(const uint64_t* bases_raw, const uint64_t* scalars, size_t n, uint64_t* out_affine, int threads) {
  Aff a = jac_to_aff(msm((const Aff*)bases_raw, (const Fr*)scalars, n, threads));
  memcpy(out_affine, &a, 96);
  return 0;
}

// powers_of_g[i] = [x^i] ([g_scalar] G), i < n  (PublicParameters::setup, srs.rs:74-87)
int cref_srs_from_secret(size_t n, const uint64_t* x_mont, const uint64_t* g_scalar_mont, uint64_t* out_raw, int threads) {
  Aff gen;
  static const uint64_t GX[6] = {0xfb3af00adb22c6bbull, 0x6c55e83ff97a1aefull, 0xa14e3a3f171bac58ull, 0xc3688c4f9774b905ull, 0x2695638c4fa9ac0full, 0x17f1d3a73197d794ull};
  static const uint64_t GY[6] = {0x0caa232946c5e7e1ull, 0xd03cc744a2888ae4ull, 0x00db18cb2c04b3edull, 0xfcf5e095d5d00af6ull, 0xa09e30ed741d8ae4ull, 0x08b3f481e3aaa0f1ull};
  memcpy(gen.x.l, GX, 48); memcpy(gen.y.l, GY, 48); gen.x = gen.x.to_mont(); gen.y = gen.y.to_mont();
  Fr x, gs; memcpy(x.l, x_mont, 32); memcpy(gs.l, g_scalar_mont, 32);
  Jac gj; gj.x = gen.x; gj.y = gen.y; gj.z = Fp::one();
  Jac g = jac_mul(gj, gs);
  std::vector<Fr> pw(n);
  Fr p = Fr::one();
  for (size_t i = 0; i < n; i++) { pw[i] = p; p = p * x; }
  Aff* out = (Aff*)out_raw;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 64)
  for (size_t i = 0; i < n; i++) out[i] = jac_to_aff(jac_mul(g, pw[i]));
  return 0;
}

void* cref_prover_new(const uint8_t* label, size_t label_len, size_t constraints, const uint64_t* selectors /*11 x constraints*/,
                      const uint32_t* wires /*4 x constraints*/, size_t n_witnesses, const uint64_t* srs_raw, size_t n_srs, int threads) {
  Prover* P = new Prover();
  P->threads = threads;
  P->label.assign(label, label + label_len);
  P->constraints = constraints;
  size_t n_trim = 1; while (n_trim < constraints + 6) n_trim <<= 1;       // compiler.rs:121-124
  size_t keep = n_trim + 6; if (keep == 1) keep = 2;                      // key.rs:336-355
  if (keep > n_srs - 1) { delete P; return nullptr; }
  P->ck.assign((const Aff*)srs_raw, (const Aff*)srs_raw + keep + 1);
  size_t size = 1; while (size < constraints) size <<= 1;
  P->size = size;
  Domain dn(size), d8(8 * size);
  const Fr* sel = (const Fr*)selectors;
  for (int k = 0; k < 4; k++) P->wires[k].assign(wires + k * constraints, wires + (k + 1) * constraints);
  // selector polynomials (compiler.rs:149-211)
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int k = 0; k < 11; k++) {
    std::vector<Fr> col(size, Fr::zero());
    for (size_t i = 0; i < constraints; i++) col[i] = sel[k * constraints + i];
    P->polys[k] = dom_fft(dn, col.data(), size, true, false, 1);
    trim(P->polys[k]);
  }
  // sigma permutations (permutation.rs:106-211)
  std::vector<std::vector<uint64_t>> wmap(n_witnesses);
  for (size_t g = 0; g < constraints; g++) for (int k = 0; k < 4; k++) wmap[P->wires[k][g]].push_back(((uint64_t)k << 40) | g);
  std::vector<uint64_t> sig[4];
  for (int k = 0; k < 4; k++) { sig[k].resize(size); for (size_t i = 0; i < size; i++) sig[k][i] = ((uint64_t)k << 40) | i; }
  for (auto& lst : wmap) for (size_t i = 0; i < lst.size(); i++) { uint64_t cur = lst[i], nxt = lst[(i + 1) % lst.size()]; sig[cur >> 40][cur & 0xffffffffffull] = nxt; }
  std::vector<Fr> roots(size); { Fr w = Fr::one(); for (size_t i = 0; i < size; i++) { roots[i] = w; w = w * dn.gen; } }
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int k = 0; k < 4; k++) {
    std::vector<Fr> lag(size);
    for (size_t i = 0; i < size; i++) lag[i] = K(sig[k][i] >> 40) * roots[sig[k][i] & 0xffffffffffull];
    P->polys[S1 + k] = dom_fft(dn, lag.data(), size, true, false, 1);
    trim(P->polys[S1 + k]);
  }
  bool ok = true;
  for (int k = 0; k < N_POLY; k++) {
    Aff c;
    if (!commit(P->ck, P->polys[k].data(), P->polys[k].size(), &c, threads)) { if (k >= S1) ok = false; c.x = Fp::zero(); c.y = Fp::zero(); }
    P->comms[k] = c;
  }
  if (!ok) { delete P; return nullptr; }
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int k = 0; k < N_POLY; k++) P->evals8[k] = dom_fft(d8, P->polys[k].data(), P->polys[k].size(), false, true, 1);
  { Fr lin[2] = {Fr::zero(), Fr::one()}; P->linear8 = dom_fft(d8, lin, 2, false, true, threads); }
  { // compute_vanishing_poly_over_coset (domain.rs:340-351)
    P->vh8.resize(8 * size);
    Fr point = d8.coset_gen.pow64(size), step = d8.gen.pow64(size);
    for (size_t i = 0; i < 8 * size; i++) { P->vh8[i] = point - Fr::one(); point = point * step; }
    for (int i = 0; i < 8; i++) P->vh_inv[i] = P->vh8[i].inv();
  }
  for (int k = 0; k < 4; k++) P->sigma_evals[k] = dom_fft(dn, P->polys[S1 + k].data(), P->polys[S1 + k].size(), false, false, threads);
  return P;
}

void cref_prover_free(void* p) { delete (Prover*)p; }

// 15 commitments (48 bytes each) in SELECTORS + sigma order, for cross-checking the preprocessing.
void cref_prover_commitments(void* p, uint8_t* out) { Prover* P = (Prover*)p; for (int k = 0; k < N_POLY; k++) g1_compress(P->comms[k], out + 48 * k); }

// Prover::prove_inner, V3.  blinders: 14 Fr (Montgomery) in RNG order a0,a1,b0,b1,c0,c1,d0,d1, z0,z1,z2, b12,b13,b14.
// Returns 0, or -5 for Error::CircuitUnsatisfied.
int cref_prove(void* p, const uint64_t* witnesses, const uint64_t* pi_idx, const uint64_t* pi_vals, size_t n_pi,
               const uint64_t* blinders, uint8_t* out_proof) {
  Prover* P = (Prover*)p;
  const int T = P->threads;
  const size_t n = P->size, n8 = 8 * n;
  Domain dn(n), d8(n8);
  const Fr* W = (const Fr*)witnesses;
  const Fr* BL = (const Fr*)blinders;
  const Fr* PIV = (const Fr*)pi_vals;
  Transcript tr(P->label.data(), P->label.size());
  tr.circuit_domain_sep(P->constraints);
  static const char* LBL[N_POLY] = {"q_m", "q_l", "q_r", "q_o", "q_c", "q_f", "q_arith", "q_range", "q_logic", "q_variable_group_add", "q_fixed_group_add", "s_sigma_1", "s_sigma_2", "s_sigma_3", "s_sigma_4"};
  static const int ORD[N_POLY] = {Q_M, Q_L, Q_R, Q_O, Q_C, Q_F, Q_ARITH, Q_RANGE, Q_LOGIC, Q_VAR, Q_FIXED, S1, S2, S3, S4};
  for (int i = 0; i < N_POLY; i++) tr.append_commitment(LBL[i], P->comms[ORD[i]]);
  tr.circuit_domain_sep(P->constraints);
  std::vector<Fr> dense_pi(n, Fr::zero());
  for (size_t i = 0; i < n_pi; i++) { dense_pi[pi_idx[i]] = PIV[i]; tr.append_scalar("pi", PIV[i]); }

  // round 1
  std::vector<Fr> wv[4];
  for (int k = 0; k < 4; k++) { wv[k].assign(n, Fr::zero()); for (size_t i = 0; i < P->constraints; i++) wv[k][i] = W[P->wires[k][i]]; }
  std::vector<Fr> wp[4];
  auto blind = [&](const std::vector<Fr>& vals, const Fr* b, int nb, int th) {
    std::vector<Fr> c = dom_fft(dn, vals.data(), n, true, false, th);
    for (int i = 0; i < nb; i++) { c[i] = c[i] - b[i]; c.push_back(b[i]); }
    trim(c);
    return c;
  };
  const int T4 = std::max(1, T / 4);
  omp_set_max_active_levels(2);
#pragma omp parallel for num_threads(std::min(4, T)) schedule(static, 1)
  for (int k = 0; k < 4; k++) wp[k] = blind(wv[k], BL + 2 * k, 2, T4);
  Aff wc[4];
  bool ok = true;
#pragma omp parallel for num_threads(std::min(4, T)) schedule(static, 1)
  for (int k = 0; k < 4; k++) if (!commit(P->ck, wp[k].data(), wp[k].size(), &wc[k], T4)) ok = false;
  if (!ok) return -3;
  tr.append_commitment("a_comm", wc[0]); tr.append_commitment("b_comm", wc[1]); tr.append_commitment("c_comm", wc[2]); tr.append_commitment("d_comm", wc[3]);

  // round 2
  Fr beta = tr.challenge_scalar("beta"); tr.append_scalar("beta", beta);
  Fr gamma = tr.challenge_scalar("gamma");
  std::vector<Fr> perm(n);
  {
    std::vector<Fr> num(n), den(n), roots(n);
    { Fr w = Fr::one(); for (size_t i = 0; i < n; i++) { roots[i] = w; w = w * dn.gen; } }
#pragma omp parallel for num_threads(T) schedule(static)
    for (size_t i = 0; i < n; i++) {
      Fr br = beta * roots[i];
      num[i] = (wv[0][i] + br + gamma) * (wv[1][i] + br * K(1) + gamma) * (wv[2][i] + br * K(2) + gamma) * (wv[3][i] + br * K(3) + gamma);
      den[i] = (wv[0][i] + beta * P->sigma_evals[0][i] + gamma) * (wv[1][i] + beta * P->sigma_evals[1][i] + gamma) * (wv[2][i] + beta * P->sigma_evals[2][i] + gamma) * (wv[3][i] + beta * P->sigma_evals[3][i] + gamma);
    }
    // batch_inversion (util.rs:87-118)
    std::vector<Fr> prod(n); Fr tmp = Fr::one();
    for (size_t i = 0; i < n; i++) { tmp = tmp * den[i]; prod[i] = tmp; }
    tmp = tmp.inv();
    for (size_t i = n; i-- > 0;) { Fr s = i ? prod[i - 1] : Fr::one(); Fr nt = tmp * den[i]; den[i] = tmp * s; tmp = nt; }
    Fr product = Fr::one();
    for (size_t i = 0; i < n; i++) { perm[i] = product; if (i + 1 < n) product = product * (num[i] * den[i]); }
  }
  std::vector<Fr> zp = blind(perm, BL + 8, 3, T);
  Aff zc;
  if (!commit(P->ck, zp.data(), zp.size(), &zc, T)) return -3;
  tr.append_commitment("z_comm", zc);

  // round 3
  Fr alpha = tr.challenge_scalar("alpha");
  Fr ch_range = tr.challenge_scalar("range separation challenge");
  Fr ch_logic = tr.challenge_scalar("logic separation challenge");
  Fr ch_fixed = tr.challenge_scalar("fixed base separation challenge");
  Fr ch_var = tr.challenge_scalar("variable base separation challenge");
  std::vector<Fr> pi_poly = dom_fft(dn, dense_pi.data(), n, true, false, T); trim(pi_poly);
  std::vector<Fr> e8[5];
  const std::vector<Fr>* src5[5] = {&zp, &wp[0], &wp[1], &wp[2], &wp[3]};
  const int T5 = std::max(1, T / 5);
#pragma omp parallel for num_threads(std::min(5, T)) schedule(static, 1)
  for (int k = 0; k < 5; k++) e8[k] = dom_fft(d8, src5[k]->data(), src5[k]->size(), false, true, T5);
  std::vector<Fr> pi8 = dom_fft(d8, pi_poly.data(), pi_poly.size(), false, true, T);
  std::vector<Fr> quotient(n8);
  {
    // first_lagrange_coset_evaluations (quotient_poly.rs:265-284)
    std::vector<Fr> l1(n8);
    for (size_t i = 0; i < n8; i++) l1[i] = P->linear8[i] - Fr::one();
    { std::vector<Fr> prod(n8); Fr tmp = Fr::one(); for (size_t i = 0; i < n8; i++) { tmp = tmp * l1[i]; prod[i] = tmp; } tmp = tmp.inv();
      for (size_t i = n8; i-- > 0;) { Fr s = i ? prod[i - 1] : Fr::one(); Fr nt = tmp * l1[i]; l1[i] = tmp * s; tmp = nt; } }
    Fr psi = d8.size_inv * fr_small(8), a2 = alpha.sqr();
#pragma omp parallel for num_threads(T) schedule(static)
    for (size_t i = 0; i < n8; i++) {
      size_t iw = (i + 8) % n8;
      Wire8 v = {e8[1][i], e8[2][i], e8[3][i], e8[4][i], e8[1][iw], e8[2][iw], e8[4][iw]};
      const Fr &z = e8[0][i], &z_w = e8[0][iw];
      auto q = [&](int k) -> const Fr& { return P->evals8[k][i]; };
      Fr t = (v.a * v.b * q(Q_M) + v.a * q(Q_L) + v.b * q(Q_R) + v.c * q(Q_O) + v.d * q(Q_F) + q(Q_C)) * q(Q_ARITH);
      t = t + w_range(ch_range, v) * q(Q_RANGE);
      t = t + w_logic(ch_logic, q(Q_C), v) * q(Q_LOGIC);
      t = t + w_fixed(ch_fixed, q(Q_L), q(Q_R), q(Q_C), v) * q(Q_FIXED);
      t = t + w_var(ch_var, v) * q(Q_VAR);
      t = t + pi8[i];
      const Fr& x = P->linear8[i];
      Fr ident = (v.a + beta * x + gamma) * (v.b + beta * K(1) * x + gamma) * (v.c + beta * K(2) * x + gamma) * (v.d + beta * K(3) * x + gamma) * z * alpha;
      Fr copy = (v.a + beta * q(S1) + gamma) * (v.b + beta * q(S2) + gamma) * (v.c + beta * q(S3) + gamma) * (v.d + beta * q(S4) + gamma) * z_w * alpha;
      Fr l1a = l1[i] * (P->vh8[i] * psi) * a2;
      t = t + ident - copy + (z - Fr::one()) * l1a;
      quotient[i] = t * P->vh_inv[i & 7];
    }
  }
  std::vector<Fr> tp = dom_fft(d8, quotient.data(), n8, true, true, T); trim(tp);
  if (tp.size() > 7 * n) return -5;
  std::vector<Fr> tq[4];
  for (int k = 0; k < 4; k++) {
    size_t lo = k * n, hi = k == 3 ? std::max(tp.size(), lo) : (k + 1) * n;
    for (size_t i = lo; i < hi; i++) tq[k].push_back(i < tp.size() ? tp[i] : Fr::zero());
  }
  if (tq[3].empty()) return -4;
  tq[0].push_back(BL[11]); tq[1][0] = tq[1][0] - BL[11];
  tq[1].push_back(BL[12]); tq[2][0] = tq[2][0] - BL[12];
  tq[2].push_back(BL[13]); tq[3][0] = tq[3][0] - BL[13];
  for (int k = 0; k < 4; k++) trim(tq[k]);
  Aff tc[4];
#pragma omp parallel for num_threads(std::min(4, T)) schedule(static, 1)
  for (int k = 0; k < 4; k++) if (!commit(P->ck, tq[k].data(), tq[k].size(), &tc[k], T4)) ok = false;
  if (!ok) return -3;
  tr.append_commitment("t_low_comm", tc[0]); tr.append_commitment("t_mid_comm", tc[1]); tr.append_commitment("t_high_comm", tc[2]); tr.append_commitment("t_fourth_comm", tc[3]);

  // round 4
  Fr z_ch = tr.challenge_scalar("z_challenge"), zw = z_ch * dn.gen;
  Fr ev[15];
  enum { E_A, E_B, E_C, E_D, E_AW, E_BW, E_DW, E_QARITH, E_QC, E_QL, E_QR, E_S1, E_S2, E_S3, E_Z };
  {
    const std::vector<Fr>* ps[15] = {&wp[0], &wp[1], &wp[2], &wp[3], &wp[0], &wp[1], &wp[3], &P->polys[Q_ARITH], &P->polys[Q_C], &P->polys[Q_L], &P->polys[Q_R], &P->polys[S1], &P->polys[S2], &P->polys[S3], &zp};
    const Fr* pts[15] = {&z_ch, &z_ch, &z_ch, &z_ch, &zw, &zw, &zw, &z_ch, &z_ch, &z_ch, &z_ch, &z_ch, &z_ch, &z_ch, &zw};
#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
    for (int k = 0; k < 15; k++) ev[k] = poly_eval(*ps[k], *pts[k]);
  }
  tr.append_scalar("a_eval", ev[E_A]); tr.append_scalar("b_eval", ev[E_B]); tr.append_scalar("c_eval", ev[E_C]); tr.append_scalar("d_eval", ev[E_D]);
  tr.append_scalar("s_sigma_1_eval", ev[E_S1]); tr.append_scalar("s_sigma_2_eval", ev[E_S2]); tr.append_scalar("s_sigma_3_eval", ev[E_S3]);
  tr.append_scalar("z_eval", ev[E_Z]);
  tr.append_scalar("a_w_eval", ev[E_AW]); tr.append_scalar("b_w_eval", ev[E_BW]); tr.append_scalar("d_w_eval", ev[E_DW]);
  tr.append_scalar("q_arith_eval", ev[E_QARITH]); tr.append_scalar("q_c_eval", ev[E_QC]); tr.append_scalar("q_l_eval", ev[E_QL]); tr.append_scalar("q_r_eval", ev[E_QR]);

  // round 5 (linearization_poly.rs:168-264; the constant PI term cannot influence W_z and is omitted)
  Fr v_ch = tr.challenge_scalar("v_challenge");
  Wire8 ve = {ev[E_A], ev[E_B], ev[E_C], ev[E_D], ev[E_AW], ev[E_BW], ev[E_DW]};
  size_t rlen = 0;
  for (int k = 0; k < N_POLY; k++) rlen = std::max(rlen, P->polys[k].size());
  rlen = std::max(rlen, zp.size());
  for (int k = 0; k < 4; k++) rlen = std::max(rlen, tq[k].size());
  std::vector<Fr> r(rlen, Fr::zero());
  auto axpy = [&](const std::vector<Fr>& p, const Fr& k) {
#pragma omp parallel for num_threads(T) schedule(static)
    for (size_t i = 0; i < p.size(); i++) r[i] = r[i] + p[i] * k;
  };
  axpy(P->polys[Q_M], ev[E_A] * ev[E_B] * ev[E_QARITH]); axpy(P->polys[Q_L], ev[E_A] * ev[E_QARITH]); axpy(P->polys[Q_R], ev[E_B] * ev[E_QARITH]);
  axpy(P->polys[Q_O], ev[E_C] * ev[E_QARITH]); axpy(P->polys[Q_F], ev[E_D] * ev[E_QARITH]); axpy(P->polys[Q_C], ev[E_QARITH]);
  axpy(P->polys[Q_RANGE], w_range(ch_range, ve)); axpy(P->polys[Q_LOGIC], w_logic(ch_logic, ev[E_QC], ve));
  axpy(P->polys[Q_FIXED], w_fixed(ch_fixed, ev[E_QL], ev[E_QR], ev[E_QC], ve)); axpy(P->polys[Q_VAR], w_var(ch_var, ve));
  Fr bz = beta * z_ch;
  Fr s_ident = (ev[E_A] + bz + gamma) * (ev[E_B] + K(1) * bz + gamma) * (ev[E_C] + K(2) * bz + gamma) * (ev[E_D] + K(3) * bz + gamma) * alpha;
  Fr s_copy = (ev[E_A] + beta * ev[E_S1] + gamma) * (ev[E_B] + beta * ev[E_S2] + gamma) * (ev[E_C] + beta * ev[E_S3] + gamma) * (beta * ev[E_Z]) * alpha;
  Fr z_n = z_ch.pow64(n);
  Fr l1_z = (z_n - Fr::one()) * dn.size_inv * (z_ch - Fr::one()).inv();
  axpy(zp, s_ident + l1_z * alpha.sqr());
  axpy(P->polys[S4], s_copy.neg());
  Fr zh = (z_n - Fr::one()).neg();
  axpy(tq[0], zh); axpy(tq[1], zh * z_n); axpy(tq[2], zh * z_n.sqr()); axpy(tq[3], zh * z_n.sqr() * z_n);
  trim(r);
  auto aggregate = [&](std::vector<const std::vector<Fr>*> polys, const Fr& point, const Fr& v) {
    size_t mx = 0; for (auto p : polys) mx = std::max(mx, p->size());
    std::vector<Fr> c(mx, Fr::zero());
    Fr pw = Fr::one();
    for (auto p : polys) {
#pragma omp parallel for num_threads(T) schedule(static)
      for (size_t i = 0; i < p->size(); i++) c[i] = c[i] + (*p)[i] * pw;
      pw = pw * v;
    }
    trim(c);
    std::vector<Fr> q; q.reserve(c.size());  // ruffini (polynomial.rs:345-367)
    Fr k = Fr::zero();
    for (size_t i = c.size(); i-- > 0;) { Fr t = c[i] + k; q.push_back(t); k = point * t; }
    if (!q.empty()) q.pop_back();
    std::reverse(q.begin(), q.end());
    trim(q);
    return q;
  };
  std::vector<Fr> wz = aggregate({&r, &wp[0], &wp[1], &wp[2], &wp[3], &P->polys[S1], &P->polys[S2], &P->polys[S3], &P->polys[Q_ARITH], &P->polys[Q_C], &P->polys[Q_L], &P->polys[Q_R]}, z_ch, v_ch);
  Aff wzc; if (!commit(P->ck, wz.data(), wz.size(), &wzc, T)) return -3;
  Fr v_w = tr.challenge_scalar("v_w_challenge");
  std::vector<Fr> wzw = aggregate({&zp, &wp[0], &wp[1], &wp[3]}, zw, v_w);
  Aff wzwc; if (!commit(P->ck, wzw.data(), wzw.size(), &wzwc, T)) return -3;

  const Aff* cs[11] = {&wc[0], &wc[1], &wc[2], &wc[3], &zc, &tc[0], &tc[1], &tc[2], &tc[3], &wzc, &wzwc};
  for (int i = 0; i < 11; i++) g1_compress(*cs[i], out_proof + 48 * i);
  for (int i = 0; i < 15; i++) fr_store(out_proof + 528 + 32 * i, ev[i]);
  return 0;
}
}
