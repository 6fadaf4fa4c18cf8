// Device-resident PLONK prover: Prover::new / Prover::prove (PlonkVersion::V3) of the reference
// (src/compiler/prover.rs:53-115, 415-761; preprocessing as Compiler::preprocess,
// src/compiler.rs:132-461) with every polynomial kept in HBM between rounds.
//
// The two hot kernels (NTT: ntt.cu, G1 MSM: msm.cu) are driven exactly where the reference calls
// domain.{ifft,coset_fft,coset_ifft} and commit_key.commit; the O(n) glue between them
//  runs as small streaming kernels here so that per proof only the
// witnesses go up and 11 x 96 B commitments + 15 x 32 B evaluations come down:
//   round 1  gather wires -> 4 x iNTT(n) -> blind -> 4 commitments            prover.rs:446-479
//   round 2  grand product (batched inversion + prefix-product scan) -> iNTT -> commit  :483-505,
//            composer/permutation.rs:213-294
//   round 3  6 x coset NTT(8n) -> fused gate/permutation quotient kernel -> coset iNTT(8n)
//            -> split + blind -> 4 commitments            quotient_poly.rs:20-310, prover.rs:545-589
//   round 4  15 evaluations (chunked Horner + tree sum)                          prover.rs:593-658
//   round 5  linearisation + both aggregate witnesses as one linear combination per opening,
//            division by (X - z) as a suffix scan, 2 commitments   linearization_poly.rs:168-231,
//            key.rs:394-417, polynomial.rs:345-367
// The Fiat-Shamir transcript (transcript.h) and a handful of scalar formulas run on the host
// between rounds.
#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "host_field.h"
#include "transcript.h"

struct pb200_srs;

namespace pb {

int ntt_run(const uint64_t* d_in, size_t in_len, uint64_t* d_out, uint32_t log_n, int inverse, int coset,
            uint32_t batch, size_t in_stride, size_t out_stride, cudaStream_t st, Arena* ar);
int msm_run(const pb200_srs* srs, size_t first, const uint64_t* d_scalars, size_t n, uint32_t batch,
            size_t stride, uint64_t* out_affine_host, cudaStream_t st, Arena* ar);
size_t msm_workspace_bytes(const pb200_srs* srs, size_t n, uint32_t batch);
int srs_upload(const uint8_t* raw, size_t n_points, pb200_srs** out);
void srs_free(pb200_srs* s);
size_t srs_len(const pb200_srs* s);
const uint4* srs_points(const pb200_srs* s);
int srs_from_device(const uint4* d_points, size_t n_points, pb200_srs** out, int window_bits);
int msm_window_for(size_t n_points);
extern thread_local int t_msm_throughput_hint;
// pb200_throughput_mode(1): treat every proof as one of many in flight (a measurement aid: bench.py times the
// dominant kernel with single proofs but wants the launch shape of its timed region)
static std::atomic<int> g_force_throughput{0};
int g1_check_raw(const uint8_t* raw, size_t n);
int raw_commit_key_parse(const uint8_t* bytes, size_t len, int checked, size_t* n_points, uint8_t* out_raw);
int lagrange_key_dev(const uint4* d_in, int log_n, uint4* d_out, cudaStream_t st);
int get_twiddles(int logm, bool inverse, cudaStream_t st, const uint4** out);
int fill_powers(uint4* out, size_t n, const Fr& base, const Fr& scale, cudaStream_t st);
Fr ntt_group_gen(int log_n, bool inverse);
Fr ntt_size_inv(int log_n);
Fr ntt_coset_gen(bool inverse);

enum { Q_M, Q_L, Q_R, Q_O, Q_F, Q_C, Q_ARITH, Q_RANGE, Q_LOGIC, Q_FIXED, Q_VAR, S1, S2, S3, S4, N_POLY };

PB_D Fr ldg_fr(const uint4* p, size_t i) {
  uint4 a = __ldg(p + 2 * i), b = __ldg(p + 2 * i + 1);
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
PB_D Fr ld_fr_plain(const uint4* p, size_t i) {
  uint4 a = p[2 * i], b = p[2 * i + 1];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
PB_D void stg_fr(uint4* p, size_t i, const Fr& r) {
  p[2 * i] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  p[2 * i + 1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
PB_D Fr lds_pair(const uint4* s) {  // two consecutive uint4 in shared memory
  uint4 a = s[0], b = s[1];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
PB_D Fr fr_small(uint32_t x) {  // Montgomery form of a small constant
  Fr r = Fr::zero();
  r.v[0] = x;
  return r.to_mont();
}

// ---------------------------------------------------------------------------------------------
// small streaming kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_gather_wires(const uint4* wit, const uint32_t* wires, size_t constraints, size_t n, uint4* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned k = blockIdx.y;
  Fr v = Fr::zero();
  if (i < constraints) v = ldg_fr(wit, wires[k * constraints + i]);
  stg_fr(out, (size_t)k * n + i, v);
}

// coeffs[i] -= b_i ; coeffs[n + i] = b_i  (Prover::blind_poly_with_blinders, prover.rs:139-152)
struct BlindArgs {
  Fr b[4][3];
  int nb;
  int npoly;
};
// Scalars of the blinder terms of a Lagrange-basis wire commitment: blinding adds b_k X^k (X^n - 1), so
// after the n wire values come -b_0, -b_1 (against [1]G, [x]G) and b_0, b_1 (against [x^n]G, [x^(n+1)]G).
__global__ void k_lagrange_tail(uint4* sc, size_t stride, size_t n, BlindArgs a) {
  const int t = threadIdx.x;
  if (t >= 2 * a.npoly) return;
  const int p = t >> 1, k = t & 1;
  stg_fr(sc, (size_t)p * stride + n + k, a.b[p][k].neg());
  stg_fr(sc, (size_t)p * stride + n + 2 + k, a.b[p][k]);
}
__global__ void k_blind(uint4* polys, size_t stride, size_t n, BlindArgs a) {
  const int t = threadIdx.x;
  if (t >= a.npoly * a.nb) return;
  const int p = t / a.nb, i = t % a.nb;
  uint4* c = polys + 2 * (size_t)p * stride;
  stg_fr(c, i, ld_fr_plain(c, i) - a.b[p][i]);
  stg_fr(c, n + i, a.b[p][i]);
}

// BlsScalar::from_bytes for a whole array (canonical little-endian integers -> Montgomery form); a value
// >= r is not canonical and raises `flag` (dusk_bytes::Error::InvalidData in the reference).
__global__ void k_fr_from_canonical(uint4* p, size_t n_elems, unsigned* flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_elems) return;
  const Fr x = ld_fr_plain(p, i);
  bool lt = false;
#pragma unroll
  for (int k = 7; k >= 0; k--) {
    const uint32_t m = FrParams::MOD(k);
    if (x.v[k] != m) {
      lt = x.v[k] < m;
      break;
    }
  }
  if (!lt) atomicOr(flag, 1u);
  stg_fr(p, i, x.to_mont());
}

__global__ void k_zero(uint4* p, size_t n_elems) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_elems) stg_fr(p, i, Fr::zero());
}

// sigma Lagrange values: K_col * w^idx for the encoded permutation target (col << 40 | idx)
__global__ void k_sigma_lagrange(const unsigned long long* sig, size_t n, const uint4* w_half, uint4* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * n) return;
  const unsigned long long e = sig[i];
  const unsigned col = (unsigned)(e >> 40);
  const size_t idx = (size_t)(e & 0xffffffffffull);
  Fr root = (idx < n / 2 || n == 1) ? ldg_fr(w_half, idx) : ldg_fr(w_half, idx - n / 2).neg();
  const uint32_t ks[4] = {1, 7, 13, 17};
  stg_fr(out, i, root * fr_small(ks[col]));
}

// numerators and denominators of the grand product (permutation.rs:252-294)
__global__ void k_perm_terms(const uint4* wv, const uint4* sigma, const uint4* w_half, size_t n, Fr beta, Fr gamma,
                             uint4* num, uint4* den) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr root = (i < n / 2 || n == 1) ? ldg_fr(w_half, i) : ldg_fr(w_half, i - n / 2).neg();
  Fr br = beta * root;
  Fr a = ldg_fr(wv, i), b = ldg_fr(wv, n + i), c = ldg_fr(wv, 2 * n + i), d = ldg_fr(wv, 3 * n + i);
  Fr nu = (a + br + gamma) * (b + br * fr_small(7) + gamma) * (c + br * fr_small(13) + gamma) * (d + br * fr_small(17) + gamma);
  Fr de = (a + beta * ldg_fr(sigma, i) + gamma) * (b + beta * ldg_fr(sigma, n + i) + gamma) *
          (c + beta * ldg_fr(sigma, 2 * n + i) + gamma) * (d + beta * ldg_fr(sigma, 3 * n + i) + gamma);
  stg_fr(num, i, nu);
  stg_fr(den, i, de);
}

// out[i] = a[i] / b[i] (a may be null: out = 1/b).  Montgomery's trick over chunks of 8 per thread;
// zeros are left as zeros like util::batch_inversion (util.rs:87-118).
__global__ void k_batch_div(const uint4* a, const uint4* b, size_t n, uint4* out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t base = t * 8;
  if (base >= n) return;
  Fr x[8], p[8];
  Fr acc = Fr::one();
#pragma unroll
  for (int k = 0; k < 8; k++) {
    x[k] = (base + k < n) ? ld_fr_plain(b, base + k) : Fr::one();
    if (!x[k].is_zero()) acc = acc * x[k];
    p[k] = acc;
  }
  Fr inv = acc.inv();
#pragma unroll
  for (int k = 7; k >= 0; k--) {
    if (base + k >= n) continue;
    Fr r;
    if (x[k].is_zero()) {
      r = Fr::zero();
    } else {
      r = (k > 0) ? inv * p[k - 1] : inv;
      inv = inv * x[k];
    }
    if (a) r = r * ld_fr_plain(a, base + k);
    stg_fr(out, base + k, r);
  }
}

// ---------------------------------------------------------------------------------------------
// Scan over Fr (op = multiply or add), exclusive, optional reversed index order.
// Phase 1: 256 threads x 8 elements per CTA; phase 2: scan of the CTA totals; phase 3: fix-up.
// ---------------------------------------------------------------------------------------------
template <bool MUL>
PB_D Fr scan_op(const Fr& a, const Fr& b) {
  return MUL ? a * b : a + b;
}
template <bool MUL>
PB_D Fr scan_id() {
  return MUL ? Fr::one() : Fr::zero();
}

template <bool MUL, bool REV>
__global__ void __launch_bounds__(256) k_scan_local(const uint4* in, size_t n, uint4* out, uint4* totals) {
  __shared__ uint4 sh[2][256][2];
  const int tid = threadIdx.x;
  const size_t base = ((size_t)blockIdx.x * 256 + tid) * 8;
  Fr p[8];
  Fr acc = scan_id<MUL>();
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const size_t i = base + k;
    Fr x = (i < n) ? ld_fr_plain(in, REV ? (n - 1 - i) : i) : scan_id<MUL>();
    acc = scan_op<MUL>(acc, x);
    p[k] = acc;
  }
  int cur = 0;
  sh[0][tid][0] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
  sh[0][tid][1] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    Fr v = lds_pair(sh[cur][tid]);
    if (tid >= d) v = scan_op<MUL>(lds_pair(sh[cur][tid - d]), v);
    sh[cur ^ 1][tid][0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    sh[cur ^ 1][tid][1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
    cur ^= 1;
    __syncthreads();
  }
  Fr prefix = (tid > 0) ? lds_pair(sh[cur][tid - 1]) : scan_id<MUL>();
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const size_t i = base + k;
    if (i < n) {
      Fr v = (k > 0) ? scan_op<MUL>(prefix, p[k - 1]) : prefix;
      stg_fr(out, REV ? (n - 1 - i) : i, v);
    }
  }
  if (tid == 255 && totals) {
    Fr tot = lds_pair(sh[cur][255]);
    stg_fr(totals, blockIdx.x, tot);
  }
}

template <bool MUL, bool REV>
__global__ void k_scan_fixup(uint4* out, size_t n, const uint4* block_prefix) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t blk = i / 2048;
  if (blk == 0) return;
  const size_t j = REV ? (n - 1 - i) : i;
  stg_fr(out, j, scan_op<MUL>(ld_fr_plain(block_prefix, blk), ld_fr_plain(out, j)));
}

// ---------------------------------------------------------------------------------------------
// Quotient numerator on the 8n coset (quotient_poly.rs:160-310 + all widget compute_quotient_i)
// ---------------------------------------------------------------------------------------------
struct WidgetCh {
  Fr ch, k, k2, k3, k4;
};

struct QuotArgs {
  const uint4* w8;      // [6][8n]: z, a, b, c, d, pi coset evaluations
  const uint4* key8;    // [15][8n] prover-key coset evaluations (enum order)
  const uint4* linear8; // [8n]
  const uint4* l1_8;    // [8n] L_1 on the coset (without alpha^2)
  uint4* out;           // [8n]
  size_t n8;
  Fr alpha, beta, gamma, alpha_sq;
  // separation challenges with their powers (kappa = ch^2, kappa^2, ...), computed once on the host
  // instead of once per coset point
  WidgetCh ch_range, ch_logic, ch_fixed, ch_var;
  Fr edwards_d;  // dusk_jubjub::EDWARDS_D, Montgomery form
  Fr vh_inv[8];
  int has_range, has_logic, has_fixed, has_var;
};

// The quotient kernel evaluates ~90 Fr products per coset point.  Fully inlined that is ~23 k
// instructions per thread (~360 KB of SASS), several times what the instruction cache holds: ncu
// showed its warps waiting for instructions (stall no_instruction 2.0 per issue, multiply pipe 29 %
// busy).  Inside this kernel every product therefore goes through one out-of-line routine; `Q` is
// Fr with that product.
__device__ __noinline__ Fr fr_mul_outlined(Fr a, Fr b) { return a * b; }

struct Q {
  Fr v;
  PB_D Q() {}
  PB_D Q(const Fr& f) : v(f) {}
  static PB_D Q one() { return Q(Fr::one()); }
  static PB_D Q zero() { return Q(Fr::zero()); }
  friend PB_D Q operator+(const Q& x, const Q& y) { return Q(x.v + y.v); }
  friend PB_D Q operator-(const Q& x, const Q& y) { return Q(x.v - y.v); }
  friend PB_D Q operator*(const Q& x, const Q& y) { return Q(fr_mul_outlined(x.v, y.v)); }
  PB_D Q sqr() const { return Q(fr_mul_outlined(v, v)); }
  PB_D Q dbl() const { return Q(v.dbl()); }
};

template <class F>
PB_D F delta4(const F& f) {  // f (f - 1)(f - 2)(f - 3) = g (g + 2) with g = f (f - 3): two products
  const F one = F::one();
  const F g = f * (f - one - one - one);
  return g * (g + one + one);
}
template <class F>
PB_D F mul_small(const F& x, int k) {  // k * x for small positive k by additions
  F acc = F::zero(), p = x;
  while (k) {
    if (k & 1) acc = acc + p;
    p = p.dbl();
    k >>= 1;
  }
  return acc;
}

struct WireVals {
  Q a, b, c, d, a_w, b_w, d_w;
};
PB_D Q widget_range(const WidgetCh& s, const WireVals& v) {  // range/proverkey.rs:32-57 (without selector)
  const Q &ch = s.ch, &k = s.k, &k2 = s.k2, &k3 = s.k3;
  Q t = delta4(v.c - mul_small(v.d, 4)) + delta4(v.b - mul_small(v.c, 4)) * k + delta4(v.a - mul_small(v.b, 4)) * k2 +
         delta4(v.d_w - mul_small(v.a, 4)) * k3;
  return t * ch;
}
PB_D Q widget_logic(const WidgetCh& s, const Q& q_c, const WireVals& v) {  // logic/proverkey.rs:34-71, 120-144
  const Q &ch = s.ch, &k = s.k, &k2 = s.k2, &k3 = s.k3, &k4 = s.k4;
  Q A = v.a_w - mul_small(v.a, 4), B = v.b_w - mul_small(v.b, 4), D = v.d_w - mul_small(v.d, 4);
  const Q& w = v.c;
  Q ab = A + B;
  Q F = w * (w * (mul_small(w, 4) - mul_small(ab, 18) + mul_small(Q::one(), 81)) + mul_small(A.sqr() + B.sqr(), 18) - mul_small(ab, 81) + mul_small(Q::one(), 83));
  Q E = mul_small(ab + D, 3) - F.dbl();
  Q Bq = q_c * (mul_small(D, 9) - mul_small(ab, 3));
  Q t = delta4(A) + delta4(B) * k + delta4(D) * k2 + (w - A * B) * k3 + (Bq + E) * k4;
  return t * ch;
}
PB_D Q widget_fixed(const WidgetCh& s, const Q& ed, const Q& q_l, const Q& q_r, const Q& q_c, const WireVals& v) {  // fixed_base/proverkey.rs:39-103
  const Q one = Q::one();
  const Q &ch = s.ch, &k = s.k, &k2 = s.k2, &k3 = s.k3;
  Q bit = v.d_w - v.d - v.d;
  Q bit_c = bit * (bit - one) * (bit + one);
  Q y_alpha = bit.sqr() * (q_r - one) + one, x_alpha = bit * q_l;
  Q xy = (bit * q_c - v.c) * k;
  Q t = v.c * v.a * v.b * ed;
  Q xa = ((v.a_w + v.a_w * t) - (v.a * y_alpha + v.b * x_alpha)) * k2;
  Q ya = ((v.b_w - v.b_w * t) - (v.b * y_alpha + v.a * x_alpha)) * k3;
  return (bit_c + xa + ya + xy) * ch;
}
PB_D Q widget_var(const WidgetCh& s, const Q& ed, const WireVals& v) {  // curve_addition/proverkey.rs:33-79
  const Q &ch = s.ch, &k = s.k;
  const Q &x1 = v.a, &x3 = v.a_w, &y1 = v.b, &y3 = v.b_w, &x2 = v.c, &y2 = v.d, &x1y2 = v.d_w;
  Q xy = x1 * y2 - x1y2, y1x2 = y1 * x2, y1y2 = y1 * y2, x1x2 = x1 * x2;
  Q t = ed * x1y2 * y1x2;
  Q x3c = ((x1y2 + y1x2) - (x3 + x3 * t)) * k;
  Q y3c = ((y1y2 + x1x2) - (y3 - y3 * t)) * s.k2;
  return (xy + x3c + y3c) * ch;
}

// One point of the quotient.  The six witness rows (z, a, b, c, d, pi) have stride `ws` and are read at
// column i, the "next row" values (X -> omega X) at column iw; the prover-key tables (stride n8),
// the point itself, L_1 and 1/Z_H are read at index ki of the 8n coset; the result goes to out[oi].
//   8n coset (the reference's schedule): ws = 8n, iw = i + 8 mod 8n, ki = oi = i
//   4n coset (its even points):          ws = 4n, iw = i + 4 mod 4n, ki = 2i, oi = i
//   single points (Horner-evaluated):    ws = 16, iw = i + 8,        ki = 1 + n i, oi = i
PB_D void quotient_point(const QuotArgs& q, size_t ws, size_t i, size_t iw, size_t ki, size_t oi) {
  const size_t n8 = q.n8;
  WireVals v;
  const Q z = ldg_fr(q.w8, i), z_w = ldg_fr(q.w8, iw);
  v.a = ldg_fr(q.w8, ws + i); v.a_w = ldg_fr(q.w8, ws + iw);
  v.b = ldg_fr(q.w8, 2 * ws + i); v.b_w = ldg_fr(q.w8, 2 * ws + iw);
  v.c = ldg_fr(q.w8, 3 * ws + i);
  v.d = ldg_fr(q.w8, 4 * ws + i); v.d_w = ldg_fr(q.w8, 4 * ws + iw);
  const Q pi = ldg_fr(q.w8, 5 * ws + i);
#define KEY(k) ldg_fr(q.key8, (size_t)(k) * n8 + ki)
  const Q q_l = KEY(Q_L), q_r = KEY(Q_R), q_c = KEY(Q_C);
  // arithmetic/proverkey.rs:44-69
  Q t = (v.a * v.b * KEY(Q_M) + v.a * q_l + v.b * q_r + v.c * KEY(Q_O) + v.d * KEY(Q_F) + q_c) * KEY(Q_ARITH);
  if (q.has_range) t = t + widget_range(q.ch_range, v) * KEY(Q_RANGE);
  if (q.has_logic) t = t + widget_logic(q.ch_logic, q_c, v) * KEY(Q_LOGIC);
  if (q.has_fixed) t = t + widget_fixed(q.ch_fixed, q.edwards_d, q_l, q_r, q_c, v) * KEY(Q_FIXED);
  if (q.has_var) t = t + widget_var(q.ch_var, q.edwards_d, v) * KEY(Q_VAR);
  t = t + pi;
  // permutation/proverkey.rs:40-125
  const Q x = ldg_fr(q.linear8, ki);
  const Q alpha = q.alpha, beta = q.beta, gamma = q.gamma;
  const Q bx = beta * x;
  Q ident = (v.a + bx + gamma) * (v.b + mul_small(bx, 7) + gamma) * (v.c + mul_small(bx, 13) + gamma) *
            (v.d + mul_small(bx, 17) + gamma) * z * alpha;
  Q copy = (v.a + beta * KEY(S1) + gamma) * (v.b + beta * KEY(S2) + gamma) * (v.c + beta * KEY(S3) + gamma) *
           (v.d + beta * KEY(S4) + gamma) * z_w * alpha;
#undef KEY
  Q l1 = Q(ldg_fr(q.l1_8, ki)) * Q(q.alpha_sq);
  t = t + ident - copy + (z - Q::one()) * l1;
  stg_fr(q.out, oi, (t * Q(q.vh_inv[ki & 7])).v);
}

__global__ void __launch_bounds__(128) k_quotient(QuotArgs q) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q.n8) return;
  quotient_point(q, q.n8, i, (i + 8) & (q.n8 - 1), i, i);
}
// The same on the 4n coset g*H_4n = the even points of the 8n coset (PB200_QUOT4N=1, see prove_dev).
// MINB = resident CTAs per SM the register allocation is made for: 2 -> 240 registers, no spills; 3 -> 168
// registers and ~300 bytes of spills per thread (PB200_QUOT_OCC selects; measured).
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_quotient_4n(QuotArgs q) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = q.n8 >> 1;
  if (i >= n4) return;
  quotient_point(q, n4, i, (i + 4) & (n4 - 1), 2 * i, i);
}
// ... and at the eight odd points 1 + n k of the 8n coset, the witness values coming from Horner
// evaluations laid out as [6][16] (columns 0..7: x_k, columns 8..15: omega x_k).
__global__ void __launch_bounds__(128) k_quotient_pts(QuotArgs q) {
  const size_t k = threadIdx.x;
  if (k >= 8) return;
  quotient_point(q, 16, k, k + 8, 1 + (q.n8 >> 3) * k, k);
}

// flag |= any nonzero element in p[lo, hi)
__global__ void k_any_nonzero(const uint4* p, size_t lo, size_t hi, unsigned* flag) {
  size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  uint4 a = p[2 * i], b = p[2 * i + 1];
  if (a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) atomicOr(flag, 1u);
}

// split t(X) into four polynomials of stride `stride` and apply b12..b14 (prover.rs:545-574)
__global__ void k_split_quotient(const uint4* t, size_t n, size_t n8, size_t stride, Fr b12, Fr b13, Fr b14, uint4* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= stride) return;
  const unsigned k = blockIdx.y;
  Fr v = Fr::zero();
  if (k < 3) {
    if (i < n) v = ld_fr_plain(t, (size_t)k * n + i);
    if (i == n) v = (k == 0) ? b12 : (k == 1 ? b13 : b14);
    if (i == 0 && k == 1) v = v - b12;
    if (i == 0 && k == 2) v = v - b13;
  } else {
    if (3 * n + i < n8) v = ld_fr_plain(t, 3 * n + i);
    if (i == 0) v = v - b14;
  }
  stg_fr(out, (size_t)k * stride + i, v);
}

// Polynomial::evaluate (polynomial.rs:120-137) for batches of (polynomial, point) jobs.  A CTA evaluates a
// block of 2048 coefficients: P_blk(x) = sum_{i < 2048} c[2048 blk + i] x^i - every thread runs Horner over
// its 8 coefficients, then a shared-memory tree folds the 256 thread values with the powers x^8, x^16, ...
// (about 27 products per thread; the first version raised x to each thread's offset with a 64-bit
// exponent, ~90 products).  k_sum_rows then runs Horner over the blocks with x^2048.
PB_D void cta_poly_eval(const uint4* poly, unsigned len, const Fr& x, uint4* out_slot) {
  __shared__ uint4 sh[256][2];
  const int tid = threadIdx.x;
  const size_t base = ((size_t)blockIdx.x * 256 + tid) * 8;
  Fr acc = Fr::zero();
  if (base < len) {
#pragma unroll
    for (int k = 7; k >= 0; k--) {
      const Fr c = (base + k < len) ? ld_fr_plain(poly, base + k) : Fr::zero();
      acc = acc * x + c;
    }
  }
  sh[tid][0] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
  sh[tid][1] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
  Fr pw = x.sqr().sqr().sqr();  // x^8: the weight of the neighbouring thread's value
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    if ((tid & (2 * d - 1)) == 0) {
      const Fr v = lds_pair(sh[tid]) + lds_pair(sh[tid + d]) * pw;
      sh[tid][0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
      sh[tid][1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
      pw = pw.sqr();
    }
    __syncthreads();
  }
  if (tid == 0) stg_fr(out_slot, 0, lds_pair(sh[0]));
}

struct EvalJobs {
  const uint4* poly[16];
  unsigned len[16];
  Fr point[16];
  int njobs;
};
__global__ void __launch_bounds__(256) k_poly_eval(EvalJobs jobs, uint4* partial, unsigned nblocks) {
  const int j = blockIdx.y;
  cta_poly_eval(jobs.poly[j], jobs.len[j], jobs.point[j], partial + 2 * ((size_t)j * nblocks + blockIdx.x));
}
// The same for one set of points and several polynomials laid out with a fixed stride: job (y, z) evaluates
// base + z * row_stride at point[y]; partial is [z][y][nblocks].
struct EvalRows {
  const uint4* base;
  size_t row_stride;  // elements
  unsigned len;
  Fr point[16];
};
__global__ void __launch_bounds__(256) k_poly_eval_rows(EvalRows jobs, uint4* partial, unsigned nblocks) {
  cta_poly_eval(jobs.base + 2 * (size_t)blockIdx.z * jobs.row_stride, jobs.len, jobs.point[blockIdx.y],
                partial + 2 * (((size_t)blockIdx.z * gridDim.y + blockIdx.y) * nblocks + blockIdx.x));
}
// out[j] = sum_blk partial[j][blk] x_j^(2048 blk), x_j = pts.p[j mod npts]: Horner over the blocks
struct EvalPoints {
  Fr p[16];
};
__global__ void k_sum_rows(const uint4* partial, unsigned nblocks, EvalPoints pts, int npts, uint4* out) {
  const int j = blockIdx.x;
  if (threadIdx.x != 0) return;
  Fr step = pts.p[j % npts];
#pragma unroll 1
  for (int k = 0; k < 11; k++) step = step.sqr();  // x^2048
  Fr acc = Fr::zero();
#pragma unroll 1
  for (unsigned b = nblocks; b-- > 0;) acc = acc * step + ld_fr_plain(partial, (size_t)j * nblocks + b);
  stg_fr(out, j, acc);
}

// Step 3 + 4 of the 4n-coset quotient (see prove_dev): thread j < 8 computes
//   t_hi[j] = (h^-j / 8) * sum_k e_k w8^(-jk),   e_k = (t(x_k) - u(x_k)) / (-2 g^4n),
// then patches t's coefficients j and 4n + j (j < 7); a non-zero t_hi[7] raises `flag`.
struct Quot4nFix {
  Fr c;             // 1 / (-2 g^4n)
  Fr g4n;
  Fr hinv8[8];      // h^-j / 8
  Fr w8inv_pow[8];  // w8^-m
};
__global__ void k_quot4n_fix(const uint4* tx, const uint4* ux, uint4* tcoef, size_t n4, Quot4nFix f, unsigned* flag) {
  const int j = threadIdx.x;
  if (j >= 8) return;
  Fr acc = Fr::zero();
  for (int k = 0; k < 8; k++) {
    const Fr e = (ld_fr_plain(tx, k) - ld_fr_plain(ux, k)) * f.c;
    acc = acc + e * f.w8inv_pow[(j * k) & 7];
  }
  const Fr t_hi = acc * f.hinv8[j];
  if (j == 7) {
    if (!t_hi.is_zero()) atomicOr(flag, 1u);
    return;
  }
  stg_fr(tcoef, j, ld_fr_plain(tcoef, j) - f.g4n * t_hi);
  stg_fr(tcoef, n4 + j, t_hi);
}

// out[i] = sum_k coef[k] * poly[k][i]
struct LinArgs {
  const uint4* poly[24];
  unsigned len[24];
  Fr coef[24];
  int nterms;
};
__global__ void k_lincomb(LinArgs a, size_t n, uint4* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr acc = Fr::zero();
  for (int k = 0; k < a.nterms; k++)
    if (i < a.len[k]) acc = acc + ld_fr_plain(a.poly[k], i) * a.coef[k];
  stg_fr(out, i, acc);
}

// d[i] = c[i] * pw[i]
__global__ void k_mul_pointwise(const uint4* a, const uint4* b, size_t n, uint4* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stg_fr(out, i, ld_fr_plain(a, i) * ld_fr_plain(b, i));
}
// out[i] -= 1 ;  out[i] *= c[i & 7]
__global__ void k_sub_one(uint4* p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stg_fr(p, i, ld_fr_plain(p, i) - Fr::one());
}
struct Period8 {
  Fr c[8];
};
__global__ void k_scale_period8(uint4* p, size_t n, Period8 c) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stg_fr(p, i, ld_fr_plain(p, i) * c.c[i & 7]);
}

struct Quot4nConsts {
  Fr xs[16];  // x_k = h w8^k (k < 8), then omega x_k
  Quot4nFix fix;
};

}  // namespace pb

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
using pbh::HFr;

struct pb200_prover {
  std::vector<uint8_t> label;
  size_t constraints = 0, n = 0, n8 = 0;
  int log_n = 0;
  pb200_srs* srs = nullptr;
  // (unless PB200_LAGRANGE=0) [L_0(x)]G .. [L_{n-1}(x)]G, then [1]G, [x]G, [x^n]G, [x^(n+1)]G - the wire polynomials
  // are committed through their values (short scalars) plus the two blinder terms
  pb200_srs* srs_lag = nullptr;
  uint32_t* d_wires = nullptr;  // [4][constraints]
  uint4* d_polys = nullptr;     // [15][n]
  uint4* d_key8 = nullptr;      // [15][8n]
  uint4* d_linear8 = nullptr;   // [8n]
  uint4* d_l1_8 = nullptr;      // [8n]
  uint4* d_sigma = nullptr;     // [4][n]
  HFr vh_inv[8];
  pb::Fr edwards_d;       // dusk_jubjub::EDWARDS_D, Montgomery form
  pb::Quot4nConsts q4;    // constants of the 4n-coset quotient (round 3)
  int has_widget[4] = {0, 0, 0, 0};
  uint8_t comm[pb::N_POLY][48];
  size_t n_witnesses = 0;
  // scratch arenas, one per proof in flight (allocated on first use, then recycled)
  mutable std::mutex ws_mu;
  mutable std::vector<pb::Arena> ws_free;
  mutable std::vector<char*> ws_all;
  size_t ws_bytes = 0;
  mutable std::atomic<int> active{0};  // proofs in flight on this prover (all host threads)
};

namespace pb {

static Fr to_dev(const HFr& x) {
  Fr r;
  memcpy(r.v, x.v, 32);
  return r;
}
static HFr to_host(const Fr& x) {
  HFr r;
  memcpy(r.v, x.v, 32);
  return r;
}
static HFr hfr_pow(const HFr& x, uint64_t e) { return x.pow(&e, 1); }

// Exclusive scan of n elements: CTA-local scans of 2048 elements, the scan of the CTA totals (by the same
// routine, so any length works: 2^22 + 8 elements are 2049 totals, two levels), then the fix-up.
template <bool MUL, bool REV>
static int fr_scan(const uint4* in, size_t n, uint4* out, cudaStream_t st, Arena* ar) {
  const unsigned nblk = div_up(n, 2048);
  uint4 *tot = nullptr, *tot_scan = nullptr;
  ScratchScope scope(ar, st);
  PB_ALLOC(scope, tot, (size_t)nblk * 32);
  PB_ALLOC(scope, tot_scan, (size_t)nblk * 32);
  PB_LAUNCH((k_scan_local<MUL, REV>), nblk, 256, 0, st, in, n, out, tot);
  if (nblk > 1) {
    PB_TRY((fr_scan<MUL, false>((const uint4*)tot, (size_t)nblk, tot_scan, st, ar)));
    PB_LAUNCH((k_scan_fixup<MUL, REV>), div_up(n, 256), 256, 0, st, out, n, (const uint4*)tot_scan);
  }
  PB_CUDA(cudaGetLastError());
  return 0;
}

static void compress_affine(const uint64_t* raw, uint8_t out[48]) { pbh::g1_compress_raw(raw, out); }

void prover_free(pb200_prover* P);

// Stream-ordered scratch that is returned to the pool on every exit path.
struct PoolBlock {
  void* p = nullptr;
  cudaStream_t st;
  explicit PoolBlock(cudaStream_t s) : st(s) {}
  ~PoolBlock() {
    if (p) cudaFreeAsync(p, st);
  }
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes, st); }
  PoolBlock(const PoolBlock&) = delete;
  PoolBlock& operator=(const PoolBlock&) = delete;
};

// A prover key read from Prover::to_bytes: coefficient-form polynomials (canonical scalars) and commitments,
// both already in this file's enum order.
struct LoadedProverKey {
  const uint8_t* poly[N_POLY];
  size_t poly_len[N_POLY];
  uint8_t comm[N_POLY][48];
};

static int prover_build(pb200_prover* P, const uint8_t* label, size_t label_len, size_t constraints, const uint64_t* selectors,
                        const uint32_t* wires, size_t n_witnesses, const uint8_t* srs_raw, size_t n_srs, cudaStream_t st,
                        const LoadedProverKey* loaded = nullptr) {
  P->label.assign(label, label + label_len);
  P->constraints = constraints;
  P->n_witnesses = n_witnesses;
  size_t n_trim = 1;
  while (n_trim < constraints + 6) n_trim <<= 1;  // compiler.rs:121-124
  size_t keep = n_trim + 6;                       // srs.rs:188-196
  if (keep + 1 > n_srs) return fail(PB200_ERR_DEGREE_TOO_LARGE, "public parameters too small for this circuit (TruncatedDegreeTooLarge)");
  size_t n = 1;
  int log_n = 0;
  while (n < constraints) {
    n <<= 1;
    log_n++;
  }
  P->n = n;
  P->n8 = 8 * n;
  P->log_n = log_n;
  if (log_n + 3 >= 32) return fail(PB200_ERR_INVALID_DOMAIN, "quotient domain too large");
  PB_TRY(srs_upload(srs_raw, keep + 1, &P->srs));
  {
    static const bool lagrange_env = [] {
      const char* e = getenv("PB200_LAGRANGE");
      return !e || atoi(e) != 0;
    }();
    if (lagrange_env && n >= 2 && keep + 1 >= n + 2) {
      uint4* comb = nullptr;  // n Lagrange points + 4 monomial ones
      PB_CUDA(cudaMalloc((void**)&comb, (n + 4) * 96));
      const uint4* mono = srs_points(P->srs);
      int rc = lagrange_key_dev(mono, log_n, comb, st);
      const size_t idx[4] = {0, 1, n, n + 1};
      cudaError_t e = cudaSuccess;
      for (int k = 0; k < 4 && e == cudaSuccess; k++)
        e = cudaMemcpyAsync(comb + 6 * (n + k), mono + 6 * idx[k], 96, cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      // Window of the Lagrange-form key.  Its scalars are witness VALUES - mostly zero or a single small digit -
      // so the bucket reduction (~2.3 additions per bucket whatever the scalars) outweighs the accumulation
      // unless the window is narrower than the monomial key's: PB200_LAG_C overrides the default.
      int lag_c = std::min(12, msm_window_for(n + 4));  // measured on BenchCircuit<2^16>: 12 bits 185 proofs/s, 16 bits 180
      if (const char* env = getenv("PB200_LAG_C")) lag_c = atoi(env);
      if (lag_c < 2 || lag_c > 20) lag_c = 0;
      if (rc == 0 && e == cudaSuccess) rc = srs_from_device(comb, n + 4, &P->srs_lag, lag_c);
      cudaFree(comb);
      PB_TRY(rc);
      PB_CUDA(e);
    }
  }
  const size_t n8 = P->n8;
  PB_CUDA(cudaMalloc((void**)&P->d_wires, 4 * constraints * 4));
  PB_CUDA(cudaMalloc((void**)&P->d_polys, (size_t)N_POLY * n * 32));
  PB_CUDA(cudaMalloc((void**)&P->d_key8, (size_t)N_POLY * n8 * 32));
  PB_CUDA(cudaMalloc((void**)&P->d_linear8, n8 * 32));
  PB_CUDA(cudaMalloc((void**)&P->d_l1_8, n8 * 32));
  PB_CUDA(cudaMalloc((void**)&P->d_sigma, 4 * n * 32));
  PB_CUDA(cudaMemcpyAsync(P->d_wires, wires, 4 * constraints * 4, cudaMemcpyHostToDevice, st));

  if (loaded) {
    // Prover::try_from_bytes: the polynomials arrive in coefficient form and the commitments as stored - no
    // interpolation, no MSM
    for (size_t g = 0; g < constraints; g++)
      for (int k = 0; k < 4; k++)
        if (wires[(size_t)k * constraints + g] >= n_witnesses) return fail(PB200_ERR_INVALID_ARG, "wire index out of range");
    PB_CUDA(cudaMemsetAsync(P->d_polys, 0, (size_t)N_POLY * n * 32, st));
    for (int k = 0; k < N_POLY; k++)
      if (loaded->poly_len[k])
        PB_CUDA(cudaMemcpyAsync(P->d_polys + 2 * (size_t)k * n, loaded->poly[k], loaded->poly_len[k] * 32, cudaMemcpyHostToDevice, st));
    PoolBlock flag_block(st);
    PB_CUDA(flag_block.alloc(4));
    unsigned* d_flag = (unsigned*)flag_block.p;
    PB_CUDA(cudaMemsetAsync(d_flag, 0, 4, st));
    PB_LAUNCH(k_fr_from_canonical, div_up((size_t)N_POLY * n, 256), 256, 0, st, P->d_polys, (size_t)N_POLY * n, d_flag);
    unsigned h_flag = 0;
    PB_CUDA(cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    if (h_flag) return fail(PB200_ERR_POINT_MALFORMED, "InvalidData: a prover-key scalar is not canonical");
    memcpy(P->comm, loaded->comm, sizeof P->comm);
    const int widget_sel[4] = {Q_RANGE, Q_LOGIC, Q_FIXED, Q_VAR};
    for (int w = 0; w < 4; w++) P->has_widget[w] = (P->comm[widget_sel[w]][0] & 0x40) ? 0 : 1;
  } else {
    // selector columns, zero padded to n, then iNTT -> coefficient form (compiler.rs:149-211)
    PoolBlock cols_block(st);
    PB_CUDA(cols_block.alloc((size_t)N_POLY * n * 32));
    uint4* cols = (uint4*)cols_block.p;
    PB_CUDA(cudaMemsetAsync(cols, 0, (size_t)N_POLY * n * 32, st));
    PB_CUDA(cudaMemcpy2DAsync(cols, n * 32, selectors, constraints * 32, constraints * 32, 11, cudaMemcpyHostToDevice, st));
    // sigma permutation on the host (composer/permutation.rs:106-141), Lagrange values on the device
    {
      std::vector<std::vector<uint64_t>> wmap(n_witnesses);
      for (size_t g = 0; g < constraints; g++)
        for (int k = 0; k < 4; k++) {
          const uint32_t w = wires[(size_t)k * constraints + g];
          if (w >= n_witnesses) {
            return fail(PB200_ERR_INVALID_ARG, "wire index out of range");
          }
          wmap[w].push_back(((uint64_t)k << 40) | g);
        }
      std::vector<unsigned long long> sig(4 * n);
      for (int k = 0; k < 4; k++)
        for (size_t i = 0; i < n; i++) sig[(size_t)k * n + i] = ((uint64_t)k << 40) | i;
      for (auto& lst : wmap)
        for (size_t i = 0; i < lst.size(); i++) {
          const uint64_t cur = lst[i], nxt = lst[(i + 1) % lst.size()];
          sig[(size_t)(cur >> 40) * n + (cur & 0xffffffffffull)] = nxt;
        }
      PoolBlock sig_block(st);
      PB_CUDA(sig_block.alloc(4 * n * 8));
      unsigned long long* d_sig = (unsigned long long*)sig_block.p;
      PB_CUDA(cudaMemcpyAsync(d_sig, sig.data(), 4 * n * 8, cudaMemcpyHostToDevice, st));
      const uint4* w_half = nullptr;
      PB_TRY(get_twiddles(log_n, false, st, &w_half));
      PB_LAUNCH(k_sigma_lagrange, div_up(4 * n, 256), 256, 0, st, d_sig, n, w_half, cols + 2 * (size_t)S1 * n);
      PB_CUDA(cudaStreamSynchronize(st));  // sig (host vector) must outlive the copy
    }
    PB_TRY(ntt_run((const uint64_t*)cols, n, (uint64_t*)P->d_polys, log_n, 1, 0, N_POLY, n, n, st, nullptr));
    // commitments (compiler.rs:213-232): an all-zero selector commits to the identity
    {
      std::vector<uint64_t> aff((size_t)N_POLY * 12);
      PB_TRY(msm_run(P->srs, 0, (const uint64_t*)P->d_polys, n, N_POLY, n, aff.data(), st, nullptr));
      for (int k = 0; k < N_POLY; k++) compress_affine(aff.data() + 12 * k, P->comm[k]);
      const int widget_sel[4] = {Q_RANGE, Q_LOGIC, Q_FIXED, Q_VAR};
      for (int w = 0; w < 4; w++) P->has_widget[w] = (P->comm[widget_sel[w]][0] & 0x40) ? 0 : 1;
    }
  }
  // coset evaluations over 8n (compiler.rs:306-377)
  PB_TRY(ntt_run((const uint64_t*)P->d_polys, n, (uint64_t*)P->d_key8, log_n + 3, 0, 1, N_POLY, n, n8, st, nullptr));
  {
    HFr lin[2] = {HFr::zero(), HFr::one()};
    PoolBlock lin_block(st);
    PB_CUDA(lin_block.alloc(64));
    uint4* d_lin = (uint4*)lin_block.p;
    PB_CUDA(cudaMemcpyAsync(d_lin, lin, 64, cudaMemcpyHostToDevice, st));
    PB_TRY(ntt_run((const uint64_t*)d_lin, 2, (uint64_t*)P->d_linear8, log_n + 3, 0, 1, 1, 2, n8, st, nullptr));
    PB_CUDA(cudaStreamSynchronize(st));
  }
  // vanishing polynomial on the coset has period 8 (domain.rs:340-351); its inverses are cached
  // (prover.rs:78-91).  L_1 on the coset: vh[i] * (8 / 8n) / (x_i - 1)  (quotient_poly.rs:265-284).
  {
    const HFr g = to_host(ntt_coset_gen(false)), w8n = to_host(ntt_group_gen(log_n + 3, false));
    HFr point = hfr_pow(g, n), step = hfr_pow(w8n, n);
    const HFr psi = to_host(ntt_size_inv(log_n + 3)) * HFr::from_u64(8);
    Period8 c;
    for (int i = 0; i < 8; i++) {
      const HFr vh = point - HFr::one();
      P->vh_inv[i] = vh.inv();
      c.c[i] = to_dev(vh * psi);
      point = point * step;
    }
    PB_CUDA(cudaMemcpyAsync(P->d_l1_8, P->d_linear8, n8 * 32, cudaMemcpyDeviceToDevice, st));
    PB_LAUNCH(k_sub_one, div_up(n8, 256), 256, 0, st, P->d_l1_8, n8);
    PB_LAUNCH(k_batch_div, div_up(div_up(n8, 8), 128), 128, 0, st, (const uint4*)nullptr, (const uint4*)P->d_l1_8, n8, P->d_l1_8);
    PB_LAUNCH(k_scale_period8, div_up(n8, 256), 256, 0, st, P->d_l1_8, n8, c);
  }
  {  // per-domain constants of round 3, computed once (each costs a host-side inversion or power)
    P->edwards_d = to_dev((HFr::from_u64(10240) * HFr::from_u64(10241).inv()).neg());
    const HFr g = to_host(ntt_coset_gen(false)), w8n = to_host(ntt_group_gen(log_n + 3, false));
    const uint64_t e_n[1] = {(uint64_t)n}, e_4n[1] = {(uint64_t)(4 * n)};
    const HFr h = g * w8n, w8r = w8n.pow(e_n, 1), wn = to_host(ntt_group_gen(log_n, false)), g4n = g.pow(e_4n, 1);
    HFr xs[16];
    xs[0] = h;
    for (int k = 1; k < 8; k++) xs[k] = xs[k - 1] * w8r;
    for (int k = 0; k < 8; k++) xs[8 + k] = xs[k] * wn;
    for (int k = 0; k < 16; k++) P->q4.xs[k] = to_dev(xs[k]);
    P->q4.fix.c = to_dev((g4n.dbl().neg()).inv());
    P->q4.fix.g4n = to_dev(g4n);
    const HFr h_inv = h.inv(), w8i = w8r.inv();
    HFr hj = HFr::from_u64(8).inv(), wp = HFr::one();
    for (int j = 0; j < 8; j++) {
      P->q4.fix.hinv8[j] = to_dev(hj);
      P->q4.fix.w8inv_pow[j] = to_dev(wp);
      hj = hj * h_inv;
      wp = wp * w8i;
    }
  }
  // sigma evaluations over n (prover.rs:95-100)
  PB_TRY(ntt_run((const uint64_t*)(P->d_polys + 2 * (size_t)S1 * n), n, (uint64_t*)P->d_sigma, log_n, 0, 0, 4, n, n, st, nullptr));
  PB_CUDA(cudaGetLastError());
  PB_CUDA(cudaStreamSynchronize(st));
  {  // arena size for one proof: prover scratch + the larger of (NTT scratch, MSM scratch)
    const size_t stride = n + 8;
    const size_t elems = 8 * n + 16 * stride + 64 * n + 64 + 16 * (size_t)div_up(stride, 2048) + 4 * (size_t)div_up(stride, 2048);
    const size_t ntt_tmp = 6 * n8 * 32;
    size_t msm_ws = msm_workspace_bytes(P->srs, std::min(stride, srs_len(P->srs)), 4);
    if (P->srs_lag) msm_ws = std::max(msm_ws, msm_workspace_bytes(P->srs_lag, n + 4, 4));
    P->ws_bytes = elems * 32 + std::max(ntt_tmp, msm_ws) + (size_t)64 * 256 + (1 << 20);
  }
  return 0;
}

int prover_new(const uint8_t* label, size_t label_len, size_t constraints, const uint64_t* selectors,
               const uint32_t* wires, size_t n_witnesses, const uint8_t* srs_raw, size_t n_srs, pb200_prover** out) {
  if (constraints == 0) return fail(PB200_ERR_INVALID_ARG, "empty circuit");
  pb200_prover* P = new pb200_prover();
  const int rc = prover_build(P, label, label_len, constraints, selectors, wires, n_witnesses, srs_raw, n_srs, thread_stream());
  if (rc != 0) {
    prover_free(P);  // releases whatever had been allocated; the error message is already set
    return rc;
  }
  *out = P;
  return 0;
}

// Prover::try_from_bytes (src/compiler/prover.rs:265-350); the layout is spelled out at pb200_prover_from_bytes
// in include/plonk_b200.h.
int prover_from_bytes(const uint8_t* bytes, size_t len, const uint32_t* wires, size_t n_witnesses, pb200_prover** out) {
  auto be64 = [](const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
    return v;
  };
  auto le64 = [](const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
    return v;
  };
  const int short_rc = PB200_ERR_INVALID_ARG, bad_rc = PB200_ERR_POINT_MALFORMED;
  if (len < 48) return fail(short_rc, "NotEnoughBytes: serialized prover shorter than its header");
  const uint64_t label_len = be64(bytes), pk_len = be64(bytes + 8), ck_len = be64(bytes + 16), vk_len = be64(bytes + 24),
                 size = be64(bytes + 32), constraints = be64(bytes + 40);
  const uint8_t* p = bytes + 48;
  size_t left = len - 48;
  if (label_len > left || pk_len > left - label_len || ck_len > left - label_len - pk_len || vk_len > left - label_len - pk_len - ck_len)
    return fail(short_rc, "NotEnoughBytes: serialized prover shorter than its sections");
  size_t pow2 = 1;
  while (pow2 < constraints) pow2 <<= 1;
  if (constraints == 0 || pow2 != size) return fail(bad_rc, "InvalidData: size is not the next power of two of the constraint count");
  const uint8_t* label = p;
  const uint8_t* pk = label + label_len;
  const uint8_t* ck = pk + pk_len;
  const uint8_t* vk = ck + ck_len;
  // ProverKey::to_var_bytes (widget.rs:347-445); file order of the 15 polynomials -> this file's enum
  static const int file_order[N_POLY] = {Q_M, Q_L, Q_R, Q_O, Q_F, Q_C, Q_ARITH, Q_LOGIC, Q_RANGE, Q_FIXED, Q_VAR, S1, S2, S3, S4};
  LoadedProverKey key;
  {
    size_t off = 0;
    auto need = [&](size_t k) { return k <= pk_len - off; };
    if (pk_len < 16) return fail(short_rc, "NotEnoughBytes: prover key");
    const uint64_t n = le64(pk), eval_size = le64(pk + 8);
    off = 16;
    if (n != size) return fail(bad_rc, "InvalidData: prover key domain differs from the prover's size");
    if (eval_size != 8 * n * 32 + 172) return fail(bad_rc, "InvalidData: evaluations are not over the 8n domain");
    for (int i = 0; i < N_POLY; i++) {
      if (!need(8)) return fail(short_rc, "NotEnoughBytes: prover key polynomial header");
      const uint64_t cnt = le64(pk + off);
      off += 8;
      if (cnt > n) return fail(bad_rc, "InvalidData: polynomial longer than the domain");
      if (!need(cnt * 32)) return fail(short_rc, "NotEnoughBytes: prover key polynomial");
      key.poly[file_order[i]] = pk + off;
      key.poly_len[file_order[i]] = (size_t)cnt;
      off += cnt * 32;
      if (!need(eval_size)) return fail(short_rc, "NotEnoughBytes: prover key evaluations");
      off += eval_size;  // recomputed on the device
    }
    if (!need(2 * eval_size)) return fail(short_rc, "NotEnoughBytes: linear / vanishing evaluations");
  }
  // VerifierKey::to_bytes (widget.rs:84-111)
  if (vk_len < 8 + 15 * 48) return fail(short_rc, "NotEnoughBytes: verifier key");
  if (le64(vk) != size) return fail(bad_rc, "InvalidData: verifier key domain differs from the prover's size");
  for (int i = 0; i < N_POLY; i++) memcpy(key.comm[file_order[i]], vk + 8 + 48 * i, 48);
  // CommitKey::from_raw_var_bytes: validated points
  size_t n_pts = 0;
  PB_TRY(raw_commit_key_parse(ck, ck_len, 1, &n_pts, nullptr));
  std::vector<uint8_t> raw(n_pts * 96);
  PB_TRY(raw_commit_key_parse(ck, ck_len, 1, &n_pts, raw.data()));
  PB_TRY(g1_check_raw(raw.data(), n_pts));
  pb200_prover* P = new pb200_prover();
  const int rc = prover_build(P, label, label_len, constraints, nullptr, wires, n_witnesses, raw.data(), n_pts, thread_stream(), &key);
  if (rc != 0) {
    prover_free(P);
    return rc;
  }
  *out = P;
  return 0;
}

void prover_free(pb200_prover* P) {
  if (!P) return;
  if (P->srs) srs_free(P->srs);
  if (P->srs_lag) srs_free(P->srs_lag);
  cudaFree(P->d_wires); cudaFree(P->d_polys); cudaFree(P->d_key8); cudaFree(P->d_linear8); cudaFree(P->d_l1_8); cudaFree(P->d_sigma);
  for (char* w : P->ws_all) cudaFree(w);
  delete P;
}

static pbh::Transcript base_transcript(const pb200_prover* P) {  // transcript.rs:131-145, widget.rs:218-257
  pbh::Transcript tr(P->label.data(), P->label.size());
  tr.circuit_domain_sep(P->constraints);
  static const char* lbl[N_POLY] = {"q_m", "q_l", "q_r", "q_o", "q_c", "q_f", "q_arith", "q_range", "q_logic",
                                    "q_variable_group_add", "q_fixed_group_add", "s_sigma_1", "s_sigma_2", "s_sigma_3", "s_sigma_4"};
  static const int ord[N_POLY] = {Q_M, Q_L, Q_R, Q_O, Q_C, Q_F, Q_ARITH, Q_RANGE, Q_LOGIC, Q_VAR, Q_FIXED, S1, S2, S3, S4};
  for (int i = 0; i < N_POLY; i++) tr.append_commitment(lbl[i], P->comm[ord[i]]);
  tr.circuit_domain_sep(P->constraints);
  return tr;
}

// Prover::prove_inner, V3.  d_wit: n_witnesses Fr on the device; pi_*: host.
int prove_dev(const pb200_prover* P, const uint64_t* d_wit, const uint64_t* pi_idx, const uint64_t* pi_vals, size_t n_pi,
              const uint64_t* blinders_host, uint8_t* out_proof, cudaStream_t st) {
  const size_t n = P->n, n8 = P->n8, stride = n + 8;
  const int log_n = P->log_n;
  // With several proofs in flight the dense MSMs give up their latency-oriented bucket splitting (msm.cu)
  struct InFlight {
    const pb200_prover* P;
    explicit InFlight(const pb200_prover* p) : P(p) { P->active.fetch_add(1, std::memory_order_relaxed); }
    ~InFlight() {
      P->active.fetch_sub(1, std::memory_order_relaxed);
      t_msm_throughput_hint = 0;
    }
  } in_flight(P);
  static const int hint_at = [] {  // PB200_THROUGHPUT_AT=<k>: proofs in flight from which the hint is given (0 = never)
    const char* e = getenv("PB200_THROUGHPUT_AT");
    return e ? atoi(e) : 2;  // measured on BenchCircuit<2^16>, 12 in flight: never 193.5, from 4: 204.9, from 2: 208.7 proofs/s
  }();
  const HFr* BL = (const HFr*)blinders_host;
  pbh::Transcript tr = base_transcript(P);
  const HFr* PIV = (const HFr*)pi_vals;
  if (n_pi && (!pi_idx || !pi_vals)) return fail(PB200_ERR_INVALID_ARG, "public inputs announced but not given");
  for (size_t i = 0; i < n_pi; i++) {
    if (pi_idx[i] >= P->constraints) return fail(PB200_ERR_INVALID_ARG, "public input index out of range");
    // the reference keeps public inputs in a BTreeMap keyed by gate index (composer.rs:465-480): ascending, no duplicates
    if (i && pi_idx[i] <= pi_idx[i - 1]) return fail(PB200_ERR_INVALID_ARG, "public input positions must be strictly increasing");
    tr.append_scalar("pi", PIV[i]);
  }
  const uint4* w_half = nullptr;
  PB_TRY(get_twiddles(log_n, false, st, &w_half));

  // workspace: one arena per proof in flight
  Arena arena;
  {
    std::lock_guard<std::mutex> lk(P->ws_mu);
    if (!P->ws_free.empty()) {
      arena = P->ws_free.back();
      P->ws_free.pop_back();
    }
  }
  if (!arena.base) {
    char* mem = nullptr;
    PB_CUDA(cudaMalloc((void**)&mem, P->ws_bytes));
    arena.base = mem;
    arena.size = P->ws_bytes;
    std::lock_guard<std::mutex> lk(P->ws_mu);
    P->ws_all.push_back(mem);
  }
  arena.off = 0;
  struct Release {
    const pb200_prover* P;
    Arena* a;
    cudaStream_t st;
    ~Release() {
      stream_wait(st);  // error paths may leave work in flight
      a->off = 0;
      std::lock_guard<std::mutex> lk(P->ws_mu);
      P->ws_free.push_back(*a);
    }
  } release{P, &arena, st};
  Arena* ar = &arena;
  ScratchScope scope(ar, st);
  uint4 *wv, *wp, *zp, *num, *den, *w8, *quot, *tcoef, *tq, *pi_dense, *agg, *pw, *scratch, *evals_d, *partial;
  unsigned* flag;
  const unsigned eval_blocks = div_up(stride, 2048);
  PB_ALLOC(scope, wv, 5 * n * 32);       // wire values a, b, c, d and the dense public-input vector
  PB_ALLOC(scope, zp, 6 * stride * 32);  // [z, a, b, c, d, pi] coefficient form: one coset-NTT batch in round 3
  wp = zp + 2 * stride;
  PB_ALLOC(scope, num, n * 32);
  PB_ALLOC(scope, den, n * 32);
  PB_ALLOC(scope, w8, 6 * n8 * 32);
  PB_ALLOC(scope, quot, n8 * 32);
  PB_ALLOC(scope, tcoef, n8 * 32);
  PB_ALLOC(scope, tq, 4 * stride * 32);
  pi_dense = wv + 2 * 4 * n;
  PB_ALLOC(scope, agg, 2 * stride * 32);
  PB_ALLOC(scope, pw, 2 * stride * 32);
  PB_ALLOC(scope, scratch, 2 * stride * 32);
  PB_ALLOC(scope, evals_d, 16 * 32);
  PB_ALLOC(scope, partial, (size_t)16 * eval_blocks * 32);
  PB_ALLOC(scope, flag, 4);

  uint64_t aff[4 * 12];
  uint8_t c48[11][48];

  // ---- round 1 -------------------------------------------------------------------------------
  PB_LAUNCH(k_gather_wires, dim3(div_up(n, 256), 4), 256, 0, st, (const uint4*)d_wit, (const uint32_t*)P->d_wires, P->constraints, n, wv);
  // dense public-input vector (prover.rs:434-438) rides along with the wire iNTTs (prover.rs:519-521)
  if (n_pi) {
    PB_CUDA(cudaMemsetAsync(pi_dense, 0, n * 32, st));
    for (size_t i = 0; i < n_pi; i++)
      PB_CUDA(cudaMemcpyAsync(pi_dense + 2 * pi_idx[i], PIV + i, 32, cudaMemcpyHostToDevice, st));
  }
  PB_CUDA(cudaMemsetAsync(zp, 0, 6 * stride * 32, st));
  PB_TRY(ntt_run((const uint64_t*)wv, n, (uint64_t*)wp, log_n, 1, 0, n_pi ? 5 : 4, n, stride, st, ar));
  {
    BlindArgs ba;
    ba.nb = 2;
    ba.npoly = 4;
    for (int p = 0; p < 4; p++)
      for (int i = 0; i < 2; i++) ba.b[p][i] = to_dev(BL[2 * p + i]);
    PB_LAUNCH(k_blind, 1, 32, 0, st, wp, stride, n, ba);
  }
  if (P->srs_lag) {
    // the same four group elements from the wire VALUES: sum_i w_i [L_i(x)]G + b_0 ([x^n]G - [1]G) +
    // b_1 ([x^(n+1)]G - [x]G); w8 is free until round 3 and stages the scalars
    uint4* sc = w8;
    PB_CUDA(cudaMemcpy2DAsync(sc, stride * 32, wv, n * 32, n * 32, 4, cudaMemcpyDeviceToDevice, st));
    BlindArgs ba;
    ba.nb = 2;
    ba.npoly = 4;
    for (int p = 0; p < 4; p++)
      for (int i = 0; i < 2; i++) ba.b[p][i] = to_dev(BL[2 * p + i]);
    PB_LAUNCH(k_lagrange_tail, 1, 32, 0, st, sc, stride, n, ba);
    t_msm_throughput_hint = 0;  // sparse scalars: long buckets want their lanes
    PB_TRY(msm_run(P->srs_lag, 0, (const uint64_t*)sc, n + 4, 4, stride, aff, st, ar));
  } else {
    t_msm_throughput_hint = (hint_at > 0 && (g_force_throughput.load(std::memory_order_relaxed) || P->active.load(std::memory_order_relaxed) >= hint_at)) ? 1 : 0;
    PB_TRY(msm_run(P->srs, 0, (const uint64_t*)wp, n + 2, 4, stride, aff, st, ar));
  }
  for (int k = 0; k < 4; k++) compress_affine(aff + 12 * k, c48[k]);
  tr.append_commitment("a_comm", c48[0]);
  tr.append_commitment("b_comm", c48[1]);
  tr.append_commitment("c_comm", c48[2]);
  tr.append_commitment("d_comm", c48[3]);

  // ---- round 2 -------------------------------------------------------------------------------
  const HFr beta = tr.challenge_scalar("beta");
  tr.append_scalar("beta", beta);
  const HFr gamma = tr.challenge_scalar("gamma");
  PB_LAUNCH(k_perm_terms, div_up(n, 128), 128, 0, st, (const uint4*)wv, (const uint4*)P->d_sigma, w_half, n, to_dev(beta), to_dev(gamma), num, den);
  PB_LAUNCH(k_batch_div, div_up(div_up(n, 8), 128), 128, 0, st, (const uint4*)num, (const uint4*)den, n, num);
  PB_TRY((fr_scan<true, false>(num, n, den, st, ar)));  // den <- permutation vector z[i] = prod_{j<i} num_j/den_j
  PB_TRY(ntt_run((const uint64_t*)den, n, (uint64_t*)zp, log_n, 1, 0, 1, n, stride, st, ar));
  {
    BlindArgs ba;
    ba.nb = 3;
    ba.npoly = 1;
    for (int i = 0; i < 3; i++) ba.b[0][i] = to_dev(BL[8 + i]);
    PB_LAUNCH(k_blind, 1, 32, 0, st, zp, stride, n, ba);
  }
  t_msm_throughput_hint = (hint_at > 0 && (g_force_throughput.load(std::memory_order_relaxed) || P->active.load(std::memory_order_relaxed) >= hint_at)) ? 1 : 0;
  PB_TRY(msm_run(P->srs, 0, (const uint64_t*)zp, n + 3, 1, stride, aff, st, ar));
  compress_affine(aff, c48[4]);
  tr.append_commitment("z_comm", c48[4]);

  // ---- round 3 -------------------------------------------------------------------------------
  const HFr alpha = tr.challenge_scalar("alpha");
  const HFr ch_range = tr.challenge_scalar("range separation challenge");
  const HFr ch_logic = tr.challenge_scalar("logic separation challenge");
  const HFr ch_fixed = tr.challenge_scalar("fixed base separation challenge");
  const HFr ch_var = tr.challenge_scalar("variable base separation challenge");
  // t(X) has at most 4n + 7 coefficients, so the six coset transforms, the pointwise pass and the
  // inverse transform run on the 4n coset (the even points of the 8n one) and the top seven coefficients
  // are recovered from eight further points; the algebra is validated against the oracle in
  // tests/models/quotient_4n_model.py.  PB200_QUOT4N=0 restores the reference's 8n schedule
  // (quotient_poly.rs:50-137), which the parity suite keeps running as well.
  static const bool quot4n_env = [] {
    const char* e = getenv("PB200_QUOT4N");
    return !e || atoi(e) != 0;
  }();
  const bool quot4n = quot4n_env && n >= 16;  // the small arrays of step 2 live in the upper half of w8
  const size_t n4 = 4 * n;
  size_t t_len = n8;  // coefficients of t(X) present in tcoef
  QuotArgs q;
  {
    q.w8 = w8; q.key8 = P->d_key8; q.linear8 = P->d_linear8; q.l1_8 = P->d_l1_8; q.out = quot; q.n8 = n8;
    q.alpha = to_dev(alpha); q.beta = to_dev(beta); q.gamma = to_dev(gamma); q.alpha_sq = to_dev(alpha.sqr());
    auto powers = [](const HFr& ch) {
      const HFr k = ch.sqr(), k2 = k.sqr(), k3 = k2 * k;
      WidgetCh w;
      w.ch = to_dev(ch); w.k = to_dev(k); w.k2 = to_dev(k2); w.k3 = to_dev(k3); w.k4 = to_dev(k3 * k);
      return w;
    };
    q.ch_range = powers(ch_range); q.ch_logic = powers(ch_logic); q.ch_fixed = powers(ch_fixed); q.ch_var = powers(ch_var);
    q.edwards_d = P->edwards_d;
    for (int i = 0; i < 8; i++) q.vh_inv[i] = to_dev(P->vh_inv[i]);
    q.has_range = P->has_widget[0]; q.has_logic = P->has_widget[1]; q.has_fixed = P->has_widget[2]; q.has_var = P->has_widget[3];
  }
  unsigned char* stage = (unsigned char*)pinned_scratch(1024, 1);
  if (!stage) return fail(PB200_ERR_CUDA, "pinned staging buffer");
  volatile unsigned& h_flag = *(volatile unsigned*)(stage + 512);
  if (!quot4n) {
    // coset evaluations of z, a, b, c, d and the public-input polynomial in one batch
    // (quotient_poly.rs:50-59, 177)
    PB_TRY(ntt_run((const uint64_t*)zp, n + 3, (uint64_t*)w8, log_n + 3, 0, 1, n_pi ? 6 : 5, stride, n8, st, ar));
    if (!n_pi) PB_CUDA(cudaMemsetAsync(w8 + 2 * 5 * n8, 0, n8 * 32, st));  // empty PI polynomial: 8n zeros
    PB_LAUNCH(k_quotient, div_up(n8, 128), 128, 0, st, q);
    PB_TRY(ntt_run((const uint64_t*)quot, n8, (uint64_t*)tcoef, log_n + 3, 1, 1, 1, n8, n8, st, ar));
    // quotient_poly.len() > 7n  =>  CircuitUnsatisfied (quotient_poly.rs:132-134); coefficients past
    // 4n+7 cannot be committed with this key either
    PB_CUDA(cudaMemsetAsync(flag, 0, 4, st));
    PB_LAUNCH(k_any_nonzero, div_up(n8 - 7 * n, 256), 256, 0, st, (const uint4*)tcoef, 7 * n, n8, flag);
    PB_CUDA(cudaMemcpyAsync(stage + 512, flag, 4, cudaMemcpyDeviceToHost, st));
  } else {
    // 1. u(X) = t(X) mod (X^4n - g^4n) from the 4n coset
    PB_TRY(ntt_run((const uint64_t*)zp, n + 3, (uint64_t*)w8, log_n + 2, 0, 1, n_pi ? 6 : 5, stride, n4, st, ar));
    if (!n_pi) PB_CUDA(cudaMemsetAsync(w8 + 2 * 5 * n4, 0, n4 * 32, st));
    static const int quot_occ = [] {
      const char* e = getenv("PB200_QUOT_OCC");
      return e ? atoi(e) : 3;  // measured: 195.9 proofs/s at 3 CTAs/SM (168 registers, ~300 B spilled) against 194.0 at 2
    }();
    if (quot_occ == 3)
      PB_LAUNCH(k_quotient_4n<3>, div_up(n4, 128), 128, 0, st, q);
    else
      PB_LAUNCH(k_quotient_4n<2>, div_up(n4, 128), 128, 0, st, q);
    PB_TRY(ntt_run((const uint64_t*)quot, n4, (uint64_t*)tcoef, log_n + 2, 1, 1, 1, n4, n4, st, ar));
    // 2. t at the eight points x_k = h w8^k, h = g w_8n (indices 1 + n k of the 8n coset): the witness
    //    polynomials by Horner at x_k and omega x_k, the prover key from its 8n tables.  The upper half
    //    of w8 is free in this mode and holds the small arrays.
    uint4* wpts = w8 + 2 * 6 * n4;   // [6][16]
    uint4* tx = wpts + 2 * 96;       // t(x_k)
    uint4* ux = tx + 2 * 8;          // u(x_k)
    uint4* part4 = ux + 2 * 8;       // partial sums of the evaluations
    const unsigned blocks4 = div_up(n4, 2048);
    const Quot4nConsts& qc = P->q4;
    const int rows = n_pi ? 6 : 5;
    {
      EvalRows er;
      er.base = zp; er.row_stride = stride; er.len = (unsigned)n + 3;
      for (int j = 0; j < 16; j++) er.point[j] = qc.xs[j];
      EvalPoints ep;
      for (int j = 0; j < 16; j++) ep.p[j] = qc.xs[j];
      PB_LAUNCH(k_poly_eval_rows, dim3(eval_blocks, 16, rows), 256, 0, st, er, part4, eval_blocks);
      PB_LAUNCH(k_sum_rows, 16 * rows, 32, 0, st, (const uint4*)part4, eval_blocks, ep, 16, wpts);
    }
    if (!n_pi) PB_CUDA(cudaMemsetAsync(wpts + 2 * 16 * 5, 0, 16 * 32, st));
    QuotArgs qp = q;
    qp.w8 = wpts;
    qp.out = tx;
    PB_LAUNCH(k_quotient_pts, 1, 128, 0, st, qp);
    {
      EvalRows er;
      er.base = tcoef; er.row_stride = 0; er.len = (unsigned)n4;
      for (int j = 0; j < 16; j++) er.point[j] = qc.xs[j];
      EvalPoints ep;
      for (int j = 0; j < 16; j++) ep.p[j] = qc.xs[j];
      PB_LAUNCH(k_poly_eval_rows, dim3(blocks4, 8, 1), 256, 0, st, er, part4, blocks4);
      PB_LAUNCH(k_sum_rows, 8, 32, 0, st, (const uint4*)part4, blocks4, ep, 16, ux);
    }
    // 3. t_hi(x_k) = (t(x_k) - u(x_k)) / (x_k^4n - g^4n), and x_k^4n = -g^4n for every k; the 8-point
    //    inverse DFT on h*H_8 gives its coefficients, the eighth of which must vanish (N divisible by Z_H:
    //    replaces the reference's len > 7n test, quotient_poly.rs:132-134);  4. t = (u - g^4n t_hi) + X^4n t_hi.
    //    Both on the device (eight threads), so round 3 has no host synchronisation of its own.
    PB_CUDA(cudaMemsetAsync(flag, 0, 4, st));
    PB_LAUNCH(k_quot4n_fix, 1, 32, 0, st, (const uint4*)tx, (const uint4*)ux, tcoef, n4, qc.fix, flag);
    PB_CUDA(cudaMemcpyAsync(stage + 512, flag, 4, cudaMemcpyDeviceToHost, st));
    t_len = n4 + 7;
  }
  PB_LAUNCH(k_split_quotient, dim3(div_up(stride, 256), 4), 256, 0, st, (const uint4*)tcoef, n, t_len, stride, to_dev(BL[11]), to_dev(BL[12]), to_dev(BL[13]), tq);
  const size_t key_len = srs_len(P->srs);
  const size_t tlen = std::min(stride, key_len);
  t_msm_throughput_hint = (hint_at > 0 && (g_force_throughput.load(std::memory_order_relaxed) || P->active.load(std::memory_order_relaxed) >= hint_at)) ? 1 : 0;
  PB_TRY(msm_run(P->srs, 0, (const uint64_t*)tq, tlen, 4, stride, aff, st, ar));  // synchronises the stream
  if (h_flag) return fail(PB200_ERR_UNSATISFIED, "CircuitUnsatisfied");
  for (int k = 0; k < 4; k++) compress_affine(aff + 12 * k, c48[5 + k]);
  tr.append_commitment("t_low_comm", c48[5]);
  tr.append_commitment("t_mid_comm", c48[6]);
  tr.append_commitment("t_high_comm", c48[7]);
  tr.append_commitment("t_fourth_comm", c48[8]);

  // ---- round 4 -------------------------------------------------------------------------------
  const HFr z_ch = tr.challenge_scalar("z_challenge");
  const HFr zw = z_ch * to_host(ntt_group_gen(log_n, false));
  enum { E_A, E_B, E_C, E_D, E_AW, E_BW, E_DW, E_QARITH, E_QC, E_QL, E_QR, E_S1, E_S2, E_S3, E_Z };
  HFr ev[15];
  {
    EvalJobs jobs;
    jobs.njobs = 15;
    auto poly = [&](int k) { return (const uint4*)(P->d_polys + 2 * (size_t)k * n); };
    const uint4* ps[15] = {wp, wp + 2 * stride, wp + 4 * stride, wp + 6 * stride, wp, wp + 2 * stride, wp + 6 * stride,
                           poly(Q_ARITH), poly(Q_C), poly(Q_L), poly(Q_R), poly(S1), poly(S2), poly(S3), zp};
    const unsigned ls[15] = {(unsigned)n + 2, (unsigned)n + 2, (unsigned)n + 2, (unsigned)n + 2, (unsigned)n + 2, (unsigned)n + 2, (unsigned)n + 2,
                             (unsigned)n, (unsigned)n, (unsigned)n, (unsigned)n, (unsigned)n, (unsigned)n, (unsigned)n, (unsigned)n + 3};
    const bool shifted[15] = {0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 1};
    for (int j = 0; j < 15; j++) {
      jobs.poly[j] = ps[j];
      jobs.len[j] = ls[j];
      jobs.point[j] = to_dev(shifted[j] ? zw : z_ch);
    }
    EvalPoints ep;
    for (int j = 0; j < 15; j++) ep.p[j] = jobs.point[j];
    ep.p[15] = Fr::zero();
    PB_LAUNCH(k_poly_eval, dim3(eval_blocks, 15), 256, 0, st, jobs, partial, eval_blocks);
    PB_LAUNCH(k_sum_rows, 15, 32, 0, st, (const uint4*)partial, eval_blocks, ep, 16, evals_d);
    PB_CUDA(cudaMemcpyAsync(stage, evals_d, 15 * 32, cudaMemcpyDeviceToHost, st));
    PB_CUDA(stream_wait(st));
    memcpy(ev, stage, 15 * 32);
  }
  tr.append_scalar("a_eval", ev[E_A]); tr.append_scalar("b_eval", ev[E_B]); tr.append_scalar("c_eval", ev[E_C]); tr.append_scalar("d_eval", ev[E_D]);
  tr.append_scalar("s_sigma_1_eval", ev[E_S1]); tr.append_scalar("s_sigma_2_eval", ev[E_S2]); tr.append_scalar("s_sigma_3_eval", ev[E_S3]);
  tr.append_scalar("z_eval", ev[E_Z]);
  tr.append_scalar("a_w_eval", ev[E_AW]); tr.append_scalar("b_w_eval", ev[E_BW]); tr.append_scalar("d_w_eval", ev[E_DW]);
  tr.append_scalar("q_arith_eval", ev[E_QARITH]); tr.append_scalar("q_c_eval", ev[E_QC]); tr.append_scalar("q_l_eval", ev[E_QL]); tr.append_scalar("q_r_eval", ev[E_QR]);

  // ---- round 5 -------------------------------------------------------------------------------
  const HFr v = tr.challenge_scalar("v_challenge");
  const HFr v_w = tr.challenge_scalar("v_w_challenge");  // nothing is appended in between (prover.rs:680-730)
  {
    // host-side widget scalars of the linearisation polynomial (all widget compute_linearization)
    auto h4 = [](const HFr& x) { return x.dbl().dbl(); };
    auto delta = [](const HFr& f) { HFr o = HFr::one(); HFr f1 = f - o, f2 = f1 - o, f3 = f2 - o; return f * f1 * f2 * f3; };
    auto small = [](uint64_t k) { return HFr::from_u64(k); };
    const HFr &a = ev[E_A], &b = ev[E_B], &c = ev[E_C], &d = ev[E_D], &a_w = ev[E_AW], &b_w = ev[E_BW], &d_w = ev[E_DW];
    HFr s_range, s_logic, s_fixed, s_var;
    {
      HFr k = ch_range.sqr(), k2 = k.sqr(), k3 = k2 * k;
      s_range = (delta(c - h4(d)) + delta(b - h4(c)) * k + delta(a - h4(b)) * k2 + delta(d_w - h4(a)) * k3) * ch_range;
    }
    {
      HFr k = ch_logic.sqr(), k2 = k.sqr(), k3 = k2 * k, k4 = k3 * k;
      HFr A = a_w - h4(a), B = b_w - h4(b), D = d_w - h4(d);
      const HFr& w = c;
      HFr F = w * (w * (h4(w) - small(18) * (A + B) + small(81)) + small(18) * (A.sqr() + B.sqr()) - small(81) * (A + B) + small(83));
      HFr E = small(3) * (A + B + D) - F.dbl();
      HFr Bq = ev[E_QC] * (small(9) * D - small(3) * (A + B));
      s_logic = (delta(A) + delta(B) * k + delta(D) * k2 + (w - A * B) * k3 + (Bq + E) * k4) * ch_logic;
    }
    const HFr ed = (small(10240) * small(10241).inv()).neg();
    {
      HFr one = HFr::one(), k = ch_fixed.sqr(), k2 = k.sqr(), k3 = k2 * k;
      HFr bit = d_w - d - d;
      HFr bit_c = bit * (bit - one) * (bit + one);
      HFr y_alpha = bit.sqr() * (ev[E_QR] - one) + one, x_alpha = bit * ev[E_QL];
      HFr xy = (bit * ev[E_QC] - c) * k;
      HFr t = c * a * b * ed;
      HFr xa = ((a_w + a_w * t) - (a * y_alpha + b * x_alpha)) * k2;
      HFr ya = ((b_w - b_w * t) - (b * y_alpha + a * x_alpha)) * k3;
      s_fixed = (bit_c + xa + ya + xy) * ch_fixed;
    }
    {
      HFr k = ch_var.sqr();
      HFr xy = a * d - d_w, y1x2 = b * c, y1y2 = b * d, x1x2 = a * c;
      HFr t = ed * d_w * y1x2;
      HFr x3c = ((d_w + y1x2) - (a_w + a_w * t)) * k;
      HFr y3c = ((y1y2 + x1x2) - (b_w - b_w * t)) * k.sqr();
      s_var = (xy + x3c + y3c) * ch_var;
    }
    const HFr bz = beta * z_ch;
    const HFr s_ident = (a + bz + gamma) * (b + small(7) * bz + gamma) * (c + small(13) * bz + gamma) * (d + small(17) * bz + gamma) * alpha;
    const HFr s_copy = (a + beta * ev[E_S1] + gamma) * (b + beta * ev[E_S2] + gamma) * (c + beta * ev[E_S3] + gamma) * (beta * ev[E_Z]) * alpha;
    const HFr z_n = hfr_pow(z_ch, n);
    // domain of z_poly.degree() - 2 is the proving domain n (permutation/proverkey.rs:156-163)
    const HFr l1_z = (z_n - HFr::one()) * to_host(ntt_size_inv(log_n)) * (z_ch - HFr::one()).inv();
    const HFr zh = (z_n - HFr::one()).neg();
    HFr vp[12];
    vp[0] = HFr::one();
    for (int i = 1; i < 12; i++) vp[i] = vp[i - 1] * v;
    auto poly = [&](int k) { return (const uint4*)(P->d_polys + 2 * (size_t)k * n); };
    const HFr qa = ev[E_QARITH];
    // W_z numerator: r + v a + v^2 b + v^3 c + v^4 d + v^5 s1 + v^6 s2 + v^7 s3 + v^8 q_arith + v^9 q_c + v^10 q_l + v^11 q_r
    LinArgs la;
    int t = 0;
    auto term = [&](const uint4* p, size_t len, const HFr& coef) { la.poly[t] = p; la.len[t] = (unsigned)len; la.coef[t] = to_dev(coef); t++; };
    term(poly(Q_M), n, a * b * qa);
    term(poly(Q_L), n, a * qa + vp[10]);
    term(poly(Q_R), n, b * qa + vp[11]);
    term(poly(Q_O), n, c * qa);
    term(poly(Q_F), n, d * qa);
    term(poly(Q_C), n, qa + vp[9]);
    term(poly(Q_ARITH), n, vp[8]);
    term(poly(Q_RANGE), n, s_range);
    term(poly(Q_LOGIC), n, s_logic);
    term(poly(Q_FIXED), n, s_fixed);
    term(poly(Q_VAR), n, s_var);
    term(zp, n + 3, s_ident + l1_z * alpha.sqr());
    term(poly(S4), n, s_copy.neg());
    term(tq, stride, zh);
    term(tq + 2 * stride, stride, zh * z_n);
    term(tq + 4 * stride, stride, zh * z_n.sqr());
    term(tq + 6 * stride, stride, zh * z_n.sqr() * z_n);
    term(wp, n + 2, vp[1]);
    term(wp + 2 * stride, n + 2, vp[2]);
    term(wp + 4 * stride, n + 2, vp[3]);
    term(wp + 6 * stride, n + 2, vp[4]);
    term(poly(S1), n, vp[5]);
    term(poly(S2), n, vp[6]);
    term(poly(S3), n, vp[7]);
    la.nterms = t;
    PB_LAUNCH(k_lincomb, div_up(stride, 128), 128, 0, st, la, stride, agg);
    // W_zw numerator: z + v_w a + v_w^2 b + v_w^3 d
    LinArgs lb;
    t = 0;
    auto term2 = [&](const uint4* p, size_t len, const HFr& coef) { lb.poly[t] = p; lb.len[t] = (unsigned)len; lb.coef[t] = to_dev(coef); t++; };
    term2(zp, n + 3, HFr::one());
    term2(wp, n + 2, v_w);
    term2(wp + 2 * stride, n + 2, v_w.sqr());
    term2(wp + 6 * stride, n + 2, v_w.sqr() * v_w);
    lb.nterms = t;
    PB_LAUNCH(k_lincomb, div_up(stride, 128), 128, 0, st, lb, stride, agg + 2 * stride);
    // ruffini (polynomial.rs:345-367): q_j = z^-(j+1) * sum_{i>j} c_i z^i
    const HFr pts[2] = {z_ch, zw};
    for (int w = 0; w < 2; w++) {
      uint4* c_w = agg + 2 * (size_t)w * stride;
      uint4* pw_w = pw + 2 * (size_t)w * stride;
      uint4* sc_w = scratch + 2 * (size_t)w * stride;
      PB_TRY(fill_powers(pw_w, stride, to_dev(pts[w]), Fr::one(), st));
      PB_LAUNCH(k_mul_pointwise, div_up(stride, 128), 128, 0, st, (const uint4*)c_w, (const uint4*)pw_w, stride, sc_w);
      PB_TRY((fr_scan<false, true>(sc_w, stride, c_w, st, ar)));  // exclusive suffix sums
      const HFr zi = pts[w].inv();
      PB_TRY(fill_powers(pw_w, stride, to_dev(zi), to_dev(zi), st));
      PB_LAUNCH(k_mul_pointwise, div_up(stride, 128), 128, 0, st, (const uint4*)c_w, (const uint4*)pw_w, stride, c_w);
    }
    const size_t wlen = std::min(stride, key_len);
    t_msm_throughput_hint = (hint_at > 0 && (g_force_throughput.load(std::memory_order_relaxed) || P->active.load(std::memory_order_relaxed) >= hint_at)) ? 1 : 0;
    PB_TRY(msm_run(P->srs, 0, (const uint64_t*)agg, wlen, 2, stride, aff, st, ar));
    compress_affine(aff, c48[9]);
    compress_affine(aff + 12, c48[10]);
  }
  // Proof::to_bytes (proof.rs:137-162, linearization_poly.rs:98-124)
  for (int i = 0; i < 11; i++) memcpy(out_proof + 48 * i, c48[i], 48);
  for (int i = 0; i < 15; i++) {
    HFr cnon = ev[i].from_mont();
    memcpy(out_proof + 528 + 32 * i, cnon.v, 32);
  }
  return 0;
}

}  // namespace pb

using namespace pb;

extern "C" {

int pb200_prover_new(const uint8_t* label, size_t label_len, size_t n_constraints, const uint64_t* selectors,
                     const uint32_t* wires, size_t n_witnesses, const uint8_t* srs_raw, size_t n_srs_points,
                     pb200_prover_t** out) {
  PB_TRY(ensure_init());
  if (!selectors || !wires || !srs_raw || !out) return fail(PB200_ERR_INVALID_ARG, "null argument");
  return prover_new(label, label_len, n_constraints, selectors, wires, n_witnesses, srs_raw, n_srs_points, out);
}

int pb200_throughput_mode(int on) {
  g_force_throughput.store(on ? 1 : 0, std::memory_order_relaxed);
  return 0;
}

int pb200_prover_from_bytes(const uint8_t* bytes, size_t len, const uint32_t* wires, size_t n_witnesses, pb200_prover_t** out) {
  PB_TRY(ensure_init());
  if (!bytes || !wires || !out) return fail(PB200_ERR_INVALID_ARG, "null argument");
  return prover_from_bytes(bytes, len, wires, n_witnesses, out);
}

void pb200_prover_free(pb200_prover_t* p) {
  if (!p) return;
  ensure_init();  // a thread that has made no other pb200 call yet must free on the library's device
  prover_free(p);
}

int pb200_prover_commitments(const pb200_prover_t* p, uint8_t* out /* 15 x 48 */) {
  if (!p || !out) return fail(PB200_ERR_INVALID_ARG, "null argument");
  memcpy(out, p->comm, sizeof p->comm);
  return 0;
}

int pb200_prove(const pb200_prover_t* p, const uint64_t* witnesses, size_t n_witnesses, const uint64_t* pi_idx,
                const uint64_t* pi_vals, size_t n_pi, const uint64_t* blinders, uint8_t* out_proof) {
  PB_TRY(ensure_init());
  if (!p || !witnesses || !blinders || !out_proof) return fail(PB200_ERR_INVALID_ARG, "null argument");
  if (n_witnesses != p->n_witnesses) return fail(PB200_ERR_INVALID_ARG, "witness count differs from the compiled circuit");
  cudaStream_t st = thread_stream();
  uint64_t* d_wit = nullptr;
  PB_CUDA(cudaMallocAsync((void**)&d_wit, n_witnesses * 32, st));
  int rc;
  const cudaError_t e = cudaMemcpyAsync(d_wit, witnesses, n_witnesses * 32, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess)
    rc = fail(PB200_ERR_CUDA, "witness upload", cudaGetErrorString(e));
  else
    rc = prove_dev(p, d_wit, pi_idx, pi_vals, n_pi, blinders, out_proof, st);
  cudaFreeAsync(d_wit, st);
  return rc;
}

int pb200_prove_dev(const pb200_prover_t* p, const uint64_t* d_witnesses, size_t n_witnesses, const uint64_t* pi_idx,
                    const uint64_t* pi_vals, size_t n_pi, const uint64_t* blinders, uint8_t* out_proof, void* stream) {
  PB_TRY(ensure_init());
  if (!p || !d_witnesses || !blinders || !out_proof) return fail(PB200_ERR_INVALID_ARG, "null argument");
  if (n_witnesses != p->n_witnesses) return fail(PB200_ERR_INVALID_ARG, "witness count differs from the compiled circuit");
  cudaStream_t st = stream ? (cudaStream_t)stream : thread_stream();
  return prove_dev(p, d_witnesses, pi_idx, pi_vals, n_pi, blinders, out_proof, st);
}
}
