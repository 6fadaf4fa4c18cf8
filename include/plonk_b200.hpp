// C++ mirror of the reference's interface for the hot path, over the C ABI in plonk_b200.h.
//
// The reference is a compiled (Rust) crate; this header is the compiled-language host side a C++
// caller links against, with the reference's names, argument meaning and error behaviour:
//
//   plonk_b200::EvaluationDomain   src/fft/domain.rs:35-232      new / fft / ifft / coset_fft / coset_ifft
//   plonk_b200::CommitKey          src/commitment_scheme/kzg10/key.rs:36-41, 362-388   commit, max_degree
//   plonk_b200::Commitment         src/commitment_scheme/kzg10/commitment.rs:77-106    to_bytes (48 B)
//   plonk_b200::Prover             src/compiler/prover.rs:53-115, 352-362              prove
//   plonk_b200::Composer           src/composer.rs:72-495 + src/composer/{bits,range,logic,truncate,select,
//                                  point,fixed_base}.rs (host-side circuit front end, plonk_b200_composer.h)
//   plonk_b200::Error              src/error.rs:21-120 (the variants this path can produce)
//
// BlsScalar is the reference's in-memory layout: 4 x u64 little-endian limbs, Montgomery form.
// Infallible reference functions (the NTT family asserts/panics, domain.rs:394) throw
// BackendFailure on a device error - there is no CPU fallback.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <memory>
#include <string>
#include <vector>

#include "plonk_b200.h"
#include "plonk_b200_composer.h"

namespace plonk_b200 {

using BlsScalar = std::array<uint64_t, 4>;

struct Error : std::runtime_error {
  enum Kind {
    InvalidEvalDomainSize, PolynomialDegreeTooLarge, CircuitUnsatisfied, InvalidArgument, BackendFailure,
    JubJubPointNotTorsionFree, JubJubGeneratorNotPrimeOrder, JubJubScalarMalformed
  };
  Kind kind;
  Error(Kind k, const std::string& what) : std::runtime_error(what), kind(k) {}
};

inline void check(int rc) {
  if (rc == PB200_OK) return;
  const std::string msg = pb200_last_error();
  switch (rc) {
    case PB200_ERR_INVALID_DOMAIN: throw Error(Error::InvalidEvalDomainSize, msg);
    case PB200_ERR_DEGREE_TOO_LARGE: throw Error(Error::PolynomialDegreeTooLarge, msg);
    case PB200_ERR_UNSATISFIED: throw Error(Error::CircuitUnsatisfied, msg);
    case PB200_ERR_INVALID_ARG: throw Error(Error::InvalidArgument, msg);
    case PB200_ERR_JUBJUB_POINT: throw Error(Error::JubJubPointNotTorsionFree, msg);
    case PB200_ERR_JUBJUB_GENERATOR: throw Error(Error::JubJubGeneratorNotPrimeOrder, msg);
    case PB200_ERR_JUBJUB_SCALAR: throw Error(Error::JubJubScalarMalformed, msg);
    default: throw Error(Error::BackendFailure, msg);
  }
}

class EvaluationDomain {
 public:
  // EvaluationDomain::new(num_coeffs): size = next power of two; log size must stay below TWO_ADACITY = 32.
  explicit EvaluationDomain(size_t num_coeffs) {
    size_ = 1;
    log_ = 0;
    while (size_ < num_coeffs) {
      size_ <<= 1;
      log_++;
    }
    if (log_ >= 32) throw Error(Error::InvalidEvalDomainSize, "log_size_of_group >= TWO_ADACITY");
  }
  size_t size() const { return size_; }
  std::vector<BlsScalar> fft(const std::vector<BlsScalar>& coeffs) const { return run(coeffs, 0, 0); }
  std::vector<BlsScalar> ifft(const std::vector<BlsScalar>& evals) const { return run(evals, 1, 0); }
  std::vector<BlsScalar> coset_fft(const std::vector<BlsScalar>& coeffs) const { return run(coeffs, 0, 1); }
  std::vector<BlsScalar> coset_ifft(const std::vector<BlsScalar>& evals) const { return run(evals, 1, 1); }

 private:
  size_t size_;
  uint32_t log_;
  std::vector<BlsScalar> run(const std::vector<BlsScalar>& v, int inverse, int coset) const {
    std::vector<BlsScalar> out(size_);
    check(pb200_ntt(v.empty() ? nullptr : v[0].data(), v.size(), out[0].data(), log_, inverse, coset, 1, v.size(), size_));
    return out;
  }
};

struct Commitment {
  std::array<uint64_t, 12> raw{};  // x, y Montgomery limbs; identity = zeros
  std::array<uint8_t, 48> to_bytes() const {
    std::array<uint8_t, 48> b;
    check(pb200_g1_compress(raw.data(), b.data()));
    return b;
  }
  bool operator==(const Commitment& o) const { return raw == o.raw; }
};

class CommitKey {
 public:
  // powers_of_g as 96-byte raw points (CommitKey::to_raw_var_bytes without length prefix / flags)
  CommitKey(const uint8_t* raw_points, size_t n_points) : n_(n_points) { check(pb200_srs_upload(raw_points, n_points, &h_)); }
  // CommitKey::from_raw_var_bytes (key.rs:258-298: every point validated, on the GPU) and
  // CommitKey::from_slice_unchecked (key.rs:242-256: trusted bytes) for CommitKey::to_raw_var_bytes
  static std::unique_ptr<CommitKey> from_raw_var_bytes(const uint8_t* bytes, size_t len) { return from_raw(bytes, len, 1); }
  static std::unique_ptr<CommitKey> from_slice_unchecked(const uint8_t* bytes, size_t len) { return from_raw(bytes, len, 0); }
  ~CommitKey() { pb200_srs_free(h_); }
  CommitKey(const CommitKey&) = delete;
  CommitKey& operator=(const CommitKey&) = delete;
  size_t max_degree() const { return n_ - 1; }
  // CommitKey::commit: trailing zero coefficients are trimmed first (Polynomial::from_coefficients_vec),
  // then Error::PolynomialDegreeTooLarge if degree > max_degree.
  Commitment commit(std::vector<BlsScalar> polynomial) const {
    while (!polynomial.empty() && polynomial.back() == BlsScalar{0, 0, 0, 0}) polynomial.pop_back();
    const size_t degree = polynomial.empty() ? 0 : polynomial.size() - 1;
    if (degree > max_degree()) throw Error(Error::PolynomialDegreeTooLarge, "polynomial degree exceeds the commit key");
    Commitment c;
    check(pb200_msm_g1(h_, polynomial.empty() ? nullptr : polynomial[0].data(), polynomial.size(), 1, polynomial.size(), c.raw.data()));
    return c;
  }

 private:
  pb200_srs_t* h_ = nullptr;
  size_t n_;
  static std::unique_ptr<CommitKey> from_raw(const uint8_t* bytes, size_t len, int checked) {
    size_t n = 0;
    check(pb200_raw_commit_key_points(bytes, len, checked, &n));
    std::vector<uint8_t> raw(n * 96 + 1);
    check(pb200_commit_key_from_raw_var_bytes(bytes, len, checked, raw.data()));
    return std::unique_ptr<CommitKey>(new CommitKey(raw.data(), n));
  }
};

// Flat circuit description (what Compiler::preprocess reads out of the Composer, compiler.rs:132-170)
struct Circuit {
  std::vector<BlsScalar> selectors;  // 11 columns x n_constraints, column-major (order of plonk_b200.h)
  std::vector<uint32_t> wires;       // 4 columns x n_constraints witness indices
  size_t n_constraints = 0;
  size_t n_witnesses = 0;
};

// ---- circuit front end ------------------------------------------------------------------------
typedef uint32_t Witness;
struct WitnessPoint {
  Witness x, y;
};
struct JubJubAffine {
  BlsScalar u, v;
  // dusk_jubjub::GENERATOR
  static JubJubAffine generator() {
    uint64_t uv[8];
    check(pb200_jubjub_generator(uv));
    return from_raw(uv);
  }
  // scalar: canonical little-endian limbs of a JubJubScalar
  JubJubAffine mul(const std::array<uint64_t, 4>& scalar) const {
    uint64_t in[8], out[8];
    to_raw(in);
    check(pb200_jubjub_mul(in, scalar.data(), out));
    return from_raw(out);
  }
  void to_raw(uint64_t uv[8]) const {
    for (int i = 0; i < 4; i++) uv[i] = u[i], uv[4 + i] = v[i];
  }
  static JubJubAffine from_raw(const uint64_t uv[8]) {
    JubJubAffine p;
    for (int i = 0; i < 4; i++) p.u[i] = uv[i], p.v[i] = uv[4 + i];
    return p;
  }
};

// One gate under construction (constraint.rs:97-230); coefficients default to zero, wires to
// Composer::ZERO.  Scalars are Montgomery-form BlsScalar values.
struct Constraint {
  std::array<BlsScalar, 11> q{};  // q_m, q_l, q_r, q_o, q_f, q_c, q_arith, q_range, q_logic, q_fixed_group_add, q_variable_group_add
  std::array<Witness, 4> w{};     // a, b, c, d
  BlsScalar pi{};
  bool has_pi = false;
  Constraint& mult(const BlsScalar& s) { q[0] = s; return *this; }
  Constraint& left(const BlsScalar& s) { q[1] = s; return *this; }
  Constraint& right(const BlsScalar& s) { q[2] = s; return *this; }
  Constraint& output(const BlsScalar& s) { q[3] = s; return *this; }
  Constraint& fourth(const BlsScalar& s) { q[4] = s; return *this; }
  Constraint& constant(const BlsScalar& s) { q[5] = s; return *this; }
  Constraint& pub(const BlsScalar& s) { pi = s; has_pi = true; return *this; }
  Constraint& a(Witness x) { w[0] = x; return *this; }
  Constraint& b(Witness x) { w[1] = x; return *this; }
  Constraint& c(Witness x) { w[2] = x; return *this; }
  Constraint& d(Witness x) { w[3] = x; return *this; }
};

// BlsScalar::from(u64), Montgomery form (x * 2^256 mod r computed by 256 modular doublings)
inline BlsScalar scalar_from_u64(uint64_t x) {
  static const uint64_t r[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
  uint64_t v[4] = {x, 0, 0, 0};
  for (int k = 0; k < 256; k++) {
    const uint64_t top = v[3] >> 63;
    for (int i = 3; i > 0; i--) v[i] = (v[i] << 1) | (v[i - 1] >> 63);
    v[0] <<= 1;
    bool ge = top != 0;
    if (!ge) {
      ge = true;
      for (int i = 3; i >= 0; i--)
        if (v[i] != r[i]) { ge = v[i] > r[i]; break; }
    }
    if (ge) {
      unsigned __int128 borrow = 0;
      for (int i = 0; i < 4; i++) {
        const unsigned __int128 d = (unsigned __int128)v[i] - r[i] - borrow;
        v[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
      }
    }
  }
  return {v[0], v[1], v[2], v[3]};
}

class Composer {
 public:
  static constexpr Witness ZERO = 0, ONE = 1;
  static constexpr WitnessPoint IDENTITY = {0, 1};

  Composer() { check(pb200_composer_new(&h_)); }  // Composer::initialized()
  ~Composer() { pb200_composer_free(h_); }
  Composer(const Composer&) = delete;
  Composer& operator=(const Composer&) = delete;

  size_t constraints() const { return pb200_composer_constraints(h_); }
  BlsScalar operator[](Witness w) const { BlsScalar v; check(pb200_composer_witness_value(h_, w, v.data())); return v; }

  Witness append_witness(const BlsScalar& v) { Witness w; check(pb200_composer_append_witness(h_, v.data(), &w)); return w; }
  Witness append_witness(uint64_t v) { return append_witness(scalar_from_u64(v)); }
  void append_gate(const Constraint& c) { check(pb200_composer_append_gate(h_, c.q[0].data(), c.w.data(), c.has_pi ? c.pi.data() : nullptr, 0)); }
  void append_custom_gate(const Constraint& c) { check(pb200_composer_append_gate(h_, c.q[0].data(), c.w.data(), c.has_pi ? c.pi.data() : nullptr, 1)); }
  Witness gate_add(const Constraint& c) { Witness w; check(pb200_composer_gate_add(h_, c.q[0].data(), c.w.data(), c.has_pi ? c.pi.data() : nullptr, &w)); return w; }
  Witness gate_mul(const Constraint& c) { return gate_add(c); }
  Witness append_constant(const BlsScalar& v) { Witness w; check(pb200_composer_append_constant(h_, v.data(), &w)); return w; }
  Witness append_constant(uint64_t v) { return append_constant(scalar_from_u64(v)); }
  Witness append_public(const BlsScalar& v) { Witness w; check(pb200_composer_append_public(h_, v.data(), &w)); return w; }
  void assert_equal(Witness a, Witness b) { check(pb200_composer_assert_equal(h_, a, b)); }
  void assert_equal_constant(Witness a, const BlsScalar& constant, const BlsScalar* pi = nullptr) {
    check(pb200_composer_assert_equal_constant(h_, a, constant.data(), pi ? pi->data() : nullptr));
  }
  void component_boolean(Witness a) { check(pb200_composer_component_boolean(h_, a)); }
  template <size_t N>
  std::array<Witness, N> component_decomposition(Witness scalar) {
    std::array<Witness, N> bits;
    check(pb200_composer_component_decomposition(h_, scalar, (uint32_t)N, bits.data()));
    return bits;
  }
  template <size_t BITS>
  void component_range_bits(Witness w) { check(pb200_composer_component_range_bits(h_, w, (uint32_t)BITS)); }
  template <size_t BIT_PAIRS>
  Witness append_logic_and(Witness a, Witness b) { Witness w; check(pb200_composer_append_logic(h_, a, b, (uint32_t)BIT_PAIRS, 0, &w)); return w; }
  template <size_t BIT_PAIRS>
  Witness append_logic_xor(Witness a, Witness b) { Witness w; check(pb200_composer_append_logic(h_, a, b, (uint32_t)BIT_PAIRS, 1, &w)); return w; }
  template <size_t N>
  Witness component_truncate(Witness x) { Witness w; check(pb200_composer_component_truncate(h_, x, (uint32_t)N, &w)); return w; }
  Witness component_select(Witness bit, Witness a, Witness b) { Witness w; check(pb200_composer_component_select(h_, bit, a, b, &w)); return w; }
  Witness component_select_one(Witness bit, Witness v) { Witness w; check(pb200_composer_component_select_one(h_, bit, v, &w)); return w; }
  Witness component_select_zero(Witness bit, Witness v) { Witness w; check(pb200_composer_component_select_zero(h_, bit, v, &w)); return w; }

  WitnessPoint append_point(const JubJubAffine& p) { return point_in(p, 0); }
  WitnessPoint append_constant_point(const JubJubAffine& p) { return point_in(p, 1); }
  WitnessPoint append_public_point(const JubJubAffine& p) { return point_in(p, 2); }
  void assert_equal_point(WitnessPoint a, WitnessPoint b) { check(pb200_composer_assert_equal_point(h_, &a.x, &b.x)); }
  void assert_equal_public_point(WitnessPoint p, const JubJubAffine& pub) {
    uint64_t uv[8];
    pub.to_raw(uv);
    check(pb200_composer_assert_equal_public_point(h_, &p.x, uv));
  }
  WitnessPoint assert_torsion_free_point(WitnessPoint p) { check(pb200_composer_assert_torsion_free_point(h_, &p.x)); return p; }
  WitnessPoint component_add_point(WitnessPoint a, WitnessPoint b) { WitnessPoint r; check(pb200_composer_point_op(h_, PB200_POINT_ADD, &a.x, &b.x, &r.x)); return r; }
  WitnessPoint component_sub_point(WitnessPoint a, WitnessPoint b) { WitnessPoint r; check(pb200_composer_point_op(h_, PB200_POINT_SUB, &a.x, &b.x, &r.x)); return r; }
  WitnessPoint component_neg_point(WitnessPoint a) { WitnessPoint r; check(pb200_composer_point_op(h_, PB200_POINT_NEG, &a.x, nullptr, &r.x)); return r; }
  WitnessPoint component_select_identity(Witness bit, WitnessPoint a) { WitnessPoint r; check(pb200_composer_component_select_identity(h_, bit, &a.x, &r.x)); return r; }
  WitnessPoint component_select_point(Witness bit, WitnessPoint a, WitnessPoint b) { WitnessPoint r; check(pb200_composer_component_select_point(h_, bit, &a.x, &b.x, &r.x)); return r; }
  WitnessPoint component_mul_point(Witness jubjub, WitnessPoint p) { WitnessPoint r; check(pb200_composer_component_mul_point(h_, jubjub, &p.x, &r.x)); return r; }
  WitnessPoint component_mul_generator(Witness jubjub, const JubJubAffine& generator) {
    uint64_t uv[8];
    generator.to_raw(uv);
    WitnessPoint r;
    check(pb200_composer_component_mul_generator(h_, jubjub, uv, &r.x));
    return r;
  }

  // What Compiler::compile and Prover::prove read back from the composer.
  struct Export {
    std::vector<BlsScalar> selectors, witnesses, pi_vals;
    std::vector<uint32_t> wires;
    std::vector<uint64_t> pi_idx;
    size_t n_constraints = 0;
  };
  Export finish() const {
    Export e;
    e.n_constraints = constraints();
    const size_t n_w = pb200_composer_witnesses(h_), n_pi = pb200_composer_public_inputs(h_);
    e.selectors.resize(11 * e.n_constraints);
    e.wires.resize(4 * e.n_constraints);
    e.witnesses.resize(n_w);
    e.pi_idx.resize(n_pi);
    e.pi_vals.resize(n_pi);
    check(pb200_composer_export(h_, e.selectors[0].data(), e.wires.data(), e.witnesses[0].data(), e.pi_idx.data(),
                                n_pi ? e.pi_vals[0].data() : nullptr));
    return e;
  }

 private:
  pb200_composer_t* h_ = nullptr;
  WitnessPoint point_in(const JubJubAffine& p, int kind) {
    uint64_t uv[8];
    p.to_raw(uv);
    WitnessPoint r;
    check(pb200_composer_append_point(h_, uv, kind, &r.x));
    return r;
  }
};

inline Circuit circuit_of(const Composer::Export& e) {
  Circuit c;
  c.selectors = e.selectors;
  c.wires = e.wires;
  c.n_constraints = e.n_constraints;
  c.n_witnesses = e.witnesses.size();
  return c;
}

class Prover {
 public:
  static constexpr size_t PROOF_SIZE = 1008;  // Proof::SIZE
  Prover(const std::string& label, const Circuit& c, const uint8_t* srs_raw, size_t n_srs_points) : n_witnesses_(c.n_witnesses) {
    check(pb200_prover_new((const uint8_t*)label.data(), label.size(), c.n_constraints, c.selectors[0].data(), c.wires.data(),
                           c.n_witnesses, srs_raw, n_srs_points, &h_));
  }
  // Prover::try_from_bytes (prover.rs:265-350) for the bytes of Prover::to_bytes; the wiring of the circuit
  // (4 x constraints witness indices) is not part of that format and comes alongside
  static std::unique_ptr<Prover> try_from_bytes(const uint8_t* bytes, size_t len, const std::vector<uint32_t>& wires, size_t n_witnesses) {
    std::unique_ptr<Prover> p(new Prover());
    p->n_witnesses_ = n_witnesses;
    check(pb200_prover_from_bytes(bytes, len, wires.data(), n_witnesses, &p->h_));
    return p;
  }
  ~Prover() { pb200_prover_free(h_); }
  Prover(const Prover&) = delete;
  Prover& operator=(const Prover&) = delete;
  // Prover::prove: `blinders` are the 14 BlsScalar::random draws of prove_inner, in its order.
  std::array<uint8_t, PROOF_SIZE> prove(const std::vector<BlsScalar>& witnesses, const std::vector<uint64_t>& pi_idx,
                                        const std::vector<BlsScalar>& pi_vals, const std::array<BlsScalar, 14>& blinders) const {
    std::array<uint8_t, PROOF_SIZE> proof;
    check(pb200_prove(h_, witnesses[0].data(), witnesses.size(), pi_idx.data(), pi_vals.empty() ? nullptr : pi_vals[0].data(),
                      pi_idx.size(), blinders[0].data(), proof.data()));
    return proof;
  }

 private:
  Prover() : n_witnesses_(0) {}
  pb200_prover_t* h_ = nullptr;
  size_t n_witnesses_;
};

}  // namespace plonk_b200
