// C++ mirror of the reference's interface for the hot path, over the C ABI in plonk_b200.h.
//
// The reference is a compiled (Rust) crate; this header is the compiled-language host side a C++
// caller links against, with the reference's names, argument meaning and error behaviour:
//
//   plonk_b200::EvaluationDomain   src/fft/domain.rs:35-232      new / fft / ifft / coset_fft / coset_ifft
//   plonk_b200::CommitKey          src/commitment_scheme/kzg10/key.rs:36-41, 362-388   commit, max_degree
//   plonk_b200::Commitment         src/commitment_scheme/kzg10/commitment.rs:77-106    to_bytes (48 B)
//   plonk_b200::Prover             src/compiler/prover.rs:53-115, 352-362              prove
//   plonk_b200::Error              src/error.rs:21-120 (the variants this path can produce)
//
// BlsScalar is the reference's in-memory layout: 4 x u64 little-endian limbs, Montgomery form.
// Infallible reference functions (the NTT family asserts/panics, domain.rs:394) throw
// BackendFailure on a device error - there is no CPU fallback.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "plonk_b200.h"

namespace plonk_b200 {

using BlsScalar = std::array<uint64_t, 4>;

struct Error : std::runtime_error {
  enum Kind { InvalidEvalDomainSize, PolynomialDegreeTooLarge, CircuitUnsatisfied, InvalidArgument, BackendFailure };
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
};

// Flat circuit description (what Compiler::preprocess reads out of the Composer, compiler.rs:132-170)
struct Circuit {
  std::vector<BlsScalar> selectors;  // 11 columns x n_constraints, column-major (order of plonk_b200.h)
  std::vector<uint32_t> wires;       // 4 columns x n_constraints witness indices
  size_t n_constraints = 0;
  size_t n_witnesses = 0;
};

class Prover {
 public:
  static constexpr size_t PROOF_SIZE = 1008;  // Proof::SIZE
  Prover(const std::string& label, const Circuit& c, const uint8_t* srs_raw, size_t n_srs_points) : n_witnesses_(c.n_witnesses) {
    check(pb200_prover_new((const uint8_t*)label.data(), label.size(), c.n_constraints, c.selectors[0].data(), c.wires.data(),
                           c.n_witnesses, srs_raw, n_srs_points, &h_));
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
  pb200_prover_t* h_ = nullptr;
  size_t n_witnesses_;
};

}  // namespace plonk_b200
