// Exercises include/plonk_b200.hpp end to end through libplonk_b200.so.
// Usage: api_check <file>   where <file> holds (little-endian u64 counts, then raw arrays):
//   n_srs, n_constraints, n_witnesses, n_pi, label_len | srs | selectors | wires | witnesses | pi_idx | pi_vals | blinders | label
// Prints the proof as hex, "UNSATISFIED", or an error.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>

#include "../../include/plonk_b200.hpp"

using namespace plonk_b200;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const uint8_t* p = buf.data();
  auto u64 = [&]() { uint64_t v; memcpy(&v, p, 8); p += 8; return v; };
  const size_t n_srs = u64(), n_c = u64(), n_w = u64(), n_pi = u64(), label_len = u64();
  const uint8_t* srs = p; p += n_srs * 96;
  Circuit c;
  c.n_constraints = n_c; c.n_witnesses = n_w;
  c.selectors.resize(11 * n_c); memcpy(c.selectors.data(), p, 11 * n_c * 32); p += 11 * n_c * 32;
  c.wires.resize(4 * n_c); memcpy(c.wires.data(), p, 4 * n_c * 4); p += 4 * n_c * 4;
  std::vector<BlsScalar> wit(n_w); memcpy(wit.data(), p, n_w * 32); p += n_w * 32;
  std::vector<uint64_t> pi_idx(n_pi); memcpy(pi_idx.data(), p, n_pi * 8); p += n_pi * 8;
  std::vector<BlsScalar> pi_vals(n_pi); memcpy(pi_vals.data(), p, n_pi * 32); p += n_pi * 32;
  std::array<BlsScalar, 14> bl; memcpy(bl.data(), p, 14 * 32); p += 14 * 32;
  std::string label((const char*)p, label_len);
  try {
    check(pb200_init(0));
    // level 1: domain + commit key
    EvaluationDomain dom(8);
    std::vector<BlsScalar> v(5, BlsScalar{1, 2, 3, 4});
    if (dom.ifft(dom.fft(v)).size() != 8) return 3;
    CommitKey key(srs, n_srs);
    try {
      key.commit(std::vector<BlsScalar>(n_srs + 1, BlsScalar{1, 0, 0, 0}));
      std::puts("missing PolynomialDegreeTooLarge");
      return 4;
    } catch (const Error& e) {
      if (e.kind != Error::PolynomialDegreeTooLarge) throw;
    }
    // level 2: prover
    Prover prover(label, c, srs, n_srs);
    auto proof = prover.prove(wit, pi_idx, pi_vals, bl);
    for (uint8_t b : proof) std::printf("%02x", b);
    std::puts("");
  } catch (const Error& e) {
    if (e.kind == Error::CircuitUnsatisfied) { std::puts("UNSATISFIED"); return 0; }
    std::printf("error %d: %s\n", (int)e.kind, e.what());
    return 1;
  }
  return 0;
}
