// The reference's benchmark circuit (benches/plonk.rs:12-82) written against the C++ mirror
// (include/plonk_b200.hpp), gadget by gadget, as a C++ user of libplonk_b200 would.
//
//   bench_circuit export <degree>                 prints constraints, witnesses and an FNV-1a hash of
//                                                 the exported arrays (no GPU needed)
//   bench_circuit prove <degree> <file>           <file>: u64 n_srs | srs (96 B points) | 14 blinders;
//                                                 prints the 1008-byte proof as hex (GPU)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>

#include "../../include/plonk_b200.hpp"

using namespace plonk_b200;

struct BenchCircuit {
  uint64_t a = 2, b = 3, x = 6, y = 7;
  JubJubAffine z = JubJubAffine::generator().mul({7, 0, 0, 0});
  size_t degree;

  void circuit(Composer& composer) const {
    const BlsScalar one = scalar_from_u64(1);
    const Witness w_a = composer.append_witness(a);
    const Witness w_b = composer.append_witness(b);
    const Witness w_x = composer.append_witness(x);
    const Witness w_y = composer.append_witness(y);
    const WitnessPoint w_z = composer.append_point(z);
    size_t diff = 0, prev = composer.constraints();
    while (prev + diff < degree) {
      const Witness r_w = composer.gate_mul(Constraint().mult(one).a(w_a).b(w_b));
      composer.append_constant(15);
      composer.append_constant_point(z);
      composer.assert_equal(w_x, r_w);
      composer.assert_equal_point(w_z, w_z);
      composer.gate_add(Constraint().left(one).right(one).a(w_a).b(w_b));
      composer.component_add_point(w_z, w_z);
      composer.append_logic_and<127>(w_a, w_b);
      composer.append_logic_xor<127>(w_a, w_b);
      composer.component_boolean(Composer::ONE);
      composer.component_decomposition<254>(w_a);
      composer.component_mul_generator(w_y, JubJubAffine::generator());
      composer.component_mul_point(w_y, w_z);
      composer.component_range_bits<256>(w_a);
      composer.component_select(Composer::ONE, w_a, w_b);
      composer.component_select_identity(Composer::ONE, w_z);
      composer.component_select_one(Composer::ONE, w_a);
      composer.component_select_point(Composer::ONE, w_z, w_z);
      composer.component_select_zero(Composer::ONE, w_a);
      diff = composer.constraints() - prev;
      prev = composer.constraints();
    }
  }
};

static uint64_t fnv(uint64_t h, const void* data, size_t n) {
  const uint8_t* p = (const uint8_t*)data;
  for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 0x100000001b3ull;
  return h;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  try {
    BenchCircuit bc;
    bc.degree = strtoull(argv[2], nullptr, 10);
    Composer composer;
    bc.circuit(composer);
    const Composer::Export e = composer.finish();
    if (!strcmp(argv[1], "export")) {
      uint64_t h = 0xcbf29ce484222325ull;
      h = fnv(h, e.selectors.data(), e.selectors.size() * 32);
      h = fnv(h, e.wires.data(), e.wires.size() * 4);
      h = fnv(h, e.witnesses.data(), e.witnesses.size() * 32);
      std::printf("%zu %zu %016llx\n", e.n_constraints, e.witnesses.size(), (unsigned long long)h);
      return 0;
    }
    if (argc < 4) return 2;
    std::ifstream f(argv[3], std::ios::binary);
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    uint64_t n_srs;
    memcpy(&n_srs, buf.data(), 8);
    const uint8_t* srs = buf.data() + 8;
    std::array<BlsScalar, 14> blinders;
    memcpy(blinders.data(), srs + n_srs * 96, 14 * 32);
    check(pb200_init(0));
    Prover prover("dusk-network", circuit_of(e), srs, n_srs);
    const auto proof = prover.prove(e.witnesses, e.pi_idx, e.pi_vals, blinders);
    for (uint8_t b : proof) std::printf("%02x", b);
    std::puts("");
  } catch (const Error& err) {
    std::printf("error %d: %s\n", (int)err.kind, err.what());
    return 1;
  }
  return 0;
}
