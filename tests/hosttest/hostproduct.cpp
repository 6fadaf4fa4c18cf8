// Host-only build (g++) of the product's CPU-side pieces - merlin transcript, 64-bit-limb fields, G1
// helpers (plonk_b200/csrc/{transcript.h,host_field.*}) - so that `-m "not gpu"` tests can compare
// them with the oracle without a GPU.
#include <stddef.h>
#include <string.h>

#include "../../plonk_b200/csrc/host_field.cpp"
#include "../../plonk_b200/csrc/transcript.h"

using namespace pbh;

extern "C" {

// merlin: Transcript::new(label); append_message(l1, m1); challenge_bytes-like scalar under l2
int hp_transcript_scalar(const char* label, const char* l1, const uint8_t* m1, size_t n1, const char* l2, uint64_t out[4]) {
  Transcript t((const uint8_t*)label, strlen(label));
  t.append_message(l1, m1, n1);
  HFr c = t.challenge_scalar(l2);
  memcpy(out, c.v, 32);
  return 0;
}

// the prover's base transcript steps with scalars and u64s mixed in
int hp_transcript_mix(const uint64_t scalar_mont[4], uint64_t n, uint64_t out[4]) {
  Transcript t((const uint8_t*)"mix", 3);
  t.circuit_domain_sep(n);
  HFr s;
  memcpy(s.v, scalar_mont, 32);
  t.append_scalar("s", s);
  uint8_t comm[48];
  memset(comm, 0, 48);
  comm[0] = 0xC0;
  t.append_commitment("c", comm);
  HFr a = t.challenge_scalar("a");
  t.append_scalar("a", a);
  HFr b = t.challenge_scalar("b");
  HFr r = a * b + s;
  memcpy(out, r.v, 32);
  return 0;
}

int hp_fp_inv(const uint64_t a[6], uint64_t out[6]) {
  HFp x;
  memcpy(x.v, a, 48);
  HFp r = x.inv();
  memcpy(out, r.v, 48);
  return 0;
}

// the binary-GCD inverses (host_field.h: inv_bingcd), as used by the circuit front end
int hp_fp_inv_bingcd(const uint64_t a[6], uint64_t out[6]) {
  HFp x;
  memcpy(x.v, a, 48);
  HFp r = x.inv_bingcd();
  memcpy(out, r.v, 48);
  return 0;
}
int hp_fr_inv_bingcd(const uint64_t a[4], uint64_t out[4]) {
  HFr x;
  memcpy(x.v, a, 32);
  HFr r = x.inv_bingcd();
  memcpy(out, r.v, 32);
  return 0;
}

int hp_g1_compress(const uint64_t raw[12], uint8_t out[48]) {
  g1_compress_raw(raw, out);
  return 0;
}
}
