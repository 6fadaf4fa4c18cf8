// Fiat-Shamir transcript of the prover: merlin 3.0 framing over STROBE-128 / Keccak-f[1600], plus
// the TranscriptProtocol helpers of the reference (src/transcript.rs:89-145).  Host-side product
// code (tiny, latency-bound, runs between GPU rounds); independent of oracle/.
#pragma once
#include <stdint.h>
#include <string.h>

#include "host_field.h"

namespace pbh {

inline void keccak_f1600(uint64_t s[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
      0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
      0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
      0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
      0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  // rho offsets indexed [x + 5y]
  static const int RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  for (int round = 0; round < 24; round++) {
    uint64_t c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
    for (int x = 0; x < 5; x++) {
      const uint64_t n = c[(x + 1) % 5];
      d[x] = c[(x + 4) % 5] ^ ((n << 1) | (n >> 63));
    }
    for (int i = 0; i < 25; i++) s[i] ^= d[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) {
        const int i = x + 5 * y, r = RHO[i];
        const uint64_t v = s[i];
        b[y + 5 * ((2 * x + 3 * y) % 5)] = r ? ((v << r) | (v >> (64 - r))) : v;
      }
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++) s[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    s[0] ^= RC[round];
  }
}

class Strobe128 {
 public:
  explicit Strobe128(const char* protocol_label) {
    memset(st_, 0, sizeof st_);
    const uint8_t init[6] = {1, kRate + 2, 1, 0, 1, 96};
    memcpy(st_, init, 6);
    memcpy(st_ + 6, "STROBEv1.0.2", 12);
    permute();
    pos_ = pos_begin_ = 0;
    meta_ad((const uint8_t*)protocol_label, strlen(protocol_label), false);
  }
  void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(kM | kA, more); absorb(d, n); }
  void ad(const uint8_t* d, size_t n, bool more) { begin_op(kA, more); absorb(d, n); }
  void prf(uint8_t* out, size_t n) { begin_op(kI | kA | kC, false); squeeze(out, n); }

 private:
  enum { kRate = 166, kI = 1, kA = 2, kC = 4, kT = 8, kM = 16, kK = 32 };
  uint8_t st_[200];
  int pos_, pos_begin_;
  void permute() {
    uint64_t w[25];
    memcpy(w, st_, 200);
    keccak_f1600(w);
    memcpy(st_, w, 200);
  }
  void run_f() {
    st_[pos_] ^= (uint8_t)pos_begin_;
    st_[pos_ + 1] ^= 0x04;
    st_[kRate + 1] ^= 0x80;
    permute();
    pos_ = pos_begin_ = 0;
  }
  void absorb(const uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
      st_[pos_++] ^= d[i];
      if (pos_ == kRate) run_f();
    }
  }
  void squeeze(uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
      d[i] = st_[pos_];
      st_[pos_++] = 0;
      if (pos_ == kRate) run_f();
    }
  }
  void begin_op(int flags, bool more) {
    if (more) return;
    const uint8_t hdr[2] = {(uint8_t)pos_begin_, (uint8_t)flags};
    pos_begin_ = pos_ + 1;
    absorb(hdr, 2);
    if ((flags & (kC | kK)) && pos_ != 0) run_f();
  }
};

class Transcript {
 public:
  Transcript(const uint8_t* label, size_t n) : s_("Merlin v1.0") { append_message("dom-sep", label, n); }
  void append_message(const char* label, const uint8_t* m, size_t n) {
    const uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    s_.meta_ad((const uint8_t*)label, strlen(label), false);
    s_.meta_ad(len, 4, true);
    s_.ad(m, n, false);
  }
  void append_u64(const char* label, uint64_t x) {
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
    append_message(label, b, 8);
  }
  // TranscriptProtocol (reference src/transcript.rs:89-108)
  void append_commitment(const char* label, const uint8_t compressed[48]) { append_message(label, compressed, 48); }
  void append_scalar(const char* label, const HFr& x) {
    HFr c = x.from_mont();
    append_message(label, (const uint8_t*)c.v, 32);
  }
  HFr challenge_scalar(const char* label) {  // 64 bytes -> BlsScalar::from_bytes_wide
    const uint8_t len[4] = {64, 0, 0, 0};
    uint8_t buf[64];
    s_.meta_ad((const uint8_t*)label, strlen(label), false);
    s_.meta_ad(len, 4, true);
    s_.prf(buf, 64);
    HFr lo, hi, r2;
    memcpy(lo.v, buf, 32);
    memcpy(hi.v, buf + 32, 32);
    memcpy(r2.v, kFrMod.r2, 32);
    const HFr r3 = r2 * r2;
    return lo * r2 + hi * r3;
  }
  void circuit_domain_sep(uint64_t n) {
    append_message("dom-sep", (const uint8_t*)"circuit_size", 12);
    append_u64("n", n);
  }

 private:
  Strobe128 s_;
};

}  // namespace pbh
