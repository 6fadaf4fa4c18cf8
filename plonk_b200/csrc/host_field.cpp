// Definitions for host_field.h (host-side Montgomery constants, XYZZ helpers, G1 compression).
#include "host_field.h"

namespace pbh {

const Mod64<6> kFpMod = {
    {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull},
    0x89f3fffcfffcfffdull,
    {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull},
    {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull, 0x67eb88a9939d83c0ull, 0x9a793e85b519952dull, 0x11988fe592cae3aaull}};
const Mod64<4> kFrMod = {
    {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull},
    0xfffffffeffffffffull,
    {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full},
    {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull}};

HXyzz hxyzz_dbl(const HXyzz& p) {
  if (p.is_inf()) return p;
  if (p.y.is_zero()) return HXyzz::identity();
  HXyzz r;
  HFp u = p.y.dbl();
  HFp v = u.sqr();
  HFp w = u * v;
  HFp s = p.x * v;
  HFp xx = p.x.sqr();
  HFp m = xx.dbl() + xx;
  r.x = m.sqr() - s.dbl();
  r.y = m * (s - r.x) - w * p.y;
  r.zz = v * p.zz;
  r.zzz = w * p.zzz;
  return r;
}

void hxyzz_add(HXyzz& acc, const HXyzz& o) {
  if (o.is_inf()) return;
  if (acc.is_inf()) {
    acc = o;
    return;
  }
  HFp u1 = acc.x * o.zz, u2 = o.x * acc.zz;
  HFp s1 = acc.y * o.zzz, s2 = o.y * acc.zzz;
  HFp p = u2 - u1, r = s2 - s1;
  if (p.is_zero()) {
    if (r.is_zero())
      acc = hxyzz_dbl(acc);
    else
      acc = HXyzz::identity();
    return;
  }
  HFp pp = p.sqr();
  HFp ppp = p * pp;
  HFp q = u1 * pp;
  HFp x3 = r.sqr() - ppp - q.dbl();
  HFp y3 = r * (q - x3) - s1 * ppp;
  acc.x = x3;
  acc.y = y3;
  acc.zz = acc.zz * o.zz * pp;
  acc.zzz = acc.zzz * o.zzz * ppp;
}

bool hxyzz_to_affine(const HXyzz& p, HFp* x, HFp* y) {
  if (p.is_inf()) {
    *x = HFp::zero();
    *y = HFp::zero();
    return false;
  }
  HFp i = (p.zz * p.zzz).inv();
  *x = p.x * (i * p.zzz);
  *y = p.y * (i * p.zz);
  return true;
}

// G1Affine::to_bytes (zcash compressed form): big-endian x, bit7 = compressed, bit6 = infinity,
// bit5 = y lexicographically largest.
void g1_compress_raw(const uint64_t* raw, uint8_t out[48]) {
  HFp x, y;
  memcpy(x.v, raw, 48);
  memcpy(y.v, raw + 6, 48);
  if (x.is_zero() && y.is_zero()) {
    memset(out, 0, 48);
    out[0] = 0xC0;
    return;
  }
  HFp xc = x.from_mont();
  for (int i = 0; i < 6; i++)
    for (int b = 0; b < 8; b++) out[47 - (8 * i + b)] = (uint8_t)(xc.v[i] >> (8 * b));
  out[0] |= 0x80;
  if (y.canonical_gt_half()) out[0] |= 0x20;
}

}  // namespace pbh
