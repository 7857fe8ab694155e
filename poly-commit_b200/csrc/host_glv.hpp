// GLV decomposition of ONE scalar on the host: k = k1 + k2 * lambda (mod r) with |k1|, |k2| < 2^129, for the curves whose
// order-r group carries the endomorphism phi(x, y) = (zeta x, y) = [lambda](x, y) (params_gen.cuh: glv_zeta, glv_lambda and
// the reduced lattice basis (a1, b1), (a2, b2), a1 b2 - a2 b1 = r).  Used by the IPA key fold (ipa.cuh, G1FoldGlvBody): every
// point of a round is multiplied by the SAME challenge (ipa_pc/mod.rs:699-701), so the split costs microseconds per round and
// halves the doublings of all n/2 scalar multiplications.
//   c1 = round(b2 k / r), c2 = round(-b1 k / r), k1 = k - c1 a1 - c2 a2, k2 = -c1 b1 - c2 b2      (Gallant-Lambert-Vanstone)
// The result is VERIFIED in Fr (k1 + k2 lambda == k) and bounded before use; callers fall back to the plain ladder if not ok.
#pragma once
#include "host_ec.hpp"

namespace pcgpu {
namespace host {

struct UBig {  // 512-bit unsigned, little-endian limbs
  uint64_t l[8];
  static UBig zero() { UBig r; memset(r.l, 0, sizeof r.l); return r; }
  static UBig from(const uint64_t *p, int n) { UBig r = zero(); for (int i = 0; i < n; i++) r.l[i] = p[i]; return r; }
  bool is_zero() const { uint64_t o = 0; for (int i = 0; i < 8; i++) o |= l[i]; return o == 0; }
  int bits() const { for (int i = 7; i >= 0; i--) if (l[i]) return 64 * i + 64 - __builtin_clzll(l[i]); return 0; }
};
inline int ucmp(const UBig &a, const UBig &b) { for (int i = 7; i >= 0; i--) { if (a.l[i] != b.l[i]) return a.l[i] > b.l[i] ? 1 : -1; } return 0; }
inline UBig uadd(const UBig &a, const UBig &b) { UBig r; uint64_t c = 0; for (int i = 0; i < 8; i++) { u128 s = (u128)a.l[i] + b.l[i] + c; r.l[i] = (uint64_t)s; c = (uint64_t)(s >> 64); } return r; }
inline UBig usub(const UBig &a, const UBig &b) { UBig r; uint64_t br = 0; for (int i = 0; i < 8; i++) { u128 d = (u128)a.l[i] - b.l[i] - br; r.l[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; } return r; }
inline UBig umul(const UBig &a, const UBig &b) {  // low 512 bits of the product
  UBig r = UBig::zero();
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; i + j < 8; j++) { u128 s = (u128)a.l[i] * b.l[j] + r.l[i + j] + c; r.l[i + j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
  }
  return r;
}
// floor(a / d), d != 0 (binary long division; a few microseconds, once per IPA round)
inline UBig udiv(const UBig &a, const UBig &d) {
  UBig q = UBig::zero(), rem = UBig::zero();
  for (int i = a.bits() - 1; i >= 0; i--) {
    for (int j = 7; j > 0; j--) rem.l[j] = (rem.l[j] << 1) | (rem.l[j - 1] >> 63);
    rem.l[0] = (rem.l[0] << 1) | ((a.l[i >> 6] >> (i & 63)) & 1);
    if (ucmp(rem, d) >= 0) { rem = usub(rem, d); q.l[i >> 6] |= (uint64_t)1 << (i & 63); }
  }
  return q;
}
struct SBig { UBig m; bool neg; };
inline SBig sadd(const SBig &a, const SBig &b) {
  if (a.neg == b.neg) return SBig{uadd(a.m, b.m), a.neg};
  int c = ucmp(a.m, b.m);
  if (c == 0) return SBig{UBig::zero(), false};
  return c > 0 ? SBig{usub(a.m, b.m), a.neg} : SBig{usub(b.m, a.m), b.neg};
}
inline SBig smul(const SBig &a, const SBig &b) { return SBig{umul(a.m, b.m), a.neg != b.neg}; }

struct GlvSplit {
  uint32_t k1[5], k2[5];   // magnitudes, 160 bits
  uint32_t neg1, neg2;     // signs
  uint32_t nbits;          // max bit length of the two magnitudes
  // joint sparse form of (|k1|, |k2|) (Solinas): digits in {-1, 0, 1}, at most one of any two consecutive columns is
  // non-zero in the joint sense -- on average half of the columns are (0, 0).  Bit j of *_nz / *_sg = digit j non-zero / negative.
  uint32_t u1_nz[5], u1_sg[5], u2_nz[5], u2_sg[5];
  uint32_t jsf_len;        // number of columns (<= nbits + 1)
  bool ok;
};

// Joint sparse form of two non-negative integers of at most 158 bits (Solinas 2001; Handbook of Elliptic and Hyperelliptic Curve
// Cryptography, Alg. 9.27).  Returns the number of columns.
inline uint32_t jsf_recode(UBig a, UBig b, uint32_t *u1_nz, uint32_t *u1_sg, uint32_t *u2_nz, uint32_t *u2_sg) {
  for (int i = 0; i < 5; i++) u1_nz[i] = u1_sg[i] = u2_nz[i] = u2_sg[i] = 0;
  auto shr1 = [](UBig &v) { for (int j = 0; j < 7; j++) v.l[j] = (v.l[j] >> 1) | (v.l[j + 1] << 63); v.l[7] >>= 1; };
  uint32_t d1 = 0, d2 = 0, j = 0;
  while ((!a.is_zero() || d1 || !b.is_zero() || d2) && j < 160) {
    const uint32_t l1 = (uint32_t)((a.l[0] & 7) + d1) & 7, l2 = (uint32_t)((b.l[0] & 7) + d2) & 7;
    int u1 = 0, u2 = 0;
    if (l1 & 1) { u1 = 2 - (int)(l1 & 3); if ((l1 == 3 || l1 == 5) && (l2 & 3) == 2) u1 = -u1; }
    if (l2 & 1) { u2 = 2 - (int)(l2 & 3); if ((l2 == 3 || l2 == 5) && (l1 & 3) == 2) u2 = -u2; }
    if ((int)(2 * d1) == 1 + u1) d1 = 1 - d1;
    if ((int)(2 * d2) == 1 + u2) d2 = 1 - d2;
    if (u1) { u1_nz[j >> 5] |= 1u << (j & 31); if (u1 < 0) u1_sg[j >> 5] |= 1u << (j & 31); }
    if (u2) { u2_nz[j >> 5] |= 1u << (j & 31); if (u2 < 0) u2_sg[j >> 5] |= 1u << (j & 31); }
    shr1(a); shr1(b);
    j++;
  }
  return j;
}

template <class C>
inline GlvSplit glv_decompose(const uint64_t *k_canonical) {
  using R = typename C::Fr;
  GlvSplit out;
  memset(&out, 0, sizeof out);
  auto mag128 = [](uint32_t (*f)(int)) { uint64_t v[2] = {(uint64_t)f(0) | ((uint64_t)f(1) << 32), (uint64_t)f(2) | ((uint64_t)f(3) << 32)}; return UBig::from(v, 2); };
  const SBig a1{mag128(R::glv_a1), R::GLV_A1_NEG != 0}, b1{mag128(R::glv_b1), R::GLV_B1_NEG != 0};
  const SBig a2{mag128(R::glv_a2), R::GLV_A2_NEG != 0}, b2{mag128(R::glv_b2), R::GLV_B2_NEG != 0};
  uint64_t rl[4];
  for (int i = 0; i < 4; i++) rl[i] = HFp<R>::mod(i);
  const UBig r = UBig::from(rl, 4), k = UBig::from(k_canonical, 4);
  UBig half = r;
  for (int j = 0; j < 7; j++) half.l[j] = (half.l[j] >> 1) | (half.l[j + 1] << 63);
  half.l[7] >>= 1;
  // c1 = round(b2 k / r), c2 = round(-b1 k / r)
  const SBig c1{udiv(uadd(umul(b2.m, k), half), r), b2.neg};
  const SBig c2{udiv(uadd(umul(b1.m, k), half), r), !b1.neg};
  const SBig ks{k, false};
  SBig t1 = smul(c1, a1); t1.neg = !t1.neg;
  SBig t2 = smul(c2, a2); t2.neg = !t2.neg;
  const SBig k1 = sadd(sadd(ks, t1), t2);
  SBig u1 = smul(c1, b1); u1.neg = !u1.neg;
  SBig u2 = smul(c2, b2); u2.neg = !u2.neg;
  const SBig k2 = sadd(u1, u2);
  if (k1.m.bits() > 132 || k2.m.bits() > 132) return out;
  // verification in Fr: k1 + k2 lambda == k
  auto to_fr = [](const SBig &v) {
    HFp<R> x = HFp<R>::zero(), r2;
    for (int i = 0; i < 4; i++) { x.l[i] = v.m.l[i]; r2.l[i] = (uint64_t)R::r2(2 * i) | ((uint64_t)R::r2(2 * i + 1) << 32); }
    x = mul<R>(x, r2);
    return v.neg ? sub<R>(HFp<R>::zero(), x) : x;
  };
  HFp<R> lam;
  for (int i = 0; i < 4; i++) lam.l[i] = (uint64_t)R::glv_lambda(2 * i) | ((uint64_t)R::glv_lambda(2 * i + 1) << 32);
  if (!(add<R>(to_fr(k1), mul<R>(to_fr(k2), lam)) == to_fr(ks))) return out;
  for (int i = 0; i < 5; i++) {
    out.k1[i] = (uint32_t)(k1.m.l[i >> 1] >> (32 * (i & 1)));
    out.k2[i] = (uint32_t)(k2.m.l[i >> 1] >> (32 * (i & 1)));
  }
  out.neg1 = k1.neg && !k1.m.is_zero(); out.neg2 = k2.neg && !k2.m.is_zero();
  out.nbits = (uint32_t)(k1.m.bits() > k2.m.bits() ? k1.m.bits() : k2.m.bits());
  out.jsf_len = jsf_recode(k1.m, k2.m, out.u1_nz, out.u1_sg, out.u2_nz, out.u2_sg);
  {   // the recoding is checked before use: sum_j u[j] 2^j == |k| for both halves (signed accumulation over 192 bits)
    for (int h = 0; h < 2; h++) {
      const uint32_t *nz = h ? out.u2_nz : out.u1_nz, *sg = h ? out.u2_sg : out.u1_sg;
      SBig acc{UBig::zero(), false};
      for (uint32_t j = 0; j < out.jsf_len; j++) {
        if (!((nz[j >> 5] >> (j & 31)) & 1)) continue;
        UBig w = UBig::zero(); w.l[j >> 6] = (uint64_t)1 << (j & 63);
        acc = sadd(acc, SBig{w, ((sg[j >> 5] >> (j & 31)) & 1) != 0});
      }
      const UBig &want = h ? k2.m : k1.m;
      if (acc.neg || ucmp(acc.m, want) != 0) return out;
    }
  }
  out.ok = true;
  return out;
}

}  // namespace host
}  // namespace pcgpu
