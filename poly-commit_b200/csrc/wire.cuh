// G1 wire formats: the bytes ark-serialize's CanonicalSerialize / CanonicalDeserialize read and write for the G1Affine
// elements inside kzg10::Powers (kzg10/data_structures.rs:142-177), UniversalParams.powers_of_g (:57-112),
// kzg10::Commitment (:315-328) and kzg10::Proof.w (:479-495) -- SURVEY.md section 8(f) rank 1 ("SRS ingestion & wire
// formats").  The encodings themselves live in un-vendored crates (ark-serialize / ark-ec 0.5.0, ark-bls12-381 0.5.0) and
// are restated from their published behaviour:
//   * generic short-Weierstrass (BN254, Pallas): x little-endian in ceil((bits + 2) / 8) bytes, SWFlags in the two top bits
//     of the LAST byte (bit 7 = YIsNegative: y > -y as integers, bit 6 = PointAtInfinity, both = invalid); uncompressed is
//     x in ceil(bits / 8) bytes followed by y with the same flags.
//   * BLS12-381 (ark-bls12-381 overrides the generic form with the ZCash encoding): 48 / 96 bytes BIG-endian, flags in the
//     three top bits of the FIRST byte (bit 7 = compressed, bit 6 = infinity, bit 5 = y is the lexicographically largest).
// Decompression is the data-parallel hot part: one square root in Fq per point ((p+1)/4 exponent for BLS12-381 / BN254,
// Tonelli-Shanks for Pallas whose p - 1 has 2-adicity 32) and, with validation, the subgroup check (BLS12-381 only:
// sigma(P) = -[z^2] P, the endomorphism test of ark-bls12-381's is_in_correct_subgroup_assuming_on_curve; the other two
// curves have cofactor 1).  One thread per point; every kernel here is bound by the integer-multiply pipe like the MSM.
#pragma once
#include "ec.cuh"
#include "hash.cuh"
#include "msm.cuh"

namespace pcgpu {

enum { WIRE_OK = 0, WIRE_BAD_FLAGS = 1, WIRE_NOT_CANONICAL = 2, WIRE_NOT_ON_CURVE = 3, WIRE_NOT_IN_SUBGROUP = 4 };

template <class C> PCGPU_HD constexpr int wire_is_zcash() { return C::ID == 0; }
template <class C> PCGPU_HD constexpr int wire_x_bytes_flagged() { return wire_is_zcash<C>() ? 48 : (C::Fq::BITS + 2 + 7) / 8; }
template <class C> PCGPU_HD constexpr int wire_x_bytes_plain() { return wire_is_zcash<C>() ? 48 : (C::Fq::BITS + 7) / 8; }
template <class C> PCGPU_HD constexpr int wire_size(bool compressed) {
  return compressed ? wire_x_bytes_flagged<C>() : wire_x_bytes_plain<C>() + wire_x_bytes_flagged<C>();
}

// a^e for a compile-time exponent given as limbs (uniform control flow across the warp)
template <class P, class ExpFn>
PCGPU_DEV Fp<P> fp_pow_limbs(const Fp<P> &a, ExpFn e, int bits) {
  Fp<P> acc = Fp<P>::one();
  for (int i = bits - 1; i >= 0; i--) {
    acc = fp_sqr<P>(acc);
    if ((e(i / 32) >> (i % 32)) & 1) acc = fp_mul<P>(acc, a);
  }
  return acc;
}

// canonical integer comparison helpers (a, b < 2^(32N))
template <class P>
PCGPU_DEV bool limbs_gt(const uint32_t *a, const uint32_t *b) {
  for (int i = P::N - 1; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
  return false;
}
template <class P>
PCGPU_DEV bool limbs_ge_mod(const uint32_t *a) {
  uint32_t m[P::N];
  for (int i = 0; i < P::N; i++) m[i] = P::mod(i);
  return !limbs_gt<P>(m, a);
}
// y > -y as integers  <=>  canonical(y) > (p - 1) / 2
template <class P>
PCGPU_DEV bool fp_is_larger_half(const Fp<P> &y_mont) {
  Fp<P> c = fp_from_mont<P>(y_mont);
  uint32_t h[P::N];
  for (int i = 0; i < P::N; i++) h[i] = P::half(i);
  return limbs_gt<P>(c.l, h);
}

// square root in Fq; returns false when a is a non-residue.  Any root may be returned (callers order the pair).
template <class P>
PCGPU_DEV bool fp_sqrt(const Fp<P> &a, Fp<P> &out) {
  if (a.is_zero()) { out = a; return true; }
  if (!P::SQRT_TONELLI) {
    Fp<P> y = fp_pow_limbs<P>(a, [](int i) { return P::sqrt_exp(i); }, P::SQRT_EXP_BITS);
    out = y;
    return fp_sqr<P>(y) == a;
  }
  // Tonelli-Shanks, p - 1 = 2^s t:  w = a^((t-1)/2), x = a w, b = x w = a^t, z = g^t
  Fp<P> w = fp_pow_limbs<P>(a, [](int i) { return P::sqrt_exp(i); }, P::SQRT_EXP_BITS);
  Fp<P> x = fp_mul<P>(a, w);
  Fp<P> b = fp_mul<P>(x, w);
  Fp<P> z;
  for (int i = 0; i < P::N; i++) z.l[i] = P::ts_root(i);
  const Fp<P> one = Fp<P>::one();
  int v = P::FQ_TWO_ADICITY;
  while (!(b == one)) {
    int k = 0;
    Fp<P> t = b;
    while (!(t == one)) { t = fp_sqr<P>(t); k++; if (k == v) return false; }
    Fp<P> ww = z;
    for (int j = 0; j < v - k - 1; j++) ww = fp_sqr<P>(ww);
    z = fp_sqr<P>(ww);
    b = fp_mul<P>(b, z);
    x = fp_mul<P>(x, ww);
    v = k;
  }
  out = x;
  return true;
}

template <class C>
PCGPU_DEV Fp<typename C::Fq> curve_rhs(const Fp<typename C::Fq> &x) {
  using Q = typename C::Fq;
  Fp<Q> b;
  for (int i = 0; i < Q::N; i++) b.l[i] = Q::curve_b(i);
  return fp_add<Q>(fp_mul<Q>(fp_sqr<Q>(x), x), b);
}

// [k] P for a 64-bit k, P in XYZZ (left-to-right double-and-add; k is a compile-time constant so the flow is uniform)
template <class C>
PCGPU_DEV XYZZ<C> xyzz_mul_u64(const XYZZ<C> &p, unsigned long long k) {
  XYZZ<C> acc = XYZZ<C>::inf();
  for (int i = 63; i >= 0; i--) {
    acc = xyzz_dbl<C>(acc);
    if ((k >> i) & 1) xyzz_add<C>(acc, p);
  }
  return acc;
}

// membership in the prime-order subgroup for a point already known to be on the curve
template <class C>
PCGPU_DEV bool g1_in_subgroup(const Affine<C> &a) {
  using Q = typename C::Fq;
  if (Q::COFACTOR_ONE || a.is_inf()) return true;
  XYZZ<C> xp = xyzz_mul_u64<C>(xyzz_from_affine<C>(a), Q::SUBGROUP_X);
  // [x] P == P with P != O: the order of P divides x - 1, a factor of the cofactor
  if (!xp.is_inf() && fp_mul<Q>(a.x, xp.zz) == xp.x && fp_mul<Q>(a.y, xp.zzz) == xp.y) return false;
  XYZZ<C> x2p = xyzz_mul_u64<C>(xp, Q::SUBGROUP_X);
  if (x2p.is_inf()) return false;
  Fp<Q> beta;
  for (int i = 0; i < Q::N; i++) beta.l[i] = Q::beta(i);
  // sigma(P) = (beta x, y) must equal -[x^2] P = (X, -Y, ZZ, ZZZ)
  return fp_mul<Q>(fp_mul<Q>(beta, a.x), x2p.zz) == x2p.x && fp_mul<Q>(a.y, x2p.zzz) == fp_neg<Q>(x2p.y);
}

// ---- byte <-> limb helpers: `src` holds nbytes of one field element in the curve's byte order, flags already masked ----
template <class P>
PCGPU_DEV void limbs_from_bytes(const uint8_t *src, int nbytes, bool big_endian, uint32_t *l) {
  for (int i = 0; i < P::N; i++) l[i] = 0;
  for (int k = 0; k < nbytes && k < 4 * P::N; k++) {
    uint32_t byte = big_endian ? src[nbytes - 1 - k] : src[k];
    l[k / 4] |= byte << (8 * (k % 4));
  }
}
template <class P>
PCGPU_DEV void limbs_to_bytes(const uint32_t *l, int nbytes, bool big_endian, uint8_t *dst) {
  for (int k = 0; k < nbytes; k++) {
    uint8_t byte = k < 4 * P::N ? (uint8_t)(l[k / 4] >> (8 * (k % 4))) : 0;
    if (big_endian) dst[nbytes - 1 - k] = byte; else dst[k] = byte;
  }
}

// CanonicalDeserialize for G1Affine: bytes -> Montgomery x||y + infinity byte + status
template <class C>
struct G1DecodeBody {
  const uint8_t *bytes; uint32_t *out_xy; uint8_t *out_inf; uint8_t *status; int compressed; int validate;
  PCGPU_KERNEL_DEV void operator()(size_t i) const {
    using Q = typename C::Fq;
    constexpr int N = Q::N;
    constexpr bool BE = wire_is_zcash<C>();
    constexpr int XF = wire_x_bytes_flagged<C>(), XP = wire_x_bytes_plain<C>();
    const int sz = wire_size<C>(compressed != 0);
    const uint8_t *src = bytes + i * (size_t)sz;
    uint8_t buf[2 * 48 + 2];
    for (int k = 0; k < sz; k++) buf[k] = src[k];
    uint32_t *oxy = out_xy + i * (size_t)(2 * N);
    auto finish = [&](int st, bool inf, const Fp<Q> &x, const Fp<Q> &y) {
      for (int j = 0; j < N; j++) { oxy[j] = (st || inf) ? 0u : x.l[j]; oxy[N + j] = (st || inf) ? 0u : y.l[j]; }
      out_inf[i] = (!st && inf) ? 1 : 0;
      status[i] = (uint8_t)st;
    };
    Fp<Q> x = Fp<Q>::zero(), y = Fp<Q>::zero();
    bool is_inf, y_flag;
    if (BE) {
      const uint8_t f = buf[0];
      const bool fc = (f & 0x80) != 0;
      is_inf = (f & 0x40) != 0; y_flag = (f & 0x20) != 0;
      if (fc != (compressed != 0)) { finish(WIRE_BAD_FLAGS, false, x, y); return; }
      // EncodingFlags::get_flags (ark-bls12-381 curves/util.rs): the sort flag is only legal on a compressed finite point
      if (y_flag && (!fc || is_inf)) { finish(WIRE_BAD_FLAGS, false, x, y); return; }
      buf[0] &= 0x1f;
      if (is_inf) {   // read_g1_{compressed,uncompressed}: the payload of the identity must be all zero (canonical encoding only)
        uint8_t any = 0;
        for (int k = 0; k < sz; k++) any |= buf[k];
        if (any) { finish(WIRE_NOT_CANONICAL, false, x, y); return; }
      }
    } else {
      // flags sit in the last byte of the flagged element (x when compressed, y otherwise)
      uint8_t &last = buf[sz - 1];
      const bool neg = (last & 0x80) != 0;
      is_inf = (last & 0x40) != 0; y_flag = neg;
      if (neg && is_inf) { finish(WIRE_BAD_FLAGS, false, x, y); return; }
      last &= 0x3f;
    }
    if (is_inf && BE) { finish(WIRE_OK, true, x, y); return; }
    uint32_t xl[N], yl[N];
    limbs_from_bytes<Q>(buf, compressed ? XF : XP, BE, xl);
    // the generic form carries one byte beyond the limbs when bits + 2 > 8 * 4N' (Pallas: 33 bytes).  ark-ff's
    // deserialize_with_flags strips the flag bits from that byte and then converts only the 8N'-byte limb buffer
    // (SerBuffer::to_bigint), so its six low bits are IGNORED, not checked -- mirrored here (limbs_from_bytes stops at 4N bytes).
    if (limbs_ge_mod<Q>(xl)) { finish(WIRE_NOT_CANONICAL, false, x, y); return; }
    if (!compressed) {
      const uint8_t *yb = buf + XP;
      limbs_from_bytes<Q>(yb, BE ? XP : XF, BE, yl);
      if (limbs_ge_mod<Q>(yl)) { finish(WIRE_NOT_CANONICAL, false, x, y); return; }
    }
    if (is_inf) { finish(WIRE_OK, true, x, y); return; }
    for (int j = 0; j < N; j++) x.l[j] = xl[j];
    x = fp_to_mont<Q>(x);
    if (compressed) {
      Fp<Q> r;
      if (!fp_sqrt<Q>(curve_rhs<C>(x), r)) { finish(WIRE_NOT_ON_CURVE, false, x, y); return; }
      // (smaller, larger) as integers; BLS12-381: flag = take the larger; generic: YIsNegative = take the larger
      const bool r_large = fp_is_larger_half<Q>(r);
      y = (r_large == y_flag) ? r : fp_neg<Q>(r);
    } else {
      for (int j = 0; j < N; j++) y.l[j] = yl[j];
      y = fp_to_mont<Q>(y);
      if (validate && !(fp_sqr<Q>(y) == curve_rhs<C>(x))) { finish(WIRE_NOT_ON_CURVE, false, x, y); return; }
    }
    if (validate) {
      Affine<C> a; a.x = x; a.y = y;
      if (!g1_in_subgroup<C>(a)) { finish(WIRE_NOT_IN_SUBGROUP, false, x, y); return; }
    }
    finish(WIRE_OK, false, x, y);
  }
};

// CanonicalSerialize for G1Affine: Montgomery x||y + infinity byte -> bytes
template <class C>
struct G1EncodeBody {
  const uint32_t *xy; const uint8_t *inf; uint8_t *bytes; int compressed;
  PCGPU_KERNEL_DEV void operator()(size_t i) const {
    using Q = typename C::Fq;
    constexpr int N = Q::N;
    constexpr bool BE = wire_is_zcash<C>();
    constexpr int XF = wire_x_bytes_flagged<C>(), XP = wire_x_bytes_plain<C>();
    const int sz = wire_size<C>(compressed != 0);
    uint8_t buf[2 * 48 + 2];
    Fp<Q> x, y;
    const uint32_t *p = xy + i * (size_t)(2 * N);
    for (int j = 0; j < N; j++) { x.l[j] = p[j]; y.l[j] = p[N + j]; }
    // device convention: (0, 0) is the identity as well (ec.cuh)
    const bool is_inf = (inf && inf[i]) || (x.is_zero() && y.is_zero());
    bool large = false;
    if (is_inf) { x = Fp<Q>::zero(); y = x; } else large = fp_is_larger_half<Q>(y);
    Fp<Q> xc = fp_from_mont<Q>(x), yc = fp_from_mont<Q>(y);
    if (BE) {
      limbs_to_bytes<Q>(xc.l, XP, true, buf);
      if (!compressed) limbs_to_bytes<Q>(yc.l, XP, true, buf + XP);
      buf[0] |= (compressed ? 0x80 : 0) | (is_inf ? 0x40 : 0) | ((compressed && !is_inf && large) ? 0x20 : 0);
    } else {
      const uint8_t flags = is_inf ? 0x40 : (large ? 0x80 : 0);
      if (compressed) { limbs_to_bytes<Q>(xc.l, XF, false, buf); buf[XF - 1] |= flags; }
      else { limbs_to_bytes<Q>(xc.l, XP, false, buf); limbs_to_bytes<Q>(yc.l, XF, false, buf + XP); buf[XP + XF - 1] |= flags; }
    }
    uint8_t *dst = bytes + i * (size_t)sz;
    for (int k = 0; k < sz; k++) dst[k] = buf[k];
  }
};

// ---------------------------------------------------------------------------------------------------------------------------
// Generator sampling: InnerProductArgPC::sample_generators (ipa_pc/mod.rs:302-325) and HyraxPC::setup's
// copy of it (hyrax/mod.rs:143-163).  Generator i is  Blake2s(PROTOCOL_NAME || i_le64)  fed to G::from_random_bytes, and
// while that returns None  Blake2s(PROTOCOL_NAME || i_le64 || j_le64)  for j = 0, 1, ...; then mul_by_cofactor (1 on the
// curves these schemes are instantiated on here) and normalize_batch (a no-op on affine output).
// G::from_random_bytes (ark-ec short_weierstrass Affine, un-vendored, restated): the digest is read as an x-coordinate with
// SWFlags -- Fp::from_random_bytes_with_flags keeps MODULUS_BIT_SIZE bits of the little-endian bytes and takes the flags from
// the top two bits of byte ceil((bits + 2) / 8) - 1 of the (zero-extended) input, so a 32-byte digest carries flag bits only
// when bits + 2 <= 256 (BN254) and none on a 255-bit field (Pallas: always "positive"); x >= p, both flags, or the infinity
// flag on a non-zero x give None; otherwise the point is (x, y) with y the lexicographically LARGER root when the flag says
// positive (get_point_from_x_unchecked(x, greatest = y_is_positive)) and None when x^3 + b has no root.
// One thread per generator: one or a few 64-byte Blake2s blocks and one square root in Fq -- bound by the multiply pipe.
// ---------------------------------------------------------------------------------------------------------------------------
enum { SAMPLE_NAME_MAX = 40 };

template <class C>
struct SampleGeneratorsBody {
  uint8_t name[SAMPLE_NAME_MAX]; uint32_t name_len; uint64_t first; uint32_t *out_xy;
  PCGPU_DEV static bool hash_to_point(const uint32_t *digest, Affine<C> &pt) {
    using Q = typename C::Fq;
    constexpr int N = Q::N;
    constexpr int FLAG_BYTE = (Q::BITS + 2 + 7) / 8 - 1;          // 31 for a 254-bit field, 32 for a 255-bit one
    uint32_t flagbits = 0;
    if (FLAG_BYTE < 32) flagbits = (digest[FLAG_BYTE / 4] >> (8 * (FLAG_BYTE % 4))) & 0xC0u;
    uint32_t xl[N];
    for (int i = 0; i < N; i++) xl[i] = i < 8 ? digest[i] : 0u;
    if (Q::BITS % 32) xl[N - 1] &= (1u << (Q::BITS % 32)) - 1;    // keep MODULUS_BIT_SIZE bits
    if (limbs_ge_mod<Q>(xl)) return false;
    const bool neg = (flagbits & 0x80u) != 0, inf = (flagbits & 0x40u) != 0;
    if (neg && inf) return false;
    bool zero = true;
    for (int i = 0; i < N; i++) zero &= xl[i] == 0;
    if (inf) { if (!zero) return false; pt = Affine<C>::inf(); return true; }
    Fp<Q> x;
    for (int i = 0; i < N; i++) x.l[i] = xl[i];
    x = fp_to_mont<Q>(x);
    Fp<Q> r;
    if (!fp_sqrt<Q>(curve_rhs<C>(x), r)) return false;
    const bool want_larger = !neg;                                 // greatest = y_is_positive
    pt.x = x;
    pt.y = (fp_is_larger_half<Q>(r) == want_larger) ? r : fp_neg<Q>(r);
    return true;
  }
  PCGPU_KERNEL_DEV void operator()(size_t t) const {
    constexpr int N = C::Fq::N;
    const uint64_t i = first + t;
    Affine<C> pt;
    for (uint64_t attempt = 0;; attempt++) {
      // message: name || i (8 bytes LE) [|| j = attempt - 1 (8 bytes LE)] -- at most 56 bytes, one final block
      uint8_t msg[64];
      uint32_t len = 0;
      for (uint32_t k = 0; k < name_len; k++) msg[len++] = name[k];
      for (int k = 0; k < 8; k++) msg[len++] = (uint8_t)(i >> (8 * k));
      if (attempt) { const uint64_t j = attempt - 1; for (int k = 0; k < 8; k++) msg[len++] = (uint8_t)(j >> (8 * k)); }
      Blake2s b;
      b.init();
      for (int w = 0; w < 16; w++) {
        uint32_t v = 0;
        for (int k = 0; k < 4; k++) { const uint32_t pos = 4 * w + k; if (pos < len) v |= (uint32_t)msg[pos] << (8 * k); }
        b.m[w] = v;
      }
      b.compress(len, true);
      if (hash_to_point(b.h, pt)) break;
    }
    uint32_t *o = out_xy + t * (size_t)(2 * N);
    for (int k = 0; k < N; k++) { o[k] = pt.x.l[k]; o[N + k] = pt.y.l[k]; }
  }
};

}  // namespace pcgpu
