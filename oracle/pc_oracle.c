/*
 * TEST INFRASTRUCTURE -- CPU oracle for the pcgpu hot path.  NOT product code:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libpcgpu.so) never links
 * or calls it.
 *
 * PARITY STATUS: the reference's tests hold no golden vectors for commitments or
 * proofs and its arithmetic lives in un-vendored crates
 * (ark-ff / ark-ec / ark-poly 0.5.0) that cannot be built here (no Rust toolchain),
 * so no output of the reference itself backs this file: "parity unpinned" by the
 * reference.  It is pinned instead by (o) vectors PUBLISHED outside this
 * repository (tests/golden/external_kats.json: EIP-196 / go-ethereum 2G, 3G and an
 * ecMul vector through the MSM, EIP-2537 / ZCash BLS12-381 multiples, zkcrypto
 * Montgomery constants, the published Fr roots of unity and the domain-generator
 * rule; tests/test_external_kats.py), (i) an independent Python big-integer
 * implementation (oracle/pyref.py) via the committed fixtures in tests/golden/,
 * (ii) group-law identities (r*G = O, on-curve, linearity), (iii) the reference's
 * own small-integer KATs for the Fr helpers (utils.rs:274-286 test_row_mul;
 * linear_codes/utils.rs:303-331 test_reed_solomon's fft == evaluate property).
 *
 * Every exported function cites the reference call site whose dataflow it follows.
 * Data conventions (include/pcgpu.h): little-endian u64 limbs; field elements in
 * Montgomery form unless "canonical" is in the name; affine point = x||y plus a
 * separate infinity byte.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "params_gen.h"

typedef unsigned __int128 u128;

#define PFX bls
#define NQ 6
#define NR 4
#define QMOD BLS12_381_FQ_MOD
#define QONE BLS12_381_FQ_ONE
#define QR2 BLS12_381_FQ_R2
#define QM0 BLS12_381_FQ_M0
#define QB BLS12_381_FQ_B
#define RMOD BLS12_381_FR_MOD
#define RONE BLS12_381_FR_ONE
#define RR2 BLS12_381_FR_R2
#define RM0 BLS12_381_FR_M0
#define RBITS BLS12_381_FR_BITS
#include "curve_impl.inc"
#undef PFX
#undef NQ
#undef NR
#undef QMOD
#undef QONE
#undef QR2
#undef QM0
#undef QB
#undef RMOD
#undef RONE
#undef RR2
#undef RM0
#undef RBITS

#define PFX bn
#define NQ 4
#define NR 4
#define QMOD BN254_FQ_MOD
#define QONE BN254_FQ_ONE
#define QR2 BN254_FQ_R2
#define QM0 BN254_FQ_M0
#define QB BN254_FQ_B
#define RMOD BN254_FR_MOD
#define RONE BN254_FR_ONE
#define RR2 BN254_FR_R2
#define RM0 BN254_FR_M0
#define RBITS BN254_FR_BITS
#include "curve_impl.inc"
#undef PFX
#undef NQ
#undef NR
#undef QMOD
#undef QONE
#undef QR2
#undef QM0
#undef QB
#undef RMOD
#undef RONE
#undef RR2
#undef RM0
#undef RBITS

#define PFX pal
#define NQ 4
#define NR 4
#define QMOD PALLAS_FQ_MOD
#define QONE PALLAS_FQ_ONE
#define QR2 PALLAS_FQ_R2
#define QM0 PALLAS_FQ_M0
#define QB PALLAS_FQ_B
#define RMOD PALLAS_FR_MOD
#define RONE PALLAS_FR_ONE
#define RR2 PALLAS_FR_R2
#define RM0 PALLAS_FR_M0
#define RBITS PALLAS_FR_BITS
#include "curve_impl.inc"
#undef PFX
#undef NQ
#undef NR

enum { ORC_BLS12_381 = 0, ORC_BN254 = 1, ORC_PALLAS = 2 };

#define DISPATCH(curve, call_bls, call_bn, call_pal) \
  switch (curve) {                                   \
    case ORC_BLS12_381: call_bls; break;             \
    case ORC_BN254: call_bn; break;                  \
    case ORC_PALLAS: call_pal; break;                \
    default: return -1;                              \
  }

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int orc_fq_limbs(int curve) { return curve == ORC_BLS12_381 ? 6 : 4; }
int orc_fr_limbs(int curve) { (void)curve; return 4; }

/* ---- scalar field / base field element-wise helpers (unit-test surface) ---- */
#define FIELD_BINOP(NAME, T, OP)                                                                   \
  int NAME(int curve, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {            \
    DISPATCH(curve,                                                                                \
             for (size_t i = 0; i < n; i++) bls_##T##_##OP((bls_##T *)out + i, (const bls_##T *)a + i, (const bls_##T *)b + i), \
             for (size_t i = 0; i < n; i++) bn_##T##_##OP((bn_##T *)out + i, (const bn_##T *)a + i, (const bn_##T *)b + i),    \
             for (size_t i = 0; i < n; i++) pal_##T##_##OP((pal_##T *)out + i, (const pal_##T *)a + i, (const pal_##T *)b + i)) \
    return 0;                                                                                      \
  }
FIELD_BINOP(orc_fq_mul, fq, mul)
FIELD_BINOP(orc_fq_add, fq, add)
FIELD_BINOP(orc_fq_sub, fq, sub)
FIELD_BINOP(orc_fr_mul, fr, mul)
FIELD_BINOP(orc_fr_add, fr, add)
FIELD_BINOP(orc_fr_sub, fr, sub)

#define FIELD_UNOP(NAME, T, OP)                                                                    \
  int NAME(int curve, const uint64_t *a, uint64_t *out, size_t n) {                               \
    DISPATCH(curve,                                                                                \
             for (size_t i = 0; i < n; i++) bls_##T##_##OP((bls_##T *)out + i, (const bls_##T *)a + i), \
             for (size_t i = 0; i < n; i++) bn_##T##_##OP((bn_##T *)out + i, (const bn_##T *)a + i),    \
             for (size_t i = 0; i < n; i++) pal_##T##_##OP((pal_##T *)out + i, (const pal_##T *)a + i)) \
    return 0;                                                                                      \
  }
/* F::into_bigint -- kzg10/mod.rs:463-470 convert_to_bigints; ipa_pc/mod.rs:60-62; hyrax/mod.rs:88-90 */
FIELD_UNOP(orc_fr_from_mont, fr, from_mont)
FIELD_UNOP(orc_fr_to_mont, fr, to_mont)
FIELD_UNOP(orc_fq_from_mont, fq, from_mont)
FIELD_UNOP(orc_fq_to_mont, fq, to_mont)
FIELD_UNOP(orc_fq_inv, fq, inv)
FIELD_UNOP(orc_fr_inv, fr, inv)

/* ---- group helpers ---- */
int orc_g1_on_curve(int curve, const uint64_t *xy, const uint8_t *inf, size_t n) {
  int bad = 0;
#define BODY(P) for (size_t i = 0; i < n; i++) { P##_aff a; P##_aff_load(&a, xy, inf, i); bad += !P##_aff_on_curve(&a); }
  DISPATCH(curve, BODY(bls), BODY(bn), BODY(pal))
#undef BODY
  return bad;
}

int orc_g1_generator(int curve, uint64_t *xy) {
  switch (curve) {
    case ORC_BLS12_381: memcpy(xy, BLS12_381_FQ_GX, 48); memcpy(xy + 6, BLS12_381_FQ_GY, 48); break;
    case ORC_BN254: memcpy(xy, BN254_FQ_GX, 32); memcpy(xy + 4, BN254_FQ_GY, 32); break;
    case ORC_PALLAS: memcpy(xy, PALLAS_FQ_GX, 32); memcpy(xy + 4, PALLAS_FQ_GY, 32); break;
    default: return -1;
  }
  return 0;
}

/* out = k * P, k canonical */
int orc_g1_mul(int curve, const uint64_t *p_xy, const uint8_t *p_inf, const uint64_t *k, uint64_t *out_xy, uint8_t *out_inf) {
#define BODY(P) { P##_aff a, o; P##_aff_load(&a, p_xy, p_inf, 0); P##_jac j; P##_jac_mul(&j, &a, k, 4); P##_jac_to_aff(&o, &j); P##_aff_store(&o, out_xy, out_inf, 0); }
  DISPATCH(curve, BODY(bls), BODY(bn), BODY(pal))
#undef BODY
  return 0;
}

/* out = sum of n affine points (used to check the multi-GPU point-sum) */
int orc_g1_sum(int curve, const uint64_t *xy, const uint8_t *inf, size_t n, uint64_t *out_xy, uint8_t *out_inf) {
#define BODY(P) { P##_jac acc; P##_jac_set_inf(&acc); for (size_t i = 0; i < n; i++) { P##_aff a; P##_aff_load(&a, xy, inf, i); P##_jac_add_aff(&acc, &acc, &a, 0); } \
                  P##_aff o; P##_jac_to_aff(&o, &acc); P##_aff_store(&o, out_xy, out_inf, 0); }
  DISPATCH(curve, BODY(bls), BODY(bn), BODY(pal))
#undef BODY
  return 0;
}

/* key_l[i] += chal * key_r[i]; normalize_batch  (ipa_pc/mod.rs:699-707).  key: 2m affine points, chal canonical;
 * out: m affine points. */
int orc_g1_fold(int curve, const uint64_t *key_xy, size_t m, const uint64_t *chal, uint64_t *out_xy) {
#define BODY(P) { for (size_t i = 0; i < m; i++) { P##_aff l, r, o; P##_aff_load(&l, key_xy, NULL, i); P##_aff_load(&r, key_xy, NULL, m + i); \
      if (P##_fq_is_zero(&l.x) && P##_fq_is_zero(&l.y)) l.inf = 1; if (P##_fq_is_zero(&r.x) && P##_fq_is_zero(&r.y)) r.inf = 1; \
      P##_jac j; P##_jac_mul(&j, &r, chal, 4); P##_jac_add_aff(&j, &j, &l, 0); P##_jac_to_aff(&o, &j); P##_aff_store(&o, out_xy, NULL, i); } }
  DISPATCH(curve, BODY(bls), BODY(bn), BODY(pal))
#undef BODY
  return 0;
}

/* VariableBaseMSM::msm_bigint restated at definition level (sum of double-and-add products) */
int orc_msm_naive(int curve, const uint64_t *bases, const uint8_t *inf, const uint64_t *scalars, size_t n,
                  uint64_t *out_xy, uint8_t *out_inf) {
  DISPATCH(curve, bls_msm_naive(bases, inf, scalars, n, out_xy, out_inf), bn_msm_naive(bases, inf, scalars, n, out_xy, out_inf),
           pal_msm_naive(bases, inf, scalars, n, out_xy, out_inf))
  return 0;
}

/* VariableBaseMSM::msm_bigint, Pippenger structure -- kzg10/mod.rs:175-178, :255-258; ipa_pc/mod.rs:64; hyrax/mod.rs:92 */
int orc_msm_pippenger(int curve, const uint64_t *bases, const uint8_t *inf, const uint64_t *scalars, size_t n,
                      uint64_t *out_xy, uint8_t *out_inf, int nthreads) {
  if (nthreads <= 0) nthreads = orc_num_threads();
  if (n == 0) { /* empty MSM = identity (kzg10/mod.rs:197-203 non-hiding call) */
    memset(out_xy, 0, (size_t)orc_fq_limbs(curve) * 16); if (out_inf) *out_inf = 1; return 0; }
  DISPATCH(curve, bls_msm_pippenger(bases, inf, scalars, n, out_xy, out_inf, nthreads),
           bn_msm_pippenger(bases, inf, scalars, n, out_xy, out_inf, nthreads),
           pal_msm_pippenger(bases, inf, scalars, n, out_xy, out_inf, nthreads))
  return 0;
}

/* g.batch_mul(&scalars) -- kzg10/mod.rs:76 (setup); used to make synthetic SRSs */
int orc_fixed_base_batch_mul(int curve, const uint64_t *base_xy, const uint64_t *scalars, size_t n,
                             uint64_t *out_xy, uint8_t *out_inf, int nthreads) {
  if (nthreads <= 0) nthreads = orc_num_threads();
  DISPATCH(curve, bls_fixed_base_batch_mul(base_xy, scalars, n, out_xy, out_inf, nthreads),
           bn_fixed_base_batch_mul(base_xy, scalars, n, out_xy, out_inf, nthreads),
           pal_fixed_base_batch_mul(base_xy, scalars, n, out_xy, out_inf, nthreads))
  return 0;
}

/* ---- Fr vector work around the MSM ---- */

/* powers_of_beta: out[i] = beta^i, CANONICAL form (kzg10/mod.rs:66-73 setup) ; beta Montgomery */
int orc_fr_powers_canonical(int curve, const uint64_t *beta, size_t n, uint64_t *out) {
#define BODY(P) { P##_fr cur, b; memcpy(&b, beta, 32); P##_fr_set_one(&cur); \
    for (size_t i = 0; i < n; i++) { P##_fr_from_mont((P##_fr *)out + i, &cur); P##_fr_mul(&cur, &cur, &b); } }
  DISPATCH(curve, BODY(bls), BODY(bn), BODY(pal))
#undef BODY
  return 0;
}

/* y += c * x  -- DensePolynomial AddAssign<(F,&P)>, marlin_pc/mod.rs:286; ipa_pc/mod.rs:691-697 */
int orc_fr_axpy(int curve, uint64_t *y, const uint64_t *c, const uint64_t *x, size_t n) {
#define BODY(P) { P##_fr cc, t; memcpy(&cc, c, 32); for (size_t i = 0; i < n; i++) { P##_fr_mul(&t, &cc, (const P##_fr *)x + i); P##_fr_add((P##_fr *)y + i, (P##_fr *)y + i, &t); } }
  DISPATCH(curve, BODY(bls), BODY(bn), BODY(pal))
#undef BODY
  return 0;
}

/* quotient of p(X) by (X - z): q[n-2] = p[n-1]; q[i-1] = p[i] + z*q[i]   (kzg10/mod.rs:222-226).
 * p has n coefficients (low degree first), q receives n-1; *rem (optional) receives p(z). */
int orc_fr_div_linear(int curve, const uint64_t *p, size_t n, const uint64_t *z, uint64_t *q, uint64_t *rem) {
  if (n == 0) return 0;
#define BODY(P) { P##_fr zz, carry, t; memcpy(&zz, z, 32); memset(&carry, 0, 32);              \
    for (size_t i = n; i-- > 0;) { /* carry = q[i] (zero above the top) */                      \
      P##_fr_mul(&t, &zz, &carry); P##_fr_add(&t, &t, (const P##_fr *)p + i);                    \
      if (i > 0) ((P##_fr *)q)[i - 1] = t; else if (rem) memcpy(rem, &t, 32);                    \
      carry = t; } }
  DISPATCH(curve, BODY(bls), BODY(bn), BODY(pal))
#undef BODY
  return 0;
}

/* Horner evaluation p(z)  (Polynomial::evaluate; ipa_pc/mod.rs:561, kzg10/mod.rs:264) */
int orc_fr_eval(int curve, const uint64_t *p, size_t n, const uint64_t *z, uint64_t *out) {
#define BODY(P) { P##_fr zz, acc; memcpy(&zz, z, 32); memset(&acc, 0, 32); \
    for (size_t i = n; i-- > 0;) { P##_fr_mul(&acc, &acc, &zz); P##_fr_add(&acc, &acc, (const P##_fr *)p + i); } memcpy(out, &acc, 32); }
  DISPATCH(curve, BODY(bls), BODY(bn), BODY(pal))
#undef BODY
  return 0;
}

/* <a, b>  -- utils.rs:150-155 inner_product */
int orc_fr_inner_product(int curve, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out) {
#define BODY(P) { P##_fr acc, t; memset(&acc, 0, 32); for (size_t i = 0; i < n; i++) { P##_fr_mul(&t, (const P##_fr *)a + i, (const P##_fr *)b + i); P##_fr_add(&acc, &acc, &t); } memcpy(out, &acc, 32); }
  DISPATCH(curve, BODY(bls), BODY(bn), BODY(pal))
#undef BODY
  return 0;
}

/* v * M  (row vector times matrix; M is rows x cols row-major) -- utils.rs:127-146 Matrix::row_mul */
int orc_fr_row_mul(int curve, const uint64_t *v, const uint64_t *m, size_t rows, size_t cols, uint64_t *out) {
#define BODY(P) { for (size_t c = 0; c < cols; c++) { P##_fr acc, t; memset(&acc, 0, 32);                 \
      for (size_t r = 0; r < rows; r++) { P##_fr_mul(&t, (const P##_fr *)v + r, (const P##_fr *)m + r * cols + c); P##_fr_add(&acc, &acc, &t); } \
      memcpy(out + 4 * c, &acc, 32); } }
  DISPATCH(curve, BODY(bls), BODY(bn), BODY(pal))
#undef BODY
  return 0;
}

/* Domain generator for size 2^logn: root_of_unity^(2^(two_adicity-logn)) [ark-dep, from memory];
 * out Montgomery. */
int orc_fr_domain_generator(int curve, int logn, uint64_t *out) {
#define BODY(P, ROOT, TA) { if (logn > TA) return -2; P##_fr w; memcpy(&w, ROOT, 32); for (int i = logn; i < TA; i++) P##_fr_sqr(&w, &w); memcpy(out, &w, 32); }
  DISPATCH(curve, BODY(bls, BLS12_381_FR_ROOT, BLS12_381_FR_TWO_ADICITY), BODY(bn, BN254_FR_ROOT, BN254_FR_TWO_ADICITY),
           BODY(pal, PALLAS_FR_ROOT, PALLAS_FR_TWO_ADICITY))
#undef BODY
  return 0;
}

/* EvaluationDomain::fft semantics at its only call site (linear_codes/utils.rs:112-127):
 * zero-pad the n_in coefficients to N = 2^logn, out[j] = p(w^j), natural order.
 * Definition-level O(N * n_in) Horner per output (small N only). */
int orc_fr_ntt_naive(int curve, const uint64_t *in, size_t n_in, int logn, uint64_t *out) {
  uint64_t w[4];
  int rc = orc_fr_domain_generator(curve, logn, w);
  if (rc) return rc;
  size_t N = (size_t)1 << logn;
#define BODY(P) { P##_fr ww, x; memcpy(&ww, w, 32); P##_fr_set_one(&x);                         \
    for (size_t j = 0; j < N; j++) { orc_fr_eval(curve, in, n_in, x.l, out + 4 * j); P##_fr_mul(&x, &x, &ww); } }
  DISPATCH(curve, BODY(bls), BODY(bn), BODY(pal))
#undef BODY
  return 0;
}

/* Same transform, O(N log N): recursive decimation-in-time radix-2 (textbook Cooley-Tukey). */
#define DEF_NTT_REC(P)                                                                                 \
  static void P##_ntt_rec(P##_fr *a, size_t n, const P##_fr *w, P##_fr *tmp) {                          \
    if (n == 1) return;                                                                                \
    size_t h = n / 2;                                                                                  \
    for (size_t i = 0; i < h; i++) { tmp[i] = a[2 * i]; tmp[h + i] = a[2 * i + 1]; }                   \
    memcpy(a, tmp, n * sizeof *a);                                                                     \
    P##_fr w2; P##_fr_sqr(&w2, w);                                                                     \
    P##_ntt_rec(a, h, &w2, tmp); P##_ntt_rec(a + h, h, &w2, tmp);                                      \
    P##_fr x, t; P##_fr_set_one(&x);                                                                   \
    for (size_t i = 0; i < h; i++) {                                                                   \
      P##_fr_mul(&t, &x, &a[h + i]); P##_fr e = a[i];                                                  \
      P##_fr_add(&a[i], &e, &t); P##_fr_sub(&a[h + i], &e, &t); P##_fr_mul(&x, &x, w); } }
DEF_NTT_REC(bls)
DEF_NTT_REC(bn)
DEF_NTT_REC(pal)

int orc_fr_ntt(int curve, const uint64_t *in, size_t n_in, int logn, uint64_t *out) {
  uint64_t w[4];
  int rc = orc_fr_domain_generator(curve, logn, w);
  if (rc) return rc;
  size_t N = (size_t)1 << logn;
  if (n_in > N) return -3;
  memset(out, 0, N * 32); memcpy(out, in, n_in * 32);
  void *tmp = malloc(N * 32);
#define BODY(P) { P##_fr ww; memcpy(&ww, w, 32); P##_ntt_rec((P##_fr *)out, N, &ww, (P##_fr *)tmp); }
  switch (curve) { case ORC_BLS12_381: BODY(bls) break; case ORC_BN254: BODY(bn) break; case ORC_PALLAS: BODY(pal) break; default: free(tmp); return -1; }
#undef BODY
  free(tmp);
  return 0;
}

/* ---- KZG10 dataflow (non-hiding and hiding) ---- */

static size_t count_leading_zero_coeffs(const uint64_t *c, size_t n) {
  size_t k = 0;
  while (k < n && !(c[4 * k] | c[4 * k + 1] | c[4 * k + 2] | c[4 * k + 3])) k++;
  return k;
}

/* KZG10::commit -- kzg10/mod.rs:157-210.  powers_of_g: n_powers affine points; coeffs: n Fr (Montgomery).
 * Hiding part: blinding coefficients are an INPUT (n_blind of them, may be 0; the reference samples
 * them from its RNG, :182-195) committed over powers_of_gamma_g (:199-203) and added (:206).
 * Returns -4 when the degree is too large (check_degree_is_too_large, :163). */
int orc_kzg_commit(int curve, const uint64_t *powers_of_g, size_t n_powers, const uint64_t *coeffs, size_t n,
                   const uint64_t *powers_of_gamma_g, size_t n_gamma, const uint64_t *blind, size_t n_blind,
                   uint64_t *out_xy, uint8_t *out_inf, int nthreads) {
  while (n > 0 && !(coeffs[4 * (n - 1)] | coeffs[4 * (n - 1) + 1] | coeffs[4 * (n - 1) + 2] | coeffs[4 * (n - 1) + 3])) n--; /* degree() */
  if (n > n_powers) return -4;
  if (n_blind > n_gamma) return -5;
  size_t lz = count_leading_zero_coeffs(coeffs, n); /* skip_leading_zeros_and_convert_to_bigints :452-461 */
  size_t m = n - lz;
  int nq = orc_fq_limbs(curve);
  uint64_t *ints = (uint64_t *)malloc((m + n_blind + 1) * 32);
  orc_fr_from_mont(curve, coeffs + 4 * lz, ints, m);
  uint64_t pts[2 * 12]; uint8_t infs[2];
  orc_msm_pippenger(curve, powers_of_g + (size_t)2 * nq * lz, NULL, ints, m, pts, &infs[0], nthreads);
  orc_fr_from_mont(curve, blind, ints, n_blind);
  orc_msm_pippenger(curve, powers_of_gamma_g, NULL, ints, n_blind, pts + 2 * nq, &infs[1], nthreads);
  free(ints);
  return orc_g1_sum(curve, pts, infs, 2, out_xy, out_inf);
}

/* KZG10::open -- kzg10/mod.rs:287-310 -> compute_witness_polynomial :217-240 -> open_with_witness_polynomial :243-284.
 * Hiding: blind (n_blind coefficients) is the Randomness' blinding polynomial; random_v = blind(z) (:264). */
int orc_kzg_open(int curve, const uint64_t *powers_of_g, size_t n_powers, const uint64_t *coeffs, size_t n,
                 const uint64_t *z, const uint64_t *powers_of_gamma_g, size_t n_gamma, const uint64_t *blind, size_t n_blind,
                 uint64_t *out_w_xy, uint8_t *out_w_inf, uint64_t *out_random_v, int nthreads) {
  if (n > n_powers + 1 && n > 0) { /* degree check on p, :292 */ }
  while (n > 0 && !(coeffs[4 * (n - 1)] | coeffs[4 * (n - 1) + 1] | coeffs[4 * (n - 1) + 2] | coeffs[4 * (n - 1) + 3])) n--;
  if (n > n_powers) return -4;
  int nq = orc_fq_limbs(curve);
  size_t nw = n > 0 ? n - 1 : 0;
  uint64_t *wit = (uint64_t *)malloc((nw + 1) * 32);
  orc_fr_div_linear(curve, coeffs, n, z, wit, NULL);
  while (nw > 0 && !(wit[4 * (nw - 1)] | wit[4 * (nw - 1) + 1] | wit[4 * (nw - 1) + 2] | wit[4 * (nw - 1) + 3])) nw--;
  size_t lz = count_leading_zero_coeffs(wit, nw);
  size_t m = nw - lz;
  uint64_t *ints = (uint64_t *)malloc((m + n_blind + 1) * 32);
  orc_fr_from_mont(curve, wit + 4 * lz, ints, m);
  uint64_t pts[2 * 12]; uint8_t infs[2];
  orc_msm_pippenger(curve, powers_of_g + (size_t)2 * nq * lz, NULL, ints, m, pts, &infs[0], nthreads);
  size_t nbw = 0;
  if (n_blind > 0) {
    if (n_blind > n_gamma + 1) { free(ints); free(wit); return -5; }
    nbw = n_blind - 1;
    uint64_t *bw = (uint64_t *)malloc((nbw + 1) * 32);
    orc_fr_div_linear(curve, blind, n_blind, z, bw, NULL);
    orc_fr_from_mont(curve, bw, ints, nbw);
    free(bw);
    if (out_random_v) orc_fr_eval(curve, blind, n_blind, z, out_random_v);
  }
  orc_msm_pippenger(curve, powers_of_gamma_g, NULL, ints, nbw, pts + 2 * nq, &infs[1], nthreads);
  free(ints); free(wit);
  return orc_g1_sum(curve, pts, infs, 2, out_w_xy, out_w_inf);
}
