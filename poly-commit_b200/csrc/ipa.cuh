// Device-resident state and kernels for the InnerProductArgPC::open halving loop
// (ipa_pc/mod.rs:665-711): per round two MSMs over the current (folded) key halves, two inner products,
// the folds  coeffs_l += chal^-1 coeffs_r,  z_l += chal z_r,  key_l[i] += chal * key_r[i]  followed by
// normalize_batch (:699-707).  The Fiat-Shamir hash between rounds stays on the host (it consumes two affine
// points and produces one scalar); everything O(n) stays in HBM across the 18 rounds of cfg3.
#pragma once
#include "frops.cuh"
#include "msm.cuh"

namespace pcgpu {

// out[i] = z^i (ipa_pc/mod.rs:643-648)
template <class R>
struct FrPowersBody {
  const uint32_t *z; uint32_t *out;
  PCGPU_KERNEL_DEV void operator()(size_t i) const {
    Fp<R> acc = Fp<R>::one(), base = load_fr<R>(z, 0);
    size_t e = i;
    while (e) { if (e & 1) acc = fp_mul<R>(acc, base); base = fp_sqr<R>(base); e >>= 1; }
    store_fr<R>(out, i, acc);
  }
};

// key[i] = affine(key[i] + chal * key[i + m])   -- k_l += k_r.mul(chal); normalize_batch  (:699-707)
template <class C>
struct G1FoldBody {
  Affine<C> *key; uint32_t m; uint32_t chal[8];  // canonical scalar, same for every point
  PCGPU_KERNEL_DEV void operator()(size_t i) const {
    Affine<C> r = load_affine<C>(key + m + i);
    XYZZ<C> acc = XYZZ<C>::inf();
    for (int b = 255; b >= 0; b--) {
      acc = xyzz_dbl<C>(acc);
      if ((chal[b >> 5] >> (b & 31)) & 1) xyzz_madd<C>(acc, r, false);
    }
    Affine<C> l = load_affine<C>(key + i);
    xyzz_madd<C>(acc, l, false);
    key[i] = xyzz_to_affine<C>(acc);
  }
};

// SuccinctCheckPolynomial::compute_coeffs (ipa_pc/data_structures.rs:204-220): coeffs[idx] = product of challenge_i over
// the set bits of idx, challenge_1 on the top bit.  One thread per coefficient (<= log_d multiplications).
template <class R>
struct FrCheckCoeffsBody {
  const uint32_t *challenges; uint32_t log_d; uint32_t *out;
  PCGPU_KERNEL_DEV void operator()(size_t idx) const {
    Fp<R> acc = Fp<R>::one();
    for (uint32_t i = 1; i <= log_d; i++)
      if ((idx >> (log_d - i)) & 1) acc = fp_mul<R>(acc, load_fr<R>(challenges, i - 1));
    store_fr<R>(out, idx, acc);
  }
};

}  // namespace pcgpu
