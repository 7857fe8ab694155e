// SRS-side kernels: window-folding tables built once at registration, and fixed-base batch
// multiplication (KZG10::setup's g.batch_mul(&powers_of_beta), kzg10/mod.rs:76 -- used here to
// produce synthetic SRSs on the device for tests and benchmarks).
//
// Table group k holds 2^(c*k) * P_i (affine).  With all W groups present every Pippenger window
// shares ONE bucket set (window w of scalar i adds group-w's copy of base i), so the MSM tail needs
// no doublings and the bucket reduction shrinks from W*2^(c-1) to 2^(c-1) buckets.  Cost: W x the
// SRS footprint (1.6 GB for 2^20 BLS12-381 bases at c = 16 -- small against 180 GB of HBM3e).
#pragma once
#include "ec.cuh"
#include "msm.cuh"
#include "rt.cuh"

namespace pcgpu {

enum : size_t { SRS_PRECOMPUTE_MIN_N = 1u << 12 };

inline uint32_t srs_precompute_window(size_t n) {
  uint32_t lg = ilog2_floor(n ? n : 1);
  if (lg >= 18) return 16;
  if (lg >= 15) return 14;
  return 12;
}

template <class C>
struct SrsGroupsBody {
  Affine<C> *tables; size_t n; uint32_t c; uint32_t groups;
  PCGPU_KERNEL_DEV void operator()(size_t i) const {
    Affine<C> a = load_affine<C>(tables + i);
    for (uint32_t k = 1; k < groups; k++) {
      XYZZ<C> p = xyzz_dbl_affine<C>(a);
      for (uint32_t j = 1; j < c; j++) p = xyzz_dbl<C>(p);
      a = xyzz_to_affine<C>(p);
      tables[(size_t)k * n + i] = a;
    }
  }
};

template <class C>
inline int srs_build_groups(Affine<C> *tables, size_t n, uint32_t c, uint32_t groups, rt::Arena &, rt::stream_t st) {
  return rt::launch<128>(SrsGroupsBody<C>{tables, n, c, groups}, n, st);
}

// out[i] = k_i * P for canonical scalars k_i (thread per scalar, 4-bit fixed window over a table of
// 1..15 multiples of 2^(4j) P built by the same kernel family).
template <class C>
struct FixedBaseTableBody {  // thread j builds row j: {d * 16^j * P : d = 1..15}
  Affine<C> base; Affine<C> *table;
  PCGPU_KERNEL_DEV void operator()(size_t j) const {
    XYZZ<C> w = xyzz_from_affine<C>(base);
    for (size_t k = 0; k < 4 * j; k++) w = xyzz_dbl<C>(w);
    XYZZ<C> acc = w;
    for (uint32_t d = 1; d <= 15; d++) {
      table[j * 15 + (d - 1)] = xyzz_to_affine<C>(acc);
      xyzz_add<C>(acc, w);
    }
  }
};

template <class C>
struct FixedBaseMulBody {
  const Affine<C> *table; const uint32_t *scalars; Affine<C> *out;
  PCGPU_KERNEL_DEV void operator()(size_t i) const {
    uint32_t k[8];
    load_scalar<C>(scalars, i, false, k);
    XYZZ<C> acc = XYZZ<C>::inf();
    for (uint32_t j = 0; j < 64; j++) {
      uint32_t d = (k[j >> 3] >> ((j & 7) * 4)) & 15;
      if (d) { Affine<C> a = load_affine<C>(table + j * 15 + (d - 1)); xyzz_madd<C>(acc, a, false); }
    }
    out[i] = xyzz_to_affine<C>(acc);
  }
};

}  // namespace pcgpu
