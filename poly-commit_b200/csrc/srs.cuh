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
  if (const char *e = getenv("PCGPU_SRS_C")) { int v = atoi(e); if (v >= 8 && v <= 22) return (uint32_t)v; }   // tuning knob
  uint32_t lg = ilog2_floor(n ? n : 1);
  if (lg >= 18) return 16;   // c = 17 (15 windows) measured equal on the pair rounds and slower in scan / reduce (2^16 buckets): profiles/r02_msm_ab_*
  if (lg >= 15) return 14;
  return 12;
}

// rows of the base array whose ABI infinity byte is set are zeroed: (0, 0) is the device's identity encoding (ec.cuh)
struct SrsZeroIdentityBody {
  uint32_t *tables; const uint8_t *inf; uint32_t words;
  PCGPU_KERNEL_DEV void operator()(size_t i) const {
    if (inf[i]) for (uint32_t k = 0; k < words; k++) tables[i * (size_t)words + k] = 0;
  }
};

template <class C>
struct SrsGroupsBody {   // raw bases (packed x||y) -> all W groups in the aligned table layout (group 0 = the bases themselves)
  const Affine<C> *raw; uint32_t *folded; size_t n; uint32_t c; uint32_t groups; uint32_t pt_words, y_words;
  PCGPU_KERNEL_DEV void operator()(size_t i) const {
    Affine<C> a = load_affine<C>(raw + i);
    store_table_point<C>(folded, i, pt_words, y_words, a);
    for (uint32_t k = 1; k < groups; k++) {
      XYZZ<C> p = xyzz_dbl_affine<C>(a);
      for (uint32_t j = 1; j < c; j++) p = xyzz_dbl<C>(p);
      a = xyzz_to_affine<C>(p);
      store_table_point<C>(folded, (size_t)k * n + i, pt_words, y_words, a);
    }
  }
};

template <class C>
inline int srs_build_groups(const Affine<C> *raw, uint32_t *folded, size_t n, uint32_t c, uint32_t groups, rt::stream_t st) {
  return rt::launch<128>(SrsGroupsBody<C>{raw, folded, n, c, groups, aligned_pt_words<C>(), aligned_y_words<C>()}, n, st);
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
    load_scalar_plain<C>(scalars, i, false, k);
    XYZZ<C> acc = XYZZ<C>::inf();
    for (uint32_t j = 0; j < 64; j++) {
      uint32_t d = (k[j >> 3] >> ((j & 7) * 4)) & 15;
      if (d) { Affine<C> a = load_affine<C>(table + j * 15 + (d - 1)); xyzz_madd<C>(acc, a, false); }
    }
    out[i] = xyzz_to_affine<C>(acc);
  }
};

// ---------------------------------------------------------------------------------------------
// Batched MSM over SHARED bases (HyraxPC::commit: one Pedersen commitment per matrix row, all rows over the
// same com_key -- hyrax/mod.rs:233-242): fixed-base comb.  comb[(j*W + w)*NBk + d-1] = d * 2^(c w) * G_j for
// d in 1..2^(c-1), so a row costs n*W mixed additions and no bucket work at all.  Signed digits halve the table.
// ---------------------------------------------------------------------------------------------
struct CombGeom {
  uint32_t n_bases, c, W, NBk;  // NBk = 2^(c-1) table entries per (base, window)
  uint32_t n, count;            // row length, number of rows
  uint32_t segs, seg_len;       // each row is split into `segs` segments of `seg_len` scalars
  uint32_t scalar_bits, scalars_mont;
};

template <class C>
struct CombTableBody {
  const Affine<C> *bases; CombGeom g; Affine<C> *table;
  PCGPU_KERNEL_DEV void operator()(size_t t) const {
    uint32_t j = (uint32_t)(t / g.W), w = (uint32_t)(t % g.W);
    XYZZ<C> p = xyzz_from_affine<C>(load_affine<C>(bases + j));
    for (uint32_t k = 0; k < w * g.c; k++) p = xyzz_dbl<C>(p);
    XYZZ<C> acc = p;
    Affine<C> *row = table + t * (size_t)g.NBk;
    for (uint32_t d = 1; d <= g.NBk; d++) {
      row[d - 1] = xyzz_to_affine<C>(acc);
      xyzz_add<C>(acc, p);
    }
  }
};

template <class C>
struct CombAccumulateBody {
  const Affine<C> *table; const uint32_t *scalars; CombGeom g; XYZZ<C> *partial; uint32_t *err;
  PCGPU_KERNEL_DEV void operator()(size_t t) const {
    uint32_t row = (uint32_t)(t / g.segs), seg = (uint32_t)(t % g.segs);
    uint32_t lo = seg * g.seg_len, hi = lo + g.seg_len < g.n ? lo + g.seg_len : g.n;
    MsmGeom mg; mg.c = g.c; mg.W = g.W;
    XYZZ<C> acc = XYZZ<C>::inf();
    const Affine<C> *tab = table; const CombGeom gg = g;
    for (uint32_t i = lo; i < hi; i++) {
      uint32_t k[8];
      load_scalar_plain<C>(scalars, (size_t)row * g.n + i, g.scalars_mont != 0, k);
      if (!scalar_in_range(k, g.scalar_bits)) { rt::atomic_or(err, 1u); continue; }
      for_each_digit(k, mg, [&](uint32_t w, uint32_t mag, bool neg) {
        Affine<C> a = load_affine<C>(tab + ((size_t)i * gg.W + w) * gg.NBk + (mag - 1));
        xyzz_madd<C>(acc, a, neg);
      });
    }
    store_xyzz<C>(partial + t, acc);
  }
};

template <class C>
struct CombRowSumBody {
  const XYZZ<C> *partial; uint32_t segs; Affine<C> *out;
  PCGPU_KERNEL_DEV void operator()(size_t row) const {
    XYZZ<C> acc = XYZZ<C>::inf();
    for (uint32_t s = 0; s < segs; s++) { XYZZ<C> p = load_xyzz<C>(partial + row * segs + s); xyzz_add<C>(acc, p); }
    out[row] = xyzz_to_affine<C>(acc);
  }
};

}  // namespace pcgpu
