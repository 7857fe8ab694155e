// Scalar-field (Fr) vector kernels that surround the MSM on the reference's prover paths.
// All elements are Montgomery-form 8 x uint32 (32 bytes), natural order, unless stated.
//
//   fr_from_mont      F::into_bigint over a slice        kzg10/mod.rs:463-470 (convert_to_bigints)
//   fr_axpy           p += (c, q)                        marlin_pc/mod.rs:286; ipa_pc/mod.rs:691-697
//   fr_div_linear     p / (X - z), remainder p(z)        kzg10/mod.rs:222-226 (compute_witness_polynomial)
//   fr_inner_product  <a, b>                             utils.rs:150-155
//   fr_row_mul        v * M                              utils.rs:127-146 (Matrix::row_mul)
//
// These are the HBM-bound members of the path (64 B/elem for conversion and
// division, 96 B/elem for axpy).
#pragma once
#include "ec.cuh"
#include "msm.cuh"
#include "rt.cuh"

namespace pcgpu {

template <class R>
PCGPU_DEV Fp<R> load_fr(const uint32_t *base, size_t i) {
  const u32x4 *p = reinterpret_cast<const u32x4 *>(base) + 2 * i;
  u32x4 lo = p[0], hi = p[1];
  Fp<R> v;
  v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w;
  v.l[4] = hi.x; v.l[5] = hi.y; v.l[6] = hi.z; v.l[7] = hi.w;
  return v;
}
template <class R>
PCGPU_DEV void store_fr(uint32_t *base, size_t i, const Fp<R> &v) {
  u32x4 *p = reinterpret_cast<u32x4 *>(base) + 2 * i;
  u32x4 lo, hi;
  lo.x = v.l[0]; lo.y = v.l[1]; lo.z = v.l[2]; lo.w = v.l[3];
  hi.x = v.l[4]; hi.y = v.l[5]; hi.z = v.l[6]; hi.w = v.l[7];
  p[0] = lo; p[1] = hi;
}

template <class R>
struct FrFromMontBody {
  const uint32_t *in; uint32_t *out;
  PCGPU_KERNEL_DEV void operator()(size_t i) const { store_fr<R>(out, i, fp_from_mont<R>(load_fr<R>(in, i))); }
};
template <class R>
struct FrToMontBody {
  const uint32_t *in; uint32_t *out;
  PCGPU_KERNEL_DEV void operator()(size_t i) const { store_fr<R>(out, i, fp_to_mont<R>(load_fr<R>(in, i))); }
};

// y[i] += c * x[i]
template <class R>
struct FrAxpyBody {
  uint32_t *y; const uint32_t *c; const uint32_t *x;
  PCGPU_KERNEL_DEV void operator()(size_t i) const {
    Fp<R> cc = load_fr<R>(c, 0);
    store_fr<R>(y, i, fp_add<R>(load_fr<R>(y, i), fp_mul<R>(cc, load_fr<R>(x, i))));
  }
};

// ---------------------------------------------------------------------------------------------
// Division by (X - z) as a three-level scan of the affine maps  t -> p_i + z * t.
//   level 1: chunks of DIV_K coefficients  (thread per chunk)  -> local Horner value
//   level 2: groups of DIV_G chunks        (thread per group)
//   level 3: one thread walks the groups
// then the carries are pushed back down and every chunk replays its recurrence writing q.
// q[i-1] = p[i] + z*q[i]; carry into a chunk = q[hi-1] with q[n-1] := 0; remainder = p(z).
// ---------------------------------------------------------------------------------------------
enum { DIV_K = 32, DIV_G = 64 };

template <class R>
struct DivPowersBody {  // zp[0] = z^DIV_K, zp[1] = z^(DIV_K*DIV_G)
  const uint32_t *z; uint32_t *zp;
  PCGPU_KERNEL_DEV void operator()(size_t) const {
    Fp<R> a = load_fr<R>(z, 0);
    for (int k = DIV_K; k > 1; k >>= 1) a = fp_sqr<R>(a);
    store_fr<R>(zp, 0, a);
    for (int k = DIV_G; k > 1; k >>= 1) a = fp_sqr<R>(a);
    store_fr<R>(zp, 1, a);
  }
};

template <class R>
struct DivChunkLocalBody {
  const uint32_t *p; size_t n; const uint32_t *z; uint32_t *local;
  PCGPU_KERNEL_DEV void operator()(size_t c) const {
    size_t lo = c * DIV_K, hi = lo + DIV_K < n ? lo + DIV_K : n;
    Fp<R> zz = load_fr<R>(z, 0), acc = Fp<R>::zero();
    for (size_t i = hi; i-- > lo;) acc = fp_add<R>(fp_mul<R>(acc, zz), load_fr<R>(p, i));
    store_fr<R>(local, c, acc);
  }
};

// group local value: fold chunk locals from the top chunk of the group down
template <class R>
struct DivGroupLocalBody {
  const uint32_t *local; size_t nchunks; const uint32_t *zp; uint32_t *glocal;
  PCGPU_KERNEL_DEV void operator()(size_t g) const {
    size_t lo = g * DIV_G, hi = lo + DIV_G < nchunks ? lo + DIV_G : nchunks;
    Fp<R> zk = load_fr<R>(zp, 0), acc = Fp<R>::zero();
    for (size_t c = hi; c-- > lo;) acc = fp_add<R>(fp_mul<R>(acc, zk), load_fr<R>(local, c));
    store_fr<R>(glocal, g, acc);
  }
};

// carries into groups (gcarry[g] = value entering group g from above); rem = value leaving group 0
template <class R>
struct DivGroupCarryBody {
  const uint32_t *glocal; size_t ngroups; const uint32_t *zp; uint32_t *gcarry; uint32_t *rem;
  PCGPU_KERNEL_DEV void operator()(size_t) const {
    Fp<R> zg = load_fr<R>(zp, 1), t = Fp<R>::zero();
    for (size_t g = ngroups; g-- > 0;) {
      store_fr<R>(gcarry, g, t);
      t = fp_add<R>(fp_mul<R>(t, zg), load_fr<R>(glocal, g));
    }
    store_fr<R>(rem, 0, t);
  }
};

// carries into chunks
template <class R>
struct DivChunkCarryBody {
  const uint32_t *local; size_t nchunks; const uint32_t *zp; const uint32_t *gcarry; uint32_t *ccarry;
  PCGPU_KERNEL_DEV void operator()(size_t g) const {
    size_t lo = g * DIV_G, hi = lo + DIV_G < nchunks ? lo + DIV_G : nchunks;
    Fp<R> zk = load_fr<R>(zp, 0), t = load_fr<R>(gcarry, g);
    for (size_t c = hi; c-- > lo;) {
      store_fr<R>(ccarry, c, t);
      t = fp_add<R>(fp_mul<R>(t, zk), load_fr<R>(local, c));
    }
  }
};

template <class R>
struct DivChunkWriteBody {
  const uint32_t *p; size_t n; const uint32_t *z; const uint32_t *ccarry; uint32_t *q;
  PCGPU_KERNEL_DEV void operator()(size_t c) const {
    size_t lo = c * DIV_K, hi = lo + DIV_K < n ? lo + DIV_K : n;
    Fp<R> zz = load_fr<R>(z, 0), t = load_fr<R>(ccarry, c);
    for (size_t i = hi; i-- > lo;) {
      t = fp_add<R>(fp_mul<R>(t, zz), load_fr<R>(p, i));
      if (i > 0) store_fr<R>(q, i - 1, t);
    }
  }
};

inline size_t div_scratch_words(size_t n) {
  size_t nchunks = (n + DIV_K - 1) / DIV_K, ngroups = (nchunks + DIV_G - 1) / DIV_G;
  return 8 * (2 * nchunks + 2 * ngroups + 4);
}

// p: n coefficients, q: n-1 coefficients (n >= 1), rem: 1 element (may alias scratch), z: 1 element; all device.
template <class R>
inline int fr_div_linear(const uint32_t *p, size_t n, const uint32_t *z, uint32_t *q, uint32_t *rem,
                         uint32_t *scratch, rt::stream_t st) {
  if (n == 0) return rt::dev_memset(rem, 0, 32, st);
  size_t nchunks = (n + DIV_K - 1) / DIV_K, ngroups = (nchunks + DIV_G - 1) / DIV_G;
  uint32_t *zp = scratch, *local = zp + 16, *ccarry = local + 8 * nchunks, *glocal = ccarry + 8 * nchunks,
           *gcarry = glocal + 8 * ngroups;
  int rc;
  if ((rc = rt::launch<32>(DivPowersBody<R>{z, zp}, 1, st))) return rc;
  if ((rc = rt::launch<128>(DivChunkLocalBody<R>{p, n, z, local}, nchunks, st))) return rc;
  if ((rc = rt::launch<64>(DivGroupLocalBody<R>{local, nchunks, zp, glocal}, ngroups, st))) return rc;
  if ((rc = rt::launch<32>(DivGroupCarryBody<R>{glocal, ngroups, zp, gcarry, rem}, 1, st))) return rc;
  if ((rc = rt::launch<64>(DivChunkCarryBody<R>{local, nchunks, zp, gcarry, ccarry}, ngroups, st))) return rc;
  return rt::launch<128>(DivChunkWriteBody<R>{p, n, z, ccarry, q}, nchunks, st);
}

// ---------------------------------------------------------------------------------------------
// inner product: per-thread strided partial sums, then a single-thread fold of the partials
// ---------------------------------------------------------------------------------------------
enum { IP_THREADS = 4096 };
template <class R>
struct IpPartialBody {
  const uint32_t *a; const uint32_t *b; size_t n; uint32_t *partial;
  PCGPU_KERNEL_DEV void operator()(size_t t) const {
    Fp<R> acc = Fp<R>::zero();
    for (size_t i = t; i < n; i += IP_THREADS) acc = fp_add<R>(acc, fp_mul<R>(load_fr<R>(a, i), load_fr<R>(b, i)));
    store_fr<R>(partial, t, acc);
  }
};
template <class R>
struct FrTreeAddBody {
  uint32_t *a; uint32_t m; uint32_t half;
  PCGPU_KERNEL_DEV void operator()(size_t i) const {
    store_fr<R>(a, i, fp_add<R>(load_fr<R>(a, i), load_fr<R>(a, i + half)));
  }
};
template <class R>
inline int fr_inner_product(const uint32_t *a, const uint32_t *b, size_t n, uint32_t *out, uint32_t *scratch, rt::stream_t st) {
  int rc;
  if ((rc = rt::launch<128>(IpPartialBody<R>{a, b, n, scratch}, IP_THREADS, st))) return rc;
  for (uint32_t m = IP_THREADS; m > 1;) {
    uint32_t half = (m + 1) / 2;
    if ((rc = rt::launch<128>(FrTreeAddBody<R>{scratch, m, half}, m - half, st))) return rc;
    m = half;
  }
  return rt::copy_d2d(out, scratch, 32, st);
}

// out[c] = sum_r v[r] * M[r*cols + c]   (thread per column; coalesced across columns)
template <class R>
struct FrRowMulBody {
  const uint32_t *v; const uint32_t *m; size_t rows, cols; uint32_t *out;
  PCGPU_KERNEL_DEV void operator()(size_t c) const {
    Fp<R> acc = Fp<R>::zero();
    for (size_t r = 0; r < rows; r++) acc = fp_add<R>(acc, fp_mul<R>(load_fr<R>(v, r), load_fr<R>(m, r * cols + c)));
    store_fr<R>(out, c, acc);
  }
};

}  // namespace pcgpu
