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

// a0 b0 + a1 b1: one Montgomery reduction for the two products when the modulus leaves the headroom (BN254, Pallas: 3r < 2^256;
// BLS12-381's r does not), two ordinary products otherwise
template <class R>
PCGPU_DEV Fp<R> fr_dot2(const Fp<R> &a0, const Fp<R> &b0, const Fp<R> &a1, const Fp<R> &b1) {
  if constexpr (mont_mul2_supported<R>()) return fp_mul2<R>(a0, b0, a1, b1);
  else return fp_add<R>(fp_mul<R>(a0, b0), fp_mul<R>(a1, b1));
}

template <class R>
struct FrFromMontBody {
  const uint32_t *in; uint32_t *out;
  PCGPU_KERNEL_DEV void operator()(size_t i) const { store_fr<R>(out, i, fp_from_mont<R>(load_fr<R>(in, i))); }
};
// out[i] = a[i] * b[i]
template <class R>
struct FrMulBody {
  const uint32_t *a, *b; uint32_t *out;
  PCGPU_KERNEL_DEV void operator()(size_t i) const { store_fr<R>(out, i, fp_mul<R>(load_fr<R>(a, i), load_fr<R>(b, i))); }
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
// Division by (X - z) as a multi-level scan of the affine maps  t -> p_i + z * t.
//   level 0: chunks of DIV_K coefficients (thread per chunk) -> local Horner value
//   level l: nodes of DIV_F children each, folded with the factor z^(span of a child)
//   the top level (<= DIV_F nodes) is walked by one thread; carries are then pushed back down level by
//   level and every chunk replays its recurrence writing q.
// q[i-1] = p[i] + z*q[i]; carry into a chunk = q[hi-1] with q[n-1] := 0; remainder = p(z).
// ---------------------------------------------------------------------------------------------
enum { DIV_K = 32, DIV_F = 32, DIV_MAX_LEVELS = 8 };

template <class R>
struct DivPowersBody {  // zp[0] = z^DIV_K, zp[l] = zp[l-1]^DIV_F
  const uint32_t *z; uint32_t *zp; uint32_t levels;
  PCGPU_KERNEL_DEV void operator()(size_t) const {
    Fp<R> a = load_fr<R>(z, 0);
    for (int k = DIV_K; k > 1; k >>= 1) a = fp_sqr<R>(a);
    store_fr<R>(zp, 0, a);
    for (uint32_t l = 1; l < levels; l++) {
      for (int k = DIV_F; k > 1; k >>= 1) a = fp_sqr<R>(a);
      store_fr<R>(zp, l, a);
    }
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

// parent local value: fold the children's locals from the top child down (factor = z^(span of one child))
template <class R>
struct DivFoldBody {
  const uint32_t *child; size_t nchild; const uint32_t *factor; uint32_t *parent;
  PCGPU_KERNEL_DEV void operator()(size_t g) const {
    size_t lo = g * DIV_F, hi = lo + DIV_F < nchild ? lo + DIV_F : nchild;
    Fp<R> f = load_fr<R>(factor, 0), acc = Fp<R>::zero();
    for (size_t c = hi; c-- > lo;) acc = fp_add<R>(fp_mul<R>(acc, f), load_fr<R>(child, c));
    store_fr<R>(parent, g, acc);
  }
};

// carries into the children of node g (carry[c] = value entering child c from above); with parent_carry == null the
// node is the (single) root whose incoming carry is zero and whose outgoing value is the remainder
template <class R>
struct DivCarryBody {
  const uint32_t *child; size_t nchild; const uint32_t *factor; const uint32_t *parent_carry; uint32_t *carry; uint32_t *rem;
  PCGPU_KERNEL_DEV void operator()(size_t g) const {
    size_t lo = g * DIV_F, hi = lo + DIV_F < nchild ? lo + DIV_F : nchild;
    if (!parent_carry) { lo = 0; hi = nchild; }
    Fp<R> f = load_fr<R>(factor, 0), t = parent_carry ? load_fr<R>(parent_carry, g) : Fp<R>::zero();
    for (size_t c = hi; c-- > lo;) {
      store_fr<R>(carry, c, t);
      t = fp_add<R>(fp_mul<R>(t, f), load_fr<R>(child, c));
    }
    if (rem) store_fr<R>(rem, 0, t);
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

// Chunk carries in ONE launch (replaces the fold / carry levels above for up to DIV_SCAN_BLOCK * 4096 chunks): the chunk values
// obey  out_c = local_c + z^K * out_{c+1}  (out beyond the top chunk = 0), a reverse linear recurrence.  One block: thread t walks
// its segment of S consecutive chunks (Horner, top down), the DIV_SCAN_BLOCK segment values are combined by a reverse
// Hillis-Steele scan in shared memory (E_t += W^(2^k) E_(t + 2^k), W = z^(K S)), and every thread replays its segment writing
// the value that ENTERS each chunk (`carry`), which DivChunkWriteBody consumes.  The value leaving chunk 0 is the remainder p(z).
enum { DIV_SCAN_BLOCK = 1024 };
template <class R>
struct DivBlockScanBody {
  const uint32_t *local; size_t nchunks; const uint32_t *z; uint32_t *carry; uint32_t *rem;
  PCGPU_KERNEL_DEV void operator()(size_t, uint32_t *smem) const {
    const size_t S = (nchunks + DIV_SCAN_BLOCK - 1) / DIV_SCAN_BLOCK;     // chunks per thread
    uint32_t *bufA = smem, *bufB = smem + 8 * DIV_SCAN_BLOCK, *pw = smem + 16 * DIV_SCAN_BLOCK;   // pw[0] = z^K, pw[1] = W^(2^k)
    PCGPU_BLOCK_FOR(t, 1) {
      Fp<R> a = load_fr<R>(z, 0);
      for (int k = DIV_K; k > 1; k >>= 1) a = fp_sqr<R>(a);
      store_fr<R>(pw, 0, a);                                              // z^K
      Fp<R> w = Fp<R>::one(), b = a;                                      // W = (z^K)^S by square-and-multiply
      for (size_t e = S; e; e >>= 1) { if (e & 1) w = fp_mul<R>(w, b); b = fp_sqr<R>(b); }
      store_fr<R>(pw, 1, w);
    }
    PCGPU_BLOCK_SYNC();
    PCGPU_BLOCK_FOR(t, DIV_SCAN_BLOCK) {
      const size_t lo = (size_t)t * S, hi = lo + S < nchunks ? lo + S : nchunks;
      const Fp<R> zk = load_fr<R>(pw, 0);
      Fp<R> acc = Fp<R>::zero();
      for (size_t c = hi; c-- > lo;) acc = fp_add<R>(fp_mul<R>(acc, zk), load_fr<R>(local, c));
      store_fr<R>(bufA, t, acc);
    }
    PCGPU_BLOCK_SYNC();
    uint32_t *x = bufA, *y = bufB;
    for (uint32_t d = 1; d < DIV_SCAN_BLOCK; d <<= 1) {
      PCGPU_BLOCK_FOR(t, DIV_SCAN_BLOCK) {
        Fp<R> v = load_fr<R>(x, t);
        if (t + d < DIV_SCAN_BLOCK) v = fp_add<R>(v, fp_mul<R>(load_fr<R>(pw, 1), load_fr<R>(x, t + d)));
        store_fr<R>(y, t, v);
      }
      PCGPU_BLOCK_SYNC();
      PCGPU_BLOCK_FOR(t, 1) { store_fr<R>(pw, 1, fp_sqr<R>(load_fr<R>(pw, 1))); }
      PCGPU_BLOCK_SYNC();
      uint32_t *tmp = x; x = y; y = tmp;
    }
    // x[t] = value leaving segment t's lowest chunk; the value entering segment t from above is x[t + 1]
    PCGPU_BLOCK_FOR(t, DIV_SCAN_BLOCK) {
      const size_t lo = (size_t)t * S, hi = lo + S < nchunks ? lo + S : nchunks;
      const Fp<R> zk = load_fr<R>(pw, 0);
      Fp<R> in = t + 1 < DIV_SCAN_BLOCK ? load_fr<R>(x, t + 1) : Fp<R>::zero();
      for (size_t c = hi; c-- > lo;) {
        store_fr<R>(carry, c, in);
        in = fp_add<R>(fp_mul<R>(in, zk), load_fr<R>(local, c));
      }
      if (t == 0 && rem) store_fr<R>(rem, 0, in);
    }
  }
};

inline size_t div_scratch_words(size_t n) {
  size_t cnt = (n + DIV_K - 1) / DIV_K, tot = 0;
  for (int l = 0; l < DIV_MAX_LEVELS; l++) { tot += 2 * cnt; if (cnt <= DIV_F) break; cnt = (cnt + DIV_F - 1) / DIV_F; }
  return 8 * (tot + DIV_MAX_LEVELS + 4);
}

// p: n coefficients, q: n-1 coefficients (n >= 1), rem: 1 element, z: 1 element; all device.
template <class R>
inline int fr_div_linear(const uint32_t *p, size_t n, const uint32_t *z, uint32_t *q, uint32_t *rem,
                         uint32_t *scratch, rt::stream_t st) {
  if (n == 0) return rt::dev_memset(rem, 0, 32, st);
  size_t cnt[DIV_MAX_LEVELS]; uint32_t *local[DIV_MAX_LEVELS], *carry[DIV_MAX_LEVELS];
  uint32_t *zp = scratch, *cur = scratch + 8 * DIV_MAX_LEVELS;
  int levels = 0;
  for (size_t c = (n + DIV_K - 1) / DIV_K;; c = (c + DIV_F - 1) / DIV_F) {
    cnt[levels] = c; local[levels] = cur; cur += 8 * c; carry[levels] = cur; cur += 8 * c;
    levels++;
    if (c <= DIV_F || levels == DIV_MAX_LEVELS) break;
  }
  int rc;
  // (a three-launch variant with ONE block scanning all chunk carries was measured slower: 0.99 ms vs 0.36 ms at 2^22 -- 128
  // dependent products per thread twice over; the level tree below keeps every chain at DIV_F = 32)
  if (getenv("PCGPU_DIV_BLOCK_SCAN") && cnt[0] <= (size_t)DIV_SCAN_BLOCK * 4096) {
    if ((rc = rt::launch<128>(DivChunkLocalBody<R>{p, n, z, local[0]}, cnt[0], st))) return rc;
    if ((rc = rt::launch_blocks<DIV_SCAN_BLOCK>(DivBlockScanBody<R>{local[0], cnt[0], z, carry[0], rem}, 1, (16 * DIV_SCAN_BLOCK + 32) * 4, st))) return rc;
    return rt::launch<128>(DivChunkWriteBody<R>{p, n, z, carry[0], q}, cnt[0], st);
  }
  if ((rc = rt::launch<32>(DivPowersBody<R>{z, zp, (uint32_t)levels}, 1, st))) return rc;
  if ((rc = rt::launch<128>(DivChunkLocalBody<R>{p, n, z, local[0]}, cnt[0], st))) return rc;
  for (int l = 1; l < levels; l++)
    if ((rc = rt::launch<64>(DivFoldBody<R>{local[l - 1], cnt[l - 1], zp + 8 * (l - 1), local[l]}, cnt[l], st))) return rc;
  // root: one thread walks the top level (<= DIV_F nodes unless DIV_MAX_LEVELS was hit)
  if ((rc = rt::launch<32>(DivCarryBody<R>{local[levels - 1], cnt[levels - 1], zp + 8 * (levels - 1), nullptr, carry[levels - 1], rem}, 1, st))) return rc;
  for (int l = levels - 1; l >= 1; l--)
    if ((rc = rt::launch<64>(DivCarryBody<R>{local[l - 1], cnt[l - 1], zp + 8 * (l - 1), carry[l], carry[l - 1], nullptr}, cnt[l], st))) return rc;
  return rt::launch<128>(DivChunkWriteBody<R>{p, n, z, carry[0], q}, cnt[0], st);
}

// ---------------------------------------------------------------------------------------------
// inner product: strided per-thread partial sums (coalesced), then block-level shared-memory trees
// ---------------------------------------------------------------------------------------------
enum { IP_THREADS = 65536, IP_BLOCK = 256 };
template <class R>
struct IpPartialBody {
  const uint32_t *a; const uint32_t *b; size_t n; uint32_t *partial;
  PCGPU_KERNEL_DEV void operator()(size_t t) const {
    // two elements per iteration: four independent loads in flight and ONE Montgomery reduction for the pair
    // (fr_dot2 = a0 b0 + a1 b1 with a single reduction where 3r < 2^256: 192 instead of 256 wide multiplies per pair)
    Fp<R> acc = Fp<R>::zero();
    size_t i = t;
    for (; i + IP_THREADS < n; i += 2 * (size_t)IP_THREADS) {
      const Fp<R> a0 = load_fr<R>(a, i), b0 = load_fr<R>(b, i), a1 = load_fr<R>(a, i + IP_THREADS), b1 = load_fr<R>(b, i + IP_THREADS);
      acc = fp_add<R>(acc, fr_dot2<R>(a0, b0, a1, b1));
    }
    if (i < n) acc = fp_add<R>(acc, fp_mul<R>(load_fr<R>(a, i), load_fr<R>(b, i)));
    store_fr<R>(partial, t, acc);
  }
};
// block b sums in[b*IP_BLOCK .. +IP_BLOCK) (count m) into out[b]
template <class R>
struct FrBlockSumBody {
  const uint32_t *in; size_t m; uint32_t *out;
  PCGPU_KERNEL_DEV void operator()(size_t b, uint32_t *smem) const {
    PCGPU_BLOCK_FOR(i, IP_BLOCK) {
      size_t idx = b * IP_BLOCK + i;
      Fp<R> v = idx < m ? load_fr<R>(in, idx) : Fp<R>::zero();
#pragma unroll
      for (int l = 0; l < 8; l++) smem[l * IP_BLOCK + i] = v.l[l];
    }
    PCGPU_BLOCK_SYNC();
    for (uint32_t half = IP_BLOCK / 2; half >= 1; half >>= 1) {
      PCGPU_BLOCK_FOR(i, half) {
        Fp<R> x, y;
#pragma unroll
        for (int l = 0; l < 8; l++) { x.l[l] = smem[l * IP_BLOCK + i]; y.l[l] = smem[l * IP_BLOCK + i + half]; }
        x = fp_add<R>(x, y);
#pragma unroll
        for (int l = 0; l < 8; l++) smem[l * IP_BLOCK + i] = x.l[l];
      }
      PCGPU_BLOCK_SYNC();
    }
    PCGPU_BLOCK_FOR(i, 1) {
      Fp<R> v;
#pragma unroll
      for (int l = 0; l < 8; l++) v.l[l] = smem[l * IP_BLOCK];
      store_fr<R>(out, b, v);
    }
  }
};
// scratch: (IP_THREADS + IP_THREADS / IP_BLOCK) elements
template <class R>
inline int fr_inner_product(const uint32_t *a, const uint32_t *b, size_t n, uint32_t *out, uint32_t *scratch, rt::stream_t st) {
  int rc;
  uint32_t *lvl1 = scratch + 8 * IP_THREADS;
  if ((rc = rt::launch<128>(IpPartialBody<R>{a, b, n, scratch}, IP_THREADS, st))) return rc;
  if ((rc = rt::launch_blocks<IP_BLOCK>(FrBlockSumBody<R>{scratch, IP_THREADS, lvl1}, IP_THREADS / IP_BLOCK, IP_BLOCK * 32, st))) return rc;
  return rt::launch_blocks<IP_BLOCK>(FrBlockSumBody<R>{lvl1, IP_THREADS / IP_BLOCK, out}, 1, IP_BLOCK * 32, st);
}

// out[c] = sum_r v[r] * M[r*cols + c]   (thread per column; coalesced across columns)
template <class R>
struct FrRowMulBody {
  const uint32_t *v; const uint32_t *m; size_t rows, cols; uint32_t *out;
  PCGPU_KERNEL_DEV void operator()(size_t c) const {
    Fp<R> acc = Fp<R>::zero();
    size_t r = 0;
    for (; r + 2 <= rows; r += 2)      // pairs of rows: one reduction per two products (fr_dot2)
      acc = fp_add<R>(acc, fr_dot2<R>(load_fr<R>(v, r), load_fr<R>(m, r * cols + c), load_fr<R>(v, r + 1), load_fr<R>(m, (r + 1) * cols + c)));
    if (r < rows) acc = fp_add<R>(acc, fp_mul<R>(load_fr<R>(v, r), load_fr<R>(m, r * cols + c)));
    store_fr<R>(out, c, acc);
  }
};

}  // namespace pcgpu
