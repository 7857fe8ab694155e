// Column hashing and the Merkle tree of the linear-code commitments (Ligero / Brakedown) -- SURVEY.md section 8f rank 4:
//   leaves[j] = H::evaluate(col_hash_params, ext_mat.col(j))            linear_codes/mod.rs:255-262
//   col_tree  = create_merkle_tree(leaves padded to a power of two)      linear_codes/mod.rs:268-272, :507-523
// for the hashers the reference's own tests and benches instantiate (linear_codes/*/tests.rs, bench-templates/src/lib.rs):
//   H            = FieldToBytesColHasher<F, Blake2s256> (utils.rs:208-236): D(to_bytes!(column)), i.e. the digest of
//                  u64 little-endian length || n canonical 32-byte little-endian field elements        [SHA-256 selectable]
//   LeafHash     = LeafIdentityHasher (utils.rs:187-205): the leaf digest is the 32 column-hash bytes themselves
//   TwoToOneHash = ark-crypto-primitives crh::sha256::Sha256: SHA-256(left || right); at the leaf level each side first goes
//                  through ByteDigestConverter = to_uncompressed_bytes!(Vec<u8>) = u64 length || bytes, so a pair of 32-byte
//                  leaves hashes 80 bytes and a padding leaf (Vec::default(), linear_codes/mod.rs:519) contributes 8 zero bytes
// (ark-crypto-primitives 0.5.0 is un-vendored: restated from its published behaviour; digests are checked against hashlib.)
//
// One thread per column: the rows of `ext_mat` are contiguous, so the 32 threads of a warp read 32 adjacent elements of a row
// (1 KB, fully coalesced) -- the encoded matrix is consumed in place, straight out of pcgpu_ntt_batch, with F::into_bigint
// fused into the load.  HBM-bound: 32 bytes per element read once, 32 bytes per column written.
#pragma once
#include "frops.cuh"
#include "rt.cuh"

namespace pcgpu {

enum { HASH_BLAKE2S = 0, HASH_SHA256 = 1 };

PCGPU_DEV uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
PCGPU_DEV uint32_t bswap32(uint32_t x) { return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24); }

// ---- BLAKE2s-256 (RFC 7693), unkeyed ----------------------------------------------------------------------------------------
struct Blake2s {
  uint32_t h[8], m[16];
  uint32_t fill;        // words in m
  uint64_t t;           // bytes compressed so far
  PCGPU_DEV static uint32_t iv(int i) {
    const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    return IV[i];
  }
  PCGPU_DEV void init() {
    for (int i = 0; i < 8; i++) h[i] = iv(i);
    h[0] ^= 0x01010020u;   // digest length 32, no key, fanout 1, depth 1
    fill = 0; t = 0;
  }
  PCGPU_DEV void compress(uint32_t nbytes, bool last) {
    // message schedule, 4 bits per index
    const uint64_t SIGMA[10] = {0xfedcba9876543210ull, 0x357b20c16df984aeull, 0x491763eadf250c8bull, 0x8f04a562ebcd1397ull,
                                0xd386cb1efa427509ull, 0x91ef57d438b0a6c2ull, 0xb8293670a4def15cull, 0xa2684f05931ce7bdull,
                                0x5a417d2c803b9ef6ull, 0x0dc3e9bf5167482aull};
    t += nbytes;
    uint32_t v[16];
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[8 + i] = iv(i); }
    v[12] ^= (uint32_t)t; v[13] ^= (uint32_t)(t >> 32);
    if (last) v[14] = ~v[14];
#define PCGPU_B2S_G(a, b, c, d, x, y)                                         \
    v[a] = v[a] + v[b] + (x); v[d] = rotr32(v[d] ^ v[a], 16);                 \
    v[c] = v[c] + v[d];       v[b] = rotr32(v[b] ^ v[c], 12);                 \
    v[a] = v[a] + v[b] + (y); v[d] = rotr32(v[d] ^ v[a], 8);                  \
    v[c] = v[c] + v[d];       v[b] = rotr32(v[b] ^ v[c], 7);
#pragma unroll
    for (int r = 0; r < 10; r++) {
      const uint64_t s = SIGMA[r];
#define PCGPU_B2S_M(k) m[(s >> (4 * (k))) & 15]
      PCGPU_B2S_G(0, 4, 8, 12, PCGPU_B2S_M(0), PCGPU_B2S_M(1))
      PCGPU_B2S_G(1, 5, 9, 13, PCGPU_B2S_M(2), PCGPU_B2S_M(3))
      PCGPU_B2S_G(2, 6, 10, 14, PCGPU_B2S_M(4), PCGPU_B2S_M(5))
      PCGPU_B2S_G(3, 7, 11, 15, PCGPU_B2S_M(6), PCGPU_B2S_M(7))
      PCGPU_B2S_G(0, 5, 10, 15, PCGPU_B2S_M(8), PCGPU_B2S_M(9))
      PCGPU_B2S_G(1, 6, 11, 12, PCGPU_B2S_M(10), PCGPU_B2S_M(11))
      PCGPU_B2S_G(2, 7, 8, 13, PCGPU_B2S_M(12), PCGPU_B2S_M(13))
      PCGPU_B2S_G(3, 4, 9, 14, PCGPU_B2S_M(14), PCGPU_B2S_M(15))
#undef PCGPU_B2S_M
    }
#undef PCGPU_B2S_G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[8 + i];
  }
  // little-endian 32-bit words of the message, in order.  A full buffer is compressed only when more data arrives, so the
  // last block (full or partial) is always the one flagged final.
  PCGPU_DEV void push(uint32_t w) {
    if (fill == 16) { compress(64, false); fill = 0; }
    m[fill++] = w;
  }
  PCGPU_DEV void finish(uint32_t *out8) {
    const uint32_t nbytes = 4 * fill;
    for (uint32_t i = fill; i < 16; i++) m[i] = 0;
    compress(nbytes, true);
    for (int i = 0; i < 8; i++) out8[i] = h[i];   // digest bytes = the words little-endian
  }
  // static-index interface (hot loops): word slot `idx` of the current block, a full non-final block, the final block
  PCGPU_DEV void put(int idx, uint32_t le_word) { m[idx] = le_word; }
  PCGPU_DEV void full_block() { compress(64, false); }
  PCGPU_DEV void final_block(int words, uint32_t *out8) { fill = (uint32_t)words; finish(out8); }
};

// ---- SHA-256 (FIPS 180-4) ---------------------------------------------------------------------------------------------------
struct Sha256 {
  uint32_t h[8], w[16];
  uint32_t fill;        // words in w (message words are pushed little-endian and swapped here)
  uint64_t nbytes;
  PCGPU_DEV static uint32_t k(int i) {
    const uint32_t K[64] = {
        0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u,
        0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu,
        0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u,
        0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
        0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u,
        0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u,
        0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
    return K[i];
  }
  PCGPU_DEV void init() {
    const uint32_t H0[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    for (int i = 0; i < 8; i++) h[i] = H0[i];
    fill = 0; nbytes = 0;
  }
  PCGPU_DEV void compress() {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    uint32_t ws[16];
    for (int i = 0; i < 16; i++) ws[i] = w[i];
#pragma unroll
    for (int i = 0; i < 64; i++) {
      if (i >= 16) {
        uint32_t w15 = ws[(i + 1) & 15], w2 = ws[(i + 14) & 15];
        uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3), s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
        ws[i & 15] = ws[i & 15] + s0 + ws[(i + 9) & 15] + s1;
      }
      uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25), ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + k(i) + ws[i & 15];
      uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    fill = 0;
  }
  PCGPU_DEV void push(uint32_t le_word) {   // 4 message bytes given as a little-endian word
    w[fill++] = bswap32(le_word);
    nbytes += 4;
    if (fill == 16) compress();
  }
  PCGPU_DEV void finish(uint32_t *out8) {   // out words are little-endian packed digest bytes (memcpy-able)
    const uint64_t bits = nbytes * 8;
    w[fill++] = 0x80000000u;
    if (fill == 16) compress();
    if (fill > 14) { while (fill < 16) w[fill++] = 0; compress(); }
    while (fill < 14) w[fill++] = 0;
    w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits; fill = 16;
    compress();
    for (int i = 0; i < 8; i++) out8[i] = bswap32(h[i]);
  }
  PCGPU_DEV void put(int idx, uint32_t le_word) { w[idx] = bswap32(le_word); }
  PCGPU_DEV void full_block() { nbytes += 64; compress(); }
  PCGPU_DEV void final_block(int words, uint32_t *out8) { fill = (uint32_t)words; nbytes += 4 * (uint64_t)words; finish(out8); }
};

// leaves[j] = D(u64 n_rows || canonical(ext_mat[0][j]) || ... || canonical(ext_mat[n_rows-1][j]))
template <class R, class D>
struct ColumnHashBody {
  const uint32_t *mat; uint64_t n_rows, n_cols; uint32_t *leaves; int mont;
  PCGPU_KERNEL_DEV void operator()(size_t j) const {
    D d;
    d.init();
    // the stream is 2 prefix words then 8 words per element, so every 64-byte block is (2 carried words | element | 6 words of
    // the next element): two rows per iteration keep every buffer index a compile-time constant (registers, no local memory)
    d.put(0, (uint32_t)n_rows); d.put(1, (uint32_t)(n_rows >> 32));
    uint64_t r = 0;
    for (; r + 2 <= n_rows; r += 2) {
      Fp<R> a = load_fr<R>(mat, r * n_cols + j), b = load_fr<R>(mat, (r + 1) * n_cols + j);
      if (mont) { a = fp_from_mont<R>(a); b = fp_from_mont<R>(b); }
#pragma unroll
      for (int l = 0; l < 8; l++) d.put(2 + l, a.l[l]);
#pragma unroll
      for (int l = 0; l < 6; l++) d.put(10 + l, b.l[l]);
      d.full_block();
      d.put(0, b.l[6]); d.put(1, b.l[7]);
    }
    uint32_t out[8];
    if (r < n_rows) {
      Fp<R> a = load_fr<R>(mat, r * n_cols + j);
      if (mont) a = fp_from_mont<R>(a);
#pragma unroll
      for (int l = 0; l < 8; l++) d.put(2 + l, a.l[l]);
      d.final_block(10, out);
    } else {
      d.final_block(2, out);
    }
    for (int l = 0; l < 8; l++) leaves[j * 8 + l] = out[l];
  }
};

// Merkle levels.  Nodes are stored in heap order (root = node 0, children of i are 2i+1 and 2i+2) like ark-crypto-primitives'
// MerkleTree::non_leaf_nodes; `P` leaves (a power of two >= 2) give P - 1 inner nodes.
//   leaf level: node (P/2 - 1 + i) = SHA-256(len(L) || L || len(R) || R) for the leaf pair (2i, 2i+1); leaves >= n_leaves are
//               the empty padding leaf (8 zero length bytes, no payload)
struct MerkleLeafLevelBody {
  const uint32_t *leaves; uint64_t n_leaves, P; uint32_t *nodes;
  PCGPU_KERNEL_DEV void operator()(size_t i) const {
    Sha256 s;
    s.init();
    for (int side = 0; side < 2; side++) {
      const uint64_t leaf = 2 * i + side;
      const bool real = leaf < n_leaves;
      s.push(real ? 32u : 0u); s.push(0u);
      if (real) for (int l = 0; l < 8; l++) s.push(leaves[leaf * 8 + l]);
    }
    uint32_t out[8];
    s.finish(out);
    const uint64_t node = P / 2 - 1 + i;
    for (int l = 0; l < 8; l++) nodes[node * 8 + l] = out[l];
  }
};
//   inner level of `count` nodes starting at heap index `first`: node = SHA-256(left child digest || right child digest)
struct MerkleInnerLevelBody {
  uint32_t *nodes; uint64_t first;
  PCGPU_KERNEL_DEV void operator()(size_t i) const {
    const uint64_t node = first + i;
    Sha256 s;
    s.init();
    for (int l = 0; l < 8; l++) s.push(nodes[(2 * node + 1) * 8 + l]);
    for (int l = 0; l < 8; l++) s.push(nodes[(2 * node + 2) * 8 + l]);
    uint32_t out[8];
    s.finish(out);
    for (int l = 0; l < 8; l++) nodes[node * 8 + l] = out[l];
  }
};

// nodes: (P - 1) * 8 words on the device
inline int merkle_build(const uint32_t *d_leaves, uint64_t n_leaves, uint64_t P, uint32_t *d_nodes, rt::stream_t st) {
  int rc = rt::launch<128>(MerkleLeafLevelBody{d_leaves, n_leaves, P, d_nodes}, P / 2, st);
  if (rc) return rc;
  for (uint64_t count = P / 4; count >= 1; count /= 2) {
    if ((rc = rt::launch<128>(MerkleInnerLevelBody{d_nodes, count - 1}, count, st))) return rc;
    if (count == 1) break;
  }
  return rt::OK;
}

}  // namespace pcgpu
