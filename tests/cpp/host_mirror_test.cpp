// Drives the C++ host mirror (poly-commit_b200/host/pcgpu.hpp) the way the reference's kzg10 tests drive KZG10
// (kzg10/mod.rs:546-575 end_to_end_test_template): commit, open, and the degree-too-large error -- on inputs
// read from a file written by tests/test_gpu_parity.py, which compares the output with the oracle.
//   in : u32 curve, u32 n_powers, u32 n_coeffs, u32 n_gamma, u32 n_blind, then powers xy, coeffs, z, gamma xy, blind (u64 LE)
//   out: commitment xy, witness xy, hiding commitment xy, hiding witness xy, random_v, u32 error kind seen for an oversized poly
#include <cstdio>
#include <vector>

#include "../../poly-commit_b200/host/pcgpu.hpp"

using namespace pcgpu;

static std::vector<uint64_t> rd(FILE *f, size_t n) { std::vector<uint64_t> v(n); if (n && fread(v.data(), 8, n, f) != n) throw std::runtime_error("short read"); return v; }

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  uint32_t hdr[5];
  if (fread(hdr, 4, 5, f) != 5) return 2;
  Curve curve = (Curve)hdr[0];
  size_t nq = fq_limbs(curve), np = hdr[1], nc = hdr[2], ng = hdr[3], nb = hdr[4];
  auto powers = rd(f, np * 2 * nq), coeffs = rd(f, nc * 4), z = rd(f, 4), gamma = rd(f, ng * 2 * nq), blind = rd(f, nb * 4);
  fclose(f);
  try {
    Context ctx(0);
    Bases pg(ctx, curve, powers.data(), nullptr, np), gg(ctx, curve, gamma.data(), nullptr, ng, 0);
    kzg10::Powers pw{pg, &gg};
    std::vector<Fr> poly(nc), bl(nb);
    memcpy(poly.data(), coeffs.data(), nc * 32); memcpy(bl.data(), blind.data(), nb * 32);
    Fr point; memcpy(point.data(), z.data(), 32);
    auto [comm, rand0] = kzg10::KZG10::commit(ctx, pw, poly);
    auto proof = kzg10::KZG10::open(ctx, pw, poly, point, rand0);
    auto [hcomm, hrand] = kzg10::KZG10::commit(ctx, pw, poly, bl);
    auto hproof = kzg10::KZG10::open(ctx, pw, poly, point, hrand);
    uint32_t kind = 99;
    try {
      std::vector<Fr> big(np + 1, poly[0]); big.back()[0] |= 1;   // non-zero leading coefficient: degree = np
      kzg10::KZG10::commit(ctx, pw, big);
    } catch (const Error &e) { kind = (uint32_t)e.kind; }
    FILE *o = fopen(argv[2], "wb");
    fwrite(comm.point.xy.data(), 8, 2 * nq, o); fwrite(proof.w.xy.data(), 8, 2 * nq, o);
    fwrite(hcomm.point.xy.data(), 8, 2 * nq, o); fwrite(hproof.w.xy.data(), 8, 2 * nq, o);
    fwrite(hproof.random_v->data(), 8, 4, o); fwrite(&kind, 4, 1, o);
    fclose(o);
  } catch (const std::exception &e) { fprintf(stderr, "host_mirror_test: %s\n", e.what()); return 1; }
  return 0;
}
