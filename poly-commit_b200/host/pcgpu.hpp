// pcgpu.hpp -- C++17 host-side mirror of the reference's operator interface for the hot path, over the C ABI
// (include/pcgpu.h).  Header-only; link with -lpcgpu.
//
// The reference is Rust and cannot be compiled in this image, so this header plays the role its patched call sites
// would: same names, argument meaning and error behaviour as
//   kzg10::Powers / Commitment / Randomness / Proof         poly-commit/src/kzg10/data_structures.rs:124-129, :325-328, :400-404, :489-495
//   kzg10::KZG10::commit / open                              poly-commit/src/kzg10/mod.rs:157-210, :287-310
//   Error::TooManyCoefficients / HidingBoundToolarge         poly-commit/src/error.rs:36, kzg10/mod.rs:392-422
//   VariableBaseMSM::msm_bigint                              (ark-ec; call sites kzg10/mod.rs:175-178 ...)
// Polynomials are dense coefficient vectors of Montgomery-form Fr (what DensePolynomial<Fr>::coeffs holds).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pcgpu.h"

namespace pcgpu {

enum class Curve : int { Bls12_381 = PCGPU_BLS12_381, Bn254 = PCGPU_BN254, Pallas = PCGPU_PALLAS };
constexpr size_t fq_limbs(Curve c) { return c == Curve::Bls12_381 ? 6 : 4; }

// mirrors poly-commit/src/error.rs for the variants this path can raise
struct Error : std::runtime_error {
  enum Kind { TooManyCoefficients, HidingBoundToolarge, LengthMismatch, ScalarOutOfRange, Device } kind;
  int code;
  Error(Kind k, int c, const std::string &m) : std::runtime_error(m), kind(k), code(c) {}
  static void check(int rc) {
    if (rc == PCGPU_OK) return;
    Kind k = rc == PCGPU_E_DEGREE ? TooManyCoefficients : rc == PCGPU_E_HIDING ? HidingBoundToolarge
           : rc == PCGPU_E_LEN ? LengthMismatch : rc == PCGPU_E_RANGE ? ScalarOutOfRange : Device;
    throw Error(k, rc, pcgpu_strerror(rc));
  }
};

using Fr = std::array<uint64_t, 4>;  // Montgomery form unless a function says "bigint"

struct G1Affine {            // x || y Montgomery limbs + infinity flag (ark-ec Affine { x, y, infinity })
  std::array<uint64_t, 12> xy{};
  bool infinity = false;
};

class Context {              // one CUDA device + stream; not copyable
 public:
  explicit Context(int device = 0) { Error::check(pcgpu_init(device, &ctx_)); }
  ~Context() { pcgpu_destroy(ctx_); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  pcgpu_ctx *raw() const { return ctx_; }
 private:
  pcgpu_ctx *ctx_ = nullptr;
};

class Bases {                // device-resident affine bases (a Powers vector, an IPA comm_key, a Hyrax com_key)
 public:
  Bases(Context &ctx, Curve curve, const uint64_t *xy, const uint8_t *inf, size_t n, uint32_t flags = PCGPU_SRS_PRECOMPUTE)
      : ctx_(ctx), curve_(curve), n_(n) { Error::check(pcgpu_srs_register(ctx.raw(), (int)curve, xy, inf, n, flags, &srs_)); }
  ~Bases() { pcgpu_srs_release(ctx_.raw(), srs_); }
  Bases(const Bases &) = delete;
  Bases &operator=(const Bases &) = delete;
  size_t len() const { return n_; }
  Curve curve() const { return curve_; }
  const pcgpu_srs *raw() const { return srs_; }
 private:
  Context &ctx_; Curve curve_; size_t n_; pcgpu_srs *srs_ = nullptr;
};

// <G as VariableBaseMSM>::msm_bigint(&bases[base_offset..], bigints)
inline G1Affine msm_bigint(Context &ctx, const Bases &bases, size_t base_offset, const std::vector<Fr> &bigints) {
  G1Affine out; uint8_t inf = 0;
  Error::check(pcgpu_msm(ctx.raw(), bases.raw(), base_offset, bigints.data(), bigints.size(), 0, out.xy.data(), &inf));
  out.infinity = inf != 0;
  return out;
}

namespace kzg10 {

struct Powers {              // kzg10/data_structures.rs:124-129
  const Bases &powers_of_g;
  const Bases *powers_of_gamma_g;   // may be null when nothing is hiding
  size_t size() const { return powers_of_g.len(); }
};
struct Commitment { G1Affine point; };                       // :325-328
struct Randomness {                                          // :400-404
  std::vector<Fr> blinding_polynomial;
  bool is_hiding() const { return !blinding_polynomial.empty(); }
  static Randomness empty() { return {}; }
};
struct Proof { G1Affine w; std::optional<Fr> random_v; };     // :489-495

struct KZG10 {
  // KZG10::commit(powers, polynomial, hiding_bound, rng)  kzg10/mod.rs:157-210.  The reference samples the blinding
  // polynomial from its RNG (:182-195); here the caller passes it (empty = not hiding) so results are reproducible.
  static std::pair<Commitment, Randomness> commit(Context &ctx, const Powers &powers, const std::vector<Fr> &polynomial,
                                                  const std::vector<Fr> &blinding_polynomial = {}) {
    Commitment c; uint8_t inf = 0;
    Error::check(pcgpu_kzg_commit(ctx.raw(), powers.powers_of_g.raw(), polynomial.data(), polynomial.size(),
                                  powers.powers_of_gamma_g ? powers.powers_of_gamma_g->raw() : nullptr,
                                  blinding_polynomial.data(), blinding_polynomial.size(), 0, c.point.xy.data(), &inf));
    c.point.infinity = inf != 0;
    return {c, Randomness{blinding_polynomial}};
  }
  // KZG10::open(powers, p, point, rand)  kzg10/mod.rs:287-310
  static Proof open(Context &ctx, const Powers &powers, const std::vector<Fr> &p, const Fr &point, const Randomness &rand) {
    Proof pr; uint8_t inf = 0; Fr rv{};
    Error::check(pcgpu_kzg_open(ctx.raw(), powers.powers_of_g.raw(), p.data(), p.size(), point.data(),
                                powers.powers_of_gamma_g ? powers.powers_of_gamma_g->raw() : nullptr,
                                rand.blinding_polynomial.data(), rand.blinding_polynomial.size(), 0, pr.w.xy.data(), &inf,
                                rv.data()));
    pr.w.infinity = inf != 0;
    if (rand.is_hiding()) pr.random_v = rv;
    return pr;
  }
};

}  // namespace kzg10
}  // namespace pcgpu
