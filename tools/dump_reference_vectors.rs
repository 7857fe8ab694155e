//! Dumps (bases, coefficients, z, commitment, witness) of one KZG10 commit + open computed by the REFERENCE
//! (arkworks-rs/poly-commit @ a05ec99, `kzg10::KZG10::{setup, commit, open}`, kzg10/mod.rs:47-310) into the flat
//! binary file `tests/test_external_kats.py::_read_dump` replays against both oracles and the CUDA library.
//!
//! This image has no Rust toolchain (no cargo / rustc / registry), so the file is committed unbuilt; on any box with
//! cargo, drop it into the reference checkout as `poly-commit/examples/dump_reference_vectors.rs` and run
//!     cargo run --release --example dump_reference_vectors -- 4096 tests/golden/reference_dump.bin
//! (dev-dependencies ark-bls12-381 / ark-std already present, poly-commit/Cargo.toml:62-65).
//!
//! Layout (little-endian u64 words, the packed ABI layout of include/pcgpu.h):
//!   "PCGPUREF", version = 1, curve id (0 = BLS12-381), n,
//!   n x [x (6 words, Montgomery), y (6 words), infinity (1 word)]      powers_of_g[0..n]
//!   n x [4 words]                                                       polynomial coefficients, Montgomery Fr
//!   [4 words]                                                           z, Montgomery
//!   [x, y, infinity]                                                    commitment
//!   [x, y, infinity]                                                    proof.w
use ark_bls12_381::{Bls12_381, Fr, G1Affine};
use ark_ec::AffineRepr;
use ark_poly::{univariate::DensePolynomial, DenseUVPolynomial};
use ark_poly_commit::kzg10::{Powers, KZG10};
use ark_std::{test_rng, UniformRand};
use std::{borrow::Cow, fs::File, io::Write};

type Poly = DensePolynomial<Fr>;

fn put_point(out: &mut Vec<u64>, p: &G1Affine) {
    match p.xy() {
        // Fp(pub BigInt<N>, PhantomData): `.0 .0` is the in-memory Montgomery limb array
        Some((x, y)) => { out.extend_from_slice(&x.0 .0); out.extend_from_slice(&y.0 .0); out.push(0); }
        None => { out.extend_from_slice(&[0u64; 12]); out.push(1); }
    }
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let n: usize = args.get(1).map(|s| s.parse().unwrap()).unwrap_or(4096);
    let path = args.get(2).cloned().unwrap_or_else(|| "reference_dump.bin".to_string());
    let rng = &mut test_rng();
    let degree = n - 1;
    let pp = KZG10::<Bls12_381, Poly>::setup(degree, false, rng).unwrap();            // kzg10/mod.rs:47-127
    let powers_of_g = pp.powers_of_g[..=degree].to_vec();
    let powers_of_gamma_g = (0..=1).map(|i| pp.powers_of_gamma_g[&i]).collect::<Vec<_>>();
    let powers = Powers { powers_of_g: Cow::Owned(powers_of_g.clone()), powers_of_gamma_g: Cow::Owned(powers_of_gamma_g) };
    let p = Poly::rand(degree, rng);
    let z = Fr::rand(rng);
    let (comm, rand) = KZG10::<Bls12_381, Poly>::commit(&powers, &p, None, None).unwrap();   // :157-210, non-hiding
    let proof = KZG10::<Bls12_381, Poly>::open(&powers, &p, z, &rand).unwrap();               // :287-310

    let mut w: Vec<u64> = Vec::new();
    w.push(u64::from_le_bytes(*b"PCGPUREF")); w.push(1); w.push(0); w.push(n as u64);
    for b in &powers_of_g { put_point(&mut w, b); }
    for c in p.coeffs() { w.extend_from_slice(&c.0 .0); }
    for _ in p.coeffs().len()..n { w.extend_from_slice(&[0u64; 4]); }                  // DensePolynomial truncates trailing zeros
    w.extend_from_slice(&z.0 .0);
    put_point(&mut w, &comm.0);
    put_point(&mut w, &proof.w);
    let mut f = File::create(&path).unwrap();
    for v in &w { f.write_all(&v.to_le_bytes()).unwrap(); }
    eprintln!("wrote {} words to {}", w.len(), path);
}
