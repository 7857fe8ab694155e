// Kernel + host-template instantiations for Bls12381 (one translation unit per curve so they build in parallel).
#include "impl.cuh"

PCGPU_INSTANTIATE(Bls12381, )
