// Kernel + host-template instantiations for Bn254 (one translation unit per curve so they build in parallel).
#include "impl.cuh"

PCGPU_INSTANTIATE(Bn254, )
