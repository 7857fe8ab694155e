// Kernel + host-template instantiations for Pallas (one translation unit per curve so they build in parallel).
#include "impl.cuh"

PCGPU_INSTANTIATE(Pallas, )
