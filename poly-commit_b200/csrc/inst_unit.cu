// Kernel + host-template instantiations, one translation unit per (curve, group):
//   nvcc -DPCGPU_UNIT_CURVE=Bls12381 -DPCGPU_UNIT_GROUP=1 ... (build.py).  Group 0 defines every group (host emulation).
#include "impl.cuh"

#ifndef PCGPU_UNIT_CURVE
#error "compile with -DPCGPU_UNIT_CURVE=<Bls12381|Bn254|Pallas> -DPCGPU_UNIT_GROUP=<0..5>"
#endif
#define PCGPU_UC PCGPU_UNIT_CURVE

#if PCGPU_UNIT_GROUP == 0
PCGPU_INSTANTIATE(PCGPU_UC, )
#else
#if PCGPU_UNIT_GROUP == 1
PCGPU_INST_PIPE(PCGPU_UC, )
#else
PCGPU_INST_PIPE(PCGPU_UC, extern)
#endif
#if PCGPU_UNIT_GROUP == 2
PCGPU_INST_SMALL(PCGPU_UC, )
#else
PCGPU_INST_SMALL(PCGPU_UC, extern)
#endif
#if PCGPU_UNIT_GROUP == 3
PCGPU_INST_SRS(PCGPU_UC, )
#endif
#if PCGPU_UNIT_GROUP == 4
PCGPU_INST_FR(PCGPU_UC, )
#endif
#if PCGPU_UNIT_GROUP == 5
PCGPU_INST_IPA(PCGPU_UC, )
#endif
#endif
