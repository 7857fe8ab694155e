// Kernel + host-template instantiations, one translation unit per (curve, group):
//   nvcc -DPCGPU_UNIT_CURVE=Bls12381 -DPCGPU_UNIT_GROUP=1 ... (build.py).  Group 0 defines every group (host emulation).
//   1 pipeline host logic + digit / plan kernels   2 small MSM   3 SRS + MSM / KZG entry points   4 Fr, NTT, hashes
//   5 IPA + wire   6 pair-round kernels   8 XYZZ accumulate   9 bucket reduction
#include "impl.cuh"

#ifndef PCGPU_UNIT_CURVE
#error "compile with -DPCGPU_UNIT_CURVE=<Bls12381|Bn254|Pallas> -DPCGPU_UNIT_GROUP=<0..9>"
#endif
#define PCGPU_UC PCGPU_UNIT_CURVE
#define PCGPU_DEF_OR_EXTERN(GROUP, MACRO) PCGPU_DEF_OR_EXTERN_##GROUP(MACRO)

#if PCGPU_UNIT_GROUP == 0
PCGPU_INSTANTIATE(PCGPU_UC, )
#else
// the helpers other groups call are declared extern everywhere except in their own unit
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
#if PCGPU_UNIT_GROUP == 6
PCGPU_INST_PAIR1(PCGPU_UC, )
#else
PCGPU_INST_PAIR1(PCGPU_UC, extern)
#endif
#if PCGPU_UNIT_GROUP == 8
PCGPU_INST_ACC(PCGPU_UC, )
#else
PCGPU_INST_ACC(PCGPU_UC, extern)
#endif
#if PCGPU_UNIT_GROUP == 9
PCGPU_INST_REDUCE(PCGPU_UC, )
#else
PCGPU_INST_REDUCE(PCGPU_UC, extern)
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
