// Adjoint sweep 1 (second generation), compile-time viscous + non-composite dissipation: instantiations.
#include "fused_adjoint1.cuh"

int mg_fused_adjoint1_hot_launch(const void* argsv, int nD, int R, int tileY, int nChunks, cudaStream_t st,
                                 const void* tensorMapW) {
  const FusedArgs& a = *static_cast<const FusedArgs*>(argsv);
  if (tensorMapW) {   // TMA-fed k-queue: 3-D, 16 x 12 tiles, R <= 3
    const CUtensorMap* tm = static_cast<const CUtensorMap*>(tensorMapW);
    if (nD == 3 && R == 3 && tileY == 12) return dispatchAdj1v2<3, 3, -2, 4, -1, 4, true, 12, true>(a, nChunks, st, tm);
#ifndef MG_DEV_ONLY_33
    if (nD == 3 && R == 2 && tileY == 12) return dispatchAdj1v2<3, 2, -1, 3, -1, 3, true, 12, true>(a, nChunks, st, tm);
#endif
    return -1;
  }
#define MG_J(ND_, R_, DLO, DN, TLO, TN, TY_) \
  if (nD == ND_ && R == R_ && tileY == TY_) return dispatchAdj1v2<ND_, R_, DLO, DN, TLO, TN, true, TY_>(a, nChunks, st);
#ifndef MG_DEV_ONLY_33
  MG_J(2, 2, -1, 3, -1, 3, 16)
  MG_J(2, 3, -2, 4, -1, 4, 16)
  MG_J(2, 4, -2, 5, -2, 5, 16)
  MG_J(3, 2, -1, 3, -1, 3, 16)
  MG_J(3, 4, -2, 5, -2, 5, 16)
#endif
  MG_J(3, 3, -2, 4, -1, 4, 16)
  MG_J(3, 3, -2, 4, -1, 4, 12)
#ifndef MG_DEV_ONLY_33
  MG_J(3, 2, -1, 3, -1, 3, 12)
#endif
#undef MG_J
  return -1;
}
