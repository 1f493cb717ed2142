// Adjoint sweep 1 (second generation), runtime switches: instantiations.
#include "fused_adjoint1.cuh"

int mg_fused_adjoint1_gen_launch(const void* argsv, int nD, int R, int tileY, int nChunks, cudaStream_t st) {
  const FusedArgs& a = *static_cast<const FusedArgs*>(argsv);
  (void)tileY;
#define MG_J(ND_, R_, DLO, DN, TLO, TN, TY_) \
  if (nD == ND_ && R == R_ && tileY == TY_) return dispatchAdj1v2<ND_, R_, DLO, DN, TLO, TN, false, TY_>(a, nChunks, st);
#ifndef MG_DEV_ONLY_33
  MG_J(2, 2, -1, 3, -1, 3, 16)
  MG_J(2, 3, -2, 4, -1, 4, 16)
  MG_J(2, 4, -2, 5, -2, 5, 16)
  MG_J(3, 2, -1, 3, -1, 3, 16)
  MG_J(3, 4, -2, 5, -2, 5, 16)
#endif
  MG_J(3, 3, -2, 4, -1, 4, 16)
#undef MG_J
  return -1;
}
