// Forward sweep B + dissipation (fused_sweepbd.cuh), runtime switches: instantiations.
#include "fused_sweepbd.cuh"

int mg_fused_sweepbd_gen_launch(const void* argsv, int nD, int R, int tileY, int nChunks, cudaStream_t st) {
  const FusedArgs& a = *static_cast<const FusedArgs*>(argsv);
#define MG_B(ND_, R_, DLO, DN, TLO, TN, TY_) \
  if (nD == ND_ && R == R_ && tileY == TY_) return dispatchBD<ND_, R_, DLO, DN, TLO, TN, false, TY_>(a, nChunks, st);
#ifndef MG_DEV_ONLY_33
  MG_B(2, 2, -1, 3, -1, 3, 8)
  MG_B(2, 3, -2, 4, -1, 4, 8)
  MG_B(2, 4, -2, 5, -2, 5, 8)
  MG_B(3, 2, -1, 3, -1, 3, 8)
  MG_B(3, 4, -2, 5, -2, 5, 8)
#endif
  MG_B(3, 3, -2, 4, -1, 4, 8)
#undef MG_B
  return -1;
}
