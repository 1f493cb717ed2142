// Launch interface of the generic stencil kernels (stencil_apply.cu).
#pragma once
#include "mg_common.h"

struct ApplyArgs {
  const double* in = nullptr;    // first interior point of component 0
  double* out = nullptr;         // must not alias `in`
  size_t inCompStride = 0, outCompStride = 0;
  int nComp = 1;
  int n[3] = {1, 1, 1};          // local interior extents
  int interiorOnly = 0;          // 1: applyAtInteriorPoints (closure rows left untouched)
  int padded = 0;                // 1: out-of-range neighbours are read in place (ghost planes of a padded field)
  const double* ghostPrev = nullptr;   // explicit ghost buffers (nGhost, normalPlaneSize, nComp), reference layout
  const double* ghostNext = nullptr;
  // PLANE periodicity of the coordinate itself (reference src/GridImpl.f90:694-736): component
  // `shiftComp` gets -shiftLen on wrapped/left-ghost reads and +shiftLen on right ones.
  int shiftComp = -1;
  double shiftLen = 0.0;
  int shiftPrev = 0, shiftNext = 0;
  cudaStream_t stream = nullptr;
};

int mg_apply_launch(mg_stencil* s, const ApplyArgs& a);
int mg_norm_launch(mg_stencil* s, double* x, size_t compStride, int nComp, const int n[3], int inverse,
                   cudaStream_t st);
int mg_boundary_launch(mg_stencil* s, const double* in, double* out, size_t compStride, int nComp,
                       const int n[3], int face, int applyThenProject, cudaStream_t st);

int mg_stencil_create_impl(const char* scheme, mg_stencil** out);
int mg_stencil_update_impl(mg_stencil* s, int direction, const int procDims[3], const int procCoords[3],
                           const int periodic[3], int overlap);
int mg_stencil_get_adjoint_impl(const mg_stencil* s, mg_stencil** out);
int mg_stencil_negate_impl(mg_stencil* s);
int mg_stencil_clone_impl(const mg_stencil* s, mg_stencil** out);
