// Fused hot-path sweeps (rhs_fused.cu).
#pragma once
#include "grid.h"

// Returns 1 when the fused kernels cover this state's configuration including the RK4 substep (no patches, no
// sources); mg_fused_rhs_supported: the sweeps can produce the RHS, patches and sources are applied afterwards.
int mg_fused_supported(const mg_state* s, int mode);
int mg_fused_rhs_supported(const mg_state* s, int mode);

// Second-generation kernels living in their own translation units (args = the caller's FusedArgs block).
int mg_fused_adjoint1_hot_launch(const void* args, int nD, int R, int tileY, int nChunks, cudaStream_t st,
                                 const void* tensorMapW);
int mg_fused_adjoint1_gen_launch(const void* args, int nD, int R, int tileY, int nChunks, cudaStream_t st);
int mg_fused_sweepbd_hot_launch(const void* args, int nD, int R, int tileY, int nChunks, cudaStream_t st);
int mg_fused_sweepbd_gen_launch(const void* args, int nD, int R, int tileY, int nChunks, cudaStream_t st);
