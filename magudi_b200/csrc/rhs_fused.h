// Fused hot-path sweeps (rhs_fused.cu).
#pragma once
#include "grid.h"

// Returns 1 when the fused kernels cover this state's configuration.
int mg_fused_supported(const mg_state* s, int mode);
