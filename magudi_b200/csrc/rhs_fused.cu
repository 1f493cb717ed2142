// Fused hot-path sweeps (placeholder until the streaming kernels land).
#include "rhs_fused.h"
int mg_fused_supported(const mg_state* s, int mode) { (void)s; (void)mode; return 0; }
