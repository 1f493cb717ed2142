// Solution limits, solution filter and the Jameson RK3 integrator: the per-timestep companions of the
// RHS evaluation in the reference's time loop (src/SolverImpl.f90:805-880).
//
//   findMinimum / findMaximum / isVariableWithinRange   src/GridImpl.f90:1423-1623
//   checkSolutionLimits                                  src/SolverImpl.f90:189-304
//   computeSolutionLimitPenalty                          src/RegionImpl.f90:1001-1092
//   addSolutionLimitPenaltyAdjointForcing                src/RegionImpl.f90:1094-1221
//   applyFilter                                          src/GridImpl.f90:1625-1663
//   substepForwardJamesonRK3                             src/JamesonRK3IntegratorImpl.f90:56-131
#include <cfloat>
#include <cmath>
#include <cstring>
#include <string>

#include "../../include/magudi_gpu.h"
#include "grid.h"
#include "mg_common.h"

#include "stencil_apply.h"

namespace {

constexpr int EXT_BLOCKS = 592, EXT_THREADS = 256;

struct ExtArgs {
  const double* Q;
  size_t csQ, N;
  int nD, which;          // which: 0 density, 1 temperature
  double gamma;
  double *vMin, *vMax;    // [gridDim.x]
  long long *iMin, *iMax;
};

// value at point p of the variable the limits look at
__device__ __forceinline__ double limited_variable(const double* Q, size_t cs, size_t p, int nD, int which,
                                                   double gamma) {
  const double rho = Q[p];
  if (which == 0) return rho;
  const double v = 1.0 / rho;
  double usq = 0.0;
  for (int i = 0; i < nD; ++i) {
    const double u = v * Q[(size_t)(i + 1) * cs + p];
    usq = (i == 0) ? u * u : usq + u * u;
  }
  const double pr = (gamma - 1.0) * (Q[(size_t)(nD + 1) * cs + p] - 0.5 * rho * usq);
  return gamma / (gamma - 1.0) * pr * v;
}

// extremum and the FIRST point (memory order) that attains it, as the reference's strict comparisons give
__global__ void __launch_bounds__(EXT_THREADS) k_extrema(ExtArgs a) {
  double lo = DBL_MAX, hi = -DBL_MAX;
  long long ilo = -1, ihi = -1;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < a.N; p += (size_t)gridDim.x * blockDim.x) {
    const double f = limited_variable(a.Q, a.csQ, p, a.nD, a.which, a.gamma);
    if (f < lo) { lo = f; ilo = (long long)p; }
    if (f > hi) { hi = f; ihi = (long long)p; }
  }
  __shared__ double slo[EXT_THREADS], shi[EXT_THREADS];
  __shared__ long long silo[EXT_THREADS], sihi[EXT_THREADS];
  slo[threadIdx.x] = lo; shi[threadIdx.x] = hi; silo[threadIdx.x] = ilo; sihi[threadIdx.x] = ihi;
  __syncthreads();
  for (int h = EXT_THREADS / 2; h > 0; h >>= 1) {
    if (threadIdx.x < h) {
      const int o = threadIdx.x + h;
      if (silo[o] >= 0 && (silo[threadIdx.x] < 0 || slo[o] < slo[threadIdx.x] ||
                           (slo[o] == slo[threadIdx.x] && silo[o] < silo[threadIdx.x]))) {
        slo[threadIdx.x] = slo[o]; silo[threadIdx.x] = silo[o];
      }
      if (sihi[o] >= 0 && (sihi[threadIdx.x] < 0 || shi[o] > shi[threadIdx.x] ||
                           (shi[o] == shi[threadIdx.x] && sihi[o] < sihi[threadIdx.x]))) {
        shi[threadIdx.x] = shi[o]; sihi[threadIdx.x] = sihi[o];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a.vMin[blockIdx.x] = slo[0]; a.iMin[blockIdx.x] = silo[0];
    a.vMax[blockIdx.x] = shi[0]; a.iMax[blockIdx.x] = sihi[0];
  }
}

struct PenArgs {
  const double* Q;
  size_t csQ, cs, N;
  int nD, which;
  double gamma, lo, hi, factor;
  const int* iblank;
  double* f;       // penalty integrand (k_limit_integrand)
  double* rhs;     // adjoint forcing (k_limit_forcing)
  int rhoOut, tOut;
};

// f = x - max above the range, log(min / x) below it, 0 inside and in holes (src/RegionImpl.f90:1052-1078)
__global__ void k_limit_integrand(PenArgs a) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  const double x = limited_variable(a.Q, a.csQ, p, a.nD, a.which, a.gamma);
  double f = 0.0;
  if (x > a.hi) f = x - a.hi;
  else if (x < a.lo) f = log(a.lo / x);
  if (a.iblank && a.iblank[p] == 0) f = 0.0;
  a.f[p] = f;
}

struct ForceArgs {
  const double* Q;
  size_t csQ, cs, N;
  int nD, rhoOut, tOut;
  double gamma, rhoMin, rhoMax, TMin, TMax, factor;
  const int* iblank;
  double* rhs;
};

// addSolutionLimitPenaltyAdjointForcing (src/RegionImpl.f90:1143-1214)
__global__ void k_limit_forcing(ForceArgs a) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  const double rho = a.Q[p];
  const double v = 1.0 / rho;
  double u[3] = {0.0, 0.0, 0.0}, usq = 0.0;
  for (int i = 0; i < a.nD; ++i) {
    u[i] = v * a.Q[(size_t)(i + 1) * a.csQ + p];
    usq = (i == 0) ? u[i] * u[i] : usq + u[i] * u[i];
  }
  const double rhoE = a.Q[(size_t)(a.nD + 1) * a.csQ + p];
  const double T = a.gamma / (a.gamma - 1.0) * ((a.gamma - 1.0) * (rhoE - 0.5 * rho * usq)) * v;
  double fRho = 0.0, dfRho = 0.0, fT = 0.0, dfT = 0.0;
  if (a.rhoOut) {
    if (rho > a.rhoMax) { fRho = rho - a.rhoMax; dfRho = 1.0; }
    else if (rho < a.rhoMin) { fRho = log(a.rhoMin / rho); dfRho = -1.0 / rho; }
  }
  if (a.tOut) {
    if (T > a.TMax) { fT = T - a.TMax; dfT = 1.0; }
    else if (T < a.TMin) { fT = log(a.TMin / T); dfT = -1.0 / T; }
  }
  if (a.iblank && a.iblank[p] == 0) { fRho = dfRho = fT = dfT = 0.0; }
  double r0 = a.rhs[p] - a.factor * 2.0 * fRho * dfRho;
  if (a.tOut) {
    r0 = r0 - a.factor * 2.0 * fT * dfT * a.gamma * (usq - rhoE / rho) / rho;
    for (int k = 0; k < a.nD; ++k)
      a.rhs[(size_t)(k + 1) * a.cs + p] -= a.factor * 2.0 * fT * dfT * (-a.gamma * u[k] / rho);
    a.rhs[(size_t)(a.nD + 1) * a.cs + p] -= a.factor * 2.0 * fT * dfT * a.gamma / rho;
  }
  a.rhs[p] = r0;
}

struct Rk3Args {
  double *Q, *b1, *b2;
  const double* R;
  size_t cs, N;
  int nU, stage;
  double dt;
};

// substepForwardJamesonRK3 axpys (src/JamesonRK3IntegratorImpl.f90:85-125)
__global__ void k_rk3(Rk3Args a) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  for (int c = 0; c < a.nU; ++c) {
    const size_t q = (size_t)c * a.cs + p;
    const double R = a.R[q];
    if (a.stage == 1) {
      const double Q0 = a.Q[q];
      a.b1[q] = Q0;
      const double Q1 = Q0 + a.dt * R;
      a.Q[q] = Q1;
      a.b2[q] = Q1;
    } else if (a.stage == 2) {
      a.Q[q] = (a.b1[q] + a.Q[q]) / 2.0 + a.dt * R / 2.0;
    } else {
      a.Q[q] = (a.b1[q] + a.b2[q]) / 2.0 + a.dt * R / 2.0;
    }
  }
}

inline unsigned nblocks(size_t n) { return (unsigned)((n + 255) / 256); }

// integrands of computeRegionIntegral / computeAdjointXmomentum (src/RegionImpl.f90:605-730): out = the quantity, the
// norm-weighted sum is the grid inner product with 1
__global__ void k_bf_integrand(const double* Q, size_t csQ, const double* W, size_t csW, int nD, int which, size_t N,
                               double* out, double* one) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= N) return;
  one[p] = 1.0;
  double f = 1.0;
  if (which == 1) f = Q[csQ + p];
  else if (which == 2) f = W[csW + p];
  else if (which == 3) f = W[(size_t)(nD + 1) * csW + p] * ((1.0 / Q[p]) * Q[csQ + p]);
  out[p] = f;
}

struct BfArgs {
  const double *Q, *W;
  size_t csQ, csW, cs, N;
  int nD, mode, stage1;
  double mL, aL, stage1Term;
  const int* iblank;
  double* rhs;
};

// addBodyForce, pointwise part (src/RegionImpl.f90:787-847)
__global__ void k_body_force(BfArgs a) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  if (a.iblank && a.iblank[p] == 0) return;          // the RHS of a hole is zeroed afterwards in the reference
  const double v = 1.0 / a.Q[p];
  const double ux = v * a.Q[a.csQ + p];
  const size_t e = (size_t)(a.nD + 1);
  if (a.mode == MG_FORWARD) {
    a.rhs[a.cs + p] = a.rhs[a.cs + p] + a.mL;
    a.rhs[e * a.cs + p] = a.rhs[e * a.cs + p] + a.mL * ux;
  } else if (a.mode == MG_ADJOINT) {
    const double temp = a.mL * v * a.W[e * a.csW + p];
    double r1 = a.rhs[a.cs + p] - temp;
    a.rhs[p] = a.rhs[p] + temp * ux;
    if (a.stage1) r1 = r1 - a.stage1Term;
    a.rhs[a.cs + p] = r1;
  } else {
    double temp = (0.0 - ux) * a.W[p] + a.W[a.csW + p];
    temp = temp * v;
    a.rhs[a.cs + p] = a.rhs[a.cs + p] + a.aL;
    a.rhs[e * a.cs + p] = a.rhs[e * a.cs + p] + a.aL * ux + a.mL * temp;
  }
}

}  // namespace

// which: 0 volume, 1 integral of rho u, 2 integral of the second adjoint (or perturbation) variable,
// 3 <w_E, u_x> (computeAdjointXmomentum); local to this rank
int mg_state_integral_impl(mg_state* s, int which, double* value) {
  mg_grid* g = s->grid;
  if (!g->updated) MG_FAIL("region integral: grid metrics have not been computed (mg_grid_update)");
  const MgField& Q = s->Q[s->cur];
  const MgField& W = s->W[s->curW];
  if (which != 0 && !Q.p) MG_FAIL("region integral: conserved variables have not been set");
  if (which >= 2 && !W.p) MG_FAIL("region integral: adjoint variables have not been set");
  MG_TRY(mg_halo_wait_pending());
  MgField& B = g->scratchB;
  if (B.nComp < 2) MG_FAIL("region integral: scratch field is too small");
  { k_bf_integrand<<<nblocks(g->N), 256, 0, mg_stream()>>>(Q.p ? Q.comp(0) : nullptr, Q.compStride, W.p ? W.comp(0) : nullptr,
                                                          W.compStride, s->nD, which, g->N, B.comp(0), B.comp(1)); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return mg_grid_inner_product_dev(g, B.comp(0), B.comp(1), nullptr, B.compStride, 1, value);
}

int mg_state_add_body_force_impl(mg_state* s, int mode, double momentumLoss, double adjointMomentumLoss, bool stage1,
                                 double stage1Term) {
  mg_grid* g = s->grid;
  BfArgs a;
  a.Q = s->Q[s->cur].comp(0); a.csQ = s->Q[s->cur].compStride;
  a.W = s->W[s->curW].p ? s->W[s->curW].comp(0) : nullptr; a.csW = s->W[s->curW].compStride;
  if (mode != MG_FORWARD && !a.W) MG_FAIL("body force: adjoint variables have not been set");
  a.cs = s->rhs.compStride; a.N = g->N;
  a.nD = s->nD; a.mode = mode; a.stage1 = stage1 ? 1 : 0;
  a.mL = momentumLoss; a.aL = adjointMomentumLoss; a.stage1Term = stage1Term;
  a.iblank = g->iblank;
  a.rhs = s->rhs.comp(0);
  { k_body_force<<<nblocks(g->N), 256, 0, mg_stream()>>>(a); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return 0;
}

// Local extrema of density (which = 0) or temperature (1) of the conserved variables, with the 1-based GLOBAL
// (i, j, k) of the first point attaining each; the caller combines ranks (MPI_Allgather + minloc in the reference).
int mg_state_extrema_impl(mg_state* s, int which, double* vMin, int ijkMin[3], double* vMax, int ijkMax[3]) {
  mg_grid* g = s->grid;
  if (which != 0 && which != 1) MG_FAIL("mg_state_extrema: variable must be 0 (density) or 1 (temperature)");
  const MgField& Q = s->Q[s->cur];
  if (!Q.p) MG_FAIL("mg_state_extrema: conserved variables have not been set");
  MG_TRY(mg_halo_wait_pending());
  static double *dV = nullptr, *hV = nullptr;
  static long long *dI = nullptr, *hI = nullptr;
  if (!dV) {
    MG_CUDA(cudaMalloc(&dV, 2 * EXT_BLOCKS * sizeof(double)));
    MG_CUDA(cudaMalloc(&dI, 2 * EXT_BLOCKS * sizeof(long long)));
    MG_CUDA(cudaMallocHost(&hV, 2 * EXT_BLOCKS * sizeof(double)));
    MG_CUDA(cudaMallocHost(&hI, 2 * EXT_BLOCKS * sizeof(long long)));
  }
  ExtArgs a;
  a.Q = Q.comp(0); a.csQ = Q.compStride; a.N = g->N;
  a.nD = s->nD; a.which = which; a.gamma = s->opt.ratioOfSpecificHeats;
  a.vMin = dV; a.vMax = dV + EXT_BLOCKS; a.iMin = dI; a.iMax = dI + EXT_BLOCKS;
  { k_extrema<<<EXT_BLOCKS, EXT_THREADS, 0, mg_stream()>>>(a); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  MG_CUDA(cudaMemcpyAsync(hV, dV, 2 * EXT_BLOCKS * sizeof(double), cudaMemcpyDeviceToHost, mg_stream()));
  MG_CUDA(cudaMemcpyAsync(hI, dI, 2 * EXT_BLOCKS * sizeof(long long), cudaMemcpyDeviceToHost, mg_stream()));
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  double lo = DBL_MAX, hi = -DBL_MAX;
  long long ilo = -1, ihi = -1;
  for (int b = 0; b < EXT_BLOCKS; ++b) {
    if (hI[b] >= 0 && (ilo < 0 || hV[b] < lo || (hV[b] == lo && hI[b] < ilo))) { lo = hV[b]; ilo = hI[b]; }
    const int o = EXT_BLOCKS + b;
    if (hI[o] >= 0 && (ihi < 0 || hV[o] > hi || (hV[o] == hi && hI[o] < ihi))) { hi = hV[o]; ihi = hI[o]; }
  }
  auto toIjk = [&](long long p, int* ijk) {
    if (!ijk) return;
    const long long nx = g->localSize[0], ny = g->localSize[1];
    ijk[0] = (int)(p % nx) + g->offset[0] + 1;
    ijk[1] = (int)((p / nx) % ny) + g->offset[1] + 1;
    ijk[2] = (int)(p / (nx * ny)) + g->offset[2] + 1;
  };
  if (vMin) *vMin = lo;
  if (vMax) *vMax = hi;
  toIjk(ilo < 0 ? 0 : ilo, ijkMin);
  toIjk(ihi < 0 ? 0 : ihi, ijkMax);
  return 0;
}

// This rank's share of computeSolutionLimitPenalty BEFORE the factor: sum over the out-of-range variables of
// <f, f> (norm-weighted).  rhoOut / tOut: the range test of the WHOLE grid (all ranks) failed for that variable.
int mg_state_limit_penalty_impl(mg_state* s, const double densityRange[2], const double temperatureRange[2],
                                int rhoOut, int tOut, double* value) {
  mg_grid* g = s->grid;
  if (!g->updated) MG_FAIL("solution-limit penalty: grid metrics have not been computed (mg_grid_update)");
  const MgField& Q = s->Q[s->cur];
  *value = 0.0;
  MgField& A = g->scratchA;
  for (int which = 0; which < 2; ++which) {
    if (!(which == 0 ? rhoOut : tOut)) continue;
    PenArgs a;
    std::memset(&a, 0, sizeof(a));
    a.Q = Q.comp(0); a.csQ = Q.compStride; a.N = g->N;
    a.nD = s->nD; a.which = which; a.gamma = s->opt.ratioOfSpecificHeats;
    a.lo = which == 0 ? densityRange[0] : temperatureRange[0];
    a.hi = which == 0 ? densityRange[1] : temperatureRange[1];
    a.iblank = g->iblank;
    a.f = A.comp(0);
    { k_limit_integrand<<<nblocks(g->N), 256, 0, mg_stream()>>>(a); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
    double r = 0.0;
    MG_TRY(mg_grid_inner_product_dev(g, A.comp(0), A.comp(0), nullptr, A.compStride, 1, &r));
    *value += r;
  }
  return 0;
}

int mg_state_limit_forcing_impl(mg_state* s, const double densityRange[2], const double temperatureRange[2],
                                int rhoOut, int tOut, double penaltyFactor) {
  mg_grid* g = s->grid;
  if (!rhoOut && !tOut) return 0;
  const MgField& Q = s->Q[s->cur];
  ForceArgs a;
  a.Q = Q.comp(0); a.csQ = Q.compStride; a.cs = s->rhs.compStride; a.N = g->N;
  a.nD = s->nD; a.rhoOut = rhoOut; a.tOut = tOut;
  a.gamma = s->opt.ratioOfSpecificHeats;
  a.rhoMin = densityRange[0]; a.rhoMax = densityRange[1];
  a.TMin = temperatureRange[0]; a.TMax = temperatureRange[1];
  a.factor = ((s->opt.useContinuousAdjoint || s->opt.steadyStateSimulation) ? 1.0 : s->adjointForcingFactor) * penaltyFactor;
  a.iblank = g->iblank;
  a.rhs = s->rhs.comp(0);
  { k_limit_forcing<<<nblocks(g->N), 256, 0, mg_stream()>>>(a); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return 0;
}

// Filter operators of the grid: "<filteringScheme> filter" along every direction with more than one point
// (src/GridImpl.f90:603-615); NULL scheme = drop the filters.
int mg_grid_setup_filter_impl(mg_grid* g, const char* filteringScheme) {
  int periodic[3];
  for (int i = 0; i < 3; ++i) periodic[i] = g->periodicityType[i] != MG_PERIODIC_NONE;
  for (int i = 0; i < g->nD; ++i) {
    if (g->filter[i]) { mg_stencil_destroy(g->filter[i]); g->filter[i] = nullptr; }
    if (!filteringScheme) continue;
    const std::string name = g->globalSize[i] > 1 ? std::string(filteringScheme) + " filter" : std::string("null matrix");
    MG_TRY(mg_stencil_create_impl(name.c_str(), &g->filter[i]));
    MG_TRY(mg_stencil_update_impl(g->filter[i], i + 1, g->procDims, g->procCoords, periodic,
                                  g->periodicityType[i] == MG_PERIODIC_OVERLAP));
  }
  return 0;
}

// applyFilter: the directions are visited in an order that rotates with the timestep
int mg_grid_apply_filter_impl(mg_grid* g, MgField* f, int timestep) {
  const int nD = g->nD;
  for (int i = 0; i < nD; ++i)
    if (!g->filter[i]) MG_FAIL("applyFilter: the grid has no filter operators (mg_grid_setup_filter)");
  static const int d1[1] = {1}, d2[2] = {12, 21}, d3[6] = {123, 231, 312, 132, 321, 213};
  const int* dirs = nD == 1 ? d1 : (nD == 2 ? d2 : d3);
  const int nDirs = nD == 1 ? 1 : (nD == 2 ? 2 : 6);
  if (timestep < 0) MG_FAIL("applyFilter: negative timestep");
  const int code = dirs[timestep % nDirs];
  MgField& A = g->scratchA;
  if (A.nComp < f->nComp) MG_FAIL("applyFilter: scratch field is too small");
  MG_TRY(mg_halo_wait_pending());
  bool inScratch = false;
  int pw = 1;
  for (int i = 0; i < nD; ++i) {
    const int j = (code / pw) % 10;
    pw *= 10;
    const MgField& src = inScratch ? A : *f;
    const MgField& dst = inScratch ? *f : A;
    MG_TRY(mg_grid_apply(g, g->filter[j - 1], src.comp(0), src.compStride, dst.comp(0), dst.compStride, f->nComp));
    inScratch = !inScratch;
  }
  if (inScratch) {
    for (int c = 0; c < f->nComp; ++c)
      MG_CUDA(cudaMemcpyAsync(f->comp(c), A.comp(c), g->N * sizeof(double), cudaMemcpyDeviceToDevice, mg_stream()));
  }
  return 0;
}

// substepForwardJamesonRK3 for one state; the RHS has been evaluated by the caller unless !rhsReady
int mg_rk3_substep_impl(mg_state* s, double* time, double dt, int stage) {
  mg_grid* g = s->grid;
  if (stage < 1 || stage > 3) MG_FAIL("rk3 substep: stage must be 1..3");
  if (stage == 1) s->timeProgressive = *time + dt / 2.0;
  if (stage == 2) { *time += dt / 2.0; s->time = *time; s->timeProgressive = *time + dt / 2.0; }
  if (stage == 3) { *time += dt / 2.0; s->time = *time; }
  if (!s->rhsReady) MG_TRY(mg_state_compute_rhs_impl(s, MG_FORWARD));
  s->rhsReady = false;
  MG_TRY(mg_state_make_exclusive(s, &s->Q[s->cur], true));
  MG_TRY(mg_state_make_exclusive(s, &s->rk1, true));
  Rk3Args a;
  a.Q = s->Q[s->cur].comp(0);
  a.b1 = s->rk1.comp(0);
  a.b2 = s->rk2.comp(0);
  a.R = s->rhs.comp(0);
  a.cs = s->rhs.compStride;
  a.N = g->N;
  a.nU = s->nU;
  a.stage = stage;
  a.dt = dt;
  { k_rk3<<<nblocks(g->N), 256, 0, mg_stream()>>>(a); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  s->dependentValid = false;
  s->fusedValid = false;
  return 0;
}
