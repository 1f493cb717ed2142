// Internal grid / state / region interfaces (host side).
#pragma once
#include "mg_common.h"
#include "cns_device.cuh"

int mg_grid_create_impl(int index, int nD, const int globalSize[3], const int localSize[3], const int offset[3],
                        const int periodicityType[3], const double periodicLength[3], int isCurvilinear,
                        const int procDims[3], const int procCoords[3], mg_grid** out);
void mg_grid_destroy_impl(mg_grid* g);
int mg_grid_setup_discretization_impl(mg_grid* g, const char* const schemes[3], int dissipationOn,
                                      int compositeDissipation, int useContinuousAdjoint);
int mg_grid_update_impl(mg_grid* g, int* hasNegativeJacobian);
int mg_grid_apply(mg_grid* g, mg_stencil* op, const double* in, size_t inCs, double* out, size_t outCs,
                  int nComp);
int mg_grid_gradient_dev(mg_grid* g, const double* f, size_t fCs, int nComp, MgField* out, MgField* scratch);
int mg_grid_inner_product_dev(mg_grid* g, const double* f, const double* gg, const double* weight, size_t cs,
                              int nComp, double* result);

struct mg_patch;
constexpr int MG_STATE_ACCUMULATORS = 2;

struct mg_state {
  mg_grid* grid = nullptr;
  mg_options_t opt;
  int nD = 3, nU = 5;
  MgField Q[2];                 // conserved variables, double-buffered (fused RK writes the other one)
  int cur = 0;
  MgField W[2];                 // adjoint variables
  int curW = 0;
  MgField target, rhs;
  MgField meanPressure;         // mean pressure of the acoustic-noise functional (AcousticNoise data_)
  MgField meanVelocity;         // mean velocity of the Reynolds-stress functional (ReynoldsStress data_)
  MgField specificVolume, velocity, pressure, temperature, mu, lambda, kappa, stressTensor, heatFlux;
  MgField rk1, rk2;             // RK4 buffers (reference RK4IntegratorImpl.f90:32-35)
  MgField viscFluxCart;         // Cartesian viscous fluxes (nU*nD), kept only when a patch needs them
  bool keepViscousFluxes = false;
  // fused path: outputs of sweep A (unique stress entries + heat flux; dissipation term)
  MgField tauq, dissTerm;
  void* fusedOps[2] = {nullptr, nullptr};   // device operator tables of the fused closure path (fwd, adjoint)
  // Device-resident forward substep states (adjoint replay).  A slot is a VIEW of the nU-component buffer
  // that held Q when it was stored: Q, W, rk1 and the slots draw their storage from `pool`, storing / loading
  // a checkpoint is a pointer assignment, and a shared buffer is replaced by a free one right before anything
  // writes to it (mg_state_make_exclusive).
  std::vector<MgField> checkpoints;
  std::vector<double*> pool;
  size_t poolCursor = 0;        // where the search for a free pooled buffer resumes
  // inputs of the NEXT step, copied from the host into free pool buffers while the current step computes
  // (mg_state_stage_async); mg_state_adopt_staged makes them the conserved / adjoint variables (pointer swap)
  MgField staged[2];
  cudaEvent_t stagedReady[2] = {nullptr, nullptr};
  // device->host reads still in flight from pooled buffers: such a buffer is not handed out as free storage
  struct PendingRead { double* p; cudaEvent_t done; };
  std::vector<PendingRead> pendingReads;
  bool fusedValid = false;
  bool dissValid = false;      // dissTerm holds the dissipation of the current Q
  int useFused = 1;
  double time = 0.0, timeProgressive = 0.0, adjointForcingFactor = 1.0;
  struct Source { double loc[3], amplitude, angularFrequency, gaussianFactor, phase; };
  std::vector<Source> acousticSources;
  std::vector<mg_patch*> patches;
  bool bodyForce = false;           // the region adds the x-momentum body force after the RHS: no RK-fused sweeps
  double* accumulators = nullptr;   // device-resident time quadratures: [0] cost functional, [1] sensitivity
  bool dependentValid = false;
  bool rhsReady = false;            // the region has already evaluated the RHS of this substep (block interfaces)
  // soft solution limits (reference src/RegionImpl.f90:1094-1221, :2002-2005): adjoint forcing of the penalty
  struct SolutionLimits {
    bool soft = false, forcingSwitch = true;
    double densityRange[2] = {0.0, 0.0}, temperatureRange[2] = {0.0, 0.0}, penaltyFactor = 0.0;
    int rhoOut = -1, tOut = -1;     // range test of the whole grid given by the host (decomposed grids); -1: test here
  } limits;
  PhysParams phys() const {
    PhysParams p;
    p.gamma = opt.ratioOfSpecificHeats;
    p.ReInv = opt.reynoldsNumberInverse;
    p.PrInv = opt.prandtlNumberInverse;
    p.powerLaw = opt.powerLawExponent;
    p.bulkRatio = opt.bulkViscosityRatio;
    p.viscous = opt.viscosityOn;
    return p;
  }
};

// computeCoordinateDerivatives (reference src/GridImpl.f90:621-744): d(coordinates)/d(xi_dir), nD components
int mg_grid_coordinate_derivatives(mg_grid* g, int dir, MgField* out);
int mg_state_create_impl(mg_grid* g, const mg_options_t* opt, mg_state** out);
void mg_state_destroy_impl(mg_state* s);
int mg_state_update_impl(mg_state* s, const MgField* Qoverride);
int mg_state_make_exclusive(mg_state* s, MgField* f, bool keepContents);
void mg_state_pool_trim(mg_state* s);
double* mg_state_pool_acquire(mg_state* s, size_t bytes);        // a buffer nothing refers to and nothing is reading
void mg_state_note_pending_read(mg_state* s, double* p, cudaStream_t readStream);
int mg_state_rhs_forward_general(mg_state* s);
int mg_state_rhs_adjoint_general(mg_state* s);
int mg_state_rhs_linearized_general(mg_state* s);
int mg_state_compute_rhs_impl(mg_state* s, int mode);
int mg_state_rhs_pre(mg_state* s, int mode);
int mg_state_rhs_post(mg_state* s, int mode, bool alreadyTimesJacobian = false);
int mg_state_adjoint_finish(mg_state* s, MgField* src, double sign, bool timesJacobian = false);
int mg_state_dependents_from_fused(mg_state* s);
int mg_state_ensure_dependents(mg_state* s);
bool mg_state_uses_fused_rhs(const mg_state* s, int mode);
void mg_rk4_set_times(mg_state* s, int mode, double time, double dt, int stage);
int mg_rk4_substep_impl(mg_state* s, int mode, double* time, double dt, int timestep, int stage);
int mg_state_cfl_dt_impl(mg_state* s, int wantDt, double given, double* result);
int mg_state_extrema_impl(mg_state* s, int which, double* vMin, int ijkMin[3], double* vMax, int ijkMax[3]);
int mg_state_limit_penalty_impl(mg_state* s, const double densityRange[2], const double temperatureRange[2],
                                int rhoOut, int tOut, double* value);
int mg_state_limit_forcing_impl(mg_state* s, const double densityRange[2], const double temperatureRange[2],
                                int rhoOut, int tOut, double penaltyFactor);
int mg_grid_setup_filter_impl(mg_grid* g, const char* filteringScheme);
int mg_grid_apply_filter_impl(mg_grid* g, MgField* f, int timestep);
int mg_rk3_substep_impl(mg_state* s, double* time, double dt, int stage);
int mg_state_integral_impl(mg_state* s, int which, double* value);
int mg_state_add_body_force_impl(mg_state* s, int mode, double momentumLoss, double adjointMomentumLoss, bool stage1,
                                 double stage1Term);
int mg_patches_apply(mg_state* s, int mode);
int mg_patches_collect_viscous(mg_state* s);
int mg_patches_farfield_adjoint_sources(mg_state* s, MgField* temp1);
bool mg_patches_have_farfield(const mg_state* s);
int mg_fused_sweepA(mg_state* s);
int mg_fused_dissipation(mg_state* s);
int mg_fused_sweepB(mg_state* s, int fuseRk, int stage, double dt);
int mg_fused_adjoint1(mg_state* s);
int mg_fused_adjoint2(mg_state* s, int fuseRk, int stage, double dt);
void mg_count_launches(int n);
#define MG_P2P_MAX_COMP 16
void mg_profile_begin(const char* name);
void mg_profile_end();
