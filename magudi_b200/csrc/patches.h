// Patch handle (t_Patch family).  Public constants mirror include/magudi_gpu.h.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "grid.h"

enum {
  MG_PATCH_FARFIELD = 1,
  MG_PATCH_SPONGE = 2,
  MG_PATCH_SLIP_WALL = 3,
  MG_PATCH_ISOTHERMAL_WALL = 4,
  MG_PATCH_COST_TARGET = 5,
  MG_PATCH_ACTUATOR = 6,
  MG_PATCH_BLOCK_INTERFACE = 7,
  MG_PATCH_KOLMOGOROV_FORCING = 8,
  MG_PATCH_JET_EXCITATION = 9,
  MG_PATCH_PROBE = 10,
  MG_PATCH_ADIABATIC_WALL = 11,
};
constexpr int MG_JET_MAX_MODES = 99;    // src/JetExcitationPatchImpl.f90:53

struct mg_patch {
  struct Array { double* p = nullptr; int nComp = 0; };
  mg_state* state = nullptr;
  int type = 0;
  std::string name;
  int normalDirection = 0;
  int extent[6] = {1, 1, 1, 1, 1, 1};
  int globalSize[3] = {1, 1, 1};
  int localLo[3] = {0, 0, 0}, localSize[3] = {0, 0, 0}, patchOffset[3] = {0, 0, 0};
  int nPatchPoints = 0;
  double inviscidPenaltyAmount = 0.0, viscousPenaltyAmount = 0.0;   // signed, already / normBoundary(1)
  double spongeAmount = 1.0;          // SPONGE: patches/<name>/sponge_amount, sponge_exponent (src/SpongePatchImpl.f90:40-45)
  int spongeExponent = 2;
  // JET_EXCITATION (src/JetExcitationPatchImpl.f90): a sponge-shaped patch that adds eigenmode perturbations
  std::vector<double> angularFrequencies;
  // PROBE (src/ProbePatchImpl.f90): ring of collected solutions, (nPatchPoints, nUnknowns, probeCapacity)
  double* probeBuffer = nullptr;
  int probeCapacity = 0, probeCount = 0;
  // ACTUATOR: gradientBuffer (reference include/ActuatorPatch.f90), (nPatchPoints, gradientCapacity) on the device
  double* gradientBuffer = nullptr;
  int gradientCapacity = 0, gradientCount = 0;
  std::map<std::string, Array> arrays;     // patch-point arrays, (nPatchPoints, nComp) point fastest
  bool AplusReady = false;
  int AplusIncoming = 0;
  // SAT_BLOCK_INTERFACE (reference include/BlockInterfacePatch.f90): the conforming patch of the other block, this
  // patch's index reordering, and the penalty amounts / normal directions of both sides (METRICS exchange)
  mg_patch* partner = nullptr;
  struct mg_p2p* remote = nullptr;     // the conforming patch lives in another process: two-party P2P link
  int reorder[3] = {1, 2, 3};
  bool metricsReady = false;
  double sigmaIL = 0.0, sigmaIR = 0.0, sigmaVL = 0.0, sigmaVR = 0.0;
  int normalL = 0, normalR = 0;
};

int mg_patch_create_impl(mg_state* s, int type, const char* name, int normalDirection, const int extent[6],
                         mg_patch** out);
void mg_patch_destroy_impl(mg_patch* p);
int mg_patch_alloc_array(mg_patch* p, const std::string& name, int nComp, double** out);
int mg_patch_set_array_impl(mg_patch* p, const char* name, int nComp, const double* host);
int mg_patch_get_array_impl(mg_patch* p, const char* name, int nComp, double* host);
int mg_patch_collect_impl(mg_patch* p, const MgField* f, int nComp, const char* name);
int mg_patch_disperse_impl(mg_patch* p, const char* name, int nComp, MgField* f);
int mg_patches_update_impl(mg_state* s);
int mg_patches_sponge_strengths_impl(mg_state* s);
int mg_patches_sponge_arc_length_impl(mg_state* s, int dir, double* hostOut);
int mg_patches_sponge_strengths_gathered_impl(mg_state* s, int dir, const double* arcGathered);
int mg_patch_kolmogorov_setup_impl(mg_patch* p, double amplitude, int wavenumber);
int mg_patch_probe_setup_impl(mg_patch* p, int bufferSize);
int mg_patch_probe_record_impl(mg_patch* p, int mode, int* full);
int mg_patch_probe_flush_impl(mg_patch* p, double* host, int* count);

// block interfaces (interface.cu)
bool mg_state_has_interfaces(const mg_state* s);
int mg_interface_link(mg_patch* a, mg_patch* b, const int reorderA[3]);
int mg_interface_link_remote(mg_patch* p, const int reorder[3], double partnerInviscidAmount,
                             double partnerViscousAmount, int partnerNormalDirection, struct mg_p2p** link);
int mg_interfaces_exchange(const std::vector<mg_state*>& states, int mode);
int mg_interface_apply(mg_state* s, mg_patch* p, int mode);
int mg_interfaces_adjoint_sources(mg_state* s, MgField* temp1);

// functionals
int mg_functional_quadrature_impl(mg_state* s, int patchType, const double* integrandDevice, double* value);
int mg_functional_acoustic_noise_impl(mg_state* s, double timeRampFactor, double* value);
int mg_functional_acoustic_noise_forcing_impl(mg_state* s, double timeRampFactor);
int mg_functional_actuator_sensitivity_impl(mg_state* s, double timeRampFactor, double* value);
int mg_functional_actuator_gradient_impl(mg_patch* p, double timeRampFactor, double* hostOut);
int mg_functional_pressure_drag_impl(mg_state* s, const double direction[3], double* value);
int mg_functional_pressure_drag_forcing_impl(mg_state* s, const double direction[3]);
int mg_functional_accumulate_impl(mg_state* s, int which, double weight, double timeRampFactor);
int mg_functional_accumulator_get_impl(mg_state* s, int which, double* value, int reset);
int mg_patch_gradient_buffer_setup_impl(mg_patch* p, int nSlots);
int mg_functional_actuator_gradient_record_impl(mg_patch* p, double timeRampFactor, int* full);
int mg_patch_gradient_buffer_flush_impl(mg_patch* p, double* host, int* count);
int mg_patch_control_forcing_from_buffer_impl(mg_patch* p, int slot, int firstComponent, int nComponents);
int mg_functional_drag_force_impl(mg_state* s, const double direction[3], double* value);
int mg_functional_reynolds_stress_impl(mg_state* s, const double d1[3], const double d2[3], double* value);
int mg_functional_reynolds_stress_forcing_impl(mg_state* s, const double d1[3], const double d2[3]);
int mg_functional_momentum_actuator_sensitivity_impl(mg_state* s, int direction, double* value);
int mg_functional_momentum_actuator_gradient_impl(mg_patch* p, int direction, double* hostOut);
