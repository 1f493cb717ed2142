/* libmagudi_gpu -- C ABI of the B200-native RHS / adjoint / RK4 engine for magudi.
 *
 * Drop-in boundary for ONE hot path of dreamer2368/magudi: the per-stage right-hand-side evaluation
 * of the compressible Navier-Stokes equations and of its discrete adjoint, advanced by RK4.
 * The reference has no FFI seam for this path (everything is Fortran type-bound procedures); each
 * entry point below replaces one of them and is what a thin `iso_c_binding` layer in the Fortran
 * host binds.  All paths cited are relative to the reference repository root.
 *
 * Conventions
 *   - every function returns 0 on success and a negative code on error; mg_last_error() then holds
 *     the message (the Fortran shim turns that into `gracefulExit`, src/ErrorHandlerImpl.f90:162-204);
 *   - arrays are Fortran-ordered `A(N, nComp)`: point index i + nx*(j + ny*k) fastest, fp64
 *     (`SCALAR_TYPE` real64, include/config.h.in:53-60); N is the LOCAL point count of the rank;
 *   - indices in extents are 1-based and inclusive like bc.dat (negative values already resolved);
 *   - one host thread per handle; the library owns device-resident mirrors of all fields, host
 *     arrays are only touched by the *_set / *_get calls;
 *   - there is NO CPU fallback: every numerical entry point runs CUDA kernels on the current device
 *     and fails if none is available.
 */
#ifndef MAGUDI_GPU_H
#define MAGUDI_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mg_stencil mg_stencil;   /* t_StencilOperator, include/StencilOperator.f90:9-30 */
typedef struct mg_grid mg_grid;         /* t_Grid,            include/Grid.f90:32-73          */
typedef struct mg_state mg_state;       /* t_State,           include/State.f90:51-87         */
typedef struct mg_patch mg_patch;       /* t_Patch family,    include/Patch.f90:9-59          */
typedef struct mg_region mg_region;     /* t_Region,          include/Region.f90:29-64        */

/* Region_enum, include/Region.f90:8-11 */
#define MG_MODE_FORWARD 1
#define MG_MODE_ADJOINT (-1)
#define MG_MODE_LINEARIZED 0

/* Grid_enum periodicity types, include/Grid.f90:17-20 */
#define MG_NONE 0
#define MG_PLANE 1
#define MG_OVERLAP 2

/* patch types: the keys of PatchFactoryImpl.f90:43-87 that lie on the hot path */
#define MG_SAT_FAR_FIELD 1
#define MG_SPONGE 2
#define MG_SAT_SLIP_WALL 3
#define MG_SAT_ISOTHERMAL_WALL 4
#define MG_COST_TARGET 5
#define MG_ACTUATOR 6
#define MG_SAT_BLOCK_INTERFACE 7
#define MG_KOLMOGOROV_FORCING 8   /* src/KolmogorovForcingPatchImpl.f90 */
#define MG_JET_EXCITATION 9       /* src/JetExcitationPatchImpl.f90 (a sponge-shaped patch) */
#define MG_PROBE 10               /* src/ProbePatchImpl.f90 */
#define MG_SAT_ADIABATIC_WALL 11  /* src/AdiabaticWallImpl.f90: its viscous penalties are zero in the reference */

/* field ids for mg_state_set / mg_state_get (t_State members, include/State.f90:59-62) and
 * mg_grid_get / mg_grid_set (t_Grid members, include/Grid.f90:38-40) */
enum {
  MG_Q_CONSERVED = 0, MG_Q_ADJOINT = 1, MG_Q_TARGET = 2, MG_Q_RHS = 3,
  MG_Q_SPECIFIC_VOLUME = 4, MG_Q_VELOCITY = 5, MG_Q_PRESSURE = 6, MG_Q_TEMPERATURE = 7,
  MG_Q_DYNAMIC_VISCOSITY = 8, MG_Q_SECOND_VISCOSITY = 9, MG_Q_THERMAL_DIFFUSIVITY = 10,
  MG_Q_STRESS_TENSOR = 11, MG_Q_HEAT_FLUX = 12,
  /* outputs of the fused sweep A: unique stress entries + heat flux (nD(nD+1)/2 + nD), dissipation term (nU) */
  MG_Q_FUSED_TAUQ = 13, MG_Q_FUSED_DISSIPATION = 14,
  /* direction-3 block of the adjoint diffusion written by the first fused adjoint sweep (nU-1 comps) */
  MG_Q_FUSED_ADJOINT_DIFFUSION3 = 15,
  /* mean pressure of the acoustic-noise functional (t_AcousticNoise data_, src/AcousticNoiseImpl.f90:60-121) */
  MG_Q_MEAN_PRESSURE = 16,
  /* mean velocity of the Reynolds-stress functional (t_ReynoldsStress data_, src/ReynoldsStressImpl.f90:31-66), nD */
  MG_Q_MEAN_VELOCITY = 17,
  MG_G_COORDINATES = 100, MG_G_METRICS = 101, MG_G_JACOBIAN = 102, MG_G_NORM = 103,
  MG_G_ARC_LENGTHS = 104, MG_G_TARGET_MOLLIFIER = 105, MG_G_CONTROL_MOLLIFIER = 106
};

/* t_SolverOptions / t_SimulationFlags members read by the hot path
 * (src/SolverOptionsImpl.f90:46-136, src/SimulationFlagsImpl.f90:24-44) */
typedef struct mg_options {
  double ratioOfSpecificHeats;
  int viscosityOn;
  double reynoldsNumberInverse;
  double prandtlNumberInverse;
  double powerLawExponent;
  double bulkViscosityRatio;
  int dissipationOn;
  int compositeDissipation;
  double dissipationAmount;
  int useTargetState;
  int useContinuousAdjoint;
  int steadyStateSimulation;   /* steady_state_simulation: the adjoint forcing factors are 1 (src/CostTargetPatchImpl.f90:108) */
} mg_options;

/* ------------------------------------------------------------------ library */
int mg_init(int device);                      /* select the CUDA device of this rank */
const char* mg_last_error(void);
int mg_version(void);
int mg_synchronize(void);

/* ------------------------------------------------------------------ t_StencilOperator */
/* %setup(scheme): src/StencilOperatorImpl.f90:1111-2193 */
int mg_stencil_create(const char* scheme, mg_stencil** out);
/* %update(cartesianCommunicator, direction, isPeriodicityOverlapping): :2195-2256.  The Cartesian
 * communicator is replaced by explicit process-grid dims / coordinates / periodicity. */
int mg_stencil_update(mg_stencil* s, int direction, const int procDims[3], const int procCoords[3],
                      const int periodic[3], int overlap);
/* %getAdjoint(adjointOperator): :2285-2372 */
int mg_stencil_get_adjoint(const mg_stencil* s, mg_stencil** out);
/* %cleanup: :2258-2283 */
int mg_stencil_destroy(mg_stencil* s);
/* members: info = {symmetryType, interiorWidth, boundaryWidth, boundaryDepth, nGhost(1:2),
 * periodicOffset(1:2), hasDomainBoundary(1:2), lbound(rhsInterior), size(rhsInterior)} */
int mg_stencil_info(const mg_stencil* s, int info[12]);
/* rhsInterior(size), rhsBoundary1/2(boundaryWidth, boundaryDepth) column-major, normBoundary */
int mg_stencil_coefficients(const mg_stencil* s, double* rhsInterior, double* rhsBoundary1,
                            double* rhsBoundary2, double* normBoundary);
/* %apply(x, gridSize): :2374-2410 -> :35-252.  In place on x(N, nComp); x may be a host or a device
 * pointer.  Ghost points come from this rank itself (single-rank fillGhostPoints). */
int mg_stencil_apply(mg_stencil* s, double* x, int nComp, const int gridSize[3]);
/* %apply with the fillGhostPoints exchange (src/MPIHelperImpl.f90:113-389) done by the caller:
 * ghostPrev / ghostNext are the received buffers (nGhost, normalPlaneSize, nComp), NULL where the
 * operator has no ghost points on that side. */
int mg_stencil_apply_ghosted(mg_stencil* s, double* x, int nComp, const int gridSize[3],
                             const double* ghostPrev, const double* ghostNext);
/* %applyAtInteriorPoints: :254-457 (closure rows of x are left untouched) */
int mg_stencil_apply_interior(mg_stencil* s, double* x, int nComp, const int gridSize[3]);
/* %applyNorm / %applyNormInverse: :838-1107 */
int mg_stencil_apply_norm(mg_stencil* s, double* x, int nComp, const int gridSize[3]);
int mg_stencil_apply_norm_inverse(mg_stencil* s, double* x, int nComp, const int gridSize[3]);
/* %applyAndProjectOnBoundary / %projectOnBoundaryAndApply: :459-836 */
int mg_stencil_apply_and_project_on_boundary(mg_stencil* s, double* x, int nComp, const int gridSize[3],
                                             int faceOrientation);
int mg_stencil_project_on_boundary_and_apply(mg_stencil* s, double* x, int nComp, const int gridSize[3],
                                             int faceOrientation);

/* ------------------------------------------------------------------ t_Grid */
/* %setup: src/GridImpl.f90:142-291.  localSize/offset follow pigeonhole (src/MPIHelperImpl.f90:3-19);
 * only slab decompositions along direction 3 are accepted (procDims = {1,1,P}). */
int mg_grid_create(int index, int nDimensions, const int globalSize[3], const int localSize[3],
                   const int offset[3], const int periodicityType[3], const double periodicLength[3],
                   int isCurvilinear, const int procDims[3], const int procCoords[3], mg_grid** out);
int mg_grid_destroy(mg_grid* g);
/* %setupSpatialDiscretization: :487-619; scheme[d] is e.g. "SBP 3-6" */
int mg_grid_setup_spatial_discretization(mg_grid* g, const char* scheme1, const char* scheme2,
                                         const char* scheme3, int dissipationOn, int compositeDissipation,
                                         int useContinuousAdjoint);
int mg_grid_set(mg_grid* g, int field, const double* host);     /* coordinates, mollifiers */
int mg_grid_get(mg_grid* g, int field, double* host);           /* metrics, jacobian, norm, ... */
int mg_grid_set_iblank(mg_grid* g, const int* iblank);
/* %update: :746-1065 (metrics, Jacobian, norm; jacobian holds 1/det afterwards, :1063) */
int mg_grid_update(mg_grid* g, int* hasNegativeJacobian);
/* %computeGradient: :1172-1421.  f(N, nComp) -> gradF(N, nD*nComp); host or device pointers */
int mg_grid_gradient(mg_grid* g, const double* f, int nComp, double* gradF);
/* %computeInnerProduct: :1067-1170 (local sum; the caller reduces across ranks) */
int mg_grid_inner_product(mg_grid* g, const double* f, const double* gvec, const double* weight,
                          int nComp, double* result);
mg_stencil* mg_grid_operator(mg_grid* g, int which, int direction);   /* 0 first, 1 adjoint, 2 diss, 3 dissT */
/* ghost planes of a slab-decomposed grid: pack the `width` interior planes next to a k-face of a
 * grid/state field into a contiguous device buffer (nComp*width*nx*ny doubles), or unpack a received
 * buffer into the ghost planes of that face.  side 0 = low k, 1 = high k. */
int mg_halo_pack(mg_grid* g, void* owner, int field, int side, int width, double* deviceBuffer);
int mg_halo_unpack(mg_grid* g, void* owner, int field, int side, int width, const double* deviceBuffer);
/* Direct peer-to-peer ghost-plane exchange over NVLink (replaces fillGhostPoints, src/MPIHelperImpl.f90:113-389,
 * for the slab decomposition): each rank creates an exchanger, publishes its CUDA IPC handle
 * (mg_p2p_handle_size bytes), connects to the previous (side 0) / next (side 1) rank's handle (NULL = no
 * neighbour on that side; sameAsOther = 1 when both neighbours are the same rank), and then calls
 * mg_p2p_exchange for every field whose ghost planes are needed.  Exchanges are asynchronous on the library
 * stream and synchronise with the neighbours on the device (sequence flags); mg_p2p_check reports a timeout. */
typedef struct mg_p2p mg_p2p;
int mg_p2p_create(mg_grid* g, int maxComp, int width, mg_p2p** out);
/* the same for bricks split along direction `direction` (0-based): 0 / 1 = faces normal to i / j travel as packed
 * buffers (reference fillGhostPoints, src/MPIHelperImpl.f90:175-296), 2 = mg_p2p_create.  Serves the operator-by-
 * operator path (every operator application along a decomposed direction fills its ghost points through it); the
 * fused sweeps need slabs along direction 3.  Connect with mg_p2p_connect to the previous / next rank ALONG that
 * direction. */
int mg_p2p_create_dir(mg_grid* g, int direction, int maxComp, int width, mg_p2p** out);
int mg_p2p_handle_size(void);
int mg_p2p_get_handle(mg_p2p* h, void* handleOut);
int mg_p2p_connect(mg_p2p* h, int side, const void* peerHandle, int sameAsOther);
int mg_p2p_exchange(mg_p2p* h, void* owner, int field, int width);
/* the same exchange on a second stream, ordered after the work enqueued so far: the next fused sweep launches its
 * interior k-chunks first (they read no ghost plane), then waits for the exchange and launches the first and the
 * last chunk -- the counterpart of "overlapped with interior stencil work" for fillGhostPoints */
int mg_p2p_exchange_overlapped(mg_p2p* h, void* owner, int field, int width);
/* exchange only the components selected by compMask (bit c = component c of the field): e.g. the tau / q field of
 * the fused sweeps before sweep B on a rectilinear grid needs tau_13, tau_23, tau_33, q_3 in its ghost planes
 * (mask 0x134), and the adjoint sweeps read no tau / q ghost plane at all.  overlapped != 0: on the halo stream. */
int mg_p2p_exchange_masked(mg_p2p* h, void* owner, int field, int width, unsigned compMask, int overlapped);
int mg_p2p_check(mg_p2p* h);
int mg_p2p_destroy(mg_p2p* h);

/* ------------------------------------------------------------------ functionals / sensitivities
 * computeQuadratureOnPatches: src/PatchFactoryImpl.f90:376-444 (mask of the patches of one type, iblank, norm);
 * t_AcousticNoise%compute / %computeAdjointForcing: src/AcousticNoiseImpl.f90:123-280 (needs MG_Q_MEAN_PRESSURE
 * and MG_G_TARGET_MOLLIFIER; the forcing fills "adjointForcing" of every COST_TARGET patch);
 * t_ThermalActuator%computeSensitivity / %updateGradient: src/ThermalActuatorImpl.f90:83-159, 383-443
 * (needs MG_G_CONTROL_MOLLIFIER; the gradient sample w_E * mollifier at the patch points, patch order).
 * Values are the LOCAL sums (the caller reduces across ranks); patch type ids as in mg_patch_create. */
int mg_functional_quadrature_on_patches(mg_state* s, int patchType, const double* integrand, double* value);
int mg_functional_acoustic_noise(mg_state* s, double timeRampFactor, double* value);
int mg_functional_acoustic_noise_forcing(mg_state* s, double timeRampFactor);
int mg_functional_actuator_sensitivity(mg_state* s, double timeRampFactor, double* value);
int mg_functional_actuator_gradient(mg_patch* p, double timeRampFactor, double* hostOut);
/* The same quantities WITHOUT a host synchronisation per substep (launch-bound cases such as the 201 x 201
 * AcousticMonopole): the time quadratures J += norm(i) dt I (src/SolverImpl.f90:837-841) and
 * sensitivity += norm(i) dt S (:1181-1185) are accumulated on the device -- which = 0: acoustic noise, 1: thermal
 * actuator sensitivity; value added = weight * functional -- and read (and optionally reset) once at the end; the
 * gradient samples go to a device-side gradientBuffer (src/ActuatorPatchImpl.f90:226-458) read back in blocks of
 * controller_buffer_size; the control forcing of a substep is selected on the device from the uploaded
 * "controlForcingBuffer" array ((nPatchPoints, nComponents, nSlots) via mg_patch_set_array) into components
 * [firstComponent, firstComponent + nComponents) of "controlForcing", the others zero (updateForcing,
 * src/ThermalActuatorImpl.f90:161-233).  Bit-identical to the synchronising calls. */
int mg_functional_accumulate(mg_state* s, int which, double weight, double timeRampFactor);
int mg_functional_accumulator_get(mg_state* s, int which, double* value, int reset);
int mg_patch_gradient_buffer_setup(mg_patch* p, int nSlots);
int mg_functional_actuator_gradient_record(mg_patch* p, double timeRampFactor, int* bufferIsFull);
int mg_patch_gradient_buffer_flush(mg_patch* p, double* host, int* nRecords);
int mg_patch_control_forcing_from_buffer(mg_patch* p, int slot, int firstComponent, int nComponents);
/* t_PressureDrag%compute / %computeAdjointForcing (src/PressureDragImpl.f90:61-132, 148-267; magudi.inp
 * drag_direction_x/y/z, normalised here): J = sum over the COST_TARGET patches (on a boundary face) of
 * -(p - 1/gamma) patch%norm (metrics_k . direction) / normBoundary(1); the forcing (discrete or continuous
 * adjoint per the state's options) is written to the patches' "adjointForcing". */
int mg_functional_pressure_drag(mg_state* s, const double direction[3], double* value);
int mg_functional_pressure_drag_forcing(mg_state* s, const double direction[3]);
/* t_DragForce%compute (src/DragForceImpl.f90:61-146): the viscous drag sum_l direction_l (metrics_k . tau_l) /
 * normBoundary(1) on the COST_TARGET patches, weighted by the target mollifier; 0 for an inviscid state.  (The
 * reference's computeDragForceAdjointForcing, :159-209, is hard-wired to metrics(:,1) / metrics(:,5) of a 3-D grid;
 * the host composes it from mg_stencil_project_boundary_and_apply / mg_stencil_apply_norm.) */
int mg_functional_drag_force(mg_state* s, const double direction[3], double* value);
/* t_ReynoldsStress%compute / %computeAdjointForcing (src/ReynoldsStressImpl.f90:121-195, 211-284): needs
 * MG_Q_MEAN_VELOCITY and MG_G_TARGET_MOLLIFIER; directions as reynolds_stress_direction1/2_x/y/z, normalised here.
 * The forcing reproduces the reference's assignments literally (its second pair overwrites the first). */
int mg_functional_reynolds_stress(mg_state* s, const double direction1[3], const double direction2[3], double* value);
int mg_functional_reynolds_stress_forcing(mg_state* s, const double direction1[3], const double direction2[3]);
/* t_MomentumActuator%computeSensitivity / %updateGradient (src/MomentumActuatorImpl.f90:81-163, 351-412):
 * direction 0 = all momentum components (nD gradient components per patch point), d > 0 = component d only,
 * -1 = t_GenericActuator (src/GenericActuatorImpl.f90:77-149, 319-376): every unknown (nUnknowns components). */
int mg_functional_momentum_actuator_sensitivity(mg_state* s, int direction, double* value);
int mg_functional_momentum_actuator_gradient(mg_patch* p, int direction, double* hostOut);

/* ------------------------------------------------------------------ t_State */
/* %setup: src/StateImpl.f90:71-170 */
int mg_state_create(mg_grid* g, const mg_options* options, mg_state** out);
int mg_state_destroy(mg_state* s);
int mg_state_set(mg_state* s, int field, const double* host);
int mg_state_get(mg_state* s, int field, double* host);
/* Transfers that overlap the sweeps (the reference keeps its arrays in host RAM, so it has no counterpart):
 * _async variants run on a separate copy stream, ordered after everything enqueued so far on the compute
 * stream.  set_async: call mg_transfer_fence() before the first sweep that uses the field.  get_async /
 * checkpoint_get_async: the source must stay unmodified until mg_transfer_wait() (a checkpoint slot always
 * does).  Host buffers should be pinned. */
int mg_state_set_async(mg_state* s, int field, const double* pinnedHost);
int mg_state_get_async(mg_state* s, int field, double* pinnedHost);
int mg_state_checkpoint_get_async(mg_state* s, int slot, double* pinnedHost);
/* Double-buffered inputs for back-to-back steps: stage the NEXT step's conserved / adjoint variables into a free
 * device buffer while the current step computes (host->device on its own stream, device->host reads on another:
 * both copy engines run beside the sweeps), then adopt them (pointer swap, no copy) when the step begins.  A buffer
 * a pending device->host read still uses is never handed out as free storage. */
int mg_state_stage_async(mg_state* s, int field, const double* pinnedHost);
int mg_state_adopt_staged(mg_state* s, int field);
int mg_transfer_fence(void);   /* compute stream waits for the transfers issued so far */
int mg_transfer_wait(void);    /* host waits for the transfers issued so far */
int mg_state_set_time(mg_state* s, double time);
/* acoustic sources: src/StateImpl.f90:135-148, src/AcousticSourceImpl.f90:3-32 */
int mg_state_add_acoustic_source(mg_state* s, const double location[3], double amplitude, double frequency,
                                 double radius, double phase);
/* %update: src/StateImpl.f90:466-537 */
int mg_state_update(mg_state* s);
/* t_State%computeCfl / %computeTimeStepSize (reference include/State.f90:84-85, src/StateImpl.f90:548-600,
 * src/CNSHelperImpl.f90:842-982) for the current conserved variables, local to this rank: the caller takes the
 * MAX (cfl) / MIN (time step) over ranks as t_Region%getCfl / %getTimeStepSize do (src/RegionImpl.f90:1837-1873). */
int mg_state_cfl(mg_state* s, double timeStepSize, double* cfl);
int mg_state_dt(mg_state* s, double cfl, double* timeStepSize);
/* device-resident substep buffer of the UniformCheckpointer (src/UniformCheckpointerImpl.f90:78-208):
 * keep / restore the conserved variables of a substep in HBM instead of host RAM */
int mg_state_checkpoint_store(mg_state* s, int slot);
int mg_state_checkpoint_load(mg_state* s, int slot);
int mg_state_checkpoint_clear(mg_state* s);

/* ------------------------------------------------------------------ t_Patch */
/* %setup: src/PatchImpl.f90:3-151 and the derived types' setup; amounts are the magudi.inp values
 * (defaults/inviscid_penalty_amount, .../viscous_penalty_amount); sign and 1/normBoundary(1) are
 * applied here as in src/FarFieldPatchImpl.f90:53-69.  SPONGE: sponge_amount, sponge_exponent. */
int mg_patch_create(mg_state* s, int type, const char* name, int normalDirection, const int extent[6],
                    double inviscidPenaltyAmount, double viscousPenaltyAmount, mg_patch** out);
int mg_patch_num_points(const mg_patch* p, int* nPatchPoints, int localSize[3], int patchOffset[3]);
/* named patch arrays (nPatchPoints, nComp): "spongeStrength", "temperature", "adjointForcing",
 * "controlForcing", "targetViscousFluxes", ... */
int mg_patch_set_array(mg_patch* p, const char* name, int nComp, const double* host);
int mg_patch_get_array(mg_patch* p, const char* name, int nComp, double* host);
/* %collect: src/PatchImpl.f90:187-316 -- grid field of the owning state -> patch array */
int mg_patch_collect(mg_patch* p, int field, const char* name);
/* Block interfaces (t_BlockInterfacePatch, src/BlockInterfacePatchImpl.f90:3-929): "a conforms_with b" of
 * readPatchInterfaceInformation (src/InterfaceHelperImpl.f90:3-112, the bc.dat / *interface_index_reorder keys):
 * links two MG_SAT_BLOCK_INTERFACE patches of two states of ONE region; indexReorderingA is patch a's reordering
 * (NULL = 1,2,3), patch b gets the inverted one.  The METRICS pseudo-exchange (src/SolverImpl.f90:571-603) and
 * the per-stage exchangeInterfaceData (src/InterfaceHelperImpl.f90:115-239) run inside mg_region_compute_rhs:
 * both blocks live on this process's device, the exchange is one gather kernel per patch with the reordering
 * applied on the fly. */
int mg_patch_link_interface(mg_patch* a, mg_patch* b, const int indexReorderingA[3]);
/* The conforming patch belongs to a block held by ANOTHER process (another GPU of the node; the reference's
 * exchangeInterfaceData between block communicators, src/InterfaceHelperImpl.f90:115-239): creates a two-party
 * P2P link of 2 * nUnknowns * nPatchPoints doubles each way.  The host swaps the links' IPC handles
 * (mg_p2p_get_handle) between the two processes and connects BOTH sides of the link to the partner's handle
 * (mg_p2p_connect(link, 0, h, 0); mg_p2p_connect(link, 1, h, 1)).  indexReordering is THIS patch's reordering
 * (the inverse of the partner's); the partner's penalty amounts (its mg_patch_penalty_amounts) and normal
 * direction are what the METRICS pseudo-exchange carries besides the metrics
 * (src/BlockInterfacePatchImpl.f90:667-703).  Per stage the collected face arrays travel as remote NVLink stores
 * sequenced by device flags; every process pushes on all its links before it waits on any.  The link is destroyed
 * with the patch. */
/* the patch's penalty amounts as the SAT kernels use them: signed by the normal direction and divided by the
 * boundary norm weight (src/BlockInterfacePatchImpl.f90:70-96); viscous = 0 when the state's viscosity is off */
int mg_patch_penalty_amounts(mg_patch* p, double* inviscid, double* viscous);
int mg_patch_link_interface_remote(mg_patch* p, const int indexReordering[3], double partnerInviscidPenaltyAmount,
                                   double partnerViscousPenaltyAmount, int partnerNormalDirection, mg_p2p** link);

/* ------------------------------------------------------------------ t_Region / t_RK4Integrator */
int mg_region_create(mg_region** out);
int mg_region_destroy(mg_region* r);
int mg_region_add_state(mg_region* r, mg_state* s);
/* updatePatchFactories: src/PatchFactoryImpl.f90:446-574 (target viscous fluxes of far-field patches) */
int mg_region_update_patches(mg_region* r);
/* computeSpongeStrengths: src/PatchFactoryImpl.f90:161-374 -- fills the "spongeStrength" array of every SPONGE
 * patch from the grid's arc lengths.  For SPONGE patches mg_patch_create's two amounts are sponge_amount and
 * sponge_exponent (src/SpongePatchImpl.f90:40-45; reference defaults 1.0 and 2).  Sponges along a decomposed
 * direction are refused here: use the two calls below around the host's gatherAlongDirection. */
int mg_region_compute_sponge_strengths(mg_region* r);
/* computeSpongeStrengths on a grid decomposed along `direction` (1-based), split around the reference's
 * gatherAlongDirection (src/PatchFactoryImpl.f90:213-226, src/MPIHelperImpl.f90:298-402):
 * mg_state_sponge_arc_length writes the rank's arc lengths sqrt(sum (d coordinates / d xi_direction)^2) to the host
 * array arcLength (nGridPoints, i fastest); the host gathers them along the direction;
 * mg_state_sponge_strengths_gathered takes the gathered lines (local sizes in the other directions,
 * globalSize(direction) along it, i fastest) and fills "spongeStrength" of the state's SPONGE / JET_EXCITATION patches
 * whose normal is +-direction (src/PatchFactoryImpl.f90:232-363).  Also valid on an undecomposed direction. */
int mg_state_sponge_arc_length(mg_state* s, int direction, double* arcLength);
int mg_state_sponge_strengths_gathered(mg_state* s, int direction, const double* arcLengthsAlongDirection);
/* %computeRhs(mode, timestep, stage): src/RegionImpl.f90:1877-2027.  MG_MODE_LINEARIZED evaluates
 * computeRhsLinearized (src/RhsHelperImpl.f90:598-829) and the LINEARIZED branches of the patches: the perturbation
 * is the state's adjointVariables, as in the reference. */
int mg_region_compute_rhs(mg_region* r, int mode, int timestep, int stage);
/* t_RK4Integrator%substepForward / %substepAdjoint / %substepLinearized: src/RK4IntegratorImpl.f90:65-369
 * (mode selects the scheme; LINEARIZED advances the adjointVariables with the forward weights).  Includes the
 * states%update the reference's drivers issue after every substep (src/SolverImpl.f90:831-834) when
 * updateStates != 0. */
int mg_rk4_substep(mg_region* r, int mode, double* time, double dt, int timestep, int stage, int updateStates);
/* the adjoint substep in two phases, so that a slab-decomposed host can exchange ghost planes in between:
 * phase 1 = first adjoint sweep (needs ghost planes of the adjoint variables), phase 2 = second sweep +
 * RK4 update (needs ghost planes of MG_Q_FUSED_ADJOINT_DIFFUSION3).  Fused path only. */
int mg_rk4_substep_adjoint_phase(mg_region* r, int phase, double* time, double dt, int timestep, int stage);
/* enable_body_force (src/SimulationFlagsImpl.f90:42, src/SolverImpl.f90:760-765, addBodyForce src/RegionImpl.f90:732-851):
 * the x-momentum conserving body force that mg_region_compute_rhs / mg_rk4_substep add after the sources.  Call after
 * the states (with their grids updated) have been added; initialMomentumPerVolume is body_force/initial_momentum.  The
 * stage argument of mg_region_compute_rhs is the reference's: the loss is re-evaluated at stage 1, the adjoint
 * bookkeeping runs over stages 4..1.  One process per grid (the region integrals are not reduced over ranks). */
int mg_region_set_body_force(mg_region* r, int enable, double initialMomentumPerVolume, double timeStepSize);
int mg_region_get_body_force(mg_region* r, double* momentumLossPerVolume, double* adjointMomentumLossPerVolume);
/* t_JamesonRK3Integrator%substepForward (src/JamesonRK3IntegratorImpl.f90:56-131; the reference's adjoint and
 * linearized RK3 substeps are empty): stage 1..3, same conventions as mg_rk4_substep. */
int mg_rk3_substep(mg_region* r, double* time, double dt, int timestep, int stage, int updateStates);

/* ------------------------------------------------------------------ remaining patch types, limits, filter
 * KOLMOGOROV_FORCING: forcePerUnitMass = amplitude sin(2 pi wavenumber y) at the patch points
 * (setupKolmogorovForcingPatch, src/KolmogorovForcingPatchImpl.f90:3-68; the keys patches/<name>/amplitude,
 * /wavenumber); mg_patch_set_array("forcePerUnitMass") overrides it.  updateRhs: FORWARD / ADJOINT / LINEARIZED. */
int mg_patch_kolmogorov_setup(mg_patch* p, double amplitude, int wavenumber);
/* JET_EXCITATION (src/JetExcitationPatchImpl.f90:3-187): create it like a SPONGE (amounts = amplitude, exponent;
 * mg_region_compute_sponge_strengths fills its spongeStrength), give the eigenmodes with
 * mg_patch_set_array("perturbationReal" | "perturbationImag", nUnknowns * nModes) -- (nPatchPoints, nUnknowns, nModes)
 * as read from <prefix>-NN.eigenmode_real/imag.q -- and their angular frequencies (aux(2) of those files) here. */
int mg_patch_set_jet_modes(mg_patch* p, int nModes, const double* angularFrequencies);
/* PROBE (src/ProbePatchImpl.f90:3-183, saveProbeData src/RegionImpl.f90:2211-2281): the probe buffer
 * (nPatchPoints, nUnknowns, probe_buffer_size) lives on the device.  record collects the conserved (FORWARD) or
 * adjoint (ADJOINT) variables into the next slot and tells when the buffer is full; flush copies the nRecords filled
 * slots to the host (the caller appends them to <prefix>.probe_<name>.dat) and empties the buffer. */
int mg_patch_probe_setup(mg_patch* p, int probeBufferSize);
int mg_patch_probe_record(mg_patch* p, int mode, int* bufferIsFull);
int mg_patch_probe_flush(mg_patch* p, double* host, int* nRecords);
/* findMinimum / findMaximum (src/GridImpl.f90:1423-1577) of the density (variable 0) or the temperature (1) of the
 * conserved variables on this rank: values and the 1-based global (i, j, k) of the first point attaining them.  The
 * host combines ranks and applies isVariableWithinRange / checkSolutionLimits (src/GridImpl.f90:1579-1623,
 * src/SolverImpl.f90:189-304: out of range when min <= minValue or max >= maxValue). */
int mg_state_extrema(mg_state* s, int variable, double* vMin, int ijkMin[3], double* vMax, int ijkMax[3]);
/* this rank's share of computeSolutionLimitPenalty BEFORE the penalty factor (src/RegionImpl.f90:1001-1092): the
 * norm-weighted <f, f> of every variable whose range test failed on the whole grid */
int mg_state_solution_limit_penalty(mg_state* s, const double densityRange[2], const double temperatureRange[2],
                                    int densityOutOfRange, int temperatureOutOfRange, double* value);
/* soft_solution_limits: computeRhs(ADJOINT) adds addSolutionLimitPenaltyAdjointForcing after the patch penalties
 * (src/RegionImpl.f90:2002-2005, :1094-1221) while the switch is on (the adjoint driver turns it off for the terminal
 * step).  Call after the states have been added.  The range test is done on this rank unless the host gives the
 * result for the whole (decomposed) grid with mg_state_set_solution_limit_flags (-1, -1 = test locally again). */
int mg_region_set_solution_limits(mg_region* r, int soft, const double densityRange[2],
                                  const double temperatureRange[2], double penaltyFactor);
int mg_region_solution_limit_forcing_switch(mg_region* r, int on);
int mg_state_set_solution_limit_flags(mg_state* s, int densityOutOfRange, int temperatureOutOfRange);
/* filter_solution (src/GridImpl.f90:603-615, applyFilter :1625-1663): "<filteringScheme> filter" operators
 * ("Standard 5-point", "DRP 9-point"; NULL removes them) and their application to the conserved or adjoint
 * variables, the directions visited in the order that rotates with the timestep */
int mg_grid_setup_filter(mg_grid* g, const char* filteringScheme);
int mg_state_apply_filter(mg_state* s, int field, int timestep);

/* select the implementation: 0 = general operator-by-operator path, 1 = fused sweeps (default when
 * the configuration is covered) */
int mg_region_set_fused(mg_region* r, int enable);
int mg_region_uses_fused(mg_region* r, int mode);
/* 1 when the RHS of every state is evaluated by the fused sweeps: either mg_region_uses_fused, or fused sweeps
 * followed by the patch / source epilogue (states with SAT patches, sponges, cost-target / actuator patches, acoustic
 * sources on one rank; the RK4 update is then a separate pointwise kernel). */
int mg_region_uses_fused_rhs(mg_region* r, int mode);
/* number of hot-path kernels this library has launched since mg_init (bench accounting) */
long long mg_kernel_launch_count(void);
/* the CUDA stream (cudaStream_t) all kernels of this rank are launched on, for event timing */
void* mg_stream_handle(void);
/* per-kernel device timing with CUDA events on that stream: enable, run, then query the summed
 * duration (ms) and launch count of one kernel family ("sweepA", "sweepB", ...); reset with enable */
int mg_profile_enable(int enable);
int mg_profile_get(const char* name, double* milliseconds, long long* launches);
/* tuning switches of the fused sweeps (no reference counterpart): "MG_FWD" (2 = dissipation folded into sweep
 * B, 1 = separate dissipation sweep), "MG_ADJ1" (2 / 1 = generation of adjoint sweep 1), "MG_BD_TY",
 * "MG_ADJ1_TY" (tile heights), "MG_CHUNKS" (k-chunks per tile column, 0 = automatic), "MG_PREFETCH",
 * "MG_PREFETCH_ADJ", "MG_PF_MASK" (L2 prefetch).  A value set here overrides the environment variable of the
 * same name; mg_tuning_clear() forgets every override. */
int mg_tuning_set(const char* name, int value);
int mg_tuning_clear(void);

#ifdef __cplusplus
}
#endif
#endif /* MAGUDI_GPU_H */
