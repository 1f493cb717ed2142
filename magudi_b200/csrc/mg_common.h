// Internal declarations shared by the host-side C++ and the CUDA kernels of libmagudi_gpu.
// Public C ABI: include/magudi_gpu.h.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#define MG_MAX_INTERIOR 9    // SBP 4-8: offsets -4..4
#define MG_MAX_BWIDTH 16     // adjoint of SBP 4-8: boundaryWidth 12 + 4
#define MG_MAX_BDEPTH 12     // adjoint of SBP 4-8: boundaryDepth 12
#define MG_GHOST_K 4         // ghost planes kept on each side of k for 3-D fields

enum { MG_SYMMETRIC = 0, MG_SKEW_SYMMETRIC = 1, MG_ASYMMETRIC = 2 };
enum { MG_FORWARD = 1, MG_ADJOINT = -1, MG_LINEARIZED = 0 };
enum { MG_PERIODIC_NONE = 0, MG_PERIODIC_PLANE = 1, MG_PERIODIC_OVERLAP = 2 };

void mg_set_error(const std::string& msg);
int mg_cuda_fail(cudaError_t e, const char* file, int line);

#define MG_CUDA(call)                                                      \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) return mg_cuda_fail(e__, __FILE__, __LINE__);  \
  } while (0)
#define MG_TRY(call)            \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != 0) return rc__; \
  } while (0)
#define MG_FAIL(msg)   \
  do {                 \
    mg_set_error(msg); \
    return -1;         \
  } while (0)

// Device-side view of one 1-D operator (t_StencilOperator, reference include/StencilOperator.f90:9-16).
// b1[m][i]: coefficient of the i-th point from the left boundary in closure row m (left side);
// b2[m][i]: coefficient of point (n - width + i) in closure row m counted from the right boundary.
struct MgDevOp {
  int symmetryType, interiorWidth, boundaryWidth, boundaryDepth;
  int lo, nInterior;           // interior offsets lo .. lo+nInterior-1
  int nGhost[2], periodicOffset[2], hasDomainBoundary[2];
  int normDepth;
  double interior[MG_MAX_INTERIOR];
  double normBoundary[MG_MAX_BDEPTH];
  double b1[MG_MAX_BDEPTH][MG_MAX_BWIDTH];
  double b2[MG_MAX_BDEPTH][MG_MAX_BWIDTH];
};

struct mg_stencil {
  std::string scheme;
  MgDevOp op{};
  int direction = 1;
  int procDim = 1, procCoord = 0, isPeriodic = 0;
  MgDevOp* d_op = nullptr;     // lazily uploaded copy
  bool dirty = true;
};

int mg_stencil_upload(mg_stencil* s);

struct mg_options_t {
  double ratioOfSpecificHeats = 1.4;
  int viscosityOn = 0;
  double reynoldsNumberInverse = 0.0;
  double prandtlNumberInverse = 1.0 / 0.72;
  double powerLawExponent = 0.666;
  double bulkViscosityRatio = 0.6;
  int dissipationOn = 0;
  int compositeDissipation = 1;
  double dissipationAmount = 0.0;
  int useTargetState = 1;
  int useContinuousAdjoint = 0;
  int steadyStateSimulation = 0;
};

// A device field: nComp components, each (nz + 2*gk) planes of nx*ny doubles.
struct MgField {
  double* p = nullptr;
  int nComp = 0;
  size_t compStride = 0;       // doubles between components
  size_t interiorOffset = 0;   // doubles from a component's start to its first interior point
  bool owned = false;
  double* comp(int c) const { return p + (size_t)c * compStride + interiorOffset; }
};

struct mg_grid {
  int index = 1;
  int nD = 3;
  int globalSize[3] = {1, 1, 1}, localSize[3] = {1, 1, 1}, offset[3] = {0, 0, 0};
  int periodicityType[3] = {0, 0, 0};
  double periodicLength[3] = {0, 0, 0};
  int isCurvilinear = 1;
  int procDims[3] = {1, 1, 1}, procCoords[3] = {0, 0, 0};
  size_t N = 0;                 // local interior points
  int gk = 0;                   // ghost planes per side in k
  size_t plane = 0;             // nx*ny
  mg_stencil* firstDerivative[3] = {nullptr, nullptr, nullptr};
  mg_stencil* adjointFirstDerivative[3] = {nullptr, nullptr, nullptr};
  mg_stencil* dissipation[3] = {nullptr, nullptr, nullptr};
  mg_stencil* dissipationTranspose[3] = {nullptr, nullptr, nullptr};
  mg_stencil* filter[3] = {nullptr, nullptr, nullptr};      // applyFilter (src/GridImpl.f90:603-615, 1625-1663)
  int compositeDissipation = 1;
  int dissipationOn = 0;
  MgField coordinates, metrics, jacobian, norm, arcLengths, targetMollifier, controlMollifier;
  int* iblank = nullptr;        // device, N ints (nullptr = no holes)
  bool hasHoles = false;
  bool updated = false;
  // scratch
  MgField scratchA, scratchB;
  // peer-to-peer halo of a slab-decomposed grid (set by mg_p2p_create): lets the operator-by-operator path fill
  // the ghost planes of whatever array an operator is applied to along k
  struct mg_p2p* halo = nullptr;
  struct mg_p2p* haloDir[2] = {nullptr, nullptr};   // the same for bricks split along i / j (packed faces)
};
int mg_p2p_exchange_faces(struct mg_p2p* h, const double* in, size_t inCs, int nComp, int width,
                          const double** ghostPrev, const double** ghostNext);
int mg_p2p_create_pair(size_t capacity, struct mg_p2p** out);
double* mg_p2p_pair_outbox(struct mg_p2p* h);
const double* mg_p2p_pair_inbox(struct mg_p2p* h);
int mg_p2p_exchange_pair(struct mg_p2p* h, size_t count, int phase);   // phase 1: push, 2: wait + unpack
extern "C" int mg_p2p_destroy(struct mg_p2p* h);
int mg_p2p_check_all();       // fails when a halo exchange of any live handle has timed out
int mg_p2p_exchange_view(struct mg_p2p* h, const double* comp0, size_t compStride, int nComp, int width);

int mg_field_alloc(const mg_grid* g, int nComp, MgField* f);
void mg_field_free(MgField* f);
int mg_field_zero(const mg_grid* g, MgField* f);
int mg_field_upload(const mg_grid* g, MgField* f, const double* host);     // host (N, nComp) column-major
int mg_field_download(const mg_grid* g, const MgField* f, double* host);

cudaStream_t mg_stream();
// Second (high-priority) stream for overlapped halo exchanges, and the event of the last exchange still in
// flight on it: the next fused sweep takes it (launch_split), anything else waits through mg_halo_wait_pending.
cudaStream_t mg_halo_stream();
void mg_halo_set_pending(cudaEvent_t ev, double bytesPerSide = 0.0);
double mg_halo_pending_bytes();   // payload per face of the exchange in flight (sizes the overlap decision)
bool mg_halo_take_pending(cudaEvent_t* ev);
bool mg_halo_is_pending();
int mg_halo_wait_pending();       // main stream waits for the exchange AND the boundary chunks in flight
int mg_halo_mark_boundary();      // record "boundary chunks enqueued" on the halo stream
int mg_halo_wait_boundary();      // main stream waits for them
void mg_profile_suppress(bool on);
// Tuning switches of the fused kernels (kernel generation, tile heights, k-chunks, L2 prefetch distance):
// mg_tuning_set (C ABI) wins over the environment variable of the same name, which wins over the default.
int mg_tuning_get(const char* name, int dflt);
bool mg_tuning_has(const char* name);
int mg_num_sms();
void mg_count_launches(int n);   // kernels launched by the library (mg_kernel_launch_count)
