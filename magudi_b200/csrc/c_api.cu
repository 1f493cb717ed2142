// extern "C" entry points of libmagudi_gpu (declared in include/magudi_gpu.h).
#include <atomic>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <cstdlib>

#include "../../include/magudi_gpu.h"
#include "grid.h"
#include "patches.h"
#include "rhs_fused.h"
#include "stencil_apply.h"

namespace {
thread_local std::string g_error;
cudaStream_t g_stream = nullptr;
cudaStream_t g_copyStream = nullptr;   // host -> device transfers that overlap the sweeps (mg_state_*_async)
cudaStream_t g_readStream = nullptr;   // device -> host transfers: the two directions use separate copy engines
int g_device = -1;
int g_sms = 0;
std::atomic<long long> g_launches{0};
}  // namespace

void mg_set_error(const std::string& msg) { g_error = msg; }
int mg_cuda_fail(cudaError_t e, const char* file, int line) {
  g_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + file + ":" + std::to_string(line);
  return -2;
}
cudaStream_t mg_stream() { return g_stream; }

namespace {
cudaStream_t g_haloStream = nullptr;
cudaEvent_t g_haloPending = nullptr;
bool g_haloIsPending = false;
}  // namespace
cudaStream_t mg_halo_stream() {
  if (!g_haloStream) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&g_haloStream, cudaStreamNonBlocking, hi) != cudaSuccess) {
      cudaGetLastError();
      g_haloStream = g_stream;
    }
  }
  return g_haloStream;
}
double g_haloPendingBytes = 0.0;
void mg_halo_set_pending(cudaEvent_t ev, double bytesPerSide) { g_haloPending = ev; g_haloIsPending = true; g_haloPendingBytes = bytesPerSide; }
bool mg_halo_is_pending() { return g_haloIsPending; }
double mg_halo_pending_bytes() { return g_haloIsPending ? g_haloPendingBytes : 0.0; }
bool mg_halo_take_pending(cudaEvent_t* ev) {
  if (!g_haloIsPending) return false;
  *ev = g_haloPending;
  g_haloIsPending = false;
  return true;
}
namespace {
cudaEvent_t g_boundaryEv[4] = {nullptr, nullptr, nullptr, nullptr};
int g_boundaryNext = 0;
cudaEvent_t g_boundaryPending = nullptr;
bool g_profSuppress = false;
}  // namespace
void mg_profile_suppress(bool on) { g_profSuppress = on; }
int mg_halo_mark_boundary() {
  if (!g_boundaryEv[0])
    for (auto& e : g_boundaryEv) MG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaEvent_t e = g_boundaryEv[g_boundaryNext];
  g_boundaryNext = (g_boundaryNext + 1) % 4;
  MG_CUDA(cudaEventRecord(e, mg_halo_stream()));
  g_boundaryPending = e;
  return 0;
}
int mg_halo_wait_boundary() {
  if (g_boundaryPending) {
    MG_CUDA(cudaStreamWaitEvent(g_stream, g_boundaryPending, 0));
    g_boundaryPending = nullptr;
  }
  return 0;
}
int mg_halo_wait_pending() {
  cudaEvent_t ev;
  if (mg_halo_take_pending(&ev)) MG_CUDA(cudaStreamWaitEvent(g_stream, ev, 0));
  return mg_halo_wait_boundary();
}

namespace {
std::map<std::string, int> g_tuning;
std::mutex g_tuningMutex;
}  // namespace
bool mg_tuning_has(const char* name) {
  std::lock_guard<std::mutex> lock(g_tuningMutex);
  return g_tuning.count(name) || getenv(name);
}
int mg_tuning_get(const char* name, int dflt) {
  std::lock_guard<std::mutex> lock(g_tuningMutex);
  auto it = g_tuning.find(name);
  if (it != g_tuning.end()) return it->second;
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
int mg_num_sms() { return g_sms; }
void mg_count_launches(int n) { g_launches += n; }

// ---- optional per-kernel event timing (bench.py roofline accounting)
namespace {
struct ProfEntry { std::string name; cudaEvent_t e0, e1; };
std::vector<ProfEntry> g_prof;
bool g_profOn = false;
}  // namespace
bool mg_profile_on() { return g_profOn; }
void mg_profile_begin(const char* name) {
  if (!g_profOn || g_profSuppress) return;
  ProfEntry p;
  p.name = name;
  cudaEventCreate(&p.e0);
  cudaEventCreate(&p.e1);
  cudaEventRecord(p.e0, g_stream);
  g_prof.push_back(p);
}
void mg_profile_end() {
  if (!g_profOn || g_profSuppress || g_prof.empty()) return;
  cudaEventRecord(g_prof.back().e1, g_stream);
}

struct mg_region {
  std::vector<mg_state*> states;
  int fused = 1;
  // x-momentum conserving body force (reference include/Region.f90:44-45, src/RegionImpl.f90:732-851)
  struct BodyForce {
    bool enabled = false;
    double initialXmomentum = 0.0, oneOverVolume = 0.0, dt = 0.0;
    double momentumLoss = 0.0, adjointMomentumLoss = 0.0;
  } bf;
};

static bool is_device_pointer(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Stage x(N, nComp) on the device (copy if it is a host array), run f(dev_in, dev_out), copy back.
template <typename F>
static int with_device_array(double* x, size_t count, bool needOut, F f) {
  if (g_device < 0) MG_FAIL("libmagudi_gpu: mg_init has not been called (no CUDA device selected)");
  const bool dev = is_device_pointer(x);
  double *din = nullptr, *dout = nullptr;
  const size_t bytes = count * sizeof(double);
  if (dev) din = x;
  else {
    MG_CUDA(cudaMalloc(&din, bytes));
    MG_CUDA(cudaMemcpyAsync(din, x, bytes, cudaMemcpyHostToDevice, g_stream));
  }
  if (needOut) MG_CUDA(cudaMalloc(&dout, bytes));
  int rc = f(din, dout);
  if (rc == 0) {
    const double* src = needOut ? dout : din;
    if (!dev || needOut) {
      cudaError_t e = cudaMemcpyAsync(x, src, bytes, cudaMemcpyDefault, g_stream);
      if (e != cudaSuccess) rc = mg_cuda_fail(e, __FILE__, __LINE__);
    }
  }
  cudaError_t e = cudaStreamSynchronize(g_stream);
  if (rc == 0 && e != cudaSuccess) rc = mg_cuda_fail(e, __FILE__, __LINE__);
  if (!dev) cudaFree(din);
  if (dout) cudaFree(dout);
  return rc;
}

extern "C" {

int mg_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    MG_FAIL("mg_init: no CUDA device available (libmagudi_gpu has no CPU fallback)");
  }
  if (device < 0 || device >= count) MG_FAIL("mg_init: invalid device index");
  MG_CUDA(cudaSetDevice(device));
  if (g_stream && g_device != device) { cudaStreamDestroy(g_stream); g_stream = nullptr; }
  if (!g_stream) MG_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  if (g_copyStream && g_device != device) { cudaStreamDestroy(g_copyStream); g_copyStream = nullptr; }
  if (!g_copyStream) MG_CUDA(cudaStreamCreateWithFlags(&g_copyStream, cudaStreamNonBlocking));
  if (g_readStream && g_device != device) { cudaStreamDestroy(g_readStream); g_readStream = nullptr; }
  if (!g_readStream) MG_CUDA(cudaStreamCreateWithFlags(&g_readStream, cudaStreamNonBlocking));
  cudaDeviceProp prop;
  MG_CUDA(cudaGetDeviceProperties(&prop, device));
  g_sms = prop.multiProcessorCount;
  g_device = device;
  return 0;
}

const char* mg_last_error(void) { return g_error.c_str(); }
int mg_version(void) { return 100; }
int mg_synchronize(void) {
  if (g_device < 0) MG_FAIL("mg_synchronize: mg_init has not been called");
  MG_TRY(mg_halo_wait_pending());
  MG_CUDA(cudaStreamSynchronize(g_stream));
  return mg_p2p_check_all();
}
long long mg_kernel_launch_count(void) { return g_launches.load(); }
void* mg_stream_handle(void) { return (void*)g_stream; }
int mg_profile_enable(int enable) {
  for (auto& p : g_prof) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
  g_prof.clear();
  g_profOn = enable != 0;
  return 0;
}
int mg_tuning_set(const char* name, int value) {
  if (!name) MG_FAIL("mg_tuning_set: null name");
  std::lock_guard<std::mutex> lock(g_tuningMutex);
  g_tuning[name] = value;
  return 0;
}
int mg_tuning_clear(void) {
  std::lock_guard<std::mutex> lock(g_tuningMutex);
  g_tuning.clear();
  return 0;
}
int mg_profile_get(const char* name, double* ms, long long* launches) {
  if (!name || !ms || !launches) MG_FAIL("mg_profile_get: null argument");
  if (g_device < 0) MG_FAIL("mg_profile_get: mg_init has not been called");
  MG_CUDA(cudaStreamSynchronize(g_stream));
  double total = 0.0;
  long long n = 0;
  for (auto& p : g_prof)
    if (p.name == name) {
      float t = 0.f;
      MG_CUDA(cudaEventElapsedTime(&t, p.e0, p.e1));
      total += t;
      ++n;
    }
  *ms = total;
  *launches = n;
  return 0;
}

// ----------------------------------------------------------------------------- stencil
int mg_stencil_create(const char* scheme, mg_stencil** out) { return mg_stencil_create_impl(scheme, out); }
int mg_stencil_update(mg_stencil* s, int direction, const int procDims[3], const int procCoords[3],
                      const int periodic[3], int overlap) {
  return mg_stencil_update_impl(s, direction, procDims, procCoords, periodic, overlap);
}
int mg_stencil_get_adjoint(const mg_stencil* s, mg_stencil** out) { return mg_stencil_get_adjoint_impl(s, out); }
int mg_stencil_destroy(mg_stencil* s) {
  if (!s) return 0;
  if (s->d_op) cudaFree(s->d_op);
  delete s;
  return 0;
}
int mg_stencil_info(const mg_stencil* s, int info[12]) {
  if (!s) MG_FAIL("mg_stencil_info: null handle");
  const MgDevOp& o = s->op;
  const int v[12] = {o.symmetryType, o.interiorWidth, o.boundaryWidth, o.boundaryDepth, o.nGhost[0], o.nGhost[1],
                     o.periodicOffset[0], o.periodicOffset[1], o.hasDomainBoundary[0], o.hasDomainBoundary[1],
                     o.lo, o.nInterior};
  std::memcpy(info, v, sizeof(v));
  return 0;
}
int mg_stencil_coefficients(const mg_stencil* s, double* rhsInterior, double* b1, double* b2, double* norm) {
  if (!s) MG_FAIL("mg_stencil_coefficients: null handle");
  const MgDevOp& o = s->op;
  if (rhsInterior) for (int k = 0; k < o.nInterior; ++k) rhsInterior[k] = o.interior[k];
  for (int m = 0; m < o.boundaryDepth; ++m)
    for (int i = 0; i < o.boundaryWidth; ++i) {
      if (b1) b1[i + o.boundaryWidth * m] = o.b1[m][i];
      if (b2) b2[i + o.boundaryWidth * m] = o.b2[m][i];
    }
  if (norm) for (int m = 0; m < o.normDepth; ++m) norm[m] = o.normBoundary[m];
  return 0;
}

static int stencil_apply_common(mg_stencil* s, double* x, int nComp, const int n[3], int interiorOnly,
                                const double* ghostPrev, const double* ghostNext) {
  if (!s || !x) MG_FAIL("mg_stencil_apply: null argument");
  if (nComp <= 0 || n[0] <= 0 || n[1] <= 0 || n[2] <= 0) MG_FAIL("mg_stencil_apply: invalid sizes");
  const size_t N = (size_t)n[0] * n[1] * n[2];
  const size_t plane = N / n[s->direction - 1];
  // explicit ghost buffers may be host arrays
  double *dPrev = nullptr, *dNext = nullptr;
  bool freePrev = false, freeNext = false;
  auto stage = [&](const double* h, int g, double** d, bool* owned) -> int {
    if (!h || g <= 0) return 0;
    if (is_device_pointer(h)) { *d = const_cast<double*>(h); return 0; }
    const size_t bytes = sizeof(double) * (size_t)g * plane * nComp;
    MG_CUDA(cudaMalloc(d, bytes));
    MG_CUDA(cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, g_stream));
    *owned = true;
    return 0;
  };
  if (g_device < 0) MG_FAIL("libmagudi_gpu: mg_init has not been called (no CUDA device selected)");
  MG_TRY(stage(ghostPrev, s->op.nGhost[0], &dPrev, &freePrev));
  MG_TRY(stage(ghostNext, s->op.nGhost[1], &dNext, &freeNext));
  int rc = with_device_array(x, N * nComp, true, [&](double* din, double* dout) -> int {
    MG_CUDA(cudaMemcpyAsync(dout, din, N * nComp * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    ApplyArgs a;
    a.in = din;
    a.out = dout;
    a.inCompStride = a.outCompStride = N;
    a.nComp = nComp;
    for (int i = 0; i < 3; ++i) a.n[i] = n[i];
    a.interiorOnly = interiorOnly;
    a.ghostPrev = dPrev;
    a.ghostNext = dNext;
    return mg_apply_launch(s, a);
  });
  if (freePrev) cudaFree(dPrev);
  if (freeNext) cudaFree(dNext);
  return rc;
}

int mg_stencil_apply(mg_stencil* s, double* x, int nComp, const int gridSize[3]) {
  if (s && s->procDim > 1) MG_FAIL("mg_stencil_apply: operator spans several ranks; use mg_stencil_apply_ghosted");
  return stencil_apply_common(s, x, nComp, gridSize, 0, nullptr, nullptr);
}
int mg_stencil_apply_ghosted(mg_stencil* s, double* x, int nComp, const int gridSize[3], const double* ghostPrev,
                             const double* ghostNext) {
  if (s && ((s->op.nGhost[0] > 0 && !ghostPrev) || (s->op.nGhost[1] > 0 && !ghostNext)))
    MG_FAIL("mg_stencil_apply_ghosted: missing ghost buffer");
  return stencil_apply_common(s, x, nComp, gridSize, 0, ghostPrev, ghostNext);
}
int mg_stencil_apply_interior(mg_stencil* s, double* x, int nComp, const int gridSize[3]) {
  return stencil_apply_common(s, x, nComp, gridSize, 1, nullptr, nullptr);
}
static int norm_common(mg_stencil* s, double* x, int nComp, const int n[3], int inverse) {
  if (!s || !x) MG_FAIL("mg_stencil_apply_norm: null argument");
  const size_t N = (size_t)n[0] * n[1] * n[2];
  return with_device_array(x, N * nComp, false, [&](double* din, double*) -> int {
    return mg_norm_launch(s, din, N, nComp, n, inverse, g_stream);
  });
}
int mg_stencil_apply_norm(mg_stencil* s, double* x, int nComp, const int n[3]) { return norm_common(s, x, nComp, n, 0); }
int mg_stencil_apply_norm_inverse(mg_stencil* s, double* x, int nComp, const int n[3]) { return norm_common(s, x, nComp, n, 1); }
static int boundary_common(mg_stencil* s, double* x, int nComp, const int n[3], int face, int applyThenProject) {
  if (!s || !x) MG_FAIL("mg_stencil boundary apply: null argument");
  if (face == 0) MG_FAIL("mg_stencil boundary apply: faceOrientation must be non-zero");
  const size_t N = (size_t)n[0] * n[1] * n[2];
  return with_device_array(x, N * nComp, true, [&](double* din, double* dout) -> int {
    return mg_boundary_launch(s, din, dout, N, nComp, n, face, applyThenProject, g_stream);
  });
}
int mg_stencil_apply_and_project_on_boundary(mg_stencil* s, double* x, int nComp, const int n[3], int face) {
  return boundary_common(s, x, nComp, n, face, 1);
}
int mg_stencil_project_on_boundary_and_apply(mg_stencil* s, double* x, int nComp, const int n[3], int face) {
  return boundary_common(s, x, nComp, n, face, 0);
}

// -------------------------------------------------------------------------------- grid
int mg_grid_create(int index, int nD, const int globalSize[3], const int localSize[3], const int offset[3],
                   const int periodicityType[3], const double periodicLength[3], int isCurvilinear,
                   const int procDims[3], const int procCoords[3], mg_grid** out) {
  if (g_device < 0) MG_FAIL("libmagudi_gpu: mg_init has not been called (no CUDA device selected)");
  return mg_grid_create_impl(index, nD, globalSize, localSize, offset, periodicityType, periodicLength,
                             isCurvilinear, procDims, procCoords, out);
}
int mg_grid_destroy(mg_grid* g) { mg_grid_destroy_impl(g); return 0; }
int mg_grid_setup_spatial_discretization(mg_grid* g, const char* s1, const char* s2, const char* s3,
                                         int dissipationOn, int compositeDissipation, int useContinuousAdjoint) {
  if (!g) MG_FAIL("mg_grid_setup_spatial_discretization: null handle");
  const char* sch[3] = {s1 ? s1 : "SBP 4-8", s2 ? s2 : (s1 ? s1 : "SBP 4-8"), s3 ? s3 : (s1 ? s1 : "SBP 4-8")};
  return mg_grid_setup_discretization_impl(g, sch, dissipationOn, compositeDissipation, useContinuousAdjoint);
}

static MgField* grid_field(mg_grid* g, int field, int* nComp, bool alloc) {
  MgField* f = nullptr;
  int nc = 0;
  switch (field) {
    case MG_G_COORDINATES: f = &g->coordinates; nc = g->nD; break;
    case MG_G_METRICS: f = &g->metrics; nc = g->nD * g->nD; break;
    case MG_G_JACOBIAN: f = &g->jacobian; nc = 1; break;
    case MG_G_NORM: f = &g->norm; nc = 1; break;
    case MG_G_ARC_LENGTHS: f = &g->arcLengths; nc = g->nD; break;
    case MG_G_TARGET_MOLLIFIER: f = &g->targetMollifier; nc = 1; break;
    case MG_G_CONTROL_MOLLIFIER: f = &g->controlMollifier; nc = 1; break;
    default: return nullptr;
  }
  if (!f->p && alloc) { if (mg_field_alloc(g, nc, f) != 0) return nullptr; }
  *nComp = nc;
  return f->p ? f : nullptr;
}

int mg_grid_set(mg_grid* g, int field, const double* host) {
  int nc = 0;
  MgField* f = g ? grid_field(g, field, &nc, true) : nullptr;
  if (!f) MG_FAIL("mg_grid_set: unknown field or null handle");
  if (field == MG_G_COORDINATES) g->updated = false;
  return mg_field_upload(g, f, host);
}
int mg_grid_get(mg_grid* g, int field, double* host) {
  int nc = 0;
  MgField* f = g ? grid_field(g, field, &nc, false) : nullptr;
  if (!f) MG_FAIL("mg_grid_get: unknown or unset field");
  return mg_field_download(g, f, host);
}
int mg_grid_set_iblank(mg_grid* g, const int* iblank) {
  if (!g) MG_FAIL("mg_grid_set_iblank: null handle");
  bool holes = false;
  for (size_t p = 0; p < g->N; ++p) holes = holes || iblank[p] == 0;
  if (!holes) { if (g->iblank) { cudaFree(g->iblank); g->iblank = nullptr; } return 0; }
  if (!g->iblank) MG_CUDA(cudaMalloc(&g->iblank, g->N * sizeof(int)));
  MG_CUDA(cudaMemcpy(g->iblank, iblank, g->N * sizeof(int), cudaMemcpyHostToDevice));
  g->updated = false;
  return 0;
}
int mg_grid_update(mg_grid* g, int* hasNegativeJacobian) {
  if (!g) MG_FAIL("mg_grid_update: null handle");
  return mg_grid_update_impl(g, hasNegativeJacobian);
}
int mg_grid_gradient(mg_grid* g, const double* f, int nComp, double* gradF) {
  if (!g || !f || !gradF) MG_FAIL("mg_grid_gradient: null argument");
  if (!g->updated) MG_FAIL("mg_grid_gradient: grid metrics have not been computed");
  MgField in, out, scratch;
  MG_TRY(mg_field_alloc(g, nComp, &in));
  MG_TRY(mg_field_alloc(g, g->nD * nComp, &out));
  MG_TRY(mg_field_alloc(g, g->nD * nComp, &scratch));
  int rc = mg_field_upload(g, &in, f);
  if (rc == 0) rc = mg_grid_gradient_dev(g, in.comp(0), in.compStride, nComp, &out, &scratch);
  if (rc == 0) rc = mg_field_download(g, &out, gradF);
  mg_field_free(&in);
  mg_field_free(&out);
  mg_field_free(&scratch);
  return rc;
}
int mg_grid_inner_product(mg_grid* g, const double* f, const double* gv, const double* weight, int nComp,
                          double* result) {
  if (!g || !f || !gv || !result) MG_FAIL("mg_grid_inner_product: null argument");
  MgField a, b, w;
  MG_TRY(mg_field_alloc(g, nComp, &a));
  MG_TRY(mg_field_alloc(g, nComp, &b));
  int rc = mg_field_upload(g, &a, f);
  if (rc == 0) rc = mg_field_upload(g, &b, gv);
  if (rc == 0 && weight) {
    rc = mg_field_alloc(g, 1, &w);
    if (rc == 0) rc = mg_field_upload(g, &w, weight);
  }
  if (rc == 0)
    rc = mg_grid_inner_product_dev(g, a.comp(0), b.comp(0), weight ? w.comp(0) : nullptr, a.compStride, nComp, result);
  mg_field_free(&a);
  mg_field_free(&b);
  mg_field_free(&w);
  return rc;
}
mg_stencil* mg_grid_operator(mg_grid* g, int which, int direction) {
  if (!g || direction < 1 || direction > 3) return nullptr;
  switch (which) {
    case 0: return g->firstDerivative[direction - 1];
    case 1: return g->adjointFirstDerivative[direction - 1];
    case 2: return g->dissipation[direction - 1];
    case 3: return g->dissipationTranspose[direction - 1];
  }
  return nullptr;
}

// ------------------------------------------------------------------------------- state
static MgField* state_field(mg_state* s, int field) {
  switch (field) {
    case MG_Q_CONSERVED: return &s->Q[s->cur];
    case MG_Q_ADJOINT: return &s->W[s->curW];
    case MG_Q_TARGET: return &s->target;
    case MG_Q_RHS: return &s->rhs;
    case MG_Q_SPECIFIC_VOLUME: return &s->specificVolume;
    case MG_Q_VELOCITY: return &s->velocity;
    case MG_Q_PRESSURE: return &s->pressure;
    case MG_Q_TEMPERATURE: return &s->temperature;
    case MG_Q_DYNAMIC_VISCOSITY: return &s->mu;
    case MG_Q_SECOND_VISCOSITY: return &s->lambda;
    case MG_Q_THERMAL_DIFFUSIVITY: return &s->kappa;
    case MG_Q_STRESS_TENSOR: return &s->stressTensor;
    case MG_Q_HEAT_FLUX: return &s->heatFlux;
    case MG_Q_MEAN_PRESSURE:
      if (!s->meanPressure.p && mg_field_alloc(s->grid, 1, &s->meanPressure) != 0) return nullptr;
      return &s->meanPressure;
    case MG_Q_MEAN_VELOCITY:
      if (!s->meanVelocity.p && mg_field_alloc(s->grid, s->nD, &s->meanVelocity) != 0) return nullptr;
      return &s->meanVelocity;
    case MG_Q_FUSED_TAUQ: return &s->tauq;
    case MG_Q_FUSED_DISSIPATION:
      if (s->fusedValid && !s->dissValid) mg_fused_dissipation(s);
      return &s->dissTerm;
    case MG_Q_FUSED_ADJOINT_DIFFUSION3: {
      static thread_local MgField view;
      view = s->grid->scratchA;
      view.owned = false;
      if (!view.p || s->nD != 3) return nullptr;
      view.p += (size_t)(2 * (s->nU - 1)) * view.compStride;
      view.nComp = s->nU - 1;
      return &view;
    }
  }
  return nullptr;
}

int mg_state_create(mg_grid* g, const mg_options* o, mg_state** out) {
  if (!g || !o || !out) MG_FAIL("mg_state_create: null argument");
  mg_options_t t;
  t.ratioOfSpecificHeats = o->ratioOfSpecificHeats;
  t.viscosityOn = o->viscosityOn;
  t.reynoldsNumberInverse = o->reynoldsNumberInverse;
  t.prandtlNumberInverse = o->prandtlNumberInverse;
  t.powerLawExponent = o->powerLawExponent;
  t.bulkViscosityRatio = o->bulkViscosityRatio;
  t.dissipationOn = o->dissipationOn;
  t.compositeDissipation = o->compositeDissipation;
  t.dissipationAmount = o->dissipationAmount;
  t.useTargetState = o->useTargetState;
  t.useContinuousAdjoint = o->useContinuousAdjoint;
  t.steadyStateSimulation = o->steadyStateSimulation;
  if (!(t.ratioOfSpecificHeats > 1.0)) MG_FAIL("mg_state_create: ratio of specific heats must exceed 1");
  if (t.viscosityOn && !(t.reynoldsNumberInverse > 0.0)) MG_FAIL("mg_state_create: viscous terms need a positive Reynolds number");
  if (t.dissipationOn != g->dissipationOn || (t.dissipationOn && t.compositeDissipation != g->compositeDissipation))
    MG_FAIL("mg_state_create: dissipation flags differ from the grid's spatial discretization");
  return mg_state_create_impl(g, &t, out);
}
int mg_state_destroy(mg_state* s) {
  if (!s) return 0;
  for (mg_patch* p : s->patches) mg_patch_destroy_impl(p);
  for (MgField& f : s->checkpoints) mg_field_free(&f);
  mg_state_destroy_impl(s);
  return 0;
}
int mg_state_set(mg_state* s, int field, const double* host) {
  MgField* f = s ? state_field(s, field) : nullptr;
  if (!f || !f->p) MG_FAIL("mg_state_set: unknown or unallocated field");
  if (field == MG_Q_CONSERVED) { s->dependentValid = false; s->fusedValid = false; s->dissValid = false; }
  if (field == MG_Q_CONSERVED || field == MG_Q_ADJOINT) MG_TRY(mg_state_make_exclusive(s, f, false));
  if (field == MG_Q_TARGET)
    for (mg_patch* p : s->patches) p->AplusReady = false;
  return mg_field_upload(s->grid, f, host);
}
// The fused sweeps keep the dependent variables in registers / in the compact tau-q field: the reference-layout
// arrays (specific volume ... heat flux) are materialised on demand, before anything reads them.
static int refresh_dependents(mg_state* s, int field) {
  if (field >= MG_Q_SPECIFIC_VOLUME && field <= MG_Q_HEAT_FLUX && !s->dependentValid)
    MG_TRY(mg_state_ensure_dependents(s));
  return 0;
}
int mg_state_get(mg_state* s, int field, double* host) {
  MgField* f = s ? state_field(s, field) : nullptr;
  if (!f || !f->p) MG_FAIL("mg_state_get: unknown or unallocated field");
  MG_TRY(refresh_dependents(s, field));
  MG_TRY(mg_field_download(s->grid, f, host));
  return mg_p2p_check_all();
}
// computeCfl / computeTimeStepSize (reference src/StateImpl.f90:548-600 -> src/CNSHelperImpl.f90:842-982)
int mg_state_cfl(mg_state* s, double timeStepSize, double* cfl) {
  if (!s || !cfl) MG_FAIL("mg_state_cfl: null argument");
  if (!(timeStepSize > 0.0)) MG_FAIL("mg_state_cfl: the time step size must be positive");
  return mg_state_cfl_dt_impl(s, 0, timeStepSize, cfl);
}
int mg_state_dt(mg_state* s, double cfl, double* timeStepSize) {
  if (!s || !timeStepSize) MG_FAIL("mg_state_dt: null argument");
  if (!(cfl > 0.0)) MG_FAIL("mg_state_dt: the CFL number must be positive");
  return mg_state_cfl_dt_impl(s, 1, cfl, timeStepSize);
}
// ---- transfers on the copy stream, overlapping the sweeps ------------------------------------------------
namespace {
// make stream `waiter` wait for everything enqueued so far on stream `on`
int stream_wait(cudaStream_t waiter, cudaStream_t on) {
  cudaEvent_t e;
  MG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  MG_CUDA(cudaEventRecord(e, on));
  MG_CUDA(cudaStreamWaitEvent(waiter, e, 0));
  MG_CUDA(cudaEventDestroy(e));     // released once the event has completed
  return 0;
}
int copy_field_async(const mg_grid* g, MgField* f, double* host, bool toDevice) {
  cudaStream_t cs = toDevice ? g_copyStream : g_readStream;
  MG_TRY(stream_wait(cs, g_stream));
  // (the two directions never meet on one buffer: a buffer with a read in flight is not handed out for writing,
  // mg_state_make_exclusive / mg_state_pool_acquire, and a field that is being written is read only after
  // mg_transfer_fence, which orders it through the compute stream)
  for (int c = 0; c < f->nComp; ++c) {
    if (toDevice)
      MG_CUDA(cudaMemcpyAsync(f->comp(c), host + (size_t)c * g->N, g->N * sizeof(double), cudaMemcpyDefault, cs));
    else
      MG_CUDA(cudaMemcpyAsync(host + (size_t)c * g->N, f->comp(c), g->N * sizeof(double), cudaMemcpyDefault, cs));
  }
  return 0;
}
}  // namespace

int mg_state_set_async(mg_state* s, int field, const double* pinnedHost) {
  MgField* f = s ? state_field(s, field) : nullptr;
  if (!f || !f->p || !pinnedHost) MG_FAIL("mg_state_set_async: unknown or unallocated field");
  if (field == MG_Q_CONSERVED) { s->dependentValid = false; s->fusedValid = false; s->dissValid = false; }
  if (field == MG_Q_CONSERVED || field == MG_Q_ADJOINT) MG_TRY(mg_state_make_exclusive(s, f, false));
  if (field == MG_Q_TARGET)
    for (mg_patch* p : s->patches) p->AplusReady = false;
  return copy_field_async(s->grid, f, const_cast<double*>(pinnedHost), true);
}
int mg_state_get_async(mg_state* s, int field, double* pinnedHost) {
  MgField* f = s ? state_field(s, field) : nullptr;
  if (!f || !f->p || !pinnedHost) MG_FAIL("mg_state_get_async: unknown or unallocated field");
  MG_TRY(copy_field_async(s->grid, f, pinnedHost, false));
  if (field == MG_Q_CONSERVED || field == MG_Q_ADJOINT) mg_state_note_pending_read(s, f->p, g_readStream);
  return 0;
}
int mg_state_checkpoint_get_async(mg_state* s, int slot, double* pinnedHost) {
  if (!s || !pinnedHost || slot < 0 || (size_t)slot >= s->checkpoints.size() || !s->checkpoints[slot].p)
    MG_FAIL("mg_state_checkpoint_get_async: empty slot");
  MG_TRY(copy_field_async(s->grid, &s->checkpoints[slot], pinnedHost, false));
  mg_state_note_pending_read(s, s->checkpoints[slot].p, g_readStream);
  return 0;
}
// Double-buffered inputs: copy the NEXT step's conserved / adjoint variables into a free buffer of the state's pool
// while the current step computes (nothing on the compute stream reads or writes that buffer), then adopt it.
int mg_state_stage_async(mg_state* s, int field, const double* pinnedHost) {
  if (!s || !pinnedHost || (field != MG_Q_CONSERVED && field != MG_Q_ADJOINT))
    MG_FAIL("mg_state_stage_async: only the conserved and the adjoint variables can be staged");
  const int w = field == MG_Q_ADJOINT ? 1 : 0;
  MgField* cur = state_field(s, field);
  if (!cur || !cur->p) MG_FAIL("mg_state_stage_async: unallocated field");
  MgField& st = s->staged[w];
  if (!st.p) {
    st = *cur;
    st.owned = false;
    st.p = nullptr;
    st.p = mg_state_pool_acquire(s, cur->compStride * (size_t)cur->nComp * sizeof(double));
    if (!st.p) MG_FAIL("mg_state_stage_async: out of device memory");
  }
  // the buffer is unreferenced, but sweeps enqueued earlier may still be reading it under its previous role
  MG_TRY(stream_wait(g_copyStream, g_stream));
  const mg_grid* g = s->grid;
  for (int c = 0; c < st.nComp; ++c)
    MG_CUDA(cudaMemcpyAsync(st.comp(c), pinnedHost + (size_t)c * g->N, g->N * sizeof(double), cudaMemcpyDefault, g_copyStream));
  if (!s->stagedReady[w]) MG_CUDA(cudaEventCreateWithFlags(&s->stagedReady[w], cudaEventDisableTiming));
  MG_CUDA(cudaEventRecord(s->stagedReady[w], g_copyStream));
  return 0;
}
int mg_state_adopt_staged(mg_state* s, int field) {
  if (!s || (field != MG_Q_CONSERVED && field != MG_Q_ADJOINT)) MG_FAIL("mg_state_adopt_staged: invalid argument");
  const int w = field == MG_Q_ADJOINT ? 1 : 0;
  if (!s->staged[w].p) MG_FAIL("mg_state_adopt_staged: nothing has been staged for this field");
  MG_CUDA(cudaStreamWaitEvent(g_stream, s->stagedReady[w], 0));
  MgField* cur = state_field(s, field);
  cur->p = s->staged[w].p;          // the previous buffer goes back to the pool (unless a slot still views it)
  s->staged[w].p = nullptr;
  if (field == MG_Q_CONSERVED) { s->dependentValid = false; s->fusedValid = false; s->dissValid = false; }
  return 0;
}
int mg_transfer_fence(void) {
  if (g_device < 0) MG_FAIL("mg_transfer_fence: mg_init has not been called");
  MG_TRY(stream_wait(g_stream, g_copyStream));
  return stream_wait(g_stream, g_readStream);
}
int mg_transfer_wait(void) {
  if (g_device < 0) MG_FAIL("mg_transfer_wait: mg_init has not been called");
  MG_CUDA(cudaStreamSynchronize(g_copyStream));
  MG_CUDA(cudaStreamSynchronize(g_readStream));
  return 0;
}

int mg_state_set_time(mg_state* s, double time) {
  if (!s) MG_FAIL("mg_state_set_time: null handle");
  s->time = time;
  return 0;
}
int mg_state_add_acoustic_source(mg_state* s, const double location[3], double amplitude, double frequency,
                                 double radius, double phase) {
  if (!s) MG_FAIL("mg_state_add_acoustic_source: null handle");
  mg_state::Source src;
  for (int i = 0; i < 3; ++i) src.loc[i] = location[i];
  src.amplitude = amplitude;
  src.angularFrequency = 2.0 * (4.0 * atan(1.0)) * frequency;
  src.gaussianFactor = 9.0 / (2.0 * radius * radius);
  src.phase = phase;
  s->acousticSources.push_back(src);
  return 0;
}
int mg_state_update(mg_state* s) {
  if (!s) MG_FAIL("mg_state_update: null handle");
  if (!s->grid->updated) MG_FAIL("mg_state_update: grid metrics have not been computed");
  if (mg_state_uses_fused_rhs(s, MG_FORWARD)) return mg_fused_sweepA(s);
  return mg_state_update_impl(s, nullptr);
}

int mg_state_checkpoint_store(mg_state* s, int slot) {
  if (!s || slot < 0) MG_FAIL("mg_state_checkpoint_store: invalid argument");
  if ((size_t)slot >= s->checkpoints.size()) s->checkpoints.resize(slot + 1);
  // no copy: the slot becomes a view of the buffer holding Q; that buffer is never written again while the
  // slot refers to it (mg_state_make_exclusive swaps in a free buffer first)
  s->checkpoints[slot] = s->Q[s->cur];
  s->checkpoints[slot].owned = false;
  return 0;
}
int mg_state_checkpoint_load(mg_state* s, int slot) {
  if (!s || slot < 0 || (size_t)slot >= s->checkpoints.size() || !s->checkpoints[slot].p)
    MG_FAIL("mg_state_checkpoint_load: empty slot");
  s->Q[s->cur] = s->checkpoints[slot];
  s->dependentValid = false;
  s->fusedValid = false;
  s->dissValid = false;
  return 0;
}
int mg_state_checkpoint_clear(mg_state* s) {
  if (!s) MG_FAIL("mg_state_checkpoint_clear: null handle");
  s->checkpoints.clear();
  // The buffers stay in the state's pool: the next checkpoint window reuses them (cudaFree / cudaMalloc of a window's
  // worth of buffers costs more than the window's march on small grids).  MG_POOL_TRIM=1 gives the memory back.
  if (mg_tuning_get("MG_POOL_TRIM", 0)) mg_state_pool_trim(s);
  return 0;
}
// ------------------------------------------------------------------------------- patch
int mg_patch_create(mg_state* s, int type, const char* name, int normalDirection, const int extent[6],
                    double inviscidPenaltyAmount, double viscousPenaltyAmount, mg_patch** out) {
  if (!s || !out) MG_FAIL("mg_patch_create: null argument");
  MG_TRY(mg_patch_create_impl(s, type, name, normalDirection, extent, out));
  mg_patch* p = *out;
  if (type == MG_PATCH_SPONGE || type == MG_PATCH_JET_EXCITATION) {     // the two amounts carry sponge_amount (jet
    // excitation: patches/<name>/amplitude, src/JetExcitationPatchImpl.f90:47-49) and sponge_exponent
    p->spongeAmount = inviscidPenaltyAmount;
    p->spongeExponent = (int)std::lround(viscousPenaltyAmount);
  }
  const int ad = std::abs(normalDirection);
  if (ad >= 1 && ad <= s->nD && s->grid->firstDerivative[ad - 1]) {
    const double h = s->grid->firstDerivative[ad - 1]->op.normBoundary[0];
    const double sgn = normalDirection > 0 ? 1.0 : -1.0;
    p->inviscidPenaltyAmount = sgn * std::fabs(inviscidPenaltyAmount) / h;
    if (type == MG_PATCH_ISOTHERMAL_WALL)   // src/IsothermalWallImpl.f90:61-72: unsigned, times 1/Re
      p->viscousPenaltyAmount = s->opt.viscosityOn ? viscousPenaltyAmount / h * s->opt.reynoldsNumberInverse : 0.0;
    else
      p->viscousPenaltyAmount = s->opt.viscosityOn ? sgn * std::fabs(viscousPenaltyAmount) / h : 0.0;
  }
  return 0;
}
int mg_patch_num_points(const mg_patch* p, int* n, int localSize[3], int patchOffset[3]) {
  if (!p) MG_FAIL("mg_patch_num_points: null handle");
  if (n) *n = p->nPatchPoints;
  for (int i = 0; i < 3; ++i) {
    if (localSize) localSize[i] = p->localSize[i];
    if (patchOffset) patchOffset[i] = p->patchOffset[i];
  }
  return 0;
}
int mg_patch_set_array(mg_patch* p, const char* name, int nComp, const double* host) {
  if (!p || !name || !host) MG_FAIL("mg_patch_set_array: null argument");
  return mg_patch_set_array_impl(p, name, nComp, host);
}
int mg_patch_get_array(mg_patch* p, const char* name, int nComp, double* host) {
  if (!p || !name || !host) MG_FAIL("mg_patch_get_array: null argument");
  return mg_patch_get_array_impl(p, name, nComp, host);
}
int mg_patch_collect(mg_patch* p, int field, const char* name) {
  if (!p || !name) MG_FAIL("mg_patch_collect: null argument");
  if (p && p->state) MG_TRY(refresh_dependents(p->state, field));
  MgField* f = state_field(p->state, field);
  int nc = 0;
  if (!f) f = grid_field(p->state->grid, field, &nc, false);
  if (!f || !f->p) MG_FAIL("mg_patch_collect: unknown or unallocated field");
  return mg_patch_collect_impl(p, f, f->nComp, name);
}

// ------------------------------------------------------------------------------ functionals
int mg_functional_quadrature_on_patches(mg_state* s, int patchType, const double* integrand, double* value) {
  if (!s || !integrand || !value) MG_FAIL("mg_functional_quadrature_on_patches: null argument");
  // the integrand may live on the host: stage it on the device
  cudaPointerAttributes at;
  const bool onDevice = cudaPointerGetAttributes(&at, integrand) == cudaSuccess && at.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  if (onDevice) return mg_functional_quadrature_impl(s, patchType, integrand, value);
  double* d = nullptr;
  MG_CUDA(cudaMalloc(&d, s->grid->N * sizeof(double)));
  cudaMemcpyAsync(d, integrand, s->grid->N * sizeof(double), cudaMemcpyHostToDevice, g_stream);
  const int rc = mg_functional_quadrature_impl(s, patchType, d, value);
  cudaFree(d);
  return rc;
}
int mg_functional_acoustic_noise(mg_state* s, double timeRampFactor, double* value) {
  if (!s || !value) MG_FAIL("mg_functional_acoustic_noise: null argument");
  return mg_functional_acoustic_noise_impl(s, timeRampFactor, value);
}
int mg_functional_acoustic_noise_forcing(mg_state* s, double timeRampFactor) {
  if (!s) MG_FAIL("mg_functional_acoustic_noise_forcing: null handle");
  return mg_functional_acoustic_noise_forcing_impl(s, timeRampFactor);
}
int mg_functional_actuator_sensitivity(mg_state* s, double timeRampFactor, double* value) {
  if (!s || !value) MG_FAIL("mg_functional_actuator_sensitivity: null argument");
  return mg_functional_actuator_sensitivity_impl(s, timeRampFactor, value);
}
int mg_functional_pressure_drag(mg_state* s, const double direction[3], double* value) {
  if (!s || !direction || !value) MG_FAIL("mg_functional_pressure_drag: null argument");
  return mg_functional_pressure_drag_impl(s, direction, value);
}
int mg_functional_pressure_drag_forcing(mg_state* s, const double direction[3]) {
  if (!s || !direction) MG_FAIL("mg_functional_pressure_drag_forcing: null argument");
  return mg_functional_pressure_drag_forcing_impl(s, direction);
}
int mg_functional_actuator_gradient(mg_patch* p, double timeRampFactor, double* hostOut) {
  if (!p || !hostOut) MG_FAIL("mg_functional_actuator_gradient: null argument");
  return mg_functional_actuator_gradient_impl(p, timeRampFactor, hostOut);
}

// ------------------------------------------------------------------------------ region
int mg_region_create(mg_region** out) { *out = new mg_region(); return 0; }
int mg_region_destroy(mg_region* r) { delete r; return 0; }
int mg_region_add_state(mg_region* r, mg_state* s) {
  if (!r || !s) MG_FAIL("mg_region_add_state: null argument");
  r->states.push_back(s);
  s->useFused = r->fused;
  return 0;
}
int mg_region_update_patches(mg_region* r) {
  if (!r) MG_FAIL("mg_region_update_patches: null handle");
  for (mg_state* s : r->states) MG_TRY(mg_patches_update_impl(s));
  return 0;
}
int mg_region_compute_sponge_strengths(mg_region* r) {
  if (!r) MG_FAIL("mg_region_compute_sponge_strengths: null handle");
  for (mg_state* s : r->states) MG_TRY(mg_patches_sponge_strengths_impl(s));
  return 0;
}
int mg_state_sponge_arc_length(mg_state* s, int direction, double* arcLength) {
  if (!s || !arcLength) MG_FAIL("mg_state_sponge_arc_length: null argument");
  if (direction < 1 || direction > s->nD) MG_FAIL("mg_state_sponge_arc_length: direction out of range");
  return mg_patches_sponge_arc_length_impl(s, direction - 1, arcLength);
}
int mg_state_sponge_strengths_gathered(mg_state* s, int direction, const double* arcLengthsAlongDirection) {
  if (!s || !arcLengthsAlongDirection) MG_FAIL("mg_state_sponge_strengths_gathered: null argument");
  if (direction < 1 || direction > s->nD) MG_FAIL("mg_state_sponge_strengths_gathered: direction out of range");
  return mg_patches_sponge_strengths_gathered_impl(s, direction - 1, arcLengthsAlongDirection);
}
int mg_region_set_fused(mg_region* r, int enable) {
  if (!r) MG_FAIL("mg_region_set_fused: null handle");
  r->fused = enable;
  for (mg_state* s : r->states) s->useFused = enable;
  return 0;
}
int mg_region_uses_fused(mg_region* r, int mode) {
  if (!r) return 0;
  if (!r->fused) return 0;
  for (mg_state* s : r->states) if (!mg_fused_supported(s, mode)) return 0;
  return 1;
}
int mg_region_uses_fused_rhs(mg_region* r, int mode) {
  if (!r || !r->fused) return 0;
  for (mg_state* s : r->states) if (!mg_state_uses_fused_rhs(s, mode)) return 0;
  return 1;
}
static bool region_has_interfaces(const mg_region* r) {
  for (const mg_state* s : r->states) if (mg_state_has_interfaces(s)) return true;
  return false;
}
// t_Region%computeRhs (reference src/RegionImpl.f90:1877-2027).  With block interfaces the evaluation is staged
// over all the grids of the region as in the reference: RHS of every grid, interface exchange, then the viscous
// interface adjoint penalty, 1/J, patch penalties and sources of every grid.
// addBodyForce (reference src/RegionImpl.f90:732-851): region integrals on the device (two-stage deterministic
// reductions, one host synchronisation each, as the reference's MPI reductions), pointwise adds in one kernel per state
static int region_add_body_force(mg_region* r, int mode, int stage) {
  mg_region::BodyForce& bf = r->bf;
  auto integral = [&](int which, double* out) -> int {
    *out = 0.0;
    for (mg_state* s : r->states) {
      double v = 0.0;
      MG_TRY(mg_state_integral_impl(s, which, &v));
      *out += v;
    }
    return 0;
  };
  if (stage == 1) {
    double current = 0.0;
    MG_TRY(integral(1, &current));
    bf.momentumLoss = bf.oneOverVolume / bf.dt * (bf.initialXmomentum - current);
    if (mode == MG_LINEARIZED) {
      double dmom = 0.0;
      MG_TRY(integral(2, &dmom));
      bf.adjointMomentumLoss = 0.0 - bf.oneOverVolume / bf.dt * dmom;
    }
  }
  double stage1Term = 0.0;
  if (mode == MG_ADJOINT) {
    const double factor = (stage == 2 || stage == 3) ? 2.0 : 1.0;
    double wmom = 0.0, wx = 0.0;
    MG_TRY(integral(2, &wmom));
    MG_TRY(integral(3, &wx));
    bf.adjointMomentumLoss = bf.adjointMomentumLoss - factor * (wmom + wx);
    if (stage == 1) stage1Term = bf.adjointMomentumLoss * bf.oneOverVolume / bf.dt;
  }
  for (mg_state* s : r->states)
    MG_TRY(mg_state_add_body_force_impl(s, mode, bf.momentumLoss, bf.adjointMomentumLoss, stage == 1, stage1Term));
  if (mode == MG_ADJOINT && stage == 1) bf.adjointMomentumLoss = 0.0;
  return 0;
}

static int region_compute_rhs(mg_region* r, int mode, int stage = 1) {
  if (!region_has_interfaces(r)) {
    for (mg_state* s : r->states) MG_TRY(mg_state_compute_rhs_impl(s, mode));
  } else {
    for (mg_state* s : r->states) MG_TRY(mg_state_rhs_pre(s, mode));
    MG_TRY(mg_interfaces_exchange(r->states, mode));
    for (mg_state* s : r->states) MG_TRY(mg_state_rhs_post(s, mode));
  }
  // the body force joins after the sources; hole points stay zero (src/RegionImpl.f90:2012-2023)
  if (r->bf.enabled) MG_TRY(region_add_body_force(r, mode, stage));
  return 0;
}
int mg_region_set_body_force(mg_region* r, int enable, double initialMomentumPerVolume, double timeStepSize) {
  if (!r) MG_FAIL("mg_region_set_body_force: null handle");
  r->bf = mg_region::BodyForce();
  for (mg_state* s : r->states) s->bodyForce = enable != 0;
  if (!enable) return 0;
  if (!(timeStepSize > 0.0)) MG_FAIL("mg_region_set_body_force: the time step size must be positive");
  for (mg_state* s : r->states) {
    const mg_grid* g = s->grid;
    if (g->procDims[0] * g->procDims[1] * g->procDims[2] != 1)
      MG_FAIL("mg_region_set_body_force: region integrals of a decomposed grid are not reduced over ranks yet");
  }
  double volume = 0.0;
  for (mg_state* s : r->states) {
    double v = 0.0;
    MG_TRY(mg_state_integral_impl(s, 0, &v));
    volume += v;
  }
  // src/SolverImpl.f90:760-765: body_force/initial_momentum is per unit volume
  r->bf.enabled = true;
  r->bf.initialXmomentum = initialMomentumPerVolume * volume;
  r->bf.oneOverVolume = 1.0 / volume;
  r->bf.dt = timeStepSize;
  return 0;
}
int mg_region_get_body_force(mg_region* r, double* momentumLossPerVolume, double* adjointMomentumLossPerVolume) {
  if (!r) MG_FAIL("mg_region_get_body_force: null handle");
  if (momentumLossPerVolume) *momentumLossPerVolume = r->bf.momentumLoss;
  if (adjointMomentumLossPerVolume) *adjointMomentumLossPerVolume = r->bf.adjointMomentumLoss;
  return 0;
}
int mg_region_compute_rhs(mg_region* r, int mode, int timestep, int stage) {
  (void)timestep;
  if (!r) MG_FAIL("mg_region_compute_rhs: null handle");
  return region_compute_rhs(r, mode, stage);
}
int mg_patch_link_interface(mg_patch* a, mg_patch* b, const int indexReorderingA[3]) {
  return mg_interface_link(a, b, indexReorderingA);
}
int mg_patch_penalty_amounts(mg_patch* p, double* inviscid, double* viscous) {
  if (!p || !inviscid || !viscous) MG_FAIL("mg_patch_penalty_amounts: null argument");
  *inviscid = p->inviscidPenaltyAmount;
  *viscous = p->state->opt.viscosityOn ? p->viscousPenaltyAmount : 0.0;
  return 0;
}
int mg_patch_link_interface_remote(mg_patch* p, const int indexReordering[3], double partnerInviscidPenaltyAmount,
                                   double partnerViscousPenaltyAmount, int partnerNormalDirection, mg_p2p** link) {
  return mg_interface_link_remote(p, indexReordering, partnerInviscidPenaltyAmount, partnerViscousPenaltyAmount,
                                  partnerNormalDirection, link);
}
int mg_rk4_substep(mg_region* r, int mode, double* time, double dt, int timestep, int stage, int updateStates) {
  if (!r || !time) MG_FAIL("mg_rk4_substep: null argument");
  double t = *time;
  if (region_has_interfaces(r) || r->bf.enabled) {
    if (stage < 1 || stage > 4) MG_FAIL("rk4 substep: stage must be 1..4");
    for (mg_state* s : r->states) mg_rk4_set_times(s, mode, *time, dt, stage);
    MG_TRY(region_compute_rhs(r, mode, stage));
    for (mg_state* s : r->states) s->rhsReady = true;
  }
  for (mg_state* s : r->states) {
    t = *time;
    MG_TRY(mg_rk4_substep_impl(s, mode, &t, dt, timestep, stage));
    if (updateStates && mode == MG_FORWARD) {
      if (mg_state_uses_fused_rhs(s, MG_FORWARD)) MG_TRY(mg_fused_sweepA(s));
      else MG_TRY(mg_state_update_impl(s, nullptr));
    }
  }
  *time = t;
  return 0;
}

// device-resident time quadratures and control buffers (no host synchronisation per substep)
int mg_functional_accumulate(mg_state* s, int which, double weight, double timeRampFactor) {
  if (!s) MG_FAIL("mg_functional_accumulate: null handle");
  return mg_functional_accumulate_impl(s, which, weight, timeRampFactor);
}
int mg_functional_accumulator_get(mg_state* s, int which, double* value, int reset) {
  if (!s || !value) MG_FAIL("mg_functional_accumulator_get: null argument");
  return mg_functional_accumulator_get_impl(s, which, value, reset);
}
int mg_patch_gradient_buffer_setup(mg_patch* p, int nSlots) {
  if (!p) MG_FAIL("mg_patch_gradient_buffer_setup: null handle");
  return mg_patch_gradient_buffer_setup_impl(p, nSlots);
}
int mg_functional_actuator_gradient_record(mg_patch* p, double timeRampFactor, int* bufferIsFull) {
  if (!p) MG_FAIL("mg_functional_actuator_gradient_record: null handle");
  return mg_functional_actuator_gradient_record_impl(p, timeRampFactor, bufferIsFull);
}
int mg_patch_gradient_buffer_flush(mg_patch* p, double* host, int* nRecords) {
  if (!p) MG_FAIL("mg_patch_gradient_buffer_flush: null handle");
  return mg_patch_gradient_buffer_flush_impl(p, host, nRecords);
}
int mg_patch_control_forcing_from_buffer(mg_patch* p, int slot, int firstComponent, int nComponents) {
  if (!p) MG_FAIL("mg_patch_control_forcing_from_buffer: null handle");
  return mg_patch_control_forcing_from_buffer_impl(p, slot, firstComponent, nComponents);
}
// ------------------------------------------------------------------------------ SURVEY 8 f4
int mg_functional_drag_force(mg_state* s, const double direction[3], double* value) {
  if (!s || !direction || !value) MG_FAIL("mg_functional_drag_force: null argument");
  double d[3] = {direction[0], s->nD >= 2 ? direction[1] : 0.0, s->nD == 3 ? direction[2] : 0.0};
  const double n2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  if (n2 <= 2.220446049250313e-16) MG_FAIL("Unable to determine a unit vector for computing drag force!");
  for (double& v : d) v /= std::sqrt(n2);
  return mg_functional_drag_force_impl(s, d, value);
}
int mg_functional_reynolds_stress(mg_state* s, const double direction1[3], const double direction2[3], double* value) {
  if (!s || !direction1 || !direction2 || !value) MG_FAIL("mg_functional_reynolds_stress: null argument");
  return mg_functional_reynolds_stress_impl(s, direction1, direction2, value);
}
int mg_functional_reynolds_stress_forcing(mg_state* s, const double direction1[3], const double direction2[3]) {
  if (!s || !direction1 || !direction2) MG_FAIL("mg_functional_reynolds_stress_forcing: null argument");
  return mg_functional_reynolds_stress_forcing_impl(s, direction1, direction2);
}
int mg_functional_momentum_actuator_sensitivity(mg_state* s, int direction, double* value) {
  if (!s || !value) MG_FAIL("mg_functional_momentum_actuator_sensitivity: null argument");
  return mg_functional_momentum_actuator_sensitivity_impl(s, direction, value);
}
int mg_functional_momentum_actuator_gradient(mg_patch* p, int direction, double* hostOut) {
  if (!p || !hostOut) MG_FAIL("mg_functional_momentum_actuator_gradient: null argument");
  return mg_functional_momentum_actuator_gradient_impl(p, direction, hostOut);
}
int mg_patch_kolmogorov_setup(mg_patch* p, double amplitude, int wavenumber) {
  if (!p) MG_FAIL("mg_patch_kolmogorov_setup: null handle");
  return mg_patch_kolmogorov_setup_impl(p, amplitude, wavenumber);
}
int mg_patch_set_jet_modes(mg_patch* p, int nModes, const double* angularFrequencies) {
  if (!p || p->type != MG_PATCH_JET_EXCITATION) MG_FAIL("mg_patch_set_jet_modes: not a JET_EXCITATION patch");
  if (nModes < 0 || (nModes > 0 && !angularFrequencies)) MG_FAIL("mg_patch_set_jet_modes: invalid argument");
  nModes = std::min(nModes, MG_JET_MAX_MODES);        // src/JetExcitationPatchImpl.f90:53
  p->angularFrequencies.assign(angularFrequencies, angularFrequencies + nModes);
  return 0;
}
int mg_patch_probe_setup(mg_patch* p, int probeBufferSize) {
  if (!p) MG_FAIL("mg_patch_probe_setup: null handle");
  return mg_patch_probe_setup_impl(p, probeBufferSize);
}
int mg_patch_probe_record(mg_patch* p, int mode, int* bufferIsFull) {
  if (!p) MG_FAIL("mg_patch_probe_record: null handle");
  return mg_patch_probe_record_impl(p, mode, bufferIsFull);
}
int mg_patch_probe_flush(mg_patch* p, double* host, int* nRecords) {
  if (!p) MG_FAIL("mg_patch_probe_flush: null handle");
  return mg_patch_probe_flush_impl(p, host, nRecords);
}
int mg_state_extrema(mg_state* s, int variable, double* vMin, int ijkMin[3], double* vMax, int ijkMax[3]) {
  if (!s) MG_FAIL("mg_state_extrema: null handle");
  return mg_state_extrema_impl(s, variable, vMin, ijkMin, vMax, ijkMax);
}
int mg_state_solution_limit_penalty(mg_state* s, const double densityRange[2], const double temperatureRange[2],
                                    int densityOutOfRange, int temperatureOutOfRange, double* value) {
  if (!s || !densityRange || !temperatureRange || !value) MG_FAIL("mg_state_solution_limit_penalty: null argument");
  return mg_state_limit_penalty_impl(s, densityRange, temperatureRange, densityOutOfRange, temperatureOutOfRange, value);
}
int mg_region_set_solution_limits(mg_region* r, int soft, const double densityRange[2],
                                  const double temperatureRange[2], double penaltyFactor) {
  if (!r) MG_FAIL("mg_region_set_solution_limits: null handle");
  if (soft && (!densityRange || !temperatureRange)) MG_FAIL("mg_region_set_solution_limits: null range");
  for (mg_state* s : r->states) {
    s->limits.soft = soft != 0;
    if (soft) {
      for (int i = 0; i < 2; ++i) { s->limits.densityRange[i] = densityRange[i]; s->limits.temperatureRange[i] = temperatureRange[i]; }
      s->limits.penaltyFactor = penaltyFactor;
    }
  }
  return 0;
}
int mg_region_solution_limit_forcing_switch(mg_region* r, int on) {
  if (!r) MG_FAIL("mg_region_solution_limit_forcing_switch: null handle");
  for (mg_state* s : r->states) s->limits.forcingSwitch = on != 0;
  return 0;
}
int mg_state_set_solution_limit_flags(mg_state* s, int densityOutOfRange, int temperatureOutOfRange) {
  if (!s) MG_FAIL("mg_state_set_solution_limit_flags: null handle");
  s->limits.rhoOut = densityOutOfRange;
  s->limits.tOut = temperatureOutOfRange;
  return 0;
}
int mg_grid_setup_filter(mg_grid* g, const char* filteringScheme) {
  if (!g) MG_FAIL("mg_grid_setup_filter: null handle");
  return mg_grid_setup_filter_impl(g, filteringScheme);
}
int mg_state_apply_filter(mg_state* s, int field, int timestep) {
  if (!s) MG_FAIL("mg_state_apply_filter: null handle");
  if (field != MG_Q_CONSERVED && field != MG_Q_ADJOINT) MG_FAIL("mg_state_apply_filter: field must be the conserved or adjoint variables");
  MgField* f = state_field(s, field);
  if (!f || !f->p) MG_FAIL("mg_state_apply_filter: the field has not been set");
  MG_TRY(mg_state_make_exclusive(s, f, true));
  MG_TRY(mg_grid_apply_filter_impl(s->grid, f, timestep));
  if (field == MG_Q_CONSERVED) { s->dependentValid = false; s->fusedValid = false; }
  return 0;
}
// t_JamesonRK3Integrator%substepForward (reference src/JamesonRK3IntegratorImpl.f90:56-131); the adjoint and
// linearized substeps of the reference are empty.
int mg_rk3_substep(mg_region* r, double* time, double dt, int timestep, int stage, int updateStates) {
  (void)timestep;
  if (!r || !time) MG_FAIL("mg_rk3_substep: null argument");
  if (stage < 1 || stage > 3) MG_FAIL("mg_rk3_substep: stage must be 1..3");
  double t = *time;
  if (region_has_interfaces(r) || r->bf.enabled) {
    for (mg_state* s : r->states) {
      if (stage == 1) s->timeProgressive = *time + dt / 2.0;
      if (stage == 2) { s->time = *time + dt / 2.0; s->timeProgressive = *time + dt; }
      if (stage == 3) s->time = *time + dt / 2.0;
    }
    MG_TRY(region_compute_rhs(r, MG_FORWARD, stage));
    for (mg_state* s : r->states) s->rhsReady = true;
  }
  for (mg_state* s : r->states) {
    t = *time;
    MG_TRY(mg_rk3_substep_impl(s, &t, dt, stage));
    if (updateStates) {
      if (mg_state_uses_fused_rhs(s, MG_FORWARD)) MG_TRY(mg_fused_sweepA(s));
      else MG_TRY(mg_state_update_impl(s, nullptr));
    }
  }
  *time = t;
  return 0;
}

int mg_rk4_substep_adjoint_phase(mg_region* r, int phase, double* time, double dt, int timestep, int stage) {
  (void)timestep;
  if (!r || !time) MG_FAIL("mg_rk4_substep_adjoint_phase: null argument");
  if (stage < 1 || stage > 4) MG_FAIL("mg_rk4_substep_adjoint_phase: stage must be 1..4");
  double t = *time;
  for (mg_state* s : r->states) {
    if (!(s->useFused && mg_fused_supported(s, MG_ADJOINT)))
      MG_FAIL("mg_rk4_substep_adjoint_phase: the fused adjoint path does not cover this configuration");
    t = *time;
    if (phase == 1) {
      const double factor[5] = {0.0, 1.0, 0.5, 1.0, 2.0};
      s->adjointForcingFactor = factor[stage];
      if (stage == 4) s->timeProgressive = t - dt / 2.0;
      MG_TRY(mg_fused_adjoint1(s));
    } else {
      MG_TRY(mg_fused_adjoint2(s, 1, stage, dt));
      if (stage == 4 || stage == 2) s->timeProgressive = t;
      if (stage == 3 || stage == 1) { t -= dt / 2.0; s->time = t; }
    }
  }
  *time = t;
  return 0;
}

MgField* mg_lookup_field(mg_grid* g, void* owner, int field) {
  int nc = 0;
  if (field >= 100) return g ? grid_field(g, field, &nc, false) : nullptr;
  return owner ? state_field((mg_state*)owner, field) : nullptr;
}

int mg_halo_pack(mg_grid* g, void* owner, int field, int side, int width, double* buf) {
  if (!g || !buf) MG_FAIL("mg_halo_pack: null argument");
  int nc = 0;
  MgField* f = field >= 100 ? grid_field(g, field, &nc, false) : state_field((mg_state*)owner, field);
  if (!f || !f->p) MG_FAIL("mg_halo_pack: unknown field");
  if (width > g->gk || width > g->localSize[2]) MG_FAIL("mg_halo_pack: width exceeds ghost capacity");
  const size_t chunk = g->plane * (size_t)width;
  for (int c = 0; c < f->nComp; ++c) {
    const double* src = f->comp(c) + (side == 0 ? 0 : g->plane * (size_t)(g->localSize[2] - width));
    MG_CUDA(cudaMemcpyAsync(buf + chunk * c, src, chunk * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
  }
  return 0;   // asynchronous on the library stream (mg_stream_handle): the exchange is ordered on that stream
}
int mg_halo_unpack(mg_grid* g, void* owner, int field, int side, int width, const double* buf) {
  if (!g || !buf) MG_FAIL("mg_halo_unpack: null argument");
  int nc = 0;
  MgField* f = field >= 100 ? grid_field(g, field, &nc, false) : state_field((mg_state*)owner, field);
  if (!f || !f->p) MG_FAIL("mg_halo_unpack: unknown field");
  if (width > g->gk) MG_FAIL("mg_halo_unpack: width exceeds ghost capacity");
  const size_t chunk = g->plane * (size_t)width;
  for (int c = 0; c < f->nComp; ++c) {
    double* dst = f->comp(c) + (side == 0 ? -(ptrdiff_t)chunk : (ptrdiff_t)(g->plane * (size_t)g->localSize[2]));
    MG_CUDA(cudaMemcpyAsync(dst, buf + chunk * c, chunk * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
  }
  return 0;
}

}  // extern "C"
