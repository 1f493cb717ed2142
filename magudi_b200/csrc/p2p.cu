// Direct peer-to-peer halo exchange over NVLink for the slab (direction-3) decomposition: the replacement of
// fillGhostPoints (reference src/MPIHelperImpl.f90:113-389) on the hot path.
//
// Every rank owns one IPC-shared device allocation holding, per k-face, two staging buffers (double
// buffering) and the flags that sequence them.  An exchange is two kernels on the library stream and no
// host synchronisation at all (push and unpack are one launch each, blockIdx.y = face):
//   push   : read my boundary planes from the field and STORE them straight into the neighbour's staging
//            buffer (remote NVLink stores), then release a sequence flag in the neighbour's memory;
//   unpack : spin until the neighbour's flag for this exchange has arrived (acquire, system scope), copy the
//            staging buffer into my ghost planes, then tell the sender the buffer is free (ack flag).
// The host only enqueues; the GPUs synchronise pairwise through the flags, so ranks may run ahead of each
// other by up to two exchanges per face.
#include <cstdint>
#include <algorithm>
#include <vector>
#include <cstring>

#include "../../include/magudi_gpu.h"
#include "grid.h"
#include "mg_common.h"

extern "C" MgField* mg_lookup_field(mg_grid* g, void* owner, int field);   // c_api.cu

namespace {

constexpr int P2P_BLOCKS = 144, P2P_THREADS = 256;
// A wait that lasts longer than this turns a protocol bug or a dead neighbour into an error instead of a hang.
// Legitimate rank skew (first-use allocations, host I/O, a profiler) can be long: the default is 120 s and
// MG_P2P_TIMEOUT_S (environment / mg_tuning_set) overrides it.  The flag is host-mapped and sticky: it is checked at
// every host synchronisation point of the library (mg_synchronize, mg_state_get, the functionals), so a timed-out
// exchange cannot yield a result silently; the handle is dead afterwards (the use counts no longer agree).
__device__ long long g_spinLimitCycles = 240LL * 1000 * 1000 * 1000;

struct Shared {                         // layout of the flag block at the head of the shared allocation
  unsigned long long data[2][2];        // [my ghost face][parity]: uses of recv[face][parity] completed by the sender
  unsigned long long ack[2][2];         // [my send direction][parity]: uses consumed by that receiver
  unsigned long long pad[8];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// thread 0 of the block spins until *flag >= want; returns false on timeout
__device__ __forceinline__ bool block_wait(const unsigned long long* flag, unsigned long long want, int* error) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = 1;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < want) {
      if (clock64() - t0 > g_spinLimitCycles) {
        ok = 0;
        *(volatile int*)error = 1;
        __threadfence_system();
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  return ok != 0;
}

struct PushArgs {
  const double* src[MG_P2P_MAX_COMP];   // first boundary plane of every component
  int nComp;
  size_t chunk;                         // doubles per component (width * plane)
  double* dst;                          // neighbour's staging buffer
  unsigned long long* dstFlag;          // neighbour's data flag for this buffer
  const unsigned long long* ackFlag;    // my ack flag for this buffer (written by the neighbour)
  unsigned long long use;               // number of earlier uses of this buffer
  unsigned int* counter;                // local block counter
  int* error;
  int vec;                              // every pointer 16-byte aligned and chunk even: double2 copies
};

__device__ __forceinline__ void copy_chunk(double* d, const double* s, size_t n, bool vec, bool viaL2) {
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    const double2* s2 = reinterpret_cast<const double2*>(s);
    double2* d2 = reinterpret_cast<double2*>(d);
    for (size_t i = t0; i < n / 2; i += nt) d2[i] = viaL2 ? __ldcg(s2 + i) : s2[i];
  } else {
    for (size_t i = t0; i < n; i += nt) d[i] = viaL2 ? __ldcg(s + i) : s[i];
  }
}

struct PushPair { PushArgs s[2]; };

__global__ void __launch_bounds__(P2P_THREADS) k_push(const __grid_constant__ PushPair pp) {
  const PushArgs& a = pp.s[blockIdx.y];
  if (!a.dst) return;
  // the previous use of this staging buffer must have been consumed by the neighbour
  const bool ok = block_wait(a.ackFlag, a.use, a.error);
  if (ok)
    for (int c = 0; c < a.nComp; ++c) copy_chunk(a.dst + (size_t)c * a.chunk, a.src[c], a.chunk, a.vec != 0, false);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    // the block counter is left at zero whatever happened (a timed-out launch does not poison the next one)
    const unsigned int done = atomicAdd(a.counter, 1u);
    if (done == gridDim.x - 1) {
      *a.counter = 0;
      __threadfence_system();
      if (*(volatile int*)a.error == 0) st_release_sys(a.dstFlag, a.use + 1);
    }
  }
}

struct UnpackArgs {
  double* dst[MG_P2P_MAX_COMP];         // first ghost plane of every component
  int nComp;
  size_t chunk;
  const double* src;                    // my staging buffer
  const unsigned long long* dataFlag;   // my data flag for this buffer
  unsigned long long* ackFlag;          // the sender's ack flag
  unsigned long long use;
  unsigned int* counter;
  int* error;
  int vec;
};

struct UnpackPair { UnpackArgs s[2]; };

__global__ void __launch_bounds__(P2P_THREADS) k_unpack(const __grid_constant__ UnpackPair pp) {
  const UnpackArgs& a = pp.s[blockIdx.y];
  if (!a.src) return;
  const bool ok = block_wait(a.dataFlag, a.use + 1, a.error);
  // the staging buffer was written by the peer: read it through L2 (bypass L1)
  if (ok)
    for (int c = 0; c < a.nComp; ++c) copy_chunk(a.dst[c], a.src + (size_t)c * a.chunk, a.chunk, a.vec != 0, true);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(a.counter, 1u);
    if (done == gridDim.x - 1) {
      *a.counter = 0;
      __threadfence_system();
      if (*(volatile int*)a.error == 0) st_release_sys(a.ackFlag, a.use + 1);
    }
  }
}

}  // namespace

struct mg_p2p {
  mg_grid* grid = nullptr;
  int direction = 2;                    // 0 / 1: faces normal to i / j travel as packed buffers; 2: contiguous k planes
  double* faceSend[2] = {nullptr, nullptr};   // packed first / last `width` points of every line (directions 0, 1)
  double* faceRecv[2] = {nullptr, nullptr};   // ghost buffers in the layout k_apply reads (reference fillGhostPoints)
  size_t capacity = 0;                  // doubles per staging buffer
  char* base = nullptr;                 // my shared allocation
  char* peer[2] = {nullptr, nullptr};   // mapped allocations of prev (0) / next (1)
  bool peerMapped[2] = {false, false};
  unsigned long long uses[2] = {0, 0};  // exchanges done so far (same count on both faces)
  unsigned int* counters = nullptr;     // [4] local block counters
  int* error = nullptr;                 // device error flag (spin timeout)
  size_t bytes = 0;
  cudaEvent_t evProduced = nullptr, evDone[4] = {nullptr, nullptr, nullptr, nullptr};
  int nextDone = 0;
  Shared* flags() const { return reinterpret_cast<Shared*>(base); }
  static size_t bufOffset(size_t cap, int face, int parity) {
    return sizeof(Shared) + ((size_t)face * 2 + parity) * cap * sizeof(double);
  }
};

// every live handle, so that the library's host synchronisation points can look at the error flags
static std::vector<mg_p2p*> g_handles;
static void mg_p2p_register(mg_p2p* h, bool add) {
  if (add) g_handles.push_back(h);
  else g_handles.erase(std::remove(g_handles.begin(), g_handles.end(), h), g_handles.end());
}
int mg_p2p_check_all() {
  for (mg_p2p* h : g_handles)
    if (h->error && *(volatile int*)h->error)
      MG_FAIL("mg_p2p: a halo exchange timed out waiting for its neighbour: ghost planes are stale (MG_P2P_TIMEOUT_S)");
  return 0;
}

int mg_p2p_create(mg_grid* g, int maxComp, int width, mg_p2p** out) {
  if (!g || !out) MG_FAIL("mg_p2p_create: null argument");
  if (maxComp < 1 || maxComp > MG_P2P_MAX_COMP) MG_FAIL("mg_p2p_create: component count out of range");
  if (width < 1 || width > g->gk) MG_FAIL("mg_p2p_create: width exceeds the ghost capacity");
  mg_p2p* h = new mg_p2p;
  h->grid = g;
  h->capacity = g->plane * (size_t)width * (size_t)maxComp;
  h->capacity += h->capacity % 2;       // keeps every staging buffer 16-byte aligned
  h->bytes = sizeof(Shared) + 4 * h->capacity * sizeof(double);
  MG_CUDA(cudaMalloc(&h->base, h->bytes));
  MG_CUDA(cudaMemset(h->base, 0, h->bytes));
  MG_CUDA(cudaMalloc(&h->counters, 4 * sizeof(unsigned int)));
  MG_CUDA(cudaMemset(h->counters, 0, 4 * sizeof(unsigned int)));
  MG_CUDA(cudaHostAlloc(&h->error, sizeof(int), cudaHostAllocMapped));     // host-mapped: checked without a copy
  *h->error = 0;
  {
    const long long seconds = mg_tuning_get("MG_P2P_TIMEOUT_S", 120);
    int khz = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const long long cycles = seconds * (long long)(khz > 0 ? khz : 2000000) * 1000LL;
    MG_CUDA(cudaMemcpyToSymbol(g_spinLimitCycles, &cycles, sizeof(cycles)));
  }
  mg_p2p_register(h, true);
  MG_CUDA(cudaDeviceSynchronize());
  g->halo = h;          // operator applications along k on this grid fill their ghost planes through it
  *out = h;
  return 0;
}

int mg_p2p_handle_size(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int mg_p2p_get_handle(mg_p2p* h, void* handleOut) {
  if (!h || !handleOut) MG_FAIL("mg_p2p_get_handle: null argument");
  cudaIpcMemHandle_t mh;
  MG_CUDA(cudaIpcGetMemHandle(&mh, h->base));
  std::memcpy(handleOut, &mh, sizeof(mh));
  return 0;
}

// side 0: the previous rank along k, 1: the next one.  `sameAsOther` != 0 when both neighbours are the same
// process (two ranks, periodic): the mapping of the other side is reused.
int mg_p2p_connect(mg_p2p* h, int side, const void* peerHandle, int sameAsOther) {
  if (!h || side < 0 || side > 1) MG_FAIL("mg_p2p_connect: invalid argument");
  if (!peerHandle) { h->peer[side] = nullptr; return 0; }
  if (sameAsOther && h->peer[1 - side]) { h->peer[side] = h->peer[1 - side]; return 0; }
  cudaIpcMemHandle_t mh;
  std::memcpy(&mh, peerHandle, sizeof(mh));
  void* p = nullptr;
  MG_CUDA(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
  h->peer[side] = (char*)p;
  h->peerMapped[side] = true;
  return 0;
}

static int p2p_exchange_on(mg_p2p* h, void* owner, int field, int width, cudaStream_t st);
static int p2p_exchange_field(mg_p2p* h, const MgField* f, int width, cudaStream_t st, unsigned compMask = 0xffffffffu);

// Exchange `width` ghost planes of a field with both k-neighbours.  Asynchronous on the library stream.
int mg_p2p_exchange(mg_p2p* h, void* owner, int field, int width) {
  MG_TRY(mg_halo_wait_pending());
  return p2p_exchange_on(h, owner, field, width, mg_stream());
}

static int p2p_overlapped(mg_p2p* h, void* owner, int field, int width, unsigned compMask);

// The same for a subset of the field's components (bit c of compMask = component c): a consumer that reads only some
// components in the ghost planes -- the k-direction flux of sweep B on a rectilinear grid takes tau_13, tau_23,
// tau_33 and q_3 of the nine tau / q components -- moves only those.
int mg_p2p_exchange_masked(mg_p2p* h, void* owner, int field, int width, unsigned compMask, int overlapped) {
  if (!h) MG_FAIL("mg_p2p_exchange_masked: null handle");
  if (overlapped && mg_halo_stream() != mg_stream()) return p2p_overlapped(h, owner, field, width, compMask);
  MG_TRY(mg_halo_wait_pending());
  MgField* f = mg_lookup_field(h->grid, owner, field);
  if (!f || !f->p) MG_FAIL("mg_p2p_exchange_masked: unknown field");
  return p2p_exchange_field(h, f, width, mg_stream(), compMask);
}

// The same exchange on the library's halo stream, ordered after everything enqueued so far on the main stream:
// it overlaps with the interior k-chunks of the next fused sweep, which takes the completion event
// (fused_common.cuh: launch_split); any other consumer waits for it in mg_synchronize / mg_p2p_exchange.
int mg_p2p_exchange_overlapped(mg_p2p* h, void* owner, int field, int width) {
  if (!h) MG_FAIL("mg_p2p_exchange_overlapped: null handle");
  if (mg_halo_stream() == mg_stream()) return mg_p2p_exchange(h, owner, field, width);
  return p2p_overlapped(h, owner, field, width, 0xffffffffu);
}

static int p2p_overlapped(mg_p2p* h, void* owner, int field, int width, unsigned compMask) {
  cudaStream_t hs = mg_halo_stream();
  if (!h->evProduced) {
    MG_CUDA(cudaEventCreateWithFlags(&h->evProduced, cudaEventDisableTiming));
    for (int i = 0; i < 4; ++i) MG_CUDA(cudaEventCreateWithFlags(&h->evDone[i], cudaEventDisableTiming));
  }
  MG_CUDA(cudaEventRecord(h->evProduced, mg_stream()));
  MG_CUDA(cudaStreamWaitEvent(hs, h->evProduced, 0));
  double bytes = 0.0;
  {
    MgField* f = mg_lookup_field(h->grid, owner, field);
    if (!f || !f->p) MG_FAIL("mg_p2p_exchange: unknown field");
    MG_TRY(p2p_exchange_field(h, f, width, hs, compMask));
    int n = 0;
    for (int c = 0; c < f->nComp && c < 32; ++c) n += (compMask >> c) & 1u;
    bytes = (double)n * width * (double)h->grid->plane * sizeof(double);
  }
  cudaEvent_t done = h->evDone[h->nextDone];
  h->nextDone = (h->nextDone + 1) % 4;
  MG_CUDA(cudaEventRecord(done, hs));
  mg_halo_set_pending(done, bytes);     // a later exchange on the same stream supersedes an earlier one
  return 0;
}

static int p2p_exchange_on(mg_p2p* h, void* owner, int field, int width, cudaStream_t st) {
  if (!h) MG_FAIL("mg_p2p_exchange: null handle");
  MgField* f = mg_lookup_field(h->grid, owner, field);
  if (!f || !f->p) MG_FAIL("mg_p2p_exchange: unknown field");
  return p2p_exchange_field(h, f, width, st);
}

// fillGhostPoints for an arbitrary (padded) array of the grid: what every operator application along the
// decomposed direction does in the reference (src/StencilOperatorImpl.f90:66, src/MPIHelperImpl.f90:113-389).  Used
// by the operator-by-operator path (mg_grid_apply); components are sent in batches that fit the staging buffers.
int mg_p2p_exchange_view(mg_p2p* h, const double* comp0, size_t compStride, int nComp, int width) {
  if (!h || !comp0) MG_FAIL("mg_p2p_exchange_view: null argument");
  if (width <= 0) return 0;
  MG_TRY(mg_halo_wait_pending());
  const size_t chunk = h->grid->plane * (size_t)width;
  int batch = (int)std::min<size_t>(MG_P2P_MAX_COMP, h->capacity / chunk);
  if (batch < 1) MG_FAIL("mg_p2p_exchange_view: staging buffers are smaller than one component");
  for (int c0 = 0; c0 < nComp; c0 += batch) {
    MgField v;
    v.p = const_cast<double*>(comp0) + (size_t)c0 * compStride;
    v.nComp = std::min(batch, nComp - c0);
    v.compStride = compStride;
    v.interiorOffset = 0;
    MG_TRY(p2p_exchange_field(h, &v, width, mg_stream()));
  }
  return 0;
}

static int p2p_exchange_field(mg_p2p* h, const MgField* fAll, int width, cudaStream_t st, unsigned compMask) {
  mg_grid* g = h->grid;
  // compMask selects the components whose ghost planes the consumer reads (bit c = component c)
  int comps[MG_P2P_MAX_COMP * 2];
  int nSel = 0;
  for (int c = 0; c < fAll->nComp && c < 32; ++c)
    if ((compMask >> c) & 1u) {
      if (nSel >= MG_P2P_MAX_COMP) MG_FAIL("mg_p2p_exchange: too many components");
      comps[nSel++] = c;
    }
  struct View { const MgField* f; const int* comps; int nComp; double* comp(int i) const { return f->comp(comps[i]); } };
  const View view{fAll, comps, nSel};
  const View* f = &view;
  const size_t chunk = g->plane * (size_t)width;
  // OVERLAP periodicity: the first and the last point of the direction coincide, so the first rank skips its first
  // plane and the last rank its last one (periodicOffset of src/MPIHelperImpl.f90:203-296, set by updateOperator)
  const bool ovl = g->periodicityType[2] == MG_PERIODIC_OVERLAP;
  const int ovLo = (ovl && g->procCoords[2] == 0) ? 1 : 0, ovHi = (ovl && g->procCoords[2] == g->procDims[2] - 1) ? 1 : 0;
  if (width > g->gk || width + std::max(ovLo, ovHi) > g->localSize[2]) MG_FAIL("mg_p2p_exchange: width exceeds ghost capacity");
  if (f->nComp > MG_P2P_MAX_COMP || chunk * (size_t)f->nComp > h->capacity)
    MG_FAIL("mg_p2p_exchange: field exceeds the staging capacity");
  const unsigned long long n = h->uses[0];
  const int parity = (int)(n & 1);
  const unsigned long long use = n >> 1;
  // push my low planes to prev's HIGH-ghost staging, my high planes to next's LOW-ghost staging (one launch,
  // blockIdx.y = side)
  PushPair pp;
  std::memset(&pp, 0, sizeof(pp));
  for (int side = 0; side < 2; ++side) {
    if (!h->peer[side]) continue;
    PushArgs& a = pp.s[side];
    a.nComp = f->nComp;
    a.chunk = chunk;
    for (int c = 0; c < f->nComp; ++c)
      a.src[c] = f->comp(c) + (side == 0 ? g->plane * (size_t)ovLo : g->plane * (size_t)(g->localSize[2] - width - ovHi));
    const int peerFace = 1 - side;
    a.dst = reinterpret_cast<double*>(h->peer[side] + mg_p2p::bufOffset(h->capacity, peerFace, parity));
    a.dstFlag = &reinterpret_cast<Shared*>(h->peer[side])->data[peerFace][parity];
    a.ackFlag = &h->flags()->ack[side][parity];
    a.use = use;
    a.counter = h->counters + side;
    a.error = h->error;
    a.vec = (chunk % 2 == 0) && ((uintptr_t)a.dst % 16 == 0);
    for (int c = 0; c < f->nComp; ++c) a.vec = a.vec && ((uintptr_t)a.src[c] % 16 == 0);
  }
  k_push<<<dim3(P2P_BLOCKS, 2), P2P_THREADS, 0, st>>>(pp);
  UnpackPair up;
  std::memset(&up, 0, sizeof(up));
  for (int face = 0; face < 2; ++face) {
    if (!h->peer[face]) continue;      // ghost face `face` is filled by neighbour `face`
    UnpackArgs& a = up.s[face];
    a.nComp = f->nComp;
    a.chunk = chunk;
    for (int c = 0; c < f->nComp; ++c)
      a.dst[c] = f->comp(c) + (face == 0 ? -(ptrdiff_t)chunk : (ptrdiff_t)(g->plane * (size_t)g->localSize[2]));
    a.src = reinterpret_cast<const double*>(h->base + mg_p2p::bufOffset(h->capacity, face, parity));
    a.dataFlag = &h->flags()->data[face][parity];
    // the sender pushed through its side (1 - face): that is where it waits for the acknowledgement
    a.ackFlag = &reinterpret_cast<Shared*>(h->peer[face])->ack[1 - face][parity];
    a.use = use;
    a.counter = h->counters + 2 + face;
    a.error = h->error;
    a.vec = (chunk % 2 == 0) && ((uintptr_t)a.src % 16 == 0);
    for (int c = 0; c < f->nComp; ++c) a.vec = a.vec && ((uintptr_t)a.dst[c] % 16 == 0);
  }
  k_unpack<<<dim3(P2P_BLOCKS, 2), P2P_THREADS, 0, st>>>(up);
  MG_CUDA(cudaGetLastError());
  mg_count_launches(2);
  h->uses[0] = h->uses[1] = n + 1;
  return 0;
}

// Exchange explicit device buffers with the two neighbours: sendLo -> prev's recvHi, sendHi -> next's recvLo
// (count doubles each).  Same staging / flag protocol as the plane exchange.
static int p2p_exchange_buffers(mg_p2p* h, const double* sendLo, const double* sendHi, double* recvLo, double* recvHi,
                                size_t count, cudaStream_t st, bool pairOnly = false, int phase = 3) {
  // pairOnly: a two-party link (block interface): both parties push through side 0 into the other's face-1 staging.
  // phase: bit 0 = push, bit 1 = unpack (a process with several links pushes on all of them before it waits on any)
  if (count > h->capacity) MG_FAIL("mg_p2p: face payload exceeds the staging capacity");
  const unsigned long long n = h->uses[0];
  const int parity = (int)(n & 1);
  const unsigned long long use = n >> 1;
  PushPair pp;
  std::memset(&pp, 0, sizeof(pp));
  for (int side = 0; side < 2; ++side) {
    if (!h->peer[side] || (pairOnly && side == 1)) continue;
    PushArgs& a = pp.s[side];
    a.nComp = 1;
    a.chunk = count;
    a.src[0] = side == 0 ? sendLo : sendHi;
    const int peerFace = 1 - side;
    a.dst = reinterpret_cast<double*>(h->peer[side] + mg_p2p::bufOffset(h->capacity, peerFace, parity));
    a.dstFlag = &reinterpret_cast<Shared*>(h->peer[side])->data[peerFace][parity];
    a.ackFlag = &h->flags()->ack[side][parity];
    a.use = use;
    a.counter = h->counters + side;
    a.error = h->error;
    a.vec = (count % 2 == 0) && ((uintptr_t)a.dst % 16 == 0) && ((uintptr_t)a.src[0] % 16 == 0);
  }
  if (phase & 1) {
    k_push<<<dim3(P2P_BLOCKS, 2), P2P_THREADS, 0, st>>>(pp);
    mg_count_launches(1);
  }
  if (!(phase & 2)) { MG_CUDA(cudaGetLastError()); return 0; }
  UnpackPair up;
  std::memset(&up, 0, sizeof(up));
  for (int face = 0; face < 2; ++face) {
    if (!h->peer[face] || (pairOnly && face == 0)) continue;
    UnpackArgs& a = up.s[face];
    a.nComp = 1;
    a.chunk = count;
    a.dst[0] = face == 0 ? recvLo : recvHi;
    a.src = reinterpret_cast<const double*>(h->base + mg_p2p::bufOffset(h->capacity, face, parity));
    a.dataFlag = &h->flags()->data[face][parity];
    a.ackFlag = &reinterpret_cast<Shared*>(h->peer[face])->ack[1 - face][parity];
    a.use = use;
    a.counter = h->counters + 2 + face;
    a.error = h->error;
    a.vec = (count % 2 == 0) && ((uintptr_t)a.src % 16 == 0) && ((uintptr_t)a.dst[0] % 16 == 0);
  }
  k_unpack<<<dim3(P2P_BLOCKS, 2), P2P_THREADS, 0, st>>>(up);
  MG_CUDA(cudaGetLastError());
  mg_count_launches(1);
  h->uses[0] = h->uses[1] = n + 1;
  return 0;
}

namespace {
// pack the first (side 0) and last (side 1) `w` points of every grid line along DIR, for nComp components, in the
// ghost-buffer layout of the reference (src/MPIHelperImpl.f90:175-296): q + w * (line + lines * component)
template <int DIR>
__global__ void k_pack_faces(const double* x, size_t cs, int nComp, long nx, long ny, long nz, int w, int ovLo, int ovHi,
                             double* lo, double* hi) {
  const long n = DIR == 0 ? nx : ny;
  const long lines = DIR == 0 ? ny * nz : nx * nz;
  const long total = (long)w * lines * nComp;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int q = (int)(t % w);
    const long line = (t / w) % lines;
    const int l = (int)(t / ((long)w * lines));
    long p0;      // index of the line's first point
    long stride;
    if (DIR == 0) { p0 = nx * line; stride = 1; }                       // line = j + ny k
    else { const long i = line % nx, k = line / nx; p0 = i + nx * ny * k; stride = nx; }   // line = i + nx k
    const double* xl = x + (size_t)l * cs;
    lo[t] = xl[p0 + (long)(q + ovLo) * stride];
    hi[t] = xl[p0 + (n - w + q - ovHi) * stride];
  }
}
}  // namespace

// fillGhostPoints along direction 0 or 1 (reference src/MPIHelperImpl.f90:113-296) for an operator application:
// returns device pointers to the received ghost buffers (null on a side without a neighbour).
int mg_p2p_exchange_faces(mg_p2p* h, const double* in, size_t inCs, int nComp, int width, const double** ghostPrev,
                          const double** ghostNext) {
  if (!h || !in) MG_FAIL("mg_p2p_exchange_faces: null argument");
  mg_grid* g = h->grid;
  const int dir = h->direction;
  if (dir > 1) MG_FAIL("mg_p2p_exchange_faces: the handle serves direction 3 (plane exchange)");
  *ghostPrev = nullptr;
  *ghostNext = nullptr;
  if (width <= 0) return 0;
  MG_TRY(mg_halo_wait_pending());
  const long lines = (long)(g->N / g->localSize[dir]);
  const size_t count = (size_t)width * lines * nComp;
  if (count > h->capacity) MG_FAIL("mg_p2p_exchange_faces: more components than the halo was created for");
  // OVERLAP periodicity: the first rank's first point and the last rank's last point are one point; the duplicate
  // does not travel (periodicOffset of src/MPIHelperImpl.f90:203-296)
  const bool ovl = g->periodicityType[dir] == MG_PERIODIC_OVERLAP;
  const int ovLo = (ovl && g->procCoords[dir] == 0) ? 1 : 0, ovHi = (ovl && g->procCoords[dir] == g->procDims[dir] - 1) ? 1 : 0;
  if (width + std::max(ovLo, ovHi) > g->localSize[dir]) MG_FAIL("mg_p2p_exchange_faces: the rank's extent is smaller than the stencil half-width");
  cudaStream_t st = mg_stream();
  const unsigned blocks = (unsigned)std::min<size_t>((count + 255) / 256, 4096);
  if (dir == 0)
    k_pack_faces<0><<<blocks, 256, 0, st>>>(in, inCs, nComp, g->localSize[0], g->localSize[1], g->localSize[2], width,
                                            ovLo, ovHi, h->faceSend[0], h->faceSend[1]);
  else
    k_pack_faces<1><<<blocks, 256, 0, st>>>(in, inCs, nComp, g->localSize[0], g->localSize[1], g->localSize[2], width,
                                            ovLo, ovHi, h->faceSend[0], h->faceSend[1]);
  MG_CUDA(cudaGetLastError());
  mg_count_launches(1);
  MG_TRY(p2p_exchange_buffers(h, h->faceSend[0], h->faceSend[1], h->faceRecv[0], h->faceRecv[1], count, st));
  if (h->peer[0]) *ghostPrev = h->faceRecv[0];
  if (h->peer[1]) *ghostNext = h->faceRecv[1];
  return 0;
}

// A halo handle for direction `direction` (0, 1: packed faces; 2: k planes = mg_p2p_create).
int mg_p2p_create_dir(mg_grid* g, int direction, int maxComp, int width, mg_p2p** out) {
  if (direction == 2) return mg_p2p_create(g, maxComp, width, out);
  if (!g || !out) MG_FAIL("mg_p2p_create_dir: null argument");
  if (direction < 0 || direction > 1) MG_FAIL("mg_p2p_create_dir: direction must be 0, 1 or 2");
  if (maxComp < 1 || maxComp > MG_P2P_MAX_COMP) MG_FAIL("mg_p2p_create_dir: component count out of range");
  if (width < 1 || width > g->localSize[direction]) MG_FAIL("mg_p2p_create_dir: width exceeds the local extent");
  mg_p2p* h = new mg_p2p;
  h->grid = g;
  h->direction = direction;
  h->capacity = (g->N / g->localSize[direction]) * (size_t)width * (size_t)maxComp;
  h->capacity += h->capacity % 2;
  h->bytes = sizeof(Shared) + 4 * h->capacity * sizeof(double);
  MG_CUDA(cudaMalloc(&h->base, h->bytes));
  MG_CUDA(cudaMemset(h->base, 0, h->bytes));
  MG_CUDA(cudaMalloc(&h->counters, 4 * sizeof(unsigned int)));
  MG_CUDA(cudaMemset(h->counters, 0, 4 * sizeof(unsigned int)));
  MG_CUDA(cudaHostAlloc(&h->error, sizeof(int), cudaHostAllocMapped));
  *h->error = 0;
  for (int i = 0; i < 2; ++i) {
    MG_CUDA(cudaMalloc(&h->faceSend[i], h->capacity * sizeof(double)));
    MG_CUDA(cudaMalloc(&h->faceRecv[i], h->capacity * sizeof(double)));
  }
  mg_p2p_register(h, true);
  MG_CUDA(cudaDeviceSynchronize());
  g->haloDir[direction] = h;
  *out = h;
  return 0;
}

// A two-party link for block interfaces whose blocks live in different processes: `capacity` doubles each way.
// Connect BOTH sides to the partner (mg_p2p_connect(h, 0, partner, 0); mg_p2p_connect(h, 1, partner, 1)).
int mg_p2p_create_pair(size_t capacity, mg_p2p** out) {
  if (!out || capacity == 0) MG_FAIL("mg_p2p_create_pair: invalid argument");
  mg_p2p* h = new mg_p2p;
  h->grid = nullptr;
  h->direction = -1;
  h->capacity = capacity + capacity % 2;
  h->bytes = sizeof(Shared) + 4 * h->capacity * sizeof(double);
  MG_CUDA(cudaMalloc(&h->base, h->bytes));
  MG_CUDA(cudaMemset(h->base, 0, h->bytes));
  MG_CUDA(cudaMalloc(&h->counters, 4 * sizeof(unsigned int)));
  MG_CUDA(cudaMemset(h->counters, 0, 4 * sizeof(unsigned int)));
  MG_CUDA(cudaHostAlloc(&h->error, sizeof(int), cudaHostAllocMapped));
  *h->error = 0;
  for (int i = 0; i < 2; ++i) {
    MG_CUDA(cudaMalloc(&h->faceSend[i], h->capacity * sizeof(double)));
    MG_CUDA(cudaMalloc(&h->faceRecv[i], h->capacity * sizeof(double)));
  }
  mg_p2p_register(h, true);
  MG_CUDA(cudaDeviceSynchronize());
  *out = h;
  return 0;
}
double* mg_p2p_pair_outbox(mg_p2p* h) { return h ? h->faceSend[0] : nullptr; }
const double* mg_p2p_pair_inbox(mg_p2p* h) { return h ? h->faceRecv[1] : nullptr; }
// phase 1: send the first `count` doubles of the outbox; phase 2: the partner's land in the inbox (stream-ordered,
// no host synchronisation).  A process pushes on all its links before it waits on any of them.
int mg_p2p_exchange_pair(mg_p2p* h, size_t count, int phase) {
  if (!h || !h->peer[0] || !h->peer[1]) MG_FAIL("mg_p2p_exchange_pair: the link is not connected to its partner");
  MG_TRY(mg_halo_wait_pending());
  return p2p_exchange_buffers(h, h->faceSend[0], nullptr, nullptr, h->faceRecv[1], count, mg_stream(), true, phase);
}

// 0 when no spin-wait timed out so far (synchronises the library stream)
int mg_p2p_check(mg_p2p* h) {
  if (!h) MG_FAIL("mg_p2p_check: null handle");
  MG_TRY(mg_halo_wait_pending());
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  if (*(volatile int*)h->error) MG_FAIL("mg_p2p: a halo exchange timed out waiting for its neighbour (MG_P2P_TIMEOUT_S)");
  return 0;
}

int mg_p2p_destroy(mg_p2p* h) {
  if (!h) return 0;
  cudaDeviceSynchronize();
  if (h->grid && h->grid->halo == h) h->grid->halo = nullptr;
  for (int d = 0; d < 2; ++d)
    if (h->grid && h->grid->haloDir[d] == h) h->grid->haloDir[d] = nullptr;
  for (int i = 0; i < 2; ++i) { cudaFree(h->faceSend[i]); cudaFree(h->faceRecv[i]); }
  for (int s = 0; s < 2; ++s)
    if (h->peerMapped[s] && h->peer[s]) cudaIpcCloseMemHandle(h->peer[s]);
  cudaFree(h->base);
  cudaFree(h->counters);
  mg_p2p_register(h, false);
  cudaFreeHost(h->error);
  delete h;
  return 0;
}
