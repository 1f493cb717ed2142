// Shared declarations of the fused hot-path sweeps (rhs_fused.cu, fused_adjoint1.cu, fused_sweepbd.cu):
// argument block, operator tables, tile placement, flux evaluation and the host-side launch helpers.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>

#include "grid.h"
#include "rhs_fused.h"
#include "stencil_apply.h"

namespace {

constexpr int TX = 16, TY = 16, NT = TX * TY;

struct LineOp {                 // one 1-D operator: interior stencil + closure tables
  int sym, lo, nInt;
  double c[MG_MAX_INTERIOR];
  int depth, width, hasB0, hasB1;
  const double* b1;             // device [depth][MG_MAX_BWIDTH]
  const double* b2;
};

struct DirInfo {
  int n, periodic, o1, o2;      // extent, wraps?, periodicOffset(1:2)
  int normDepth, hasB0, hasB1;
  double norm[MG_MAX_BDEPTH];   // first-derivative norm (applyNormInverse in the dissipation)
};

struct DevOps;
constexpr int MG_PF_MAX = 48;

struct FusedArgs {
  int nx, ny, nz;
  long plane;
  size_t cs;                    // component stride of every field
  int wrapK;                    // 1: single rank, periodic in k -> wrap plane index; 0: read ghost planes
  int ghostK;                   // ghost planes below plane 0 in storage (TMA coordinates count from the storage start)
  int kBeg, kEnd, kChunk;
  int zOff, zMul;               // k-chunk of a block = zOff + blockIdx.z * zMul (interior / boundary chunk launches)
  int overlapPays;              // choose_chunks: splitting the sweep around the pending exchange is the cheaper plan
  int curvilinear, viscous;
  DirInfo dir[3];
  LineOp D[3], Dd[3], Dt[3];    // first derivative, dissipation, dissipation transpose
  PhysParams pp;
  double dissAmount;
  const double *Q, *m, *jac, *arc;
  double *tauq, *diss;          // sweep A outputs
  const double *tauqIn, *dissIn;
  // sweep B outputs
  double *rhs;                  // when !fuseRk
  const double *b1in; double *b1out, *b2, *Qout;
  int fuseRk, stage;
  double dt;
  double rkB, rkQ;              // dt x RK4 weights of the accumulator / state update of this stage
  int prefetch;                 // planes of L2 prefetch distance (0 = off)
  int composite;                // composite dissipation (single operator) instead of Dt(-arc Dd)
  const struct DevOps* ops;     // device copy of the operators for the (out-of-line) closure path
  // L2 prefetch streams: pf[0..pfIn) are read at the arriving plane, pf[pfIn..pfAll) at the output plane
  const double* pf[MG_PF_MAX];
  int pfIn, pfAll;
  // adjoint sweeps
  const double *Win, *diffIn, *rhsIn;
  double* diffOut;
  int dissOn;
};

// L2 prefetch of the 128-byte line holding p: used to pull the planes the march will need two steps
// ahead while the current plane is being processed (costs no registers, unlike a software pipeline).
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Pull the rows of the planes the march needs `prefetch` steps ahead into L2.  The 16 lanes of a tile row
// share the row's streams (one 128-byte line per stream and row), so a thread issues pfAll/16 prefetches
// per plane instead of one lane issuing all of them.
__device__ __forceinline__ void prefetch_streams(const FusedArgs& a, int kIn, bool okIn, int kOut, bool okOut,
                                                 long rowOff, int lane16, bool rowOk) {
  if (!rowOk) return;
  for (int sid = lane16; sid < a.pfAll; sid += 16) {
    const bool in = sid < a.pfIn;
    if (in ? okIn : okOut) prefetch_l2(a.pf[sid] + (long)(in ? kIn : kOut) * a.plane + rowOff);
  }
}

// The same, with everything that does not change along the march hoisted out of it: a thread keeps the base
// (stream + tile row) of its NPF streams; one step costs a select, a multiply-add and the prefetch per stream
// (the loop above re-derived stream, row and validity every plane: ~70 instructions per thread and plane).
template <int NPF, bool KEEP = true>
struct PfItems {
  const double* base[KEEP ? NPF : 1];
  long rowOff;
  int lane;
  unsigned inMask;
  bool ok;
  // KEEP: the stream bases live in registers (2 per stream).  !KEEP: for the kernels that have no registers to
  // spare the base is re-read from the argument block (one indexed constant load) every plane.
  __device__ __forceinline__ void init(const FusedArgs& a, long rowOffset, int lane16, bool rowOk) {
    inMask = 0u;
    rowOff = rowOffset;
    lane = lane16;
    ok = rowOk;
    if constexpr (KEEP) {
#pragma unroll
      for (int n = 0; n < NPF; ++n) {
        const int sid = lane16 + 16 * n;
        base[n] = (rowOk && sid < a.pfAll) ? a.pf[sid] + rowOffset : nullptr;
        if (sid < a.pfIn) inMask |= 1u << n;
      }
    }
  }
  // stateless form (nothing kept across planes): the caller passes the tile-row offset again
  __device__ static __forceinline__ void issue_stateless(const FusedArgs& a, int kIn, bool okIn, int kOut, bool okOut,
                                                         long rowOffset, int lane16, bool rowOk) {
    if (!rowOk) return;
    const long offIn = rowOffset + (long)kIn * a.plane, offOut = rowOffset + (long)kOut * a.plane;
#pragma unroll
    for (int n = 0; n < NPF; ++n) {
      const int sid = lane16 + 16 * n;
      const bool in = sid < a.pfIn;
      if (sid < a.pfAll && (in ? okIn : okOut)) prefetch_l2(a.pf[sid] + (in ? offIn : offOut));
    }
  }
  __device__ __forceinline__ void issue(const FusedArgs& a, int kIn, bool okIn, int kOut, bool okOut) const {
    if constexpr (KEEP) {
#pragma unroll
      for (int n = 0; n < NPF; ++n) {
        const bool in = (inMask >> n) & 1u;
        if (base[n] != nullptr && (in ? okIn : okOut)) prefetch_l2(base[n] + (long)(in ? kIn : kOut) * a.plane);
      }
    } else {
      if (!ok) return;
      const long offIn = rowOff + (long)kIn * a.plane, offOut = rowOff + (long)kOut * a.plane;
#pragma unroll
      for (int n = 0; n < NPF; ++n) {
        const int sid = lane + 16 * n;
        const bool in = sid < a.pfIn;
        if (sid < a.pfAll && (in ? okIn : okOut)) prefetch_l2(a.pf[sid] + (in ? offIn : offOut));
      }
    }
  }
};

__device__ __forceinline__ int wrap_index(int c, const DirInfo& d) {
  if (c >= 0 && c < d.n) return c;
  if (!d.periodic) return -1;
  return c < 0 ? d.n + c - d.o1 : c - d.n + d.o2;
}

// Apply a 1-D operator at coordinate c of a line of extent n; get(cc) returns the value at coordinate cc.
template <class G>
__device__ __forceinline__ double line_apply(const LineOp& op, int c, int n, G&& get) {
  if (op.hasB0 && c < op.depth) {
    double r = 0.0;
    for (int s = 0; s < op.width; ++s) r += op.b1[c * MG_MAX_BWIDTH + s] * get(s);
    return r;
  }
  if (op.hasB1 && c >= n - op.depth) {
    const int m = n - 1 - c, first = n - op.width;
    double r = 0.0;
    for (int s = 0; s < op.width; ++s) r += op.b2[m * MG_MAX_BWIDTH + s] * get(first + s);
    return r;
  }
  double r = 0.0;
  if (op.sym == MG_SKEW_SYMMETRIC) {
    const int h = op.nInt / 2;
    for (int q = 1; q <= h; ++q) r += op.c[q - op.lo] * (get(c + q) - get(c - q));
  } else if (op.sym == MG_SYMMETRIC) {
    const int h = op.nInt / 2;
    for (int q = 1; q <= h; ++q) r += op.c[q - op.lo] * (get(c + q) + get(c - q));
    r += op.c[0 - op.lo] * get(c);
  } else {
    for (int q = 0; q < op.nInt; ++q) r += op.c[q] * get(c + op.lo + q);
  }
  return r;
}

// Non-composite dissipation along a line at coordinate c (reference src/RhsHelperImpl.f90:68-77):
//   H^-1 Dt ( -arc * (Dd q) ); getq(cc) the field, getarc(cc) the arc length.
template <class GQ, class GA>
__device__ __forceinline__ double line_dissipation(const LineOp& Dd, const LineOp& Dt, const DirInfo& di, int c,
                                                   GQ&& getq, GA&& getarc) {
  const int n = di.n;
  double r = line_apply(Dt, c, n, [&](int cc) { return -getarc(cc) * line_apply(Dd, cc, n, getq); });
  if (di.hasB0 && c < di.normDepth) r = r / di.norm[c];
  if (di.hasB1 && c >= n - di.normDepth) r = r / di.norm[n - 1 - c];
  return r;
}

// Device-resident copy of the operator tables, used by the out-of-line closure path below (keeps the
// boundary-tile code out of the hot kernels' instruction stream and register allocation).
struct DevOps {
  LineOp D[3], Dd[3], Dt[3];
  DirInfo dir[3];
};

// line[(cc - c0) * stride] holds the value at coordinate cc of the line (shared-memory tile, read
// through generic addresses: this is the slow, rarely taken path).
__device__ __noinline__ double g_line_apply(const LineOp* op, int c, int n, const double* line, int stride, int c0) {
  return line_apply(*op, c, n, [&](int cc) { return line[(long)(cc - c0) * stride]; });
}
__device__ __noinline__ double g_line_dissipation(const LineOp* Dd, const LineOp* Dt, const DirInfo* di, int c,
                                                  const double* q, const double* arc, int stride, int c0) {
  return line_dissipation(*Dd, *Dt, *di, c, [&](int cc) { return q[(long)(cc - c0) * stride]; },
                          [&](int cc) { return arc[(long)(cc - c0) * stride]; });
}

// Tile layout [field][H][W]; dirIdx 0: along columns (i), 1: along rows (j); c0 = coordinate of index 0.
template <int W, int H>
__device__ __forceinline__ double tile_line_apply(const LineOp* op, int c, int n, const double* T0, int f, int row,
                                                  int col, int dirIdx, int c0) {
  return dirIdx == 0 ? g_line_apply(op, c, n, T0 + ((size_t)f * H + row) * W, 1, c0)
                     : g_line_apply(op, c, n, T0 + (size_t)f * H * W + col, W, c0);
}
template <int W, int H>
__device__ __forceinline__ double tile_line_dissipation(const DevOps* ops, int d, int c, const double* T0, int fq,
                                                        int fa, int row, int col, int c0) {
  return d == 0 ? g_line_dissipation(&ops->Dd[d], &ops->Dt[d], &ops->dir[d], c, T0 + ((size_t)fq * H + row) * W,
                                     T0 + ((size_t)fa * H + row) * W, 1, c0)
                : g_line_dissipation(&ops->Dd[d], &ops->Dt[d], &ops->dir[d], c, T0 + (size_t)fq * H * W + col,
                                     T0 + (size_t)fa * H * W + col, W, c0);
}
template <int L>
__device__ __forceinline__ double strided_line_apply(const LineOp* op, int c, int n, const double* line, int stride,
                                                     int c0) {
  return g_line_apply(op, c, n, line, stride, c0);
}

// Tile placement: tiles are anchored at the origin except the last one of a direction, which is
// anchored at the far boundary so that it always contains the whole right closure block.
__device__ __forceinline__ void tile_origin(int t, int n, int T, int& c0, bool& isLast) {
  const int nt = (n + T - 1) / T;
  isLast = t == nt - 1;
  c0 = (isLast && n >= T) ? n - T : t * T;
}
__device__ __forceinline__ bool owns(int c, int n, int T, bool isLast) {
  if (c >= n) return false;
  if (isLast) return true;
  return (n < T) || (c < n - T) || (n % T == 0);
}


// ------------------------------------------------------------------------------- sweep B
// index of the unique stress entry (l, c) in the compact layout written by sweep A
template <int ND>
__device__ __forceinline__ constexpr int tau_index(int l, int c) {
  const int r0 = l < c ? l : c, c0 = l < c ? c : l;
  return r0 * ND - r0 * (r0 - 1) / 2 + (c0 - r0);
}

// Raw inputs of the flux evaluation at one point; loading and computing are separate so that all
// global loads of an iteration are issued back to back (one exposed memory latency, not one per use).
template <int ND>
struct RawPoint {
  double Q[ND + 2];
  double tq[ND * (ND + 1) / 2 + ND];
  double m[ND * ND];
};

template <int ND, int DIRS>
__device__ __forceinline__ constexpr bool needs_tq(int e) {
  constexpr int NTAU = ND * (ND + 1) / 2;
  for (int d = 0; d < ND; ++d) {
    if (!((DIRS >> d) & 1)) continue;
    if (e == NTAU + d) return true;
    for (int c = 0; c < ND; ++c)
      if (tau_index<ND>(d, c) == e) return true;
  }
  return false;
}

// DIRS: bit d set -> the flux along direction d will be needed.
template <int ND, int DIRS, bool CURV>
__device__ __forceinline__ void load_raw(const FusedArgs& a, long off, RawPoint<ND>& r) {
  constexpr int NU = ND + 2;
  constexpr int NTQ = ND * (ND + 1) / 2 + ND;
  const double* __restrict__ Qp = a.Q + off;
#pragma unroll
  for (int c = 0; c < NU; ++c) r.Q[c] = __ldg(Qp + (size_t)c * a.cs);
  if (a.viscous) {
    const double* __restrict__ tq = a.tauqIn + off;
#pragma unroll
    for (int e = 0; e < NTQ; ++e)
      if (CURV || needs_tq<ND, DIRS>(e)) r.tq[e] = __ldg(tq + (size_t)e * a.cs);
  }
  const double* __restrict__ mp = a.m + off;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    if (!((DIRS >> d) & 1)) continue;
    if constexpr (CURV) {
#pragma unroll
      for (int l = 0; l < ND; ++l) r.m[l + ND * d] = __ldg(mp + (size_t)(l + ND * d) * a.cs);
    } else {
      r.m[d + ND * d] = __ldg(mp + (size_t)(d + ND * d) * a.cs);
    }
  }
}

// Contravariant total fluxes from the raw inputs (reference CNSHelperImpl.f90:563-689 Cartesian
// inviscid - viscous, :772-840 metric transform).
template <int ND, int DIRS, bool CURV>
__device__ __forceinline__ void fluxes_from_raw(const FusedArgs& a, const RawPoint<ND>& r, double (*Fh)[ND + 2]) {
  constexpr int NU = ND + 2;
  constexpr int NTAU = ND * (ND + 1) / 2;
  const double* Q = r.Q;
  Prim<ND> s;
  dependent<ND>(Q, a.pp.gamma, s);
  if constexpr (!CURV) {
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      if (!((DIRS >> d) & 1)) continue;
      double F[NU];
      F[0] = Q[d + 1];
#pragma unroll
      for (int c = 0; c < ND; ++c) {
        if (c == d) F[c + 1] = Q[d + 1] * s.u[d] + s.p;
        else F[c + 1] = Q[(c < d ? c : d) + 1] * s.u[c < d ? d : c];
      }
      F[NU - 1] = s.u[d] * (Q[NU - 1] + s.p);
      if (a.viscous) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < ND; ++c) {
          const double t = r.tq[tau_index<ND>(d, c)];
          F[c + 1] = F[c + 1] - t;
          acc = (c == 0) ? s.u[0] * t : acc + s.u[c] * t;
        }
        F[NU - 1] = F[NU - 1] - (acc - r.tq[NTAU + d]);
      }
      const double md = r.m[d + ND * d];
#pragma unroll
      for (int c = 0; c < NU; ++c) Fh[d][c] = md * F[c];
    }
  } else {
    double tau[ND * ND], q[ND], Fc[ND][NU], Fv[NU];
    if (a.viscous) {
#pragma unroll
      for (int l = 0; l < ND; ++l)
#pragma unroll
        for (int c = 0; c < ND; ++c) tau[l + ND * c] = r.tq[tau_index<ND>(l, c)];
#pragma unroll
      for (int e = 0; e < ND; ++e) q[e] = r.tq[NTAU + e];
    }
#pragma unroll
    for (int l = 0; l < ND; ++l) cartesian_flux<ND>(l, Q, s, a.viscous, tau, q, Fc[l], Fv);
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      if (!((DIRS >> d) & 1)) continue;
#pragma unroll
      for (int l = 0; l < ND; ++l) {
        const double ml = r.m[l + ND * d];
#pragma unroll
        for (int c = 0; c < NU; ++c) Fh[d][c] = (l == 0) ? ml * Fc[0][c] : Fh[d][c] + ml * Fc[l][c];
      }
    }
  }
}

// --------------------------------------------------------------------------------- host
struct SchemeInfo { int R, dlo, dn, tlo, tn; };

bool scheme_of(const mg_grid* g, SchemeInfo* si) {
  // all directions must share one scheme family
  int R = -1;
  for (int d = 0; d < g->nD; ++d) {
    const mg_stencil* D = g->firstDerivative[d];
    if (!D || D->op.symmetryType != MG_SKEW_SYMMETRIC) return false;
    const int h = D->op.interiorWidth / 2;
    if (R < 0) R = h;
    if (h != R) return false;
  }
  if (R == 2) *si = {2, -1, 3, -1, 3};
  else if (R == 3) *si = {3, -2, 4, -1, 4};
  else if (R == 4) *si = {4, -2, 5, -2, 5};
  else return false;
  if (g->dissipationOn && !g->compositeDissipation)
    for (int d = 0; d < g->nD; ++d) {
      const MgDevOp& o = g->dissipation[d]->op;
      const MgDevOp& t = g->dissipationTranspose[d]->op;
      if (o.lo != si->dlo || o.nInterior != si->dn || t.lo != si->tlo || t.nInterior != si->tn) return false;
    }
  if (g->dissipationOn && g->compositeDissipation)
    for (int d = 0; d < g->nD; ++d)
      if (g->dissipation[d]->op.interiorWidth / 2 != R) return false;
  return true;
}

int fill_lineop(mg_stencil* s, LineOp* o) {
  MG_TRY(mg_stencil_upload(s));
  const MgDevOp& op = s->op;
  o->sym = op.symmetryType;
  o->lo = op.lo;
  o->nInt = op.nInterior;
  for (int k = 0; k < MG_MAX_INTERIOR; ++k) o->c[k] = op.interior[k];
  o->depth = op.boundaryDepth;
  o->width = op.boundaryWidth;
  o->hasB0 = op.hasDomainBoundary[0];
  o->hasB1 = op.hasDomainBoundary[1];
  o->b1 = &s->d_op->b1[0][0];
  o->b2 = &s->d_op->b2[0][0];
  return 0;
}

int fill_args(mg_state* s, FusedArgs* a) {
  mg_grid* g = s->grid;
  std::memset(a, 0, sizeof(*a));
  a->nx = g->localSize[0];
  a->ny = g->localSize[1];
  a->nz = g->localSize[2];
  a->plane = (long)g->plane;
  a->cs = s->rhs.compStride;
  a->wrapK = (g->nD == 3 && g->procDims[2] == 1) ? 1 : 0;
  a->ghostK = g->gk;
  a->kBeg = 0;
  a->kEnd = g->localSize[2];
  a->kChunk = g->localSize[2];
  a->zOff = 0;
  a->zMul = 1;
  a->curvilinear = g->isCurvilinear;
  a->viscous = s->opt.viscosityOn;
  for (int d = 0; d < g->nD; ++d) {
    const MgDevOp& op = g->firstDerivative[d]->op;
    DirInfo& di = a->dir[d];
    di.n = g->localSize[d];
    di.periodic = g->periodicityType[d] != MG_PERIODIC_NONE;
    di.o1 = op.periodicOffset[0];
    di.o2 = op.periodicOffset[1];
    di.normDepth = op.normDepth;
    di.hasB0 = op.hasDomainBoundary[0];
    di.hasB1 = op.hasDomainBoundary[1];
    for (int m = 0; m < MG_MAX_BDEPTH; ++m) di.norm[m] = op.normBoundary[m];
    MG_TRY(fill_lineop(g->firstDerivative[d], &a->D[d]));
    if (g->dissipationOn) {
      MG_TRY(fill_lineop(g->dissipation[d], &a->Dd[d]));
      if (!g->compositeDissipation) MG_TRY(fill_lineop(g->dissipationTranspose[d], &a->Dt[d]));
    }
  }
  a->composite = (g->compositeDissipation || !g->dissipationOn) ? 1 : 0;
  a->pp = s->phys();
  // L2 prefetch distance in planes (measured on B200: 1 is best for sweeps A and B, 2 for the dissipation sweep)
  const int pf = mg_tuning_get("MG_PREFETCH", 1);
  a->prefetch = pf;
  a->dissAmount = s->opt.dissipationAmount;
  a->m = g->metrics.comp(0);
  a->jac = g->jacobian.comp(0);
  a->arc = g->arcLengths.comp(0);
  return 0;
}

// L2 prefetch stream table (see prefetch_streams)
struct PfList {
  std::vector<const double*> in, out;
  size_t cs;
  void addIn(const double* base, int n) { if (base) for (int c = 0; c < n; ++c) in.push_back(base + (size_t)c * cs); }
  void addOut(const double* base, int n) { if (base) for (int c = 0; c < n; ++c) out.push_back(base + (size_t)c * cs); }
  void finish(FusedArgs* a) {
    // MG_PF_MASK (tuning): bit 0 = prefetch the arriving-plane streams, bit 1 = the output-plane streams
    const int mask = mg_tuning_get("MG_PF_MASK", 3);
    a->pfIn = a->pfAll = 0;
    if (mask & 1)
      for (const double* p : in) if (a->pfAll < MG_PF_MAX) { a->pf[a->pfAll++] = p; a->pfIn = a->pfAll; }
    if (mask & 2)
      for (const double* p : out) if (a->pfAll < MG_PF_MAX) a->pf[a->pfAll++] = p;
  }
};

// Device copy of the operator tables for the out-of-line closure path (built once per state and mode).
int upload_ops(mg_state* s, int which, FusedArgs* a) {
  if (!s->fusedOps[which]) {
    DevOps h;
    std::memset(&h, 0, sizeof(h));
    for (int d = 0; d < 3; ++d) { h.D[d] = a->D[d]; h.Dd[d] = a->Dd[d]; h.Dt[d] = a->Dt[d]; h.dir[d] = a->dir[d]; }
    MG_CUDA(cudaMalloc(&s->fusedOps[which], sizeof(DevOps)));
    MG_CUDA(cudaMemcpy(s->fusedOps[which], &h, sizeof(DevOps), cudaMemcpyHostToDevice));
  }
  a->ops = (const DevOps*)s->fusedOps[which];
  return 0;
}

// Split k into chunks so that the CTA count fills whole waves of (SMs x resident CTAs); every chunk
// re-streams 2R warm-up planes, so fewer, longer chunks are preferred when the fill is equal.
int choose_chunks(FusedArgs* a, int R, int residentPerSm, int tileX = TX, int tileY = TY) {
  const int tilesXY = ((a->nx + tileX - 1) / tileX) * ((a->ny + tileY - 1) / tileY);
  int best = 1;
  if (a->nz > 1) {
    const int forced = mg_tuning_get("MG_CHUNKS", 0);
    if (forced > 0) best = forced;
    else {
      const double slots = (double)mg_num_sms() * residentPerSm;
      // When an overlapped halo exchange is in flight the sweep may run as TWO launches (launch_split: interior
      // chunks at once, the first and the last chunk behind the exchange; the second launch fills most of the tail
      // of the first (measured: 512 x 512 x 64 slabs run fastest as 1 interior + 2 boundary chunks)) so that the exchange, E microseconds for its payload,
      // hides behind the interior chunks; or as one launch after the exchange.  Costs in CTA-plane units (one
      // plane of one wave ~ 5 us on B200 for these sweeps); MG_CHUNKS / MG_OVERLAP override the choice.
      const bool pending = mg_halo_is_pending() && a->kBeg == 0;
      const double tau = 5.0;                                           // us per plane and wave
      const double E = pending ? (30.0 + mg_halo_pending_bytes() / 0.4e6) / tau : 0.0;
      double bestCost = 1e300;
      for (int n = 1; n <= 32 && n * 4 * R <= a->nz; ++n) {
        const int chunk = (a->nz + n - 1) / n;
        const double perCta = chunk + 0.6 * 2 * R;               // planes a CTA streams (warm-up planes are cheaper)
        const double waves = tilesXY * (double)n / slots;
        double cost = ceil(waves) * perCta + E;                  // one launch, after the exchange
        if (pending && n >= 3 && chunk >= R) {
          const double interior = tilesXY * (double)(n - 2) / slots * perCta;
          const double splitCost = ceil(waves) * perCta + (E > interior ? E - interior : 0.0);
          if (splitCost < cost) cost = splitCost;
        }
        if (cost < bestCost - 1e-9) { bestCost = cost; best = n; }
      }
      a->overlapPays = 1;
      if (pending) {
        const int chunk = (a->nz + best - 1) / best;
        const double perCta = chunk + 0.6 * 2 * R, waves = tilesXY * (double)best / slots;
        const double interior = tilesXY * (double)(best - 2) / slots * perCta;
        a->overlapPays = (best >= 3 && chunk >= R &&
                          ceil(waves) * perCta + (E > interior ? E - interior : 0.0) < ceil(waves) * perCta + E) ? 1 : 0;
      }
    }
  }
  a->kChunk = (a->nz + best - 1) / best;
  return (a->nz + a->kChunk - 1) / a->kChunk;
}

dim3 tiles(const FusedArgs& a, int nChunks) {
  return dim3((a.nx + TX - 1) / TX, (a.ny + TY - 1) / TY, nChunks);
}

// true when an in-plane direction has a domain boundary (SBP closures): selects the kernel variant that
// carries the out-of-line closure path; fully periodic in-plane grids run the variant without it.
bool has_closures(const FusedArgs& a) {
  return a.dir[0].hasB0 || a.dir[0].hasB1 || a.dir[1].hasB0 || a.dir[1].hasB1;
}

// Launch a sweep of nChunks k-chunks.  When an overlapped halo exchange is in flight on the halo stream
// (mg_p2p_exchange_overlapped), the interior chunks - which read no ghost plane - go to the main stream at once
// and the first and the last chunk are queued on the HALO stream behind the exchange: they start as soon as the
// planes have arrived and fill the SMs the interior launch frees in its tail.  The event recorded after them is
// what the next launch on the main stream waits for.  go(args, zBlocks, stream) launches the kernel.
template <class F>
int launch_split(FusedArgs& a, int nChunks, int R, F&& go) {
  cudaEvent_t ev;
  const bool exchange = mg_halo_take_pending(&ev);
  MG_TRY(mg_halo_wait_boundary());          // the previous sweep's boundary chunks (halo stream) must be done
  if (!exchange) return go(a, nChunks, mg_stream());
  const bool split = nChunks >= 3 && a.kChunk >= R && a.kBeg == 0 && (a.overlapPays || mg_tuning_get("MG_CHUNKS", 0) > 0);
  if (!split) {
    MG_CUDA(cudaStreamWaitEvent(mg_stream(), ev, 0));
    return go(a, nChunks, mg_stream());
  }
  a.zOff = 1; a.zMul = 1;
  int rc = go(a, nChunks - 2, mg_stream());
  a.zOff = 0; a.zMul = nChunks - 1;
  mg_profile_suppress(true);                // the timing events live on the main stream
  if (rc == 0) rc = go(a, 2, mg_halo_stream());
  mg_profile_suppress(false);
  a.zOff = 0; a.zMul = 1;
  if (rc != 0) return rc;
  return mg_halo_mark_boundary();
}

// ------------------------------------------------------------------------------ TMA (bulk tensor copies)
// One elected thread asks the Tensor Memory Accelerator for a whole box of a field (cp.async.bulk.tensor ->
// SASS UTMALDG); the bytes land in shared memory asynchronously and complete an mbarrier transaction the
// consumers wait on.  No registers, no per-thread address arithmetic, and the request is issued one plane ahead.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MG_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MG_DONE_%=;\n"
      "bra MG_WAIT_%=;\n"
      "MG_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3) : "memory");
}

// Tensor map over a library field: dims (i, j, storage plane incl. ghost planes, component), fp64, no swizzle.
// Returns false when the layout does not meet the TMA alignment rules (odd nx) or the driver entry point is
// missing; the caller then launches the plain-load variant of the kernel.
inline bool make_field_tensor_map(const mg_grid* g, const MgField& f, int boxX, int boxY, int boxC, CUtensorMap* out) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeFn>(fn);
    else
      cudaGetLastError();
  }
  if (!encode) return false;
  const cuuint64_t nx = g->localSize[0], ny = g->localSize[1];
  const cuuint64_t planes = f.compStride / g->plane;         // nz + 2 * ghost planes
  if ((nx * sizeof(double)) % 16 || ((size_t)f.p % 16) || (f.compStride * sizeof(double)) % 16) return false;
  const cuuint64_t dims[4] = {nx, ny, planes, (cuuint64_t)f.nComp};
  const cuuint64_t strides[3] = {nx * sizeof(double), (cuuint64_t)g->plane * sizeof(double),
                                 (cuuint64_t)f.compStride * sizeof(double)};
  const cuuint32_t box[4] = {(cuuint32_t)boxX, (cuuint32_t)boxY, 1u, (cuuint32_t)boxC};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, f.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace
