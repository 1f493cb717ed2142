// Generic application of a 1-D SBP operator along one direction of a 3-D array: the CUDA
// counterpart of t_StencilOperator%apply / applyAtInteriorPoints / applyNorm / applyNormInverse /
// applyAndProjectOnBoundary / projectOnBoundaryAndApply
// (reference: src/StencilOperatorImpl.f90:35-457, :459-836, :838-1107) with fillGhostPoints
// (src/MPIHelperImpl.f90:113-389) folded into the neighbour fetch.
//
// This is the general path (any scheme, any closure, any direction, 1..16 components); the fused
// RHS sweeps in rhs_fused.cu are the hot path.  HBM-bound: one thread per point, components looped,
// i-contiguous so every warp reads/writes full 128-byte lines for all three directions.
#include "mg_common.h"
#include "stencil_apply.h"

int mg_stencil_upload(mg_stencil* s) {
  if (!s->d_op) MG_CUDA(cudaMalloc(&s->d_op, sizeof(MgDevOp)));
  if (s->dirty) {
    MG_CUDA(cudaMemcpyAsync(s->d_op, &s->op, sizeof(MgDevOp), cudaMemcpyHostToDevice, mg_stream()));
    MG_CUDA(cudaStreamSynchronize(mg_stream()));
    s->dirty = false;
  }
  return 0;
}

namespace {

template <int DIR>
__global__ void __launch_bounds__(256) k_apply(const MgDevOp* __restrict__ opp, ApplyArgs a) {
  const MgDevOp& op = *opp;
  const long nx = a.n[0], ny = a.n[1], nz = a.n[2];
  const long total = nx * ny * nz;
  const long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (p >= total) return;
  const long i = p % nx, j = (p / nx) % ny, k = p / (nx * ny);
  const long c = DIR == 0 ? i : (DIR == 1 ? j : k);
  const long n = a.n[DIR];
  const long stride = DIR == 0 ? 1 : (DIR == 1 ? nx : nx * ny);
  const long base = p - c * stride;        // index of the line's first point
  // index of this line among the lines normal to DIR (for explicit ghost buffers)
  const long lineIdx = DIR == 0 ? (j + ny * k) : (DIR == 1 ? (i + nx * k) : (i + nx * j));
  const long planeSize = total / n;
  const int g0 = op.nGhost[0], g1 = op.nGhost[1];

  int kind = 0;   // 0 interior, 1 left closure, 2 right closure
  long m = 0;
  // on a line shorter than two closure blocks the right closure wins, as in the reference, which writes the left
  // rows first and the right rows after them (src/StencilOperatorImpl.f90:73-102)
  if (op.hasDomainBoundary[1] && c >= n - op.boundaryDepth) { kind = 2; m = n - 1 - c; }
  else if (op.hasDomainBoundary[0] && c < op.boundaryDepth) { kind = 1; m = c; }
  // nGhost == 0 happens only on a domain-boundary side, where the closure rows cover the points
  // whose interior stencil would reach outside: every remaining point is an interior point.

  for (int l = 0; l < a.nComp; ++l) {
    const double* __restrict__ x = a.in + (size_t)l * a.inCompStride;
    double* __restrict__ y = a.out + (size_t)l * a.outCompStride;
    double r;
    if (kind == 1) {
      if (a.interiorOnly == 1) continue;
      r = 0.0;
      for (int s = 0; s < op.boundaryWidth; ++s) r += op.b1[m][s] * x[base + s * stride];
    } else if (kind == 2) {
      if (a.interiorOnly == 1) continue;
      r = 0.0;
      const long first = n - op.boundaryWidth;
      for (int s = 0; s < op.boundaryWidth; ++s) r += op.b2[m][s] * x[base + (first + s) * stride];
    } else {
      const double sPrev = (l == a.shiftComp && a.shiftPrev) ? a.shiftLen : 0.0;
      const double sNext = (l == a.shiftComp && a.shiftNext) ? a.shiftLen : 0.0;
      auto fetch = [&](long cc) -> double {
        if (cc < 0) {
          if (a.ghostPrev) return a.ghostPrev[(g0 + cc) + g0 * (lineIdx + planeSize * l)] - sPrev;
          if (a.padded) return x[base + cc * stride] - sPrev;
          cc = n + cc - op.periodicOffset[0];
          return x[base + cc * stride] - sPrev;
        } else if (cc >= n) {
          if (a.ghostNext) return a.ghostNext[(cc - n) + g1 * (lineIdx + planeSize * l)] + sNext;
          if (a.padded) return x[base + cc * stride] + sNext;
          cc = cc - n + op.periodicOffset[1];
          return x[base + cc * stride] + sNext;
        }
        return x[base + cc * stride];
      };
      r = 0.0;
      const int h = op.interiorWidth / 2;
      if (op.symmetryType == MG_SKEW_SYMMETRIC) {
        for (int q = 1; q <= h; ++q) r += op.interior[q - op.lo] * (fetch(c + q) - fetch(c - q));
      } else if (op.symmetryType == MG_SYMMETRIC) {
        for (int q = 1; q <= h; ++q) r += op.interior[q - op.lo] * (fetch(c + q) + fetch(c - q));
        r += op.interior[0 - op.lo] * x[p];
      } else {
        for (int q = 0; q < op.nInterior; ++q) r += op.interior[q] * fetch(c + op.lo + q);
      }
    }
    y[p] = r;
  }
}

// mode 0: multiply by norm, 1: divide
template <int DIR>
__global__ void __launch_bounds__(256) k_norm(const MgDevOp* __restrict__ opp, double* x, size_t compStride,
                                              int nComp, long nx, long ny, long nz, int inverse) {
  const MgDevOp& op = *opp;
  const long total = nx * ny * nz;
  const long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (p >= total) return;
  const long i = p % nx, j = (p / nx) % ny, k = p / (nx * ny);
  const long c = DIR == 0 ? i : (DIR == 1 ? j : k);
  const long n = DIR == 0 ? nx : (DIR == 1 ? ny : nz);
  double w = 1.0;
  bool touch = false;
  if (op.hasDomainBoundary[0] && c < op.normDepth) { w = op.normBoundary[c]; touch = true; }
  // NB: both closures can cover the same point only on grids too small to be valid
  if (op.hasDomainBoundary[1] && c >= n - op.normDepth) {
    const double w2 = op.normBoundary[n - 1 - c];
    w = touch ? w * w2 : w2;
    touch = true;
  }
  if (!touch) return;
  for (int l = 0; l < nComp; ++l) {
    double* y = x + (size_t)l * compStride;
    y[p] = inverse ? y[p] / w : y[p] * w;
  }
}

// face > 0: left boundary; apply == 1: applyAndProjectOnBoundary, apply == 0: projectOnBoundaryAndApply
template <int DIR>
__global__ void __launch_bounds__(256) k_boundary(const MgDevOp* __restrict__ opp, const double* in, double* out,
                                                  size_t compStride, int nComp, long nx, long ny, long nz,
                                                  int face, int applyThenProject) {
  const MgDevOp& op = *opp;
  const long total = nx * ny * nz;
  const long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (p >= total) return;
  const long i = p % nx, j = (p / nx) % ny, k = p / (nx * ny);
  const long c = DIR == 0 ? i : (DIR == 1 ? j : k);
  const long n = DIR == 0 ? nx : (DIR == 1 ? ny : nz);
  const long stride = DIR == 0 ? 1 : (DIR == 1 ? nx : nx * ny);
  const long base = p - c * stride;
  const bool left = face > 0;
  const bool active = left ? op.hasDomainBoundary[0] : op.hasDomainBoundary[1];
  for (int l = 0; l < nComp; ++l) {
    const double* x = in + (size_t)l * compStride;
    double r = 0.0;
    if (active) {
      if (applyThenProject) {
        if (left && c == 0) {
          for (int s = 0; s < op.boundaryWidth; ++s) r += op.b1[0][s] * x[base + s * stride];
        } else if (!left && c == n - 1) {
          const long first = n - op.boundaryWidth;
          for (int s = 0; s < op.boundaryWidth; ++s) r += op.b2[0][s] * x[base + (first + s) * stride];
        }
      } else {
        if (left && c < op.boundaryDepth) r = op.b1[c][0] * x[base];
        else if (!left && c >= n - op.boundaryDepth)
          r = op.b2[n - 1 - c][op.boundaryWidth - 1] * x[base + (n - 1) * stride];
      }
    }
    out[(size_t)l * compStride + p] = r;
  }
}

}  // namespace

int mg_apply_launch(mg_stencil* s, const ApplyArgs& a) {
  MG_TRY(mg_stencil_upload(s));
  const int d = s->direction - 1;
  const long total = (long)a.n[0] * a.n[1] * a.n[2];
  if (total <= 0) return 0;
  const MgDevOp& op = s->op;
  // Sanity: a closure must fit in the local extent (reference asserts the same in Debug builds).
  if ((op.hasDomainBoundary[0] || op.hasDomainBoundary[1]) && a.n[d] < op.boundaryWidth && op.interiorWidth > 0)
    MG_FAIL("stencil apply: local grid size smaller than the boundary stencil width");
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  cudaStream_t st = a.stream ? a.stream : mg_stream();
  if (d == 0) { k_apply<0><<<blocks, threads, 0, st>>>(s->d_op, a); mg_count_launches(1); }
  else if (d == 1) { k_apply<1><<<blocks, threads, 0, st>>>(s->d_op, a); mg_count_launches(1); }
  else { k_apply<2><<<blocks, threads, 0, st>>>(s->d_op, a); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return 0;
}

int mg_norm_launch(mg_stencil* s, double* x, size_t compStride, int nComp, const int n[3], int inverse,
                   cudaStream_t st) {
  MG_TRY(mg_stencil_upload(s));
  const int d = s->direction - 1;
  const long total = (long)n[0] * n[1] * n[2];
  if (total <= 0) return 0;
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (!st) st = mg_stream();
  if (d == 0) { k_norm<0><<<blocks, threads, 0, st>>>(s->d_op, x, compStride, nComp, n[0], n[1], n[2], inverse); mg_count_launches(1); }
  else if (d == 1) { k_norm<1><<<blocks, threads, 0, st>>>(s->d_op, x, compStride, nComp, n[0], n[1], n[2], inverse); mg_count_launches(1); }
  else { k_norm<2><<<blocks, threads, 0, st>>>(s->d_op, x, compStride, nComp, n[0], n[1], n[2], inverse); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return 0;
}

int mg_boundary_launch(mg_stencil* s, const double* in, double* out, size_t compStride, int nComp,
                       const int n[3], int face, int applyThenProject, cudaStream_t st) {
  MG_TRY(mg_stencil_upload(s));
  const int d = s->direction - 1;
  const long total = (long)n[0] * n[1] * n[2];
  if (total <= 0) return 0;
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (!st) st = mg_stream();
  if (d == 0)
    { k_boundary<0><<<blocks, threads, 0, st>>>(s->d_op, in, out, compStride, nComp, n[0], n[1], n[2], face, applyThenProject); mg_count_launches(1); }
  else if (d == 1)
    { k_boundary<1><<<blocks, threads, 0, st>>>(s->d_op, in, out, compStride, nComp, n[0], n[1], n[2], face, applyThenProject); mg_count_launches(1); }
  else
    { k_boundary<2><<<blocks, threads, 0, st>>>(s->d_op, in, out, compStride, nComp, n[0], n[1], n[2], face, applyThenProject); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return 0;
}
