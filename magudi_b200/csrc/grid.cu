// t_Grid on the device: field storage, metrics / Jacobian / norm (updateGrid), gradient and SBP
// inner product.  Reference: src/GridImpl.f90:142-291 (setup), :487-619 (operators), :621-744
// (coordinate derivatives), :746-1065 (updateGrid), :1067-1170 (inner products), :1172-1421 (gradient).
#include <cmath>
#include <algorithm>
#include <cstring>

#include "mg_common.h"
#include "stencil_apply.h"
#include "grid.h"

// ---------------------------------------------------------------------------------- fields
int mg_field_alloc(const mg_grid* g, int nComp, MgField* f) {
  mg_field_free(f);
  f->nComp = nComp;
  f->compStride = g->plane * (size_t)(g->localSize[2] + 2 * g->gk);
  f->interiorOffset = g->plane * (size_t)g->gk;
  const size_t bytes = f->compStride * (size_t)nComp * sizeof(double);
  MG_CUDA(cudaMalloc(&f->p, bytes));
  MG_CUDA(cudaMemsetAsync(f->p, 0, bytes, mg_stream()));
  f->owned = true;
  return 0;
}

void mg_field_free(MgField* f) {
  if (f->p && f->owned) cudaFree(f->p);
  f->p = nullptr;
  f->owned = false;
  f->nComp = 0;
}

int mg_field_zero(const mg_grid* g, MgField* f) {
  (void)g;
  MG_CUDA(cudaMemsetAsync(f->p, 0, f->compStride * (size_t)f->nComp * sizeof(double), mg_stream()));
  return 0;
}

int mg_field_upload(const mg_grid* g, MgField* f, const double* host) {
  MG_TRY(mg_halo_wait_pending());   // boundary chunks / exchanges still in flight on the halo stream
  for (int c = 0; c < f->nComp; ++c)
    MG_CUDA(cudaMemcpyAsync(f->comp(c), host + (size_t)c * g->N, g->N * sizeof(double), cudaMemcpyDefault,
                            mg_stream()));
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  return 0;
}

int mg_field_download(const mg_grid* g, const MgField* f, double* host) {
  MG_TRY(mg_halo_wait_pending());   // boundary chunks / exchanges still in flight on the halo stream
  for (int c = 0; c < f->nComp; ++c)
    MG_CUDA(cudaMemcpyAsync(host + (size_t)c * g->N, f->comp(c), g->N * sizeof(double), cudaMemcpyDefault,
                            mg_stream()));
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  return 0;
}

// ------------------------------------------------------------------------ elementwise helpers
namespace {

__global__ void k_mul(double* out, const double* a, const double* b, size_t n) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p < n) out[p] = a[p] * b[p];
}
__global__ void k_sub(double* out, const double* a, size_t n) {   // out -= a
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p < n) out[p] -= a[p];
}
__global__ void k_fill(double* out, double v, size_t n) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p < n) out[p] = v;
}

struct MetricArgs {
  const double* Ji[9];     // Ji[j*nD + i] = d x_i / d xi_j  (reference "jacobianMatrixInverse", column-major)
  double* m[9];
  double* jac;
  double* arc[3];
  const int* iblank;
  size_t N;
  int nD, curvilinear, planeFormulas;
};

// Pointwise part of updateGrid (reference src/GridImpl.f90:805-900 and :1010-1029).
__global__ void k_metrics(MetricArgs a) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  const bool hole = a.iblank && a.iblank[p] == 0;
  double J[9];
  for (int q = 0; q < a.nD * a.nD; ++q) J[q] = hole ? 0.0 : a.Ji[q][p];
  if (a.nD == 1) {
    a.jac[p] = J[0];
    a.m[0][p] = 1.0;
    a.arc[0][p] = 1.0;
  } else if (a.nD == 2) {
    if (a.curvilinear) {
      a.jac[p] = J[0] * J[3] - J[1] * J[2];
      const double m0 = J[3], m1 = -J[2], m2 = -J[1], m3 = J[0];
      a.m[0][p] = m0; a.m[1][p] = m1; a.m[2][p] = m2; a.m[3][p] = m3;
      a.arc[0][p] = sqrt(m0 * m0 + m1 * m1);
      a.arc[1][p] = sqrt(m2 * m2 + m3 * m3);
    } else {
      a.jac[p] = J[0] * J[3];
      a.m[0][p] = J[3]; a.m[1][p] = 0.0; a.m[2][p] = 0.0; a.m[3][p] = J[0];
      a.arc[0][p] = fabs(J[3]);
      a.arc[1][p] = fabs(J[0]);
    }
  } else {
    if (a.curvilinear)
      a.jac[p] = J[0] * (J[4] * J[8] - J[7] * J[5]) + J[3] * (J[7] * J[2] - J[1] * J[8]) +
                 J[6] * (J[1] * J[5] - J[4] * J[2]);
    else
      a.jac[p] = J[0] * J[4] * J[8];
    if (a.planeFormulas) {
      double m[9];
      if (a.curvilinear) {
        m[0] = J[4] * J[8] - J[7] * J[5];
        m[1] = J[6] * J[5] - J[3] * J[8];
        m[2] = J[3] * J[7] - J[6] * J[4];
        m[3] = J[7] * J[2] - J[1] * J[8];
        m[4] = J[0] * J[8] - J[6] * J[2];
        m[5] = J[6] * J[1] - J[0] * J[7];
        m[6] = J[1] * J[5] - J[4] * J[2];
        m[7] = J[3] * J[2] - J[0] * J[5];
        m[8] = J[0] * J[4] - J[3] * J[1];
      } else {
        for (int q = 0; q < 9; ++q) m[q] = 0.0;
        m[0] = J[4] * J[8];
        m[4] = J[0] * J[8];
        m[8] = J[0] * J[4];
      }
      for (int q = 0; q < 9; ++q) a.m[q][p] = m[q];
    }
  }
  if (hole) a.jac[p] = 1.0;
}

// 3-D arc lengths from the final metrics (reference :1012-1026), hole masking of the metrics (:1006-1008).
__global__ void k_arc3(double* const* mptr, double* a0, double* a1, double* a2, const int* iblank, size_t N,
                       int curvilinear) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= N) return;
  double m[9];
  const bool hole = iblank && iblank[p] == 0;
  for (int q = 0; q < 9; ++q) {
    if (hole) mptr[q][p] = 0.0;
    m[q] = mptr[q][p];
  }
  if (curvilinear) {
    a0[p] = sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    a1[p] = sqrt(m[3] * m[3] + m[4] * m[4] + m[5] * m[5]);
    a2[p] = sqrt(m[6] * m[6] + m[7] * m[7] + m[8] * m[8]);
  } else {
    a0[p] = fabs(m[0]);
    a1[p] = fabs(m[4]);
    a2[p] = fabs(m[8]);
  }
}

__global__ void k_norm_finish(double* norm, double* jac, size_t N, int* negFlag) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= N) return;
  const double J = jac[p];
  if (!(J > 0.0)) atomicExch(negFlag, 1);
  norm[p] = norm[p] * J;
  jac[p] = 1.0 / J;
}

struct GradArgs {
  const double* d[3];     // d f / d xi_i, nComp components each (component stride cs)
  const double* m[9];
  const double* jac;
  double* out;            // (nD*nComp) components, stride cs
  size_t cs, N;
  int nD, nComp, curvilinear;
};

// gradF(:, j + nD*c) = (1/J) sum_i M_ij d f_c / d xi_i (reference :1357-1413)
__global__ void k_gradient(GradArgs a) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  const double J = a.jac[p];
  for (int c = 0; c < a.nComp; ++c) {
    double dx[3];
    for (int i = 0; i < a.nD; ++i) dx[i] = a.d[i][(size_t)c * a.cs + p];
    for (int j = 0; j < a.nD; ++j) {
      double r;
      if (a.curvilinear) {
        r = a.m[j][p] * dx[0];
        for (int i = 1; i < a.nD; ++i) r += a.m[j + a.nD * i][p] * dx[i];
        r = J * r;
      } else {
        r = J * a.m[j + a.nD * j][p] * dx[j];
      }
      a.out[(size_t)(j + a.nD * c) * a.cs + p] = r;
    }
  }
}

// Deterministic two-stage reduction: sum_p f g norm [weight]  (reference :1067-1170)
__global__ void __launch_bounds__(256) k_inner(const double* f, const double* g, const double* norm,
                                               const double* weight, size_t cs, int nComp, size_t N,
                                               double* partial) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < N; p += (size_t)gridDim.x * blockDim.x) {
    const double w = weight ? norm[p] * weight[p] : norm[p];
    for (int c = 0; c < nComp; ++c) acc += f[(size_t)c * cs + p] * w * g[(size_t)c * cs + p];
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

inline unsigned nblocks(size_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

// ---------------------------------------------------------------------------------- grid
int mg_grid_create_impl(int index, int nD, const int globalSize[3], const int localSize[3], const int offset[3],
                        const int periodicityType[3], const double periodicLength[3], int isCurvilinear,
                        const int procDims[3], const int procCoords[3], mg_grid** out) {
  if (nD < 1 || nD > 3) MG_FAIL("mg_grid_create: nDimensions must be 1, 2 or 3");
  auto* g = new mg_grid();
  g->index = index;
  g->nD = nD;
  for (int i = 0; i < 3; ++i) {
    g->globalSize[i] = globalSize[i];
    g->localSize[i] = localSize[i];
    g->offset[i] = offset[i];
    g->periodicityType[i] = periodicityType[i];
    g->periodicLength[i] = periodicLength[i];
    g->procDims[i] = procDims[i];
    g->procCoords[i] = procCoords[i];
    if (localSize[i] <= 0) { delete g; MG_FAIL("mg_grid_create: local size must be positive"); }
    if (procDims[i] < 1 || procCoords[i] < 0 || procCoords[i] >= procDims[i]) { delete g; MG_FAIL("mg_grid_create: invalid process grid"); }
  }
  for (int i = nD; i < 3; ++i)
    if (globalSize[i] != 1) { delete g; MG_FAIL("mg_grid_create: extent beyond nDimensions must be 1"); }
  g->isCurvilinear = isCurvilinear;
  g->plane = (size_t)localSize[0] * localSize[1];
  g->N = g->plane * localSize[2];
  g->gk = (nD == 3) ? MG_GHOST_K : 0;
  MG_TRY(mg_field_alloc(g, nD, &g->coordinates));
  MG_TRY(mg_field_alloc(g, nD * nD, &g->metrics));
  MG_TRY(mg_field_alloc(g, 1, &g->jacobian));
  MG_TRY(mg_field_alloc(g, 1, &g->norm));
  MG_TRY(mg_field_alloc(g, nD, &g->arcLengths));
  *out = g;
  return 0;
}

void mg_grid_destroy_impl(mg_grid* g) {
  if (!g) return;
  mg_field_free(&g->coordinates);
  mg_field_free(&g->metrics);
  mg_field_free(&g->jacobian);
  mg_field_free(&g->norm);
  mg_field_free(&g->arcLengths);
  mg_field_free(&g->targetMollifier);
  mg_field_free(&g->controlMollifier);
  mg_field_free(&g->scratchA);
  mg_field_free(&g->scratchB);
  if (g->iblank) cudaFree(g->iblank);
  for (int i = 0; i < 3; ++i) {
    for (mg_stencil* s : {g->firstDerivative[i], g->adjointFirstDerivative[i], g->dissipation[i],
                          g->dissipationTranspose[i], g->filter[i]})
      if (s) {
        if (s->d_op) cudaFree(s->d_op);
        delete s;
      }
  }
  delete g;
}

// setupSpatialDiscretization (reference src/GridImpl.f90:487-619)
int mg_grid_setup_discretization_impl(mg_grid* g, const char* const schemes[3], int dissipationOn,
                                      int compositeDissipation, int useContinuousAdjoint) {
  int periodic[3];
  for (int i = 0; i < 3; ++i) periodic[i] = g->periodicityType[i] != MG_PERIODIC_NONE;
  g->dissipationOn = dissipationOn;
  g->compositeDissipation = compositeDissipation;
  for (int i = 0; i < g->nD; ++i) {
    const bool big = g->globalSize[i] > 1;
    const int ov = g->periodicityType[i] == MG_PERIODIC_OVERLAP;
    const std::string sch = schemes[i];
    const std::string name = big ? sch + " first derivative" : std::string("null matrix");
    MG_TRY(mg_stencil_create_impl(name.c_str(), &g->firstDerivative[i]));
    MG_TRY(mg_stencil_update_impl(g->firstDerivative[i], i + 1, g->procDims, g->procCoords, periodic, ov));
    if (useContinuousAdjoint || !big) {
      MG_TRY(mg_stencil_clone_impl(g->firstDerivative[i], &g->adjointFirstDerivative[i]));
      MG_TRY(mg_stencil_negate_impl(g->adjointFirstDerivative[i]));
    } else {
      MG_TRY(mg_stencil_get_adjoint_impl(g->firstDerivative[i], &g->adjointFirstDerivative[i]));
    }
    MG_TRY(mg_stencil_update_impl(g->adjointFirstDerivative[i], i + 1, g->procDims, g->procCoords, periodic, ov));
    if (dissipationOn) {
      const std::string dn = big ? sch + (compositeDissipation ? " composite dissipation" : " dissipation")
                                 : std::string("null matrix");
      MG_TRY(mg_stencil_create_impl(dn.c_str(), &g->dissipation[i]));
      MG_TRY(mg_stencil_update_impl(g->dissipation[i], i + 1, g->procDims, g->procCoords, periodic, ov));
      if (!compositeDissipation) {
        const std::string tn = big ? sch + " dissipation transpose" : std::string("null matrix");
        MG_TRY(mg_stencil_create_impl(tn.c_str(), &g->dissipationTranspose[i]));
        MG_TRY(mg_stencil_update_impl(g->dissipationTranspose[i], i + 1, g->procDims, g->procCoords, periodic, ov));
      }
    }
  }
  return 0;
}

// Generic operator application on a padded field (general path).
// fillGhostPoints + apply (reference src/StencilOperatorImpl.f90:52-104): along a decomposed direction the ghost
// points of the input come from the neighbours -- direction 3: ghost planes of the padded field, filled in place;
// directions 1 and 2: packed faces received into explicit ghost buffers.
static int apply_with_halo(mg_grid* g, mg_stencil* op, ApplyArgs& a) {
  const int dir = op->direction - 1;
  const int w = std::max(op->op.nGhost[0], op->op.nGhost[1]);
  if (dir >= 0 && dir < 3 && g->procDims[dir] > 1) {
    if (dir == 2) {
      a.padded = 1;
      if (!g->halo)
        MG_FAIL("operator application along a decomposed direction: no halo attached to the grid (mg_p2p_create); "
                "the NCCL fallback only serves the fused sweeps");
      MG_TRY(mg_p2p_exchange_view(g->halo, a.in, a.inCompStride, a.nComp, w));
    } else {
      if (!g->haloDir[dir])
        MG_FAIL("operator application along a decomposed direction: no halo attached to the grid for this "
                "direction (mg_p2p_create_dir)");
      MG_TRY(mg_p2p_exchange_faces(g->haloDir[dir], a.in, a.inCompStride, a.nComp, w, &a.ghostPrev, &a.ghostNext));
    }
  }
  return mg_apply_launch(op, a);
}

int mg_grid_apply(mg_grid* g, mg_stencil* op, const double* in, size_t inCs, double* out, size_t outCs,
                  int nComp) {
  ApplyArgs a;
  a.in = in;
  a.out = out;
  a.inCompStride = inCs;
  a.outCompStride = outCs;
  a.nComp = nComp;
  for (int i = 0; i < 3; ++i) a.n[i] = g->localSize[i];
  return apply_with_halo(g, op, a);
}

// computeCoordinateDerivatives (reference :621-744)
int mg_grid_coordinate_derivatives(mg_grid* g, int dir, MgField* out) {
  mg_stencil* D = g->firstDerivative[dir];
  ApplyArgs a;
  a.in = g->coordinates.comp(0);
  a.out = out->comp(0);
  a.inCompStride = g->coordinates.compStride;
  a.outCompStride = out->compStride;
  a.nComp = g->nD;
  for (int i = 0; i < 3; ++i) a.n[i] = g->localSize[i];
  if (g->periodicityType[dir] == MG_PERIODIC_PLANE) {
    a.interiorOnly = 1;
    a.shiftComp = dir;
    a.shiftLen = g->periodicLength[dir];
    a.shiftPrev = g->procCoords[dir] == 0;
    a.shiftNext = g->procCoords[dir] == g->procDims[dir] - 1;
  }
  // bricks split along i / j: the coordinates' ghost points arrive as packed faces; slabs along k keep the
  // explicit exchange of the coordinate field the caller issues before mg_grid_update (ghost planes, in place)
  if (dir < 2) return apply_with_halo(g, D, a);
  a.padded = (g->procDims[2] > 1) ? 1 : 0;
  return mg_apply_launch(D, a);
}

// updateGrid (reference :746-1065)
int mg_grid_update_impl(mg_grid* g, int* hasNegativeJacobian) {
  const int nD = g->nD;
  const size_t N = g->N;
  cudaStream_t st = mg_stream();
  for (int i = 0; i < nD; ++i)
    if (!g->firstDerivative[i]) MG_FAIL("mg_grid_update: spatial discretization has not been set up");
  MgField Ji[3], F, T;
  for (int j = 0; j < nD; ++j) {
    MG_TRY(mg_field_alloc(g, nD, &Ji[j]));
    MG_TRY(mg_grid_coordinate_derivatives(g, j, &Ji[j]));
  }
  bool anyPlane = false;
  for (int i = 0; i < nD; ++i) anyPlane = anyPlane || g->periodicityType[i] == MG_PERIODIC_PLANE;
  MetricArgs ma;
  std::memset(&ma, 0, sizeof(ma));
  for (int j = 0; j < nD; ++j)
    for (int i = 0; i < nD; ++i) ma.Ji[j * nD + i] = Ji[j].comp(i);
  for (int q = 0; q < nD * nD; ++q) ma.m[q] = g->metrics.comp(q);
  ma.jac = g->jacobian.comp(0);
  for (int i = 0; i < nD; ++i) ma.arc[i] = g->arcLengths.comp(i);
  ma.iblank = g->iblank;
  ma.N = N;
  ma.nD = nD;
  ma.curvilinear = g->isCurvilinear;
  ma.planeFormulas = anyPlane;
  { k_metrics<<<nblocks(N), 256, 0, st>>>(ma); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  if (nD == 3) {
    if (!anyPlane) {
      // conservative (curl) form, reference :905-1002: metrics(q) = D_a(Ji[x]*coord[y]) - D_b(Ji[z]*coord[w])
      MG_TRY(mg_field_alloc(g, 1, &F));
      MG_TRY(mg_field_alloc(g, 1, &T));
      struct Term { int q, d1, j1, c1, d2, j2, c2; };   // 0-based: derivative dir, Ji index, coordinate
      // Ji flat index q' = j*3 + i  (reference jacobianMatrixInverse(:, q'+1))
      static const Term terms[9] = {
          {0, 2, 4, 2, 1, 7, 2}, {1, 2, 5, 0, 1, 8, 0}, {2, 2, 3, 1, 1, 6, 1},
          {3, 0, 7, 2, 2, 1, 2}, {4, 0, 8, 0, 2, 2, 0}, {5, 0, 6, 1, 2, 0, 1},
          {6, 1, 1, 2, 0, 4, 2}, {7, 1, 2, 0, 0, 5, 0}, {8, 1, 0, 1, 0, 3, 1}};
      const bool diag[9] = {true, false, false, false, true, false, false, false, true};
      auto JiPtr = [&](int q) { return Ji[q / 3].comp(q % 3); };
      for (const Term& t : terms) {
        double* mq = g->metrics.comp(t.q);
        if (!g->isCurvilinear && !diag[t.q]) {
          { k_fill<<<nblocks(N), 256, 0, st>>>(mq, 0.0, N); mg_count_launches(1); }
          continue;
        }
        { k_mul<<<nblocks(N), 256, 0, st>>>(F.comp(0), JiPtr(t.j1), g->coordinates.comp(t.c1), N); mg_count_launches(1); }
        MG_TRY(mg_grid_apply(g, g->firstDerivative[t.d1], F.comp(0), F.compStride, mq, g->metrics.compStride, 1));
        if (g->isCurvilinear) {
          { k_mul<<<nblocks(N), 256, 0, st>>>(F.comp(0), JiPtr(t.j2), g->coordinates.comp(t.c2), N); mg_count_launches(1); }
          MG_TRY(mg_grid_apply(g, g->firstDerivative[t.d2], F.comp(0), F.compStride, T.comp(0), T.compStride, 1));
          { k_sub<<<nblocks(N), 256, 0, st>>>(mq, T.comp(0), N); mg_count_launches(1); }
        }
      }
      MG_CUDA(cudaGetLastError());
    }
    double* mp[9];
    for (int q = 0; q < 9; ++q) mp[q] = g->metrics.comp(q);
    double** d_mp = nullptr;
    MG_CUDA(cudaMalloc(&d_mp, sizeof(mp)));
    MG_CUDA(cudaMemcpyAsync(d_mp, mp, sizeof(mp), cudaMemcpyHostToDevice, st));
    { k_arc3<<<nblocks(N), 256, 0, st>>>(d_mp, g->arcLengths.comp(0), g->arcLengths.comp(1), g->arcLengths.comp(2),
                                       g->iblank, N, g->isCurvilinear); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
    MG_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_mp);
  }
  // norm = prod_dir H_dir * J, then jacobian <- 1/J (reference :1054-1063)
  { k_fill<<<nblocks(N), 256, 0, st>>>(g->norm.comp(0), 1.0, N); mg_count_launches(1); }
  for (int i = 0; i < nD; ++i)
    MG_TRY(mg_norm_launch(g->firstDerivative[i], g->norm.comp(0), g->norm.compStride, 1, g->localSize, 0, st));
  int* d_flag = nullptr;
  MG_CUDA(cudaMalloc(&d_flag, sizeof(int)));
  MG_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
  { k_norm_finish<<<nblocks(N), 256, 0, st>>>(g->norm.comp(0), g->jacobian.comp(0), N, d_flag); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  int flag = 0;
  MG_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  MG_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_flag);
  for (int j = 0; j < nD; ++j) mg_field_free(&Ji[j]);
  mg_field_free(&F);
  mg_field_free(&T);
  if (hasNegativeJacobian) *hasNegativeJacobian = flag;
  g->updated = true;
  return 0;
}

// computeGradient (reference :1172-1421): f has nComp components (stride fCs), out nD*nComp components.
int mg_grid_gradient_dev(mg_grid* g, const double* f, size_t fCs, int nComp, MgField* out, MgField* scratch) {
  const int nD = g->nD;
  if (scratch->nComp < nD * nComp || out->nComp < nD * nComp) MG_FAIL("gradient: scratch too small");
  GradArgs a;
  std::memset(&a, 0, sizeof(a));
  for (int i = 0; i < nD; ++i) {
    double* di = scratch->comp(i * nComp);
    MG_TRY(mg_grid_apply(g, g->firstDerivative[i], f, fCs, di, scratch->compStride, nComp));
    a.d[i] = di;
  }
  for (int q = 0; q < nD * nD; ++q) a.m[q] = g->metrics.comp(q);
  a.jac = g->jacobian.comp(0);
  a.out = out->comp(0);
  a.cs = out->compStride;
  a.N = g->N;
  a.nD = nD;
  a.nComp = nComp;
  a.curvilinear = g->isCurvilinear;
  if (scratch->compStride != out->compStride) MG_FAIL("gradient: stride mismatch");
  { k_gradient<<<nblocks(g->N), 256, 0, mg_stream()>>>(a); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return 0;
}

int mg_grid_inner_product_dev(mg_grid* g, const double* f, const double* gg, const double* weight, size_t cs,
                              int nComp, double* result) {
  const int blocks = 1024;
  static double* d_partial = nullptr;
  static double* h_partial = nullptr;
  if (!d_partial) {
    MG_CUDA(cudaMalloc(&d_partial, blocks * sizeof(double)));
    MG_CUDA(cudaMallocHost(&h_partial, blocks * sizeof(double)));
  }
  { k_inner<<<blocks, 256, 0, mg_stream()>>>(f, gg, g->norm.comp(0), weight, cs, nComp, g->N, d_partial); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  MG_CUDA(cudaMemcpyAsync(h_partial, d_partial, blocks * sizeof(double), cudaMemcpyDeviceToHost, mg_stream()));
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  // fixed-order pairwise sum on the host: deterministic and independent of launch timing
  int n = blocks;
  while (n > 1) {
    const int half = n / 2;
    for (int i = 0; i < half; ++i) h_partial[i] += h_partial[i + half];
    if (n & 1) h_partial[0] += h_partial[n - 1];
    n = half;
  }
  *result = h_partial[0];
  return 0;
}
