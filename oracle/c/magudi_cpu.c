/* CPU restatement in C + OpenMP of the reference's RHS / adjoint / RK4 path for patch-free configurations.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py): the product never links or calls this.
 * It is (i) a second, compiled restatement that the NumPy oracle is cross-checked against and (ii) the CPU arm
 * of bench.py ("cpu_baseline", "--impl reference"): the reference's Fortran/MPI build cannot be compiled in
 * this image (no Fortran compiler, no MPI), so its algorithm is restated here with the reference's own loop
 * structure - every operator application copies into a ghosted array, fills the ghost points, sweeps the
 * interior stencil and then the boundary closures (src/StencilOperatorImpl.f90:35-252); fluxes, Jacobians and
 * the RK4 axpys are separate full-array passes with (N, nComp) temporaries (src/RhsHelperImpl.f90:254-596,
 * src/RK4IntegratorImpl.f90:65-270) - and one OpenMP thread team over the whole domain in place of the MPI
 * ranks (no halo messages: favourable to the CPU side).  Scratch arrays are kept between calls (the reference
 * allocates and frees them inside every call: also favourable to the CPU side).  Parity unpinned against the
 * compiled reference (same status as the NumPy oracle).
 *
 * Arrays are Fortran column-major A(N, nComp): a[p + N*c], p = i + nx*(j + ny*k).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXI 9
#define MAXD 12
#define MAXW 16

enum { SYMMETRIC = 0, SKEW_SYMMETRIC = 1, ASYMMETRIC = 2 };
enum { FORWARD = 1, ADJOINT = -1 };

/* t_StencilOperator (include/StencilOperator.f90:9-16) */
typedef struct {
  int symmetryType, interiorWidth, boundaryWidth, boundaryDepth;
  int lo, nInterior;
  int nGhost[2], periodicOffset[2], hasDomainBoundary[2];
  double rhsInterior[MAXI];
  double normBoundary[MAXD];
  double rhsBoundary1[MAXD][MAXW]; /* [row m][column s] */
  double rhsBoundary2[MAXD][MAXW];
} cpu_op;

typedef struct {
  int nD, nU, n[3];
  long N;
  int curvilinear, viscous, dissipationOn, composite;
  double gamma, ReInv, PrInv, powerLaw, bulkRatio, dissipationAmount;
  cpu_op D[3], Dadj[3], Dd[3], Dt[3];
  double *metrics, *jacobian, *arc; /* (N, nD*nD), (N), (N, nD) */
  /* t_State */
  double *Q, *W, *rhs;
  double *v, *u, *p, *T, *mu, *lam, *kap, *tau, *q;
  /* RK4 buffers */
  double *b1, *b2;
  /* scratch */
  double *ghost;      /* ghosted copy for apply: (max extent) */
  double *f1, *f2;    /* (N, nU, nD) */
  double *t1;         /* (N, nU) */
  double *dxi;        /* (N, nD*nD) gradient scratch */
} cpu_ctx;

static double* dalloc(long n) {
  double* p = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  if (p) memset(p, 0, sizeof(double) * (size_t)(n > 0 ? n : 1));
  return p;
}

int cpu_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

cpu_ctx* cpu_create(int nD, const int* n, int curvilinear, int viscous, int dissipationOn, int composite,
                    double gamma, double ReInv, double PrInv, double powerLaw, double bulkRatio,
                    double dissipationAmount, const cpu_op* D, const cpu_op* Dadj, const cpu_op* Dd,
                    const cpu_op* Dt, const double* metrics, const double* jacobian, const double* arc) {
  cpu_ctx* c = (cpu_ctx*)calloc(1, sizeof(cpu_ctx));
  c->nD = nD;
  c->nU = nD + 2;
  c->N = 1;
  for (int d = 0; d < 3; ++d) { c->n[d] = d < nD ? n[d] : 1; c->N *= c->n[d]; }
  c->curvilinear = curvilinear; c->viscous = viscous; c->dissipationOn = dissipationOn; c->composite = composite;
  c->gamma = gamma; c->ReInv = ReInv; c->PrInv = PrInv; c->powerLaw = powerLaw; c->bulkRatio = bulkRatio;
  c->dissipationAmount = dissipationAmount;
  for (int d = 0; d < nD; ++d) {
    c->D[d] = D[d]; c->Dadj[d] = Dadj[d];
    if (dissipationOn) { c->Dd[d] = Dd[d]; if (!composite) c->Dt[d] = Dt[d]; }
  }
  const long N = c->N;
  const int nU = c->nU;
  c->metrics = dalloc(N * nD * nD); memcpy(c->metrics, metrics, sizeof(double) * N * nD * nD);
  c->jacobian = dalloc(N); memcpy(c->jacobian, jacobian, sizeof(double) * N);
  c->arc = dalloc(N * nD); memcpy(c->arc, arc, sizeof(double) * N * nD);
  c->Q = dalloc(N * nU); c->W = dalloc(N * nU); c->rhs = dalloc(N * nU);
  c->v = dalloc(N); c->u = dalloc(N * nD); c->p = dalloc(N); c->T = dalloc(N);
  c->mu = dalloc(N); c->lam = dalloc(N); c->kap = dalloc(N); c->tau = dalloc(N * nD * nD); c->q = dalloc(N * nD);
  c->b1 = dalloc(N * nU); c->b2 = dalloc(N * nU);
  long gmax = 0;
  for (int d = 0; d < nD; ++d) { long e = N / c->n[d] * (c->n[d] + 2 * (MAXI / 2)); if (e > gmax) gmax = e; }
  c->ghost = dalloc(gmax * (nD * nD > nU ? nD * nD : nU));
  c->f1 = dalloc(N * nU * nD); c->f2 = dalloc(N * nU * nD); c->t1 = dalloc(N * nU);
  c->dxi = dalloc(N * nD * nD);
  return c;
}

void cpu_destroy(cpu_ctx* c) {
  if (!c) return;
  double* all[] = {c->metrics, c->jacobian, c->arc, c->Q, c->W, c->rhs, c->v, c->u, c->p, c->T, c->mu, c->lam, c->kap,
                   c->tau, c->q, c->b1, c->b2, c->ghost, c->f1, c->f2, c->t1, c->dxi};
  for (unsigned i = 0; i < sizeof(all) / sizeof(all[0]); ++i) free(all[i]);
  free(c);
}

double* cpu_field(cpu_ctx* c, int id) {
  switch (id) {
    case 0: return c->Q; case 1: return c->W; case 2: return c->rhs; case 3: return c->v; case 4: return c->u;
    case 5: return c->p; case 6: return c->T; case 7: return c->mu; case 8: return c->lam; case 9: return c->kap;
    case 10: return c->tau; case 11: return c->q;
  }
  return 0;
}

/* applyOperator_{1,2,3} (src/StencilOperatorImpl.f90:35-252): in place on x(N, nComp), direction d.
 * (1) copy into the ghosted array, (2) fillGhostPoints (single rank: the periodic neighbour is this rank,
 * src/MPIHelperImpl.f90:113-389), (3) applyOperatorAtInteriorPoints (:254-457), (4) boundary closures. */
static void apply_operator(cpu_ctx* c, const cpu_op* op, double* x, int nComp, int d) {
  const int nd = c->n[d];
  long inner = 1, outer = 1;
  for (int e = 0; e < d; ++e) inner *= c->n[e];
  for (int e = d + 1; e < 3; ++e) outer *= c->n[e];
  const int g1 = op->nGhost[0], g2 = op->nGhost[1];
  const int gn = nd + g1 + g2;
  const long N = c->N, GN = inner * gn * outer;
  double* xg = c->ghost;
  const int o1 = op->periodicOffset[0], o2 = op->periodicOffset[1];
  const int fill = (g1 > 0 || g2 > 0) && g1 == g2;   /* one-sided ghosts on one rank: early return (:158) */
  /* (1) + (2) */
#pragma omp parallel for collapse(2) schedule(static)
  for (int cc = 0; cc < nComp; ++cc)
    for (long o = 0; o < outer; ++o) {
      const double* xs = x + N * cc + inner * nd * o;
      double* gs = xg + GN * cc + inner * gn * o;
      memcpy(gs + inner * g1, xs, sizeof(double) * inner * nd);
      if (fill) {
        for (int t = 0; t < g1; ++t)
          memcpy(gs + inner * t, gs + inner * (g1 + nd - g2 - o1 + t), sizeof(double) * inner);
        for (int t = 0; t < g2; ++t)
          memcpy(gs + inner * (g1 + nd + t), gs + inner * (g1 + o2 + t), sizeof(double) * inner);
      } else {
        for (int t = 0; t < g1; ++t) memset(gs + inner * t, 0, sizeof(double) * inner);
        for (int t = 0; t < g2; ++t) memset(gs + inner * (g1 + nd + t), 0, sizeof(double) * inner);
      }
    }
  /* (3) interior */
  int is = 0, ie = nd;
  if (g1 == 0) is += op->boundaryDepth;
  if (g2 == 0) ie -= op->boundaryDepth;
  const int h = op->interiorWidth / 2;
  const double* ci = op->rhsInterior;
  const int lo = op->lo;
  if (ie > is) {
    if (inner == 1) {
#pragma omp parallel for collapse(2) schedule(static)
      for (int cc = 0; cc < nComp; ++cc)
        for (long o = 0; o < outer; ++o) {
          double* xs = x + N * cc + (long)nd * o;
          const double* gs = xg + GN * cc + (long)gn * o + g1;
          if (op->symmetryType == SKEW_SYMMETRIC) {
            for (int i = is; i < ie; ++i) {
              double r = 0.0;
              for (int m = 1; m <= h; ++m) r += ci[m - lo] * (gs[i + m] - gs[i - m]);
              xs[i] = r;
            }
          } else if (op->symmetryType == SYMMETRIC) {
            for (int i = is; i < ie; ++i) {
              double r = 0.0;
              for (int m = 1; m <= h; ++m) r += ci[m - lo] * (gs[i + m] + gs[i - m]);
              xs[i] = r + ci[0 - lo] * gs[i];
            }
          } else {
            for (int i = is; i < ie; ++i) {
              double r = 0.0;
              for (int m = 0; m < op->nInterior; ++m) r += ci[m] * gs[i + lo + m];
              xs[i] = r;
            }
          }
        }
    } else {
#pragma omp parallel for collapse(3) schedule(static)
      for (int cc = 0; cc < nComp; ++cc)
        for (long o = 0; o < outer; ++o)
          for (int i = is; i < ie; ++i) {
            double* xs = x + N * cc + inner * ((long)nd * o + i);
            const double* gs = xg + GN * cc + inner * ((long)gn * o + i + g1);
            if (op->symmetryType == SKEW_SYMMETRIC) {
              for (long t = 0; t < inner; ++t) {
                double r = 0.0;
                for (int m = 1; m <= h; ++m) r += ci[m - lo] * (gs[t + inner * m] - gs[t - inner * m]);
                xs[t] = r;
              }
            } else if (op->symmetryType == SYMMETRIC) {
              for (long t = 0; t < inner; ++t) {
                double r = 0.0;
                for (int m = 1; m <= h; ++m) r += ci[m - lo] * (gs[t + inner * m] + gs[t - inner * m]);
                xs[t] = r + ci[0 - lo] * gs[t];
              }
            } else {
              for (long t = 0; t < inner; ++t) {
                double r = 0.0;
                for (int m = 0; m < op->nInterior; ++m) r += ci[m] * gs[t + inner * (lo + m)];
                xs[t] = r;
              }
            }
          }
    }
  }
  /* (4) closures (:73-102) */
  for (int side = 0; side < 2; ++side) {
    if (!op->hasDomainBoundary[side]) continue;
    const int depth = op->boundaryDepth, width = op->boundaryWidth;
#pragma omp parallel for collapse(2) schedule(static)
    for (int cc = 0; cc < nComp; ++cc)
      for (long o = 0; o < outer; ++o)
        for (int m = 0; m < depth; ++m) {
          const int row = side == 0 ? m : nd - 1 - m;
          double* xs = x + N * cc + inner * ((long)nd * o + row);
          for (long t = 0; t < inner; ++t) {
            double r = 0.0;
            for (int s = 0; s < width; ++s) {
              const int col = side == 0 ? s : nd - width + s;
              const double b = side == 0 ? op->rhsBoundary1[m][s] : op->rhsBoundary2[m][s];
              r += b * xg[GN * cc + inner * ((long)gn * o + col + g1) + t];
            }
            xs[t] = r;
          }
        }
  }
}

/* applyOperatorNormInverse_{1,2,3} (src/StencilOperatorImpl.f90:973-1107) */
static void apply_norm_inverse(cpu_ctx* c, const cpu_op* op, double* x, int nComp, int d) {
  const int nd = c->n[d];
  long inner = 1, outer = 1;
  for (int e = 0; e < d; ++e) inner *= c->n[e];
  for (int e = d + 1; e < 3; ++e) outer *= c->n[e];
  for (int side = 0; side < 2; ++side) {
    if (!op->hasDomainBoundary[side]) continue;
#pragma omp parallel for collapse(2) schedule(static)
    for (int cc = 0; cc < nComp; ++cc)
      for (long o = 0; o < outer; ++o)
        for (int m = 0; m < op->boundaryDepth; ++m) {
          const int row = side == 0 ? m : nd - 1 - m;
          double* xs = x + c->N * cc + inner * ((long)nd * o + row);
          for (long t = 0; t < inner; ++t) xs[t] /= op->normBoundary[m];
        }
  }
}

/* computeGradientOfScalar/Vector (src/GridImpl.f90:1172-1421): out(:, j + nD*c) = d f_c / d x_j */
static void compute_gradient(cpu_ctx* c, const double* f, int nc, double* out) {
  const long N = c->N;
  const int nD = c->nD;
  double* dxi = c->dxi; /* (N, nc, nD) */
  for (int i = 0; i < nD; ++i) {
    memcpy(dxi + N * nc * i, f, sizeof(double) * N * nc);
    apply_operator(c, &c->D[i], dxi + N * nc * i, nc, i);
  }
#pragma omp parallel for schedule(static)
  for (long p = 0; p < N; ++p)
    for (int cc = 0; cc < nc; ++cc)
      for (int j = 0; j < nD; ++j) {
        double acc;
        if (c->curvilinear) {
          acc = c->metrics[p + N * (j + nD * 0)] * dxi[p + N * (cc + nc * 0)];
          for (int i = 1; i < nD; ++i) acc += c->metrics[p + N * (j + nD * i)] * dxi[p + N * (cc + nc * i)];
        } else {
          acc = c->metrics[p + N * (j + nD * j)] * dxi[p + N * (cc + nc * j)];
        }
        out[p + N * (j + nD * cc)] = c->jacobian[p] * acc;
      }
}

/* updateState (src/StateImpl.f90:466-537): computeDependentVariables (src/CNSHelperImpl.f90:3-87),
 * computeTransportVariables (:89-177), velocity gradient -> computeStressTensor (:412-452), heat flux */
void cpu_update_state(cpu_ctx* c) {
  const long N = c->N;
  const int nD = c->nD;
  const double g = c->gamma;
#pragma omp parallel for schedule(static)
  for (long p = 0; p < N; ++p) {
    const double v = 1.0 / c->Q[p];
    double usq = 0.0;
    for (int i = 0; i < nD; ++i) {
      const double ui = v * c->Q[p + N * (i + 1)];
      c->u[p + N * i] = ui;
      usq += ui * ui;
    }
    const double pr = (g - 1.0) * (c->Q[p + N * (nD + 1)] - 0.5 * c->Q[p] * usq);
    c->v[p] = v;
    c->p[p] = pr;
    c->T[p] = g * pr / (g - 1.0) * v;
    if (c->viscous) {
      if (c->powerLaw <= 0.0) {
        c->mu[p] = c->ReInv;
        c->lam[p] = (c->bulkRatio - 2.0 / 3.0) * c->ReInv;
        c->kap[p] = c->ReInv * c->PrInv;
      } else {
        c->mu[p] = pow((g - 1.0) * c->T[p], c->powerLaw) * c->ReInv;
        c->lam[p] = (c->bulkRatio - 2.0 / 3.0) * c->mu[p];
        c->kap[p] = c->mu[p] * c->PrInv;
      }
    }
  }
  if (!c->viscous) return;
  compute_gradient(c, c->u, nD, c->tau);
#pragma omp parallel for schedule(static)
  for (long p = 0; p < N; ++p) {
    double gr[9], s[9];
    const double mu = c->mu[p], lam = c->lam[p];
    for (int e = 0; e < nD * nD; ++e) gr[e] = c->tau[p + N * e];
    if (nD == 1) {
      s[0] = (2.0 * mu + lam) * gr[0];
    } else if (nD == 2) {
      const double div = lam * (gr[0] + gr[3]);
      s[0] = 2.0 * mu * gr[0] + div; s[1] = mu * (gr[1] + gr[2]); s[2] = s[1]; s[3] = 2.0 * mu * gr[3] + div;
    } else {
      const double div = lam * (gr[0] + gr[4] + gr[8]);
      s[0] = 2.0 * mu * gr[0] + div; s[1] = mu * (gr[1] + gr[3]); s[2] = mu * (gr[2] + gr[6]);
      s[3] = s[1]; s[4] = 2.0 * mu * gr[4] + div; s[5] = mu * (gr[5] + gr[7]);
      s[6] = s[2]; s[7] = s[5]; s[8] = 2.0 * mu * gr[8] + div;
    }
    for (int e = 0; e < nD * nD; ++e) c->tau[p + N * e] = s[e];
  }
  compute_gradient(c, c->T, 1, c->q);
#pragma omp parallel for schedule(static)
  for (long p = 0; p < N; ++p)
    for (int i = 0; i < nD; ++i) c->q[p + N * i] = -c->kap[p] * c->q[p + N * i];
}

/* addDissipation (src/RhsHelperImpl.f90:10-87) */
static void add_dissipation(cpu_ctx* c, int mode) {
  if (!c->dissipationOn) return;
  const long N = c->N;
  const int nD = c->nD, nU = c->nU;
  const double amount = mode == ADJOINT ? -c->dissipationAmount : c->dissipationAmount;
  const double* src = mode == FORWARD ? c->Q : c->W;
  double* t = c->t1;
  for (int i = 0; i < nD; ++i) {
    memcpy(t, src, sizeof(double) * N * nU);
    apply_operator(c, &c->Dd[i], t, nU, i);
    if (!c->composite) {
#pragma omp parallel for schedule(static)
      for (long p = 0; p < N; ++p)
        for (int cc = 0; cc < nU; ++cc) t[p + N * cc] = -c->arc[p + N * i] * t[p + N * cc];
      apply_operator(c, &c->Dt[i], t, nU, i);
      apply_norm_inverse(c, &c->D[i], t, nU, i);
    }
#pragma omp parallel for schedule(static)
    for (long p = 0; p < N; ++p)
      for (int cc = 0; cc < nU; ++cc) c->rhs[p + N * cc] += amount * t[p + N * cc];
  }
}

/* computeRhsForward (src/RhsHelperImpl.f90:254-354) */
static void rhs_forward(cpu_ctx* c) {
  const long N = c->N;
  const int nD = c->nD, nU = c->nU;
  memset(c->rhs, 0, sizeof(double) * N * nU);
  double *f1 = c->f1, *f2 = c->f2; /* (N, nU, nD): index p + N*(cc + nU*l) */
  /* computeCartesianInviscidFluxes (src/CNSHelperImpl.f90:563-619) */
#pragma omp parallel for schedule(static)
  for (long p = 0; p < N; ++p)
    for (int l = 0; l < nD; ++l) {
      f1[p + N * (0 + nU * l)] = c->Q[p + N * (l + 1)];
      for (int cc = 0; cc < nD; ++cc) {
        if (cc == l) f1[p + N * (cc + 1 + nU * l)] = c->Q[p + N * (l + 1)] * c->u[p + N * l] + c->p[p];
        else {
          const int lo = cc < l ? cc : l, hi = cc < l ? l : cc;
          f1[p + N * (cc + 1 + nU * l)] = c->Q[p + N * (lo + 1)] * c->u[p + N * hi];
        }
      }
      f1[p + N * (nD + 1 + nU * l)] = c->u[p + N * l] * (c->Q[p + N * (nD + 1)] + c->p[p]);
    }
  if (c->viscous) {
    /* computeCartesianViscousFluxes (:621-689); fluxes1 -= fluxes2 */
#pragma omp parallel for schedule(static)
    for (long p = 0; p < N; ++p)
      for (int l = 0; l < nD; ++l) {
        double acc = 0.0;
        f2[p + N * (0 + nU * l)] = 0.0;
        for (int cc = 0; cc < nD; ++cc) {
          const double t = c->tau[p + N * (l + nD * cc)];
          f2[p + N * (cc + 1 + nU * l)] = t;
          acc = cc == 0 ? c->u[p] * t : acc + c->u[p + N * cc] * t;
        }
        f2[p + N * (nD + 1 + nU * l)] = acc - c->q[p + N * l];
      }
#pragma omp parallel for schedule(static)
    for (long e = 0; e < N * nU * nD; ++e) f1[e] -= f2[e];
  }
  /* transformFluxes (:772-840) */
#pragma omp parallel for schedule(static)
  for (long p = 0; p < N; ++p)
    for (int i = 0; i < nD; ++i)
      for (int cc = 0; cc < nU; ++cc) {
        double acc;
        if (c->curvilinear) {
          acc = c->metrics[p + N * (0 + nD * i)] * f1[p + N * (cc + nU * 0)];
          for (int j = 1; j < nD; ++j) acc += c->metrics[p + N * (j + nD * i)] * f1[p + N * (cc + nU * j)];
        } else {
          acc = c->metrics[p + N * (i + nD * i)] * f1[p + N * (cc + nU * i)];
        }
        f2[p + N * (cc + nU * i)] = acc;
      }
  for (int i = 0; i < nD; ++i) apply_operator(c, &c->D[i], f2 + N * nU * i, nU, i);
#pragma omp parallel for schedule(static)
  for (long p = 0; p < N; ++p)
    for (int cc = 0; cc < nU; ++cc) {
      double s = f2[p + N * cc];
      for (int i = 1; i < nD; ++i) s += f2[p + N * (cc + nU * i)];
      c->rhs[p + N * cc] -= s;
    }
  add_dissipation(c, FORWARD);
}

/* computeJacobianOfInviscidFlux{1,2,3}D (src/CNSHelperImpl.f90:984-1444) minus
 * computeFirstPartialViscousJacobian{1,2,3}D (:2344-2600): A[i][j] at one point for metrics m */
static void flux_jacobian(const cpu_ctx* c, long p, const double* m, double A[5][5]) {
  const long N = c->N;
  const int nD = c->nD, nU = c->nU;
  const double g = c->gamma, v = c->v[p], T = c->T[p];
  double u[3], uh = 0.0, usq = 0.0;
  for (int i = 0; i < nD; ++i) { u[i] = c->u[p + N * i]; uh += m[i] * u[i]; usq += u[i] * u[i]; }
  const double phi2 = 0.5 * (g - 1.0) * usq;
  for (int i = 0; i < nU; ++i) for (int j = 0; j < nU; ++j) A[i][j] = 0.0;
  for (int a = 0; a < nD; ++a) A[a + 1][0] = phi2 * m[a] - uh * u[a];
  A[nU - 1][0] = uh * ((g - 2.0) / (g - 1.0) * phi2 - T);
  for (int b = 0; b < nD; ++b) {
    A[0][b + 1] = m[b];
    for (int a = 0; a < nD; ++a)
      A[a + 1][b + 1] = a == b ? uh - (g - 2.0) * u[a] * m[a] : u[a] * m[b] - (g - 1.0) * u[b] * m[a];
    A[nU - 1][b + 1] = (T + phi2 / (g - 1.0)) * m[b] - (g - 1.0) * uh * u[b];
  }
  for (int a = 0; a < nD; ++a) A[a + 1][nU - 1] = (g - 1.0) * m[a];
  A[nU - 1][nU - 1] = g * uh;
  if (!c->viscous) return;
  double cst[3], chf = 0.0, ucst = 0.0;
  for (int cc = 0; cc < nD; ++cc) {
    double acc = 0.0;
    for (int l = 0; l < nD; ++l) acc += m[l] * c->tau[p + N * (l + nD * cc)];
    cst[cc] = acc;
    ucst += u[cc] * acc;
  }
  for (int l = 0; l < nD; ++l) chf += m[l] * c->q[p + N * l];
  const double temp1 = ucst - chf;
  double temp2 = c->powerLaw * g * v / T * (phi2 / (g - 1.0) - T / g);
  for (int cc = 0; cc < nD; ++cc) A[cc + 1][0] -= temp2 * cst[cc];
  A[nU - 1][0] -= temp2 * temp1 - v * ucst;
  for (int b = 0; b < nD; ++b) {
    temp2 = -c->powerLaw * g * v / T * u[b];
    for (int cc = 0; cc < nD; ++cc) A[cc + 1][b + 1] -= temp2 * cst[cc];
    A[nU - 1][b + 1] -= temp2 * temp1 + v * cst[b];
  }
  temp2 = c->powerLaw * g * v / T;
  for (int cc = 0; cc < nD; ++cc) A[cc + 1][nU - 1] -= temp2 * cst[cc];
  A[nU - 1][nU - 1] -= temp2 * temp1;
}

/* computeSecondPartialViscousJacobian{1,2,3}D (src/CNSHelperImpl.f90:2602-2756), x 1/J */
static void second_partial(const cpu_ctx* c, long p, const double* m1, const double* m2, double B[4][4]) {
  const long N = c->N;
  const int nD = c->nD;
  const double mu = c->mu[p], lam = c->lam[p], kap = c->kap[p], jac = c->jacobian[p];
  double u[3], temp1 = 0.0, d1 = 0.0, d2 = 0.0;
  for (int i = 0; i < nD; ++i) { u[i] = c->u[p + N * i]; temp1 += m1[i] * m2[i]; d2 += m2[i] * u[i]; d1 += m1[i] * u[i]; }
  const double temp2 = mu * d2, temp3 = lam * d1;
  for (int a = 0; a <= nD; ++a) for (int b = 0; b <= nD; ++b) B[a][b] = 0.0;
  for (int a = 0; a < nD; ++a)
    for (int b = 0; b < nD; ++b)
      B[a][b] = jac * (a == b ? mu * temp1 + (mu + lam) * m1[a] * m2[a] : mu * m1[b] * m2[a] + lam * m1[a] * m2[b]);
  for (int b = 0; b < nD; ++b) B[nD][b] = jac * (mu * temp1 * u[b] + m1[b] * temp2 + m2[b] * temp3);
  B[nD][nD] = jac * (kap * temp1);
}

/* computeRhsAdjoint (src/RhsHelperImpl.f90:356-596), patch-free */
static void rhs_adjoint(cpu_ctx* c) {
  const long N = c->N;
  const int nD = c->nD, nU = c->nU;
  memset(c->rhs, 0, sizeof(double) * N * nU);
  double* temp1 = c->f1; /* (N, nU, nD) */
  for (int i = 0; i < nD; ++i) {
    memcpy(temp1 + N * nU * i, c->W, sizeof(double) * N * nU);
    apply_operator(c, &c->Dadj[i], temp1 + N * nU * i, nU, i);
  }
#pragma omp parallel for schedule(static)
  for (long p = 0; p < N; ++p) {
    double A[5][5], m[3];
    for (int i = 0; i < nD; ++i) {
      for (int l = 0; l < nD; ++l) m[l] = c->metrics[p + N * (l + nD * i)];
      flux_jacobian(c, p, m, A);
      for (int j = 0; j < nU; ++j) {
        double acc = 0.0;
        for (int r = 0; r < nU; ++r) acc += A[r][j] * temp1[p + N * (r + nU * i)];
        c->rhs[p + N * j] += acc;
      }
    }
  }
  if (c->viscous) {
    double* diff = c->f2; /* (N, nU-1, nD) */
    const int nG = nU - 1;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < N; ++p) {
      double B[4][4], m1[3], m2[3];
      for (int j = 0; j < nD; ++j) {
        double acc[4] = {0, 0, 0, 0};
        for (int l = 0; l < nD; ++l) m2[l] = c->metrics[p + N * (l + nD * j)];
        for (int i = 0; i < nD; ++i) {
          for (int l = 0; l < nD; ++l) m1[l] = c->metrics[p + N * (l + nD * i)];
          second_partial(c, p, m1, m2, B);
          for (int b = 0; b < nG; ++b)
            for (int a = 0; a < nG; ++a) acc[b] += B[a][b] * temp1[p + N * (a + 1 + nU * i)];
        }
        for (int b = 0; b < nG; ++b) diff[p + N * (b + nG * j)] = acc[b];
      }
    }
    for (int j = 0; j < nD; ++j) apply_operator(c, &c->Dadj[j], diff + N * nG * j, nG, j);
    const double g = c->gamma;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < N; ++p) {
      double t[4];
      for (int b = 0; b < nG; ++b) {
        double s = diff[p + N * b];
        for (int j = 1; j < nD; ++j) s += diff[p + N * (b + nG * j)];
        t[b] = s;
      }
      const double v = c->v[p];
      t[nD] = g * v * t[nD];
      double ut = 0.0;
      for (int i = 0; i < nD; ++i) {
        t[i] = v * t[i] - c->u[p + N * i] * t[nD];
        ut += c->u[p + N * i] * t[i];
      }
      for (int b = 0; b < nG; ++b) c->rhs[p + N * (b + 1)] -= t[b];
      c->rhs[p] += v * c->Q[p + N * (nD + 1)] * t[nD] + ut;
    }
  }
  add_dissipation(c, ADJOINT);
}

/* computeRhs (src/RegionImpl.f90:1877-2027), one grid, no patches: bulk RHS, then x (1/J) */
void cpu_compute_rhs(cpu_ctx* c, int mode) {
  if (mode == FORWARD) rhs_forward(c); else rhs_adjoint(c);
  const long N = c->N;
#pragma omp parallel for schedule(static)
  for (long p = 0; p < N; ++p)
    for (int cc = 0; cc < c->nU; ++cc) c->rhs[p + N * cc] *= c->jacobian[p];
}

/* substepForwardRK4 / substepAdjointRK4 (src/RK4IntegratorImpl.f90:65-270); the caller issues the state
 * update after each forward substep (src/SolverImpl.f90:831-834) */
void cpu_rk4_substep(cpu_ctx* c, int mode, int stage, double dt) {
  const long n = c->N * c->nU;
  double* X = mode == FORWARD ? c->Q : c->W;
  const int st = mode == FORWARD ? stage : 5 - stage;
  const double h = mode == FORWARD ? dt : -dt;
  if (st == 1) memcpy(c->b1, X, sizeof(double) * n);
  cpu_compute_rhs(c, mode);
  const double* R = c->rhs;
  if (st == 1) {
#pragma omp parallel for schedule(static)
    for (long e = 0; e < n; ++e) { c->b2[e] = X[e] + h * R[e] / 6.0; X[e] = c->b1[e] + h * R[e] / 2.0; }
  } else if (st == 2) {
#pragma omp parallel for schedule(static)
    for (long e = 0; e < n; ++e) { c->b2[e] = c->b2[e] + h * R[e] / 3.0; X[e] = c->b1[e] + h * R[e] / 2.0; }
  } else if (st == 3) {
#pragma omp parallel for schedule(static)
    for (long e = 0; e < n; ++e) { c->b2[e] = c->b2[e] + h * R[e] / 3.0; X[e] = c->b1[e] + h * R[e]; }
  } else {
#pragma omp parallel for schedule(static)
    for (long e = 0; e < n; ++e) X[e] = c->b2[e] + h * R[e] / 6.0;
  }
}

/* One step of the benchmark: forward RK4 step storing the 4 substep states (the reference's checkpointer keeps
 * them in a RAM buffer, src/UniformCheckpointerImpl.f90:78-208), then one adjoint RK4 step that restores and
 * updates each stored state before its stage (src/SolverImpl.f90:1147-1195).  store: (4, N, nU) scratch. */
void cpu_forward_adjoint_step(cpu_ctx* c, double dt, double* store) {
  const long n = c->N * c->nU;
  for (int stage = 1; stage <= 4; ++stage) {
    memcpy(store + n * (stage - 1), c->Q, sizeof(double) * n);
    cpu_rk4_substep(c, FORWARD, stage, dt);
    cpu_update_state(c);
  }
  for (int stage = 4; stage >= 1; --stage) {
    memcpy(c->Q, store + n * (stage - 1), sizeof(double) * n);
    cpu_update_state(c);
    cpu_rk4_substep(c, ADJOINT, stage, dt);
  }
}
