// t_State on the device and the GENERAL (operator-by-operator) right-hand-side path: a direct GPU
// restatement of the reference call structure, valid for every scheme / boundary / dimension.
// The fused hot path lives in rhs_fused.cu; this path is its fallback for configurations the fused
// kernels do not cover yet and its device-side cross-check.
//
// Reference: src/StateImpl.f90:466-537 (updateState), src/RhsHelperImpl.f90:10-87 (addDissipation),
// :254-354 (computeRhsForward), :356-596 (computeRhsAdjoint), src/RegionImpl.f90:1877-2027 (computeRhs),
// src/RK4IntegratorImpl.f90:65-270.
#include <cstring>

#include "grid.h"
#include <unordered_set>

#include "patches.h"
#include "stencil_apply.h"
#include "rhs_fused.h"

namespace {

inline unsigned nblocks(size_t n) { return (unsigned)((n + 255) / 256); }

struct PtrSet {
  const double* Q; size_t csQ;
  double *v, *u, *p, *T, *mu, *lam, *kap;
  size_t cs;
  size_t N;
};

template <int ND>
__global__ void k_dependent(PtrSet a, PhysParams pp) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  double Q[ND + 2];
#pragma unroll
  for (int c = 0; c < ND + 2; ++c) Q[c] = a.Q[(size_t)c * a.csQ + p];
  Prim<ND> s;
  dependent<ND>(Q, pp.gamma, s);
  a.v[p] = s.v;
#pragma unroll
  for (int i = 0; i < ND; ++i) a.u[(size_t)i * a.cs + p] = s.u[i];
  a.p[p] = s.p;
  a.T[p] = s.T;
  if (pp.viscous) {
    double mu, lam, kap;
    transport(s.T, pp, mu, lam, kap);
    a.mu[p] = mu;
    a.lam[p] = lam;
    a.kap[p] = kap;
  }
}

template <int ND>
__global__ void k_stress(double* g, size_t cs, const double* mu, const double* lam, size_t N) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= N) return;
  double gr[ND * ND], s[ND * ND];
#pragma unroll
  for (int q = 0; q < ND * ND; ++q) gr[q] = g[(size_t)q * cs + p];
  stress_from_gradient<ND>(gr, mu[p], lam[p], s);
#pragma unroll
  for (int q = 0; q < ND * ND; ++q) g[(size_t)q * cs + p] = s[q];
}

__global__ void k_heatflux(double* q, size_t cs, int nD, const double* kap, size_t N) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= N) return;
  for (int i = 0; i < nD; ++i) q[(size_t)i * cs + p] = -kap[p] * q[(size_t)i * cs + p];
}

struct FluxArgs {
  const double *Q, *u, *pr, *tau, *q, *m;
  double* Fhat;      // component c + NU*i  (flux direction i)
  double* Fv;        // Cartesian viscous flux, component c + NU*l (may be null)
  size_t csQ, cs, N;
  int viscous, curvilinear;
};

// computeCartesianInviscid/ViscousFluxes + transformFluxes (reference CNSHelperImpl.f90:563-689, :772-840)
template <int ND>
__global__ void k_flux(FluxArgs a) {
  constexpr int NU = ND + 2;
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  double Q[NU], tau[ND * ND], q[ND], M[ND * ND];
  Prim<ND> s;
#pragma unroll
  for (int c = 0; c < NU; ++c) Q[c] = a.Q[(size_t)c * a.csQ + p];
#pragma unroll
  for (int i = 0; i < ND; ++i) s.u[i] = a.u[(size_t)i * a.cs + p];
  s.p = a.pr[p];
  if (a.viscous) {
#pragma unroll
    for (int c = 0; c < ND * ND; ++c) tau[c] = a.tau[(size_t)c * a.cs + p];
#pragma unroll
    for (int i = 0; i < ND; ++i) q[i] = a.q[(size_t)i * a.cs + p];
  }
#pragma unroll
  for (int c = 0; c < ND * ND; ++c) M[c] = a.m[(size_t)c * a.cs + p];
  double F[ND][NU], Fv[NU];
#pragma unroll
  for (int l = 0; l < ND; ++l) {
    cartesian_flux<ND>(l, Q, s, a.viscous, tau, q, F[l], Fv);
    if (a.Fv && a.viscous) {
#pragma unroll
      for (int c = 0; c < NU; ++c) a.Fv[(size_t)(c + NU * l) * a.cs + p] = Fv[c];
    }
  }
#pragma unroll
  for (int i = 0; i < ND; ++i)
#pragma unroll
    for (int c = 0; c < NU; ++c) {
      double r;
      if (a.curvilinear) {
        r = M[0 + ND * i] * F[0][c];
#pragma unroll
        for (int j = 1; j < ND; ++j) r += M[j + ND * i] * F[j][c];
      } else {
        r = M[i + ND * i] * F[i][c];
      }
      a.Fhat[(size_t)(c + NU * i) * a.cs + p] = r;
    }
}

// rhs = -(d0 + d1 + d2)   (reference RhsHelperImpl.f90:344)
__global__ void k_neg_sum(double* rhs, const double* d, size_t cs, int nU, int nD, size_t N) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= N) return;
  for (int c = 0; c < nU; ++c) {
    double s = d[(size_t)c * cs + p];
    for (int i = 1; i < nD; ++i) s += d[(size_t)(c + nU * i) * cs + p];
    rhs[(size_t)c * cs + p] = 0.0 - s;
  }
}

__global__ void k_scale_by(double* t, size_t cs, int nComp, const double* a, double sign, size_t N) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= N) return;
  for (int c = 0; c < nComp; ++c) t[(size_t)c * cs + p] = sign * a[p] * t[(size_t)c * cs + p];
}

__global__ void k_axpy(double* y, const double* x, size_t cs, size_t csx, int nComp, double a, size_t N) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= N) return;
  for (int c = 0; c < nComp; ++c) y[(size_t)c * cs + p] += a * x[(size_t)c * csx + p];
}

struct AdjArgs {
  const double *Q, *W, *v, *u, *T, *tau, *q, *mu, *lam, *kap, *m, *jac;
  const double* dW;   // adjoint derivative: component c + NU*i
  double* rhs;
  double* diff;       // adjoint diffusion: component c + (NU-1)*j
  size_t csQ, cs, N;
  int viscous, curvilinear;
  double gamma, powerLaw;
};

// Pointwise part of computeRhsAdjoint (reference RhsHelperImpl.f90:431-553)
template <int ND>
__global__ void k_adjoint_point(AdjArgs a) {
  constexpr int NU = ND + 2;
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  double Q[NU], tau[ND * ND], q[ND], M[ND * ND];
  Prim<ND> s;
#pragma unroll
  for (int c = 0; c < NU; ++c) Q[c] = a.Q[(size_t)c * a.csQ + p];
  s.v = a.v[p];
  s.T = a.T[p];
#pragma unroll
  for (int i = 0; i < ND; ++i) s.u[i] = a.u[(size_t)i * a.cs + p];
#pragma unroll
  for (int c = 0; c < ND * ND; ++c) M[c] = a.m[(size_t)c * a.cs + p];
  if (a.viscous) {
#pragma unroll
    for (int c = 0; c < ND * ND; ++c) tau[c] = a.tau[(size_t)c * a.cs + p];
#pragma unroll
    for (int i = 0; i < ND; ++i) q[i] = a.q[(size_t)i * a.cs + p];
  }
  double dW[ND][NU];
#pragma unroll
  for (int i = 0; i < ND; ++i)
#pragma unroll
    for (int c = 0; c < NU; ++c) dW[i][c] = a.dW[(size_t)(c + NU * i) * a.cs + p];
  double r[NU];
#pragma unroll
  for (int c = 0; c < NU; ++c) r[c] = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i)
    add_flux_jacobian_transpose<ND>(Q, s, &M[ND * i], a.gamma, a.viscous, a.powerLaw, tau, q, dW[i], r);
#pragma unroll
  for (int c = 0; c < NU; ++c) a.rhs[(size_t)c * a.cs + p] = r[c];
  if (a.viscous) {
    const double mu = a.mu[p], lam = a.lam[p], kap = a.kap[p], jac = a.jac[p];
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      double d[ND + 1];
#pragma unroll
      for (int c = 0; c < ND + 1; ++c) d[c] = 0.0;
#pragma unroll
      for (int i = 0; i < ND; ++i)
        add_second_partial_transpose<ND>(s.u, mu, lam, kap, jac, &M[ND * i], &M[ND * j], &dW[i][1], d);
#pragma unroll
      for (int c = 0; c < ND + 1; ++c) a.diff[(size_t)(c + (NU - 1) * j) * a.cs + p] = d[c];
    }
  }
}

struct LinArgs {
  const double *Q, *W, *v, *u, *T, *tau, *q, *mu, *lam, *kap, *m, *jac;
  const double* dT;   // derivatives of the primitive-like perturbation: component c + (NU-1)*j
  double* t;          // phase 1 output: the perturbation, NU-1 components
  double* Fhat;       // phase 2 output: linearized contravariant total flux, component c + NU*i
  double* Fv;         // linearized contravariant viscous flux, component c + NU*i (may be null)
  size_t csQ, csW, cs, N;
  int viscous;
  double gamma, powerLaw;
};

// computeRhsLinearized, perturbation of (u, T) up to the factors absorbed in the second-partial Jacobians
// (reference src/RhsHelperImpl.f90:692-714); the perturbation dQ lives in the adjoint variables.
template <int ND>
__global__ void k_linearized_primitive(LinArgs a) {
  constexpr int NU = ND + 2;
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  double dq[NU], u[ND], t[NU - 1];
#pragma unroll
  for (int c = 0; c < NU; ++c) dq[c] = a.W[(size_t)c * a.csW + p];
#pragma unroll
  for (int i = 0; i < ND; ++i) u[i] = a.u[(size_t)i * a.cs + p];
  const double v = a.v[p];
  double ut = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    t[i] = -u[i] * dq[0] + dq[i + 1];
    ut = (i == 0) ? u[0] * t[0] : ut + u[i] * t[i];
  }
  double e = -v * a.Q[(size_t)(NU - 1) * a.csQ + p] * dq[0];
  e = e - ut + dq[NU - 1];
  t[NU - 2] = e * a.gamma;
#pragma unroll
  for (int c = 0; c < NU - 1; ++c) a.t[(size_t)c * a.cs + p] = t[c] * v;
}

// Linearized contravariant fluxes (reference :660-690 inviscid, :716-790 viscous): Fhat_i = A_i dQ - B1_i dQ
// - [0; sum_j B2(m_i, m_j) d_j t]
template <int ND>
__global__ void k_linearized_flux(LinArgs a) {
  constexpr int NU = ND + 2;
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  double dq[NU], tau[ND * ND], q[ND], M[ND * ND];
  Prim<ND> s;
#pragma unroll
  for (int c = 0; c < NU; ++c) dq[c] = a.W[(size_t)c * a.csW + p];
  s.v = a.v[p];
  s.T = a.T[p];
  s.p = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) s.u[i] = a.u[(size_t)i * a.cs + p];
#pragma unroll
  for (int c = 0; c < ND * ND; ++c) M[c] = a.m[(size_t)c * a.cs + p];
  double mu = 0.0, lam = 0.0, kap = 0.0, jac = 0.0;
  double dT[ND][NU - 1];
  if (a.viscous) {
#pragma unroll
    for (int c = 0; c < ND * ND; ++c) tau[c] = a.tau[(size_t)c * a.cs + p];
#pragma unroll
    for (int i = 0; i < ND; ++i) q[i] = a.q[(size_t)i * a.cs + p];
    mu = a.mu[p]; lam = a.lam[p]; kap = a.kap[p]; jac = a.jac[p];
#pragma unroll
    for (int j = 0; j < ND; ++j)
#pragma unroll
      for (int c = 0; c < NU - 1; ++c) dT[j][c] = a.dT[(size_t)(c + (NU - 1) * j) * a.cs + p];
  }
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    double f1[NU], f2[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) { f1[c] = 0.0; f2[c] = 0.0; }
    add_flux_jacobian_apply<ND>(s, &M[ND * i], a.gamma, false, a.powerLaw, tau, q, dq, f1);        // A_i dQ
    if (a.viscous) {
      double ft[NU];
#pragma unroll
      for (int c = 0; c < NU; ++c) ft[c] = 0.0;
      add_flux_jacobian_apply<ND>(s, &M[ND * i], a.gamma, true, a.powerLaw, tau, q, dq, ft);       // (A_i - B1_i) dQ
#pragma unroll
      for (int c = 0; c < NU; ++c) f2[c] = f1[c] - ft[c];                                          // B1_i dQ
#pragma unroll
      for (int j = 0; j < ND; ++j)
        add_second_partial_apply<ND>(s.u, mu, lam, kap, jac, &M[ND * i], &M[ND * j], dT[j], &f2[1]);
      if (a.Fv) {
#pragma unroll
        for (int c = 0; c < NU; ++c) a.Fv[(size_t)(c + NU * i) * a.cs + p] = f2[c];
      }
    }
#pragma unroll
    for (int c = 0; c < NU; ++c) a.Fhat[(size_t)(c + NU * i) * a.cs + p] = f1[c] - f2[c];
  }
}

struct AdjFinishArgs {
  const double* jacScale;   // when set the contribution is multiplied by 1/J (RHS already holds R/J: fused sweeps)
  const double *Q, *v, *u;
  const double* d;    // derivative of adjoint diffusion: component c + (NU-1)*j
  double* rhs;
  size_t csQ, cs, N;
  double gamma, sign;
};

// Variable change after the second viscous sweep (reference RhsHelperImpl.f90:558-570; the far-field
// adjoint penalty reuses it with the opposite sign, :228-240).
template <int ND>
__global__ void k_adjoint_finish(AdjFinishArgs a) {
  constexpr int NU = ND + 2;
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  double t[ND + 1];
#pragma unroll
  for (int c = 0; c < ND + 1; ++c) {
    double s = a.d[(size_t)c * a.cs + p];
#pragma unroll
    for (int j = 1; j < ND; ++j) s += a.d[(size_t)(c + (NU - 1) * j) * a.cs + p];
    t[c] = s;
  }
  const double v = a.v[p];
  double u[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) u[i] = a.u[(size_t)i * a.cs + p];
  t[ND] = a.gamma * v * t[ND];
  double ut = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    t[i] = v * t[i] - u[i] * t[ND];
    ut = (i == 0) ? u[0] * t[0] : ut + u[i] * t[i];
  }
  const double f = a.jacScale ? a.sign * a.jacScale[p] : a.sign;
#pragma unroll
  for (int c = 0; c < ND + 1; ++c) a.rhs[(size_t)(c + 1) * a.cs + p] -= f * t[c];
  a.rhs[p] += f * (v * a.Q[(size_t)(ND + 1) * a.csQ + p] * t[ND] + ut);
}

__global__ void k_mul_jacobian(double* rhs, size_t cs, int nU, const double* jac, size_t N) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= N) return;
  for (int c = 0; c < nU; ++c) rhs[(size_t)c * cs + p] *= jac[p];
}

__global__ void k_mask_holes(double* rhs, size_t cs, int nU, const int* iblank, size_t N) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= N) return;
  if (iblank[p] == 0)
    for (int c = 0; c < nU; ++c) rhs[(size_t)c * cs + p] = 0.0;
}

struct SrcArgs {
  double* rhsE;
  const double* x[3];
  const int* iblank;
  double loc[3], a, gaussianFactor;
  int nD;
  size_t N;
};

// addAcousticSource (reference src/AcousticSourceImpl.f90:34-64)
__global__ void k_acoustic(SrcArgs a) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  if (a.iblank && a.iblank[p] == 0) return;
  double r = 0.0;
  for (int i = 0; i < a.nD; ++i) {
    const double d = a.x[i][p] - a.loc[i];
    r = (i == 0) ? d * d : r + d * d;
  }
  a.rhsE[p] += a.a * exp(-a.gaussianFactor * r);
}

struct RkArgs {
  const double* R;
  const double* Qin;   // state the RHS was evaluated at
  double *b1, *b2, *Qout;
  size_t cs, N;
  int nU, stage;
  double dt;           // signed: +dt forward, -dt adjoint
};

// substepForward / substepAdjoint axpys (reference src/RK4IntegratorImpl.f90:106-158, :205-264)
__global__ void k_rk4(RkArgs a) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  for (int c = 0; c < a.nU; ++c) {
    const size_t q = (size_t)c * a.cs + p;
    const double R = a.R[q];
    if (a.stage == 1) {
      const double Q0 = a.Qin[q];
      a.b1[q] = Q0;
      a.b2[q] = Q0 + a.dt * R / 6.0;
      a.Qout[q] = Q0 + a.dt * R / 2.0;
    } else if (a.stage == 2) {
      a.b2[q] = a.b2[q] + a.dt * R / 3.0;
      a.Qout[q] = a.b1[q] + a.dt * R / 2.0;
    } else if (a.stage == 3) {
      a.b2[q] = a.b2[q] + a.dt * R / 3.0;
      a.Qout[q] = a.b1[q] + a.dt * R;
    } else {
      a.Qout[q] = a.b2[q] + a.dt * R / 6.0;
    }
  }
}

struct CflArgs {
  const double* Q; size_t csQ;
  const double *m, *jac; size_t cs;
  const int* iblank;
  size_t N;
  PhysParams pp;
  double* partial;     // [gridDim.x] block maxima of the local wave speed
};

// Largest local wave speed (inviscid: J' (c |grad xi| + sum_j |u . M_j|); viscous: J'^2 |M|^2 max(2 mu, kappa)):
// CFL = dt x max, dt = CFL / max (reference src/CNSHelperImpl.f90:842-982).  The dependent variables are
// recomputed from Q (the fused sweeps do not materialise them); hole points are skipped.
template <int ND>
__global__ void __launch_bounds__(256) k_wave_speed(CflArgs a) {
  __shared__ double sh[256];
  double best = 0.0;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < a.N; p += (size_t)gridDim.x * blockDim.x) {
    if (a.iblank && a.iblank[p] == 0) continue;
    double Q[ND + 2];
#pragma unroll
    for (int c = 0; c < ND + 2; ++c) Q[c] = a.Q[(size_t)c * a.csQ + p];
    Prim<ND> s;
    dependent<ND>(Q, a.pp.gamma, s);
    const double c0 = sqrt((a.pp.gamma - 1.0) * s.T);
    double msq = 0.0, conv = 0.0;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      double dot = 0.0;
#pragma unroll
      for (int l = 0; l < ND; ++l) {
        const double m = a.m[(size_t)(l + ND * j) * a.cs + p];
        msq += m * m;
        dot += s.u[l] * m;
      }
      conv += fabs(dot);
    }
    const double J = a.jac[p];
    best = fmax(best, J * (c0 * sqrt(msq) + conv));
    if (a.pp.viscous) {
      double mu, lam, kap;
      transport(s.T, a.pp, mu, lam, kap);
      best = fmax(best, J * J * msq * fmax(2.0 * mu, kap));
    }
  }
  sh[threadIdx.x] = best;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + st]);
    __syncthreads();
  }
  if (threadIdx.x == 0) a.partial[blockIdx.x] = sh[0];
}

template <typename F>
int dispatch_nd(int nD, F f) {
  if (nD == 1) return f(std::integral_constant<int, 1>());
  if (nD == 2) return f(std::integral_constant<int, 2>());
  return f(std::integral_constant<int, 3>());
}

}  // namespace

// ------------------------------------------------------------------------------------ state
int mg_state_create_impl(mg_grid* g, const mg_options_t* opt, mg_state** out) {
  auto* s = new mg_state();
  s->grid = g;
  s->opt = *opt;
  s->nD = g->nD;
  s->nU = g->nD + 2;
  const int nD = s->nD, nU = s->nU;
  MG_TRY(mg_field_alloc(g, nU, &s->Q[0]));
  MG_TRY(mg_field_alloc(g, nU, &s->Q[1]));
  MG_TRY(mg_field_alloc(g, nU, &s->W[0]));
  MG_TRY(mg_field_alloc(g, nU, &s->W[1]));
  MG_TRY(mg_field_alloc(g, nU, &s->target));
  MG_TRY(mg_field_alloc(g, nU, &s->rhs));
  MG_TRY(mg_field_alloc(g, 1, &s->specificVolume));
  MG_TRY(mg_field_alloc(g, nD, &s->velocity));
  MG_TRY(mg_field_alloc(g, 1, &s->pressure));
  MG_TRY(mg_field_alloc(g, 1, &s->temperature));
  MG_TRY(mg_field_alloc(g, nU, &s->rk1));
  MG_TRY(mg_field_alloc(g, nU, &s->rk2));
  // Q, W and rk1 exchange roles with each other and with the checkpoint slots: pooled storage
  for (MgField* f : {&s->Q[0], &s->Q[1], &s->W[0], &s->W[1], &s->rk1}) {
    s->pool.push_back(f->p);
    f->owned = false;
  }
  if (opt->viscosityOn) {
    MG_TRY(mg_field_alloc(g, 1, &s->mu));
    MG_TRY(mg_field_alloc(g, 1, &s->lambda));
    MG_TRY(mg_field_alloc(g, 1, &s->kappa));
    MG_TRY(mg_field_alloc(g, nD * nD, &s->stressTensor));
    MG_TRY(mg_field_alloc(g, nD, &s->heatFlux));
  }
  // scratch shared by the general path: two fields of nU*nD components
  if (g->scratchA.nComp < nU * nD) MG_TRY(mg_field_alloc(g, nU * nD, &g->scratchA));
  if (g->scratchB.nComp < nU * nD) MG_TRY(mg_field_alloc(g, nU * nD, &g->scratchB));
  *out = s;
  return 0;
}

// ---- pooled nU-component buffers (Q, W, rk1, checkpoint slots)
namespace {
template <class F>
void for_each_pooled(mg_state* s, F&& fn) {
  for (MgField* f : {&s->Q[0], &s->Q[1], &s->W[0], &s->W[1], &s->rk1, &s->staged[0], &s->staged[1]}) fn(f);
  for (MgField& c : s->checkpoints) fn(&c);
}
bool pool_referenced(mg_state* s, const double* p, const MgField* except) {
  bool used = false;
  for_each_pooled(s, [&](MgField* f) { if (f != except && f->p == p) used = true; });
  return used;
}
}  // namespace

// A device->host read of pooled buffer p has been enqueued on readStream: until it completes the buffer must not
// be reused as free storage (re-storing a checkpoint slot or adopting a staged input drops the last reference).
void mg_state_note_pending_read(mg_state* s, double* p, cudaStream_t readStream) {
  cudaEvent_t e = nullptr;
  if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return;
  cudaEventRecord(e, readStream);
  s->pendingReads.push_back({p, e});
}

static bool read_pending(mg_state* s, const double* p) {
  bool pending = false;
  for (size_t i = 0; i < s->pendingReads.size();) {
    if (cudaEventQuery(s->pendingReads[i].done) == cudaSuccess) {
      cudaEventDestroy(s->pendingReads[i].done);
      s->pendingReads.erase(s->pendingReads.begin() + i);
      continue;
    }
    if (s->pendingReads[i].p == p) pending = true;
    ++i;
  }
  cudaGetLastError();      // cudaErrorNotReady from the queries is not an error
  return pending;
}

double* mg_state_pool_acquire(mg_state* s, size_t bytes) {
  // buffers are handed out in rotation: the one after the last hit is usually free (a checkpoint window takes them in
  // order), which keeps the search O(references) instead of O(pool x references) per substep
  const size_t P = s->pool.size();
  for (size_t probe = 0; probe < std::min<size_t>(P, 2); ++probe) {
    const size_t i = (s->poolCursor + probe) % P;
    double* p = s->pool[i];
    if (!pool_referenced(s, p, nullptr) && !read_pending(s, p)) { s->poolCursor = (i + 1) % P; return p; }
  }
  if (P > 2) {
    std::unordered_set<const double*> used;
    for_each_pooled(s, [&](MgField* f) { if (f->p) used.insert(f->p); });
    for (size_t k = 0; k < P; ++k) {
      const size_t i = (s->poolCursor + k) % P;
      double* p = s->pool[i];
      if (!used.count(p) && !read_pending(s, p)) { s->poolCursor = (i + 1) % P; return p; }
    }
  }
  double* fresh = nullptr;
  if (cudaMalloc(&fresh, bytes) != cudaSuccess) return nullptr;
  cudaMemsetAsync(fresh, 0, bytes, mg_stream());
  s->pool.push_back(fresh);
  s->poolCursor = 0;
  return fresh;
}

// Give `f` storage that no other pooled field refers to (a shared buffer keeps serving the others).
int mg_state_make_exclusive(mg_state* s, MgField* f, bool keepContents) {
  if (!f->p || (!pool_referenced(s, f->p, f) && !read_pending(s, f->p))) return 0;
  const size_t bytes = f->compStride * (size_t)f->nComp * sizeof(double);
  double* fresh = mg_state_pool_acquire(s, bytes);
  if (!fresh) MG_FAIL("out of device memory for a state buffer");
  if (keepContents) MG_CUDA(cudaMemcpyAsync(fresh, f->p, bytes, cudaMemcpyDeviceToDevice, mg_stream()));
  f->p = fresh;
  return 0;
}

// Release pooled buffers nothing refers to any more.
void mg_state_pool_trim(mg_state* s) {
  s->poolCursor = 0;
  std::vector<double*> keep;
  for (double* p : s->pool) {
    if (pool_referenced(s, p, nullptr) || read_pending(s, p)) keep.push_back(p);
    else cudaFree(p);
  }
  s->pool.swap(keep);
}

void mg_state_destroy_impl(mg_state* s) {
  if (!s) return;
  for (MgField* f : {&s->Q[0], &s->Q[1], &s->W[0], &s->W[1], &s->target, &s->rhs, &s->specificVolume,
                     &s->velocity, &s->pressure, &s->temperature, &s->mu, &s->lambda, &s->kappa,
                     &s->stressTensor, &s->heatFlux, &s->rk1, &s->rk2, &s->viscFluxCart, &s->tauq, &s->dissTerm,
                     &s->meanPressure, &s->meanVelocity})
    mg_field_free(f);
  for (void* p : s->fusedOps) if (p) cudaFree(p);
  for (double* p : s->pool) cudaFree(p);
  cudaFree(s->accumulators);
  delete s;
}

// updateState (reference src/StateImpl.f90:466-537)
int mg_state_update_impl(mg_state* s, const MgField* Qoverride) {
  mg_grid* g = s->grid;
  MG_TRY(mg_halo_wait_pending());
  const MgField& Q = Qoverride ? *Qoverride : s->Q[s->cur];
  const size_t N = g->N;
  cudaStream_t st = mg_stream();
  PtrSet a;
  a.Q = Q.comp(0);
  a.csQ = Q.compStride;
  a.v = s->specificVolume.comp(0);
  a.u = s->velocity.comp(0);
  a.p = s->pressure.comp(0);
  a.T = s->temperature.comp(0);
  a.mu = s->opt.viscosityOn ? s->mu.comp(0) : nullptr;
  a.lam = s->opt.viscosityOn ? s->lambda.comp(0) : nullptr;
  a.kap = s->opt.viscosityOn ? s->kappa.comp(0) : nullptr;
  a.cs = s->velocity.compStride;
  a.N = N;
  const PhysParams pp = s->phys();
  MG_TRY(dispatch_nd(s->nD, [&](auto nd) {
    { k_dependent<decltype(nd)::value><<<nblocks(N), 256, 0, st>>>(a, pp); mg_count_launches(1); }
    return 0;
  }));
  MG_CUDA(cudaGetLastError());
  if (s->opt.viscosityOn) {
    MG_TRY(mg_grid_gradient_dev(g, s->velocity.comp(0), s->velocity.compStride, s->nD, &s->stressTensor,
                                &g->scratchA));
    MG_TRY(dispatch_nd(s->nD, [&](auto nd) {
      { k_stress<decltype(nd)::value><<<nblocks(N), 256, 0, st>>>(s->stressTensor.comp(0), s->stressTensor.compStride,
                                                               s->mu.comp(0), s->lambda.comp(0), N); mg_count_launches(1); }
      return 0;
    }));
    MG_TRY(mg_grid_gradient_dev(g, s->temperature.comp(0), s->temperature.compStride, 1, &s->heatFlux,
                                &g->scratchA));
    { k_heatflux<<<nblocks(N), 256, 0, st>>>(s->heatFlux.comp(0), s->heatFlux.compStride, s->nD, s->kappa.comp(0), N); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
  }
  s->dependentValid = true;
  return 0;
}

// The fused sweep A keeps the stress tensor as its nD(nD+1)/2 unique entries followed by the heat flux: expand it
// into the reference layout (stressTensor(N, nD^2), heatFlux(N, nD)) for the patch kernels and the getters.
template <int ND>
__global__ void k_expand_tauq(const double* tq, size_t cs, double* tau, double* q, size_t N) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= N) return;
  constexpr int NTAU = ND * (ND + 1) / 2;
#pragma unroll
  for (int l = 0; l < ND; ++l)
#pragma unroll
    for (int c = 0; c < ND; ++c) {
      const int r0 = l < c ? l : c, c0 = l < c ? c : l;
      tau[(size_t)(l + ND * c) * cs + p] = tq[(size_t)(r0 * ND - r0 * (r0 - 1) / 2 + (c0 - r0)) * cs + p];
    }
#pragma unroll
  for (int d = 0; d < ND; ++d) q[(size_t)d * cs + p] = tq[(size_t)(NTAU + d) * cs + p];
}

// Dependent variables of the current state from what the fused sweep A already produced: one pointwise kernel for
// (v, u, p, T, mu, lambda, kappa) and one that expands tau / q -- no derivative is taken again.
int mg_state_dependents_from_fused(mg_state* s) {
  mg_grid* g = s->grid;
  if (!s->fusedValid) MG_FAIL("dependent variables from the fused sweep: sweep A has not run on the current state");
  const MgField& Q = s->Q[s->cur];
  const size_t N = g->N;
  cudaStream_t st = mg_stream();
  PtrSet a;
  a.Q = Q.comp(0);
  a.csQ = Q.compStride;
  a.v = s->specificVolume.comp(0);
  a.u = s->velocity.comp(0);
  a.p = s->pressure.comp(0);
  a.T = s->temperature.comp(0);
  a.mu = s->opt.viscosityOn ? s->mu.comp(0) : nullptr;
  a.lam = s->opt.viscosityOn ? s->lambda.comp(0) : nullptr;
  a.kap = s->opt.viscosityOn ? s->kappa.comp(0) : nullptr;
  a.cs = s->velocity.compStride;
  a.N = N;
  const PhysParams pp = s->phys();
  MG_TRY(dispatch_nd(s->nD, [&](auto nd) {
    { k_dependent<decltype(nd)::value><<<nblocks(N), 256, 0, st>>>(a, pp); mg_count_launches(1); }
    if (s->opt.viscosityOn) {
      k_expand_tauq<decltype(nd)::value><<<nblocks(N), 256, 0, st>>>(s->tauq.comp(0), s->tauq.compStride,
                                                                     s->stressTensor.comp(0), s->heatFlux.comp(0), N);
      mg_count_launches(1);
    }
    return 0;
  }));
  MG_CUDA(cudaGetLastError());
  s->dependentValid = true;
  return 0;
}

// addDissipation (reference src/RhsHelperImpl.f90:10-87)
static int add_dissipation_general(mg_state* s, int mode) {
  mg_grid* g = s->grid;
  if (!s->opt.dissipationOn) return 0;
  const size_t N = g->N;
  cudaStream_t st = mg_stream();
  const double amount = mode == MG_ADJOINT ? -s->opt.dissipationAmount : s->opt.dissipationAmount;
  const MgField& X = mode == MG_FORWARD ? s->Q[s->cur] : s->W[s->curW];
  MgField& A = g->scratchA;
  MgField& B = g->scratchB;
  for (int i = 0; i < s->nD; ++i) {
    MG_TRY(mg_grid_apply(g, g->dissipation[i], X.comp(0), X.compStride, A.comp(0), A.compStride, s->nU));
    const double* result = A.comp(0);
    if (!g->compositeDissipation) {
      { k_scale_by<<<nblocks(N), 256, 0, st>>>(A.comp(0), A.compStride, s->nU, g->arcLengths.comp(i), -1.0, N); mg_count_launches(1); }
      MG_TRY(mg_grid_apply(g, g->dissipationTranspose[i], A.comp(0), A.compStride, B.comp(0), B.compStride, s->nU));
      MG_TRY(mg_norm_launch(g->firstDerivative[i], B.comp(0), B.compStride, s->nU, g->localSize, 1, st));
      result = B.comp(0);
    }
    { k_axpy<<<nblocks(N), 256, 0, st>>>(s->rhs.comp(0), result, s->rhs.compStride, A.compStride, s->nU, amount, N); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
  }
  return 0;
}

// computeRhsForward (reference src/RhsHelperImpl.f90:254-354)
int mg_state_rhs_forward_general(mg_state* s) {
  mg_grid* g = s->grid;
  const size_t N = g->N;
  cudaStream_t st = mg_stream();
  const int nD = s->nD, nU = s->nU;
  MgField& Fh = g->scratchA;
  MgField& Dv = g->scratchB;
  if (s->keepViscousFluxes && s->opt.viscosityOn && s->viscFluxCart.nComp < nU * nD)
    MG_TRY(mg_field_alloc(g, nU * nD, &s->viscFluxCart));
  FluxArgs a;
  const MgField& Q = s->Q[s->cur];
  a.Q = Q.comp(0);
  a.csQ = Q.compStride;
  a.u = s->velocity.comp(0);
  a.pr = s->pressure.comp(0);
  a.tau = s->opt.viscosityOn ? s->stressTensor.comp(0) : nullptr;
  a.q = s->opt.viscosityOn ? s->heatFlux.comp(0) : nullptr;
  a.m = g->metrics.comp(0);
  a.Fhat = Fh.comp(0);
  a.Fv = (s->keepViscousFluxes && s->opt.viscosityOn) ? s->viscFluxCart.comp(0) : nullptr;
  a.cs = Fh.compStride;
  a.N = N;
  a.viscous = s->opt.viscosityOn;
  a.curvilinear = g->isCurvilinear;
  MG_TRY(dispatch_nd(nD, [&](auto nd) {
    { k_flux<decltype(nd)::value><<<nblocks(N), 256, 0, st>>>(a); mg_count_launches(1); }
    return 0;
  }));
  MG_CUDA(cudaGetLastError());
  if (s->keepViscousFluxes && s->opt.viscosityOn) MG_TRY(mg_patches_collect_viscous(s));
  for (int i = 0; i < nD; ++i) {
    MG_TRY(mg_grid_apply(g, g->firstDerivative[i], Fh.comp(nU * i), Fh.compStride, Dv.comp(nU * i), Dv.compStride, nU));
  }
  { k_neg_sum<<<nblocks(N), 256, 0, st>>>(s->rhs.comp(0), Dv.comp(0), Dv.compStride, nU, nD, N); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return add_dissipation_general(s, MG_FORWARD);
}

// adjointFirstDerivative of src(:, :, j) along j (into scratch A), summed over j, followed by the change of
// variables back to the conserved adjoint; rhs -/+= sign * result (reference src/RhsHelperImpl.f90:528-570,
// :213-247, :987-1023).
int mg_state_adjoint_finish(mg_state* s, MgField* src, double sign, bool timesJacobian) {
  mg_grid* g = s->grid;
  const size_t N = g->N;
  const int nD = s->nD, nU = s->nU;
  MgField& A = g->scratchA;
  const MgField& Q = s->Q[s->cur];
  for (int j = 0; j < nD; ++j)
    MG_TRY(mg_grid_apply(g, g->adjointFirstDerivative[j], src->comp((nU - 1) * j), src->compStride,
                         A.comp((nU - 1) * j), A.compStride, nU - 1));
  AdjFinishArgs f;
  f.jacScale = timesJacobian ? g->jacobian.comp(0) : nullptr;
  f.Q = Q.comp(0);
  f.csQ = Q.compStride;
  f.v = s->specificVolume.comp(0);
  f.u = s->velocity.comp(0);
  f.d = A.comp(0);
  f.rhs = s->rhs.comp(0);
  f.cs = A.compStride;
  f.N = N;
  f.gamma = s->opt.ratioOfSpecificHeats;
  f.sign = sign;
  MG_TRY(dispatch_nd(nD, [&](auto nd) {
    { k_adjoint_finish<decltype(nd)::value><<<nblocks(N), 256, 0, mg_stream()>>>(f); mg_count_launches(1); }
    return 0;
  }));
  MG_CUDA(cudaGetLastError());
  return 0;
}

// computeRhsAdjoint (reference src/RhsHelperImpl.f90:356-596)
int mg_state_rhs_adjoint_general(mg_state* s) {
  mg_grid* g = s->grid;
  const size_t N = g->N;
  cudaStream_t st = mg_stream();
  const int nD = s->nD, nU = s->nU;
  MgField& A = g->scratchA;
  MgField& B = g->scratchB;
  const MgField& W = s->W[s->curW];
  const MgField& Q = s->Q[s->cur];
  for (int i = 0; i < nD; ++i)
    MG_TRY(mg_grid_apply(g, g->adjointFirstDerivative[i], W.comp(0), W.compStride, A.comp(nU * i), A.compStride, nU));
  AdjArgs a;
  std::memset(&a, 0, sizeof(a));
  a.Q = Q.comp(0);
  a.csQ = Q.compStride;
  a.v = s->specificVolume.comp(0);
  a.u = s->velocity.comp(0);
  a.T = s->temperature.comp(0);
  if (s->opt.viscosityOn) {
    a.tau = s->stressTensor.comp(0);
    a.q = s->heatFlux.comp(0);
    a.mu = s->mu.comp(0);
    a.lam = s->lambda.comp(0);
    a.kap = s->kappa.comp(0);
  }
  a.m = g->metrics.comp(0);
  a.jac = g->jacobian.comp(0);
  a.dW = A.comp(0);
  a.rhs = s->rhs.comp(0);
  a.diff = B.comp(0);
  a.cs = A.compStride;
  a.N = N;
  a.viscous = s->opt.viscosityOn;
  a.curvilinear = g->isCurvilinear;
  a.gamma = s->opt.ratioOfSpecificHeats;
  a.powerLaw = s->opt.powerLawExponent;
  MG_TRY(dispatch_nd(nD, [&](auto nd) {
    { k_adjoint_point<decltype(nd)::value><<<nblocks(N), 256, 0, st>>>(a); mg_count_launches(1); }
    return 0;
  }));
  MG_CUDA(cudaGetLastError());
  if (s->opt.viscosityOn) MG_TRY(mg_state_adjoint_finish(s, &B, 1.0));
  MG_TRY(add_dissipation_general(s, MG_ADJOINT));
  if (s->opt.viscosityOn && mg_patches_have_farfield(s)) {
    // addFarFieldAdjointPenalty (reference src/RhsHelperImpl.f90:89-250)
    MG_TRY(mg_field_zero(g, &B));
    MG_TRY(mg_patches_farfield_adjoint_sources(s, &B));
    MG_TRY(mg_state_adjoint_finish(s, &B, -1.0));
  }
  return 0;
}

// wantDt = 0: result = given(dt) x max wave speed (computeCfl); 1: result = given(cfl) / max wave speed
// (computeTimeStepSize).  Local to this rank: the caller reduces over ranks (MPI_Allreduce MAX / MIN in the
// reference, src/RegionImpl.f90:1837-1873).
int mg_state_cfl_dt_impl(mg_state* s, int wantDt, double given, double* result) {
  mg_grid* g = s->grid;
  if (!g->updated) MG_FAIL("cfl / time step: grid metrics have not been computed (mg_grid_update)");
  MG_TRY(mg_halo_wait_pending());
  const int blocks = 592;
  static double* d_partial = nullptr;
  static double* h_partial = nullptr;
  if (!d_partial) {
    MG_CUDA(cudaMalloc(&d_partial, blocks * sizeof(double)));
    MG_CUDA(cudaMallocHost(&h_partial, blocks * sizeof(double)));
  }
  CflArgs a;
  a.Q = s->Q[s->cur].comp(0);
  a.csQ = s->Q[s->cur].compStride;
  a.m = g->metrics.comp(0);
  a.jac = g->jacobian.comp(0);
  a.cs = g->metrics.compStride;
  a.iblank = g->iblank;
  a.N = g->N;
  a.pp = s->phys();
  a.partial = d_partial;
  MG_TRY(dispatch_nd(s->nD, [&](auto nd) {
    { k_wave_speed<decltype(nd)::value><<<blocks, 256, 0, mg_stream()>>>(a); mg_count_launches(1); }
    return 0;
  }));
  MG_CUDA(cudaGetLastError());
  MG_CUDA(cudaMemcpyAsync(h_partial, d_partial, blocks * sizeof(double), cudaMemcpyDeviceToHost, mg_stream()));
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  double w = 0.0;
  for (int i = 0; i < blocks; ++i) w = h_partial[i] > w ? h_partial[i] : w;
  *result = wantDt ? given / w : given * w;
  return 0;
}

// computeRhs for one grid/state (reference src/RegionImpl.f90:1877-2027)
// computeRhsLinearized (reference src/RhsHelperImpl.f90:598-829)
int mg_state_rhs_linearized_general(mg_state* s) {
  mg_grid* g = s->grid;
  const size_t N = g->N;
  cudaStream_t st = mg_stream();
  const int nD = s->nD, nU = s->nU;
  MgField& A = g->scratchA;
  MgField& B = g->scratchB;
  const MgField& W = s->W[s->curW];
  const MgField& Q = s->Q[s->cur];
  const bool keep = s->keepViscousFluxes && s->opt.viscosityOn;
  if (keep && s->viscFluxCart.nComp < nU * nD) MG_TRY(mg_field_alloc(g, nU * nD, &s->viscFluxCart));
  LinArgs a;
  std::memset(&a, 0, sizeof(a));
  a.Q = Q.comp(0); a.csQ = Q.compStride;
  a.W = W.comp(0); a.csW = W.compStride;
  a.v = s->specificVolume.comp(0);
  a.u = s->velocity.comp(0);
  a.T = s->temperature.comp(0);
  a.m = g->metrics.comp(0);
  a.jac = g->jacobian.comp(0);
  a.cs = A.compStride;
  a.N = N;
  a.viscous = s->opt.viscosityOn;
  a.gamma = s->opt.ratioOfSpecificHeats;
  a.powerLaw = s->opt.powerLawExponent;
  if (s->opt.viscosityOn) {
    a.tau = s->stressTensor.comp(0);
    a.q = s->heatFlux.comp(0);
    a.mu = s->mu.comp(0);
    a.lam = s->lambda.comp(0);
    a.kap = s->kappa.comp(0);
    a.t = B.comp(0);
    MG_TRY(dispatch_nd(nD, [&](auto nd) {
      { k_linearized_primitive<decltype(nd)::value><<<nblocks(N), 256, 0, st>>>(a); mg_count_launches(1); }
      return 0;
    }));
    MG_CUDA(cudaGetLastError());
    for (int i = 0; i < nD; ++i)
      MG_TRY(mg_grid_apply(g, g->firstDerivative[i], B.comp(0), B.compStride, A.comp((nU - 1) * i), A.compStride, nU - 1));
    a.dT = A.comp(0);
  }
  a.Fhat = B.comp(0);
  a.Fv = keep ? s->viscFluxCart.comp(0) : nullptr;
  MG_TRY(dispatch_nd(nD, [&](auto nd) {
    { k_linearized_flux<decltype(nd)::value><<<nblocks(N), 256, 0, st>>>(a); mg_count_launches(1); }
    return 0;
  }));
  MG_CUDA(cudaGetLastError());
  if (keep) MG_TRY(mg_patches_collect_viscous(s));
  for (int i = 0; i < nD; ++i)
    MG_TRY(mg_grid_apply(g, g->firstDerivative[i], B.comp(nU * i), B.compStride, A.comp(nU * i), A.compStride, nU));
  { k_neg_sum<<<nblocks(N), 256, 0, st>>>(s->rhs.comp(0), A.comp(0), A.compStride, nU, nD, N); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return add_dissipation_general(s, MG_LINEARIZED);
}

// computeRhs up to (not including) the multiplication by 1/J: the part that precedes the block-interface
// exchange in the reference (src/RegionImpl.f90:1877-1925).  General path only.
int mg_state_rhs_pre(mg_state* s, int mode) {
  mg_grid* g = s->grid;
  if (!g->updated) MG_FAIL("computeRhs: grid metrics have not been computed (mg_grid_update)");
  MG_TRY(mg_halo_wait_pending());
  // The reference's callers run state%update after every substep (src/SolverImpl.f90:831-834); here
  // it is refreshed on demand.
  if (!s->dependentValid) MG_TRY(mg_state_update_impl(s, nullptr));
  if (mode == MG_FORWARD) MG_TRY(mg_state_rhs_forward_general(s));
  else if (mode == MG_ADJOINT) MG_TRY(mg_state_rhs_adjoint_general(s));
  else MG_TRY(mg_state_rhs_linearized_general(s));
  return 0;
}

// ... and from the viscous interface adjoint penalty on (src/RegionImpl.f90:1960-2027): x 1/J, patches, sources,
// hole masking.
int mg_state_rhs_post(mg_state* s, int mode, bool alreadyTimesJacobian) {
  mg_grid* g = s->grid;
  const size_t N = g->N;
  cudaStream_t st = mg_stream();
  if (mode == MG_ADJOINT && s->opt.viscosityOn && mg_state_has_interfaces(s)) {
    // addInterfaceAdjointPenalty (reference src/RhsHelperImpl.f90:831-1026)
    MG_TRY(mg_field_zero(g, &g->scratchB));
    MG_TRY(mg_interfaces_adjoint_sources(s, &g->scratchB));
    MG_TRY(mg_state_adjoint_finish(s, &g->scratchB, -1.0));
  }
  if (!alreadyTimesJacobian)
    { k_mul_jacobian<<<nblocks(N), 256, 0, st>>>(s->rhs.comp(0), s->rhs.compStride, s->nU, g->jacobian.comp(0), N); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  MG_TRY(mg_patches_apply(s, mode));
  if (mode == MG_ADJOINT && s->limits.soft && s->limits.forcingSwitch) {
    // addSolutionLimitPenaltyAdjointForcing (reference src/RegionImpl.f90:2002-2005, :1094-1221)
    int rhoOut = s->limits.rhoOut, tOut = s->limits.tOut;
    if (rhoOut < 0 || tOut < 0) {
      double lo, hi;
      MG_TRY(mg_state_extrema_impl(s, 0, &lo, nullptr, &hi, nullptr));
      rhoOut = lo <= s->limits.densityRange[0] || hi >= s->limits.densityRange[1];
      MG_TRY(mg_state_extrema_impl(s, 1, &lo, nullptr, &hi, nullptr));
      tOut = lo <= s->limits.temperatureRange[0] || hi >= s->limits.temperatureRange[1];
    }
    MG_TRY(mg_state_limit_forcing_impl(s, s->limits.densityRange, s->limits.temperatureRange, rhoOut, tOut,
                                       s->limits.penaltyFactor));
  }
  if (mode == MG_FORWARD) {
    for (const auto& src : s->acousticSources) {
      SrcArgs a;
      a.rhsE = s->rhs.comp(s->nD + 1);
      for (int i = 0; i < 3; ++i) {
        a.x[i] = i < s->nD ? g->coordinates.comp(i) : nullptr;
        a.loc[i] = src.loc[i];
      }
      a.iblank = g->iblank;
      a.a = src.amplitude * cos(src.angularFrequency * s->time + src.phase);
      a.gaussianFactor = src.gaussianFactor;
      a.nD = s->nD;
      a.N = N;
      { k_acoustic<<<nblocks(N), 256, 0, st>>>(a); mg_count_launches(1); }
    }
  }
  if (g->iblank)
    { k_mask_holes<<<nblocks(N), 256, 0, st>>>(s->rhs.comp(0), s->rhs.compStride, s->nU, g->iblank, N); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return 0;
}

static bool fused_rhs_then_patches(const mg_state* s, int mode);

// true when computeRhs of this state runs the fused sweeps (with the RK4 substep folded in, or followed by the patch
// epilogue): then state%update is sweep A, and the reference-layout dependent variables are expanded on demand
bool mg_state_uses_fused_rhs(const mg_state* s, int mode) {
  return s->useFused && (mg_fused_supported(s, mode) || fused_rhs_then_patches(s, mode));
}

// Dependent variables in the reference layout, whichever path is active: from the fused sweep A when that is what
// state%update runs for this state, else the operator-by-operator update.
int mg_state_ensure_dependents(mg_state* s) {
  if (s->dependentValid) return 0;
  if (!s->grid->updated) MG_FAIL("dependent variables requested before mg_grid_update");
  if (mg_state_uses_fused_rhs(s, MG_FORWARD)) {
    if (!s->fusedValid) MG_TRY(mg_fused_sweepA(s));
    return mg_state_dependents_from_fused(s);
  }
  return mg_state_update_impl(s, nullptr);
}

// Cartesian viscous fluxes at the patches that take them (far-field SAT), from the dependent variables
static int collect_patch_viscous_fluxes(mg_state* s) {
  mg_grid* g = s->grid;
  const int nD = s->nD, nU = s->nU;
  if (s->viscFluxCart.nComp < nU * nD) MG_TRY(mg_field_alloc(g, nU * nD, &s->viscFluxCart));
  FluxArgs a;
  const MgField& Q = s->Q[s->cur];
  a.Q = Q.comp(0);
  a.csQ = Q.compStride;
  a.u = s->velocity.comp(0);
  a.pr = s->pressure.comp(0);
  a.tau = s->stressTensor.comp(0);
  a.q = s->heatFlux.comp(0);
  a.m = g->metrics.comp(0);
  a.Fhat = g->scratchA.comp(0);
  a.Fv = s->viscFluxCart.comp(0);
  a.cs = g->scratchA.compStride;
  a.N = g->N;
  a.viscous = 1;
  a.curvilinear = g->isCurvilinear;
  MG_TRY(dispatch_nd(nD, [&](auto nd) {
    { k_flux<decltype(nd)::value><<<nblocks(g->N), 256, 0, mg_stream()>>>(a); mg_count_launches(1); }
    return 0;
  }));
  MG_CUDA(cudaGetLastError());
  return mg_patches_collect_viscous(s);
}

// The RHS of a state WITH patches / sources whose interior scheme the fused sweeps cover: two (forward) or three
// (adjoint) sweeps produce R / J, then the penalties act on the patch points exactly as after the operator-by-operator
// evaluation (they are applied after the 1/J multiplication, reference src/RegionImpl.f90:1969-1995).
static bool fused_rhs_then_patches(const mg_state* s, int mode) {
  if (!s->useFused || !mg_fused_rhs_supported(s, mode) || mg_fused_supported(s, mode)) return false;
  if (mg_state_has_interfaces(s)) return false;           // staged over the region's grids (mg_region_compute_rhs)
  const mg_grid* g = s->grid;
  if (g->procDims[0] * g->procDims[1] * g->procDims[2] != 1) return false;   // the caller of the fused sweeps owns the halos
  return mg_tuning_get("MG_FUSED_PATCHES", 1) != 0;
}

int mg_state_compute_rhs_impl(mg_state* s, int mode) {
  mg_grid* g = s->grid;
  if (!g->updated) MG_FAIL("computeRhs: grid metrics have not been computed (mg_grid_update)");
  if (mode != MG_FORWARD && mode != MG_ADJOINT && mode != MG_LINEARIZED) MG_FAIL("computeRhs: unknown mode");
  if (s->useFused && mg_fused_supported(s, mode)) {
    // fused sweeps: the dependent variables live in the sweep-A outputs (no patches on this path)
    if (!s->fusedValid) MG_TRY(mg_fused_sweepA(s));
    if (mode == MG_FORWARD) return mg_fused_sweepB(s, 0, 0, 0.0);
    MG_TRY(mg_fused_adjoint1(s));
    return mg_fused_adjoint2(s, 0, 1, 0.0);
  }
  if (fused_rhs_then_patches(s, mode)) {
    if (!s->fusedValid) MG_TRY(mg_fused_sweepA(s));
    if (mode == MG_FORWARD) {
      MG_TRY(mg_fused_sweepB(s, 0, 0, 0.0));
    } else {
      MG_TRY(mg_fused_adjoint1(s));
      MG_TRY(mg_fused_adjoint2(s, 0, 1, 0.0));
    }
    if (!s->patches.empty()) {
      // the patch kernels read the dependent variables in the reference layout: expand what sweep A produced
      MG_TRY(mg_state_ensure_dependents(s));
      if (mode == MG_FORWARD && s->keepViscousFluxes && s->opt.viscosityOn) MG_TRY(collect_patch_viscous_fluxes(s));
      if (mode == MG_ADJOINT && s->opt.viscosityOn && mg_patches_have_farfield(s)) {
        // addFarFieldAdjointPenalty (reference src/RhsHelperImpl.f90:89-250), joined after the 1/J multiplication
        bool anyViscousPenalty = false;
        for (const mg_patch* p : s->patches)
          anyViscousPenalty = anyViscousPenalty || (p->type == MG_PATCH_FARFIELD && p->viscousPenaltyAmount != 0.0);
        if (anyViscousPenalty) {
          MG_TRY(mg_field_zero(g, &g->scratchB));
          MG_TRY(mg_patches_farfield_adjoint_sources(s, &g->scratchB));
          MG_TRY(mg_state_adjoint_finish(s, &g->scratchB, -1.0, true));
        }
      }
    }
    return mg_state_rhs_post(s, mode, true);
  }
  if (mg_state_has_interfaces(s)) {
    MG_FAIL("computeRhs: a state with block-interface patches must be evaluated through its region (mg_region_compute_rhs)");
  }
  MG_TRY(mg_state_rhs_pre(s, mode));
  return mg_state_rhs_post(s, mode, false);
}

// substepForward / substepAdjoint (reference src/RK4IntegratorImpl.f90:65-270).  The state update that
// the reference's callers issue after every substep (src/SolverImpl.f90:831-834) stays with the caller.
// The bookkeeping that precedes the RHS evaluation of a substep: times seen by the sources, and the stage factor
// of the adjoint forcing (reference src/RK4IntegratorImpl.f90:106-158, :205-264).
void mg_rk4_set_times(mg_state* s, int mode, double time, double dt, int stage) {
  if (mode == MG_FORWARD) {
    if (stage == 1 || stage == 3) s->timeProgressive = time + dt / 2.0;
    if (stage == 2 || stage == 4) s->time = time + dt / 2.0;
  } else if (mode == MG_ADJOINT) {
    const double factor[5] = {0.0, 1.0, 0.5, 1.0, 2.0};
    s->adjointForcingFactor = factor[stage];
    if (stage == 4) s->timeProgressive = time - dt / 2.0;
  }
}

int mg_rk4_substep_impl(mg_state* s, int mode, double* time, double dt, int timestep, int stage) {
  (void)timestep;
  mg_grid* g = s->grid;
  const size_t N = g->N;
  cudaStream_t st = mg_stream();
  if (stage < 1 || stage > 4) MG_FAIL("rk4 substep: stage must be 1..4");
  RkArgs a;
  a.cs = s->rhs.compStride;
  a.N = N;
  a.nU = s->nU;
  a.b1 = s->rk1.comp(0);
  a.b2 = s->rk2.comp(0);
  if (mode == MG_FORWARD) {
    if (stage == 1) s->timeProgressive = *time + dt / 2.0;
    if (stage == 2 || stage == 4) { *time += dt / 2.0; s->time = *time; }
    if (stage == 3) s->timeProgressive = *time + dt / 2.0;
    if (s->useFused && mg_fused_supported(s, MG_FORWARD)) {
      if (!s->fusedValid) MG_TRY(mg_fused_sweepA(s));
      return mg_fused_sweepB(s, 1, stage, dt);
    }
    if (!s->rhsReady) MG_TRY(mg_state_compute_rhs_impl(s, MG_FORWARD));
    s->rhsReady = false;
    MG_TRY(mg_state_make_exclusive(s, &s->Q[s->cur], true));
    MG_TRY(mg_state_make_exclusive(s, &s->rk1, true));
    a.b1 = s->rk1.comp(0);
    a.R = s->rhs.comp(0);
    a.Qin = s->Q[s->cur].comp(0);
    a.Qout = s->Q[s->cur].comp(0);     // in place: the RHS is already materialised
    a.stage = stage;
    a.dt = dt;
    { k_rk4<<<nblocks(N), 256, 0, st>>>(a); mg_count_launches(1); }
    s->dependentValid = false;
    s->fusedValid = false;
  } else if (mode == MG_ADJOINT) {
    const double factor[5] = {0.0, 1.0, 0.5, 1.0, 2.0};
    s->adjointForcingFactor = factor[stage];
    if (stage == 4) s->timeProgressive = *time - dt / 2.0;
    if (s->useFused && mg_fused_supported(s, MG_ADJOINT)) {
      if (!s->fusedValid) MG_TRY(mg_fused_sweepA(s));
      MG_TRY(mg_fused_adjoint1(s));
      MG_TRY(mg_fused_adjoint2(s, 1, stage, dt));
      if (stage == 4 || stage == 2) s->timeProgressive = *time;
      if (stage == 3 || stage == 1) { *time -= dt / 2.0; s->time = *time; }
      return 0;
    }
    if (!s->rhsReady) MG_TRY(mg_state_compute_rhs_impl(s, MG_ADJOINT));
    s->rhsReady = false;
    MG_TRY(mg_state_make_exclusive(s, &s->W[s->curW], true));
    MG_TRY(mg_state_make_exclusive(s, &s->rk1, true));
    a.b1 = s->rk1.comp(0);
    a.R = s->rhs.comp(0);
    a.Qin = s->W[s->curW].comp(0);
    a.Qout = s->W[s->curW].comp(0);
    a.stage = 5 - stage;               // adjoint stage 4 plays the role of RK stage 1, etc.
    a.dt = -dt;
    { k_rk4<<<nblocks(N), 256, 0, st>>>(a); mg_count_launches(1); }
    if (stage == 4 || stage == 2) s->timeProgressive = *time;
    if (stage == 3 || stage == 1) { *time -= dt / 2.0; s->time = *time; }
  } else if (mode == MG_LINEARIZED) {
    // substepLinearizedRK4 (reference src/RK4IntegratorImpl.f90:272-369): the forward scheme on the perturbation
    if (stage == 1) s->timeProgressive = *time + dt / 2.0;
    if (stage == 2 || stage == 4) { *time += dt / 2.0; s->time = *time; }
    if (stage == 3) s->timeProgressive = *time + dt / 2.0;
    if (!s->rhsReady) MG_TRY(mg_state_compute_rhs_impl(s, MG_LINEARIZED));
    s->rhsReady = false;
    MG_TRY(mg_state_make_exclusive(s, &s->W[s->curW], true));
    MG_TRY(mg_state_make_exclusive(s, &s->rk1, true));
    a.b1 = s->rk1.comp(0);
    a.R = s->rhs.comp(0);
    a.Qin = s->W[s->curW].comp(0);
    a.Qout = s->W[s->curW].comp(0);
    a.stage = stage;
    a.dt = dt;
    { k_rk4<<<nblocks(N), 256, 0, st>>>(a); mg_count_launches(1); }
  } else {
    MG_FAIL("rk4 substep: unknown mode");
  }
  MG_CUDA(cudaGetLastError());
  return 0;
}
