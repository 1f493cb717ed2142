// SAT penalty / sponge / forcing patches on the device (t_Patch family).
// Reference: src/PatchImpl.f90:3-151 (extents), src/FarFieldPatchImpl.f90:93-286,
// src/RhsHelperImpl.f90:89-250 (viscous far-field adjoint sources), src/SpongePatchImpl.f90:65-165,
// src/ImpenetrableWallImpl.f90:60-207, src/IsothermalWallImpl.f90:103-337,
// src/CostTargetPatchImpl.f90:72-136, src/ActuatorPatchImpl.f90:108-181.
//
// One thread per patch point; patch points are numbered in the patch-local Fortran order of the
// reference (i fastest inside the part of the patch owned by this rank).
#include <cfloat>
#include <cstring>

#include "grid.h"
#include "patches.h"

namespace {

inline unsigned nblocks(size_t n) { return (unsigned)((n + 127) / 128); }

struct PatchGeom {
  int lo[3], sz[3];      // local 0-based start inside the rank's brick, local patch extents
  int nx, ny;
  int n;
  __device__ size_t gridIndex(int q) const {
    const int i = q % sz[0], j = (q / sz[0]) % sz[1], k = q / (sz[0] * sz[1]);
    return (size_t)(lo[0] + i) + (size_t)nx * ((size_t)(lo[1] + j) + (size_t)ny * (size_t)(lo[2] + k));
  }
};

struct FFArgs {
  PatchGeom g;
  const double *Q, *W, *target, *m, *jac, *v, *u, *T, *tau, *q;
  const int* iblank;
  double* rhs;
  const double* Aplus;          // (NU*NU) per patch point, row-major A[i][j], point fastest
  const double *Fv, *FvTarget;  // (NU*ND) per patch point: component c + NU*l, point fastest
  size_t csQ, csW, cs;
  int dir, mode, viscous, continuousAdjoint;
  double sigmaI, sigmaV, gamma, powerLaw;
};

template <int ND>
__global__ void k_farfield_setup(PatchGeom g, const double* target, size_t cs, const double* m, size_t csm,
                                 int dir, double gamma, int incoming, double* Aplus) {
  constexpr int NU = ND + 2;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= g.n) return;
  const size_t p = g.gridIndex(q);
  double Q[NU], mm[ND], A[NU][NU];
#pragma unroll
  for (int c = 0; c < NU; ++c) Q[c] = target[(size_t)c * cs + p];
#pragma unroll
  for (int i = 0; i < ND; ++i) mm[i] = m[(size_t)(i + ND * dir) * csm + p];
  incoming_jacobian<ND>(Q, mm, gamma, incoming, A);
#pragma unroll
  for (int i = 0; i < NU; ++i)
#pragma unroll
    for (int j = 0; j < NU; ++j) Aplus[(size_t)(i * NU + j) * g.n + q] = A[i][j];
}

// addFarFieldPenalty (reference src/FarFieldPatchImpl.f90:93-286)
template <int ND>
__global__ void k_farfield(FFArgs a) {
  constexpr int NU = ND + 2;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  const size_t p = a.g.gridIndex(q);
  if (a.iblank && a.iblank[p] == 0) return;
  const double jac = a.jac[p];
  double mm[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) mm[i] = a.m[(size_t)(i + ND * a.dir) * a.cs + p];
  double r[NU];
  if (a.mode == MG_FORWARD) {
    double dq[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) dq[c] = a.Q[(size_t)c * a.csQ + p] - a.target[(size_t)c * a.cs + p];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NU; ++j) acc += a.Aplus[(size_t)(i * NU + j) * a.g.n + q] * dq[j];
      r[i] = -a.sigmaI * jac * acc;
    }
    if (a.viscous) {
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < ND; ++l)
          acc += (a.Fv[(size_t)(c + NU * l) * a.g.n + q] - a.FvTarget[(size_t)(c + NU * l) * a.g.n + q]) * mm[l];
        r[c] += a.sigmaV * jac * acc;
      }
    }
  } else if (a.mode == MG_LINEARIZED) {
    // reference :258-268: - sigmaI (1/J) A+ dQ + sigmaV (1/J) (linearized contravariant viscous flux along dir)
    double dq[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) dq[c] = a.W[(size_t)c * a.csW + p];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NU; ++j) acc += a.Aplus[(size_t)(i * NU + j) * a.g.n + q] * dq[j];
      r[i] = -a.sigmaI * jac * acc;
    }
    if (a.viscous) {
#pragma unroll
      for (int c = 0; c < NU; ++c) r[c] += a.sigmaV * jac * a.Fv[(size_t)(c + NU * a.dir) * a.g.n + q];
    }
  } else {
    double w[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) w[c] = a.W[(size_t)c * a.csW + p];
    const double sgn = a.continuousAdjoint ? -1.0 : 1.0;
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < NU; ++i) acc += a.Aplus[(size_t)(i * NU + j) * a.g.n + q] * w[i];
      r[j] = sgn * a.sigmaI * jac * acc;
    }
    if (a.viscous) {
      // - sigmaV (1/J) B^T w with B the first-partial viscous Jacobian: obtained from the shared
      // helper as  -(A - B)^T w + A^T w  would waste work, so evaluate (A-B)^T and A^T explicitly.
      double Q[NU], tau[ND * ND], qq[ND], y1[NU], y2[NU];
      Prim<ND> s;
#pragma unroll
      for (int c = 0; c < NU; ++c) { Q[c] = a.Q[(size_t)c * a.csQ + p]; y1[c] = 0.0; y2[c] = 0.0; }
      s.v = a.v[p];
      s.T = a.T[p];
#pragma unroll
      for (int i = 0; i < ND; ++i) { s.u[i] = a.u[(size_t)i * a.cs + p]; qq[i] = a.q[(size_t)i * a.cs + p]; }
#pragma unroll
      for (int c = 0; c < ND * ND; ++c) tau[c] = a.tau[(size_t)c * a.cs + p];
      add_flux_jacobian_transpose<ND>(Q, s, mm, a.gamma, true, a.powerLaw, tau, qq, w, y1);    // (A-B)^T w
      add_flux_jacobian_transpose<ND>(Q, s, mm, a.gamma, false, a.powerLaw, tau, qq, w, y2);   // A^T w
#pragma unroll
      for (int c = 0; c < NU; ++c) r[c] -= a.sigmaV * jac * (y2[c] - y1[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < NU; ++c) a.rhs[(size_t)c * a.cs + p] += r[c];
}

struct FFSrcArgs {
  PatchGeom g;
  const double *W, *u, *mu, *lam, *kap, *m, *jac;
  const int* iblank;
  double* temp1;     // component c + (NU-1)*l
  size_t csW, cs;
  int dir;
  double sigmaV;
};

// Source of addFarFieldAdjointPenalty (reference src/RhsHelperImpl.f90:137-207)
template <int ND>
__global__ void k_farfield_adjoint_source(FFSrcArgs a) {
  constexpr int NU = ND + 2;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  const size_t p = a.g.gridIndex(q);
  if (a.iblank && a.iblank[p] == 0) return;
  double u[ND], M[ND * ND], w[ND + 1];
#pragma unroll
  for (int i = 0; i < ND; ++i) u[i] = a.u[(size_t)i * a.cs + p];
#pragma unroll
  for (int c = 0; c < ND * ND; ++c) M[c] = a.m[(size_t)c * a.cs + p];
#pragma unroll
  for (int c = 0; c < ND + 1; ++c) w[c] = a.W[(size_t)(c + 1) * a.csW + p];
  const double mu = a.mu[p], lam = a.lam[p], kap = a.kap[p], jac = a.jac[p];
#pragma unroll
  for (int l = 0; l < ND; ++l) {
    double d[ND + 1];
#pragma unroll
    for (int c = 0; c < ND + 1; ++c) d[c] = 0.0;
    add_second_partial_transpose<ND>(u, mu, lam, kap, jac, &M[ND * a.dir], &M[ND * l], w, d);
#pragma unroll
    for (int c = 0; c < ND + 1; ++c) a.temp1[(size_t)(c + (NU - 1) * l) * a.cs + p] -= a.sigmaV * d[c];
  }
}

struct SpongeArgs {
  PatchGeom g;
  const double *X, *target, *strength;
  const int* iblank;
  double* rhs;
  size_t csX, cs;
  int nU, mode;
};

// addDamping (reference src/SpongePatchImpl.f90:65-165)
__global__ void k_sponge(SpongeArgs a) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  const size_t p = a.g.gridIndex(q);
  if (a.iblank && a.iblank[p] == 0) return;
  const double s = a.strength[q];
  for (int c = 0; c < a.nU; ++c) {
    if (a.mode == MG_FORWARD)
      a.rhs[(size_t)c * a.cs + p] -= s * (a.X[(size_t)c * a.csX + p] - a.target[(size_t)c * a.cs + p]);
    else if (a.mode == MG_LINEARIZED)
      a.rhs[(size_t)c * a.cs + p] -= s * a.X[(size_t)c * a.csX + p];
    else
      a.rhs[(size_t)c * a.cs + p] += s * a.X[(size_t)c * a.csX + p];
  }
}

struct WallArgs {
  PatchGeom g;
  const double *Q, *W, *m, *jac, *v, *u, *pr, *T, *wallT;
  const int* iblank;
  double* rhs;
  size_t csQ, csW, cs;
  int dir, mode, isothermal, viscous;
  double sigmaI, sigmaV1, gamma;
};

// addImpenetrableWallPenalty (reference src/ImpenetrableWallImpl.f90:60-207) followed, for isothermal
// walls, by addIsothermalWallPenalty (src/IsothermalWallImpl.f90:103-337).
template <int ND>
__global__ void k_wall(WallArgs a) {
  constexpr int NU = ND + 2;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  const size_t p = a.g.gridIndex(q);
  if (a.iblank && a.iblank[p] == 0) return;
  const double jac = a.jac[p];
  double Q[NU], mm[ND], u[ND], r[NU];
#pragma unroll
  for (int c = 0; c < NU; ++c) Q[c] = a.Q[(size_t)c * a.csQ + p];
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    mm[i] = a.m[(size_t)(i + ND * a.dir) * a.cs + p];
    u[i] = a.u[(size_t)i * a.cs + p];
  }
  const double v = a.v[p];
  if (a.mode == MG_FORWARD) {
    double nm = 0.0;
#pragma unroll
    for (int l = 0; l < ND; ++l) nm = (l == 0) ? Q[1] * mm[0] : nm + Q[l + 1] * mm[l];
    r[0] = nm;
#pragma unroll
    for (int i = 0; i < ND; ++i) r[i + 1] = nm * u[i];
    r[NU - 1] = nm * v * (Q[NU - 1] + a.pr[p]);
#pragma unroll
    for (int c = 0; c < NU; ++c) r[c] = -a.sigmaI * jac * r[c];
    if (a.isothermal && a.viscous) {
      double pen[NU];
      pen[0] = 0.0;
#pragma unroll
      for (int c = 1; c < NU; ++c) pen[c] = Q[c];
      pen[NU - 1] = pen[NU - 1] - Q[0] * a.wallT[q] / a.gamma;
#pragma unroll
      for (int c = 0; c < NU; ++c) r[c] -= a.sigmaV1 * (jac * pen[c]);
    }
  } else {
    double w[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) { w[c] = a.W[(size_t)c * a.csW + p]; r[c] = 0.0; }
    // deltaInviscidPenalty = A(Q, m) with the pressure rows removed (reference :143-172)
    Prim<ND> s;
    s.v = v;
    s.T = a.T[p];
#pragma unroll
    for (int i = 0; i < ND; ++i) s.u[i] = v * Q[i + 1];     // velocity not passed: recomputed (:150-166)
    double y[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) y[c] = 0.0;
    if (a.mode == MG_LINEARIZED) add_flux_jacobian_apply<ND>(s, mm, a.gamma, false, 0.0, nullptr, nullptr, w, y);
    else add_flux_jacobian_transpose<ND>(Q, s, mm, a.gamma, false, 0.0, nullptr, nullptr, w, y);
    double dp[NU];
    double usq = 0.0;
#pragma unroll
    for (int i = 0; i < ND; ++i) usq = (i == 0) ? u[0] * u[0] : usq + u[i] * u[i];
    dp[0] = 0.5 * usq;
#pragma unroll
    for (int i = 0; i < ND; ++i) dp[i + 1] = -u[i];
    dp[NU - 1] = 1.0;
    if (a.mode == MG_LINEARIZED) {
      // - sigmaI (1/J) (A - [0; m dp; 0]) dQ  (reference :187-191), w holds dQ
      double dpw = 0.0;
#pragma unroll
      for (int c = 0; c < NU; ++c) dpw += (dp[c] * (a.gamma - 1.0)) * w[c];
#pragma unroll
      for (int l = 0; l < ND; ++l) y[l + 1] -= mm[l] * dpw;
#pragma unroll
      for (int c = 0; c < NU; ++c) r[c] = -a.sigmaI * jac * y[c];
      if (a.isothermal && a.viscous) {
        double pen[NU];
        pen[0] = 0.0;
#pragma unroll
        for (int c = 1; c < NU; ++c) pen[c] = w[c];
        pen[NU - 1] = pen[NU - 1] - w[0] * a.wallT[q] / a.gamma;
#pragma unroll
        for (int c = 0; c < NU; ++c) r[c] -= a.sigmaV1 * (jac * pen[c]);
      }
#pragma unroll
      for (int c = 0; c < NU; ++c) a.rhs[(size_t)c * a.cs + p] += r[c];
      return;
    }
    double mw = 0.0;
#pragma unroll
    for (int l = 0; l < ND; ++l) mw += mm[l] * w[l + 1];
#pragma unroll
    for (int c = 0; c < NU; ++c) {
      y[c] -= (dp[c] * (a.gamma - 1.0)) * mw;
      r[c] = a.sigmaI * jac * y[c];
    }
    if (a.isothermal && a.viscous) {
      double ap[NU];
      ap[0] = -w[NU - 1] * a.wallT[q] / a.gamma;
#pragma unroll
      for (int c = 1; c < NU; ++c) ap[c] = w[c];
#pragma unroll
      for (int c = 0; c < NU; ++c) r[c] += a.sigmaV1 * (jac * ap[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < NU; ++c) a.rhs[(size_t)c * a.cs + p] += r[c];
}

struct AddArgs {
  PatchGeom g;
  const double* data;      // (nComp) per patch point, point fastest
  const double* mollifier; // grid field or null
  const int* iblank;
  double* rhs;
  size_t cs;
  int nComp;
  double factor;
};

// addAdjointForcing / updateActuatorPatch (reference CostTargetPatchImpl.f90:72-136, ActuatorPatchImpl.f90:108-181)
__global__ void k_patch_add(AddArgs a) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  const size_t p = a.g.gridIndex(q);
  if (a.iblank && a.iblank[p] == 0) return;
  const double f = a.mollifier ? a.factor * a.mollifier[p] : a.factor;
  for (int c = 0; c < a.nComp; ++c) a.rhs[(size_t)c * a.cs + p] += f * a.data[(size_t)c * a.g.n + q];
}

// addKolmogorovForcing (reference src/KolmogorovForcingPatchImpl.f90:86-176): a body force per unit mass along x
struct KolmogorovArgs {
  PatchGeom g;
  const double* force;      // forcePerUnitMass at the patch points
  const double *Q, *W;
  size_t csQ, csW, cs;
  const int* iblank;
  double* rhs;
  int mode;
};

__global__ void k_kolmogorov(KolmogorovArgs a) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  const size_t p = a.g.gridIndex(q);
  if (a.iblank && a.iblank[p] == 0) return;
  const double f = a.force[q];
  if (a.mode == MG_FORWARD) a.rhs[a.cs + p] += a.Q[p] * f;
  else if (a.mode == MG_ADJOINT) a.rhs[p] -= a.W[a.csW + p] * f;
  else a.rhs[a.cs + p] += a.W[p] * f;
}

// forcePerUnitMass = amplitude sin(2 pi n y) (src/KolmogorovForcingPatchImpl.f90:47-66)
__global__ void k_kolmogorov_setup(PatchGeom g, const double* y, const int* iblank, double amplitude, double twoPiN,
                                   double* out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= g.n) return;
  const size_t p = g.gridIndex(q);
  out[q] = (iblank && iblank[p] == 0) ? 0.0 : amplitude * sin(twoPiN * y[p]);
}

// addJetExcitation (reference src/JetExcitationPatchImpl.f90:128-187)
struct JetArgs {
  PatchGeom g;
  const double *strength, *re, *im;   // re / im: (nPatchPoints, nUnknowns, nModes), point fastest
  const int* iblank;
  double* rhs;
  size_t cs;
  int nU, nModes;
  double c[MG_JET_MAX_MODES], s[MG_JET_MAX_MODES];
};

__global__ void k_jet_excitation(const __grid_constant__ JetArgs a) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  const size_t p = a.g.gridIndex(q);
  if (a.iblank && a.iblank[p] == 0) return;
  const double st = a.strength[q];
  for (int c = 0; c < a.nU; ++c) {
    double r = a.rhs[(size_t)c * a.cs + p];
    for (int l = 0; l < a.nModes; ++l) {
      const size_t e = ((size_t)l * a.nU + c) * a.g.n + q;
      r -= st * (a.re[e] * a.c[l] - a.im[e] * a.s[l]);
    }
    a.rhs[(size_t)c * a.cs + p] = r;
  }
}

__global__ void k_collect(PatchGeom g, const double* field, size_t cs, int nComp, double* out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= g.n) return;
  const size_t p = g.gridIndex(q);
  for (int c = 0; c < nComp; ++c) out[(size_t)c * g.n + q] = field[(size_t)c * cs + p];
}

__global__ void k_disperse(PatchGeom g, const double* in, size_t cs, int nComp, double* field) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= g.n) return;
  const size_t p = g.gridIndex(q);
  for (int c = 0; c < nComp; ++c) field[(size_t)c * cs + p] = in[(size_t)c * g.n + q];
}

// computeSpongeStrengths (reference src/PatchFactoryImpl.f90:161-374): fraction of the arc length, measured along
// the sponge's direction from its inner edge, raised to sponge_exponent and scaled by sponge_amount.
struct SpongeSetupArgs {
  PatchGeom g;
  const double* arc;     // sqrt(sum (d coordinates / d xi_dir)^2): the rank's grid field, or the lines gathered along dir
  int n0, n1;            // first two sizes of the arc array (the local sizes, the global one along a gathered direction)
  int dir, eLo, eHi;     // direction, 0-based extent [eLo, eHi] of the whole patch along it, in the arc array's index
  int cOffset;           // index of the rank's first point along dir in the arc array (0 unless gathered)
  int normalDirection, exponent;
  double amount;
  double* out;
};

__global__ void k_sponge_strength(SpongeSetupArgs a) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  int cl[3] = {a.g.lo[0] + q % a.g.sz[0], a.g.lo[1] + (q / a.g.sz[0]) % a.g.sz[1], a.g.lo[2] + q / (a.g.sz[0] * a.g.sz[1])};
  const int c = cl[a.dir] + a.cOffset;
  cl[a.dir] = 0;
  const long stride = a.dir == 0 ? 1 : (a.dir == 1 ? (long)a.n0 : (long)a.n0 * a.n1);
  const double* line = a.arc + ((long)cl[0] + (long)a.n0 * ((long)cl[1] + (long)a.n1 * cl[2]));   // coordinate 0 of this line
  double num = 0.0, den = 0.0;
  if (a.normalDirection > 0) {
    for (int l = a.eLo; l <= c - 1; ++l) num += line[(long)l * stride];
    for (int l = a.eLo; l <= a.eHi - 1; ++l) den += line[(long)l * stride];
  } else {
    for (int l = c + 1; l <= a.eHi; ++l) num += line[(long)l * stride];
    for (int l = a.eLo + 1; l <= a.eHi; ++l) den += line[(long)l * stride];
  }
  a.out[q] = a.amount * pow(1.0 - num / den, (double)a.exponent);
}

__global__ void k_arc_from_derivatives(const double* d, size_t cs, int nD, double* arc, size_t N) {
  size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= N) return;
  double s = 0.0;
  for (int i = 0; i < nD; ++i) s += d[(size_t)i * cs + p] * d[(size_t)i * cs + p];
  arc[p] = sqrt(s);
}

PatchGeom geom(const mg_patch* pt) {
  PatchGeom g;
  for (int i = 0; i < 3; ++i) { g.lo[i] = pt->localLo[i]; g.sz[i] = pt->localSize[i]; }
  g.nx = pt->state->grid->localSize[0];
  g.ny = pt->state->grid->localSize[1];
  g.n = pt->nPatchPoints;
  return g;
}

template <typename F>
int dispatch_nd(int nD, F f) {
  if (nD == 1) return f(std::integral_constant<int, 1>());
  if (nD == 2) return f(std::integral_constant<int, 2>());
  return f(std::integral_constant<int, 3>());
}

}  // namespace

// setupPatch (reference src/PatchImpl.f90:3-151): intersect the global extent with this rank's brick.
int mg_patch_create_impl(mg_state* s, int type, const char* name, int normalDirection, const int extent[6],
                         mg_patch** out) {
  mg_grid* g = s->grid;
  auto* p = new mg_patch();
  p->state = s;
  p->type = type;
  p->name = name ? name : "";
  p->normalDirection = normalDirection;
  for (int i = 0; i < 6; ++i) p->extent[i] = extent[i];
  bool empty = false;
  for (int d = 0; d < 3; ++d) {
    const int lo = extent[2 * d], hi = extent[2 * d + 1];     // 1-based inclusive, global
    if (lo < 1 || hi > g->globalSize[d] || lo > hi) { delete p; MG_FAIL("mg_patch_create: invalid extent"); }
    p->globalSize[d] = hi - lo + 1;
    const int a = std::max(lo, g->offset[d] + 1), b = std::min(hi, g->offset[d] + g->localSize[d]);
    if (b < a) empty = true;
    p->localLo[d] = a - 1 - g->offset[d];
    p->localSize[d] = b - a + 1;
    p->patchOffset[d] = a - lo;          // offset of the local part inside the global patch
  }
  if (empty) { for (int d = 0; d < 3; ++d) { p->localLo[d] = 0; p->localSize[d] = 0; } }
  p->nPatchPoints = p->localSize[0] * p->localSize[1] * p->localSize[2];
  const int ad = std::abs(normalDirection);
  if (type < MG_PATCH_FARFIELD || type > MG_PATCH_ADIABATIC_WALL) { delete p; MG_FAIL("mg_patch_create: unknown patch type"); }
  if (type != MG_PATCH_SPONGE && type != MG_PATCH_ACTUATOR && type != MG_PATCH_COST_TARGET &&
      type != MG_PATCH_KOLMOGOROV_FORCING && type != MG_PATCH_JET_EXCITATION && type != MG_PATCH_PROBE) {
    if (ad < 1 || ad > g->nD) { delete p; MG_FAIL("mg_patch_create: normal direction is invalid"); }
    if (extent[2 * (ad - 1)] != extent[2 * (ad - 1) + 1]) {
      delete p;
      MG_FAIL("mg_patch_create: patch extends more than 1 grid point along normal direction");
    }
  }
  if (type == MG_PATCH_KOLMOGOROV_FORCING) {     // verifyKolmogorovForcingPatchUsage (:178-237)
    if (g->nD == 1) { delete p; MG_FAIL("mg_patch_create: KOLMOGOROV_FORCING can't be used with a 1D grid"); }
    for (int d = 0; d < g->nD; ++d)
      if (extent[2 * d] == extent[2 * d + 1]) { delete p; MG_FAIL("mg_patch_create: KOLMOGOROV_FORCING expects a patch of the grid's dimension"); }
  }
  if (type == MG_PATCH_FARFIELD || type == MG_PATCH_SPONGE || type == MG_PATCH_JET_EXCITATION) {
    if (!s->opt.useTargetState) { delete p; MG_FAIL("mg_patch_create: no target state available for this patch type"); }
  }
  s->patches.push_back(p);
  if ((type == MG_PATCH_FARFIELD || type == MG_PATCH_BLOCK_INTERFACE) && s->opt.viscosityOn) s->keepViscousFluxes = true;
  *out = p;
  return 0;
}

void mg_patch_destroy_impl(mg_patch* p) {
  if (!p) return;
  for (auto& kv : p->arrays) cudaFree(kv.second.p);
  cudaFree(p->probeBuffer);
  cudaFree(p->gradientBuffer);
  if (p->remote) mg_p2p_destroy(p->remote);
  delete p;
}

int mg_patch_alloc_array(mg_patch* p, const std::string& name, int nComp, double** out) {
  auto it = p->arrays.find(name);
  if (it != p->arrays.end() && it->second.nComp == nComp) { *out = it->second.p; return 0; }
  if (it != p->arrays.end()) { cudaFree(it->second.p); p->arrays.erase(it); }
  mg_patch::Array a;
  a.nComp = nComp;
  const size_t bytes = sizeof(double) * (size_t)std::max(1, p->nPatchPoints) * nComp;
  MG_CUDA(cudaMalloc(&a.p, bytes));
  MG_CUDA(cudaMemsetAsync(a.p, 0, bytes, mg_stream()));
  p->arrays[name] = a;
  *out = a.p;
  return 0;
}

int mg_patch_set_array_impl(mg_patch* p, const char* name, int nComp, const double* host) {
  double* d = nullptr;
  MG_TRY(mg_patch_alloc_array(p, name, nComp, &d));
  if (p->nPatchPoints > 0)
    MG_CUDA(cudaMemcpyAsync(d, host, sizeof(double) * (size_t)p->nPatchPoints * nComp, cudaMemcpyDefault, mg_stream()));
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  if (std::string(name) == "Aplus") p->AplusReady = true;
  return 0;
}

int mg_patch_get_array_impl(mg_patch* p, const char* name, int nComp, double* host) {
  auto it = p->arrays.find(name);
  if (it == p->arrays.end() || it->second.nComp != nComp) MG_FAIL(std::string("mg_patch_get_array: no array '") + name + "'");
  if (p->nPatchPoints > 0)
    MG_CUDA(cudaMemcpyAsync(host, it->second.p, sizeof(double) * (size_t)p->nPatchPoints * nComp, cudaMemcpyDefault, mg_stream()));
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  return 0;
}

// collect (reference src/PatchImpl.f90:187-316): grid field -> patch array
int mg_patch_collect_impl(mg_patch* p, const MgField* f, int nComp, const char* name) {
  double* d = nullptr;
  MG_TRY(mg_patch_alloc_array(p, name, nComp, &d));
  if (p->nPatchPoints == 0) return 0;
  { k_collect<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(geom(p), f->comp(0), f->compStride, nComp, d); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return 0;
}

int mg_patch_disperse_impl(mg_patch* p, const char* name, int nComp, MgField* f) {
  auto it = p->arrays.find(name);
  if (it == p->arrays.end() || it->second.nComp != nComp) MG_FAIL(std::string("mg_patch_disperse: no array '") + name + "'");
  if (p->nPatchPoints == 0) return 0;
  { k_disperse<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(geom(p), it->second.p, f->compStride, nComp, f->comp(0)); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return 0;
}

bool mg_patches_have_farfield(const mg_state* s) {
  for (const mg_patch* p : s->patches)
    if (p->type == MG_PATCH_FARFIELD) return true;
  return false;
}

int mg_patches_collect_viscous(mg_state* s) {
  for (mg_patch* p : s->patches)
    if (p->type == MG_PATCH_FARFIELD || p->type == MG_PATCH_BLOCK_INTERFACE)
      MG_TRY(mg_patch_collect_impl(p, &s->viscFluxCart, s->nU * s->nD, "viscousFluxes"));
  return 0;
}

// updatePatchFactories, far-field part (reference src/PatchFactoryImpl.f90:531-560): target viscous
// fluxes from the target state.  NB: like the reference this leaves the state's dependent variables
// evaluated at the target state; callers update the state afterwards.
int mg_patches_update_impl(mg_state* s) {
  if (!(s->opt.viscosityOn && s->opt.useTargetState) || !mg_patches_have_farfield(s)) return 0;
  MG_TRY(mg_state_update_impl(s, &s->target));
  // Cartesian viscous fluxes of the target state via the general flux kernel
  const bool keep = s->keepViscousFluxes;
  s->keepViscousFluxes = true;
  MgField saveQ = s->Q[s->cur];
  s->Q[s->cur] = s->target;             // shallow view swap: k_flux reads Q only for the inviscid part
  s->Q[s->cur].owned = false;
  int rc = mg_state_rhs_forward_general(s);
  s->Q[s->cur] = saveQ;
  s->keepViscousFluxes = keep;
  MG_TRY(rc);
  for (mg_patch* p : s->patches)
    if (p->type == MG_PATCH_FARFIELD)
      MG_TRY(mg_patch_collect_impl(p, &s->viscFluxCart, s->nU * s->nD, "targetViscousFluxes"));
  s->dependentValid = false;
  return 0;
}

static bool sponge_along(const mg_state* s, int dir) {
  for (const mg_patch* p : s->patches)
    if ((p->type == MG_PATCH_SPONGE || p->type == MG_PATCH_JET_EXCITATION) && std::abs(p->normalDirection) == dir + 1) return true;
  return false;
}

// local arc length along dir (src/PatchFactoryImpl.f90:213-218)
static int sponge_arc_field(mg_state* s, int dir, MgField* arc) {
  mg_grid* g = s->grid;
  MgField cd;
  MG_TRY(mg_field_alloc(g, s->nD, &cd));
  MG_TRY(mg_field_alloc(g, 1, arc));
  MG_TRY(mg_grid_coordinate_derivatives(g, dir, &cd));
  { k_arc_from_derivatives<<<(unsigned)((g->N + 255) / 256), 256, 0, mg_stream()>>>(cd.comp(0), cd.compStride, s->nD, arc->comp(0), g->N); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  mg_field_free(&cd);
  return 0;
}

// strengths of the sponges along dir from arc lengths whose lines hold `lineLength` points along dir, the rank's first
// point being number cOffset of them (src/PatchFactoryImpl.f90:232-363)
static int sponge_strengths_along(mg_state* s, int dir, const double* arc, int lineLength, int cOffset) {
  mg_grid* g = s->grid;
  for (mg_patch* p : s->patches) {
    if ((p->type != MG_PATCH_SPONGE && p->type != MG_PATCH_JET_EXCITATION) || std::abs(p->normalDirection) != dir + 1 ||
        p->nPatchPoints <= 0) continue;
    double* out = nullptr;
    MG_TRY(mg_patch_alloc_array(p, "spongeStrength", 1, &out));
    SpongeSetupArgs a;
    a.g = geom(p);
    a.arc = arc;
    a.n0 = dir == 0 ? lineLength : g->localSize[0];
    a.n1 = dir == 1 ? lineLength : g->localSize[1];
    a.dir = dir;
    a.cOffset = cOffset;
    a.eLo = p->extent[2 * dir] - 1 - g->offset[dir] + cOffset;
    a.eHi = p->extent[2 * dir + 1] - 1 - g->offset[dir] + cOffset;
    a.normalDirection = p->normalDirection;
    a.exponent = p->spongeExponent;
    a.amount = p->spongeAmount;
    a.out = out;
    { k_sponge_strength<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(a); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
  }
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  return 0;
}

int mg_patches_sponge_strengths_impl(mg_state* s) {
  mg_grid* g = s->grid;
  for (int dir = 0; dir < s->nD; ++dir) {
    if (!sponge_along(s, dir)) continue;
    if (g->procDims[dir] > 1)
      MG_FAIL("computeSpongeStrengths: sponges along a decomposed direction need the arc length of the whole line: "
              "gather mg_state_sponge_arc_length along the direction and call mg_state_sponge_strengths_gathered");
    MgField arc;
    MG_TRY(sponge_arc_field(s, dir, &arc));
    const int rc = sponge_strengths_along(s, dir, arc.comp(0), g->localSize[dir], 0);
    mg_field_free(&arc);
    if (rc) return rc;
  }
  return 0;
}

int mg_patches_sponge_arc_length_impl(mg_state* s, int dir, double* hostOut) {
  MgField arc;
  MG_TRY(sponge_arc_field(s, dir, &arc));
  const cudaError_t e = cudaMemcpy(hostOut, arc.comp(0), s->grid->N * sizeof(double), cudaMemcpyDeviceToHost);
  mg_field_free(&arc);
  MG_CUDA(e);
  return 0;
}

int mg_patches_sponge_strengths_gathered_impl(mg_state* s, int dir, const double* arcGathered) {
  mg_grid* g = s->grid;
  if (!sponge_along(s, dir)) return 0;
  const size_t n = g->N / (size_t)g->localSize[dir] * (size_t)g->globalSize[dir];
  double* d = nullptr;
  MG_CUDA(cudaMalloc(&d, n * sizeof(double)));
  cudaError_t e = cudaMemcpy(d, arcGathered, n * sizeof(double), cudaMemcpyHostToDevice);
  int rc = 0;
  if (e == cudaSuccess) rc = sponge_strengths_along(s, dir, d, g->globalSize[dir], g->offset[dir]);
  cudaFree(d);
  MG_CUDA(e);
  return rc;
}

int mg_patches_farfield_adjoint_sources(mg_state* s, MgField* temp1) {
  mg_grid* g = s->grid;
  for (mg_patch* p : s->patches) {
    if (p->type != MG_PATCH_FARFIELD || p->nPatchPoints == 0) continue;
    FFSrcArgs a;
    a.g = geom(p);
    a.W = s->W[s->curW].comp(0);
    a.csW = s->W[s->curW].compStride;
    a.u = s->velocity.comp(0);
    a.mu = s->mu.comp(0);
    a.lam = s->lambda.comp(0);
    a.kap = s->kappa.comp(0);
    a.m = g->metrics.comp(0);
    a.jac = g->jacobian.comp(0);
    a.iblank = g->iblank;
    a.temp1 = temp1->comp(0);
    a.cs = temp1->compStride;
    a.dir = std::abs(p->normalDirection) - 1;
    a.sigmaV = p->viscousPenaltyAmount;
    MG_TRY(dispatch_nd(s->nD, [&](auto nd) {
      { k_farfield_adjoint_source<decltype(nd)::value><<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(a); mg_count_launches(1); }
      return 0;
    }));
    MG_CUDA(cudaGetLastError());
  }
  return 0;
}

// patch%updateRhs for every patch of the state, in creation (bc.dat) order
// (reference src/RegionImpl.f90:1985-1995)
int mg_patches_apply(mg_state* s, int mode) {
  mg_grid* g = s->grid;
  cudaStream_t st = mg_stream();
  const MgField& Q = s->Q[s->cur];
  const MgField& W = s->W[s->curW];
  for (mg_patch* p : s->patches) {
    if (p->nPatchPoints == 0) continue;
    const int dir = std::abs(p->normalDirection) - 1;
    switch (p->type) {
      case MG_PATCH_FARFIELD: {
        double* Aplus = nullptr;
        MG_TRY(mg_patch_alloc_array(p, "Aplus", s->nU * s->nU, &Aplus));
        const int incoming = (mode == MG_ADJOINT && s->opt.useContinuousAdjoint) ? -p->normalDirection : p->normalDirection;
        if (!p->AplusReady || p->AplusIncoming != incoming) {
          MG_TRY(dispatch_nd(s->nD, [&](auto nd) {
            { k_farfield_setup<decltype(nd)::value><<<nblocks(p->nPatchPoints), 128, 0, st>>>(
                geom(p), s->target.comp(0), s->target.compStride, g->metrics.comp(0), g->metrics.compStride, dir,
                s->opt.ratioOfSpecificHeats, incoming, Aplus); mg_count_launches(1); }
            return 0;
          }));
          p->AplusReady = true;
          p->AplusIncoming = incoming;
        }
        FFArgs a;
        std::memset(&a, 0, sizeof(a));
        a.g = geom(p);
        a.Q = Q.comp(0); a.csQ = Q.compStride;
        a.W = W.comp(0); a.csW = W.compStride;
        a.target = s->target.comp(0);
        a.m = g->metrics.comp(0);
        a.jac = g->jacobian.comp(0);
        a.v = s->specificVolume.comp(0);
        a.u = s->velocity.comp(0);
        a.T = s->temperature.comp(0);
        a.iblank = g->iblank;
        a.rhs = s->rhs.comp(0);
        a.cs = s->rhs.compStride;
        a.Aplus = Aplus;
        a.dir = dir;
        a.mode = mode;
        a.viscous = s->opt.viscosityOn;
        a.continuousAdjoint = s->opt.useContinuousAdjoint;
        a.sigmaI = p->inviscidPenaltyAmount;
        a.sigmaV = p->viscousPenaltyAmount;
        a.gamma = s->opt.ratioOfSpecificHeats;
        a.powerLaw = s->opt.powerLawExponent;
        if (s->opt.viscosityOn) {
          a.tau = s->stressTensor.comp(0);
          a.q = s->heatFlux.comp(0);
          if (mode == MG_FORWARD) {
            auto fv = p->arrays.find("viscousFluxes"), ft = p->arrays.find("targetViscousFluxes");
            if (fv == p->arrays.end() || ft == p->arrays.end())
              MG_FAIL("far-field patch: viscous fluxes missing (call mg_region_update_patches after setting the target state)");
            a.Fv = fv->second.p;
            a.FvTarget = ft->second.p;
          } else if (mode == MG_LINEARIZED) {
            auto fv = p->arrays.find("viscousFluxes");
            if (fv == p->arrays.end()) MG_FAIL("far-field patch: linearized viscous fluxes have not been collected");
            a.Fv = fv->second.p;
          }
        }
        MG_TRY(dispatch_nd(s->nD, [&](auto nd) {
          { k_farfield<decltype(nd)::value><<<nblocks(p->nPatchPoints), 128, 0, st>>>(a); mg_count_launches(1); }
          return 0;
        }));
        break;
      }
      case MG_PATCH_SPONGE: {
        auto it = p->arrays.find("spongeStrength");
        if (it == p->arrays.end()) MG_FAIL("sponge patch: spongeStrength has not been set");
        SpongeArgs a;
        a.g = geom(p);
        const MgField& X = mode == MG_FORWARD ? Q : W;
        a.X = X.comp(0); a.csX = X.compStride;
        a.target = s->target.comp(0);
        a.strength = it->second.p;
        a.iblank = g->iblank;
        a.rhs = s->rhs.comp(0);
        a.cs = s->rhs.compStride;
        a.nU = s->nU;
        a.mode = mode;
        { k_sponge<<<nblocks(p->nPatchPoints), 128, 0, st>>>(a); mg_count_launches(1); }
        break;
      }
      case MG_PATCH_KOLMOGOROV_FORCING: {
        auto it = p->arrays.find("forcePerUnitMass");
        if (it == p->arrays.end()) MG_FAIL("Kolmogorov forcing patch: forcePerUnitMass has not been set (mg_patch_kolmogorov_setup)");
        KolmogorovArgs a;
        a.g = geom(p);
        a.force = it->second.p;
        a.Q = Q.comp(0); a.csQ = Q.compStride;
        a.W = W.p ? W.comp(0) : nullptr; a.csW = W.compStride;
        a.iblank = g->iblank;
        a.rhs = s->rhs.comp(0);
        a.cs = s->rhs.compStride;
        a.mode = mode;
        { k_kolmogorov<<<nblocks(p->nPatchPoints), 128, 0, st>>>(a); mg_count_launches(1); }
        break;
      }
      case MG_PATCH_JET_EXCITATION: {
        if (mode != MG_FORWARD) break;          // src/JetExcitationPatchImpl.f90:160
        const int nModes = (int)p->angularFrequencies.size();
        if (nModes == 0) break;
        auto is = p->arrays.find("spongeStrength"), ir = p->arrays.find("perturbationReal"),
             ii = p->arrays.find("perturbationImag");
        if (is == p->arrays.end()) MG_FAIL("jet excitation patch: spongeStrength has not been set");
        if (ir == p->arrays.end() || ii == p->arrays.end() || ir->second.nComp != s->nU * nModes ||
            ii->second.nComp != s->nU * nModes)
          MG_FAIL("jet excitation patch: perturbationReal / perturbationImag (nPatchPoints, nUnknowns * nModes) have not been set");
        JetArgs a;
        a.g = geom(p);
        a.strength = is->second.p;
        a.re = ir->second.p;
        a.im = ii->second.p;
        a.iblank = g->iblank;
        a.rhs = s->rhs.comp(0);
        a.cs = s->rhs.compStride;
        a.nU = s->nU;
        a.nModes = nModes;
        for (int l = 0; l < nModes; ++l) {
          a.c[l] = std::cos(p->angularFrequencies[l] * s->time);
          a.s[l] = std::sin(p->angularFrequencies[l] * s->time);
        }
        { k_jet_excitation<<<nblocks(p->nPatchPoints), 128, 0, st>>>(a); mg_count_launches(1); }
        break;
      }
      case MG_PATCH_PROBE:
        break;                                  // updateProbePatch is empty (src/ProbePatchImpl.f90:71-97)
      case MG_PATCH_ADIABATIC_WALL:             // the reference's adiabatic viscous penalties are identically zero
      case MG_PATCH_SLIP_WALL:                  // (src/AdiabaticWallImpl.f90:128): what is left is the slip wall
      case MG_PATCH_ISOTHERMAL_WALL: {
        if (mode == MG_ADJOINT && s->opt.useContinuousAdjoint) break;
        WallArgs a;
        std::memset(&a, 0, sizeof(a));
        a.g = geom(p);
        a.Q = Q.comp(0); a.csQ = Q.compStride;
        a.W = W.comp(0); a.csW = W.compStride;
        a.m = g->metrics.comp(0);
        a.jac = g->jacobian.comp(0);
        a.v = s->specificVolume.comp(0);
        a.u = s->velocity.comp(0);
        a.pr = s->pressure.comp(0);
        a.T = s->temperature.comp(0);
        a.iblank = g->iblank;
        a.rhs = s->rhs.comp(0);
        a.cs = s->rhs.compStride;
        a.dir = dir;
        a.mode = mode;
        a.isothermal = p->type == MG_PATCH_ISOTHERMAL_WALL;
        a.viscous = s->opt.viscosityOn;
        a.sigmaI = p->inviscidPenaltyAmount;
        a.sigmaV1 = p->viscousPenaltyAmount;
        a.gamma = s->opt.ratioOfSpecificHeats;
        if (a.isothermal && a.viscous) {
          auto it = p->arrays.find("temperature");
          if (it == p->arrays.end()) MG_FAIL("isothermal wall patch: temperature has not been set");
          a.wallT = it->second.p;
        }
        MG_TRY(dispatch_nd(s->nD, [&](auto nd) {
          { k_wall<decltype(nd)::value><<<nblocks(p->nPatchPoints), 128, 0, st>>>(a); mg_count_launches(1); }
          return 0;
        }));
        break;
      }
      case MG_PATCH_COST_TARGET: {
        if (mode == MG_FORWARD) break;
        auto it = p->arrays.find("adjointForcing");
        if (it == p->arrays.end()) break;
        AddArgs a;
        a.g = geom(p);
        a.data = it->second.p;
        a.mollifier = nullptr;
        a.iblank = g->iblank;
        a.rhs = s->rhs.comp(0);
        a.cs = s->rhs.compStride;
        a.nComp = s->nU;
        a.factor = (s->opt.useContinuousAdjoint || s->opt.steadyStateSimulation) ? 1.0 : s->adjointForcingFactor;
        { k_patch_add<<<nblocks(p->nPatchPoints), 128, 0, st>>>(a); mg_count_launches(1); }
        break;
      }
      case MG_PATCH_ACTUATOR: {
        if (mode == MG_ADJOINT) break;
        auto it = p->arrays.find(mode == MG_FORWARD ? "controlForcing" : "deltaControlForcing");
        if (it == p->arrays.end()) break;
        if (!g->controlMollifier.p) MG_FAIL("actuator patch: control mollifier has not been set");
        AddArgs a;
        a.g = geom(p);
        a.data = it->second.p;
        a.mollifier = g->controlMollifier.comp(0);
        a.iblank = g->iblank;
        a.rhs = s->rhs.comp(0);
        a.cs = s->rhs.compStride;
        a.nComp = s->nU;
        a.factor = 1.0;
        { k_patch_add<<<nblocks(p->nPatchPoints), 128, 0, st>>>(a); mg_count_launches(1); }
        break;
      }
      case MG_PATCH_BLOCK_INTERFACE:
        MG_TRY(mg_interface_apply(s, p, mode));
        break;
      default:
        MG_FAIL("patch: unknown type");
    }
    MG_CUDA(cudaGetLastError());
  }
  return 0;
}

// setupKolmogorovForcingPatch (reference src/KolmogorovForcingPatchImpl.f90:3-68): the keys
// patches/<name>/amplitude and patches/<name>/wavenumber
int mg_patch_kolmogorov_setup_impl(mg_patch* p, double amplitude, int wavenumber) {
  if (p->type != MG_PATCH_KOLMOGOROV_FORCING) MG_FAIL("mg_patch_kolmogorov_setup: not a KOLMOGOROV_FORCING patch");
  mg_grid* g = p->state->grid;
  double* d = nullptr;
  MG_TRY(mg_patch_alloc_array(p, "forcePerUnitMass", 1, &d));
  if (p->nPatchPoints == 0) return 0;
  if (!g->coordinates.p) MG_FAIL("mg_patch_kolmogorov_setup: grid coordinates have not been set");
  const double pi = 4.0 * std::atan(1.0);
  const int n = std::max(0, wavenumber);
  { k_kolmogorov_setup<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(geom(p), g->coordinates.comp(1), g->iblank, amplitude,
                                                                          2.0 * pi * n, d); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return 0;
}

// t_ProbePatch (reference src/ProbePatchImpl.f90:3-60, saveProbeData src/RegionImpl.f90:2211-2281): the probe buffer
// lives on the device; a record is one collect kernel, the flush one device -> host copy of the filled part.
int mg_patch_probe_setup_impl(mg_patch* p, int bufferSize) {
  if (p->type != MG_PATCH_PROBE) MG_FAIL("mg_patch_probe_setup: not a PROBE patch");
  if (bufferSize < 1) MG_FAIL("mg_patch_probe_setup: probe_buffer_size must be positive");
  cudaFree(p->probeBuffer);
  p->probeBuffer = nullptr;
  p->probeCapacity = bufferSize;
  p->probeCount = 0;
  if (p->nPatchPoints > 0)
    MG_CUDA(cudaMalloc(&p->probeBuffer, sizeof(double) * (size_t)p->nPatchPoints * p->state->nU * bufferSize));
  return 0;
}

int mg_patch_probe_record_impl(mg_patch* p, int mode, int* full) {
  if (p->type != MG_PATCH_PROBE || p->probeCapacity < 1) MG_FAIL("mg_patch_probe_record: the probe has not been set up");
  if (mode != MG_FORWARD && mode != MG_ADJOINT) MG_FAIL("mg_patch_probe_record: mode must be FORWARD or ADJOINT");
  if (p->probeCount >= p->probeCapacity) MG_FAIL("mg_patch_probe_record: the probe buffer is full (flush it)");
  mg_state* s = p->state;
  const MgField& X = mode == MG_FORWARD ? s->Q[s->cur] : s->W[s->curW];
  if (p->nPatchPoints > 0) {
    if (!X.p) MG_FAIL("mg_patch_probe_record: the field has not been set");
    double* dst = p->probeBuffer + (size_t)p->probeCount * p->nPatchPoints * s->nU;
    { k_collect<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(geom(p), X.comp(0), X.compStride, s->nU, dst); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
  }
  ++p->probeCount;
  if (full) *full = p->probeCount == p->probeCapacity;
  return 0;
}

int mg_patch_probe_flush_impl(mg_patch* p, double* host, int* count) {
  if (p->type != MG_PATCH_PROBE) MG_FAIL("mg_patch_probe_flush: not a PROBE patch");
  if (count) *count = p->probeCount;
  if (p->probeCount > 0 && p->nPatchPoints > 0) {
    if (!host) MG_FAIL("mg_patch_probe_flush: null host buffer");
    MG_CUDA(cudaMemcpyAsync(host, p->probeBuffer, sizeof(double) * (size_t)p->nPatchPoints * p->state->nU * p->probeCount,
                            cudaMemcpyDefault, mg_stream()));
    MG_CUDA(cudaStreamSynchronize(mg_stream()));
  }
  p->probeCount = 0;
  return 0;
}

// ------------------------------------------------------------------- functionals
// computeQuadratureOnPatches (reference src/PatchFactoryImpl.f90:376-444), computeAcousticNoise and its adjoint
// forcing (src/AcousticNoiseImpl.f90:123-280), computeThermalActuatorSensitivity / gradient sample
// (src/ThermalActuatorImpl.f90:83-159, 383-443).  The reduction is deterministic: fixed grid, per-block partial
// sums, summed on the host in block order.
namespace {

constexpr int QUAD_BLOCKS = 592, QUAD_THREADS = 256, QUAD_MAX_PATCHES = 16;

struct QuadArgs {
  int lo[QUAD_MAX_PATCHES][3], hi[QUAD_MAX_PATCHES][3];   // local boxes [lo, hi) of the patches of the wanted type
  int nPatches;
  int nx, ny;
  size_t N;
  const int* iblank;
  const double *norm, *a, *b, *w;
  int kind;          // 0: a        1: (a - b)^2 w   (acoustic noise)        2: (a w)^2   (thermal actuator)
  double* partial;
};

__global__ void __launch_bounds__(QUAD_THREADS) k_quadrature(QuadArgs q) {
  __shared__ double red[QUAD_THREADS];
  double acc = 0.0;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < q.N; p += (size_t)gridDim.x * blockDim.x) {
    if (q.iblank && q.iblank[p] == 0) continue;
    const int i = (int)(p % q.nx), j = (int)((p / q.nx) % q.ny), k = (int)(p / ((size_t)q.nx * q.ny));
    bool in = false;
    for (int l = 0; l < q.nPatches && !in; ++l)
      in = i >= q.lo[l][0] && i < q.hi[l][0] && j >= q.lo[l][1] && j < q.hi[l][1] && k >= q.lo[l][2] && k < q.hi[l][2];
    if (!in) continue;
    double v;
    if (q.kind == 0) v = q.a[p];
    else if (q.kind == 1) { const double d = q.a[p] - q.b[p]; v = d * d * q.w[p]; }
    else { const double d = q.a[p] * q.w[p]; v = d * d; }
    acc += q.norm[p] * v;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int st = QUAD_THREADS / 2; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) q.partial[blockIdx.x] = red[0];
}

// sum of the block partials in block order (the order of the host-side sum), then acc += weight * (sum * scale): no
// fused multiply-add, so that the device-accumulated functional equals the host-accumulated one bit for bit
__global__ void __launch_bounds__(QUAD_BLOCKS > 1024 ? 1024 : QUAD_BLOCKS)
k_quad_accumulate(const double* partial, int n, double scale, double weight, double* acc) {
  __shared__ double sh[QUAD_BLOCKS];
  for (int i = threadIdx.x; i < n; i += blockDim.x) sh[i] = partial[i];      // one parallel read of the partials ...
  __syncthreads();
  if (threadIdx.x != 0) return;
  double sum = 0.0;
  for (int i = 0; i < n; ++i) sum = __dadd_rn(sum, sh[i]);                   // ... summed in block order
  *acc = __dadd_rn(*acc, __dmul_rn(weight, __dmul_rn(sum, scale)));
}

int quadrature(mg_state* s, int patchType, int kind, const double* a, const double* b, const double* w, double* value,
               double* devAcc = nullptr, double scale = 1.0, double weight = 1.0) {
  mg_grid* g = s->grid;
  QuadArgs q;
  std::memset(&q, 0, sizeof(q));
  for (mg_patch* p : s->patches) {
    if (p->type != patchType || p->nPatchPoints <= 0) continue;
    if (q.nPatches >= QUAD_MAX_PATCHES) MG_FAIL("quadrature on patches: too many patches of one type");
    for (int d = 0; d < 3; ++d) { q.lo[q.nPatches][d] = p->localLo[d]; q.hi[q.nPatches][d] = p->localLo[d] + p->localSize[d]; }
    ++q.nPatches;
  }
  if (value) *value = 0.0;
  if (q.nPatches == 0) return 0;
  q.nx = g->localSize[0];
  q.ny = g->localSize[1];
  q.N = g->N;
  q.iblank = g->iblank;
  q.norm = g->norm.comp(0);
  q.a = a; q.b = b; q.w = w; q.kind = kind;
  static double* partial = nullptr;
  if (!partial) MG_CUDA(cudaMalloc(&partial, QUAD_BLOCKS * sizeof(double)));
  q.partial = partial;
  { k_quadrature<<<QUAD_BLOCKS, QUAD_THREADS, 0, mg_stream()>>>(q); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  if (devAcc) {      // device-resident accumulation: no host synchronisation
    { k_quad_accumulate<<<1, (QUAD_BLOCKS > 1024 ? 1024 : QUAD_BLOCKS), 0, mg_stream()>>>(partial, QUAD_BLOCKS, scale, weight, devAcc); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
    return 0;
  }
  double host[QUAD_BLOCKS];
  MG_CUDA(cudaMemcpyAsync(host, partial, sizeof(host), cudaMemcpyDeviceToHost, mg_stream()));
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  double sum = 0.0;
  for (int i = 0; i < QUAD_BLOCKS; ++i) sum += host[i];
  *value = sum;
  return mg_p2p_check_all();      // a functional of a state with stale ghost planes must not come back silently
}

struct ForcingArgs {
  PatchGeom g;
  const int* iblank;
  const double *pressure, *meanPressure, *mollifier, *u;
  size_t csU;
  int nD;
  double gamma, ramp;
  double* out;       // adjointForcing (nPatchPoints, nU), point fastest
};

__global__ void k_noise_forcing(ForcingArgs a) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  const size_t p = a.g.gridIndex(q);
  if (a.iblank && a.iblank[p] == 0) return;
  const double F = -2.0 * a.mollifier[p] * a.ramp * (a.gamma - 1.0) * (a.pressure[p] - a.meanPressure[p]);
  double usq = 0.0;
  for (int d = 0; d < a.nD; ++d) {
    const double ud = a.u[(size_t)d * a.csU + p];
    a.out[(size_t)(d + 1) * a.g.n + q] = -ud * F;
    usq += ud * ud;
  }
  a.out[(size_t)(a.nD + 1) * a.g.n + q] = F;
  a.out[q] = 0.5 * usq * F;
}

__global__ void k_actuator_gradient(PatchGeom g, const double* wE, const double* mollifier, double ramp, double* out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= g.n) return;
  const size_t p = g.gridIndex(q);
  out[q] = wE[p] * (mollifier[p] * ramp);
}

// t_PressureDrag (reference src/PressureDragImpl.f90:61-267) on the COST_TARGET patches: the integrand uses the
// patch norm of updatePatchFactories (src/PatchFactoryImpl.f90:496-505) = SBP norm weights of the tangential
// directions (no Jacobian), evaluated here from the norm tables instead of being stored per patch.
constexpr int DRAG_BLOCKS = 64, DRAG_THREADS = 256;

struct DragArgs {
  PatchGeom g;
  const int* iblank;
  const double *pressure, *metricsK, *jac, *u, *Q, *W;
  const double *tau = nullptr, *mollifier = nullptr;   // kind 1 (DragForce): stress tensor and target mollifier
  size_t cs, csQ, csW;
  int kind = 0;                                         // 0: PRESSURE_DRAG, 1: DRAG (viscous)
  int nD, axis, normalDirection, continuous;
  int n[3], depth[3], hasB0[3], hasB1[3];
  double norm[3][MG_MAX_BDEPTH];
  double dirv[3], factor, gamma, sigmaI;
  double* partial;     // k_drag
  double* out;         // k_drag_forcing: adjointForcing (nPatchPoints, nU), point fastest
};

__device__ __forceinline__ double drag_patch_norm(const DragArgs& a, int q) {
  const int c[3] = {a.g.lo[0] + q % a.g.sz[0], a.g.lo[1] + (q / a.g.sz[0]) % a.g.sz[1],
                    a.g.lo[2] + q / (a.g.sz[0] * a.g.sz[1])};
  double w = 1.0;
  for (int d = 0; d < a.nD; ++d) {
    if (d == a.axis) continue;
    if (a.hasB0[d] && c[d] < a.depth[d]) w *= a.norm[d][c[d]];
    if (a.hasB1[d] && c[d] >= a.n[d] - a.depth[d]) w *= a.norm[d][a.n[d] - 1 - c[d]];
  }
  return w;
}

__global__ void __launch_bounds__(DRAG_THREADS) k_drag(DragArgs a) {
  __shared__ double red[DRAG_THREADS];
  double acc = 0.0;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < a.g.n; q += gridDim.x * blockDim.x) {
    const size_t p = a.g.gridIndex(q);
    if (a.iblank && a.iblank[p] == 0) continue;
    if (a.kind == 1) {
      // computeDragForce (reference src/DragForceImpl.f90:108-127): F = nbf sum_l dir_l (metrics_k . tau_l), weight =
      // target mollifier
      double F = 0.0;
      for (int l = 0; l < a.nD; ++l) {
        double mt = 0.0;
        for (int j = 0; j < a.nD; ++j) mt += a.metricsK[(size_t)j * a.cs + p] * a.tau[(size_t)(l * a.nD + j) * a.cs + p];
        F += a.dirv[l] * mt;
      }
      acc += (a.factor * F) * drag_patch_norm(a, q) * a.mollifier[p];
      continue;
    }
    double md = 0.0;
    for (int l = 0; l < a.nD; ++l) md = (l == 0) ? a.metricsK[p] * a.dirv[0] : md + a.metricsK[(size_t)l * a.cs + p] * a.dirv[l];
    acc += (0.0 - (a.pressure[p] - 1.0 / a.gamma)) * drag_patch_norm(a, q) * (md * a.factor);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int st = DRAG_THREADS / 2; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.partial[blockIdx.x] = red[0];
}

template <int ND>
__global__ void k_drag_forcing(DragArgs a) {
  constexpr int NU = ND + 2;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  const size_t p = a.g.gridIndex(q);
  if (a.iblank && a.iblank[p] == 0) return;
  double m[ND];
#pragma unroll
  for (int l = 0; l < ND; ++l) m[l] = a.metricsK[(size_t)l * a.cs + p];
  const double sgn = a.normalDirection > 0 ? 1.0 : -1.0;
  if (a.continuous) {
    double arc = 0.0;
#pragma unroll
    for (int l = 0; l < ND; ++l) arc = (l == 0) ? m[0] * m[0] : arc + m[l] * m[l];
    arc = sqrt(arc);
    double n[ND], Qp[NU], A[NU][NU];
    double F = 0.0;
#pragma unroll
    for (int l = 0; l < ND; ++l) {
      n[l] = m[l] / arc;
      const double t = (a.W[(size_t)(l + 1) * a.csW + p] - sgn * fabs(a.dirv[l])) * n[l];
      F = (l == 0) ? t : F + t;
    }
    F = a.jac[p] * F;
#pragma unroll
    for (int c = 0; c < NU; ++c) Qp[c] = a.Q[(size_t)c * a.csQ + p];
    incoming_jacobian<ND>(Qp, m, a.gamma, -a.normalDirection, A);
#pragma unroll
    for (int c = 0; c < NU; ++c) {
      double t = 0.0;
#pragma unroll
      for (int l = 0; l < ND; ++l) t = (l == 0) ? A[1][c] * n[0] : t + A[l + 1][c] * n[l];
      a.out[(size_t)c * a.g.n + q] = (0.0 - a.sigmaI) * F * t;
    }
    return;
  }
  double md = 0.0, usq = 0.0;
#pragma unroll
  for (int l = 0; l < ND; ++l) md = (l == 0) ? m[0] * a.dirv[0] : md + m[l] * a.dirv[l];
  const double F = a.jac[p] * (sgn * a.factor) * (a.gamma - 1.0) * md;
#pragma unroll
  for (int l = 0; l < ND; ++l) {
    const double ul = a.u[(size_t)l * a.cs + p];
    a.out[(size_t)(l + 1) * a.g.n + q] = (0.0 - ul) * F;
    usq = (l == 0) ? ul * ul : usq + ul * ul;
  }
  a.out[q] = 0.5 * usq * F;
  a.out[(size_t)(ND + 1) * a.g.n + q] = F;
}

int drag_args(mg_state* s, mg_patch* p, const double direction[3], DragArgs* a) {
  mg_grid* g = s->grid;
  std::memset(a, 0, sizeof(*a));
  const int nD = s->nD;
  const int k = std::abs(p->normalDirection);
  if (k < 1 || k > nD) MG_FAIL("pressure drag: the COST_TARGET patch needs a normal direction");
  double nrm = 0.0;
  for (int l = 0; l < nD; ++l) nrm += direction[l] * direction[l];
  if (nrm <= 2.220446049250313e-16) MG_FAIL("Unable to determine a unit vector for computing pressure drag!");
  for (int l = 0; l < nD; ++l) a->dirv[l] = direction[l] / sqrt(nrm);
  a->g = geom(p);
  a->iblank = g->iblank;
  a->pressure = s->pressure.comp(0);
  a->metricsK = g->metrics.comp(nD * (k - 1));
  a->jac = g->jacobian.comp(0);
  a->u = s->velocity.comp(0);
  a->Q = s->Q[s->cur].comp(0);
  a->W = s->W[s->curW].comp(0);
  a->cs = g->metrics.compStride;
  a->csQ = s->Q[s->cur].compStride;
  a->csW = s->W[s->curW].compStride;
  a->nD = nD;
  a->axis = k - 1;
  a->normalDirection = p->normalDirection;
  a->continuous = s->opt.useContinuousAdjoint;
  for (int d = 0; d < nD; ++d) {
    const MgDevOp& op = g->firstDerivative[d]->op;
    a->n[d] = g->localSize[d];
    a->depth[d] = op.normDepth;
    a->hasB0[d] = op.hasDomainBoundary[0];
    a->hasB1[d] = op.hasDomainBoundary[1];
    for (int m = 0; m < MG_MAX_BDEPTH; ++m) a->norm[d][m] = op.normBoundary[m];
  }
  a->factor = 1.0 / g->firstDerivative[k - 1]->op.normBoundary[0];
  a->gamma = s->opt.ratioOfSpecificHeats;
  a->sigmaI = p->inviscidPenaltyAmount;
  return 0;
}

}  // namespace

int mg_functional_quadrature_impl(mg_state* s, int patchType, const double* integrandDevice, double* value) {
  return quadrature(s, patchType, 0, integrandDevice, nullptr, nullptr, value);
}

int mg_functional_acoustic_noise_impl(mg_state* s, double timeRampFactor, double* value) {
  mg_grid* g = s->grid;
  if (!s->meanPressure.p) MG_FAIL("acoustic noise: the mean pressure has not been set");
  if (!g->targetMollifier.p) MG_FAIL("acoustic noise: the target mollifier has not been set");
  MG_TRY(mg_state_ensure_dependents(s));
  MG_TRY(quadrature(s, MG_PATCH_COST_TARGET, 1, s->pressure.comp(0), s->meanPressure.comp(0),
                    g->targetMollifier.comp(0), value));
  *value *= timeRampFactor;
  return 0;
}

int mg_functional_acoustic_noise_forcing_impl(mg_state* s, double timeRampFactor) {
  mg_grid* g = s->grid;
  if (!s->meanPressure.p) MG_FAIL("acoustic noise: the mean pressure has not been set");
  if (!g->targetMollifier.p) MG_FAIL("acoustic noise: the target mollifier has not been set");
  MG_TRY(mg_state_ensure_dependents(s));
  for (mg_patch* p : s->patches) {
    if (p->type != MG_PATCH_COST_TARGET || p->nPatchPoints <= 0) continue;
    double* out = nullptr;
    auto it = p->arrays.find("adjointForcing");
    if (it == p->arrays.end()) MG_TRY(mg_patch_alloc_array(p, "adjointForcing", s->nU, &out));
    else out = it->second.p;
    ForcingArgs a;
    a.g = geom(p);
    a.iblank = g->iblank;
    a.pressure = s->pressure.comp(0);
    a.meanPressure = s->meanPressure.comp(0);
    a.mollifier = g->targetMollifier.comp(0);
    a.u = s->velocity.comp(0);
    a.csU = s->velocity.compStride;
    a.nD = s->nD;
    a.gamma = s->opt.ratioOfSpecificHeats;
    a.ramp = timeRampFactor;
    a.out = out;
    { k_noise_forcing<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(a); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
  }
  return 0;
}

int mg_functional_actuator_sensitivity_impl(mg_state* s, double timeRampFactor, double* value) {
  mg_grid* g = s->grid;
  if (!g->controlMollifier.p) MG_FAIL("thermal actuator: the control mollifier has not been set");
  MG_TRY(quadrature(s, MG_PATCH_ACTUATOR, 2, s->W[s->curW].comp(s->nD + 1), nullptr, g->controlMollifier.comp(0), value));
  *value *= timeRampFactor * timeRampFactor;
  return 0;
}

int mg_functional_actuator_gradient_impl(mg_patch* p, double timeRampFactor, double* hostOut) {
  mg_state* s = p->state;
  mg_grid* g = s->grid;
  if (p->type != MG_PATCH_ACTUATOR) MG_FAIL("thermal actuator gradient: not an ACTUATOR patch");
  if (!g->controlMollifier.p) MG_FAIL("thermal actuator: the control mollifier has not been set");
  if (p->nPatchPoints <= 0) return 0;
  double* out = nullptr;
  auto it = p->arrays.find("gradient");
  if (it == p->arrays.end()) MG_TRY(mg_patch_alloc_array(p, "gradient", 1, &out));
  else out = it->second.p;
  { k_actuator_gradient<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(geom(p), s->W[s->curW].comp(s->nD + 1),
                                                                      g->controlMollifier.comp(0), timeRampFactor, out); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  MG_CUDA(cudaMemcpyAsync(hostOut, out, (size_t)p->nPatchPoints * sizeof(double), cudaMemcpyDeviceToHost, mg_stream()));
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  return 0;
}

// Time quadrature of the cost functional / sensitivity on the device (J += norm(i) dt I, reference
// src/SolverImpl.f90:837-841, :1181-1185): which = 0 acoustic noise, 1 thermal-actuator sensitivity.
int mg_functional_accumulate_impl(mg_state* s, int which, double weight, double timeRampFactor) {
  mg_grid* g = s->grid;
  if (which < 0 || which >= MG_STATE_ACCUMULATORS) MG_FAIL("mg_functional_accumulate: unknown accumulator");
  if (!s->accumulators) {
    MG_CUDA(cudaMalloc(&s->accumulators, MG_STATE_ACCUMULATORS * sizeof(double)));
    MG_CUDA(cudaMemsetAsync(s->accumulators, 0, MG_STATE_ACCUMULATORS * sizeof(double), mg_stream()));
  }
  if (which == 0) {
    if (!s->meanPressure.p) MG_FAIL("acoustic noise: the mean pressure has not been set");
    if (!g->targetMollifier.p) MG_FAIL("acoustic noise: the target mollifier has not been set");
    MG_TRY(mg_state_ensure_dependents(s));
    return quadrature(s, MG_PATCH_COST_TARGET, 1, s->pressure.comp(0), s->meanPressure.comp(0),
                      g->targetMollifier.comp(0), nullptr, s->accumulators + 0, timeRampFactor, weight);
  }
  if (!g->controlMollifier.p) MG_FAIL("thermal actuator: the control mollifier has not been set");
  return quadrature(s, MG_PATCH_ACTUATOR, 2, s->W[s->curW].comp(s->nD + 1), nullptr, g->controlMollifier.comp(0),
                    nullptr, s->accumulators + 1, timeRampFactor * timeRampFactor, weight);
}

int mg_functional_accumulator_get_impl(mg_state* s, int which, double* value, int reset) {
  if (which < 0 || which >= MG_STATE_ACCUMULATORS) MG_FAIL("mg_functional_accumulator: unknown accumulator");
  *value = 0.0;
  if (!s->accumulators) return 0;
  MG_CUDA(cudaMemcpyAsync(value, s->accumulators + which, sizeof(double), cudaMemcpyDeviceToHost, mg_stream()));
  if (reset) MG_CUDA(cudaMemsetAsync(s->accumulators + which, 0, sizeof(double), mg_stream()));
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  return mg_p2p_check_all();
}

// gradientBuffer of t_ActuatorPatch (reference src/ActuatorPatchImpl.f90:226-458) on the device: samples are recorded
// without a host synchronisation and read back in blocks.
int mg_patch_gradient_buffer_setup_impl(mg_patch* p, int nSlots) {
  if (p->type != MG_PATCH_ACTUATOR) MG_FAIL("mg_patch_gradient_buffer_setup: not an ACTUATOR patch");
  if (nSlots < 1) MG_FAIL("mg_patch_gradient_buffer_setup: controller_buffer_size must be positive");
  cudaFree(p->gradientBuffer);
  p->gradientBuffer = nullptr;
  p->gradientCapacity = nSlots;
  p->gradientCount = 0;
  if (p->nPatchPoints > 0) MG_CUDA(cudaMalloc(&p->gradientBuffer, sizeof(double) * (size_t)p->nPatchPoints * nSlots));
  return 0;
}

int mg_functional_actuator_gradient_record_impl(mg_patch* p, double timeRampFactor, int* full) {
  mg_state* s = p->state;
  mg_grid* g = s->grid;
  if (p->type != MG_PATCH_ACTUATOR || p->gradientCapacity < 1) MG_FAIL("actuator gradient record: the gradient buffer has not been set up");
  if (!g->controlMollifier.p) MG_FAIL("thermal actuator: the control mollifier has not been set");
  if (p->gradientCount >= p->gradientCapacity) MG_FAIL("actuator gradient record: the gradient buffer is full (flush it)");
  if (p->nPatchPoints > 0) {
    { k_actuator_gradient<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(
          geom(p), s->W[s->curW].comp(s->nD + 1), g->controlMollifier.comp(0), timeRampFactor,
          p->gradientBuffer + (size_t)p->gradientCount * p->nPatchPoints); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
  }
  ++p->gradientCount;
  if (full) *full = p->gradientCount == p->gradientCapacity;
  return 0;
}

int mg_patch_gradient_buffer_flush_impl(mg_patch* p, double* host, int* count) {
  if (p->type != MG_PATCH_ACTUATOR) MG_FAIL("mg_patch_gradient_buffer_flush: not an ACTUATOR patch");
  if (count) *count = p->gradientCount;
  if (p->gradientCount > 0 && p->nPatchPoints > 0) {
    if (!host) MG_FAIL("mg_patch_gradient_buffer_flush: null host buffer");
    MG_CUDA(cudaMemcpyAsync(host, p->gradientBuffer, sizeof(double) * (size_t)p->nPatchPoints * p->gradientCount,
                            cudaMemcpyDefault, mg_stream()));
    MG_CUDA(cudaStreamSynchronize(mg_stream()));
  }
  p->gradientCount = 0;
  return 0;
}

namespace {
// controlForcing(:, :) = 0; controlForcing(:, first : first + nComp - 1) = controlForcingBuffer(:, :, slot)
// (updateForcing of the controllers, reference src/ThermalActuatorImpl.f90:161-233, src/MomentumActuatorImpl.f90:165-224)
__global__ void k_forcing_from_buffer(int n, int nU, int first, int nComp, const double* buffer, double* forcing) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  for (int c = 0; c < nU; ++c) {
    const int b = c - first;
    forcing[(size_t)c * n + q] = (b >= 0 && b < nComp) ? buffer[(size_t)b * n + q] : 0.0;
  }
}
}  // namespace

int mg_patch_control_forcing_from_buffer_impl(mg_patch* p, int slot, int firstComponent, int nComponents) {
  mg_state* s = p->state;
  if (p->type != MG_PATCH_ACTUATOR) MG_FAIL("mg_patch_control_forcing_from_buffer: not an ACTUATOR patch");
  auto it = p->arrays.find("controlForcingBuffer");
  if (it == p->arrays.end()) MG_FAIL("mg_patch_control_forcing_from_buffer: controlForcingBuffer has not been set");
  if (nComponents < 1 || firstComponent < 0 || firstComponent + nComponents > s->nU)
    MG_FAIL("mg_patch_control_forcing_from_buffer: component range out of bounds");
  if (slot < 0 || (slot + 1) * nComponents > it->second.nComp) MG_FAIL("mg_patch_control_forcing_from_buffer: slot out of range");
  double* forcing = nullptr;
  MG_TRY(mg_patch_alloc_array(p, "controlForcing", s->nU, &forcing));
  it = p->arrays.find("controlForcingBuffer");      // the map may have rehashed nothing, but stay safe
  if (p->nPatchPoints > 0) {
    { k_forcing_from_buffer<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(
          p->nPatchPoints, s->nU, firstComponent, nComponents,
          it->second.p + (size_t)slot * nComponents * p->nPatchPoints, forcing); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
  }
  return 0;
}

// computePressureDrag (reference src/PressureDragImpl.f90:61-132); local to this rank, the caller reduces
int mg_functional_pressure_drag_impl(mg_state* s, const double direction[3], double* value) {
  MG_TRY(mg_state_ensure_dependents(s));
  static double* partial = nullptr;
  if (!partial) MG_CUDA(cudaMalloc(&partial, DRAG_BLOCKS * sizeof(double)));
  double sum = 0.0;
  for (mg_patch* p : s->patches) {
    if (p->type != MG_PATCH_COST_TARGET || p->nPatchPoints <= 0) continue;
    DragArgs a;
    MG_TRY(drag_args(s, p, direction, &a));
    a.partial = partial;
    { k_drag<<<DRAG_BLOCKS, DRAG_THREADS, 0, mg_stream()>>>(a); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
    double host[DRAG_BLOCKS];
    MG_CUDA(cudaMemcpyAsync(host, partial, sizeof(host), cudaMemcpyDeviceToHost, mg_stream()));
    MG_CUDA(cudaStreamSynchronize(mg_stream()));
    for (int i = 0; i < DRAG_BLOCKS; ++i) sum += host[i];
  }
  *value = sum;
  return 0;
}

// computeDragForce (reference src/DragForceImpl.f90:61-146): the viscous drag on the COST_TARGET patches; 0 when the
// flow is inviscid.  Local to this rank.
int mg_functional_drag_force_impl(mg_state* s, const double direction[3], double* value) {
  mg_grid* g = s->grid;
  *value = 0.0;
  if (!s->opt.viscosityOn) return 0;
  if (!g->targetMollifier.p) MG_FAIL("drag force: the target mollifier has not been set");
  MG_TRY(mg_state_ensure_dependents(s));
  static double* partial = nullptr;
  if (!partial) MG_CUDA(cudaMalloc(&partial, DRAG_BLOCKS * sizeof(double)));
  double sum = 0.0;
  for (mg_patch* p : s->patches) {
    if (p->type != MG_PATCH_COST_TARGET || p->nPatchPoints <= 0) continue;
    DragArgs a;
    MG_TRY(drag_args(s, p, direction, &a));
    if (s->stressTensor.compStride != a.cs) MG_FAIL("drag force: unexpected field layout");
    a.kind = 1;
    a.tau = s->stressTensor.comp(0);
    a.mollifier = g->targetMollifier.comp(0);
    a.partial = partial;
    { k_drag<<<DRAG_BLOCKS, DRAG_THREADS, 0, mg_stream()>>>(a); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
    double host[DRAG_BLOCKS];
    MG_CUDA(cudaMemcpyAsync(host, partial, sizeof(host), cudaMemcpyDeviceToHost, mg_stream()));
    MG_CUDA(cudaStreamSynchronize(mg_stream()));
    for (int i = 0; i < DRAG_BLOCKS; ++i) sum += host[i];
  }
  *value = sum;
  return 0;
}

namespace {
struct ReArgs {
  PatchGeom g;
  const int* iblank;
  const double *Q, *meanU, *mollifier;
  size_t csQ, csM, N;
  int nD;
  double d1[3], d2[3];
  double* out;
};

// integrand of computeReynoldsStress (reference src/ReynoldsStressImpl.f90:146-156), times the target mollifier
__global__ void k_reynolds_integrand(ReArgs a) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= a.N) return;
  const double v = 1.0 / a.Q[p];
  double s1 = 0.0, s2 = 0.0;
  for (int l = 0; l < a.nD; ++l) {
    const double du = v * a.Q[(size_t)(l + 1) * a.csQ + p] - a.meanU[(size_t)l * a.csM + p];
    s1 += du * a.d1[l];
    s2 += du * a.d2[l];
  }
  a.out[p] = 0.5 * s1 * s2 * a.mollifier[p];
}

// computeReynoldsStressAdjointForcing (reference :211-284).  As in the reference the second pair of assignments
// OVERWRITES the first: what is left is the firstDirection x F(secondDirection) term; the energy entry is not touched.
__global__ void k_reynolds_forcing(ReArgs a, int nU) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  const size_t p = a.g.gridIndex(q);
  if (a.iblank && a.iblank[p] == 0) return;
  const double v = 1.0 / a.Q[p];
  double u[3] = {0.0, 0.0, 0.0}, s2 = 0.0, d1u = 0.0;
  for (int l = 0; l < a.nD; ++l) {
    u[l] = v * a.Q[(size_t)(l + 1) * a.csQ + p];
    s2 += a.d2[l] * (u[l] - a.meanU[(size_t)l * a.csM + p]);
    d1u += a.d1[l] * u[l];
  }
  const double F = -0.5 * a.mollifier[p] * v * s2;
  for (int l = 0; l < a.nD; ++l) a.out[(size_t)(l + 1) * a.g.n + q] = a.d1[l] * F;
  a.out[q] = (0.0 - d1u) * F;
  (void)nU;
}

int reynolds_args(mg_state* s, const double d1[3], const double d2[3], ReArgs* a) {
  mg_grid* g = s->grid;
  if (!s->meanVelocity.p) MG_FAIL("Reynolds stress: the mean velocity has not been set (MG_Q_MEAN_VELOCITY)");
  if (!g->targetMollifier.p) MG_FAIL("Reynolds stress: the target mollifier has not been set");
  std::memset(a, 0, sizeof(*a));
  a->iblank = g->iblank;
  a->Q = s->Q[s->cur].comp(0); a->csQ = s->Q[s->cur].compStride;
  a->meanU = s->meanVelocity.comp(0); a->csM = s->meanVelocity.compStride;
  a->mollifier = g->targetMollifier.comp(0);
  a->N = g->N;
  a->nD = s->nD;
  double n1 = 0.0, n2 = 0.0;
  for (int l = 0; l < s->nD; ++l) { n1 += d1[l] * d1[l]; n2 += d2[l] * d2[l]; }
  if (n1 <= DBL_EPSILON || n2 <= DBL_EPSILON) MG_FAIL("Unable to determine unit vectors for computing Reynolds stress!");
  for (int l = 0; l < s->nD; ++l) { a->d1[l] = d1[l] / std::sqrt(n1); a->d2[l] = d2[l] / std::sqrt(n2); }
  return 0;
}
}  // namespace

int mg_functional_reynolds_stress_impl(mg_state* s, const double d1[3], const double d2[3], double* value) {
  mg_grid* g = s->grid;
  ReArgs a;
  MG_TRY(reynolds_args(s, d1, d2, &a));
  a.out = g->scratchA.comp(0);
  { k_reynolds_integrand<<<(unsigned)((g->N + 255) / 256), 256, 0, mg_stream()>>>(a); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return quadrature(s, MG_PATCH_COST_TARGET, 0, a.out, nullptr, nullptr, value);
}

int mg_functional_reynolds_stress_forcing_impl(mg_state* s, const double d1[3], const double d2[3]) {
  for (mg_patch* p : s->patches) {
    if (p->type != MG_PATCH_COST_TARGET || p->nPatchPoints <= 0) continue;
    ReArgs a;
    MG_TRY(reynolds_args(s, d1, d2, &a));
    a.g = geom(p);
    double* out = nullptr;
    auto it = p->arrays.find("adjointForcing");
    if (it == p->arrays.end()) {
      MG_TRY(mg_patch_alloc_array(p, "adjointForcing", s->nU, &out));
      MG_CUDA(cudaMemsetAsync(out, 0, (size_t)p->nPatchPoints * s->nU * sizeof(double), mg_stream()));
    } else out = it->second.p;
    a.out = out;
    { k_reynolds_forcing<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(a, s->nU); mg_count_launches(1); }
    MG_CUDA(cudaGetLastError());
  }
  return 0;
}

// t_MomentumActuator (reference src/MomentumActuatorImpl.f90:81-163, 351-412): direction 0 = all momentum components,
// d > 0 = component d only.  Sensitivity = quadrature over the ACTUATOR patches of sum_j (w_{j+1} mollifier)^2; a
// gradient sample = mollifier x w_{k+1} at the patch points, (nPatchPoints, nComponents).
int mg_functional_momentum_actuator_sensitivity_impl(mg_state* s, int direction, double* value) {
  mg_grid* g = s->grid;
  if (!g->controlMollifier.p) MG_FAIL("momentum actuator: the control mollifier has not been set");
  if (direction < -1 || direction > s->nD) MG_FAIL("momentum actuator: invalid direction");
  *value = 0.0;
  // direction -1: t_GenericActuator (reference src/GenericActuatorImpl.f90:77-149), every unknown
  for (int j = (direction < 0 ? 0 : 1); j <= (direction < 0 ? s->nU - 1 : s->nD); ++j) {
    if (direction > 0 && j != direction) continue;
    double r = 0.0;
    MG_TRY(quadrature(s, MG_PATCH_ACTUATOR, 2, s->W[s->curW].comp(j), nullptr, g->controlMollifier.comp(0), &r));
    *value += r;
  }
  return 0;
}

int mg_functional_momentum_actuator_gradient_impl(mg_patch* p, int direction, double* hostOut) {
  mg_state* s = p->state;
  mg_grid* g = s->grid;
  if (p->type != MG_PATCH_ACTUATOR) MG_FAIL("momentum actuator gradient: not an ACTUATOR patch");
  if (!g->controlMollifier.p) MG_FAIL("momentum actuator: the control mollifier has not been set");
  if (direction < -1 || direction > s->nD) MG_FAIL("momentum actuator: invalid direction");
  if (p->nPatchPoints <= 0) return 0;
  const int nComp = direction < 0 ? s->nU : (direction == 0 ? s->nD : 1);
  double* out = nullptr;
  MG_TRY(mg_patch_alloc_array(p, "momentumGradient", nComp, &out));
  int c = 0;
  for (int k = (direction < 0 ? 0 : 1); k <= (direction < 0 ? s->nU - 1 : s->nD); ++k) {
    if (direction > 0 && k != direction) continue;
    { k_actuator_gradient<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(geom(p), s->W[s->curW].comp(k),
                                                                        g->controlMollifier.comp(0), 1.0,
                                                                        out + (size_t)c * p->nPatchPoints); mg_count_launches(1); }
    ++c;
  }
  MG_CUDA(cudaGetLastError());
  MG_CUDA(cudaMemcpyAsync(hostOut, out, (size_t)p->nPatchPoints * nComp * sizeof(double), cudaMemcpyDeviceToHost, mg_stream()));
  MG_CUDA(cudaStreamSynchronize(mg_stream()));
  return 0;
}

// computePressureDragAdjointForcing (reference :148-267) into the COST_TARGET patches' "adjointForcing"
int mg_functional_pressure_drag_forcing_impl(mg_state* s, const double direction[3]) {
  MG_TRY(mg_state_ensure_dependents(s));
  for (mg_patch* p : s->patches) {
    if (p->type != MG_PATCH_COST_TARGET || p->nPatchPoints <= 0) continue;
    DragArgs a;
    MG_TRY(drag_args(s, p, direction, &a));
    double* out = nullptr;
    auto it = p->arrays.find("adjointForcing");
    if (it == p->arrays.end()) {
      MG_TRY(mg_patch_alloc_array(p, "adjointForcing", s->nU, &out));
      MG_CUDA(cudaMemsetAsync(out, 0, (size_t)p->nPatchPoints * s->nU * sizeof(double), mg_stream()));
    } else out = it->second.p;
    a.out = out;
    MG_TRY(dispatch_nd(s->nD, [&](auto nd) {
      { k_drag_forcing<decltype(nd)::value><<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(a); mg_count_launches(1); }
      return 0;
    }));
    MG_CUDA(cudaGetLastError());
  }
  return 0;
}
