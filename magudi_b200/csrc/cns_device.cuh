// Pointwise gas dynamics shared by the general and the fused RHS kernels: the device-function
// counterpart of CNSHelper (reference: src/CNSHelperImpl.f90).  Everything is templated on the
// number of dimensions ND (NU = ND + 2 unknowns) and operates on registers.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

struct PhysParams {
  double gamma;
  double ReInv, PrInv, powerLaw, bulkRatio;
  int viscous;
};

template <int ND>
struct Prim {          // dependent variables at one point
  double v;            // specific volume
  double u[ND];
  double p, T;
};

// computeDependentVariables (reference :3-87)
template <int ND>
__device__ __forceinline__ void dependent(const double* Q, double gamma, Prim<ND>& s) {
  s.v = 1.0 / Q[0];
  double usq = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    s.u[i] = s.v * Q[i + 1];
    usq = (i == 0) ? s.u[i] * s.u[i] : usq + s.u[i] * s.u[i];
  }
  s.p = (gamma - 1.0) * (Q[ND + 1] - 0.5 * Q[0] * usq);
  // gamma / (gamma - 1) is loop invariant: one division per kernel instead of one per point
  const double gg1 = gamma / (gamma - 1.0);
  s.T = gg1 * s.p * s.v;
}

// computeTransportVariables (reference :89-177).  FASTPOW: x^n as exp(n log x) (~2 ulp instead of pow()'s
// < 1 ulp, at well under half the instructions; T > 0 on every admissible state).  Used where it lowers the
// register pressure (adjoint sweep 1); sweep A keeps pow(), which allocates better there.
template <bool FASTPOW = false>
__device__ __forceinline__ void transport(double T, const PhysParams& pp, double& mu, double& lam, double& kap) {
  if (pp.powerLaw <= 0.0) {
    mu = pp.ReInv;
    lam = (pp.bulkRatio - 2.0 / 3.0) * pp.ReInv;
    kap = pp.ReInv * pp.PrInv;
  } else {
    if (FASTPOW) mu = exp(pp.powerLaw * log((pp.gamma - 1.0) * T)) * pp.ReInv;
    else mu = pow((pp.gamma - 1.0) * T, pp.powerLaw) * pp.ReInv;
    lam = (pp.bulkRatio - 2.0 / 3.0) * mu;
    kap = mu * pp.PrInv;
  }
}

// computeStressTensor, in-place form (reference :412-452).  g[j + ND*c] = d u_c / d x_j.
template <int ND>
__device__ __forceinline__ void stress_from_gradient(const double* g, double mu, double lam, double* s) {
  if (ND == 1) {
    s[0] = (2.0 * mu + lam) * g[0];
  } else if (ND == 2) {
    const double div = lam * (g[0] + g[3]);
    s[0] = 2.0 * mu * g[0] + div;
    s[1] = mu * (g[1] + g[2]);
    s[2] = s[1];
    s[3] = 2.0 * mu * g[3] + div;
  } else {
    const double div = lam * (g[0] + g[4] + g[8]);
    s[0] = 2.0 * mu * g[0] + div;
    s[1] = mu * (g[1] + g[3]);
    s[2] = mu * (g[2] + g[6]);
    s[3] = s[1];
    s[4] = 2.0 * mu * g[4] + div;
    s[5] = mu * (g[5] + g[7]);
    s[6] = s[2];
    s[7] = s[5];
    s[8] = 2.0 * mu * g[8] + div;
  }
}

// Cartesian total flux in direction l: inviscid (reference :563-619) minus viscous (:621-689).
// tau[l + ND*c] is the stress tensor in the reference layout, q the heat flux.
template <int ND>
__device__ __forceinline__ void cartesian_flux(int l, const double* Q, const Prim<ND>& s, bool viscous,
                                               const double* tau, const double* q, double* F, double* Fv) {
  F[0] = Q[l + 1];
#pragma unroll
  for (int c = 0; c < ND; ++c) {
    if (c == l) F[c + 1] = Q[l + 1] * s.u[l] + s.p;
    else {
      const int lo = c < l ? c : l, hi = c < l ? l : c;
      F[c + 1] = Q[lo + 1] * s.u[hi];
    }
  }
  F[ND + 1] = s.u[l] * (Q[ND + 1] + s.p);
  if (viscous) {
    double acc = 0.0;
    Fv[0] = 0.0;
#pragma unroll
    for (int c = 0; c < ND; ++c) {
      const double t = tau[l + ND * c];
      Fv[c + 1] = t;
      acc = (c == 0) ? s.u[c] * t : acc + s.u[c] * t;
    }
    Fv[ND + 1] = acc - q[l];
#pragma unroll
    for (int c = 0; c < ND + 2; ++c) F[c] = F[c] - Fv[c];
  }
}

// y += A^T x with A = Jacobian of the inviscid flux along metrics m (reference :984-1444), minus
// (if viscous) the first-partial viscous Jacobian (:2344-2600).  x, y have NU entries.
template <int ND>
__device__ __forceinline__ void flux_jacobian_matrix(const Prim<ND>& s, const double* m, double gamma, bool viscous,
                                                     double powerLaw, const double* tau, const double* q,
                                                     double (*A)[ND + 2]) {
  constexpr int NU = ND + 2;
  double uh = 0.0, usq = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    uh = (i == 0) ? m[0] * s.u[0] : uh + m[i] * s.u[i];
    usq = (i == 0) ? s.u[0] * s.u[0] : usq + s.u[i] * s.u[i];
  }
  const double phi2 = 0.5 * (gamma - 1.0) * usq;
  A[0][0] = 0.0;
#pragma unroll
  for (int a = 0; a < ND; ++a) A[a + 1][0] = phi2 * m[a] - uh * s.u[a];
  A[NU - 1][0] = uh * ((gamma - 2.0) / (gamma - 1.0) * phi2 - s.T);
#pragma unroll
  for (int b = 0; b < ND; ++b) {
    A[0][b + 1] = m[b];
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      if (a == b) A[a + 1][b + 1] = uh - (gamma - 2.0) * s.u[a] * m[a];
      else A[a + 1][b + 1] = s.u[a] * m[b] - (gamma - 1.0) * s.u[b] * m[a];
    }
    A[NU - 1][b + 1] = (s.T + phi2 / (gamma - 1.0)) * m[b] - (gamma - 1.0) * uh * s.u[b];
  }
  A[0][NU - 1] = 0.0;
#pragma unroll
  for (int a = 0; a < ND; ++a) A[a + 1][NU - 1] = (gamma - 1.0) * m[a];
  A[NU - 1][NU - 1] = gamma * uh;
  if (viscous) {
    double cst[ND];
    double chf = 0.0, ucst = 0.0;
#pragma unroll
    for (int c = 0; c < ND; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < ND; ++l) acc = (l == 0) ? m[0] * tau[0 + ND * c] : acc + m[l] * tau[l + ND * c];
      cst[c] = acc;
    }
#pragma unroll
    for (int l = 0; l < ND; ++l) chf = (l == 0) ? m[0] * q[0] : chf + m[l] * q[l];
#pragma unroll
    for (int c = 0; c < ND; ++c) ucst = (c == 0) ? s.u[0] * cst[0] : ucst + s.u[c] * cst[c];
    const double temp1 = ucst - chf;
    double temp2 = powerLaw * gamma * s.v / s.T * (phi2 / (gamma - 1.0) - s.T / gamma);
#pragma unroll
    for (int c = 0; c < ND; ++c) A[c + 1][0] -= temp2 * cst[c];
    A[NU - 1][0] -= temp2 * temp1 - s.v * ucst;
#pragma unroll
    for (int b = 0; b < ND; ++b) {
      temp2 = -powerLaw * gamma * s.v / s.T * s.u[b];
#pragma unroll
      for (int c = 0; c < ND; ++c) A[c + 1][b + 1] -= temp2 * cst[c];
      A[NU - 1][b + 1] -= temp2 * temp1 + s.v * cst[b];
    }
    temp2 = powerLaw * gamma * s.v / s.T;
#pragma unroll
    for (int c = 0; c < ND; ++c) A[c + 1][NU - 1] -= temp2 * cst[c];
    A[NU - 1][NU - 1] -= temp2 * temp1;
  }
}

template <int ND>
__device__ __forceinline__ void add_flux_jacobian_transpose(const double* Q, const Prim<ND>& s, const double* m,
                                                            double gamma, bool viscous, double powerLaw,
                                                            const double* tau, const double* q, const double* x,
                                                            double* y, double scale = 1.0) {
  constexpr int NU = ND + 2;
  (void)Q;
  double A[NU][NU];
  flux_jacobian_matrix<ND>(s, m, gamma, viscous, powerLaw, tau, q, A);
#pragma unroll
  for (int j = 0; j < NU; ++j) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < NU; ++i) acc += A[i][j] * x[i];
    y[j] += scale * acc;
  }
}

// y += A x (the linearized fluxes, reference src/RhsHelperImpl.f90:660-690, :736-760)
template <int ND>
__device__ __forceinline__ void add_flux_jacobian_apply(const Prim<ND>& s, const double* m, double gamma, bool viscous,
                                                        double powerLaw, const double* tau, const double* q,
                                                        const double* x, double* y) {
  constexpr int NU = ND + 2;
  double A[NU][NU];
  flux_jacobian_matrix<ND>(s, m, gamma, viscous, powerLaw, tau, q, A);
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < NU; ++j) acc += A[i][j] * x[j];
    y[i] += acc;
  }
}

// y += B^T x with B the second-partial viscous Jacobian for metric rows m1 (first direction) and m2
// (second direction), already times the inverse Jacobian (reference :2602-2756). x, y: ND+1 entries.
template <int ND>
__device__ __forceinline__ void add_second_partial_transpose(const double* u, double mu, double lam, double kap,
                                                             double jac, const double* m1, const double* m2,
                                                             const double* x, double* y, double scale = 1.0) {
  double temp1 = 0.0, d1 = 0.0, d2 = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    temp1 = (i == 0) ? m1[0] * m2[0] : temp1 + m1[i] * m2[i];
    d2 = (i == 0) ? m2[0] * u[0] : d2 + m2[i] * u[i];
    d1 = (i == 0) ? m1[0] * u[0] : d1 + m1[i] * u[i];
  }
  const double temp2 = mu * d2, temp3 = lam * d1;
#pragma unroll
  for (int b = 0; b < ND; ++b) {
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      const double Bab = (a == b) ? mu * temp1 + (mu + lam) * m1[a] * m2[a]
                                  : mu * m1[b] * m2[a] + lam * m1[a] * m2[b];
      acc += jac * Bab * x[a];
    }
    const double Blast = mu * temp1 * u[b] + m1[b] * temp2 + m2[b] * temp3;
    acc += jac * Blast * x[ND];
    y[b] += scale * acc;
  }
  y[ND] += scale * (jac * (kap * temp1) * x[ND]);
}

// y += B x, same matrix (linearized viscous fluxes, reference src/RhsHelperImpl.f90:762-790)
template <int ND>
__device__ __forceinline__ void add_second_partial_apply(const double* u, double mu, double lam, double kap, double jac,
                                                         const double* m1, const double* m2, const double* x,
                                                         double* y) {
  double temp1 = 0.0, d1 = 0.0, d2 = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    temp1 = (i == 0) ? m1[0] * m2[0] : temp1 + m1[i] * m2[i];
    d2 = (i == 0) ? m2[0] * u[0] : d2 + m2[i] * u[i];
    d1 = (i == 0) ? m1[0] * u[0] : d1 + m1[i] * u[i];
  }
  const double temp2 = mu * d2, temp3 = lam * d1;
  double last = 0.0;
#pragma unroll
  for (int a = 0; a < ND; ++a) {
    double acc = 0.0;
#pragma unroll
    for (int b = 0; b < ND; ++b) {
      const double Bab = (a == b) ? mu * temp1 + (mu + lam) * m1[a] * m2[a]
                                  : mu * m1[b] * m2[a] + lam * m1[a] * m2[b];
      acc += jac * Bab * x[b];
    }
    y[a] += acc;
    const double Blast = mu * temp1 * u[a] + m1[a] * temp2 + m2[a] * temp3;
    last += jac * Blast * x[a];
  }
  y[ND] += last + jac * (kap * temp1) * x[ND];
}

// ---- rectilinear specialisations: the metric row of direction D is md * e_D, so most entries of the
// Jacobians vanish.  Same arithmetic as the general forms above with the zero terms dropped.
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (N > 0) {
    static_for<N - 1>(f);
    f(std::integral_constant<int, N - 1>{});
  }
}

template <int ND, int D>
__device__ __forceinline__ void add_flux_jacobian_transpose_rect(const Prim<ND>& s, double md, double gamma,
                                                                 bool viscous, double powerLaw, const double* tau,
                                                                 const double* q, const double* x, double* y) {
  constexpr int NU = ND + 2;
  const double g1 = gamma - 1.0;
  double usq = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) usq = (i == 0) ? s.u[0] * s.u[0] : usq + s.u[i] * s.u[i];
  const double uh = md * s.u[D];
  const double phi2 = 0.5 * g1 * usq;
  double cst[ND], temp1 = 0.0, ucst = 0.0, pw = 0.0;
  if (viscous) {
#pragma unroll
    for (int c = 0; c < ND; ++c) cst[c] = md * tau[D + ND * c];
    const double chf = md * q[D];
#pragma unroll
    for (int c = 0; c < ND; ++c) ucst = (c == 0) ? s.u[0] * cst[0] : ucst + s.u[c] * cst[c];
    temp1 = ucst - chf;
    pw = powerLaw * gamma * s.v / s.T;
  }
  {  // column 0
    const double t2 = pw * (phi2 / g1 - s.T / gamma);
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      double A = (a == D) ? phi2 * md - uh * s.u[a] : -(uh * s.u[a]);
      if (viscous) A -= t2 * cst[a];
      acc += A * x[a + 1];
    }
    double Al = uh * ((gamma - 2.0) / g1 * phi2 - s.T);
    if (viscous) Al -= t2 * temp1 - s.v * ucst;
    acc += Al * x[NU - 1];
    y[0] += acc;
  }
#pragma unroll
  for (int b = 0; b < ND; ++b) {  // columns 1..ND
    const double t2 = -pw * s.u[b];
    double acc = (b == D) ? md * x[0] : 0.0;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      double A = 0.0;
      bool has = true;
      if (a == b) A = (a == D) ? uh - (gamma - 2.0) * s.u[a] * md : uh;
      else if (b == D) A = s.u[a] * md;
      else if (a == D) A = -(g1 * s.u[b] * md);
      else has = false;
      if (viscous) A = has ? A - t2 * cst[a] : -(t2 * cst[a]);
      if (has || viscous) acc += A * x[a + 1];
    }
    double Al = (b == D) ? (s.T + phi2 / g1) * md - g1 * uh * s.u[b] : -(g1 * uh * s.u[b]);
    if (viscous) Al -= t2 * temp1 + s.v * cst[b];
    acc += Al * x[NU - 1];
    y[b + 1] += acc;
  }
  {  // column NU-1
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      if (a == D) {
        double A = g1 * md;
        if (viscous) A -= pw * cst[a];
        acc += A * x[a + 1];
      } else if (viscous) {
        acc += -(pw * cst[a]) * x[a + 1];
      }
    }
    double Al = gamma * uh;
    if (viscous) Al -= pw * temp1;
    acc += Al * x[NU - 1];
    y[NU - 1] += acc;
  }
}

// Point-invariant factors of the closed-form Jacobian-transpose products below.
template <int ND>
struct JacFactors {
  double phi2;      // (gamma-1)/2 |u|^2
  double H;         // T + phi2/(gamma-1)  (total enthalpy)
  double pw;        // powerLaw gamma v / T           (viscous)
};
template <int ND>
__device__ __forceinline__ void jac_factors(const Prim<ND>& s, double gamma, bool viscous, double powerLaw,
                                            JacFactors<ND>& f) {
  const double g1 = gamma - 1.0;
  double usq = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) usq = (i == 0) ? s.u[0] * s.u[0] : usq + s.u[i] * s.u[i];
  f.phi2 = 0.5 * g1 * usq;
  f.H = s.T + f.phi2 / g1;
  f.pw = viscous ? powerLaw * gamma * s.v / s.T : 0.0;
}

// y += (A - B)^T x in closed form: A = inviscid flux Jacobian along the metric row m (reference
// src/CNSHelperImpl.f90:984-1444), B = first-partial viscous Jacobian (:2344-2600).  A^T x is the gradient of
// x . F(Q) with x frozen, which collapses to three dot products (u.x_m, m.x_m, m.u) and O(NU) work instead of
// building the NU x NU matrix.  RECT: m = m[D] e_D (off-diagonal metrics are structurally zero).
// cst[c] = sum_l m_l tau(l,c), chf = m . q   (only read when viscous).
template <int ND, bool RECT, int D>
__device__ __forceinline__ void add_flux_jacobian_transpose_cf(const Prim<ND>& s, const JacFactors<ND>& f,
                                                               const double* m, double gamma, bool viscous,
                                                               const double* cst, double chf, const double* x,
                                                               double* y) {
  constexpr int NU = ND + 2;
  const double g1 = gamma - 1.0;
  const double* xm = x + 1;
  const double xE = x[NU - 1];
  double uh, xh, ux = 0.0;
  if constexpr (RECT) {
    uh = m[D] * s.u[D];
    xh = m[D] * xm[D];
  } else {
    uh = 0.0;
    xh = 0.0;
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      uh = (i == 0) ? m[0] * s.u[0] : uh + m[i] * s.u[i];
      xh = (i == 0) ? m[0] * xm[0] : xh + m[i] * xm[i];
    }
  }
#pragma unroll
  for (int i = 0; i < ND; ++i) ux = (i == 0) ? s.u[0] * xm[0] : ux + s.u[i] * xm[i];
  const double sfac = x[0] + ux + xE * f.H;
  const double tfac = g1 * (xh + xE * uh);
  // (gamma-2)/(gamma-1) phi2 - T  ==  phi2 - H
  double y0 = f.phi2 * xh - uh * ux + xE * (uh * (f.phi2 - f.H));
  double yE = g1 * xh + gamma * (uh * xE);
  double yb[ND];
#pragma unroll
  for (int b = 0; b < ND; ++b) {
    yb[b] = uh * xm[b] - tfac * s.u[b];
    if (!RECT || b == D) yb[b] += m[b] * sfac;
  }
  if (viscous) {
    double ucst = 0.0, S = 0.0;
#pragma unroll
    for (int c = 0; c < ND; ++c) {
      ucst = (c == 0) ? s.u[0] * cst[0] : ucst + s.u[c] * cst[c];
      S = (c == 0) ? cst[0] * xm[0] : S + cst[c] * xm[c];
    }
    const double temp1 = ucst - chf;
    S += temp1 * xE;
    const double vx = s.v * xE;
    // pw (phi2/(gamma-1) - T/gamma), phi2/(gamma-1) == H - T
    const double t2 = f.pw * ((f.H - s.T) - s.T * (1.0 / gamma));
    y0 -= t2 * S - vx * ucst;
    const double pS = f.pw * S;
#pragma unroll
    for (int b = 0; b < ND; ++b) yb[b] += pS * s.u[b] - vx * cst[b];
    yE -= pS;
  }
  y[0] += y0;
#pragma unroll
  for (int b = 0; b < ND; ++b) y[b + 1] += yb[b];
  y[NU - 1] += yE;
}

// Second-partial viscous Jacobian transpose for m1 = M1 e_II, m2 = M2 e_JJ.
template <int ND, int II, int JJ>
__device__ __forceinline__ void add_second_partial_transpose_rect(const double* u, double mu, double lam, double kap,
                                                                  double jac, double M1, double M2, const double* x,
                                                                  double* y) {
  constexpr bool SAME = II == JJ;
  const double temp1 = SAME ? M1 * M2 : 0.0;
  const double temp2 = mu * (M2 * u[JJ]), temp3 = lam * (M1 * u[II]);
#pragma unroll
  for (int b = 0; b < ND; ++b) {
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      if (a == b) {
        if (SAME) {
          const double Bab = (a == II) ? mu * temp1 + (mu + lam) * M1 * M2 : mu * temp1;
          acc += jac * Bab * x[a];
        }
      } else if (b == II && a == JJ) {
        acc += jac * (mu * M1 * M2) * x[a];
      } else if (a == II && b == JJ) {
        acc += jac * (lam * M1 * M2) * x[a];
      }
    }
    double Blast = 0.0;
    bool has = false;
    if (SAME) { Blast = mu * temp1 * u[b]; has = true; }
    if (b == II) { Blast = has ? Blast + M1 * temp2 : M1 * temp2; has = true; }
    if (b == JJ) { Blast = has ? Blast + M2 * temp3 : M2 * temp3; has = true; }
    if (has) acc += jac * Blast * x[ND];
    y[b] += acc;
  }
  if (SAME) y[ND] += jac * (kap * temp1) * x[ND];
}

// Same product with the point-invariant factors folded: jm = mu/J', jl = lambda/J', jk = kappa/J' (J' = the
// reference's `jacobian`, i.e. the inverse Jacobian, already multiplied in) and MM = M1 * M2.
template <int ND, int II, int JJ>
__device__ __forceinline__ void add_second_partial_transpose_rect_f(const double* u, double jm, double jl, double jk,
                                                                    double MM, const double* x, double* y) {
  const double xE = x[ND];
  if constexpr (II == JJ) {
    const double a = jm * MM;
    double ux = 0.0;
#pragma unroll
    for (int b = 0; b < ND; ++b) y[b] += a * (x[b] + u[b] * xE);
    (void)ux;
    y[II] += (jm + jl) * MM * (x[II] + u[II] * xE);
    y[ND] += jk * MM * xE;
  } else {
    y[II] += jm * MM * (x[JJ] + u[JJ] * xE);
    y[JJ] += jl * MM * (x[II] + u[II] * xE);
  }
}

// Incoming part of the inviscid flux Jacobian, A+ = R max/min(Lambda,0) L (reference :1446-2342).
template <int ND>
__device__ void incoming_jacobian(const double* Q, const double* m, double gamma, int incomingDirection,
                                  double (*A)[ND + 2]) {
  constexpr int NU = ND + 2;
  Prim<ND> s;
  // the reference recomputes T as gamma*(v*rhoE - |u|^2/2) here (no pressure intermediate)
  s.v = 1.0 / Q[0];
  double usq = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    s.u[i] = s.v * Q[i + 1];
    usq = (i == 0) ? s.u[i] * s.u[i] : usq + s.u[i] * s.u[i];
  }
  s.T = gamma * (s.v * Q[ND + 1] - 0.5 * usq);
  double arc = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) arc = (i == 0) ? m[0] * m[0] : arc + m[i] * m[i];
  arc = (ND == 1) ? fabs(m[0]) : sqrt(arc);
  double n[ND];
  double uh = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    n[i] = m[i] / arc;
    uh = (i == 0) ? n[0] * s.u[0] : uh + n[i] * s.u[i];
  }
  const double c = sqrt((gamma - 1.0) * s.T);
  const double phi2 = 0.5 * (gamma - 1.0) * usq;
  const double g1 = gamma - 1.0;
  const double rho = Q[0], v = s.v, T = s.T;
  double ev[NU];
#pragma unroll
  for (int i = 0; i < ND; ++i) ev[i] = uh;
  ev[ND] = uh + c;
  ev[ND + 1] = uh - c;
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    ev[i] = arc * ev[i];
    if (incomingDirection * ev[i] < 0.0) ev[i] = 0.0;
  }
  double R[NU][NU], L[NU][NU];
  const double* u = s.u;
  if (ND == 1) {
    R[0][0] = 1.0; R[1][0] = u[0]; R[2][0] = phi2 / g1;
    R[0][1] = 1.0; R[1][1] = u[0] + n[0] * c; R[2][1] = T + phi2 / g1 + c * uh;
    R[0][2] = 1.0; R[1][2] = u[0] - n[0] * c; R[2][2] = T + phi2 / g1 - c * uh;
    L[0][0] = 1.0 - phi2 / (c * c);
    L[1][0] = 0.5 * (phi2 / (c * c) - uh / c);
    L[2][0] = 0.5 * (phi2 / (c * c) + uh / c);
    L[0][1] = u[0] / T;
    L[1][1] = -0.5 * (u[0] / T - n[0] / c);
    L[2][1] = -0.5 * (u[0] / T + n[0] / c);
    L[0][2] = -1.0 / T; L[1][2] = 0.5 / T; L[2][2] = 0.5 / T;
  } else if (ND == 2) {
    const double n1 = n[0], n2 = n[ND > 1 ? 1 : 0], u1 = u[0], u2 = u[ND > 1 ? 1 : 0];
    R[0][0] = 1.0; R[1][0] = u1; R[2][0] = u2; R[3 % NU][0] = phi2 / g1;
    R[0][1] = 0.0; R[1][1] = n2 * rho; R[2][1] = -n1 * rho; R[3 % NU][1] = rho * (n2 * u1 - n1 * u2);
    R[0][2] = 1.0; R[1][2] = u1 + n1 * c; R[2][2] = u2 + n2 * c; R[3 % NU][2] = T + phi2 / g1 + c * uh;
    R[0][3 % NU] = 1.0; R[1][3 % NU] = u1 - n1 * c; R[2][3 % NU] = u2 - n2 * c;
    R[3 % NU][3 % NU] = T + phi2 / g1 - c * uh;
    L[0][0] = 1.0 - phi2 / (c * c);
    L[1][0] = -v * (n2 * u1 - n1 * u2);
    L[2][0] = 0.5 * (phi2 / (c * c) - uh / c);
    L[3 % NU][0] = 0.5 * (phi2 / (c * c) + uh / c);
    L[0][1] = u1 / T; L[1][1] = v * n2;
    L[2][1] = -0.5 * (u1 / T - n1 / c);
    L[3 % NU][1] = -0.5 * (u1 / T + n1 / c);
    L[0][2] = u2 / T; L[1][2] = -v * n1;
    L[2][2] = -0.5 * (u2 / T - n2 / c);
    L[3 % NU][2] = -0.5 * (u2 / T + n2 / c);
    L[0][3 % NU] = -1.0 / T; L[1][3 % NU] = 0.0; L[2][3 % NU] = 0.5 / T; L[3 % NU][3 % NU] = 0.5 / T;
  } else {
    constexpr int I3 = 3 % NU, I4 = 4 % NU;
    const double n1 = n[0], n2 = n[1 % ND], n3 = n[2 % ND];
    const double u1 = u[0], u2 = u[1 % ND], u3 = u[2 % ND];
    R[0][0] = n1; R[1][0] = n1 * u1; R[2][0] = n1 * u2 + rho * n3; R[I3][0] = n1 * u3 - rho * n2;
    R[I4][0] = rho * (n3 * u2 - n2 * u3) + phi2 / g1 * n1;
    R[0][1] = n2; R[1][1] = n2 * u1 - rho * n3; R[2][1] = n2 * u2; R[I3][1] = n2 * u3 + rho * n1;
    R[I4][1] = rho * (n1 * u3 - n3 * u1) + phi2 / g1 * n2;
    R[0][2] = n3; R[1][2] = n3 * u1 + rho * n2; R[2][2] = n3 * u2 - rho * n1; R[I3][2] = n3 * u3;
    R[I4][2] = rho * (n2 * u1 - n1 * u2) + phi2 / g1 * n3;
    R[0][I3] = 1.0; R[1][I3] = u1 + n1 * c; R[2][I3] = u2 + n2 * c; R[I3][I3] = u3 + n3 * c;
    R[I4][I3] = T + phi2 / g1 + c * uh;
    R[0][I4] = 1.0; R[1][I4] = u1 - n1 * c; R[2][I4] = u2 - n2 * c; R[I3][I4] = u3 - n3 * c;
    R[I4][I4] = T + phi2 / g1 - c * uh;
    const double w = 1.0 - phi2 / (c * c);
    L[0][0] = n1 * w - v * (n3 * u2 - n2 * u3);
    L[1][0] = n2 * w - v * (n1 * u3 - n3 * u1);
    L[2][0] = n3 * w - v * (n2 * u1 - n1 * u2);
    L[I3][0] = 0.5 * (phi2 / (c * c) - uh / c);
    L[I4][0] = 0.5 * (phi2 / (c * c) + uh / c);
    L[0][1] = n1 * u1 / T; L[1][1] = n2 * u1 / T - v * n3; L[2][1] = n3 * u1 / T + v * n2;
    L[I3][1] = -0.5 * (u1 / T - n1 / c); L[I4][1] = -0.5 * (u1 / T + n1 / c);
    L[0][2] = n1 * u2 / T + v * n3; L[1][2] = n2 * u2 / T; L[2][2] = n3 * u2 / T - v * n1;
    L[I3][2] = -0.5 * (u2 / T - n2 / c); L[I4][2] = -0.5 * (u2 / T + n2 / c);
    L[0][I3] = n1 * u3 / T - v * n2; L[1][I3] = n2 * u3 / T + v * n1; L[2][I3] = n3 * u3 / T;
    L[I3][I3] = -0.5 * (u3 / T - n3 / c); L[I4][I3] = -0.5 * (u3 / T + n3 / c);
    L[0][I4] = -n1 / T; L[1][I4] = -n2 / T; L[2][I4] = -n3 / T; L[I3][I4] = 0.5 / T; L[I4][I4] = 0.5 / T;
  }
#pragma unroll
  for (int j = 0; j < NU; ++j)
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < NU; ++k) acc += R[i][k] * ev[k] * L[k][j];
      A[i][j] = acc;
    }
}
