// Block-interface SAT coupling on the device (t_BlockInterfacePatch + InterfaceHelper).
// Reference: src/BlockInterfacePatchImpl.f90:127-538 (addBlockInterfacePenalty, forward and discrete adjoint),
// :592-810 (collectInterfaceData / disperseInterfaceData incl. the METRICS pseudo-mode), :812-929
// (reshapeReceivedData: index reordering), src/InterfaceHelperImpl.f90:3-239 (interface links, exchange),
// src/CNSHelperImpl.f90:179-351 (Roe average and its variation), :1446-2342 (incoming Jacobian and its
// variation), src/RhsHelperImpl.f90:831-1026 (viscous interface adjoint penalty).
//
// The reference funnels every interface through the patch-master ranks (gather -> Isend/Recv -> barrier ->
// reshape -> scatter).  Here both sides of an interface live in device memory of this process: the exchange is
// ONE kernel per patch that reads the partner's collected face data and applies the index reordering on the
// fly (no staging, no barrier).  The variation of A+ with respect to the local state - which the reference
// differentiates by hand, operation by operation - is evaluated with forward-mode dual numbers over the same
// operations.
#include <cstring>

#include "grid.h"
#include "patches.h"

namespace {

inline unsigned nblocks(size_t n) { return (unsigned)((n + 127) / 128); }

struct PatchGeom {
  int lo[3], sz[3];
  int nx, ny;
  int n;
  __device__ size_t gridIndex(int q) const {
    const int i = q % sz[0], j = (q / sz[0]) % sz[1], k = q / (sz[0] * sz[1]);
    return (size_t)(lo[0] + i) + (size_t)nx * ((size_t)(lo[1] + j) + (size_t)ny * (size_t)(lo[2] + k));
  }
};

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

// ------------------------------------------------------------------------------------------ dual numbers
template <int M>
struct Dual {
  double v;
  double d[M];
};
#define MG_DUAL_LOOP for (int e_ = 0; e_ < M; ++e_)
template <int M> __device__ __forceinline__ Dual<M> dconst(double x) { Dual<M> r; r.v = x; MG_DUAL_LOOP r.d[e_] = 0.0; return r; }
template <int M> __device__ __forceinline__ Dual<M> operator+(const Dual<M>& a, const Dual<M>& b) { Dual<M> r; r.v = a.v + b.v; MG_DUAL_LOOP r.d[e_] = a.d[e_] + b.d[e_]; return r; }
template <int M> __device__ __forceinline__ Dual<M> operator-(const Dual<M>& a, const Dual<M>& b) { Dual<M> r; r.v = a.v - b.v; MG_DUAL_LOOP r.d[e_] = a.d[e_] - b.d[e_]; return r; }
template <int M> __device__ __forceinline__ Dual<M> operator-(const Dual<M>& a) { Dual<M> r; r.v = -a.v; MG_DUAL_LOOP r.d[e_] = -a.d[e_]; return r; }
template <int M> __device__ __forceinline__ Dual<M> operator*(const Dual<M>& a, const Dual<M>& b) { Dual<M> r; r.v = a.v * b.v; MG_DUAL_LOOP r.d[e_] = a.d[e_] * b.v + b.d[e_] * a.v; return r; }
template <int M> __device__ __forceinline__ Dual<M> operator/(const Dual<M>& a, const Dual<M>& b) { Dual<M> r; r.v = a.v / b.v; MG_DUAL_LOOP r.d[e_] = (a.d[e_] - b.d[e_] * r.v) / b.v; return r; }
template <int M> __device__ __forceinline__ Dual<M> operator+(const Dual<M>& a, double b) { Dual<M> r = a; r.v += b; return r; }
template <int M> __device__ __forceinline__ Dual<M> operator+(double b, const Dual<M>& a) { return a + b; }
template <int M> __device__ __forceinline__ Dual<M> operator-(const Dual<M>& a, double b) { Dual<M> r = a; r.v -= b; return r; }
template <int M> __device__ __forceinline__ Dual<M> operator-(double b, const Dual<M>& a) { Dual<M> r; r.v = b - a.v; MG_DUAL_LOOP r.d[e_] = -a.d[e_]; return r; }
template <int M> __device__ __forceinline__ Dual<M> operator*(const Dual<M>& a, double b) { Dual<M> r; r.v = a.v * b; MG_DUAL_LOOP r.d[e_] = a.d[e_] * b; return r; }
template <int M> __device__ __forceinline__ Dual<M> operator*(double b, const Dual<M>& a) { return a * b; }
template <int M> __device__ __forceinline__ Dual<M> operator/(const Dual<M>& a, double b) { Dual<M> r; r.v = a.v / b; MG_DUAL_LOOP r.d[e_] = a.d[e_] / b; return r; }
template <int M> __device__ __forceinline__ Dual<M> operator/(double b, const Dual<M>& a) { Dual<M> r; r.v = b / a.v; const double f = r.v / a.v; MG_DUAL_LOOP r.d[e_] = -a.d[e_] * f; return r; }
template <int M> __device__ __forceinline__ Dual<M> dsqrt(const Dual<M>& a) { Dual<M> r; r.v = sqrt(a.v); const double f = 0.5 / r.v; MG_DUAL_LOOP r.d[e_] = a.d[e_] * f; return r; }

// computeRoeAverage (reference src/CNSHelperImpl.f90:179-351) with the variation with respect to the LEFT state
// (deltaConservedVariablesL = identity): roe[c].d[l] = d roe_c / d (QL)_l
template <int ND>
__device__ void roe_average_dual(const double* QL, const double* QR, double gamma, Dual<ND + 2>* roe) {
  constexpr int NU = ND + 2;
  typedef Dual<NU> D;
  D ql[NU];
#pragma unroll
  for (int c = 0; c < NU; ++c) {
    ql[c] = dconst<NU>(QL[c]);
    ql[c].d[c] = 1.0;
  }
  const D sL = dsqrt(ql[0]);
  const double sR = sqrt(QR[0]);
  const D vL = 1.0 / ql[0];
  const double vR = 1.0 / QR[0];
  D mL = ql[1] * ql[1];
  double mR = QR[1] * QR[1];
#pragma unroll
  for (int i = 1; i < ND; ++i) { mL = mL + ql[i + 1] * ql[i + 1]; mR += QR[i + 1] * QR[i + 1]; }
  const D hL = gamma * ql[NU - 1] - (0.5 * (gamma - 1.0)) * (vL * mL);
  const double hR = gamma * QR[NU - 1] - 0.5 * (gamma - 1.0) * vR * mR;
  const D den = sL + sR;
  roe[0] = sL * sR;
#pragma unroll
  for (int i = 0; i < ND; ++i) roe[i + 1] = (sR * ql[i + 1] + sL * QR[i + 1]) / den;
  D h = (sR * hL + sL * hR) / den;
  D msq = roe[1] * roe[1];
#pragma unroll
  for (int i = 1; i < ND; ++i) msq = msq + roe[i + 1] * roe[i + 1];
  roe[NU - 1] = (h + (0.5 * (gamma - 1.0)) * (msq / roe[0])) / gamma;
}

// computeRoeAverage with deltaConservedVariablesL = diag(dQL), deltaConservedVariablesR = diag(dQR) (the LINEARIZED
// branch of addBlockInterfacePenalty, reference src/BlockInterfacePatchImpl.f90:470-481): summed over its columns the
// reference's deltaRoeAverage is the directional derivative of the Roe state along (dQL, dQR) = roe[c].d[0]
template <int ND>
__device__ void roe_average_directional(const double* QL, const double* QR, const double* dQL, const double* dQR,
                                        double gamma, Dual<1>* roe) {
  constexpr int NU = ND + 2;
  typedef Dual<1> D;
  D ql[NU], qr[NU];
#pragma unroll
  for (int c = 0; c < NU; ++c) {
    ql[c].v = QL[c]; ql[c].d[0] = dQL[c];
    qr[c].v = QR[c]; qr[c].d[0] = dQR[c];
  }
  const D sL = dsqrt(ql[0]), sR = dsqrt(qr[0]);
  const D vL = 1.0 / ql[0], vR = 1.0 / qr[0];
  D mL = ql[1] * ql[1], mR = qr[1] * qr[1];
#pragma unroll
  for (int i = 1; i < ND; ++i) { mL = mL + ql[i + 1] * ql[i + 1]; mR = mR + qr[i + 1] * qr[i + 1]; }
  const D hL = gamma * ql[NU - 1] - (0.5 * (gamma - 1.0)) * (vL * mL);
  const D hR = gamma * qr[NU - 1] - (0.5 * (gamma - 1.0)) * (vR * mR);
  const D den = sL + sR;
  roe[0] = sL * sR;
#pragma unroll
  for (int i = 0; i < ND; ++i) roe[i + 1] = (sR * ql[i + 1] + sL * qr[i + 1]) / den;
  D h = (sR * hL + sL * hR) / den;
  D msq = roe[1] * roe[1];
#pragma unroll
  for (int i = 1; i < ND; ++i) msq = msq + roe[i + 1] * roe[i + 1];
  roe[NU - 1] = (h + (0.5 * (gamma - 1.0)) * (msq / roe[0])) / gamma;
}

// computeIncomingJacobianOfInviscidFlux{1,2,3}D with its variation (reference :1446-2342): the operations of
// cns_device.cuh:incoming_jacobian on dual numbers.
template <int ND, int M>
__device__ void incoming_jacobian_dual(const Dual<M>* Q, const double* m, double gamma, int incomingDirection,
                                       Dual<M> (*A)[ND + 2]) {
  constexpr int NU = ND + 2;
  typedef Dual<M> D;
  const D rho = Q[0];
  const D v = 1.0 / rho;
  D u[ND];
  D usq = dconst<M>(0.0);
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    u[i] = v * Q[i + 1];
    usq = (i == 0) ? u[0] * u[0] : usq + u[i] * u[i];
  }
  const D T = gamma * (v * Q[NU - 1] - 0.5 * usq);
  double arc = 0.0;
#pragma unroll
  for (int i = 0; i < ND; ++i) arc = (i == 0) ? m[0] * m[0] : arc + m[i] * m[i];
  arc = (ND == 1) ? fabs(m[0]) : sqrt(arc);
  double n[ND];
  D uh = dconst<M>(0.0);
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    n[i] = m[i] / arc;
    uh = (i == 0) ? n[0] * u[0] : uh + n[i] * u[i];
  }
  const double g1 = gamma - 1.0;
  const D c = dsqrt(g1 * T);
  const D phi2 = (0.5 * g1) * usq;
  D ev[NU];
#pragma unroll
  for (int i = 0; i < ND; ++i) ev[i] = uh;
  ev[ND] = uh + c;
  ev[ND + 1] = uh - c;
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    ev[i] = arc * ev[i];
    if (incomingDirection * ev[i].v < 0.0) ev[i] = dconst<M>(0.0);
  }
  D R[NU][NU], L[NU][NU];
  const D zero = dconst<M>(0.0), one = dconst<M>(1.0);
#pragma unroll
  for (int i = 0; i < NU; ++i)
#pragma unroll
    for (int j = 0; j < NU; ++j) { R[i][j] = zero; L[i][j] = zero; }
  const D c2 = c * c;
  const D H = T + phi2 / g1;
  if constexpr (ND == 1) {
    R[0][0] = one; R[1][0] = u[0]; R[2][0] = phi2 / g1;
    R[0][1] = one; R[1][1] = u[0] + n[0] * c; R[2][1] = H + c * uh;
    R[0][2] = one; R[1][2] = u[0] - n[0] * c; R[2][2] = H - c * uh;
    L[0][0] = 1.0 - phi2 / c2;
    L[1][0] = 0.5 * (phi2 / c2 - uh / c);
    L[2][0] = 0.5 * (phi2 / c2 + uh / c);
    L[0][1] = u[0] / T;
    L[1][1] = -0.5 * (u[0] / T - n[0] / c);
    L[2][1] = -0.5 * (u[0] / T + n[0] / c);
    L[0][2] = -1.0 / T; L[1][2] = 0.5 / T; L[2][2] = 0.5 / T;
  } else if constexpr (ND == 2) {
    const double n1 = n[0], n2 = n[1];
    const D u1 = u[0], u2 = u[1];
    R[0][0] = one; R[1][0] = u1; R[2][0] = u2; R[3][0] = phi2 / g1;
    R[0][1] = zero; R[1][1] = n2 * rho; R[2][1] = -(n1 * rho); R[3][1] = rho * (n2 * u1 - n1 * u2);
    R[0][2] = one; R[1][2] = u1 + n1 * c; R[2][2] = u2 + n2 * c; R[3][2] = H + c * uh;
    R[0][3] = one; R[1][3] = u1 - n1 * c; R[2][3] = u2 - n2 * c; R[3][3] = H - c * uh;
    L[0][0] = 1.0 - phi2 / c2;
    L[1][0] = -(v * (n2 * u1 - n1 * u2));
    L[2][0] = 0.5 * (phi2 / c2 - uh / c);
    L[3][0] = 0.5 * (phi2 / c2 + uh / c);
    L[0][1] = u1 / T; L[1][1] = v * n2;
    L[2][1] = -0.5 * (u1 / T - n1 / c);
    L[3][1] = -0.5 * (u1 / T + n1 / c);
    L[0][2] = u2 / T; L[1][2] = -(v * n1);
    L[2][2] = -0.5 * (u2 / T - n2 / c);
    L[3][2] = -0.5 * (u2 / T + n2 / c);
    L[0][3] = -1.0 / T; L[1][3] = zero; L[2][3] = 0.5 / T; L[3][3] = 0.5 / T;
  } else {
    const double n1 = n[0], n2 = n[1], n3 = n[2];
    const D u1 = u[0], u2 = u[1], u3 = u[2];
    R[0][0] = dconst<M>(n1); R[1][0] = n1 * u1; R[2][0] = n1 * u2 + rho * n3; R[3][0] = n1 * u3 - rho * n2;
    R[4][0] = rho * (n3 * u2 - n2 * u3) + (phi2 / g1) * n1;
    R[0][1] = dconst<M>(n2); R[1][1] = n2 * u1 - rho * n3; R[2][1] = n2 * u2; R[3][1] = n2 * u3 + rho * n1;
    R[4][1] = rho * (n1 * u3 - n3 * u1) + (phi2 / g1) * n2;
    R[0][2] = dconst<M>(n3); R[1][2] = n3 * u1 + rho * n2; R[2][2] = n3 * u2 - rho * n1; R[3][2] = n3 * u3;
    R[4][2] = rho * (n2 * u1 - n1 * u2) + (phi2 / g1) * n3;
    R[0][3] = one; R[1][3] = u1 + n1 * c; R[2][3] = u2 + n2 * c; R[3][3] = u3 + n3 * c; R[4][3] = H + c * uh;
    R[0][4] = one; R[1][4] = u1 - n1 * c; R[2][4] = u2 - n2 * c; R[3][4] = u3 - n3 * c; R[4][4] = H - c * uh;
    const D w = 1.0 - phi2 / c2;
    L[0][0] = n1 * w - v * (n3 * u2 - n2 * u3);
    L[1][0] = n2 * w - v * (n1 * u3 - n3 * u1);
    L[2][0] = n3 * w - v * (n2 * u1 - n1 * u2);
    L[3][0] = 0.5 * (phi2 / c2 - uh / c);
    L[4][0] = 0.5 * (phi2 / c2 + uh / c);
    L[0][1] = n1 * u1 / T; L[1][1] = n2 * u1 / T - v * n3; L[2][1] = n3 * u1 / T + v * n2;
    L[3][1] = -0.5 * (u1 / T - n1 / c); L[4][1] = -0.5 * (u1 / T + n1 / c);
    L[0][2] = n1 * u2 / T + v * n3; L[1][2] = n2 * u2 / T; L[2][2] = n3 * u2 / T - v * n1;
    L[3][2] = -0.5 * (u2 / T - n2 / c); L[4][2] = -0.5 * (u2 / T + n2 / c);
    L[0][3] = n1 * u3 / T - v * n2; L[1][3] = n2 * u3 / T + v * n1; L[2][3] = n3 * u3 / T;
    L[3][3] = -0.5 * (u3 / T - n3 / c); L[4][3] = -0.5 * (u3 / T + n3 / c);
    L[0][4] = -(n1 / T); L[1][4] = -(n2 / T); L[2][4] = -(n3 / T); L[3][4] = 0.5 / T; L[4][4] = 0.5 / T;
  }
  for (int i = 0; i < NU; ++i)
    for (int j = 0; j < NU; ++j) {
      D acc = R[i][0] * ev[0] * L[0][j];
      for (int k = 1; k < NU; ++k) acc = acc + R[i][k] * ev[k] * L[k][j];
      A[i][j] = acc;
    }
}

// ---------------------------------------------------------------------------------------------- kernels
__global__ void k_collect(PatchGeom g, const double* f, size_t cs, int nComp, double* out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= g.n) return;
  const size_t p = g.gridIndex(q);
  for (int c = 0; c < nComp; ++c) out[(size_t)c * g.n + q] = f[(size_t)c * cs + p];
}

// viscousFluxesL(:, c) = sum_j cartesianViscousFluxesL(:, c, j) metricsL(:, j), c >= 2 (reference :627-635)
__global__ void k_normal_viscous_flux(int n, int nU, int nD, const double* Fc, const double* mL, double* out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  out[q] = 0.0;
  for (int c = 1; c < nU; ++c) {
    double acc = 0.0;
    for (int j = 0; j < nD; ++j) acc += Fc[(size_t)(c + nU * j) * n + q] * mL[(size_t)j * n + q];
    out[(size_t)c * n + q] = acc;
  }
}

// exchangeInterfaceData + reshapeReceivedData (reference src/InterfaceHelperImpl.f90:115-239,
// src/BlockInterfacePatchImpl.f90:812-929): dst(i, j, k) of this patch <- the partner's point that coincides
// with it under this patch's index reordering (o1, o2).
__global__ void k_receive(int g1, int g2, int g3, int o1, int o2, int nComp, const double* src, double* dst) {
  const int n = g1 * g2 * g3;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  int i = q % g1, j = (q / g1) % g2;
  const int k = q / (g1 * g2);
  const bool transposed = (abs(o1) == 2 && abs(o2) == 1);
  int a1 = o1, a2 = o2;
  if (transposed) { a1 = o2; a2 = o1; }
  if (a1 == -1) i = g1 - 1 - i;
  if (a2 == -2) j = g2 - 1 - j;
  const int s = transposed ? j + g2 * (i + g1 * k) : i + g1 * (j + g2 * k);
  for (int c = 0; c < nComp; ++c) dst[(size_t)c * n + q] = src[(size_t)c * n + s];
}

struct IfArgs {
  PatchGeom g;
  const double *QL, *QR, *WL, *WR, *FvL, *FvR, *mL, *mR;   // patch arrays, point fastest
  const double *m, *jac, *v, *u, *T, *tau, *q;             // grid fields
  const int* iblank;
  double* rhs;
  size_t cs;
  int dir, mode, viscous, normal, normalL, normalR;
  double sigmaI, sigmaV, sigmaIL, sigmaIR, sigmaVL, sigmaVR, gamma, powerLaw;
};

// addBlockInterfacePenalty (reference src/BlockInterfacePatchImpl.f90:127-538)
template <int ND>
__global__ void __launch_bounds__(128) k_interface(IfArgs a) {
  constexpr int NU = ND + 2;
  typedef Dual<NU> D;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  const size_t p = a.g.gridIndex(q);
  if (a.iblank && a.iblank[p] == 0) return;
  const int n = a.g.n;
  const double jac = a.jac[p];
  double QL[NU], QR[NU], r[NU];
#pragma unroll
  for (int c = 0; c < NU; ++c) { QL[c] = a.QL[(size_t)c * n + q]; QR[c] = a.QR[(size_t)c * n + q]; r[c] = 0.0; }
  D roe[NU];
  roe_average_dual<ND>(QL, QR, a.gamma, roe);
  if (a.mode == MG_FORWARD) {
    double roev[NU], mm[ND], A[NU][NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) roev[c] = roe[c].v;
#pragma unroll
    for (int i = 0; i < ND; ++i) mm[i] = a.m[(size_t)(i + ND * a.dir) * a.cs + p];
    incoming_jacobian<ND>(roev, mm, a.gamma, a.normal, A);
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NU; ++j) acc += A[i][j] * (QL[j] - QR[j]);
      r[i] = -a.sigmaI * jac * acc;
    }
    if (a.viscous) {
      const double vL = copysign(a.sigmaV, (double)a.normalL), vR = copysign(a.sigmaV, (double)a.normalR);
#pragma unroll
      for (int c = 1; c < NU; ++c) r[c] += jac * (vL * a.FvL[(size_t)c * n + q] + vR * a.FvR[(size_t)c * n + q]);
    }
  } else if (a.mode == MG_LINEARIZED) {
    // :470-516: - sigmaI J [ A+ (dQL - dQR) + dA+ (QL - QR) ] + J (sigmaVL FvL + sigmaVR FvR), the perturbations
    // travel in the adjoint-variable slots and Fv is the linearized viscous flux along the normal
    double dL[NU], dR[NU], mm[ND];
#pragma unroll
    for (int c = 0; c < NU; ++c) { dL[c] = a.WL[(size_t)c * n + q]; dR[c] = a.WR[(size_t)c * n + q]; }
#pragma unroll
    for (int i = 0; i < ND; ++i) mm[i] = a.m[(size_t)(i + ND * a.dir) * a.cs + p];
    Dual<1> roe1[NU], A1[NU][NU];
    roe_average_directional<ND>(QL, QR, dL, dR, a.gamma, roe1);
    incoming_jacobian_dual<ND, 1>(roe1, mm, a.gamma, a.normal, A1);
    for (int i = 0; i < NU; ++i) {
      double acc = 0.0;
      for (int j = 0; j < NU; ++j) acc += A1[i][j].v * (dL[j] - dR[j]) + A1[i][j].d[0] * (QL[j] - QR[j]);
      r[i] = -a.sigmaI * jac * acc;
    }
    if (a.viscous) {
      const double vL = copysign(a.sigmaV, (double)a.normalL), vR = copysign(a.sigmaV, (double)a.normalR);
#pragma unroll
      for (int c = 1; c < NU; ++c) r[c] += jac * (vL * a.FvL[(size_t)c * n + q] + vR * a.FvR[(size_t)c * n + q]);
    }
  } else {
    double dq[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) dq[c] = QL[c] - QR[c];
    for (int side = 0; side < 2; ++side) {
      const double* mp = side == 0 ? a.mL : a.mR;
      const double* wp = side == 0 ? a.WL : a.WR;
      const int inc = side == 0 ? a.normalL : a.normalR;
      const double sI = (side == 0 ? a.sigmaIL : -a.sigmaIR) * jac;
      const double sV = (side == 0 ? a.sigmaVL : -a.sigmaVR) * jac;
      double mm[ND], w[NU];
#pragma unroll
      for (int i = 0; i < ND; ++i) mm[i] = mp[(size_t)i * n + q];
#pragma unroll
      for (int c = 0; c < NU; ++c) w[c] = wp[(size_t)c * n + q];
      D A[NU][NU];
      incoming_jacobian_dual<ND, NU>(roe, mm, a.gamma, inc, A);
      // + sI [ A^T w + sum_ij dA_ij/dQL_l dq_j w_i ]
      for (int l = 0; l < NU; ++l) {
        double acc = 0.0;
        for (int i = 0; i < NU; ++i) {
          acc += A[i][l].v * w[i];
          double t = 0.0;
          for (int j = 0; j < NU; ++j) t += A[i][j].d[l] * dq[j];
          acc += t * w[i];
        }
        r[l] += sI * acc;
      }
      if (a.viscous) {
        // - sV B^T w, B = first-partial viscous Jacobian about the local state along this side's metrics
        double tau[ND * ND], qq[ND], y1[NU], y2[NU];
        Prim<ND> s;
#pragma unroll
        for (int c = 0; c < NU; ++c) { y1[c] = 0.0; y2[c] = 0.0; }
        s.v = a.v[p];
        s.T = a.T[p];
        s.p = 0.0;
#pragma unroll
        for (int i = 0; i < ND; ++i) { s.u[i] = a.u[(size_t)i * a.cs + p]; qq[i] = a.q[(size_t)i * a.cs + p]; }
#pragma unroll
        for (int c = 0; c < ND * ND; ++c) tau[c] = a.tau[(size_t)c * a.cs + p];
        add_flux_jacobian_transpose<ND>(QL, s, mm, a.gamma, true, a.powerLaw, tau, qq, w, y1);    // (A-B)^T w
        add_flux_jacobian_transpose<ND>(QL, s, mm, a.gamma, false, a.powerLaw, tau, qq, w, y2);   // A^T w
#pragma unroll
        for (int c = 0; c < NU; ++c) r[c] -= sV * (y2[c] - y1[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NU; ++c) a.rhs[(size_t)c * a.cs + p] += r[c];
}

struct IfSrcArgs {
  PatchGeom g;
  const double *WL, *WR, *mL, *mR;
  const double *u, *mu, *lam, *kap, *m, *jac;
  const int* iblank;
  double* temp1;
  size_t cs;
  double sigmaVL, sigmaVR;
};

// Sources of addInterfaceAdjointPenalty (reference src/RhsHelperImpl.f90:886-965)
template <int ND>
__global__ void k_interface_adjoint_source(IfSrcArgs a) {
  constexpr int NU = ND + 2;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.g.n) return;
  const size_t p = a.g.gridIndex(q);
  if (a.iblank && a.iblank[p] == 0) return;
  const int n = a.g.n;
  double u[ND], M[ND * ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) u[i] = a.u[(size_t)i * a.cs + p];
#pragma unroll
  for (int c = 0; c < ND * ND; ++c) M[c] = a.m[(size_t)c * a.cs + p];
  const double mu = a.mu[p], lam = a.lam[p], kap = a.kap[p], jac = a.jac[p];
  for (int side = 0; side < 2; ++side) {
    const double* mp = side == 0 ? a.mL : a.mR;
    const double* wp = side == 0 ? a.WL : a.WR;
    const double sV = side == 0 ? -a.sigmaVL : a.sigmaVR;
    double m1[ND], w[ND + 1];
#pragma unroll
    for (int i = 0; i < ND; ++i) m1[i] = mp[(size_t)i * n + q];
#pragma unroll
    for (int c = 0; c < ND + 1; ++c) w[c] = wp[(size_t)(c + 1) * n + q];
#pragma unroll
    for (int l = 0; l < ND; ++l) {
      double d[ND + 1];
#pragma unroll
      for (int c = 0; c < ND + 1; ++c) d[c] = 0.0;
      add_second_partial_transpose<ND>(u, mu, lam, kap, jac, m1, &M[ND * l], w, d);
#pragma unroll
      for (int c = 0; c < ND + 1; ++c) a.temp1[(size_t)(c + (NU - 1) * l) * a.cs + p] += sV * d[c];
    }
  }
}

double* arr(mg_patch* p, const char* name) {
  auto it = p->arrays.find(name);
  return it == p->arrays.end() ? nullptr : it->second.p;
}

int collect_to(mg_patch* p, const double* f, size_t cs, int nComp, const char* name) {
  double* d = nullptr;
  MG_TRY(mg_patch_alloc_array(p, name, nComp, &d));
  if (p->nPatchPoints == 0) return 0;
  { k_collect<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(geom(p), f, cs, nComp, d); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return 0;
}

int receive(mg_patch* p, const char* partnerName, const char* myName, int nComp) {
  mg_patch* o = p->partner;
  double* src = arr(o, partnerName);
  if (!src) MG_FAIL(std::string("block interface '") + p->name + "': partner has not collected '" + partnerName + "'");
  double* dst = nullptr;
  MG_TRY(mg_patch_alloc_array(p, myName, nComp, &dst));
  if (p->nPatchPoints == 0) return 0;
  { k_receive<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(p->globalSize[0], p->globalSize[1], p->globalSize[2],
                                                              p->reorder[0], p->reorder[1], nComp, src, dst); mg_count_launches(1); }
  MG_CUDA(cudaGetLastError());
  return 0;
}

// remote partner: ship the named patch arrays (in order) through the link, then reshape what arrived
int send_remote(mg_patch* p, const char* const* names, const int* nComps, int nArrays) {
  double* box = mg_p2p_pair_outbox(p->remote);
  size_t off = 0;
  const size_t n = (size_t)p->nPatchPoints;
  for (int a = 0; a < nArrays; ++a) {
    double* src = arr(p, names[a]);
    if (!src) MG_FAIL(std::string("block interface '") + p->name + "': '" + names[a] + "' has not been collected");
    MG_CUDA(cudaMemcpyAsync(box + off, src, n * nComps[a] * sizeof(double), cudaMemcpyDeviceToDevice, mg_stream()));
    off += n * nComps[a];
  }
  return mg_p2p_exchange_pair(p->remote, off, 1);
}

int receive_remote(mg_patch* p, const char* const* names, const int* nComps, int nArrays) {
  const double* box = mg_p2p_pair_inbox(p->remote);
  size_t off = 0;
  const size_t n = (size_t)p->nPatchPoints;
  for (int a = 0; a < nArrays; ++a) off += n * nComps[a];
  MG_TRY(mg_p2p_exchange_pair(p->remote, off, 2));
  off = 0;
  for (int a = 0; a < nArrays; ++a) {
    double* dst = nullptr;
    MG_TRY(mg_patch_alloc_array(p, names[a], nComps[a], &dst));
    if (p->nPatchPoints)
      { k_receive<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(p->globalSize[0], p->globalSize[1], p->globalSize[2],
                                                                  p->reorder[0], p->reorder[1], nComps[a], box + off, dst); mg_count_launches(1); }
    off += n * nComps[a];
  }
  MG_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

bool mg_state_has_interfaces(const mg_state* s) {
  for (const mg_patch* p : s->patches)
    if (p->type == MG_PATCH_BLOCK_INTERFACE) return true;
  return false;
}

// readPatchInterfaceInformation (reference src/InterfaceHelperImpl.f90:3-112): a conforms_with b under a's
// index reordering; b gets the inverted reordering (:96-105).
int mg_interface_link(mg_patch* a, mg_patch* b, const int reorderA[3]) {
  if (!a || !b || a->type != MG_PATCH_BLOCK_INTERFACE || b->type != MG_PATCH_BLOCK_INTERFACE)
    MG_FAIL("mg_patch_link_interface: both patches must be SAT_BLOCK_INTERFACE patches");
  const int nD = a->state->nD;
  for (int j = 0; j < 3; ++j) {
    const int o = reorderA ? reorderA[j] : j + 1;
    if (o == 0 || std::abs(o) > 3 || (j < nD && std::abs(o) > nD)) MG_FAIL("mg_patch_link_interface: invalid index reordering");
    if (j == 2 && o != 3) MG_FAIL("mg_patch_link_interface: reordering of the third index is not supported");
    a->reorder[j] = o;
  }
  for (int l = 1; l <= 3; ++l)
    for (int k = 1; k <= 3; ++k)
      if (std::abs(a->reorder[k - 1]) == l) { b->reorder[l - 1] = a->reorder[k - 1] < 0 ? -k : k; break; }
  const bool transposed = std::abs(a->reorder[0]) == 2 && std::abs(a->reorder[1]) == 1;
  for (int d = 0; d < 3; ++d) {
    const int e = transposed && d < 2 ? 1 - d : d;
    if (a->globalSize[d] != b->globalSize[e]) MG_FAIL("mg_patch_link_interface: the two patches do not conform");
  }
  if (a->state->grid->procDims[0] * a->state->grid->procDims[1] * a->state->grid->procDims[2] != 1 ||
      b->state->grid->procDims[0] * b->state->grid->procDims[1] * b->state->grid->procDims[2] != 1)
    MG_FAIL("mg_patch_link_interface: decomposed blocks are not supported (one block per process and device)");
  a->partner = b;
  b->partner = a;
  a->metricsReady = b->metricsReady = false;
  return 0;
}

// The conforming patch lives in another process (another GPU of the box): the two exchange their collected face
// data through a two-party P2P link.  `reorder` is THIS patch's index reordering (the inverse of the partner's);
// the partner's signed penalty amounts and normal direction -- what the METRICS pseudo-exchange carries besides the
// metrics (src/BlockInterfacePatchImpl.f90:667-703) -- are passed by the host, which swaps them when it swaps the
// IPC handles.
int mg_interface_link_remote(mg_patch* p, const int reorder[3], double partnerInviscidAmount,
                             double partnerViscousAmount, int partnerNormalDirection, mg_p2p** link) {
  if (!p || !link || p->type != MG_PATCH_BLOCK_INTERFACE)
    MG_FAIL("mg_patch_link_interface_remote: not a SAT_BLOCK_INTERFACE patch");
  const int nD = p->state->nD;
  for (int j = 0; j < 3; ++j) {
    const int o = reorder ? reorder[j] : j + 1;
    if (o == 0 || std::abs(o) > 3 || (j < nD && std::abs(o) > nD)) MG_FAIL("mg_patch_link_interface_remote: invalid index reordering");
    if (j == 2 && o != 3) MG_FAIL("mg_patch_link_interface_remote: reordering of the third index is not supported");
    p->reorder[j] = o;
  }
  const mg_grid* g = p->state->grid;
  if (g->procDims[0] * g->procDims[1] * g->procDims[2] != 1)
    MG_FAIL("mg_patch_link_interface_remote: a block with interfaces must live in one process");
  p->sigmaIR = partnerInviscidAmount;
  p->sigmaVR = partnerViscousAmount;
  p->normalR = partnerNormalDirection;
  p->partner = nullptr;
  p->metricsReady = false;
  if (!p->remote) MG_TRY(mg_p2p_create_pair((size_t)std::max(1, p->nPatchPoints) * 3 * p->state->nU, &p->remote));
  *link = p->remote;
  return 0;
}

// collectInterfaceData -> exchange (+ reshape) -> disperseInterfaceData for every interface patch of the region
// (reference src/RegionImpl.f90:1927-1958); the METRICS pseudo-mode (src/SolverImpl.f90:571-603) runs once.
int mg_interfaces_exchange(const std::vector<mg_state*>& states, int mode) {
  std::vector<mg_patch*> ifs;
  for (mg_state* s : states)
    for (mg_patch* p : s->patches)
      if (p->type == MG_PATCH_BLOCK_INTERFACE) {
        if (!p->partner && !p->remote)
          MG_FAIL(std::string("block interface '") + p->name + "' has no partner (mg_patch_link_interface)");
        ifs.push_back(p);
      }
  if (ifs.empty()) return 0;
  bool needMetrics = false;
  for (mg_patch* p : ifs) needMetrics = needMetrics || !p->metricsReady;
  if (needMetrics) {
    for (mg_patch* p : ifs) {
      mg_state* s = p->state;
      mg_grid* g = s->grid;
      const int dir = std::abs(p->normalDirection) - 1;
      MG_TRY(collect_to(p, g->metrics.comp(s->nD * dir), g->metrics.compStride, s->nD, "metricsL"));
      p->sigmaIL = p->inviscidPenaltyAmount;
      p->sigmaVL = s->opt.viscosityOn ? p->viscousPenaltyAmount : 0.0;
      p->normalL = p->normalDirection;
    }
    {
      static const char* const sendN[1] = {"metricsL"};
      static const char* const recvN[1] = {"metricsR"};
      for (mg_patch* p : ifs)
        if (p->remote) { const int nc[1] = {p->state->nD}; MG_TRY(send_remote(p, sendN, nc, 1)); }
      for (mg_patch* p : ifs) {
        if (p->remote) {
          const int nc[1] = {p->state->nD};
          MG_TRY(receive_remote(p, recvN, nc, 1));      // sigma / normal of the partner came with the link
        } else {
          MG_TRY(receive(p, "metricsL", "metricsR", p->state->nD));
          p->sigmaIR = p->partner->sigmaIL;
          p->sigmaVR = p->partner->sigmaVL;
          p->normalR = p->partner->normalL;
        }
        p->metricsReady = true;
      }
    }
  }
  for (mg_patch* p : ifs) {
    mg_state* s = p->state;
    const MgField& Q = s->Q[s->cur];
    MG_TRY(collect_to(p, Q.comp(0), Q.compStride, s->nU, "conservedVariablesL"));
    if (mode == MG_FORWARD) {
      if (s->opt.viscosityOn) {
        double* Fc = arr(p, "viscousFluxes");
        if (!Fc) MG_FAIL("block interface: Cartesian viscous fluxes have not been collected");
        double* out = nullptr;
        MG_TRY(mg_patch_alloc_array(p, "viscousFluxesL", s->nU, &out));
        if (p->nPatchPoints)
          { k_normal_viscous_flux<<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(p->nPatchPoints, s->nU, s->nD, Fc,
                                                                                  arr(p, "metricsL"), out); mg_count_launches(1); }
      }
    } else {
      const MgField& W = s->W[s->curW];
      MG_TRY(collect_to(p, W.comp(0), W.compStride, s->nU, "adjointVariablesL"));
      if (mode == MG_LINEARIZED && s->opt.viscosityOn) {
        // computeRhsLinearized collected the linearized viscous fluxes (nU x nD); the interface keeps the ones along
        // its normal direction (reference src/RhsHelperImpl.f90:803-806)
        double* Fc = arr(p, "viscousFluxes");
        if (!Fc) MG_FAIL("block interface: linearized viscous fluxes have not been collected");
        double* out = nullptr;
        MG_TRY(mg_patch_alloc_array(p, "viscousFluxesL", s->nU, &out));
        const int dir = std::abs(p->normalDirection) - 1;
        if (p->nPatchPoints)
          MG_CUDA(cudaMemcpyAsync(out, Fc + (size_t)s->nU * dir * p->nPatchPoints,
                                  (size_t)s->nU * p->nPatchPoints * sizeof(double), cudaMemcpyDeviceToDevice, mg_stream()));
      }
    }
  }
  // what travels per mode (reference collectInterfaceData, src/BlockInterfacePatchImpl.f90:592-700)
  static const char* const sendF[2] = {"conservedVariablesL", "viscousFluxesL"};
  static const char* const recvF[2] = {"conservedVariablesR", "viscousFluxesR"};
  static const char* const sendA[2] = {"conservedVariablesL", "adjointVariablesL"};
  static const char* const recvA[2] = {"conservedVariablesR", "adjointVariablesR"};
  static const char* const sendL[3] = {"conservedVariablesL", "adjointVariablesL", "viscousFluxesL"};
  static const char* const recvL[3] = {"conservedVariablesR", "adjointVariablesR", "viscousFluxesR"};
  for (mg_patch* p : ifs) {
    if (!p->remote) continue;
    mg_state* s = p->state;
    const int nc[3] = {s->nU, s->nU, s->nU};
    if (mode == MG_FORWARD) MG_TRY(send_remote(p, sendF, nc, s->opt.viscosityOn ? 2 : 1));
    else if (mode == MG_LINEARIZED) MG_TRY(send_remote(p, sendL, nc, s->opt.viscosityOn ? 3 : 2));
    else MG_TRY(send_remote(p, sendA, nc, 2));
  }
  for (mg_patch* p : ifs) {
    mg_state* s = p->state;
    if (p->remote) {
      const int nc[3] = {s->nU, s->nU, s->nU};
      if (mode == MG_FORWARD) MG_TRY(receive_remote(p, recvF, nc, s->opt.viscosityOn ? 2 : 1));
      else if (mode == MG_LINEARIZED) MG_TRY(receive_remote(p, recvL, nc, s->opt.viscosityOn ? 3 : 2));
      else MG_TRY(receive_remote(p, recvA, nc, 2));
      continue;
    }
    MG_TRY(receive(p, "conservedVariablesL", "conservedVariablesR", s->nU));
    if (mode == MG_FORWARD) {
      if (s->opt.viscosityOn) MG_TRY(receive(p, "viscousFluxesL", "viscousFluxesR", s->nU));
    } else {
      MG_TRY(receive(p, "adjointVariablesL", "adjointVariablesR", s->nU));
      if (mode == MG_LINEARIZED && s->opt.viscosityOn) MG_TRY(receive(p, "viscousFluxesL", "viscousFluxesR", s->nU));
    }
  }
  MG_CUDA(cudaGetLastError());
  return 0;
}

int mg_interface_apply(mg_state* s, mg_patch* p, int mode) {
  if (p->nPatchPoints == 0) return 0;
  if (mode == MG_ADJOINT && s->opt.useContinuousAdjoint)
    MG_FAIL("block interface: the continuous adjoint is disabled in the reference (src/BlockInterfacePatchImpl.f90:280)");
  mg_grid* g = s->grid;
  IfArgs a;
  std::memset(&a, 0, sizeof(a));
  a.g = geom(p);
  a.QL = arr(p, "conservedVariablesL");
  a.QR = arr(p, "conservedVariablesR");
  if (!a.QL || !a.QR) MG_FAIL("block interface: interface data have not been exchanged");
  a.WL = arr(p, "adjointVariablesL");
  a.WR = arr(p, "adjointVariablesR");
  a.FvL = arr(p, "viscousFluxesL");
  a.FvR = arr(p, "viscousFluxesR");
  a.mL = arr(p, "metricsL");
  a.mR = arr(p, "metricsR");
  if (mode != MG_FORWARD && (!a.WL || !a.WR)) MG_FAIL("block interface: adjoint interface data have not been exchanged");
  if (mode != MG_ADJOINT && s->opt.viscosityOn && (!a.FvL || !a.FvR))
    MG_FAIL("block interface: viscous interface fluxes have not been exchanged");
  a.m = g->metrics.comp(0);
  a.jac = g->jacobian.comp(0);
  a.v = s->specificVolume.comp(0);
  a.u = s->velocity.comp(0);
  a.T = s->temperature.comp(0);
  a.tau = s->opt.viscosityOn ? s->stressTensor.comp(0) : nullptr;
  a.q = s->opt.viscosityOn ? s->heatFlux.comp(0) : nullptr;
  a.iblank = g->iblank;
  a.rhs = s->rhs.comp(0);
  a.cs = s->rhs.compStride;
  a.dir = std::abs(p->normalDirection) - 1;
  a.mode = mode;
  a.viscous = s->opt.viscosityOn;
  a.normal = p->normalDirection;
  a.normalL = p->normalL;
  a.normalR = p->normalR;
  a.sigmaI = p->inviscidPenaltyAmount;
  a.sigmaV = p->viscousPenaltyAmount;
  a.sigmaIL = p->sigmaIL; a.sigmaIR = p->sigmaIR; a.sigmaVL = p->sigmaVL; a.sigmaVR = p->sigmaVR;
  a.gamma = s->opt.ratioOfSpecificHeats;
  a.powerLaw = s->opt.powerLawExponent;
  MG_TRY(dispatch_nd(s->nD, [&](auto nd) {
    { k_interface<decltype(nd)::value><<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(a); mg_count_launches(1); }
    return 0;
  }));
  MG_CUDA(cudaGetLastError());
  return 0;
}

int mg_interfaces_adjoint_sources(mg_state* s, MgField* temp1) {
  mg_grid* g = s->grid;
  for (mg_patch* p : s->patches) {
    if (p->type != MG_PATCH_BLOCK_INTERFACE || p->nPatchPoints == 0) continue;
    IfSrcArgs a;
    a.g = geom(p);
    a.WL = arr(p, "adjointVariablesL");
    a.WR = arr(p, "adjointVariablesR");
    a.mL = arr(p, "metricsL");
    a.mR = arr(p, "metricsR");
    if (!a.WL || !a.WR || !a.mL || !a.mR) MG_FAIL("block interface: adjoint interface data have not been exchanged");
    a.u = s->velocity.comp(0);
    a.mu = s->mu.comp(0);
    a.lam = s->lambda.comp(0);
    a.kap = s->kappa.comp(0);
    a.m = g->metrics.comp(0);
    a.jac = g->jacobian.comp(0);
    a.iblank = g->iblank;
    a.temp1 = temp1->comp(0);
    a.cs = temp1->compStride;
    a.sigmaVL = p->sigmaVL;
    a.sigmaVR = p->sigmaVR;
    MG_TRY(dispatch_nd(s->nD, [&](auto nd) {
      { k_interface_adjoint_source<decltype(nd)::value><<<nblocks(p->nPatchPoints), 128, 0, mg_stream()>>>(a); mg_count_launches(1); }
      return 0;
    }));
  }
  MG_CUDA(cudaGetLastError());
  return 0;
}
