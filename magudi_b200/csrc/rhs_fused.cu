// Fused hot-path sweeps for the forward right-hand side (the path BASELINE.json:north_star names).
//
// Per RK stage the reference performs ~15 separate operator applications plus pointwise passes
// (src/RhsHelperImpl.f90:254-354, src/StateImpl.f90:466-537, src/RK4IntegratorImpl.f90:65-162).
// Here a stage is TWO streaming sweeps over the grid:
//
//   sweep A  (t_State%update + the state-only part of addDissipation)
//            reads  Q (nU) + metrics/Jacobian/arc lengths
//            writes tau (nD(nD+1)/2 unique entries), q (nD), dissipation term (nU)
//   sweep B  (computeRhsForward + x 1/J + RK4 substep)
//            reads  Q, tau, q, dissipation term, metrics/Jacobian, RK buffers
//            writes RK accumulator and the next Q   (or the RHS when patches follow)
//
// Both are 2.5-D streaming kernels: a CTA owns a 16x16 (i,j) tile and marches along k.  In-plane
// neighbours come from a shared-memory tile (with halo, periodic wrap or SBP boundary closures),
// k-neighbours from a per-thread queue (registers in A, shared memory in B) so every field is read
// from HBM once per sweep.  HBM-bound fp64 stencil/pointwise work: no tensor cores.
#include <cstring>

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

struct FusedArgs {
  int nx, ny, nz;
  long plane;
  size_t cs;                    // component stride of every field
  int wrapK;                    // 1: single rank, periodic in k -> wrap plane index; 0: read ghost planes
  int kBeg, kEnd, kChunk;
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

// ------------------------------------------------------------------------------- sweep A
// Shared tile: NF fields on a (TY+2R) x (TX+2R) box (corners unused).
template <int ND, int R, bool COMPOSITE, int DLO, int DN, int TLO, int TN>
__global__ void __launch_bounds__(NT, 1) k_sweepA(FusedArgs a) {
  constexpr int NU = ND + 2;
  constexpr int NTAU = ND * (ND + 1) / 2;
  constexpr int W = TX + 2 * R, H = TY + 2 * R;
  constexpr int NF = NU + ND + 1 + 2;          // Q, u, T, arc_i, arc_j
  constexpr int RK = (ND == 3) ? R : 0;        // k half-width
  extern __shared__ double smem[];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  int i0, j0;
  bool lastI, lastJ;
  tile_origin(blockIdx.x, a.nx, TX, i0, lastI);
  tile_origin(blockIdx.y, a.ny, TY, j0, lastJ);
  const int i = i0 + tx, j = j0 + ty;
  const bool mine = owns(i, a.nx, TX, lastI) && owns(j, a.ny, TY, lastJ);
  const bool inside = i < a.nx && j < a.ny;
  const long pij = (long)i + (long)a.nx * j;
  const double gamma = a.pp.gamma;

  // halo point handled by this thread (at most one): hk 0 none, 1 i-halo, 2 j-halo
  int hcol = 0, hrow = 0, hk = 0;
  long hp = -1;
  {
    const int h = threadIdx.x;
    if (h < 2 * R * TY) {
      const int ii = h % (2 * R), row = h / (2 * R);
      const int lc = ii < R ? ii : TX + ii;               // local column in the box
      const int gi = wrap_index(i0 - R + lc, a.dir[0]);
      const int gj = j0 + row;
      hcol = lc; hrow = row + R;
      if (gi >= 0 && gj < a.ny) { hk = 1; hp = (long)gi + (long)a.nx * gj; }
    } else if (h < 2 * R * TY + 2 * R * TX) {
      const int h2 = h - 2 * R * TY;
      const int col = h2 % TX, jj = h2 / TX;
      const int lr = jj < R ? jj : TY + jj;
      const int gj = wrap_index(j0 - R + lr, a.dir[1]);
      const int gi = i0 + col;
      hcol = col + R; hrow = lr;
      if (gj >= 0 && gi < a.nx) { hk = 2; hp = (long)gi + (long)a.nx * gj; }
    }
  }

  const int kc0 = a.kBeg + blockIdx.z * a.kChunk;
  const int kc1 = min(kc0 + a.kChunk, a.kEnd);
  auto planeOf = [&](int k) -> long {
    if (ND < 3) return 0;
    int kk = k;
    if (a.wrapK) kk = (k % a.nz + a.nz) % a.nz;
    return (long)kk * a.plane;
  };

  double qq[2 * RK + 1][NU];       // Q at planes p-RK .. p+RK once the queue is primed
  for (int s = kc0 - RK; s < kc1 + RK; ++s) {
    // ---- arrival of plane s
#pragma unroll
    for (int q = 0; q < 2 * RK; ++q)
#pragma unroll
      for (int c = 0; c < NU; ++c) qq[q][c] = qq[q + 1][c];
    if (inside) {
      const long off = planeOf(s) + pij;
#pragma unroll
      for (int c = 0; c < NU; ++c) qq[2 * RK][c] = a.Q[(size_t)c * a.cs + off];
    }
    const int p = s - RK;
    if (p < kc0) continue;
    // ---- output plane p: build the in-plane tile
    double* S = smem + (size_t)((p - kc0) & 1) * NF * H * W;
    auto at = [&](int f, int row, int col) -> double& { return S[((size_t)f * H + row) * W + col]; };
    const long poff = planeOf(p);
    Prim<ND> sc;
    if (inside) {
      dependent<ND>(qq[RK], gamma, sc);
#pragma unroll
      for (int c = 0; c < NU; ++c) at(c, ty + R, tx + R) = qq[RK][c];
#pragma unroll
      for (int d = 0; d < ND; ++d) at(NU + d, ty + R, tx + R) = sc.u[d];
      at(NU + ND, ty + R, tx + R) = sc.T;
      if (!COMPOSITE) {
        at(NU + ND + 1, ty + R, tx + R) = a.arc[(size_t)0 * a.cs + poff + pij];
        at(NU + ND + 2, ty + R, tx + R) = a.arc[(size_t)1 * a.cs + poff + pij];
      }
    }
    if (hk) {
      double Qh[NU];
      const long off = poff + hp;
#pragma unroll
      for (int c = 0; c < NU; ++c) { Qh[c] = a.Q[(size_t)c * a.cs + off]; at(c, hrow, hcol) = Qh[c]; }
      Prim<ND> sh;
      dependent<ND>(Qh, gamma, sh);
#pragma unroll
      for (int d = 0; d < ND; ++d) at(NU + d, hrow, hcol) = sh.u[d];
      at(NU + ND, hrow, hcol) = sh.T;
      if (!COMPOSITE) at(NU + ND + hk, hrow, hcol) = a.arc[(size_t)(hk - 1) * a.cs + off];
    }
    __syncthreads();
    if (!mine) continue;       // NB: no barrier after this point inside the iteration (double-buffered tile)

    // ---- derivatives of (u, T) along xi, eta, zeta
    double dxi[ND][ND + 1];    // dxi[direction][field]: fields u_0..u_{ND-1}, T
#pragma unroll
    for (int f = 0; f < ND + 1; ++f) {
      dxi[0][f] = line_apply(a.D[0], i, a.nx, [&](int cc) { return at(NU + f, ty + R, cc - i0 + R); });
      dxi[1][f] = line_apply(a.D[1], j, a.ny, [&](int cc) { return at(NU + f, cc - j0 + R, tx + R); });
    }
    if constexpr (ND == 3) {
      double acc[ND + 1];
#pragma unroll
      for (int f = 0; f < ND + 1; ++f) acc[f] = 0.0;
#pragma unroll
      for (int q = 1; q <= RK; ++q) {
        Prim<ND> sp, sm;
        dependent<ND>(qq[RK + q], gamma, sp);
        dependent<ND>(qq[RK - q], gamma, sm);
        const double cq = a.D[2].c[q - a.D[2].lo];
#pragma unroll
        for (int d = 0; d < ND; ++d) acc[d] += cq * (sp.u[d] - sm.u[d]);
        acc[ND] += cq * (sp.T - sm.T);
      }
#pragma unroll
      for (int f = 0; f < ND + 1; ++f) dxi[ND - 1][f] = acc[f];
    }
    // ---- gradient in physical space (reference src/GridImpl.f90:1357-1413), stress tensor, heat flux
    const long off = poff + pij;
    const double jac = a.jac[off];
    double M[ND * ND];
    if (a.curvilinear) {
#pragma unroll
      for (int c = 0; c < ND * ND; ++c) M[c] = a.m[(size_t)c * a.cs + off];
    } else {
#pragma unroll
      for (int c = 0; c < ND * ND; ++c) M[c] = 0.0;
#pragma unroll
      for (int d = 0; d < ND; ++d) M[d + ND * d] = a.m[(size_t)(d + ND * d) * a.cs + off];
    }
    if (a.viscous) {
      double g[ND * ND], gT[ND];
#pragma unroll
      for (int c = 0; c < ND; ++c)
#pragma unroll
        for (int jx = 0; jx < ND; ++jx) {
          double r;
          if (a.curvilinear) {
            r = M[jx] * dxi[0][c];
#pragma unroll
            for (int d = 1; d < ND; ++d) r += M[jx + ND * d] * dxi[d][c];
            r = jac * r;
          } else {
            r = jac * M[jx + ND * jx] * dxi[jx][c];
          }
          g[jx + ND * c] = r;
        }
#pragma unroll
      for (int jx = 0; jx < ND; ++jx) {
        double r;
        if (a.curvilinear) {
          r = M[jx] * dxi[0][ND];
#pragma unroll
          for (int d = 1; d < ND; ++d) r += M[jx + ND * d] * dxi[d][ND];
          r = jac * r;
        } else {
          r = jac * M[jx + ND * jx] * dxi[jx][ND];
        }
        gT[jx] = r;
      }
      double mu, lam, kap, tau[ND * ND];
      transport(sc.T, a.pp, mu, lam, kap);
      stress_from_gradient<ND>(g, mu, lam, tau);
      // unique entries: (0,0),(0,1)[,(0,2)],(1,1)[,(1,2),(2,2)] in row-major upper order
      int t = 0;
#pragma unroll
      for (int r0 = 0; r0 < ND; ++r0)
#pragma unroll
        for (int c0 = r0; c0 < ND; ++c0) a.tauq[(size_t)(t++) * a.cs + off] = tau[c0 + ND * r0];
#pragma unroll
      for (int d = 0; d < ND; ++d) a.tauq[(size_t)(NTAU + d) * a.cs + off] = -kap * gT[d];
    }
    // ---- dissipation term  sum_dir Diss_dir(Q)   (reference src/RhsHelperImpl.f90:58-81)
    if (a.diss) {
      double dz[NU];
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        double r;
        if (COMPOSITE) {
          r = line_apply(a.Dd[0], i, a.nx, [&](int cc) { return at(c, ty + R, cc - i0 + R); });
          r += line_apply(a.Dd[1], j, a.ny, [&](int cc) { return at(c, cc - j0 + R, tx + R); });
        } else {
          r = line_dissipation(a.Dd[0], a.Dt[0], a.dir[0], i,
                               [&](int cc) { return at(c, ty + R, cc - i0 + R); },
                               [&](int cc) { return at(NU + ND + 1, ty + R, cc - i0 + R); });
          r += line_dissipation(a.Dd[1], a.Dt[1], a.dir[1], j,
                                [&](int cc) { return at(c, cc - j0 + R, tx + R); },
                                [&](int cc) { return at(NU + ND + 2, cc - j0 + R, tx + R); });
        }
        dz[c] = r;
      }
      if constexpr (ND == 3) {
        if (COMPOSITE) {
#pragma unroll
          for (int c = 0; c < NU; ++c) {
            double r = a.Dd[2].c[0 - a.Dd[2].lo] * qq[RK][c];
#pragma unroll
            for (int q = 1; q <= RK; ++q) r += a.Dd[2].c[q - a.Dd[2].lo] * (qq[RK + q][c] + qq[RK - q][c]);
            dz[c] += r;
          }
        } else {
          double tz[TN][NU];
#pragma unroll
          for (int e = 0; e < TN; ++e) {
            const int ko = TLO + e;         // plane offset of this t sample
            const double arc = a.arc[(size_t)2 * a.cs + planeOf(p + ko) + pij];
#pragma unroll
            for (int c = 0; c < NU; ++c) {
              double r = 0.0;
#pragma unroll
              for (int b = 0; b < DN; ++b) r += a.Dd[2].c[b] * qq[RK + ko + DLO + b][c];
              tz[e][c] = -arc * r;
            }
          }
#pragma unroll
          for (int c = 0; c < NU; ++c) {
            double r = 0.0;
#pragma unroll
            for (int e = 0; e < TN; ++e) r += a.Dt[2].c[e] * tz[e][c];
            dz[c] += r;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NU; ++c) a.diss[(size_t)c * a.cs + off] = dz[c];
    }
  }
}

// ------------------------------------------------------------------------------- sweep B
template <int ND, int R>
__global__ void __launch_bounds__(NT, 2) k_sweepB(FusedArgs a) {
  constexpr int NU = ND + 2;
  constexpr int NTAU = ND * (ND + 1) / 2;
  constexpr int W = TX + 2 * R, H = TY + 2 * R;
  constexpr int RK = (ND == 3) ? R : 0;
  constexpr int NQ = 2 * RK + 1;
  extern __shared__ double smem[];
  double* F1 = smem;                               // [NU][TY][W]   contravariant flux along xi
  double* F2 = F1 + (size_t)NU * TY * W;           // [NU][H][TX]   contravariant flux along eta
  double* F3 = F2 + (size_t)NU * H * TX;           // [NQ][NU][NT]  k-queue of the flux along zeta
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  int i0, j0;
  bool lastI, lastJ;
  tile_origin(blockIdx.x, a.nx, TX, i0, lastI);
  tile_origin(blockIdx.y, a.ny, TY, j0, lastJ);
  const int i = i0 + tx, j = j0 + ty;
  const bool mine = owns(i, a.nx, TX, lastI) && owns(j, a.ny, TY, lastJ);
  const bool inside = i < a.nx && j < a.ny;
  const long pij = (long)i + (long)a.nx * j;
  const double gamma = a.pp.gamma;

  int hcol = 0, hrow = 0, hk = 0;
  long hp = -1;
  {
    const int h = threadIdx.x;
    if (h < 2 * R * TY) {
      const int ii = h % (2 * R), row = h / (2 * R);
      const int lc = ii < R ? ii : TX + ii;
      const int gi = wrap_index(i0 - R + lc, a.dir[0]);
      const int gj = j0 + row;
      hcol = lc; hrow = row;
      if (gi >= 0 && gj < a.ny) { hk = 1; hp = (long)gi + (long)a.nx * gj; }
    } else if (h < 2 * R * TY + 2 * R * TX) {
      const int h2 = h - 2 * R * TY;
      const int col = h2 % TX, jj = h2 / TX;
      const int lr = jj < R ? jj : TY + jj;
      const int gj = wrap_index(j0 - R + lr, a.dir[1]);
      const int gi = i0 + col;
      hcol = col; hrow = lr;
      if (gj >= 0 && gi < a.nx) { hk = 2; hp = (long)gi + (long)a.nx * gj; }
    }
  }
  const int kc0 = a.kBeg + blockIdx.z * a.kChunk;
  const int kc1 = min(kc0 + a.kChunk, a.kEnd);
  auto planeOf = [&](int k) -> long {
    if (ND < 3) return 0;
    int kk = k;
    if (a.wrapK) kk = (k % a.nz + a.nz) % a.nz;
    return (long)kk * a.plane;
  };

  // contravariant flux along direction d at a point (reference CNSHelperImpl.f90:563-689, :772-840)
  auto flux_dir = [&](int d, long off, double* Fh) {
    double Q[NU], tau[ND * ND], q[ND], Fc[NU], Fv[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) Q[c] = a.Q[(size_t)c * a.cs + off];
    Prim<ND> s;
    dependent<ND>(Q, gamma, s);
    if (a.viscous) {
      int t = 0;
#pragma unroll
      for (int r0 = 0; r0 < ND; ++r0)
#pragma unroll
        for (int c0 = r0; c0 < ND; ++c0) {
          const double v = a.tauqIn[(size_t)(t++) * a.cs + off];
          tau[c0 + ND * r0] = v;
          tau[r0 + ND * c0] = v;
        }
#pragma unroll
      for (int e = 0; e < ND; ++e) q[e] = a.tauqIn[(size_t)(NTAU + e) * a.cs + off];
    }
    if (a.curvilinear) {
#pragma unroll
      for (int c = 0; c < NU; ++c) Fh[c] = 0.0;
#pragma unroll
      for (int l = 0; l < ND; ++l) {
        cartesian_flux<ND>(l, Q, s, a.viscous, tau, q, Fc, Fv);
        const double ml = a.m[(size_t)(l + ND * d) * a.cs + off];
#pragma unroll
        for (int c = 0; c < NU; ++c) Fh[c] = (l == 0) ? ml * Fc[c] : Fh[c] + ml * Fc[c];
      }
    } else {
      cartesian_flux<ND>(d, Q, s, a.viscous, tau, q, Fc, Fv);
      const double md = a.m[(size_t)(d + ND * d) * a.cs + off];
#pragma unroll
      for (int c = 0; c < NU; ++c) Fh[c] = md * Fc[c];
    }
  };

  double rxy[RK + 1][NU];          // in-plane part of -div(F) for planes s-RK .. s
  for (int s = kc0 - RK; s < kc1 + RK; ++s) {
    const long soff = planeOf(s);
    const bool planeActive = s >= kc0 && s < kc1;
    // ---- arrival of plane s: fluxes at the own point
    if (inside) {
      double Fh[NU];
      if constexpr (ND == 3) {
        flux_dir(2, soff + pij, Fh);
        const int slot = ((s % NQ) + NQ) % NQ;
#pragma unroll
        for (int c = 0; c < NU; ++c) F3[((size_t)slot * NU + c) * NT + threadIdx.x] = Fh[c];
      }
      if (planeActive) {
        flux_dir(0, soff + pij, Fh);
#pragma unroll
        for (int c = 0; c < NU; ++c) F1[((size_t)c * TY + ty) * W + tx + R] = Fh[c];
        flux_dir(1, soff + pij, Fh);
#pragma unroll
        for (int c = 0; c < NU; ++c) F2[((size_t)c * H + ty + R) * TX + tx] = Fh[c];
      }
    }
    if (planeActive && hk) {
      double Fh[NU];
      flux_dir(hk - 1, soff + hp, Fh);
      if (hk == 1) {
#pragma unroll
        for (int c = 0; c < NU; ++c) F1[((size_t)c * TY + hrow) * W + hcol] = Fh[c];
      } else {
#pragma unroll
        for (int c = 0; c < NU; ++c) F2[((size_t)c * H + hrow) * TX + hcol] = Fh[c];
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RK; ++q)
#pragma unroll
      for (int c = 0; c < NU; ++c) rxy[q][c] = rxy[q + 1][c];
    if (planeActive && mine) {
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        double r = line_apply(a.D[0], i, a.nx, [&](int cc) { return F1[((size_t)c * TY + ty) * W + cc - i0 + R]; });
        r += line_apply(a.D[1], j, a.ny, [&](int cc) { return F2[((size_t)c * H + cc - j0 + R) * TX + tx]; });
        rxy[RK][c] = r;
      }
    }
    // ---- output plane p
    const int p = s - RK;
    if (p >= kc0 && mine) {
      const long off = planeOf(p) + pij;
      double r[NU];
#pragma unroll
      for (int c = 0; c < NU; ++c) r[c] = rxy[0][c];
      if constexpr (ND == 3) {
#pragma unroll
        for (int q = 1; q <= RK; ++q) {
          const int sp = (((p + q) % NQ) + NQ) % NQ, sm = (((p - q) % NQ) + NQ) % NQ;
          const double cq = a.D[2].c[q - a.D[2].lo];
#pragma unroll
          for (int c = 0; c < NU; ++c)
            r[c] += cq * (F3[((size_t)sp * NU + c) * NT + threadIdx.x] - F3[((size_t)sm * NU + c) * NT + threadIdx.x]);
        }
      }
      const double jac = a.jac[off];
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        double rhs = 0.0 - r[c];
        if (a.dissIn) rhs += a.dissAmount * a.dissIn[(size_t)c * a.cs + off];
        r[c] = rhs * jac;
      }
      if (!a.fuseRk) {
#pragma unroll
        for (int c = 0; c < NU; ++c) a.rhs[(size_t)c * a.cs + off] = r[c];
      } else {
        // RK4 substep (reference src/RK4IntegratorImpl.f90:106-158) fused into the last RHS kernel
#pragma unroll
        for (int c = 0; c < NU; ++c) {
          const size_t qi = (size_t)c * a.cs + off;
          if (a.stage == 1) {
            const double Q0 = a.Q[qi];                  // buffer1 is the input Q buffer itself
            a.b2[qi] = Q0 + a.dt * r[c] / 6.0;
            a.Qout[qi] = Q0 + a.dt * r[c] / 2.0;
          } else if (a.stage == 2) {
            a.b2[qi] = a.b2[qi] + a.dt * r[c] / 3.0;
            a.Qout[qi] = a.b1in[qi] + a.dt * r[c] / 2.0;
          } else if (a.stage == 3) {
            a.b2[qi] = a.b2[qi] + a.dt * r[c] / 3.0;
            a.Qout[qi] = a.b1in[qi] + a.dt * r[c];
          } else {
            a.Qout[qi] = a.b2[qi] + a.dt * r[c] / 6.0;
          }
        }
      }
    }
    __syncthreads();
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
  a->kBeg = 0;
  a->kEnd = g->localSize[2];
  a->kChunk = g->localSize[2];
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
  a->pp = s->phys();
  a->dissAmount = s->opt.dissipationAmount;
  a->m = g->metrics.comp(0);
  a->jac = g->jacobian.comp(0);
  a->arc = g->arcLengths.comp(0);
  return 0;
}

template <int ND, int R, bool COMP, int DLO, int DN, int TLO, int TN>
int launchA(const FusedArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int NF = (ND + 2) + ND + 1 + 2;
  const size_t smem = 2 * sizeof(double) * NF * (TY + 2 * R) * (TX + 2 * R);
  auto kern = k_sweepA<ND, R, COMP, DLO, DN, TLO, TN>;
  static bool configured = false;
  if (!configured) {
    MG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  mg_profile_begin("sweepA");
  kern<<<grid, NT, smem, st>>>(a);
  mg_profile_end();
  MG_CUDA(cudaGetLastError());
  mg_count_launches(1);
  return 0;
}

template <int ND, int R>
int launchB(const FusedArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int NU = ND + 2;
  constexpr int NQ = (ND == 3) ? 2 * R + 1 : 1;
  const size_t smem = sizeof(double) * ((size_t)NU * TY * (TX + 2 * R) + (size_t)NU * (TY + 2 * R) * TX +
                                        (size_t)NQ * NU * NT);
  auto kern = k_sweepB<ND, R>;
  static bool configured = false;
  if (!configured) {
    MG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  mg_profile_begin("sweepB");
  kern<<<grid, NT, smem, st>>>(a);
  mg_profile_end();
  MG_CUDA(cudaGetLastError());
  mg_count_launches(1);
  return 0;
}

dim3 tiles(const FusedArgs& a, int nChunks) {
  return dim3((a.nx + TX - 1) / TX, (a.ny + TY - 1) / TY, nChunks);
}

}  // namespace

int mg_fused_supported(const mg_state* s, int mode) {
  const mg_grid* g = s->grid;
  if (mode != MG_FORWARD) return 0;
  if (g->nD < 2) return 0;
  if (g->iblank) return 0;
  if (!s->patches.empty() || !s->acousticSources.empty()) return 0;
  if (g->nD == 3 && g->periodicityType[2] != MG_PERIODIC_PLANE) return 0;
  SchemeInfo si;
  if (!scheme_of(g, &si)) return 0;
  if (g->nD == 3 && g->localSize[2] < si.R) return 0;
  for (int d = 0; d < 2; ++d)
    if (g->periodicityType[d] == MG_PERIODIC_NONE) {
      // a closure block (and its adjoint-free forward operators) must fit in one 16-wide tile + halo
      const MgDevOp& o = g->firstDerivative[d]->op;
      if (o.boundaryWidth > TX + si.R || o.boundaryDepth > TX || g->localSize[d] < 2 * o.boundaryDepth) return 0;
    } else if (g->localSize[d] < si.R + 1) {
      return 0;
    }
  return 1;
}

int mg_fused_alloc(mg_state* s) {
  mg_grid* g = s->grid;
  const int nD = s->nD;
  const int nTauQ = nD * (nD + 1) / 2 + nD;
  if (s->opt.viscosityOn && s->tauq.nComp != nTauQ) MG_TRY(mg_field_alloc(g, nTauQ, &s->tauq));
  if (s->opt.dissipationOn && s->dissTerm.nComp != s->nU) MG_TRY(mg_field_alloc(g, s->nU, &s->dissTerm));
  return 0;
}

// Sweep A: state update + dissipation term
int mg_fused_sweepA(mg_state* s) {
  mg_grid* g = s->grid;
  MG_TRY(mg_fused_alloc(s));
  FusedArgs a;
  MG_TRY(fill_args(s, &a));
  a.Q = s->Q[s->cur].comp(0);
  a.tauq = s->opt.viscosityOn ? s->tauq.comp(0) : nullptr;
  a.diss = s->opt.dissipationOn ? s->dissTerm.comp(0) : nullptr;
  if (!a.viscous && !a.diss) { s->fusedValid = true; return 0; }
  SchemeInfo si;
  scheme_of(g, &si);
  const dim3 grid = tiles(a, 1);
  cudaStream_t st = mg_stream();
  const bool comp = g->compositeDissipation || !g->dissipationOn;
  int rc = -1;
#define MG_A(ND_, R_, DLO, DN, TLO, TN)                                                 \
  if (s->nD == ND_ && si.R == R_)                                                       \
    rc = comp ? launchA<ND_, R_, true, DLO, DN, TLO, TN>(a, grid, st)                   \
              : launchA<ND_, R_, false, DLO, DN, TLO, TN>(a, grid, st);
  MG_A(2, 2, -1, 3, -1, 3)
  MG_A(2, 3, -2, 4, -1, 4)
  MG_A(2, 4, -2, 5, -2, 5)
  MG_A(3, 2, -1, 3, -1, 3)
  MG_A(3, 3, -2, 4, -1, 4)
  MG_A(3, 4, -2, 5, -2, 5)
#undef MG_A
  if (rc != 0) return rc < 0 && rc != -2 ? (mg_set_error("fused sweep A: unsupported configuration"), -1) : rc;
  s->fusedValid = true;
  return 0;
}

// Sweep B: RHS (+ RK4 substep when fuseRk)
int mg_fused_sweepB(mg_state* s, int fuseRk, int stage, double dt) {
  mg_grid* g = s->grid;
  if (!s->fusedValid) MG_FAIL("fused sweep B: state has not been updated (sweep A)");
  FusedArgs a;
  MG_TRY(fill_args(s, &a));
  a.Q = s->Q[s->cur].comp(0);
  a.tauqIn = s->opt.viscosityOn ? s->tauq.comp(0) : nullptr;
  a.dissIn = s->opt.dissipationOn ? s->dissTerm.comp(0) : nullptr;
  a.rhs = s->rhs.comp(0);
  a.fuseRk = fuseRk;
  a.stage = stage;
  a.dt = dt;
  if (fuseRk) {
    // stage 1: buffer1 := Q (pointer swap, no copy): the current Q buffer becomes buffer1 and the
    // freed buffer1 storage receives the new Q.
    if (stage == 1) {
      a.b1in = a.Q;
      a.Qout = s->rk1.comp(0);
    } else {
      a.b1in = s->rk1.comp(0);
      a.Qout = s->Q[1 - s->cur].comp(0);
    }
    a.b2 = s->rk2.comp(0);
  }
  SchemeInfo si;
  scheme_of(g, &si);
  const dim3 grid = tiles(a, 1);
  cudaStream_t st = mg_stream();
  int rc = -1;
  if (s->nD == 2 && si.R == 2) rc = launchB<2, 2>(a, grid, st);
  if (s->nD == 2 && si.R == 3) rc = launchB<2, 3>(a, grid, st);
  if (s->nD == 2 && si.R == 4) rc = launchB<2, 4>(a, grid, st);
  if (s->nD == 3 && si.R == 2) rc = launchB<3, 2>(a, grid, st);
  if (s->nD == 3 && si.R == 3) rc = launchB<3, 3>(a, grid, st);
  if (s->nD == 3 && si.R == 4) rc = launchB<3, 4>(a, grid, st);
  if (rc != 0) return rc;
  if (fuseRk) {
    if (stage == 1) {
      // swap roles: rk1 <-> Q[cur]
      MgField tmp = s->rk1;
      s->rk1 = s->Q[s->cur];
      s->Q[s->cur] = tmp;
    } else {
      s->cur = 1 - s->cur;
    }
    s->fusedValid = false;
    s->dependentValid = false;
  }
  return 0;
}
