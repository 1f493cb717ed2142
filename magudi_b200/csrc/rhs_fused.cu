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
#include "fused_common.cuh"

#ifndef MG_SWEEPA_PIPE_DEFAULT
#define MG_SWEEPA_PIPE_DEFAULT 0
#endif

namespace {

// ------------------------------------------------------------------------------- sweep A
// t_State%update: (u, T) -> gradient -> stress tensor + heat flux.  Shared memory: an in-plane tile of
// (u, T) on a (TY+2R) x (TX+2R) box (corners unused) for the output plane, plus the k-queue of (u, T) for
// the 2R+1 planes in flight (thread-private columns).  Nothing of the queue lives in registers, so three
// CTAs fit per SM.
// PIPE: the loads of the next arriving plane (own point and halo point) are issued right after the barrier that
// publishes the tile, so that they are in flight during the stencil arithmetic of the current output plane instead
// of stalling the first instruction of the next iteration (MG_SWEEPA_PIPE; same arithmetic, same results).
template <int ND, int R, bool CURV, bool CLOS, bool PIPE = false>
__global__ void __launch_bounds__(NT, 3) k_sweepA(FusedArgs a) {
  constexpr int NU = ND + 2;
  constexpr int NTAU = ND * (ND + 1) / 2;
  constexpr int W = TX + 2 * R, H = TY + 2 * R;
  constexpr int NP = ND + 1;                   // u_0..u_{ND-1}, T
  constexpr int RK = (ND == 3) ? R : 0;        // k half-width
  constexpr int NQ = 2 * RK + 1;
  extern __shared__ double smem[];
  double* const T0 = smem;                                   // [NP][H][W]
  double* const KQ = smem + (size_t)NP * H * W;              // [NQ][NP][NT]
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
  const bool fastI = !CLOS || !((a.D[0].hasB0 && i0 < a.D[0].depth) || (a.D[0].hasB1 && i0 + TX > a.nx - a.D[0].depth));
  const bool fastJ = !CLOS || !((a.D[1].hasB0 && j0 < a.D[1].depth) || (a.D[1].hasB1 && j0 + TY > a.ny - a.D[1].depth));
  double* const tc = T0 + (ty + R) * W + tx + R;             // own point; field stride H*W, row stride W
  double* const kqc = KQ + threadIdx.x;                      // slot stride NP*NT, field stride NT

  // halo point handled by this thread (at most one): hk 0 none, 1 i-halo, 2 j-halo
  int hoff = 0, hk = 0;
  long hp = -1;
  {
    const int h = threadIdx.x;
    if (h < 2 * R * TY) {
      const int ii = h % (2 * R), row = h / (2 * R);
      const int lc = ii < R ? ii : TX + ii;               // local column in the box
      const int gi = wrap_index(i0 - R + lc, a.dir[0]);
      const int gj = j0 + row;
      hoff = (row + R) * W + lc;
      if (gi >= 0 && gj < a.ny) { hk = 1; hp = (long)gi + (long)a.nx * gj; }
    } else if (h < 2 * R * TY + 2 * R * TX) {
      const int h2 = h - 2 * R * TY;
      const int col = h2 % TX, jj = h2 / TX;
      const int lr = jj < R ? jj : TY + jj;
      const int gj = wrap_index(j0 - R + lr, a.dir[1]);
      const int gi = i0 + col;
      hoff = lr * W + col + R;
      if (gj >= 0 && gi < a.nx) { hk = 2; hp = (long)gi + (long)a.nx * gj; }
    }
  }

  const int kc0 = a.kBeg + (a.zOff + (int)blockIdx.z * a.zMul) * a.kChunk;
  const int kc1 = min(kc0 + a.kChunk, a.kEnd);
  auto wrapPlane = [&](int k) -> int {
    if (ND < 3 || !a.wrapK) return k;
    int kk = k % a.nz;
    return kk < 0 ? kk + a.nz : kk;
  };
  int ks = wrapPlane(kc0 - RK);
  int slot = 0;                    // queue slot of the arriving plane s

  PfItems<1> pfi;
  pfi.init(a, (long)i0 + (long)a.nx * j, tx, i0 < a.nx && j < a.ny);
  double Qs[NU], Qh[NU];
  // loads of the arriving plane s_ (storage plane ks_) and of the halo point of plane s_ - RK
  auto issueLoads = [&](int s_, int ks_) {
    if (inside) {
      const double* __restrict__ Qp = a.Q + ((ND == 3) ? (long)ks_ * a.plane : 0) + pij;
#pragma unroll
      for (int c = 0; c < NU; ++c) Qs[c] = __ldg(Qp + (size_t)c * a.cs);
    }
    if (hk && s_ - RK >= kc0) {
      int kp_ = ks_ - RK;
      if (ND == 3 && a.wrapK && kp_ < 0) kp_ += a.nz;
      const double* __restrict__ Qp = a.Q + ((ND == 3) ? (long)kp_ * a.plane : 0) + hp;
#pragma unroll
      for (int c = 0; c < NU; ++c) Qh[c] = __ldg(Qp + (size_t)c * a.cs);
    }
  };
  if (PIPE) issueLoads(kc0 - RK, ks);
  for (int s = kc0 - RK; s < kc1 + RK; ++s) {
    // ---- arrival of plane s
    if (ND == 3 && a.prefetch) {
      int kf = ks + a.prefetch, kq = ks - RK + a.prefetch;
      if (a.wrapK) { if (kf >= a.nz) kf -= a.nz; if (kq < 0) kq += a.nz; else if (kq >= a.nz) kq -= a.nz; }
      pfi.issue(a, kf, a.wrapK || s + a.prefetch < a.nz + RK, kq, a.wrapK || (kq >= 0 && kq < a.nz));
    }
    const int p = s - RK;
    int sp0 = slot - RK;             // slot of plane p
    if (sp0 < 0) sp0 += NQ;
    int kp = ks - RK;                // storage plane of p
    if (ND == 3 && a.wrapK && kp < 0) kp += a.nz;
    const long poff = (ND == 3) ? (long)kp * a.plane : 0;
    // issue the loads of the arriving own point and of the halo point of plane p together
    const bool doHalo = hk && p >= kc0;
    if (!PIPE) issueLoads(s, ks);
    if (inside) {
      Prim<ND> sa;
      dependent<ND>(Qs, gamma, sa);
#pragma unroll
      for (int d = 0; d < ND; ++d) kqc[((size_t)slot * NP + d) * NT] = sa.u[d];
      kqc[((size_t)slot * NP + ND) * NT] = sa.T;
    }
    if (ND == 3) {
      ++ks;
      if (a.wrapK && ks >= a.nz) ks -= a.nz;
      if (++slot >= NQ) slot = 0;
    }
    if (p < kc0) {
      if (PIPE && s + 1 < kc1 + RK) issueLoads(s + 1, ks);
      continue;
    }
    // ---- output plane p: in-plane tile (own point from the queue, halo from global memory)
    double Tp = 0.0;
    if (inside) {
#pragma unroll
      for (int f = 0; f < NP; ++f) {
        const double v = kqc[((size_t)sp0 * NP + f) * NT];
        tc[f * H * W] = v;
        if (f == ND) Tp = v;
      }
    }
    if (doHalo) {
      double* const th = T0 + hoff;
      Prim<ND> sh;
      dependent<ND>(Qh, gamma, sh);
#pragma unroll
      for (int d = 0; d < ND; ++d) th[d * H * W] = sh.u[d];
      th[ND * H * W] = sh.T;
    }
    __syncthreads();
    if (PIPE && s + 1 < kc1 + RK) issueLoads(s + 1, ks);
    if (mine && a.viscous) {
      const long off = poff + pij;
      // geometry loads first (latency overlaps the stencil arithmetic)
      const double jac = __ldg(a.jac + off);
      double M[ND * ND];
#pragma unroll
      for (int c = 0; c < ND * ND; ++c)
        if (CURV || (c % ND) == (c / ND)) M[c] = __ldg(a.m + (size_t)c * a.cs + off);
      // ---- derivatives of (u, T) along xi, eta, zeta
      double dxi[ND][NP];
#pragma unroll
      for (int f = 0; f < NP; ++f) {
        if (fastI) {
          double r = 0.0;
#pragma unroll
          for (int q = 1; q <= R; ++q) r += a.D[0].c[R + q] * (tc[f * H * W + q] - tc[f * H * W - q]);
          dxi[0][f] = r;
        } else {
          dxi[0][f] = tile_line_apply<W, H>(&a.ops->D[0], i, a.nx, T0, f, ty + R, tx + R, 0, i0 - R);
        }
        if (fastJ) {
          double r = 0.0;
#pragma unroll
          for (int q = 1; q <= R; ++q) r += a.D[1].c[R + q] * (tc[f * H * W + q * W] - tc[f * H * W - q * W]);
          dxi[1][f] = r;
        } else {
          dxi[1][f] = tile_line_apply<W, H>(&a.ops->D[1], j, a.ny, T0, f, ty + R, tx + R, 1, j0 - R);
        }
      }
      if constexpr (ND == 3) {
#pragma unroll
        for (int f = 0; f < NP; ++f) dxi[ND - 1][f] = 0.0;
#pragma unroll
        for (int q = 1; q <= RK; ++q) {
          int sp = sp0 + q, sm = sp0 - q;
          if (sp >= NQ) sp -= NQ;
          if (sm < 0) sm += NQ;
          const double cq = a.D[2].c[RK + q];
#pragma unroll
          for (int f = 0; f < NP; ++f)
            dxi[ND - 1][f] += cq * (kqc[((size_t)sp * NP + f) * NT] - kqc[((size_t)sm * NP + f) * NT]);
        }
      }
      // ---- gradient in physical space (reference src/GridImpl.f90:1357-1413), stress tensor, heat flux
      double g[ND * ND], gT[ND];
      if constexpr (CURV) {
#pragma unroll
        for (int c = 0; c < NP; ++c)
#pragma unroll
          for (int jx = 0; jx < ND; ++jx) {
            double r = M[jx] * dxi[0][c];
#pragma unroll
            for (int d = 1; d < ND; ++d) r += M[jx + ND * d] * dxi[d][c];
            r = jac * r;
            if (c < ND) g[jx + ND * c] = r; else gT[jx] = r;
          }
      } else {
#pragma unroll
        for (int jx = 0; jx < ND; ++jx) {
          const double mj = jac * M[jx + ND * jx];
#pragma unroll
          for (int c = 0; c < ND; ++c) g[jx + ND * c] = mj * dxi[jx][c];
          gT[jx] = mj * dxi[jx][ND];
        }
      }
      double mu, lam, kap, tau[ND * ND];
      transport(Tp, a.pp, mu, lam, kap);   // pow(): measured faster here than exp(n log x) (0.84 vs 0.90 ms)
      stress_from_gradient<ND>(g, mu, lam, tau);
      int t = 0;
#pragma unroll
      for (int r0 = 0; r0 < ND; ++r0)
#pragma unroll
        for (int c0 = r0; c0 < ND; ++c0) a.tauq[(size_t)(t++) * a.cs + off] = tau[c0 + ND * r0];
#pragma unroll
      for (int d = 0; d < ND; ++d) a.tauq[(size_t)(NTAU + d) * a.cs + off] = -kap * gT[d];
    }
    __syncthreads();
  }
}

// --------------------------------------------------------------------------- dissipation sweep
// State-only part of addDissipation (reference src/RhsHelperImpl.f90:58-81): out = sum_dir Diss_dir(X),
// X = a.Q (NU components).  In-plane neighbours from a shared-memory tile of X and the arc lengths, the k
// neighbours from a register queue of X.  Kept apart from sweep A so that neither kernel has to hold both
// the (u, T) and the X queue on chip (the fused version spilled to local memory).
template <int ND, int R, int DLO, int DN, int TLO, int TN, bool CLOS>
__global__ void __launch_bounds__(NT, 2) k_diss(FusedArgs a) {
  const bool COMPOSITE = a.composite != 0;
  constexpr int NU = ND + 2;
  constexpr int W = TX + 2 * R, H = TY + 2 * R;
  constexpr int NF = NU + 2;                   // X, arc_i, arc_j
  constexpr int FA = NU;
  constexpr int RK = (ND == 3) ? R : 0;
  constexpr int NQ = 2 * RK + 1;
  extern __shared__ double smem[];
  double* const T0 = smem;                                   // [NF][H][W]
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  int i0, j0;
  bool lastI, lastJ;
  tile_origin(blockIdx.x, a.nx, TX, i0, lastI);
  tile_origin(blockIdx.y, a.ny, TY, j0, lastJ);
  const int i = i0 + tx, j = j0 + ty;
  const bool mine = owns(i, a.nx, TX, lastI) && owns(j, a.ny, TY, lastJ);
  const bool inside = i < a.nx && j < a.ny;
  const long pij = (long)i + (long)a.nx * j;
  auto touches = [&](int d, int c0, int T, int n) {
    int depth = a.Dd[d].depth;
    if (!COMPOSITE) depth = max(depth, max(a.Dt[d].depth + a.Dd[d].width, a.dir[d].normDepth));
    return (a.dir[d].hasB0 && c0 < depth) || (a.dir[d].hasB1 && c0 + T > n - depth);
  };
  const bool fastI = !CLOS || !touches(0, i0, TX, a.nx);
  const bool fastJ = !CLOS || !touches(1, j0, TY, a.ny);
  double* const tc = T0 + (ty + R) * W + tx + R;

  int hoff = 0, hk = 0;
  long hp = -1;
  {
    const int h = threadIdx.x;
    if (h < 2 * R * TY) {
      const int ii = h % (2 * R), row = h / (2 * R);
      const int lc = ii < R ? ii : TX + ii;
      const int gi = wrap_index(i0 - R + lc, a.dir[0]);
      const int gj = j0 + row;
      hoff = (row + R) * W + lc;
      if (gi >= 0 && gj < a.ny) { hk = 1; hp = (long)gi + (long)a.nx * gj; }
    } else if (h < 2 * R * TY + 2 * R * TX) {
      const int h2 = h - 2 * R * TY;
      const int col = h2 % TX, jj = h2 / TX;
      const int lr = jj < R ? jj : TY + jj;
      const int gj = wrap_index(j0 - R + lr, a.dir[1]);
      const int gi = i0 + col;
      hoff = lr * W + col + R;
      if (gj >= 0 && gi < a.nx) { hk = 2; hp = (long)gi + (long)a.nx * gj; }
    }
  }
  const int kc0 = a.kBeg + (a.zOff + (int)blockIdx.z * a.zMul) * a.kChunk;
  const int kc1 = min(kc0 + a.kChunk, a.kEnd);
  auto wrapPlane = [&](int k) -> int {
    if (ND < 3 || !a.wrapK) return k;
    int kk = k % a.nz;
    return kk < 0 ? kk + a.nz : kk;
  };
  int ks = wrapPlane(kc0 - RK);

  double qq[NQ][NU];               // X at planes p-RK .. p+RK once the queue is primed
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int c = 0; c < NU; ++c) qq[q][c] = 0.0;
  PfItems<1> pfi;
  pfi.init(a, (long)i0 + (long)a.nx * j, tx, i0 < a.nx && j < a.ny);
  for (int s = kc0 - RK; s < kc1 + RK; ++s) {
    if (ND == 3 && a.prefetch) {
      int kf = ks + a.prefetch, kq = ks - RK + a.prefetch;
      if (a.wrapK) { if (kf >= a.nz) kf -= a.nz; if (kq < 0) kq += a.nz; else if (kq >= a.nz) kq -= a.nz; }
      pfi.issue(a, kf, a.wrapK || s + a.prefetch < a.nz + RK, kq, a.wrapK || (kq >= 0 && kq < a.nz));
    }
#pragma unroll
    for (int q = 0; q < NQ - 1; ++q)
#pragma unroll
      for (int c = 0; c < NU; ++c) qq[q][c] = qq[q + 1][c];
    const int p = s - RK;
    int kp = ks - RK;                // storage plane of p
    if (ND == 3 && a.wrapK && kp < 0) kp += a.nz;
    const long poff = (ND == 3) ? (long)kp * a.plane : 0;
    const bool out = p >= kc0;
    if (inside) {
      const double* __restrict__ Xp = a.Q + ((ND == 3) ? (long)ks * a.plane : 0) + pij;
#pragma unroll
      for (int c = 0; c < NU; ++c) qq[NQ - 1][c] = __ldg(Xp + (size_t)c * a.cs);
    }
    if (ND == 3) {
      ++ks;
      if (a.wrapK && ks >= a.nz) ks -= a.nz;
    }
    if (!out) continue;
    // ---- output plane p: in-plane tile (halo and arc lengths from global memory)
    double xh[NU], ah = 0.0, a0 = 0.0, a1 = 0.0, ak[TN];
    if (hk) {
      const long off = poff + hp;
#pragma unroll
      for (int c = 0; c < NU; ++c) xh[c] = __ldg(a.Q + (size_t)c * a.cs + off);
      if (!COMPOSITE) ah = __ldg(a.arc + (size_t)(hk - 1) * a.cs + off);
    }
    if (inside && !COMPOSITE) {
      a0 = __ldg(a.arc + (size_t)0 * a.cs + poff + pij);
      a1 = __ldg(a.arc + (size_t)1 * a.cs + poff + pij);
      if constexpr (ND == 3) {
#pragma unroll
        for (int ea = 0; ea < TN; ++ea) {
          int kk = kp + TLO + ea;
          if (a.wrapK) { if (kk < 0) kk += a.nz; else if (kk >= a.nz) kk -= a.nz; }
          ak[ea] = __ldg(a.arc + (size_t)2 * a.cs + (long)kk * a.plane + pij);
        }
      }
    }
    if (inside) {
#pragma unroll
      for (int c = 0; c < NU; ++c) tc[c * H * W] = qq[RK][c];
      if (!COMPOSITE) { tc[(FA + 0) * H * W] = a0; tc[(FA + 1) * H * W] = a1; }
    }
    if (hk) {
      double* const th = T0 + hoff;
#pragma unroll
      for (int c = 0; c < NU; ++c) th[c * H * W] = xh[c];
      if (!COMPOSITE) th[(FA + hk - 1) * H * W] = ah;
    }
    __syncthreads();
    if (mine) {
      double dz[NU];
#pragma unroll
      for (int c = 0; c < NU; ++c) dz[c] = 0.0;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const bool fast = d == 0 ? fastI : fastJ;
        const int st = d == 0 ? 1 : W;                 // tile stride along the direction
        if (fast) {
          double e[2 * R + 1];
          if (COMPOSITE) {
#pragma unroll
            for (int m = 0; m < 2 * R + 1; ++m) e[m] = a.Dd[d].c[m];
          } else {
#pragma unroll
            for (int m = 0; m < 2 * R + 1; ++m) e[m] = 0.0;
#pragma unroll
            for (int ea = 0; ea < TN; ++ea) {
              const double w = -a.Dt[d].c[ea] * tc[(FA + d) * H * W + (TLO + ea) * st];
#pragma unroll
              for (int eb = 0; eb < DN; ++eb) e[TLO + ea + DLO + eb + R] += w * a.Dd[d].c[eb];
            }
          }
#pragma unroll
          for (int c = 0; c < NU; ++c) {
            double r = 0.0;
#pragma unroll
            for (int m = 0; m < 2 * R + 1; ++m) r += e[m] * tc[c * H * W + (m - R) * st];
            dz[c] += r;
          }
        } else {
          const int cd = d == 0 ? i : j, nd = d == 0 ? a.nx : a.ny;
#pragma unroll
          for (int c = 0; c < NU; ++c) {
            dz[c] += COMPOSITE ? tile_line_apply<W, H>(&a.ops->Dd[d], cd, nd, T0, c, ty + R, tx + R, d, (d == 0 ? i0 : j0) - R)
                               : tile_line_dissipation<W, H>(a.ops, d, cd, T0, c, FA + d, ty + R, tx + R, (d == 0 ? i0 : j0) - R);
          }
        }
      }
      if constexpr (ND == 3) {
        double e[2 * RK + 1];
        if (COMPOSITE) {
#pragma unroll
          for (int m = 0; m < 2 * RK + 1; ++m) e[m] = a.Dd[2].c[m];
        } else {
#pragma unroll
          for (int m = 0; m < 2 * RK + 1; ++m) e[m] = 0.0;
#pragma unroll
          for (int ea = 0; ea < TN; ++ea) {
            const double w = -a.Dt[2].c[ea] * ak[ea];
#pragma unroll
            for (int eb = 0; eb < DN; ++eb) e[TLO + ea + DLO + eb + RK] += w * a.Dd[2].c[eb];
          }
        }
#pragma unroll
        for (int c = 0; c < NU; ++c) {
          double r = 0.0;
#pragma unroll
          for (int m = 0; m < 2 * RK + 1; ++m) r += e[m] * qq[m][c];
          dz[c] += r;
        }
      }
      const long off = poff + pij;
#pragma unroll
      for (int c = 0; c < NU; ++c) a.diss[(size_t)c * a.cs + off] = dz[c];
    }
    __syncthreads();
  }
}


template <int ND, int R, bool CURV, bool CLOS>
__global__ void __launch_bounds__(NT, 2) k_sweepB(FusedArgs a) {
  constexpr int NU = ND + 2;
  constexpr int W = TX + 2 * R, H = TY + 2 * R;
  constexpr int RK = (ND == 3) ? R : 0;
  constexpr int NQ = 2 * RK + 1;
  constexpr int ALLDIRS = (1 << ND) - 1;
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
  // fast interior stencils unless the tile touches a closure block of that direction
  const bool fastI = !CLOS || !((a.D[0].hasB0 && i0 < a.D[0].depth) || (a.D[0].hasB1 && i0 + TX > a.nx - a.D[0].depth));
  const bool fastJ = !CLOS || !((a.D[1].hasB0 && j0 < a.D[1].depth) || (a.D[1].hasB1 && j0 + TY > a.ny - a.D[1].depth));
  double* const f1c = F1 + ty * W + tx + R;        // component stride TY*W
  double* const f2c = F2 + (ty + R) * TX + tx;     // component stride H*TX, row stride TX
  double* const f3c = F3 + threadIdx.x;            // slot stride NU*NT, component stride NT

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
  const int kc0 = a.kBeg + (a.zOff + (int)blockIdx.z * a.zMul) * a.kChunk;
  const int kc1 = min(kc0 + a.kChunk, a.kEnd);
  // plane offsets advance incrementally (no division in the loop)
  auto wrapPlane = [&](int k) -> int {
    if (ND < 3 || !a.wrapK) return k;
    int kk = k % a.nz;
    return kk < 0 ? kk + a.nz : kk;
  };
  int ks = wrapPlane(kc0 - RK);                    // storage plane of the arriving plane s
  int kp = wrapPlane(kc0 - 2 * RK);                // storage plane of the output plane p = s - RK
  int slot = 0;                                    // queue slot of plane s

  // The queue holds, for every plane in flight, the running sum of div(F): the in-plane part is added when
  // the plane arrives, the k-derivative contributions c_q (F3(p+q) - F3(p-q)) as the neighbouring planes
  // arrive (thread-private columns of shared memory: no register queue).
  if (inside) {
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int c = 0; c < NU; ++c) f3c[((size_t)q * NU + c) * NT] = 0.0;
  }
  // output of plane p (storage plane kpl, queue slot sp): x 1/J, dissipation, RK4 substep
  // inputs of the output plane (loaded at the top of an iteration, consumed by emit)
  double ejac = 0.0, dss[NU], vb1[NU], vb2[NU];
  auto emit_load = [&](int kpl) {
    const long off = ((ND == 3) ? (long)kpl * a.plane : 0) + pij;
    ejac = __ldg(a.jac + off);
#pragma unroll
    for (int c = 0; c < NU; ++c) {
      const size_t qi = (size_t)c * a.cs + off;
      dss[c] = a.dissIn ? __ldg(a.dissIn + qi) : 0.0;
      if (a.fuseRk) {
        vb1[c] = (a.stage == 1) ? a.Q[qi] : ((a.stage == 4) ? 0.0 : a.b1in[qi]);
        vb2[c] = (a.stage == 1) ? 0.0 : a.b2[qi];
      }
    }
  };
  auto emit = [&](int kpl, int sp, const double* extra) {
    const long off = ((ND == 3) ? (long)kpl * a.plane : 0) + pij;
    const double jac = ejac;
    double r[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) {
      double rhs = 0.0 - (f3c[((size_t)sp * NU + c) * NT] + extra[c]);
      if (a.dissIn) rhs += a.dissAmount * dss[c];
      r[c] = rhs * jac;
    }
    if (!a.fuseRk) {
#pragma unroll
      for (int c = 0; c < NU; ++c) a.rhs[(size_t)c * a.cs + off] = r[c];
    } else {
      // RK4 substep (reference src/RK4IntegratorImpl.f90:106-158) fused into the last RHS kernel;
      // in stage 1 buffer1 is the input Q buffer itself (vb1 = Q).  rkB / rkQ = dt x the stage weights.
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        const size_t qi = (size_t)c * a.cs + off;
        if (a.stage != 4) a.b2[qi] = ((a.stage == 1) ? vb1[c] : vb2[c]) + a.rkB * r[c];
        a.Qout[qi] = ((a.stage == 4) ? vb2[c] : vb1[c]) + a.rkQ * r[c];
      }
    }
  };
  PfItems<3> pfi;
  pfi.init(a, (long)i0 + (long)a.nx * j, tx, i0 < a.nx && j < a.ny);
  for (int s = kc0 - RK; s < kc1 + RK; ++s) {
    const long soff = (ND == 3) ? (long)ks * a.plane : 0;
    const bool planeActive = s >= kc0 && s < kc1;
    if (ND == 3 && a.prefetch) {
      int kf = ks + a.prefetch, kq = kp + a.prefetch;
      if (a.wrapK) { if (kf >= a.nz) kf -= a.nz; if (kq < 0) kq += a.nz; else if (kq >= a.nz) kq -= a.nz; }
      pfi.issue(a, kf, a.wrapK || s + a.prefetch < a.nz + RK, kq, a.wrapK || (kq >= 0 && kq < a.nz));
    }
    if constexpr (ND == 3) {
      if (s - RK >= kc0 && mine) emit_load(kp);
    }
    // ---- arrival of plane s: own point (loads issued back to back, then the flux evaluation) ...
    double f3[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) f3[c] = 0.0;
    if (inside) {
      RawPoint<ND> raw;
      double Fh[ND][NU];
      if (planeActive) {
        load_raw<ND, ALLDIRS, CURV>(a, soff + pij, raw);
        fluxes_from_raw<ND, ALLDIRS, CURV>(a, raw, Fh);
#pragma unroll
        for (int c = 0; c < NU; ++c) {
          f1c[c * TY * W] = Fh[0][c];
          f2c[c * H * TX] = Fh[1][c];
        }
      } else {
        load_raw<ND, (ND == 3 ? 4 : 0), CURV>(a, soff + pij, raw);
        fluxes_from_raw<ND, (ND == 3 ? 4 : 0), CURV>(a, raw, Fh);
      }
      if constexpr (ND == 3) {
        // scatter c_q F3(s) to the planes s-q (+) and s+q (-); plane s+RK is touched for the first time
#pragma unroll
        for (int c = 0; c < NU; ++c) f3[c] = Fh[ND - 1][c];
#pragma unroll
        for (int q = 1; q <= RK; ++q) {
          int sp = slot + q, sm = slot - q;
          if (sp >= NQ) sp -= NQ;
          if (sm < 0) sm += NQ;
          const double cq = a.D[2].c[RK + q];
#pragma unroll
          for (int c = 0; c < NU; ++c) {
            if (q == RK) f3c[((size_t)sp * NU + c) * NT] = 0.0 - cq * f3[c];
            else f3c[((size_t)sp * NU + c) * NT] -= cq * f3[c];
            if (q < RK) f3c[((size_t)sm * NU + c) * NT] += cq * f3[c];
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < NU; ++c) f3c[c * NT] = 0.0;
      }
    }
    // ... then the halo point of this thread (xi- or eta-halo)
    if (planeActive && hk) {
      RawPoint<ND> raw;
      double Fh[ND][NU];
      if (hk == 1) {
        load_raw<ND, 1, CURV>(a, soff + hp, raw);
        fluxes_from_raw<ND, 1, CURV>(a, raw, Fh);
#pragma unroll
        for (int c = 0; c < NU; ++c) F1[((size_t)c * TY + hrow) * W + hcol] = Fh[0][c];
      } else {
        load_raw<ND, 2, CURV>(a, soff + hp, raw);
        fluxes_from_raw<ND, 2, CURV>(a, raw, Fh);
#pragma unroll
        for (int c = 0; c < NU; ++c) F2[((size_t)c * H + hrow) * TX + hcol] = Fh[1][c];
      }
    }
    // ---- output plane p = s - RK (3-D): every contribution but the last (c_RK F3(s), still in registers)
    // is already in the queue
    if constexpr (ND == 3) {
      if (s - RK >= kc0 && mine) {
        int sp0 = slot - RK;
        if (sp0 < 0) sp0 += NQ;
        double last[NU];
#pragma unroll
        for (int c = 0; c < NU; ++c) last[c] = a.D[2].c[2 * RK] * f3[c];
        emit(kp, sp0, last);
      }
    }
    __syncthreads();
    if (planeActive && mine) {
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        double r;
        if (fastI) {
          r = 0.0;
#pragma unroll
          for (int q = 1; q <= R; ++q) r += a.D[0].c[R + q] * (f1c[c * TY * W + q] - f1c[c * TY * W - q]);
        } else {
          r = strided_line_apply<W>(&a.ops->D[0], i, a.nx, F1 + ((size_t)c * TY + ty) * W, 1, i0 - R);
        }
        if (fastJ) {
          double r2 = 0.0;
#pragma unroll
          for (int q = 1; q <= R; ++q)
            r2 += a.D[1].c[R + q] * (f2c[c * H * TX + q * TX] - f2c[c * H * TX - q * TX]);
          r += r2;
        } else {
          r += strided_line_apply<H>(&a.ops->D[1], j, a.ny, F2 + (size_t)c * H * TX + tx, TX, j0 - R);
        }
        f3c[((size_t)slot * NU + c) * NT] += r;
      }
      if constexpr (ND == 2) { emit_load(0); emit(0, 0, f3); }       // f3 == 0 in 2-D
    }
    __syncthreads();
    // advance plane bookkeeping
    if (ND == 3) {
      ++ks; ++kp;
      if (a.wrapK) { if (ks >= a.nz) ks -= a.nz; if (kp >= a.nz) kp -= a.nz; }
      if (++slot >= NQ) slot = 0;
    }
  }
}

// ------------------------------------------------------------------------- adjoint sweep 1
// computeRhsAdjoint, first half (reference src/RhsHelperImpl.f90:408-553) + addDissipation(ADJOINT):
//   dW_d = D+_d w (adjoint first derivative, all directions)        -> in-plane tile + k register queue
//   rhs  = sum_d (A_d - B_d)^T dW_d  - sigma * Diss(w)               -> written to `rhs`
//   diffusion_j = sum_i B2(i,j)^T dW_i(2:)   (viscous)               -> written to `diffOut` (4 x nD comps)
// Same 2.5-D streaming structure as sweep A with X = w.  a.D = adjoint first derivative operators.
template <int ND, int R, int DLO, int DN, int TLO, int TN, bool CURV, bool CLOS>
__global__ void __launch_bounds__(NT, 2) k_adjoint1(FusedArgs a) {
  const bool COMPOSITE = a.composite != 0;
  constexpr int NU = ND + 2;
  constexpr int NTAU = ND * (ND + 1) / 2;
  constexpr int W = TX + 2 * R, H = TY + 2 * R;
  constexpr int NF = NU + 2;                   // w, arc_i, arc_j
  constexpr int FA = NU;
  constexpr int RK = (ND == 3) ? R : 0;
  constexpr int NQ = 2 * RK + 1;
  extern __shared__ double smem[];
  double* const T0 = smem;                                   // [NF][H][W]
  double* const WQ = smem + (size_t)NF * H * W;              // [NQ][NU][NT] k-queue of w (thread-private columns)
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  int i0, j0;
  bool lastI, lastJ;
  tile_origin(blockIdx.x, a.nx, TX, i0, lastI);
  tile_origin(blockIdx.y, a.ny, TY, j0, lastJ);
  const int i = i0 + tx, j = j0 + ty;
  const bool mine = owns(i, a.nx, TX, lastI) && owns(j, a.ny, TY, lastJ);
  const bool inside = i < a.nx && j < a.ny;
  const long pij = (long)i + (long)a.nx * j;
  auto touches = [&](int d, int c0, int T, int n) {
    int depth = a.D[d].depth;
    if (a.dissOn) {
      depth = max(depth, a.Dd[d].depth);
      if (!COMPOSITE) depth = max(depth, max(a.Dt[d].depth + a.Dd[d].width, a.dir[d].normDepth));
    }
    return (a.dir[d].hasB0 && c0 < depth) || (a.dir[d].hasB1 && c0 + T > n - depth);
  };
  const bool dissOn = a.dissOn != 0;
  const bool fastI = !CLOS || !touches(0, i0, TX, a.nx);
  const bool fastJ = !CLOS || !touches(1, j0, TY, a.ny);
  double* const tc = T0 + (ty + R) * W + tx + R;
  double* const wqc = WQ + threadIdx.x;                      // slot stride NU*NT, component stride NT

  int hcol = 0, hrow = 0, hk = 0;
  long hp = -1;
  {
    const int h = threadIdx.x;
    if (h < 2 * R * TY) {
      const int ii = h % (2 * R), row = h / (2 * R);
      const int lc = ii < R ? ii : TX + ii;
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
  const int kc0 = a.kBeg + (a.zOff + (int)blockIdx.z * a.zMul) * a.kChunk;
  const int kc1 = min(kc0 + a.kChunk, a.kEnd);
  auto wrapPlane = [&](int k) -> int {
    if (ND < 3 || !a.wrapK) return k;
    int kk = k % a.nz;
    return kk < 0 ? kk + a.nz : kk;
  };
  int ks = wrapPlane(kc0 - RK);
  int slot = 0;                    // queue slot of the arriving plane s

  PfItems<3> pfi;
  pfi.init(a, (long)i0 + (long)a.nx * j, tx, i0 < a.nx && j < a.ny);
  for (int s = kc0 - RK; s < kc1 + RK; ++s) {
    if (ND == 3 && a.prefetch) {
      int kf = ks + a.prefetch, kq = ks - RK + a.prefetch;
      if (a.wrapK) { if (kf >= a.nz) kf -= a.nz; if (kq < 0) kq += a.nz; else if (kq >= a.nz) kq -= a.nz; }
      pfi.issue(a, kf, a.wrapK || s + a.prefetch < a.nz + RK, kq, a.wrapK || (kq >= 0 && kq < a.nz));
    }
    if (inside) {
      const double* __restrict__ Wp = a.Win + ((ND == 3) ? (long)ks * a.plane : 0) + pij;
      double wv[NU];
#pragma unroll
      for (int c = 0; c < NU; ++c) wv[c] = __ldg(Wp + (size_t)c * a.cs);
#pragma unroll
      for (int c = 0; c < NU; ++c) wqc[((size_t)slot * NU + c) * NT] = wv[c];
    }
    const int p = s - RK;
    int sp0 = slot - RK;             // slot of plane p
    if (sp0 < 0) sp0 += NQ;
    int kp = ks - RK;
    if (ND == 3 && a.wrapK && kp < 0) kp += a.nz;
    if (ND == 3) {
      ++ks;
      if (a.wrapK && ks >= a.nz) ks -= a.nz;
      if (++slot >= NQ) slot = 0;
    }
    if (p < kc0) continue;
    const long poff = (ND == 3) ? (long)kp * a.plane : 0;
    const long off = poff + pij;
    if (inside) {
#pragma unroll
      for (int c = 0; c < NU; ++c) tc[c * H * W] = wqc[((size_t)sp0 * NU + c) * NT];
      if (!COMPOSITE && dissOn) {
        tc[(FA + 0) * H * W] = a.arc[(size_t)0 * a.cs + off];
        tc[(FA + 1) * H * W] = a.arc[(size_t)1 * a.cs + off];
      }
    }
    if (hk) {
      const long hoff = poff + hp;
      double* const th = T0 + hrow * W + hcol;
#pragma unroll
      for (int c = 0; c < NU; ++c) th[c * H * W] = __ldg(a.Win + (size_t)c * a.cs + hoff);
      if (!COMPOSITE && dissOn) th[(FA + hk - 1) * H * W] = a.arc[(size_t)(hk - 1) * a.cs + hoff];
    }
    __syncthreads();
    if (mine) {
      double r[NU];
      // ---- adjoint dissipation first: r = - sigma * sum_dir Diss_dir(w)  (needs only the tile and the queue)
      if (dissOn) {
        double dz[NU];
#pragma unroll
        for (int c = 0; c < NU; ++c) dz[c] = 0.0;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const bool fast = d == 0 ? fastI : fastJ;
          const int st = d == 0 ? 1 : W;
          if (fast) {
            double e[2 * R + 1];
            if (COMPOSITE) {
#pragma unroll
              for (int m = 0; m < 2 * R + 1; ++m) e[m] = a.Dd[d].c[m];
            } else {
#pragma unroll
              for (int m = 0; m < 2 * R + 1; ++m) e[m] = 0.0;
#pragma unroll
              for (int ea = 0; ea < TN; ++ea) {
                const double w = -a.Dt[d].c[ea] * tc[(FA + d) * H * W + (TLO + ea) * st];
#pragma unroll
                for (int eb = 0; eb < DN; ++eb) e[TLO + ea + DLO + eb + R] += w * a.Dd[d].c[eb];
              }
            }
#pragma unroll
            for (int c = 0; c < NU; ++c) {
              double rr = 0.0;
#pragma unroll
              for (int m = 0; m < 2 * R + 1; ++m) rr += e[m] * tc[c * H * W + (m - R) * st];
              dz[c] += rr;
            }
          } else {
            const int cd = d == 0 ? i : j, nd = d == 0 ? a.nx : a.ny;
#pragma unroll
            for (int c = 0; c < NU; ++c) {
              dz[c] += COMPOSITE ? tile_line_apply<W, H>(&a.ops->Dd[d], cd, nd, T0, c, ty + R, tx + R, d, (d == 0 ? i0 : j0) - R)
                                 : tile_line_dissipation<W, H>(a.ops, d, cd, T0, c, FA + d, ty + R, tx + R, (d == 0 ? i0 : j0) - R);
            }
          }
        }
        if constexpr (ND == 3) {
          double e[2 * RK + 1];
          if (COMPOSITE) {
#pragma unroll
            for (int m = 0; m < 2 * RK + 1; ++m) e[m] = a.Dd[2].c[m];
          } else {
#pragma unroll
            for (int m = 0; m < 2 * RK + 1; ++m) e[m] = 0.0;
#pragma unroll
            for (int ea = 0; ea < TN; ++ea) {
              int kk = kp + TLO + ea;
              if (a.wrapK) { if (kk < 0) kk += a.nz; else if (kk >= a.nz) kk -= a.nz; }
              const double w = -a.Dt[2].c[ea] * a.arc[(size_t)2 * a.cs + (long)kk * a.plane + pij];
#pragma unroll
              for (int eb = 0; eb < DN; ++eb) e[TLO + ea + DLO + eb + RK] += w * a.Dd[2].c[eb];
            }
          }
#pragma unroll
          for (int m = 0; m < 2 * RK + 1; ++m) {
            int sm = sp0 + m - RK;
            if (sm < 0) sm += NQ; else if (sm >= NQ) sm -= NQ;
#pragma unroll
            for (int c = 0; c < NU; ++c) dz[c] += e[m] * wqc[((size_t)sm * NU + c) * NT];
          }
        }
#pragma unroll
        for (int c = 0; c < NU; ++c) r[c] = 0.0 - a.dissAmount * dz[c];
      } else {
#pragma unroll
        for (int c = 0; c < NU; ++c) r[c] = 0.0;
      }
      // own-point inputs of the pointwise part
      double Q[NU], tq[NTAU + ND], M[ND * ND], jac = 0.0;
#pragma unroll
      for (int c = 0; c < NU; ++c) Q[c] = __ldg(a.Q + (size_t)c * a.cs + off);
      if (a.viscous) jac = __ldg(a.jac + off);
#pragma unroll
      for (int c = 0; c < ND * ND; ++c) {
        const bool diag = (c % ND) == (c / ND);
        if (CURV || diag) M[c] = __ldg(a.m + (size_t)c * a.cs + off);
      }
      asm volatile("" ::: "memory");   // scheduling fence: keep the phases' live ranges apart
      // ---- pointwise Jacobian-transpose products, one direction at a time
      Prim<ND> sp;
      dependent<ND>(Q, a.pp.gamma, sp);
      double mu = 0.0, lam = 0.0, kap = 0.0;
      if (a.viscous) transport<true>(sp.T, a.pp, mu, lam, kap);
      // adjoint first derivative of w along direction d at this point (tile for xi/eta, queue for zeta)
      auto deriv = [&](auto dI, double* dW) {
        constexpr int d = dI.value;
        if (d == 0) {
#pragma unroll
          for (int f = 0; f < NU; ++f) {
            if (fastI) {
              double t = 0.0;
#pragma unroll
              for (int q = 1; q <= R; ++q) t += a.D[0].c[R + q] * (tc[f * H * W + q] - tc[f * H * W - q]);
              dW[f] = t;
            } else {
              dW[f] = tile_line_apply<W, H>(&a.ops->D[0], i, a.nx, T0, f, ty + R, tx + R, 0, i0 - R);
            }
          }
        } else if (d == 1) {
#pragma unroll
          for (int f = 0; f < NU; ++f) {
            if (fastJ) {
              double t = 0.0;
#pragma unroll
              for (int q = 1; q <= R; ++q) t += a.D[1].c[R + q] * (tc[f * H * W + q * W] - tc[f * H * W - q * W]);
              dW[f] = t;
            } else {
              dW[f] = tile_line_apply<W, H>(&a.ops->D[1], j, a.ny, T0, f, ty + R, tx + R, 1, j0 - R);
            }
          }
        } else {
#pragma unroll
          for (int f = 0; f < NU; ++f) dW[f] = 0.0;
#pragma unroll
          for (int q = 1; q <= RK; ++q) {
            int sq = sp0 + q, sm = sp0 - q;
            if (sq >= NQ) sq -= NQ;
            if (sm < 0) sm += NQ;
            const double cq = a.D[2].c[RK + q];
#pragma unroll
            for (int f = 0; f < NU; ++f)
              dW[f] += cq * (wqc[((size_t)sq * NU + f) * NT] - wqc[((size_t)sm * NU + f) * NT]);
          }
        }
      };
      // pass 1: adjoint diffusion_j = sum_i B2(i,j)^T dW_i(2:)  (needs u, mu, lambda, kappa, 1/J, metrics only)
      if (a.viscous) {
        double dd[ND][ND + 1];
#pragma unroll
        for (int jj = 0; jj < ND; ++jj)
#pragma unroll
          for (int c = 0; c < ND + 1; ++c) dd[jj][c] = 0.0;
        static_for<ND>([&](auto dI) {
          constexpr int d = dI.value;
          double dW[NU];
          deriv(dI, dW);
          if constexpr (CURV) {
#pragma unroll
            for (int jj = 0; jj < ND; ++jj)
              add_second_partial_transpose<ND>(sp.u, mu, lam, kap, jac, &M[ND * d], &M[ND * jj], &dW[1], dd[jj]);
          } else {
            static_for<ND>([&](auto jj) {
              add_second_partial_transpose_rect<ND, d, jj.value>(sp.u, mu, lam, kap, jac, M[d + ND * d],
                                                                 M[jj.value + ND * jj.value], &dW[1], dd[jj.value]);
            });
          }
          asm volatile("" ::: "memory");
        });
#pragma unroll
        for (int jj = 0; jj < ND; ++jj)
#pragma unroll
          for (int c = 0; c < ND + 1; ++c) a.diffOut[(size_t)(c + (NU - 1) * jj) * a.cs + off] = dd[jj][c];
        asm volatile("" ::: "memory");
      }
      // pass 2: r += sum_d (A_d - B_d)^T dW_d
      double tau[ND * ND], qh[ND];
      if (a.viscous) {
#pragma unroll
        for (int e = 0; e < NTAU + ND; ++e) tq[e] = __ldg(a.tauqIn + (size_t)e * a.cs + off);
#pragma unroll
        for (int l = 0; l < ND; ++l)
#pragma unroll
          for (int c = 0; c < ND; ++c) tau[l + ND * c] = tq[tau_index<ND>(l, c)];
#pragma unroll
        for (int e = 0; e < ND; ++e) qh[e] = tq[NTAU + e];
      }
      static_for<ND>([&](auto dI) {
        constexpr int d = dI.value;
        double dW[NU];
        deriv(dI, dW);
        if constexpr (CURV) {
          add_flux_jacobian_transpose<ND>(Q, sp, &M[ND * d], a.pp.gamma, a.viscous, a.pp.powerLaw, tau, qh, dW, r);
        } else {
          add_flux_jacobian_transpose_rect<ND, d>(sp, M[d + ND * d], a.pp.gamma, a.viscous, a.pp.powerLaw, tau, qh,
                                                  dW, r);
        }
        asm volatile("" ::: "memory");
      });
#pragma unroll
      for (int c = 0; c < NU; ++c) a.rhs[(size_t)c * a.cs + off] = r[c];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------- adjoint sweep 2
// computeRhsAdjoint, second half (reference src/RhsHelperImpl.f90:555-570) + x 1/J + substepAdjointRK4:
//   t = sum_j D+_j diffusion_j ; variable change ; rhs = (rhsPartial -/+ t) / J ; RK4 substep on w.
// Same structure as sweep B with G_d = diffusion_d (NU-1 components, read from memory).
template <int ND, int R, bool CLOS>
__global__ void __launch_bounds__(NT, 2) k_adjoint2(FusedArgs a) {
  constexpr int NU = ND + 2;
  constexpr int NG = NU - 1;
  constexpr int W = TX + 2 * R, H = TY + 2 * R;
  constexpr int RK = (ND == 3) ? R : 0;
  constexpr int NQ = 2 * RK + 1;
  extern __shared__ double smem[];
  double* F1 = smem;                               // [NG][TY][W]
  double* F2 = F1 + (size_t)NG * TY * W;           // [NG][H][TX]
  double* F3 = F2 + (size_t)NG * H * TX;           // [NQ][NG][NT]
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  int i0, j0;
  bool lastI, lastJ;
  tile_origin(blockIdx.x, a.nx, TX, i0, lastI);
  tile_origin(blockIdx.y, a.ny, TY, j0, lastJ);
  const int i = i0 + tx, j = j0 + ty;
  const bool mine = owns(i, a.nx, TX, lastI) && owns(j, a.ny, TY, lastJ);
  const bool inside = i < a.nx && j < a.ny;
  const long pij = (long)i + (long)a.nx * j;
  const bool fastI = !CLOS || !((a.D[0].hasB0 && i0 < a.D[0].depth) || (a.D[0].hasB1 && i0 + TX > a.nx - a.D[0].depth));
  const bool fastJ = !CLOS || !((a.D[1].hasB0 && j0 < a.D[1].depth) || (a.D[1].hasB1 && j0 + TY > a.ny - a.D[1].depth));
  double* const f1c = F1 + ty * W + tx + R;
  double* const f2c = F2 + (ty + R) * TX + tx;
  double* const f3c = F3 + threadIdx.x;
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
  const int kc0 = a.kBeg + (a.zOff + (int)blockIdx.z * a.zMul) * a.kChunk;
  const int kc1 = min(kc0 + a.kChunk, a.kEnd);
  auto wrapPlane = [&](int k) -> int {
    if (ND < 3 || !a.wrapK) return k;
    int kk = k % a.nz;
    return kk < 0 ? kk + a.nz : kk;
  };
  int ks = wrapPlane(kc0 - RK);
  int kp = wrapPlane(kc0 - 2 * RK);
  int slot = 0;
  double rxy[RK + 1][NG];
#pragma unroll
  for (int q = 0; q < RK + 1; ++q)
#pragma unroll
    for (int c = 0; c < NG; ++c) rxy[q][c] = 0.0;
  PfItems<3> pfi;
  pfi.init(a, (long)i0 + (long)a.nx * j, tx, i0 < a.nx && j < a.ny);
  for (int s = kc0 - RK; s < kc1 + RK; ++s) {
    const long soff = (ND == 3) ? (long)ks * a.plane : 0;
    const bool planeActive = s >= kc0 && s < kc1;
    if (ND == 3 && a.prefetch) {
      int kf = ks + a.prefetch, kq = kp + a.prefetch;
      if (a.wrapK) { if (kf >= a.nz) kf -= a.nz; if (kq < 0) kq += a.nz; else if (kq >= a.nz) kq -= a.nz; }
      pfi.issue(a, kf, a.wrapK || s + a.prefetch < a.nz + RK, kq, a.wrapK || (kq >= 0 && kq < a.nz));
    }
    if (a.viscous) {
      if (inside) {
        const double* __restrict__ dp = a.diffIn + soff + pij;
        if constexpr (ND == 3) {
#pragma unroll
          for (int c = 0; c < NG; ++c) f3c[((size_t)slot * NG + c) * NT] = __ldg(dp + (size_t)(c + NG * 2) * a.cs);
        }
        if (planeActive) {
#pragma unroll
          for (int c = 0; c < NG; ++c) {
            f1c[c * TY * W] = __ldg(dp + (size_t)(c + NG * 0) * a.cs);
            f2c[c * H * TX] = __ldg(dp + (size_t)(c + NG * 1) * a.cs);
          }
        }
      }
      if (planeActive && hk) {
        const double* __restrict__ dp = a.diffIn + soff + hp;
        if (hk == 1) {
#pragma unroll
          for (int c = 0; c < NG; ++c) F1[((size_t)c * TY + hrow) * W + hcol] = __ldg(dp + (size_t)(c + NG * 0) * a.cs);
        } else {
#pragma unroll
          for (int c = 0; c < NG; ++c) F2[((size_t)c * H + hrow) * TX + hcol] = __ldg(dp + (size_t)(c + NG * 1) * a.cs);
        }
      }
    }
    // inputs of the output plane p = s - RK: issued before the barrier so that their latency overlaps the
    // in-plane derivative of the arriving plane
    const int p = s - RK;
    const bool emitNow = p >= kc0 && mine;
    const long off = ((ND == 3) ? (long)kp * a.plane : 0) + pij;
    double jac = 0.0, rp[NU], vb1[NU], vb2[NU], Q[NU];
    if (emitNow) {
      jac = __ldg(a.jac + off);
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        rp[c] = a.rhsIn[(size_t)c * a.cs + off];
        vb1[c] = 0.0;
        vb2[c] = 0.0;
      }
      if (a.viscous) {
#pragma unroll
        for (int c = 0; c < NU; ++c) Q[c] = __ldg(a.Q + (size_t)c * a.cs + off);
      }
      // RK buffers: one uniform branch on the stage, not one per component
      if (a.fuseRk) {
        if (a.stage == 1) {
#pragma unroll
          for (int c = 0; c < NU; ++c) vb1[c] = a.Win[(size_t)c * a.cs + off];
        } else if (a.stage == 4) {
#pragma unroll
          for (int c = 0; c < NU; ++c) vb2[c] = a.b2[(size_t)c * a.cs + off];
        } else {
#pragma unroll
          for (int c = 0; c < NU; ++c) {
            vb1[c] = a.b1in[(size_t)c * a.cs + off];
            vb2[c] = a.b2[(size_t)c * a.cs + off];
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RK; ++q)
#pragma unroll
      for (int c = 0; c < NG; ++c) rxy[q][c] = rxy[q + 1][c];
    if (a.viscous && planeActive && mine) {
#pragma unroll
      for (int c = 0; c < NG; ++c) {
        double r;
        if (fastI) {
          r = 0.0;
#pragma unroll
          for (int q = 1; q <= R; ++q) r += a.D[0].c[R + q] * (f1c[c * TY * W + q] - f1c[c * TY * W - q]);
        } else {
          r = strided_line_apply<W>(&a.ops->D[0], i, a.nx, F1 + ((size_t)c * TY + ty) * W, 1, i0 - R);
        }
        if (fastJ) {
          double r2 = 0.0;
#pragma unroll
          for (int q = 1; q <= R; ++q)
            r2 += a.D[1].c[R + q] * (f2c[c * H * TX + q * TX] - f2c[c * H * TX - q * TX]);
          r += r2;
        } else {
          r += strided_line_apply<H>(&a.ops->D[1], j, a.ny, F2 + (size_t)c * H * TX + tx, TX, j0 - R);
        }
        rxy[RK][c] = r;
      }
    }
    if (emitNow) {
      if (a.viscous) {
        double t[NG];
#pragma unroll
        for (int c = 0; c < NG; ++c) t[c] = rxy[0][c];
        if constexpr (ND == 3) {
          int sp0 = slot - RK;
          if (sp0 < 0) sp0 += NQ;
#pragma unroll
          for (int q = 1; q <= RK; ++q) {
            int sp = sp0 + q, sm = sp0 - q;
            if (sp >= NQ) sp -= NQ;
            if (sm < 0) sm += NQ;
            const double cq = a.D[2].c[RK + q];
#pragma unroll
            for (int c = 0; c < NG; ++c)
              t[c] += cq * (f3c[((size_t)sp * NG + c) * NT] - f3c[((size_t)sm * NG + c) * NT]);
          }
        }
        // variable change (reference src/RhsHelperImpl.f90:560-570)
        Prim<ND> sq;
        dependent<ND>(Q, a.pp.gamma, sq);
        t[ND] = a.pp.gamma * sq.v * t[ND];
        double ut = 0.0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          t[d] = sq.v * t[d] - sq.u[d] * t[ND];
          ut = (d == 0) ? sq.u[0] * t[0] : ut + sq.u[d] * t[d];
        }
#pragma unroll
        for (int c = 0; c < NG; ++c) rp[c + 1] -= t[c];
        rp[0] += sq.v * Q[NU - 1] * t[ND] + ut;
      }
      double r[NU];
#pragma unroll
      for (int c = 0; c < NU; ++c) r[c] = rp[c] * jac;
      if (!a.fuseRk) {
#pragma unroll
        for (int c = 0; c < NU; ++c) a.rhs[(size_t)c * a.cs + off] = r[c];
      } else if (a.stage == 1) {
#pragma unroll
        for (int c = 0; c < NU; ++c) {
          a.b2[(size_t)c * a.cs + off] = vb1[c] + a.rkB * r[c];
          a.Qout[(size_t)c * a.cs + off] = vb1[c] + a.rkQ * r[c];
        }
      } else if (a.stage == 4) {
#pragma unroll
        for (int c = 0; c < NU; ++c) a.Qout[(size_t)c * a.cs + off] = vb2[c] + a.rkQ * r[c];
      } else {
#pragma unroll
        for (int c = 0; c < NU; ++c) {
          a.b2[(size_t)c * a.cs + off] = vb2[c] + a.rkB * r[c];
          a.Qout[(size_t)c * a.cs + off] = vb1[c] + a.rkQ * r[c];
        }
      }
    }
    __syncthreads();
    if (ND == 3) {
      ++ks; ++kp;
      if (a.wrapK) { if (ks >= a.nz) ks -= a.nz; if (kp >= a.nz) kp -= a.nz; }
      if (++slot >= NQ) slot = 0;
    }
  }
}


template <int ND, int R, bool CURV, bool CLOS>
int launchA(const FusedArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int NP = ND + 1;
  constexpr int NQ = (ND == 3) ? 2 * R + 1 : 1;
  const size_t smem = sizeof(double) * ((size_t)NP * (TY + 2 * R) * (TX + 2 * R) + (size_t)NQ * NP * NT);
  // the software-pipelined variant exists for the closure-free instantiations (the benched ones)
  const bool pipe = !CLOS && mg_tuning_get("MG_SWEEPA_PIPE", MG_SWEEPA_PIPE_DEFAULT) != 0;
  auto kern = k_sweepA<ND, R, CURV, CLOS, false>;
  if (pipe) kern = k_sweepA<ND, R, CURV, CLOS, !CLOS>;
  static int configuredDevice[2] = {-1, -1};   // the attribute is per device and per kernel
  int device = 0;
  cudaGetDevice(&device);
  if (configuredDevice[pipe] != device) {
    MG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configuredDevice[pipe] = device;
  }
  mg_profile_begin("sweepA");
  kern<<<grid, NT, smem, st>>>(a);
  mg_profile_end();
  MG_CUDA(cudaGetLastError());
  mg_count_launches(1);
  return 0;
}

template <int ND, int R, int DLO, int DN, int TLO, int TN, bool CLOS>
int launchD(const FusedArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int NF = (ND + 2) + 2;
  const size_t smem = sizeof(double) * (size_t)NF * (TY + 2 * R) * (TX + 2 * R);
  auto kern = k_diss<ND, R, DLO, DN, TLO, TN, CLOS>;
  static int configuredDevice = -1;      // the attribute is per device: a second mg_init on another GPU sets it again
  int device = 0;
  cudaGetDevice(&device);
  if (configuredDevice != device) {
    MG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configuredDevice = device;
  }
  mg_profile_begin("dissipation");
  kern<<<grid, NT, smem, st>>>(a);
  mg_profile_end();
  MG_CUDA(cudaGetLastError());
  mg_count_launches(1);
  return 0;
}

template <int ND, int R, bool CURV, bool CLOS>
int launchB(const FusedArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int NU = ND + 2;
  constexpr int NQ = (ND == 3) ? 2 * R + 1 : 1;
  const size_t smem = sizeof(double) * ((size_t)NU * TY * (TX + 2 * R) + (size_t)NU * (TY + 2 * R) * TX +
                                        (size_t)NQ * NU * NT);
  auto kern = k_sweepB<ND, R, CURV, CLOS>;
  static int configuredDevice = -1;      // the attribute is per device: a second mg_init on another GPU sets it again
  int device = 0;
  cudaGetDevice(&device);
  if (configuredDevice != device) {
    MG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configuredDevice = device;
  }
  mg_profile_begin("sweepB");
  kern<<<grid, NT, smem, st>>>(a);
  mg_profile_end();
  MG_CUDA(cudaGetLastError());
  mg_count_launches(1);
  return 0;
}


template <int ND, int R>
int launchB_(const FusedArgs& a, dim3 grid, cudaStream_t st) {
  const bool clos = has_closures(a);
  return a.curvilinear ? (clos ? launchB<ND, R, true, true>(a, grid, st) : launchB<ND, R, true, false>(a, grid, st))
                       : (clos ? launchB<ND, R, false, true>(a, grid, st) : launchB<ND, R, false, false>(a, grid, st));
}

}  // namespace

// Can direction d of the grid be covered by tiles of extent T (plus a halo of R) by the fused kernels?
static bool dir_fits_tile(const mg_grid* g, int d, int T, int R, int mode) {
  if (g->periodicityType[d] == MG_PERIODIC_NONE) {
    // a closure block (and its adjoint-free forward operators) must fit in one tile + halo
    const MgDevOp& o = (mode == MG_ADJOINT ? g->adjointFirstDerivative[d] : g->firstDerivative[d])->op;
    if (o.boundaryWidth > T + R || o.boundaryDepth > T || g->localSize[d] < 2 * o.boundaryDepth) return false;
    // The last tile of a direction is anchored at the far boundary; when the line is not a multiple of the
    // tile it takes over the points >= n - T from the first tile.  Every point of the LEFT closure region
    // (widest operator used along the direction) must stay with the first tile, whose halo holds the block.
    int depth = std::max(o.boundaryDepth, g->firstDerivative[d]->op.boundaryDepth);
    int width = o.boundaryWidth;
    if (g->dissipationOn) {
      const MgDevOp& dd = g->dissipation[d]->op;
      depth = std::max(depth, dd.boundaryDepth);
      width = std::max(width, dd.boundaryWidth);
      if (!g->compositeDissipation) {
        const MgDevOp& dt = g->dissipationTranspose[d]->op;
        depth = std::max(depth, std::max(dt.boundaryDepth + dd.boundaryWidth, g->firstDerivative[d]->op.normDepth));
        width = std::max(width, dt.boundaryWidth);
      }
    }
    const int n = g->localSize[d];
    if (n > T && n % T != 0 && n - T < depth) return false;
    if (depth > T || width > T + R) return false;
    if (n < 2 * depth) return false;
  } else if (g->localSize[d] < T) {
    // a periodic line shorter than the tile would have to wrap inside the tile: general path
    return false;
  }
  return true;
}

// The fused sweeps can evaluate the RHS of this state (interior scheme + closures); patches and sources, if any,
// are applied afterwards by the caller (state.cu: fused RHS + patch epilogue).
int mg_fused_rhs_supported(const mg_state* s, int mode) {
  const mg_grid* g = s->grid;
  if (mode != MG_FORWARD && mode != MG_ADJOINT) return 0;
  if (g->nD < 2) return 0;
  if (g->iblank) return 0;
  if (g->procDims[0] != 1 || g->procDims[1] != 1) return 0;       // bricks split along i / j: operator path
  if (g->nD == 3 && g->periodicityType[2] != MG_PERIODIC_PLANE) return 0;
  SchemeInfo si;
  if (!scheme_of(g, &si)) return 0;
  if (g->nD == 3 && g->localSize[2] < si.R) return 0;
  for (int d = 0; d < 2; ++d)
    if (!dir_fits_tile(g, d, TX, si.R, mode)) return 0;
  return 1;
}

// ... and can also fold the RK4 substep into the last sweep: nothing may touch the RHS after the sweeps
int mg_fused_supported(const mg_state* s, int mode) {
  if (!s->patches.empty() || !s->acousticSources.empty()) return 0;
  if (mode == MG_ADJOINT && s->limits.soft) return 0;        // the solution-limit forcing joins after the sweeps
  if (s->bodyForce) return 0;                                // the region's body force joins after the sweeps
  return mg_fused_rhs_supported(s, mode);
}

int mg_fused_alloc(mg_state* s) {
  mg_grid* g = s->grid;
  const int nD = s->nD;
  const int nTauQ = nD * (nD + 1) / 2 + nD;
  if (s->opt.viscosityOn && s->tauq.nComp != nTauQ) MG_TRY(mg_field_alloc(g, nTauQ, &s->tauq));
  if (s->opt.dissipationOn && s->dissTerm.nComp != s->nU) MG_TRY(mg_field_alloc(g, s->nU, &s->dissTerm));
  return 0;
}

// Sweep A: state update (stress tensor + heat flux).  The dissipation term is computed on demand by
// mg_fused_dissipation (the adjoint path never needs it for the forward state).
int mg_fused_sweepA(mg_state* s) {
  mg_grid* g = s->grid;
  MG_TRY(mg_fused_alloc(s));
  s->dissValid = false;
  FusedArgs a;
  MG_TRY(fill_args(s, &a));
  MG_TRY(upload_ops(s, 0, &a));
  a.Q = s->Q[s->cur].comp(0);
  a.tauq = s->opt.viscosityOn ? s->tauq.comp(0) : nullptr;
  if (!a.viscous) { s->fusedValid = true; return 0; }
  SchemeInfo si;
  scheme_of(g, &si);
  {  // sweep A streams
    PfList pfl; pfl.cs = a.cs;
    pfl.addIn(a.Q, s->nU);
    pfl.addOut(a.jac, 1);
    for (int d = 0; d < s->nD; ++d) pfl.addOut(a.m + (size_t)(d + s->nD * d) * a.cs, 1);
    pfl.finish(&a);
  }
  const int nChunksA = choose_chunks(&a, si.R, 3);
  cudaStream_t st = mg_stream();
  const bool clos = has_closures(a);
  (void)clos;
  auto go = [&](FusedArgs& a, int zBlocks, cudaStream_t st) -> int {
  const dim3 grid = tiles(a, zBlocks);
  int rc = -1;
#define MG_A(ND_, R_)                                                                                     \
  if (s->nD == ND_ && si.R == R_)                                                                         \
    rc = a.curvilinear ? (clos ? launchA<ND_, R_, true, true>(a, grid, st) : launchA<ND_, R_, true, false>(a, grid, st))  \
                       : (clos ? launchA<ND_, R_, false, true>(a, grid, st) : launchA<ND_, R_, false, false>(a, grid, st));
#ifndef MG_DEV_ONLY_33
  MG_A(2, 2)
#endif
#ifndef MG_DEV_ONLY_33
  MG_A(2, 3)
#endif
#ifndef MG_DEV_ONLY_33
  MG_A(2, 4)
#endif
#ifndef MG_DEV_ONLY_33
  MG_A(3, 2)
#endif
  MG_A(3, 3)
#ifndef MG_DEV_ONLY_33
  MG_A(3, 4)
#endif
#undef MG_A
  return rc;
  };
  int rc = launch_split(a, nChunksA, si.R, go);
  if (rc != 0) return rc < 0 && rc != -2 ? (mg_set_error("fused sweep A: unsupported configuration"), -1) : rc;
  s->fusedValid = true;
  return 0;
}

// Dissipation sweep: dissTerm = sum_dir Diss_dir(Q) for the current conserved variables.
int mg_fused_dissipation(mg_state* s) {
  mg_grid* g = s->grid;
  if (!s->opt.dissipationOn) { s->dissValid = true; return 0; }
  MG_TRY(mg_halo_wait_pending());
  MG_TRY(mg_fused_alloc(s));
  FusedArgs a;
  MG_TRY(fill_args(s, &a));
  MG_TRY(upload_ops(s, 0, &a));
  a.Q = s->Q[s->cur].comp(0);
  a.diss = s->dissTerm.comp(0);
  if (!mg_tuning_has("MG_PREFETCH")) a.prefetch = 2;
  SchemeInfo si;
  scheme_of(g, &si);
  {  // dissipation sweep streams
    PfList pfl; pfl.cs = a.cs;
    pfl.addIn(a.Q, s->nU);
    if (!a.composite) pfl.addOut(a.arc, s->nD);
    pfl.finish(&a);
  }
  const dim3 grid = tiles(a, choose_chunks(&a, si.R, 2));
  cudaStream_t st = mg_stream();
  int rc = -1;
  const bool clos = has_closures(a);
  (void)clos;
#define MG_D(ND_, R_, DLO, DN, TLO, TN)                                                 \
  if (s->nD == ND_ && si.R == R_)                                                       \
    rc = clos ? launchD<ND_, R_, DLO, DN, TLO, TN, true>(a, grid, st) : launchD<ND_, R_, DLO, DN, TLO, TN, false>(a, grid, st);
#ifndef MG_DEV_ONLY_33
  MG_D(2, 2, -1, 3, -1, 3)
#endif
#ifndef MG_DEV_ONLY_33
  MG_D(2, 3, -2, 4, -1, 4)
#endif
#ifndef MG_DEV_ONLY_33
  MG_D(2, 4, -2, 5, -2, 5)
#endif
#ifndef MG_DEV_ONLY_33
  MG_D(3, 2, -1, 3, -1, 3)
#endif
  MG_D(3, 3, -2, 4, -1, 4)
#ifndef MG_DEV_ONLY_33
  MG_D(3, 4, -2, 5, -2, 5)
#endif
#undef MG_D
  if (rc != 0) return rc < 0 && rc != -2 ? (mg_set_error("fused dissipation: unsupported configuration"), -1) : rc;
  s->dissValid = true;
  return 0;
}

// Sweep B: RHS (+ RK4 substep when fuseRk)
// tile height of the folded sweep B: 12 (192 threads, 2 CTAs/SM) exists for the 3-D viscous instantiations with
// non-composite dissipation and R <= 3, else 8 (128 threads)
static int bd_tile_height(const mg_state* s, int R) {
  const int tyPref = mg_tuning_get("MG_BD_TY", 12);
  const bool hot = s->opt.viscosityOn && s->opt.dissipationOn && !s->grid->compositeDissipation;
  return (tyPref == 12 && hot && s->nD == 3 && R <= 3) ? 12 : 8;
}

int mg_fused_sweepB(mg_state* s, int fuseRk, int stage, double dt) {
  mg_grid* g = s->grid;
  if (!s->fusedValid) MG_FAIL("fused sweep B: state has not been updated (sweep A)");
  // MG_FWD=1 selects the first-generation three-sweep forward stage (separate dissipation sweep; kept for A/B
  // measurements); default: dissipation folded into sweep B (fused_sweepbd.cuh), two sweeps per stage
  SchemeInfo si;
  scheme_of(g, &si);
  const int genPref = mg_tuning_get("MG_FWD", 2);
  // the folded kernel uses 16 x 12 / 16 x 8 tiles: eta closures wider than the tile stay with generation 1
  const int gen = (genPref == 2 && dir_fits_tile(g, 1, bd_tile_height(s, si.R), si.R, MG_FORWARD)) ? 2 : 1;
  if (gen != 2 && s->opt.dissipationOn && !s->dissValid) MG_TRY(mg_fused_dissipation(s));
  FusedArgs a;
  MG_TRY(fill_args(s, &a));
  MG_TRY(upload_ops(s, 0, &a));
  a.Q = s->Q[s->cur].comp(0);
  a.tauqIn = s->opt.viscosityOn ? s->tauq.comp(0) : nullptr;
  a.dissIn = (gen != 2 && s->opt.dissipationOn) ? s->dissTerm.comp(0) : nullptr;
  a.dissOn = s->opt.dissipationOn;
  a.rhs = s->rhs.comp(0);
  a.fuseRk = fuseRk;
  a.stage = stage;
  a.dt = dt;
  a.rkB = (stage == 1) ? dt / 6.0 : dt / 3.0;
  a.rkQ = (stage == 3) ? dt : ((stage == 4) ? dt / 6.0 : dt / 2.0);
  if (fuseRk) {
    // the buffer about to receive the new Q must not be shared with a checkpoint slot
    MG_TRY(mg_state_make_exclusive(s, stage == 1 ? &s->rk1 : &s->Q[1 - s->cur], false));
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
  {  // sweep B streams
    PfList pfl; pfl.cs = a.cs;
    pfl.addIn(a.Q, s->nU);
    if (a.viscous) pfl.addIn(a.tauqIn, s->nD * (s->nD + 1) / 2 + s->nD);
    for (int d = 0; d < s->nD; ++d) pfl.addIn(a.m + (size_t)(d + s->nD * d) * a.cs, 1);
    pfl.addOut(a.jac, 1);
    pfl.addOut(a.dissIn, s->nU);
    if (gen == 2 && a.dissOn && !a.composite) pfl.addIn(a.arc, s->nD);
    if (fuseRk && stage != 1) pfl.addOut(a.b2, s->nU);
    if (fuseRk && (stage == 2 || stage == 3)) pfl.addOut(a.b1in, s->nU);
    pfl.finish(&a);
  }
  cudaStream_t st = mg_stream();
  int rc = -1;
  if (gen == 2) {
    const bool hot = a.viscous && a.dissOn && !a.composite;
    const int tileY = bd_tile_height(s, si.R);
    const int resident = (tileY == 8 && si.R < 4) ? 3 : 2;
    const int nChunks = choose_chunks(&a, si.R, resident, TX, tileY);
    rc = launch_split(a, nChunks, si.R, [&](FusedArgs& aa, int zBlocks, cudaStream_t stx) -> int {
      return hot ? mg_fused_sweepbd_hot_launch(&aa, s->nD, si.R, tileY, zBlocks, stx)
                 : mg_fused_sweepbd_gen_launch(&aa, s->nD, si.R, tileY, zBlocks, stx);
    });
    if (rc == -1) MG_FAIL("fused sweep B: unsupported configuration");
  } else {
  MG_TRY(mg_halo_wait_pending());
  const dim3 grid = tiles(a, choose_chunks(&a, si.R, 2));
  const bool clos = has_closures(a);
  (void)clos;
#ifndef MG_DEV_ONLY_33
  if (s->nD == 2 && si.R == 2) rc = launchB_<2, 2>(a, grid, st);
#endif
#ifndef MG_DEV_ONLY_33
  if (s->nD == 2 && si.R == 3) rc = launchB_<2, 3>(a, grid, st);
#endif
#ifndef MG_DEV_ONLY_33
  if (s->nD == 2 && si.R == 4) rc = launchB_<2, 4>(a, grid, st);
#endif
#ifndef MG_DEV_ONLY_33
  if (s->nD == 3 && si.R == 2) rc = launchB_<3, 2>(a, grid, st);
#endif
  if (s->nD == 3 && si.R == 3) rc = launchB_<3, 3>(a, grid, st);
#ifndef MG_DEV_ONLY_33
  if (s->nD == 3 && si.R == 4) rc = launchB_<3, 4>(a, grid, st);
#endif
  }
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

// ------------------------------------------------------------------------------ adjoint host side
namespace {

template <int ND, int R, int DLO, int DN, int TLO, int TN, bool CURV, bool CLOS>
int launchAdj1(const FusedArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int NF = (ND + 2) + 2;
  constexpr int NQ = (ND == 3) ? 2 * R + 1 : 1;
  const size_t smem = sizeof(double) * ((size_t)NF * (TY + 2 * R) * (TX + 2 * R) + (size_t)NQ * (ND + 2) * NT);
  auto kern = k_adjoint1<ND, R, DLO, DN, TLO, TN, CURV, CLOS>;
  static int configuredDevice = -1;      // the attribute is per device: a second mg_init on another GPU sets it again
  int device = 0;
  cudaGetDevice(&device);
  if (configuredDevice != device) {
    MG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configuredDevice = device;
  }
  mg_profile_begin("adjoint1");
  kern<<<grid, NT, smem, st>>>(a);
  mg_profile_end();
  MG_CUDA(cudaGetLastError());
  mg_count_launches(1);
  return 0;
}

template <int ND, int R, bool CLOS>
int launchAdj2(const FusedArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int NG = ND + 1;
  constexpr int NQ = (ND == 3) ? 2 * R + 1 : 1;
  const size_t smem = sizeof(double) * ((size_t)NG * TY * (TX + 2 * R) + (size_t)NG * (TY + 2 * R) * TX +
                                        (size_t)NQ * NG * NT);
  auto kern = k_adjoint2<ND, R, CLOS>;
  static int configuredDevice = -1;      // the attribute is per device: a second mg_init on another GPU sets it again
  int device = 0;
  cudaGetDevice(&device);
  if (configuredDevice != device) {
    MG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configuredDevice = device;
  }
  mg_profile_begin("adjoint2");
  kern<<<grid, NT, smem, st>>>(a);
  mg_profile_end();
  MG_CUDA(cudaGetLastError());
  mg_count_launches(1);
  return 0;
}

int fill_args_adjoint(mg_state* s, FusedArgs* a) {
  mg_grid* g = s->grid;
  MG_TRY(fill_args(s, a));
  for (int d = 0; d < g->nD; ++d) MG_TRY(fill_lineop(g->adjointFirstDerivative[d], &a->D[d]));
  a->Q = s->Q[s->cur].comp(0);
  a->Win = s->W[s->curW].comp(0);
  a->tauqIn = s->opt.viscosityOn ? s->tauq.comp(0) : nullptr;
  a->dissOn = s->opt.dissipationOn;
  // measured on B200: adjoint sweep 1 runs faster without the L2 prefetch (sweep 2 overrides this with 1)
  // L2 prefetch distance of adjoint sweep 1: the second-generation kernel is latency bound and gains from it,
  // the first generation (shared-memory bound) ran faster without
  const int pfAdj = mg_tuning_get("MG_PREFETCH_ADJ", mg_tuning_get("MG_ADJ1", 2) == 2 ? 1 : 0);
  a->prefetch = pfAdj;
  MG_TRY(upload_ops(s, 1, a));
  return 0;
}

}  // namespace

// Adjoint sweep 1: rhs (partial) and adjoint diffusion.  Needs the forward state's sweep-A outputs.
int mg_fused_adjoint1(mg_state* s) {
  mg_grid* g = s->grid;
  if (!s->fusedValid) MG_TRY(mg_fused_sweepA(s));
  FusedArgs a;
  MG_TRY(fill_args_adjoint(s, &a));
  a.rhs = s->rhs.comp(0);
  a.diffOut = g->scratchA.comp(0);
  if (g->scratchA.compStride != a.cs) MG_FAIL("fused adjoint: scratch stride mismatch");
  SchemeInfo si;
  scheme_of(g, &si);
  {  // adjoint sweep 1 streams
    PfList pfl; pfl.cs = a.cs;
    pfl.addIn(a.Win, s->nU);
    pfl.addOut(a.Q, s->nU);
    if (a.viscous) { pfl.addOut(a.tauqIn, s->nD * (s->nD + 1) / 2 + s->nD); pfl.addOut(a.jac, 1); }
    for (int d = 0; d < s->nD; ++d) pfl.addOut(a.m + (size_t)(d + s->nD * d) * a.cs, 1);
    if (a.dissOn && !a.composite) pfl.addOut(a.arc, s->nD);
    pfl.finish(&a);
  }
  cudaStream_t st = mg_stream();
  // MG_ADJ1=1 selects the first-generation kernel (kept for A/B measurements); default: fused_adjoint1.cuh
  const int gen = mg_tuning_get("MG_ADJ1", 2);
  if (gen == 2) {
    const bool hot = a.viscous && a.dissOn && !a.composite;
    // tile height 12 (192 threads, 168 registers) exists for the 3-D hot instantiations with R <= 3
    const int tyPref = mg_tuning_get("MG_ADJ1_TY", 12);
    int tileY = (tyPref == 12 && hot && s->nD == 3 && si.R <= 3) ? 12 : 16;
    if (tileY != 16 && !dir_fits_tile(g, 1, tileY, si.R, MG_ADJOINT)) tileY = 16;
    const int nChunks2 = choose_chunks(&a, si.R, 2, TX, tileY);
    // TMA-fed k-queue (cp.async.bulk.tensor): 3-D hot instantiations with 16 x 12 tiles; MG_TMA=0 disables
    CUtensorMap tmW;
    const bool useTma = hot && tileY == 12 && s->nD == 3 && si.R <= 3 && mg_tuning_get("MG_TMA", 1) &&
                        make_field_tensor_map(g, s->W[s->curW], TX, tileY, s->nU, &tmW);
    const int rc2 = launch_split(a, nChunks2, si.R, [&](FusedArgs& aa, int zBlocks, cudaStream_t stx) -> int {
      return hot ? mg_fused_adjoint1_hot_launch(&aa, s->nD, si.R, tileY, zBlocks, stx, useTma ? &tmW : nullptr)
                 : mg_fused_adjoint1_gen_launch(&aa, s->nD, si.R, tileY, zBlocks, stx);
    });
    if (rc2 == -1) MG_FAIL("fused adjoint sweep 1: unsupported configuration");
    return rc2;
  }
  MG_TRY(mg_halo_wait_pending());
  const int nChunks = choose_chunks(&a, si.R, 2);
  const dim3 grid = tiles(a, nChunks);
  int rc = -1;
  const bool clos = has_closures(a);
  (void)clos;
#define MG_J(ND_, R_, DLO, DN, TLO, TN)                                                 \
  if (s->nD == ND_ && si.R == R_)                                                       \
    rc = a.curvilinear ? (clos ? launchAdj1<ND_, R_, DLO, DN, TLO, TN, true, true>(a, grid, st)                      \
                               : launchAdj1<ND_, R_, DLO, DN, TLO, TN, true, false>(a, grid, st))                    \
                       : (clos ? launchAdj1<ND_, R_, DLO, DN, TLO, TN, false, true>(a, grid, st)                     \
                               : launchAdj1<ND_, R_, DLO, DN, TLO, TN, false, false>(a, grid, st));
#ifndef MG_DEV_ONLY_33
  MG_J(2, 2, -1, 3, -1, 3)
#endif
#ifndef MG_DEV_ONLY_33
  MG_J(2, 3, -2, 4, -1, 4)
#endif
#ifndef MG_DEV_ONLY_33
  MG_J(2, 4, -2, 5, -2, 5)
#endif
#ifndef MG_DEV_ONLY_33
  MG_J(3, 2, -1, 3, -1, 3)
#endif
  MG_J(3, 3, -2, 4, -1, 4)
#ifndef MG_DEV_ONLY_33
  MG_J(3, 4, -2, 5, -2, 5)
#endif
#undef MG_J
  return rc;
}

// Adjoint sweep 2: second derivative sweep, variable change, x 1/J (+ RK4 substep on w when fuseRk).
// `stage` is the reference's adjoint stage (4 -> 1); dt > 0.
int mg_fused_adjoint2(mg_state* s, int fuseRk, int stage, double dt) {
  mg_grid* g = s->grid;
  FusedArgs a;
  MG_TRY(fill_args_adjoint(s, &a));
  a.rhsIn = s->rhs.comp(0);
  a.rhs = s->rhs.comp(0);
  a.diffIn = g->scratchA.comp(0);
  if (!mg_tuning_has("MG_PREFETCH_ADJ")) a.prefetch = 1;   // measured: distance 1 helps this sweep, none helps sweep 1
  a.fuseRk = fuseRk;
  const int rkStage = 5 - stage;        // adjoint stage 4 plays the role of RK stage 1, ...
  a.stage = rkStage;
  a.dt = -dt;
  a.rkB = (rkStage == 1) ? -dt / 6.0 : -dt / 3.0;
  a.rkQ = (rkStage == 3) ? -dt : ((rkStage == 4) ? -dt / 6.0 : -dt / 2.0);
  if (fuseRk) {
    MG_TRY(mg_state_make_exclusive(s, rkStage == 1 ? &s->rk1 : &s->W[1 - s->curW], false));
    if (rkStage == 1) {
      a.b1in = a.Win;
      a.Qout = s->rk1.comp(0);
    } else {
      a.b1in = s->rk1.comp(0);
      a.Qout = s->W[1 - s->curW].comp(0);
    }
    a.b2 = s->rk2.comp(0);
  }
  SchemeInfo si;
  scheme_of(g, &si);
  {  // adjoint sweep 2 streams
    PfList pfl; pfl.cs = a.cs;
    if (a.viscous) pfl.addIn(a.diffIn, (s->nU - 1) * s->nD);
    pfl.addOut(a.jac, 1);
    pfl.addOut(a.rhsIn, s->nU);
    if (a.viscous) pfl.addOut(a.Q, s->nU);
    if (fuseRk && rkStage != 1) pfl.addOut(a.b2, s->nU);
    if (fuseRk && (rkStage == 2 || rkStage == 3)) pfl.addOut(a.b1in, s->nU);
    pfl.finish(&a);
  }
  const int nChunksJ = choose_chunks(&a, si.R, 2);
  cudaStream_t st = mg_stream();
  const bool clos = has_closures(a);
  (void)clos;
  auto go = [&](FusedArgs& a, int zBlocks, cudaStream_t st) -> int {
  const dim3 grid = tiles(a, zBlocks);
  int rc = -1;
#ifndef MG_DEV_ONLY_33
  if (s->nD == 2 && si.R == 2) rc = clos ? launchAdj2<2, 2, true>(a, grid, st) : launchAdj2<2, 2, false>(a, grid, st);
#endif
#ifndef MG_DEV_ONLY_33
  if (s->nD == 2 && si.R == 3) rc = clos ? launchAdj2<2, 3, true>(a, grid, st) : launchAdj2<2, 3, false>(a, grid, st);
#endif
#ifndef MG_DEV_ONLY_33
  if (s->nD == 2 && si.R == 4) rc = clos ? launchAdj2<2, 4, true>(a, grid, st) : launchAdj2<2, 4, false>(a, grid, st);
#endif
#ifndef MG_DEV_ONLY_33
  if (s->nD == 3 && si.R == 2) rc = clos ? launchAdj2<3, 2, true>(a, grid, st) : launchAdj2<3, 2, false>(a, grid, st);
#endif
  if (s->nD == 3 && si.R == 3) rc = clos ? launchAdj2<3, 3, true>(a, grid, st) : launchAdj2<3, 3, false>(a, grid, st);
#ifndef MG_DEV_ONLY_33
  if (s->nD == 3 && si.R == 4) rc = clos ? launchAdj2<3, 4, true>(a, grid, st) : launchAdj2<3, 4, false>(a, grid, st);
#endif
  return rc;
  };
  int rc = launch_split(a, nChunksJ, si.R, go);
  if (rc != 0) return rc;
  if (fuseRk) {
    if (rkStage == 1) {
      MgField tmp = s->rk1;
      s->rk1 = s->W[s->curW];
      s->W[s->curW] = tmp;
    } else {
      s->curW = 1 - s->curW;
    }
  }
  return 0;
}
