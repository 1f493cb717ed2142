// Adjoint sweep 1, second generation (the dominant kernel of the forward+adjoint step).
//
// computeRhsAdjoint, first half (reference src/RhsHelperImpl.f90:408-553) + addDissipation(ADJOINT) (:10-87):
//   dW_d        = D+_d w                         (adjoint first derivative, every direction)
//   rhs         = sum_d (A_d - B_d)^T dW_d - sigma * Diss(w)          -> a.rhs
//   diffusion_j = sum_d B2(d,j)^T dW_d(2:)       (viscous)            -> a.diffOut (4 x nD components)
//
// Same 2.5-D streaming structure as the first version (16x16 tile marching along k, in-plane tile of w with
// halo, thread-private k-queue of w in shared memory) but the per-point work is organised to move far less
// data through shared memory and to issue far fewer instructions (round-1 ncu: 316 M shared wavefronts and
// 1.13 G warp instructions per launch made this kernel shared-memory/issue bound at 31 % DRAM throughput):
//   * ONE pass over the directions: the 2R neighbours of a component along a direction are read once and feed
//     both the adjoint first derivative and the artificial dissipation of that direction (the first version
//     read them three times: dissipation, second-partial pass, Jacobian pass);
//   * the Jacobian-transpose products are closed forms (cns_device.cuh: add_flux_jacobian_transpose_cf),
//     no NU x NU matrix is materialised;
//   * the stress / heat-flux entries of a direction are loaded right before that direction's stencil work,
//     so only one direction's viscous inputs are live at a time.
#pragma once
#include "fused_common.cuh"

namespace {

// HOT: viscous + non-composite dissipation known at compile time (the benchmark / AcousticMonopole family);
// otherwise those switches are runtime-uniform.  TYv: tile height (16 -> 256 threads, 12 -> 192 threads with a
// 168-register budget at 2 CTAs per SM).
//
// TMAQ: the arriving plane of w is fetched by the Tensor Memory Accelerator (one cp.async.bulk.tensor of a
// 16 x TY x 1 x NU box per plane, issued by thread 0 ONE PLANE AHEAD) straight into its slot of the k-queue, whose
// [NU][TY][16] layout is exactly the dense box layout; the threads only wait on the slot's mbarrier.  This removes
// the exposed load -> shared-store dependency at the top of every iteration (24 % of the stall samples of the
// plain-load variant sat on that STS) and the per-thread address arithmetic of NU loads.
template <int ND, int R, int DLO, int DN, int TLO, int TN, bool CURV, bool CLOS, bool HOT, int TYv, bool TMAQ>
__global__ void __launch_bounds__(TX * TYv, 2) k_adjoint1v2(FusedArgs a, const __grid_constant__ CUtensorMap tmW) {
  constexpr int TY = TYv, NT = TX * TYv;
  const bool COMPOSITE = HOT ? false : a.composite != 0;
  const bool viscous = HOT ? true : (a.viscous ? true : false);
  constexpr int NU = ND + 2;
  constexpr int NTAU = ND * (ND + 1) / 2;
  constexpr int W = TX + 2 * R, H = TY + 2 * R;
  constexpr int NF = NU + 2;                   // w, arc_i, arc_j
  constexpr int FA = NU;
  constexpr int RK = (ND == 3) ? R : 0;
  // ring of planes: 8 slots (wrap is a mask) when 2R+1 fits, else 2R+1 slots with a conditional wrap
  constexpr int NQ = (ND == 3) ? (2 * R + 1 <= 8 ? 8 : 2 * R + 1) : 1;
  auto ringw = [](int x) -> int {
    if constexpr ((NQ & (NQ - 1)) == 0) return x & (NQ - 1);
    else return x < 0 ? x + NQ : (x >= NQ ? x - NQ : x);
  };
  static_assert(!TMAQ || (ND == 3 && NQ == 8 && 2 * R + 1 <= 7), "TMA queue needs a free ring slot");
  extern __shared__ __align__(1024) double smem[];
  double* const WQ = smem;                                   // [NQ][NU][NT] k-queue of w (thread-private columns)
  double* const T0 = smem + (size_t)NQ * NU * NT;            // [NF][H][W]
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(T0 + (size_t)NF * H * W);   // [2]
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  int i0, j0;
  bool lastI, lastJ;
  tile_origin(blockIdx.x, a.nx, TX, i0, lastI);
  tile_origin(blockIdx.y, a.ny, TY, j0, lastJ);
  const int i = i0 + tx, j = j0 + ty;
  const bool mine = owns(i, a.nx, TX, lastI) && owns(j, a.ny, TY, lastJ);
  const bool inside = i < a.nx && j < a.ny;
  const int pij = i + a.nx * j;
  auto touches = [&](int d, int c0, int T, int n) {
    int depth = a.D[d].depth;
    if (a.dissOn) {
      depth = max(depth, a.Dd[d].depth);
      if (!COMPOSITE) depth = max(depth, max(a.Dt[d].depth + a.Dd[d].width, a.dir[d].normDepth));
    }
    return (a.dir[d].hasB0 && c0 < depth) || (a.dir[d].hasB1 && c0 + T > n - depth);
  };
  const bool dissOn = HOT ? true : a.dissOn != 0;
  const bool fastI = !CLOS || !touches(0, i0, TX, a.nx);
  const bool fastJ = !CLOS || !touches(1, j0, TY, a.ny);
  double* const tc = T0 + (ty + R) * W + tx + R;
  double* const wqc = WQ + threadIdx.x;                      // slot stride NU*NT, component stride NT

  int hcol = 0, hrow = 0, hk = 0;
  int hp = -1;
  {
    const int h = threadIdx.x;
    if (h < 2 * R * TY) {
      const int ii = h % (2 * R), row = h / (2 * R);
      const int lc = ii < R ? ii : TX + ii;
      const int gi = wrap_index(i0 - R + lc, a.dir[0]);
      const int gj = j0 + row;
      hcol = lc; hrow = row + R;
      if (gi >= 0 && gj < a.ny) { hk = 1; hp = gi + a.nx * gj; }
    } else if (h < 2 * R * TY + 2 * R * TX) {
      const int h2 = h - 2 * R * TY;
      const int col = h2 % TX, jj = h2 / TX;
      const int lr = jj < R ? jj : TY + jj;
      const int gj = wrap_index(j0 - R + lr, a.dir[1]);
      const int gi = i0 + col;
      hcol = col + R; hrow = lr;
      if (gj >= 0 && gi < a.nx) { hk = 2; hp = gi + a.nx * gj; }
    }
  }
  double* const th = T0 + hrow * W + hcol;
  const int kc0 = a.kBeg + (a.zOff + (int)blockIdx.z * a.zMul) * a.kChunk;
  const int kc1 = min(kc0 + a.kChunk, a.kEnd);
  auto wrapPlane = [&](int k) -> int {
    if (ND < 3 || !a.wrapK) return k;
    int kk = k % a.nz;
    return kk < 0 ? kk + a.nz : kk;
  };
  int ks = wrapPlane(kc0 - RK);
  int slot = 0;                    // queue slot of the arriving plane s
  const double gamma = a.pp.gamma;
  const double sigma = a.dissAmount;
  constexpr unsigned BOX_BYTES = (unsigned)(sizeof(double) * NU * NT);
  if constexpr (TMAQ) {
    if (threadIdx.x == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      mbar_fence_init();
      fence_proxy_async();
      mbar_expect_tx(&bars[0], BOX_BYTES);
      tma_load_4d(WQ, &tmW, &bars[0], i0, j0, ks + a.ghostK, 0);
    }
    __syncthreads();
  }
  int nIter = 0;

  for (int s = kc0 - RK; s < kc1 + RK; ++s, ++nIter) {
    if (ND == 3 && a.prefetch) {
      // pull the lines of the planes needed `prefetch` steps ahead into L2 (costs no registers)
      int kf = ks + a.prefetch, kq = ks - RK + a.prefetch;
      if (a.wrapK) { if (kf >= a.nz) kf -= a.nz; if (kq < 0) kq += a.nz; else if (kq >= a.nz) kq -= a.nz; }
      prefetch_streams(a, kf, a.wrapK || s + a.prefetch < a.nz + RK, kq, a.wrapK || (kq >= 0 && kq < a.nz),
                       (long)i0 + (long)a.nx * j, tx, i0 < a.nx && j < a.ny);
    }
    if constexpr (TMAQ) {
      // request plane s + 1 (its ring slot was last read two iterations ago), then wait for plane s
      if (threadIdx.x == 0 && s + 1 < kc1 + RK) {
        int kn = ks + 1;
        if (a.wrapK && kn >= a.nz) kn -= a.nz;
        unsigned long long* bar = &bars[(nIter + 1) & 1];
        fence_proxy_async();
        mbar_expect_tx(bar, BOX_BYTES);
        tma_load_4d(WQ + (size_t)ringw(slot + 1) * NU * NT, &tmW, bar, i0, j0, kn + a.ghostK, 0);
      }
      mbar_wait(&bars[nIter & 1], (unsigned)((nIter >> 1) & 1));
    } else if (inside) {
      const double* __restrict__ Wp = a.Win + ((ND == 3) ? (long)ks * a.plane : 0) + pij;
      double wv[NU];
#pragma unroll
      for (int c = 0; c < NU; ++c) wv[c] = __ldg(Wp + (size_t)c * a.cs);
#pragma unroll
      for (int c = 0; c < NU; ++c) wqc[(slot * NU + c) * NT] = wv[c];
    }
    const int p = s - RK;
    const int sp0 = ringw(slot - RK);      // slot of plane p
    int kp = ks - RK;
    if (ND == 3 && a.wrapK && kp < 0) kp += a.nz;
    if (ND == 3) {
      ++ks;
      if (a.wrapK && ks >= a.nz) ks -= a.nz;
      slot = ringw(slot + 1);
    }
    if (p < kc0) {
      if constexpr (TMAQ) __syncthreads();   // every thread has passed its wait before the barrier is re-armed
      continue;
    }
    const long poff = (ND == 3) ? (long)kp * a.plane : 0;
    const long off = poff + pij;
    // ---- in-plane tile of plane p: own point from the queue, halo (and arc lengths) from global memory
    if (inside) {
#pragma unroll
      for (int c = 0; c < NU; ++c) tc[c * H * W] = wqc[(sp0 * NU + c) * NT];
      if (!COMPOSITE && dissOn) {
        tc[(FA + 0) * H * W] = __ldg(a.arc + (size_t)0 * a.cs + off);
        tc[(FA + 1) * H * W] = __ldg(a.arc + (size_t)1 * a.cs + off);
      }
    }
    if (hk) {
      const long hoff = poff + hp;
#pragma unroll
      for (int c = 0; c < NU; ++c) th[c * H * W] = __ldg(a.Win + (size_t)c * a.cs + hoff);
      if (!COMPOSITE && dissOn) th[(FA + hk - 1) * H * W] = __ldg(a.arc + (size_t)(hk - 1) * a.cs + hoff);
    }
    // own-point inputs of the pointwise part: issued before the barrier so their latency overlaps it
    double Q[NU], M[ND * ND], jac = 0.0;
    if (mine) {
#pragma unroll
      for (int c = 0; c < NU; ++c) Q[c] = __ldg(a.Q + (size_t)c * a.cs + off);
      if (viscous) jac = __ldg(a.jac + off);
#pragma unroll
      for (int c = 0; c < ND * ND; ++c) {
        const bool diag = (c % ND) == (c / ND);
        if (CURV || diag) M[c] = __ldg(a.m + (size_t)c * a.cs + off);
      }
    }
    // 192-thread variant (168 registers): the stress / heat-flux entries are requested before the barrier too
    constexpr bool HOIST = (TYv == 12) && !CURV;
    double tqH[HOIST ? NTAU + ND : 1];
    if constexpr (HOIST) {
      if (mine && viscous) {
#pragma unroll
        for (int e = 0; e < NTAU + ND; ++e) tqH[e] = __ldg(a.tauqIn + (size_t)e * a.cs + off);
      }
    }
    __syncthreads();
    if (mine) {
      Prim<ND> sp;
      dependent<ND>(Q, gamma, sp);
      double mu = 0.0, lam = 0.0, kap = 0.0;
      if (viscous) transport<true>(sp.T, a.pp, mu, lam, kap);
      const double jm = jac * mu, jl = jac * lam, jk = jac * kap;   // rectilinear second-partial factors
      JacFactors<ND> jf;
      jac_factors<ND>(sp, gamma, viscous, a.pp.powerLaw, jf);
      double r[NU], dd[ND][ND + 1];
#pragma unroll
      for (int c = 0; c < NU; ++c) r[c] = 0.0;
#pragma unroll
      for (int jj = 0; jj < ND; ++jj)
#pragma unroll
        for (int c = 0; c < ND + 1; ++c) dd[jj][c] = 0.0;
      // curvilinear: every stress entry is needed by every direction
      double tqAll[CURV ? NTAU + ND : 1];
      if constexpr (CURV) {
        if (viscous) {
#pragma unroll
          for (int e = 0; e < NTAU + ND; ++e) tqAll[e] = __ldg(a.tauqIn + (size_t)e * a.cs + off);
        }
      }
      static_for<ND>([&](auto dI) {
        constexpr int d = dI.value;
        // ---- adjoint derivative dW and dissipation of w along direction d from ONE read of the neighbours
        double dW[NU];
        const bool fast = d == 0 ? fastI : (d == 1 ? fastJ : true);
        if (fast) {
          // dissipation weights of this point, already times -sigma: r += sum_m e[m] w(c + m - R)
          double e[2 * R + 1];
          if (dissOn) {
            if (COMPOSITE) {
#pragma unroll
              for (int m = 0; m < 2 * R + 1; ++m) e[m] = -sigma * a.Dd[d].c[m];
            } else {
#pragma unroll
              for (int m = 0; m < 2 * R + 1; ++m) e[m] = 0.0;
#pragma unroll
              for (int ea = 0; ea < TN; ++ea) {
                double arc;
                if constexpr (d == 2) {
                  int kk = kp + TLO + ea;
                  if (a.wrapK) { if (kk < 0) kk += a.nz; else if (kk >= a.nz) kk -= a.nz; }
                  arc = __ldg(a.arc + (size_t)2 * a.cs + (long)kk * a.plane + pij);
                } else {
                  arc = tc[(FA + d) * H * W + (TLO + ea) * (d == 0 ? 1 : W)];
                }
                const double w = sigma * a.Dt[d].c[ea] * arc;
#pragma unroll
                for (int eb = 0; eb < DN; ++eb) e[TLO + ea + DLO + eb + R] += w * a.Dd[d].c[eb];
              }
            }
          }
#pragma unroll
          for (int f = 0; f < NU; ++f) {
            double t = 0.0, z = dissOn ? e[R] * tc[f * H * W] : 0.0;
#pragma unroll
            for (int q = 1; q <= R; ++q) {
              double np, nm;
              if constexpr (d == 2) {
                np = wqc[(ringw(sp0 + q) * NU + f) * NT];
                nm = wqc[(ringw(sp0 - q) * NU + f) * NT];
              } else {
                constexpr int st = d == 0 ? 1 : W;
                np = tc[f * H * W + q * st];
                nm = tc[f * H * W - q * st];
              }
              t += a.D[d].c[R + q] * (np - nm);
              if (dissOn) z += e[R + q] * np + e[R - q] * nm;
            }
            dW[f] = t;
            r[f] += z;
          }
        } else {
          // closure tiles: out-of-line boundary rows (rarely taken)
          const int cd = d == 0 ? i : j, nd = d == 0 ? a.nx : a.ny;
#pragma unroll
          for (int f = 0; f < NU; ++f) {
            dW[f] = tile_line_apply<W, H>(&a.ops->D[d], cd, nd, T0, f, ty + R, tx + R, d, (d == 0 ? i0 : j0) - R);
            if (dissOn) {
              const double z = COMPOSITE
                  ? tile_line_apply<W, H>(&a.ops->Dd[d], cd, nd, T0, f, ty + R, tx + R, d, (d == 0 ? i0 : j0) - R)
                  : tile_line_dissipation<W, H>(a.ops, d, cd, T0, f, FA + d, ty + R, tx + R, (d == 0 ? i0 : j0) - R);
              r[f] -= sigma * z;
            }
          }
        }
        // ---- viscous inputs of this direction (rectilinear: row d of the stress tensor and q_d only)
        double cst[ND], chf = 0.0;
        if (viscous) {
          if constexpr (CURV) {
#pragma unroll
            for (int c = 0; c < ND; ++c) {
              double acc = 0.0;
#pragma unroll
              for (int l = 0; l < ND; ++l)
                acc = (l == 0) ? M[ND * d] * tqAll[tau_index<ND>(0, c)] : acc + M[l + ND * d] * tqAll[tau_index<ND>(l, c)];
              cst[c] = acc;
            }
#pragma unroll
            for (int l = 0; l < ND; ++l) chf = (l == 0) ? M[ND * d] * tqAll[NTAU] : chf + M[l + ND * d] * tqAll[NTAU + l];
          } else {
            if constexpr (HOIST) {
#pragma unroll
              for (int c = 0; c < ND; ++c) cst[c] = tqH[tau_index<ND>(d, c)];
              chf = tqH[NTAU + d];
            } else {
#pragma unroll
              for (int c = 0; c < ND; ++c) cst[c] = __ldg(a.tauqIn + (size_t)tau_index<ND>(d, c) * a.cs + off);
              chf = __ldg(a.tauqIn + (size_t)(NTAU + d) * a.cs + off);
            }
          }
        }
        // ---- pointwise products of this direction
        if (viscous) {
          if constexpr (CURV) {
#pragma unroll
            for (int jj = 0; jj < ND; ++jj)
              add_second_partial_transpose<ND>(sp.u, mu, lam, kap, jac, &M[ND * d], &M[ND * jj], &dW[1], dd[jj]);
          } else {
            static_for<ND>([&](auto jj) {
              add_second_partial_transpose_rect_f<ND, d, jj.value>(sp.u, jm, jl, jk, M[d + ND * d] * M[jj.value + ND * jj.value],
                                                                   &dW[1], dd[jj.value]);
            });
            const double md = M[d + ND * d];
#pragma unroll
            for (int c = 0; c < ND; ++c) cst[c] *= md;
            chf *= md;
          }
        }
        add_flux_jacobian_transpose_cf<ND, !CURV, d>(sp, jf, &M[ND * d], gamma, viscous, cst, chf, dW, r);
        asm volatile("" ::: "memory");   // scheduling fence: one direction's temporaries live at a time
      });
      if (viscous) {
#pragma unroll
        for (int jj = 0; jj < ND; ++jj)
#pragma unroll
          for (int c = 0; c < ND + 1; ++c) a.diffOut[(size_t)(c + (NU - 1) * jj) * a.cs + off] = dd[jj][c];
      }
#pragma unroll
      for (int c = 0; c < NU; ++c) a.rhs[(size_t)c * a.cs + off] = r[c];
    }
    __syncthreads();
  }
}

template <int ND, int R, int DLO, int DN, int TLO, int TN, bool CURV, bool CLOS, bool HOT, int TYv, bool TMAQ>
int launchAdj1v2(const FusedArgs& a, int nChunks, cudaStream_t st, const CUtensorMap* tmW) {
  constexpr int NF = (ND + 2) + 2;
  constexpr int NQ = (ND == 3) ? (2 * R + 1 <= 8 ? 8 : 2 * R + 1) : 1;
  constexpr int NT = TX * TYv;
  static_assert(NT >= 2 * R * (TX + TYv), "one halo point per thread");
  const size_t smem = sizeof(double) * ((size_t)NF * (TYv + 2 * R) * (TX + 2 * R) + (size_t)NQ * (ND + 2) * NT) + 16;
  auto kern = k_adjoint1v2<ND, R, DLO, DN, TLO, TN, CURV, CLOS, HOT, TYv, TMAQ>;
  static int configuredDevice = -1;      // per template instantiation and device
  int device = 0;
  cudaGetDevice(&device);
  if (configuredDevice != device) {
    MG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configuredDevice = device;
  }
  const dim3 grid((a.nx + TX - 1) / TX, (a.ny + TYv - 1) / TYv, nChunks);
  CUtensorMap none;
  std::memset(&none, 0, sizeof(none));
  mg_profile_begin("adjoint1");
  kern<<<grid, NT, smem, st>>>(a, tmW ? *tmW : none);
  mg_profile_end();
  MG_CUDA(cudaGetLastError());
  mg_count_launches(1);
  return 0;
}

template <int ND, int R, int DLO, int DN, int TLO, int TN, bool HOT, int TYv, bool TMAQ = false>
int dispatchAdj1v2(const FusedArgs& a, int nChunks, cudaStream_t st, const CUtensorMap* tmW = nullptr) {
  const bool clos = has_closures(a);
  return a.curvilinear ? (clos ? launchAdj1v2<ND, R, DLO, DN, TLO, TN, true, true, HOT, TYv, TMAQ>(a, nChunks, st, tmW)
                               : launchAdj1v2<ND, R, DLO, DN, TLO, TN, true, false, HOT, TYv, TMAQ>(a, nChunks, st, tmW))
                       : (clos ? launchAdj1v2<ND, R, DLO, DN, TLO, TN, false, true, HOT, TYv, TMAQ>(a, nChunks, st, tmW)
                               : launchAdj1v2<ND, R, DLO, DN, TLO, TN, false, false, HOT, TYv, TMAQ>(a, nChunks, st, tmW));
}

}  // namespace
