// Forward sweep B with the artificial dissipation folded in: the forward stage is TWO sweeps over the grid
// (sweep A: state update; this kernel: everything else), the 448 B/point contract of SURVEY.md section 8(d).
//
//   computeRhsForward (reference src/RhsHelperImpl.f90:254-354) + addDissipation(FORWARD) (:10-87)
//   + x 1/J (src/RegionImpl.f90:1969-1974) + substepForwardRK4 (src/RK4IntegratorImpl.f90:65-162)
//
// reads Q, tau/q, metrics, 1/J, arc lengths and the RK buffers; writes the RK accumulator and the next Q (or the
// RHS).  No dissipation term travels through HBM any more (round 1: +144 B/point and a third kernel).
//
// 2.5-D streaming: a CTA owns a 16 x TY tile and marches along k.
//   * in-plane: shared tiles of Q (with halo, feeds the in-plane dissipation), of the arc lengths and of the
//     contravariant fluxes F1 / F2 (halo columns / rows evaluated by the halo threads);
//   * k direction, fluxes: ring of running sums of div F for the 2R+1 planes in flight (accumulate form:
//     c_q F3(s) is scattered to planes s -/+ q on arrival) - thread-private columns of shared memory;
//   * k direction, dissipation Dt(-a Dd q): g = Dd q is accumulated in a REGISTER ring of DN-1 partial sums as
//     the planes arrive; a completed g(r), times -a(r), is scattered with the Dt weights into the SAME running
//     sums the flux scatter touches (no extra shared-memory traffic, no second k-queue).
// TY = 12 (192 threads, 2 CTAs/SM) or 8 (128 threads, 3 CTAs/SM): 384 points in flight per SM so that a
// 168-register budget holds the g ring, the emit inputs and the flux evaluation without spilling.
#pragma once
#include "fused_common.cuh"

namespace {

template <int ND, int R, int DLO, int DN, int TLO, int TN, bool CURV, bool CLOS, bool HOT, int TYv>
__global__ void __launch_bounds__(TX * TYv, (TYv == 8 && R < 4) ? 3 : 2) k_sweepBD(FusedArgs a) {
  constexpr int TY = TYv, NT = TX * TYv;
  constexpr int NU = ND + 2;
  constexpr int W = TX + 2 * R, H = TY + 2 * R;
  constexpr int RK = (ND == 3) ? R : 0;
  constexpr int NQ = 2 * RK + 1;
  constexpr int ALLDIRS = (1 << ND) - 1;
  constexpr int FA = NU;                              // arc-length fields follow Q in the tile block
  constexpr int HI_D = DLO + DN - 1;                  // g(r) is complete once q(r + HI_D) has arrived
  constexpr int NG = DN - 1;                          // partial g sums in flight
  static_assert(HI_D + (TLO + TN - 1) == R, "dissipation support must match the first-derivative radius");
  constexpr int NHALO = 2 * R * TY + 2 * R * TX;
  constexpr int NH = (NHALO + NT - 1) / NT;
  const bool dissOn = HOT ? true : a.dissOn != 0;
  const bool COMPOSITE = HOT ? false : a.composite != 0;
  extern __shared__ double smem[];
  double* const QT = smem;                            // [NU + 2][H][W]  Q, arc_i, arc_j
  double* const F1 = QT + (size_t)(NU + 2) * H * W;   // [NU][TY][W]     contravariant flux along xi
  double* const F2 = F1 + (size_t)NU * TY * W;        // [NU][H][TX]     contravariant flux along eta
  double* const ACC = F2 + (size_t)NU * H * TX;       // [NQ][NU][NT]    running sums (div F - sigma Diss)
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  int i0, j0;
  bool lastI, lastJ;
  tile_origin(blockIdx.x, a.nx, TX, i0, lastI);
  tile_origin(blockIdx.y, a.ny, TY, j0, lastJ);
  const int i = i0 + tx, j = j0 + ty;
  const bool mine = owns(i, a.nx, TX, lastI) && owns(j, a.ny, TY, lastJ);
  const bool inside = i < a.nx && j < a.ny;
  const long pij = (long)i + (long)a.nx * j;
  auto touchesD = [&](int d, int c0, int T, int n) {   // first derivative closures
    return (a.D[d].hasB0 && c0 < a.D[d].depth) || (a.D[d].hasB1 && c0 + T > n - a.D[d].depth);
  };
  auto touchesDiss = [&](int d, int c0, int T, int n) {
    if (!dissOn) return false;
    int depth = a.Dd[d].depth;
    if (!COMPOSITE) depth = max(depth, max(a.Dt[d].depth + a.Dd[d].width, a.dir[d].normDepth));
    return (a.dir[d].hasB0 && c0 < depth) || (a.dir[d].hasB1 && c0 + T > n - depth);
  };
  const bool fastI = !CLOS || !touchesD(0, i0, TX, a.nx), fastJ = !CLOS || !touchesD(1, j0, TY, a.ny);
  const bool fastDI = !CLOS || !touchesDiss(0, i0, TX, a.nx), fastDJ = !CLOS || !touchesDiss(1, j0, TY, a.ny);
  double* const qc = QT + (ty + R) * W + tx + R;   // own point in the Q tile; field stride H*W
  double* const f1c = F1 + ty * W + tx + R;        // component stride TY*W
  double* const f2c = F2 + (ty + R) * TX + tx;     // component stride H*TX, row stride TX
  double* const acc = ACC + threadIdx.x;           // slot stride NU*NT, component stride NT

  // halo points handled by this thread: kind 0 none, 1 xi-halo, 2 eta-halo
  int hkind[NH], hq[NH], hf[NH];
  long hp[NH];
#pragma unroll
  for (int n = 0; n < NH; ++n) {
    const int h = threadIdx.x + n * NT;
    hkind[n] = 0; hq[n] = 0; hf[n] = 0; hp[n] = 0;
    if (h < 2 * R * TY) {
      const int ii = h % (2 * R), row = h / (2 * R);
      const int lc = ii < R ? ii : TX + ii;
      const int gi = wrap_index(i0 - R + lc, a.dir[0]);
      const int gj = j0 + row;
      hq[n] = (row + R) * W + lc;
      hf[n] = row * W + lc;
      if (gi >= 0 && gj < a.ny) { hkind[n] = 1; hp[n] = (long)gi + (long)a.nx * gj; }
    } else if (h < NHALO) {
      const int h2 = h - 2 * R * TY;
      const int col = h2 % TX, jj = h2 / TX;
      const int lr = jj < R ? jj : TY + jj;
      const int gj = wrap_index(j0 - R + lr, a.dir[1]);
      const int gi = i0 + col;
      hq[n] = lr * W + col + R;
      hf[n] = lr * TX + col;
      if (gj >= 0 && gi < a.nx) { hkind[n] = 2; hp[n] = (long)gi + (long)a.nx * gj; }
    }
  }
  const int kc0 = a.kBeg + (a.zOff + (int)blockIdx.z * a.zMul) * a.kChunk;
  const int kc1 = min(kc0 + a.kChunk, a.kEnd);
  auto wrapPlane = [&](int k) -> int {
    if (ND < 3 || !a.wrapK) return k;
    int kk = k % a.nz;
    return kk < 0 ? kk + a.nz : kk;
  };
  auto ring = [&](int x) -> int { return x < 0 ? x + NQ : (x >= NQ ? x - NQ : x); };
  int ks = wrapPlane(kc0 - RK);                    // storage plane of the arriving plane s
  int kp = wrapPlane(kc0 - 2 * RK);                // storage plane of the output plane p = s - RK
  int kg = wrapPlane(kc0 - RK - HI_D);             // storage plane of the g that completes at arrival s
  int slot = 0;                                    // ring slot of plane s
  if (inside) {
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int c = 0; c < NU; ++c) acc[(q * NU + c) * NT] = 0.0;
  }
  const double sigma = a.dissAmount;
  // register ring of the partial sums g(s - HI_D + 1 + n), n = 0 .. NG-1, of Dd q along k
  double G[NG > 0 ? NG : 1][NU];
#pragma unroll
  for (int n = 0; n < NG; ++n)
#pragma unroll
    for (int c = 0; c < NU; ++c) G[n][c] = 0.0;

  for (int s = kc0 - RK; s < kc1 + RK; ++s) {
    const long soff = (ND == 3) ? (long)ks * a.plane : 0;
    const bool planeActive = s >= kc0 && s < kc1;
    const bool emitNow = (ND == 3) ? (s - RK >= kc0 && mine) : false;
    if (ND == 3 && a.prefetch) {
      // pull the lines of the planes needed `prefetch` steps ahead into L2 (costs no registers)
      int kf = ks + a.prefetch, kq = kp + a.prefetch;
      if (a.wrapK) { if (kf >= a.nz) kf -= a.nz; if (kq < 0) kq += a.nz; else if (kq >= a.nz) kq -= a.nz; }
      PfItems<2, false>::issue_stateless(a, kf, a.wrapK || s + a.prefetch < a.nz + RK, kq, a.wrapK || (kq >= 0 && kq < a.nz),
                                         pij - tx, tx, i0 < a.nx && j < a.ny);
    }
    // ---- inputs of the output plane p = s - RK (consumed by the emit below; issued first)
    double ejac = 0.0, vb1[NU], vb2[NU], arcg = 0.0;
    auto emit_load = [&](int kpl) {
      const long off = ((ND == 3) ? (long)kpl * a.plane : 0) + pij;
      ejac = __ldg(a.jac + off);
#pragma unroll
      for (int c = 0; c < NU; ++c) { vb1[c] = 0.0; vb2[c] = 0.0; }
      // RK buffers: one uniform branch on the stage, not one per component
      if (a.fuseRk) {
        if (a.stage == 1) {
#pragma unroll
          for (int c = 0; c < NU; ++c) vb1[c] = __ldg(a.Q + (size_t)c * a.cs + off);
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
    };
    auto emit = [&](int kpl, const double* total) {
      const long off = ((ND == 3) ? (long)kpl * a.plane : 0) + pij;
      double r[NU];
#pragma unroll
      for (int c = 0; c < NU; ++c) r[c] = (0.0 - total[c]) * ejac;
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
    };
    if (emitNow) emit_load(kp);
    // arc length of the g that completes now (planes below kc0 + TLO only feed outputs of another chunk)
    if (ND == 3 && dissOn && !COMPOSITE && inside && s - HI_D >= kc0 + TLO)
      arcg = __ldg(a.arc + (size_t)2 * a.cs + (long)kg * a.plane + pij);
    // ---- arrival of plane s: the halo points' inputs are requested first so that their latency overlaps the
    // own-point flux evaluation
    RawPoint<ND> rawH[NH];
    double arcH[NH];
    if (planeActive) {
#pragma unroll
      for (int n = 0; n < NH; ++n) {
        arcH[n] = 0.0;
        if (hkind[n] == 1) load_raw<ND, 1, CURV>(a, soff + hp[n], rawH[n]);
        else if (hkind[n] == 2) load_raw<ND, 2, CURV>(a, soff + hp[n], rawH[n]);
        if (hkind[n] && dissOn && !COMPOSITE) arcH[n] = __ldg(a.arc + (size_t)(hkind[n] - 1) * a.cs + soff + hp[n]);
      }
    }
    // ---- own point
    double f3[NU], last[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) { f3[c] = 0.0; last[c] = 0.0; }
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
          if (dissOn) qc[c * H * W] = raw.Q[c];
        }
        if (dissOn && !COMPOSITE) {
          qc[(FA + 0) * H * W] = __ldg(a.arc + (size_t)0 * a.cs + soff + pij);
          qc[(FA + 1) * H * W] = __ldg(a.arc + (size_t)1 * a.cs + soff + pij);
        }
      } else {
        load_raw<ND, (ND == 3 ? 4 : 0), CURV>(a, soff + pij, raw);
        fluxes_from_raw<ND, (ND == 3 ? 4 : 0), CURV>(a, raw, Fh);
      }
      if constexpr (ND == 3) {
        // k direction: c_q F3(s) to planes s-q (+) and s+q (-); the dissipation contributions ride on the
        // same read-modify-writes.  delta[m] = what plane s + m - R receives from this arrival.
#pragma unroll
        for (int c = 0; c < NU; ++c) f3[c] = Fh[ND - 1][c];
        // q(s) completes g(s - HI_D) and feeds the NG partial sums behind it (non-composite dissipation)
        double gg[NU];
#pragma unroll
        for (int c = 0; c < NU; ++c) gg[c] = 0.0;
        if (dissOn && !COMPOSITE) {
#pragma unroll
          for (int c = 0; c < NU; ++c) {
            const double g = ((NG > 0) ? G[0][c] : 0.0) + a.Dd[2].c[DN - 1] * raw.Q[c];
#pragma unroll
            for (int n = 0; n + 1 < NG; ++n) G[n][c] = G[n + 1][c] + a.Dd[2].c[DN - 2 - n] * raw.Q[c];
            if (NG > 0) G[NG - 1][c] = a.Dd[2].c[0] * raw.Q[c];
            gg[c] = sigma * arcg * g;             // -sigma x (-arc g)
          }
        }
        // what plane s + m - R receives from this arrival: flux scatter -/+ c_q F3(s), Dt weights of the
        // completed g (plane offset m = R - HI_D - TLO - ea), or the composite operator applied to q(s).
        // m = 2R initialises the slot of plane s + R, m = 0 completes plane s - R (kept in registers).
        static_for<2 * R + 1>([&](auto mI) {
          constexpr int m = mI.value;
          constexpr int ea = R - HI_D - TLO - m;
          double v[NU];
#pragma unroll
          for (int c = 0; c < NU; ++c) {
            double t = 0.0;
            if (m < R) t = a.D[2].c[2 * R - m] * f3[c];
            if (m > R) t = 0.0 - a.D[2].c[m] * f3[c];
            if (dissOn) {
              if (COMPOSITE) t -= sigma * a.Dd[2].c[2 * R - m] * raw.Q[c];
              else if (ea >= 0 && ea < TN) t += a.Dt[2].c[ea >= 0 && ea < TN ? ea : 0] * gg[c];
            }
            v[c] = t;
          }
          if constexpr (m == 0) {
#pragma unroll
            for (int c = 0; c < NU; ++c) last[c] = v[c];
          } else if constexpr (m == 2 * R) {
            const int sl = ring(slot + R);
#pragma unroll
            for (int c = 0; c < NU; ++c) acc[(sl * NU + c) * NT] = v[c];
          } else {
            if (m != R || dissOn) {
              const int sl = ring(slot + m - R);
#pragma unroll
              for (int c = 0; c < NU; ++c) acc[(sl * NU + c) * NT] += v[c];
            }
          }
        });
      } else {
#pragma unroll
        for (int c = 0; c < NU; ++c) acc[c * NT] = 0.0;
      }
    }
    // ---- halo points of this thread: Q into the tile, the flux along the halo direction into F1 / F2
    if (planeActive) {
#pragma unroll
      for (int n = 0; n < NH; ++n) {
        if (!hkind[n]) continue;
        const RawPoint<ND>& raw = rawH[n];
        double Fh[ND][NU];
        if (hkind[n] == 1) {
          fluxes_from_raw<ND, 1, CURV>(a, raw, Fh);
#pragma unroll
          for (int c = 0; c < NU; ++c) F1[c * TY * W + hf[n]] = Fh[0][c];
        } else {
          fluxes_from_raw<ND, 2, CURV>(a, raw, Fh);
#pragma unroll
          for (int c = 0; c < NU; ++c) F2[c * H * TX + hf[n]] = Fh[1][c];
        }
        if (dissOn) {
#pragma unroll
          for (int c = 0; c < NU; ++c) QT[c * H * W + hq[n]] = raw.Q[c];
          if (!COMPOSITE) QT[(FA + hkind[n] - 1) * H * W + hq[n]] = arcH[n];
        }
      }
    }
    // ---- output plane p = s - RK (3-D): everything but this arrival's share is already in the ring
    if constexpr (ND == 3) {
      if (emitNow) {
        const int sp0 = ring(slot - RK);
        double total[NU];
#pragma unroll
        for (int c = 0; c < NU; ++c) total[c] = acc[(sp0 * NU + c) * NT] + last[c];
        emit(kp, total);
      }
    }
    __syncthreads();
    if (planeActive && mine) {
      double r[NU];
      // in-plane divergence of the fluxes
#pragma unroll
      for (int c = 0; c < NU; ++c) {
        double r1;
        if (fastI) {
          r1 = 0.0;
#pragma unroll
          for (int q = 1; q <= R; ++q) r1 += a.D[0].c[R + q] * (f1c[c * TY * W + q] - f1c[c * TY * W - q]);
        } else {
          r1 = strided_line_apply<W>(&a.ops->D[0], i, a.nx, F1 + ((size_t)c * TY + ty) * W, 1, i0 - R);
        }
        if (fastJ) {
          double r2 = 0.0;
#pragma unroll
          for (int q = 1; q <= R; ++q) r2 += a.D[1].c[R + q] * (f2c[c * H * TX + q * TX] - f2c[c * H * TX - q * TX]);
          r1 += r2;
        } else {
          r1 += strided_line_apply<H>(&a.ops->D[1], j, a.ny, F2 + (size_t)c * H * TX + tx, TX, j0 - R);
        }
        r[c] = r1;
      }
      // in-plane artificial dissipation (reference src/RhsHelperImpl.f90:58-81), times -sigma
      if (dissOn) {
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const bool fast = d == 0 ? fastDI : fastDJ;
          const int st = d == 0 ? 1 : W;
          if (fast) {
            double e[2 * R + 1];
            if (COMPOSITE) {
#pragma unroll
              for (int m = 0; m < 2 * R + 1; ++m) e[m] = -sigma * a.Dd[d].c[m];
            } else {
#pragma unroll
              for (int m = 0; m < 2 * R + 1; ++m) e[m] = 0.0;
#pragma unroll
              for (int ea = 0; ea < TN; ++ea) {
                const double w = sigma * a.Dt[d].c[ea] * qc[(FA + d) * H * W + (TLO + ea) * st];
#pragma unroll
                for (int eb = 0; eb < DN; ++eb) e[TLO + ea + DLO + eb + R] += w * a.Dd[d].c[eb];
              }
            }
#pragma unroll
            for (int c = 0; c < NU; ++c) {
              double z = 0.0;
#pragma unroll
              for (int m = 0; m < 2 * R + 1; ++m) z += e[m] * qc[c * H * W + (m - R) * st];
              r[c] += z;
            }
          } else {
            const int cd = d == 0 ? i : j, nd = d == 0 ? a.nx : a.ny;
#pragma unroll
            for (int c = 0; c < NU; ++c) {
              const double z = COMPOSITE
                  ? tile_line_apply<W, H>(&a.ops->Dd[d], cd, nd, QT, c, ty + R, tx + R, d, (d == 0 ? i0 : j0) - R)
                  : tile_line_dissipation<W, H>(a.ops, d, cd, QT, c, FA + d, ty + R, tx + R, (d == 0 ? i0 : j0) - R);
              r[c] -= sigma * z;
            }
          }
        }
      }
      if constexpr (ND == 3) {
#pragma unroll
        for (int c = 0; c < NU; ++c) acc[(slot * NU + c) * NT] += r[c];
      } else {
        emit_load(0);
        emit(0, r);
      }
    }
    __syncthreads();
    if (ND == 3) {
      ++ks; ++kp; ++kg;
      if (a.wrapK) { if (ks >= a.nz) ks -= a.nz; if (kp >= a.nz) kp -= a.nz; if (kg >= a.nz) kg -= a.nz; }
      if (++slot >= NQ) slot = 0;
    }
  }
}

template <int ND, int R, int DLO, int DN, int TLO, int TN, bool CURV, bool CLOS, bool HOT, int TYv>
int launchBD(const FusedArgs& a, int nChunks, cudaStream_t st) {
  constexpr int NU = ND + 2, NT = TX * TYv;
  constexpr int NQ = (ND == 3) ? 2 * R + 1 : 1;
  constexpr int W = TX + 2 * R, H = TYv + 2 * R;
  const size_t smem = sizeof(double) * ((size_t)(NU + 2) * H * W + (size_t)NU * TYv * W + (size_t)NU * H * TX +
                                        (size_t)NQ * NU * NT);
  auto kern = k_sweepBD<ND, R, DLO, DN, TLO, TN, CURV, CLOS, HOT, TYv>;
  static int configuredDevice = -1;      // per template instantiation and device
  int device = 0;
  cudaGetDevice(&device);
  if (configuredDevice != device) {
    MG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configuredDevice = device;
  }
  const dim3 grid((a.nx + TX - 1) / TX, (a.ny + TYv - 1) / TYv, nChunks);
  mg_profile_begin("sweepB");
  kern<<<grid, NT, smem, st>>>(a);
  mg_profile_end();
  MG_CUDA(cudaGetLastError());
  mg_count_launches(1);
  return 0;
}

template <int ND, int R, int DLO, int DN, int TLO, int TN, bool HOT, int TYv>
int dispatchBD(const FusedArgs& a, int nChunks, cudaStream_t st) {
  const bool clos = has_closures(a);
  return a.curvilinear ? (clos ? launchBD<ND, R, DLO, DN, TLO, TN, true, true, HOT, TYv>(a, nChunks, st)
                               : launchBD<ND, R, DLO, DN, TLO, TN, true, false, HOT, TYv>(a, nChunks, st))
                       : (clos ? launchBD<ND, R, DLO, DN, TLO, TN, false, true, HOT, TYv>(a, nChunks, st)
                               : launchBD<ND, R, DLO, DN, TLO, TN, false, false, HOT, TYv>(a, nChunks, st));
}

}  // namespace
