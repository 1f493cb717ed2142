// Host side of the StencilOperator handle: scheme tables, update (process-grid placement),
// getAdjoint.  Mirrors t_StencilOperator%setup/update/getAdjoint
// (reference: src/StencilOperatorImpl.f90:1111-2193, :2195-2256, :2285-2372).
#include <cmath>
#include <cstring>
#include <initializer_list>
#include <map>

#include "mg_common.h"

namespace {

// rhsBoundary arrays are held as [point i][row m] here (the reference's (boundaryWidth, boundaryDepth)
// layout) and transposed into MgDevOp::b1[m][i] by finish().
struct Builder {
  int sym = MG_SYMMETRIC, iw = 0, bw = 0, bd = 0, lo = 0, nInt = 0;
  double interior[MG_MAX_INTERIOR] = {0};
  double b1[MG_MAX_BWIDTH][MG_MAX_BDEPTH] = {{0}};
  double b2[MG_MAX_BWIDTH][MG_MAX_BDEPTH] = {{0}};
  double norm[MG_MAX_BDEPTH];
  bool explicitB2 = false;

  void allocate(int s, int iw_, int bw_, int bd_) {
    sym = s; iw = iw_; bw = bw_; bd = bd_;
    if (sym != MG_ASYMMETRIC) { lo = -(iw / 2); nInt = 2 * (iw / 2) + 1; }
    for (int m = 0; m < MG_MAX_BDEPTH; ++m) norm[m] = 1.0;
  }
  void setInteriorHalf(std::initializer_list<double> half, bool hasCenter = false, double center = 0.0) {
    int m = 1;
    for (double c : half) {
      interior[m - lo] = c;
      interior[-m - lo] = (sym == MG_SKEW_SYMMETRIC) ? -c : c;
      ++m;
    }
    if (hasCenter) interior[0 - lo] = center;
  }
  void setInteriorFull(int lo_, std::initializer_list<double> v) {
    lo = lo_; nInt = (int)v.size();
    int k = 0;
    for (double c : v) interior[k++] = c;
  }
  void setNorm(const std::vector<double>& v) { for (size_t m = 0; m < v.size(); ++m) norm[m] = v[m]; }
  // rhsBoundary1(start:start+len-1, row) = vals, 1-based like the reference
  void row(int start, int r, std::initializer_list<double> vals) {
    int i = start - 1;
    for (double v : vals) b1[i++][r - 1] = v;
  }
  void row2(int start, int r, std::initializer_list<double> vals) {
    explicitB2 = true;
    int i = start - 1;
    for (double v : vals) b2[i++][r - 1] = v;
  }
  void put(int i, int j, double v) { b1[i - 1][j - 1] = v; }
  // rhsBoundary1(i, m0:m0+len-1) = vals, 0-based
  void b1Across(int i, int m0, std::initializer_list<double> vals) { for (double v : vals) b1[i][m0++] = v; }
  void b2Across(int i, int m0, std::initializer_list<double> vals) {
    explicitB2 = true;
    for (double v : vals) b2[i][m0++] = v;
  }
  void scaleInterior(double d) { for (int k = 0; k < nInt; ++k) interior[k] /= d; }
  void scaleBoundary1(double d) {
    for (int i = 0; i < bw; ++i) for (int m = 0; m < bd; ++m) b1[i][m] /= d;
  }
};

#include "stencil_tables.inc"

typedef void (*TableFn)(Builder&);
const std::map<std::string, TableFn>& tables() {
  static const std::map<std::string, TableFn> t = {
      {"null matrix", t_null},
      {"Standard 5-point filter", t_std5_filter},
      {"DRP 9-point filter", t_drp9_filter},
      {"SBP 1-2 first derivative", t_12_first},
      {"SBP 1-2 second derivative", t_12_second},
      {"SBP 1-2 composite dissipation", t_12_compdiss},
      {"SBP 2-4 first derivative", t_24_first},
      {"SBP 2-4 second derivative", t_24_second},
      {"SBP 2-4 composite dissipation", t_24_compdiss},
      {"SBP 2-4 dissipation", t_24_diss},
      {"SBP 2-4 dissipation transpose", t_24_disst},
      {"SBP 3-6 first derivative", t_36_first},
      {"SBP 3-6 second derivative", t_36_second},
      {"SBP 3-6 composite dissipation", t_36_compdiss},
      {"SBP 3-6 dissipation", t_36_diss},
      {"SBP 3-6 dissipation transpose", t_36_disst},
      {"SBP 4-8 first derivative", t_48_first},
      {"SBP 4-8 composite dissipation", t_48_compdiss},
      {"SBP 4-8 dissipation", t_48_diss},
      {"SBP 4-8 dissipation transpose", t_48_disst},
  };
  return t;
}

// Transpose the builder's [i][m] arrays into the device op and mirror the right boundary
// (reference :2181-2191).
void finish(const Builder& b, MgDevOp* op) {
  std::memset(op, 0, sizeof(*op));
  op->symmetryType = b.sym;
  op->interiorWidth = b.iw;
  op->boundaryWidth = b.bw;
  op->boundaryDepth = b.bd;
  op->normDepth = b.bd;
  op->lo = b.lo;
  op->nInterior = b.nInt;
  for (int k = 0; k < b.nInt; ++k) op->interior[k] = b.interior[k];
  for (int m = 0; m < MG_MAX_BDEPTH; ++m) op->normBoundary[m] = b.norm[m];
  for (int i = 0; i < b.bw; ++i)
    for (int m = 0; m < b.bd; ++m) {
      op->b1[m][i] = b.b1[i][m];
      if (b.sym == MG_SYMMETRIC) op->b2[m][i] = b.b1[b.bw - 1 - i][m];
      else if (b.sym == MG_SKEW_SYMMETRIC) op->b2[m][i] = -b.b1[b.bw - 1 - i][m];
      else op->b2[m][i] = b.b2[i][m];
    }
}

}  // namespace

int mg_stencil_create_impl(const char* scheme, mg_stencil** out) {
  if (!scheme || !out) MG_FAIL("mg_stencil_create: null argument");
  auto it = tables().find(scheme);
  if (it == tables().end()) MG_FAIL(std::string("mg_stencil_create: unknown stencil scheme '") + scheme + "'");
  Builder b;
  it->second(b);
  auto* s = new mg_stencil();
  s->scheme = scheme;
  finish(b, &s->op);
  *out = s;
  return 0;
}

// updateOperator (reference :2195-2256); the Cartesian communicator is replaced by explicit
// process-grid dims / coordinates / periodicity.
int mg_stencil_update_impl(mg_stencil* s, int direction, const int procDims[3], const int procCoords[3],
                           const int periodic[3], int overlap) {
  if (!s) MG_FAIL("mg_stencil_update: null handle");
  if (direction < 1 || direction > 3) MG_FAIL("mg_stencil_update: direction must be 1, 2 or 3");
  const int d = direction - 1;
  if (procDims[d] <= 0 || procCoords[d] < 0 || procCoords[d] >= procDims[d])
    MG_FAIL("mg_stencil_update: invalid process grid");
  s->direction = direction;
  s->procDim = procDims[d];
  s->procCoord = procCoords[d];
  s->isPeriodic = periodic[d] ? 1 : 0;
  const bool first = s->procCoord == 0, last = s->procCoord == s->procDim - 1;
  MgDevOp& op = s->op;
  op.hasDomainBoundary[0] = first && !s->isPeriodic;
  op.hasDomainBoundary[1] = last && !s->isPeriodic;
  op.nGhost[0] = op.nGhost[1] = op.interiorWidth / 2;
  if (!s->isPeriodic && first) op.nGhost[0] = 0;
  if (!s->isPeriodic && last) op.nGhost[1] = 0;
  op.periodicOffset[0] = op.periodicOffset[1] = 0;
  if (s->isPeriodic && overlap) {
    if (first) op.periodicOffset[1] = 1;
    if (last) op.periodicOffset[0] = 1;
  }
  s->dirty = true;
  return 0;
}

// getAdjointOperator (reference :2285-2372): H^{-1} A^T H restricted to the closure block.
int mg_stencil_get_adjoint_impl(const mg_stencil* s, mg_stencil** out) {
  if (!s || !out) MG_FAIL("mg_stencil_get_adjoint: null argument");
  const MgDevOp& o = s->op;
  if (o.symmetryType != MG_SYMMETRIC && o.symmetryType != MG_SKEW_SYMMETRIC)
    MG_FAIL("mg_stencil_get_adjoint: operator must be symmetric or skew-symmetric");
  if (o.interiorWidth <= 0 || o.boundaryWidth <= 0 || o.boundaryDepth <= 0)
    MG_FAIL("mg_stencil_get_adjoint: empty operator");
  const int h = o.interiorWidth / 2;
  const int bd = o.boundaryWidth, bw = o.boundaryWidth + h, nb = o.boundaryDepth;
  if (bw > MG_MAX_BWIDTH || bd > MG_MAX_BDEPTH) MG_FAIL("mg_stencil_get_adjoint: closure too large");
  auto* a = new mg_stencil();
  a->scheme = s->scheme + " (adjoint)";
  MgDevOp& ao = a->op;
  std::memset(&ao, 0, sizeof(ao));
  ao.symmetryType = o.symmetryType;
  ao.interiorWidth = o.interiorWidth;
  ao.boundaryDepth = bd;
  ao.boundaryWidth = bw;
  ao.lo = o.lo;
  ao.nInterior = o.nInterior;
  ao.normDepth = nb;
  for (int i = -h; i <= h; ++i) ao.interior[i - ao.lo] = o.interior[-i - o.lo];
  for (int m = 0; m < MG_MAX_BDEPTH; ++m) ao.normBoundary[m] = m < nb ? o.normBoundary[m] : 1.0;
  // A1[i][m]: coefficient of point i in adjoint row m (reference layout rhsBoundary1(i, m))
  static thread_local double A1[MG_MAX_BWIDTH][MG_MAX_BDEPTH];
  std::memset(A1, 0, sizeof(A1));
  for (int i = 0; i < nb; ++i)
    for (int m = 0; m < bd; ++m) A1[i][m] = o.b1[i][m];   // transpose(rhsBoundary1): o.b1[row i][point m]
  for (int i = nb + 1; i <= o.boundaryWidth + h; ++i)      // 1-based rows of the original interior
    for (int j = -h; j <= h; ++j) {
      if (i + j > o.boundaryWidth) break;
      A1[i - 1][i + j - 1] = o.interior[j - o.lo];
    }
  for (int i = 0; i < bw; ++i)
    for (int m = 0; m < nb; ++m) A1[i][m] = A1[i][m] / o.normBoundary[m];
  for (int m = 0; m < bd; ++m)
    for (int i = 0; i < nb; ++i) A1[i][m] = A1[i][m] * o.normBoundary[i];
  for (int i = 0; i < bw; ++i)
    for (int m = 0; m < bd; ++m) {
      ao.b1[m][i] = A1[i][m];
      ao.b2[m][i] = (ao.symmetryType == MG_SYMMETRIC ? 1.0 : -1.0) * A1[bw - 1 - i][m];
    }
  *out = a;
  return 0;
}

int mg_stencil_negate_impl(mg_stencil* s) {
  MgDevOp& o = s->op;
  for (int k = 0; k < o.nInterior; ++k) o.interior[k] = -o.interior[k];
  for (int m = 0; m < MG_MAX_BDEPTH; ++m)
    for (int i = 0; i < MG_MAX_BWIDTH; ++i) { o.b1[m][i] = -o.b1[m][i]; o.b2[m][i] = -o.b2[m][i]; }
  s->dirty = true;
  return 0;
}

int mg_stencil_clone_impl(const mg_stencil* s, mg_stencil** out) {
  auto* a = new mg_stencil(*s);
  a->d_op = nullptr;
  a->dirty = true;
  *out = a;
  return 0;
}
