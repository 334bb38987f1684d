// Second-order anisotropic inverse permittivity of the Yee grid (MxYeeFitInvEps.cpp:33-72 stencil, :126-171 triplets,
// :271-324 interface normal, :327-416 gamma / pi averaging, :420-596 matrix rows, :650-725 cell-averaged scalar 1/eps)
// as host/device row functions, like the generators of mxg_yee.h. Complex arithmetic follows std::complex as GCC
// evaluates it -- products by the plain four-multiplication formula, quotients by libgcc's __divdc3 (Smith's method with
// the subnormal-ratio alternative) -- because the host path the results are compared with computes that way.
#pragma once
#include "mxg_shape.h"
#include "mxg_yee.h"

namespace mxy {

MXY_HD Cx subS(Cx a, Cx b) { return {a.re - b.re, a.im - b.im}; }
MXY_HD Cx scaleS(double s, Cx a) { return {a.re * s, a.im * s}; }      // T * complex<T>: both parts times the scalar
MXY_HD Cx overReal(Cx a, double s) { return {a.re / s, a.im / s}; }     // complex<T> / T
// libgcc2.c __divdc3 for finite operands: (a + ib) / (c + id)
MXY_HD Cx divS(Cx num, Cx den) {
  double a = num.re, b = num.im, c = den.re, d = den.im;
  const double RBIG = 1.7976931348623157e308 / 2, RMIN = 2.2250738585072014e-308, RMIN2 = 2.220446049250313e-16,
               RMINSCAL = 4503599627370496.0, RMAX2 = RBIG * RMIN2;
  double x, y;
  if (absD(c) < absD(d)) {
    if (absD(d) >= RBIG) { a = a / 2; b = b / 2; c = c / 2; d = d / 2; }
    if (absD(d) < RMIN2) { a = a * RMINSCAL; b = b * RMINSCAL; c = c * RMINSCAL; d = d * RMINSCAL; }
    else if (((absD(a) < RMIN) && (absD(b) < RMAX2) && (absD(d) < RMAX2)) || ((absD(b) < RMIN) && (absD(a) < RMAX2) && (absD(d) < RMAX2))) {
      a = a * RMINSCAL; b = b * RMINSCAL; c = c * RMINSCAL; d = d * RMINSCAL;
    }
    const double ratio = c / d;
    const double denom = (c * ratio) + d;
    if (absD(ratio) > RMIN) {
      x = ((a * ratio) + b) / denom;
      y = ((b * ratio) - a) / denom;
    } else {
      x = ((c * (a / d)) + b) / denom;
      y = ((c * (b / d)) - a) / denom;
    }
  } else {
    if (absD(c) >= RBIG) { a = a / 2; b = b / 2; c = c / 2; d = d / 2; }
    if (absD(c) < RMIN2) { a = a * RMINSCAL; b = b * RMINSCAL; c = c * RMINSCAL; d = d * RMINSCAL; }
    else if (((absD(a) < RMIN) && (absD(b) < RMAX2) && (absD(c) < RMAX2)) || ((absD(b) < RMIN) && (absD(a) < RMAX2) && (absD(c) < RMAX2))) {
      a = a * RMINSCAL; b = b * RMINSCAL; c = c * RMINSCAL; d = d * RMINSCAL;
    }
    const double ratio = d / c;
    const double denom = (d * ratio) + c;
    if (absD(ratio) > RMIN) {
      x = ((b * ratio) + a) / denom;
      y = (b - (a * ratio)) / denom;
    } else {
      x = ((d * (b / c)) + a) / denom;
      y = (b - (d * (a / c))) / denom;
    }
  }
  return {x, y};
}

struct M3 {
  Cx a[3][3];
};
MXY_HD M3 m3Zero() {
  M3 m;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) m.a[i][j] = {0.0, 0.0};
  return m;
}
MXY_HD M3 m3Mul(const M3& x, const M3& y) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Cx s = {0.0, 0.0};
      for (int k = 0; k < 3; ++k) s = addS(s, mulS(x.a[i][k], y.a[k][j]));
      r.a[i][j] = s;
    }
  return r;
}
MXY_HD M3 m3Inv(const M3& m) {
  const Cx(*a)[3] = m.a;
  const Cx det = addS(subS(mulS(a[0][0], subS(mulS(a[1][1], a[2][2]), mulS(a[1][2], a[2][1]))),
                           mulS(a[0][1], subS(mulS(a[1][0], a[2][2]), mulS(a[1][2], a[2][0])))),
                      mulS(a[0][2], subS(mulS(a[1][0], a[2][1]), mulS(a[1][1], a[2][0]))));
  M3 r;
  r.a[0][0] = divS(subS(mulS(a[1][1], a[2][2]), mulS(a[1][2], a[2][1])), det);
  r.a[0][1] = divS(subS(mulS(a[0][2], a[2][1]), mulS(a[0][1], a[2][2])), det);
  r.a[0][2] = divS(subS(mulS(a[0][1], a[1][2]), mulS(a[0][2], a[1][1])), det);
  r.a[1][0] = divS(subS(mulS(a[1][2], a[2][0]), mulS(a[1][0], a[2][2])), det);
  r.a[1][1] = divS(subS(mulS(a[0][0], a[2][2]), mulS(a[0][2], a[2][0])), det);
  r.a[1][2] = divS(subS(mulS(a[0][2], a[1][0]), mulS(a[0][0], a[1][2])), det);
  r.a[2][0] = divS(subS(mulS(a[1][0], a[2][1]), mulS(a[1][1], a[2][0])), det);
  r.a[2][1] = divS(subS(mulS(a[0][1], a[2][0]), mulS(a[0][0], a[2][1])), det);
  r.a[2][2] = divS(subS(mulS(a[0][0], a[1][1]), mulS(a[0][1], a[1][0])), det);
  return r;
}

MXY_HD double repFrac(const Sim& s, const double* rep, int ncomp, int comp, const int cell[3]) {
  if (!rep) return 1.0;
  const int64_t fi = fullIndex(s.g, cell);
  if (fi < 0) { if (s.err) *s.err = 1; return 1.0; }
  return rep[comp + ncomp * fi];
}

// MxYeeFitInvEps.cpp:33-72: E_c0(cell) couples to D_c0(cell), four D_c1 and four D_c2 neighbours
struct EpsStencil {
  int comps[9];
  int cells[9][3];
};
MXY_HD void epsStencil(int c0, const int cell[3], EpsStencil& st) {
  const int c1 = (c0 + 1) % 3, c2 = (c1 + 1) % 3;
  st.comps[0] = c0;
  for (int i = 1; i < 5; ++i) st.comps[i] = c1;
  for (int i = 5; i < 9; ++i) st.comps[i] = c2;
  for (int i = 0; i < 9; ++i)
    for (int k = 0; k < 3; ++k) st.cells[i][k] = cell[k];
  st.cells[1][c1]--;
  st.cells[3][c0]++; st.cells[3][c1]--;
  st.cells[4][c0]++;
  st.cells[5][c2]--;
  st.cells[7][c0]++; st.cells[7][c2]--;
  st.cells[8][c0]++;
}

MXY_HD void epsAccumulate(const Cx eps[3][3], const Cx nc[3], const double lfr[3], const double afr[3], M3& aveGamma, M3& avePi) {
  M3 e, nn, eyeMinusEps;
  for (int j = 0; j < 3; ++j)
    for (int k = 0; k < 3; ++k) {
      e.a[j][k] = eps[j][k];
      nn.a[j][k] = mulS(nc[j], nc[k]);
      eyeMinusEps.a[j][k] = subS(Cx{j == k ? 1.0 : 0.0, 0.0}, eps[j][k]);
    }
  Cx nEn = {0.0, 0.0};
  for (int j = 0; j < 3; ++j) {
    Cx t = {0.0, 0.0};
    for (int k = 0; k < 3; ++k) t = addS(t, mulS(e.a[j][k], nc[k]));
    nEn = addS(nEn, mulS(nc[j], t));
  }
  M3 gamma = m3Mul(nn, eyeMinusEps);
  for (int j = 0; j < 3; ++j)
    for (int k = 0; k < 3; ++k) gamma.a[j][k] = addS(Cx{j == k ? 1.0 : 0.0, 0.0}, divS(gamma.a[j][k], nEn));
  const M3 pi = m3Mul(e, gamma);
  for (int j = 0; j < 3; ++j)      // the Cartesian projector times the fraction selects row j
    for (int k = 0; k < 3; ++k) {
      aveGamma.a[j][k] = addS(aveGamma.a[j][k], scaleS(lfr[j], gamma.a[j][k]));
      avePi.a[j][k] = addS(avePi.a[j][k], scaleS(afr[j], pi.a[j][k]));
    }
}

// gamma / pi averaging of one (E_c0, D_c1, D_c2) triplet (MxYeeFitInvEps.cpp:327-416)
MXY_HD M3 epsTupleUpdate(const Sim& s, const int comps[3], const int cells[3][3], const double n[3]) {
  const Cx nc[3] = {{n[0], 0.0}, {n[1], 0.0}, {n[2], 0.0}};
  M3 aveGamma = m3Zero(), avePi = m3Zero();
  double lsum[3] = {0, 0, 0}, asum[3] = {0, 0, 0};
  for (int d = 0; d < s.numDiel; ++d) {
    const DielectricRep& D = s.diel[d];
    double lfr[3], afr[3];
    for (int j = 0; j < 3; ++j) {
      const int c = comps[j];
      lfr[c] = repFrac(s, D.fracE, 3, c, cells[j]);
      afr[c] = repFrac(s, D.fracD, 3, c, cells[j]);
      lsum[c] += lfr[c];
      asum[c] += afr[c];
    }
    Cx eps[3][3];
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) eps[j][k] = {D.epsRe[3 * j + k], D.epsIm[3 * j + k]};
    epsAccumulate(eps, nc, lfr, afr, aveGamma, avePi);
  }
  Cx bg[3][3];
  for (int j = 0; j < 3; ++j)
    for (int k = 0; k < 3; ++k) bg[j][k] = {j == k ? 1.0 : 0.0, 0.0};     // background: vacuum
  double lfr[3], afr[3];
  for (int j = 0; j < 3; ++j) { lfr[j] = 1.0 - lsum[j]; afr[j] = 1.0 - asum[j]; }
  epsAccumulate(bg, nc, lfr, afr, aveGamma, avePi);
  return m3Mul(aveGamma, m3Inv(avePi));
}

// MxYeeFitInvEps.cpp:420-596 (no PML): row of E = invEps D. Entries whose D column is not in the map are dropped,
// explicit zeros inside the map are kept.
template <class S>
struct InvEpsRow {
  static constexpr int kMax = 9;
  const Sim* sim;
  int hasPEC;
  MXY_HD int operator()(int64_t row, int32_t* cols, S* vals) const {
    const Sim& s = *sim;
    const Field &E = s.f[FIELD_E], &Df = s.f[FIELD_D];
    int cell[3], c0;
    cellCompOf(s, E, E.gids[row], cell, c0);
    bool inDiel = false, epsIsDiag = false;
    int diel = -1;
    for (int d = 0; d < s.numDiel; ++d) {
      const double l0 = repFrac(s, s.diel[d].fracE, 3, c0, cell), a0 = repFrac(s, s.diel[d].fracD, 3, c0, cell);
      if (l0 == 1 && a0 == 1) { inDiel = true; epsIsDiag = s.diel[d].isDiag != 0; diel = d; break; }
      else if (l0 == 0 && a0 == 0) continue;
      else { inDiel = true; diel = d; break; }
    }
    if (!inDiel) epsIsDiag = true;      // the background (vacuum) is diagonal
    int n = 0;
    if (epsIsDiag) {
      const Cx e00 = inDiel ? Cx{s.diel[diel].epsRe[4 * c0], s.diel[diel].epsIm[4 * c0]} : Cx{1.0, 0.0};
      double fr, fi;
      factorOf(s, FIELD_D, c0, cell, fr, fi);
      const Cx v = divS(Cx{fr, fi}, e00);
      cols[0] = int32_t(row);
      vals[0] = fromParts<S>(v.re, v.im);
      return 1;
    }
    EpsStencil st;
    epsStencil(c0, cell, st);
    // interface normal (MxYeeFitInvEps.cpp:271-324)
    double nsum[3] = {0, 0, 0}, first[3] = {0, 0, 0};
    int numNorms = 0;
    for (int d = 0; d < s.numDiel; ++d) {
      bool cut = false;
      for (int j = 0; j < 9; ++j) {
        const double lf = repFrac(s, s.diel[d].fracE, 3, st.comps[j], st.cells[j]), af = repFrac(s, s.diel[d].fracD, 3, st.comps[j], st.cells[j]);
        if ((lf != 0 && lf != 1) || (af != 0 && af != 1)) cut = true;
      }
      if (!cut) continue;
      mxa::V3 p;
      for (int k = 0; k < 3; ++k) p.v[k] = (s.g.origin[k] + double(st.cells[0][k]) * s.g.d[k]) + E.xi[st.comps[0]][k];
      const mxa::V3 g = mxa::shapeGrad(static_cast<const mxa::ShapeNode*>(s.diel[d].shape), p);
      const mxa::V3 nrm = mxa::vDiv(g, mxa::vNorm(g));
      if (numNorms == 0) for (int k = 0; k < 3; ++k) first[k] = nrm.v[k];
      const double dt = first[0] * nrm.v[0] + first[1] * nrm.v[1] + first[2] * nrm.v[2];
      const double sg = (numNorms > 0 && dt < 0) ? -1.0 : 1.0;
      for (int k = 0; k < 3; ++k) nsum[k] = nsum[k] + sg * nrm.v[k];
      ++numNorms;
    }
    double nvec[3] = {1, 0, 0};
    if (numNorms > 0) {
      const double len = ::sqrt(nsum[0] * nsum[0] + nsum[1] * nsum[1] + nsum[2] * nsum[2]);
      for (int k = 0; k < 3; ++k) nvec[k] = nsum[k] / len;
    }
    // triplets (MxYeeFitInvEps.cpp:126-171)
    const int T[8][3] = {{0, 1, 5}, {0, 1, 6}, {0, 2, 5}, {0, 2, 6}, {0, 3, 7}, {0, 3, 8}, {0, 4, 7}, {0, 4, 8}};
    bool use[8];
    int numUsed = 0;
    for (int t = 0; t < 8; ++t) {
      use[t] = true;
      if (hasPEC)
        for (int j = 0; j < 3; ++j)
          if (repFrac(s, E.region, 3, st.comps[T[t][j]], st.cells[T[t][j]]) == 0) use[t] = false;
      if (use[t]) ++numUsed;
    }
    Cx acc[9];
    for (int i = 0; i < 9; ++i) acc[i] = {0.0, 0.0};
    for (int t = 0; t < 8; ++t) {
      if (!use[t]) continue;
      int comps[3], cells[3][3];
      for (int j = 0; j < 3; ++j) {
        comps[j] = st.comps[T[t][j]];
        for (int k = 0; k < 3; ++k) cells[j][k] = st.cells[T[t][j]][k];
      }
      const M3 ie = epsTupleUpdate(s, comps, cells, nvec);
      for (int j = 0; j < 3; ++j) acc[T[t][j]] = addS(acc[T[t][j]], overReal(ie.a[c0][comps[j]], double(numUsed)));
    }
    for (int i = 0; i < 9; ++i) {
      double fr, fi;
      factorOf(s, FIELD_D, st.comps[i], st.cells[i], fr, fi);
      const Cx v = mulS(acc[i], Cx{fr, fi});
      const int32_t l = Df.lidOf[gidOf(s.g, Df, st.comps[i], st.cells[i])];
      if (l >= 0) { cols[n] = l; vals[n] = fromParts<S>(v.re, v.im); ++n; }
    }
    return flushRow<S>(n, cols, vals);
  }
};

// MxYeeFitInvEps.cpp:650-725: cell-averaged scalar 1 / eps on the psi field (3 / trace(eps))
template <class S>
struct InvEpsVolAveRow {
  static constexpr int kMax = 1;
  const Sim* sim;
  MXY_HD int operator()(int64_t row, int32_t* cols, S* vals) const {
    const Sim& s = *sim;
    const Field& P = s.f[FIELD_PSI];
    int cell[3], c;
    cellCompOf(s, P, P.gids[row], cell, c);
    double sum = 0;
    Cx ave = {0.0, 0.0};
    for (int d = 0; d < s.numDiel; ++d) {
      const DielectricRep& D = s.diel[d];
      const double f = repFrac(s, D.fracPsi, 1, c, cell);
      sum += f;
      const Cx tr = addS(addS(Cx{D.epsRe[0], D.epsIm[0]}, Cx{D.epsRe[4], D.epsIm[4]}), Cx{D.epsRe[8], D.epsIm[8]});
      ave = addS(ave, scaleS(f, divS(Cx{3.0, 0.0}, tr)));
    }
    ave = addS(ave, scaleS(1.0 - sum, Cx{1.0, 0.0}));
    cols[0] = int32_t(row);
    vals[0] = fromParts<S>(ave.re, ave.im);
    return 1;
  }
};

}  // namespace mxy
