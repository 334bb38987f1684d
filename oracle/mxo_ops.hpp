// TEST INFRASTRUCTURE ONLY -- CPU oracle for maxwell_b200 (see mxo_geom.hpp header).
//
// CRS matrices with Epetra semantics, the Yee operator generators and the product / sum
// chains that define curlCurl, gradDiv, vecLapl, scaLapl and mRhs.
#pragma once
#include <algorithm>
#include <memory>
#include <type_traits>

#include "mxo_sim.hpp"

namespace mxo {

template <class S> inline S fromFactor(const cplx& f);
template <> inline double fromFactor<double>(const cplx& f) { return f.real(); }  // MxUtil.hpp:58-62
template <> inline cplx fromFactor<cplx>(const cplx& f) { return f; }

// Local-index CRS. Rows are stored with ascending local column index and duplicate
// insertions summed, which is what Epetra_CrsMatrix::FillComplete leaves behind
// (MxCrsMatrix.cpp:122-143,325-342). Explicit zeros are kept.
template <class S>
struct Csr {
  int64_t nrows = 0, ncols = 0;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
  std::vector<S> val;
  std::vector<int64_t> rowGid, colGid;  // the row / column maps
  int64_t nnz() const { return int64_t(col.size()); }
};

template <class S>
struct RowBuf {
  std::vector<std::pair<int32_t, S>> e;
  void add(int32_t c, S v) { e.emplace_back(c, v); }
  void flushInto(Csr<S>& m) {
    std::stable_sort(e.begin(), e.end(), [](auto& a, auto& b) { return a.first < b.first; });
    for (size_t i = 0; i < e.size();) {
      S s = e[i].second;
      size_t j = i + 1;
      while (j < e.size() && e[j].first == e[i].first) s += e[j++].second;
      m.col.push_back(e[i].first);
      m.val.push_back(s);
      i = j;
    }
    m.rowptr.push_back(int64_t(m.col.size()));
    e.clear();
  }
};

template <class S>
inline void startMatrix(Csr<S>& m, const Field& rowF, const Field& colF) {
  m.nrows = int64_t(rowF.gids.size());
  m.ncols = int64_t(colF.gids.size());
  m.rowGid = rowF.gids;
  m.colGid = colF.gids;
  m.rowptr.assign(1, 0);
}

inline int32_t colLid(const Field& f, int64_t gid) {
  const int32_t l = f.lid(gid);
  if (l < 0) throw std::runtime_error("mxo: operator references a column GID that is not in the domain map");
  return l;
}

inline void cellCompOf(const Field& f, int64_t gid, I3& cell, int& comp) {
  comp = int(gid % f.ncomp);                       // MxGridFieldIter.hpp:34,66-70
  cell = f.g->globalToCell(gid / f.ncomp);
}

// MxYeeDeyMittraCurlE.cpp:117-178
template <class S>
Csr<S> curlE(const Field& B, const Field& E) {
  Csr<S> m;
  startMatrix(m, B, E);
  const Grid& g = *B.g;
  const double id[3] = {1 / g.d[0], 1 / g.d[1], 1 / g.d[2]};
  RowBuf<S> row;
  for (int64_t gid : B.gids) {
    I3 cell; int c0;
    cellCompOf(B, gid, cell, c0);
    const int c1 = (c0 + 1) % 3, c2 = (c0 + 2) % 3;
    if (B.factor(c0, cell) != cplx(0.0)) {
      const double v[4] = {id[c1], -id[c1], -id[c2], id[c2]};
      int64_t cols[4]; cplx fac[4];
      cell[c1]++; cols[0] = E.gid(c2, cell); fac[0] = E.factor(c2, cell);
      cell[c1]--; cols[1] = E.gid(c2, cell); fac[1] = E.factor(c2, cell);
      cell[c2]++; cols[2] = E.gid(c1, cell); fac[2] = E.factor(c1, cell);
      cell[c2]--; cols[3] = E.gid(c1, cell); fac[3] = E.factor(c1, cell);
      for (int i = 0; i < 4; ++i)
        if (fac[i] != cplx(0.0)) row.add(colLid(E, cols[i]), fromFactor<S>(fac[i]) * S(v[i]));
    }
    row.flushInto(m);
  }
  return m;
}

// MxYeeDeyMittraCurlB.cpp:113-169
template <class S>
Csr<S> curlB(const Field& B, const Field& E) {
  Csr<S> m;
  startMatrix(m, E, B);
  const Grid& g = *B.g;
  const double id[3] = {1 / g.d[0], 1 / g.d[1], 1 / g.d[2]};
  RowBuf<S> row;
  for (int64_t gid : E.gids) {
    I3 cell; int c0;
    cellCompOf(E, gid, cell, c0);
    const int c1 = (c0 + 1) % 3, c2 = (c0 + 2) % 3;
    const double v[4] = {id[c1], -id[c1], -id[c2], id[c2]};
    int64_t cols[4]; cplx fac[4];
    cols[0] = B.gid(c2, cell); fac[0] = B.factor(c2, cell);
    cell[c1]--; cols[1] = B.gid(c2, cell); fac[1] = B.factor(c2, cell);
    cell[c1]++; cols[2] = B.gid(c1, cell); fac[2] = B.factor(c1, cell);
    cell[c2]--; cols[3] = B.gid(c1, cell); fac[3] = B.factor(c1, cell);
    for (int i = 0; i < 4; ++i)
      if (fac[i] != cplx(0.0)) row.add(colLid(B, cols[i]), fromFactor<S>(fac[i]) * S(v[i]));
    row.flushInto(m);
  }
  return m;
}

// MxYeeDeyMittraDivB.cpp:157-200
template <class S>
Csr<S> divB(const Field& B, const Field& Psi) {
  Csr<S> m;
  startMatrix(m, Psi, B);
  const Grid& g = *B.g;
  RowBuf<S> row;
  for (int64_t gid : Psi.gids) {
    I3 cell; int c0;
    cellCompOf(Psi, gid, cell, c0);
    int64_t cols[6]; cplx fac[6]; double v[6];
    for (int c = 0; c < 3; ++c) {
      v[2 * c] = 1.0 / g.d[c];
      v[2 * c + 1] = -1.0 / g.d[c];
      cell[c]++; cols[2 * c] = B.gid(c, cell); fac[2 * c] = B.factor(c, cell);
      cell[c]--; cols[2 * c + 1] = B.gid(c, cell); fac[2 * c + 1] = B.factor(c, cell);
    }
    for (int i = 0; i < 6; ++i)
      if (fac[i] != cplx(0.0)) row.add(colLid(B, cols[i]), fromFactor<S>(fac[i]) * S(v[i]));
    row.flushInto(m);
  }
  return m;
}

// MxYeeDeyMittraGradPsi.cpp:23-84
template <class S>
Csr<S> gradPsi(const Field& B, const Field& Psi) {
  Csr<S> m;
  startMatrix(m, B, Psi);
  const Grid& g = *B.g;
  RowBuf<S> row;
  for (int64_t gid : B.gids) {
    I3 cell; int c;
    cellCompOf(B, gid, cell, c);
    if (B.factor(c, cell) != cplx(0.0)) {
      const double inv = 1.0 / g.d[c];
      int64_t cols[2]; cplx fac[2];
      cols[0] = Psi.gid(0, cell); fac[0] = Psi.factor(0, cell);
      cell[c]--; cols[1] = Psi.gid(0, cell); fac[1] = Psi.factor(0, cell);
      const double v[2] = {inv, -inv};
      for (int i = 0; i < 2; ++i)
        if (fac[i] != cplx(0.0)) row.add(colLid(Psi, cols[i]), fromFactor<S>(fac[i]) * S(v[i]));
    }
    row.flushInto(m);
  }
  return m;
}

// MxYeeDeyMittraFracs.cpp:30-131: diagonal of PEC fractions (zeros stored explicitly).
template <class S>
Csr<S> fracs(const Field& F, bool inverse, double minFrac) {
  Csr<S> m;
  startMatrix(m, F, F);
  m.col.reserve(F.gids.size());
  m.val.reserve(F.gids.size());
  for (size_t i = 0; i < F.gids.size(); ++i) {
    I3 cell; int c;
    cellCompOf(F, F.gids[i], cell, c);
    double v = F.frac(c, cell, "pec");
    if (!inverse) {
      if (v != 0 && v < minFrac) v = minFrac;
    } else {
      if (v == 0) {}
      else if (v < minFrac) v = 1.0 / minFrac;
      else v = 1.0 / v;
    }
    m.col.push_back(int32_t(i));
    m.val.push_back(S(v));
    m.rowptr.push_back(int64_t(i + 1));
  }
  return m;
}

// MxCrsMatrix.cpp:31-45: scalar diagonal
template <class S>
Csr<S> diagonal(const Field& F, S d) {
  Csr<S> m;
  startMatrix(m, F, F);
  for (size_t i = 0; i < F.gids.size(); ++i) {
    m.col.push_back(int32_t(i));
    m.val.push_back(d);
    m.rowptr.push_back(int64_t(i + 1));
  }
  return m;
}

// C = A * B with the EpetraExt::MatrixMatrix::Multiply accumulation order
// (MxCrsMatrix.cpp:358-382): C(i,j) = sum over k in A's stored row order; structural
// zeros produced by cancellation are kept.
template <class S>
Csr<S> multiply(const Csr<S>& A, const Csr<S>& B) {
  if (A.ncols != B.nrows) throw std::runtime_error("mxo: multiply shape mismatch");
  Csr<S> C;
  C.nrows = A.nrows; C.ncols = B.ncols;
  C.rowGid = A.rowGid; C.colGid = B.colGid;
  C.rowptr.assign(A.nrows + 1, 0);
  std::vector<int32_t> cnt(A.nrows);
  auto rowProduct = [&](int64_t i, std::vector<std::pair<int32_t, S>>& acc) {
    acc.clear();
    for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) {
      const int32_t k = A.col[p];
      const S a = A.val[p];
      for (int64_t q = B.rowptr[k]; q < B.rowptr[k + 1]; ++q) {
        const int32_t j = B.col[q];
        const S prod = a * B.val[q];
        bool found = false;
        for (auto& t : acc)
          if (t.first == j) { t.second += prod; found = true; break; }
        if (!found) acc.emplace_back(j, prod);
      }
    }
    std::sort(acc.begin(), acc.end(), [](auto& x, auto& y) { return x.first < y.first; });
  };
#pragma omp parallel
  {
    std::vector<std::pair<int32_t, S>> acc;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < A.nrows; ++i) { rowProduct(i, acc); cnt[i] = int32_t(acc.size()); }
  }
  for (int64_t i = 0; i < A.nrows; ++i) C.rowptr[i + 1] = C.rowptr[i] + cnt[i];
  C.col.resize(C.rowptr[A.nrows]);
  C.val.resize(C.rowptr[A.nrows]);
#pragma omp parallel
  {
    std::vector<std::pair<int32_t, S>> acc;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < A.nrows; ++i) {
      rowProduct(i, acc);
      int64_t o = C.rowptr[i];
      for (auto& t : acc) { C.col[o] = t.first; C.val[o] = t.second; ++o; }
    }
  }
  return C;
}

// C = sa*A + sb*B on the union pattern (EpetraExt::MatrixMatrix::Add; MxCrsMatrix.cpp:401-430)
template <class S>
Csr<S> add(const Csr<S>& A, S sa, const Csr<S>& B, S sb) {
  if (A.nrows != B.nrows || A.ncols != B.ncols) throw std::runtime_error("mxo: add shape mismatch");
  Csr<S> C;
  C.nrows = A.nrows; C.ncols = A.ncols;
  C.rowGid = A.rowGid; C.colGid = A.colGid;
  C.rowptr.assign(1, 0);
  for (int64_t i = 0; i < A.nrows; ++i) {
    int64_t p = A.rowptr[i], pe = A.rowptr[i + 1], q = B.rowptr[i], qe = B.rowptr[i + 1];
    while (p < pe || q < qe) {
      if (q >= qe || (p < pe && A.col[p] < B.col[q])) { C.col.push_back(A.col[p]); C.val.push_back(sa * A.val[p]); ++p; }
      else if (p >= pe || B.col[q] < A.col[p]) { C.col.push_back(B.col[q]); C.val.push_back(sb * B.val[q]); ++q; }
      else { C.col.push_back(A.col[p]); C.val.push_back(sa * A.val[p] + sb * B.val[q]); ++p; ++q; }
    }
    C.rowptr.push_back(int64_t(C.col.size()));
  }
  return C;
}

// MxCrsMatrix.cpp:84-117: drop |v| <= 1e-12 (the reference tests the stored doubles; for
// complex storage that is each real/imaginary K-form entry separately -- an entry is kept
// if either part survives).
inline bool survivesPurge(double v) { return std::fabs(v) > 1.e-12; }
inline bool survivesPurge(const cplx& v) { return std::fabs(v.real()) > 1.e-12 || std::fabs(v.imag()) > 1.e-12; }
template <class S>
Csr<S> purgeZeros(const Csr<S>& A) {
  Csr<S> C;
  C.nrows = A.nrows; C.ncols = A.ncols;
  C.rowGid = A.rowGid; C.colGid = A.colGid;
  C.rowptr.assign(1, 0);
  for (int64_t i = 0; i < A.nrows; ++i) {
    for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p)
      if (survivesPurge(A.val[p])) { C.col.push_back(A.col[p]); C.val.push_back(A.val[p]); }
    C.rowptr.push_back(int64_t(C.col.size()));
  }
  return C;
}

template <class S>
void scaleInPlace(Csr<S>& A, S s) {
  for (auto& v : A.val) v *= s;
}

// Y = A X, column-major multivectors, each row summed in stored (ascending local column)
// order starting from zero, as Epetra_CrsMatrix::Apply does (MxCrsMatrix.cpp:347-353).
template <class S>
void spmm(const Csr<S>& A, const S* X, int64_t ldx, S* Y, int64_t ldy, int nvec) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < A.nrows; ++i) {
    for (int v = 0; v < nvec; ++v) {
      const S* x = X + v * ldx;
      S sum = S(0);
      for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) sum += A.val[p] * x[A.col[p]];
      Y[i + v * ldy] = sum;
    }
  }
}

// MxGridFieldInterpolator.cpp:28-65 (stencil), :68-122 (insertion, row-sum normalisation):
// (tri)linear interpolation of `from` at the component positions of `to`. Entries whose
// column is not in the from-map are dropped, then every row is divided by the sum of the
// magnitudes of what is left. Two latent bugs of the reference are NOT reproduced: the
// stencil weights use an unset point `p` (:39,60 -- we use the target position) and
// getRowSums never zeroes its accumulator (MxCrsMatrix.cpp:306-310); see DESIGN.md R12.
template <class S>
Csr<S> interpolator(const Field& from, const Field& to) {
  if (from.ncomp != to.ncomp) throw std::runtime_error("mxo: interpolator needs fields with equal component counts");
  Csr<S> m;
  startMatrix(m, to, from);
  const Grid& fg = *from.g;
  RowBuf<S> row;
  for (int64_t gid : to.gids) {
    I3 tcell; int comp;
    cellCompOf(to, gid, tcell, comp);
    const D3 point = to.g->nodeCoord(tcell) + to.xi[comp];
    I3 c0;
    for (int i = 0; i < 3; ++i)
      c0[i] = int(std::floor((point[i] - from.xi[comp][i] - fg.origin[i]) / fg.d[i]));
    double sum = 0.0;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 2; ++c) {
          const I3 cell{c0[0] + a, c0[1] + b, c0[2] + c};
          const D3 p0 = fg.nodeCoord(cell) + from.xi[comp];
          double w = 1.0;
          for (int i = 0; i < 3; ++i) w *= 1.0 - std::fabs(point[i] - p0[i]) / fg.d[i];
          // cells outside the guarded block cannot be interiorised meaningfully; skip them
          bool inRange = true;
          for (int i = 0; i < 3; ++i)
            if (cell[i] < -1 || cell[i] > fg.N[i] + 1) inRange = false;
          if (!inRange) continue;
          const int64_t cg = from.gid(comp, cell);
          const int32_t l = from.lid(cg);
          if (l < 0) continue;
          row.add(l, S(w));
          sum += std::fabs(w);
        }
    if (sum > 0)
      for (auto& e : row.e) e.second = e.second / S(sum);
    row.flushInto(m);
  }
  return m;
}

template <class S> inline S conjOf(const S& v);
template <> inline double conjOf<double>(const double& v) { return v; }
template <> inline cplx conjOf<cplx>(const cplx& v) { return std::conj(v); }

// (conjugate) transpose, rows sorted by column
template <class S>
Csr<S> transpose(const Csr<S>& A) {
  Csr<S> T;
  T.nrows = A.ncols; T.ncols = A.nrows;
  T.rowGid = A.colGid; T.colGid = A.rowGid;
  T.rowptr.assign(T.nrows + 1, 0);
  for (int64_t p = 0; p < A.nnz(); ++p) T.rowptr[A.col[p] + 1]++;
  for (int64_t i = 0; i < T.nrows; ++i) T.rowptr[i + 1] += T.rowptr[i];
  T.col.resize(A.nnz());
  T.val.resize(A.nnz());
  std::vector<int64_t> fill(T.rowptr.begin(), T.rowptr.end() - 1);
  for (int64_t i = 0; i < A.nrows; ++i)
    for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) {
      const int64_t o = fill[A.col[p]]++;
      T.col[o] = int32_t(i);
      T.val[o] = conjOf(A.val[p]);
    }
  return T;
}

// divide every row by the sum of the magnitudes of its entries (rows of zeros stay empty)
template <class S>
void normalizeRows(Csr<S>& A) {
  for (int64_t i = 0; i < A.nrows; ++i) {
    double sum = 0;
    for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) sum += std::abs(A.val[p]);
    if (sum > 0)
      for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) A.val[p] = A.val[p] / S(sum);
  }
}

// The reference's storage for complex scalars: real 2n x 2n "K form" with interleaved
// re/im rows and the 2x2 block [[a,-b],[b,a]] per entry (MxCrsMatrix.cpp:145-170,
// MxMap.cpp:90-108).
inline Csr<double> toKForm(const Csr<cplx>& A) {
  Csr<double> K;
  K.nrows = 2 * A.nrows; K.ncols = 2 * A.ncols;
  K.rowptr.assign(1, 0);
  for (auto g : A.rowGid) { K.rowGid.push_back(2 * g); K.rowGid.push_back(2 * g + 1); }
  for (auto g : A.colGid) { K.colGid.push_back(2 * g); K.colGid.push_back(2 * g + 1); }
  for (int64_t i = 0; i < A.nrows; ++i) {
    for (int part = 0; part < 2; ++part) {
      for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) {
        const double a = A.val[p].real(), b = A.val[p].imag();
        K.col.push_back(2 * A.col[p]);     K.val.push_back(part == 0 ? a : b);
        K.col.push_back(2 * A.col[p] + 1); K.val.push_back(part == 0 ? -b : a);
      }
      K.rowptr.push_back(int64_t(K.col.size()));
    }
  }
  return K;
}

// ---------------------------------------------------------------------------------------
// Simulation container (MxEMSim.cpp:54-225) and named operators (MxEMOps.cpp:39-168,
// MxMagWaveOp.cpp:137-245).
// ---------------------------------------------------------------------------------------
// MxDielectric.{hpp,cpp}: a shape with a (possibly anisotropic, possibly complex) permittivity tensor
struct Dielectric {
  std::string name;
  std::shared_ptr<Shape> shape;
  cplx eps[3][3];
  bool isDiag() const {  // MxDielectric.cpp:27-36 (dEps = 0)
    return !(std::abs(eps[0][1]) > 0 || std::abs(eps[0][2]) > 0 || std::abs(eps[1][2]) > 0);
  }
};

struct Sim {
  Grid grid;
  BCType lower[3] = {PERIODIC, PERIODIC, PERIODIC}, upper[3] = {PERIODIC, PERIODIC, PERIODIC};
  double phaseShifts[3] = {0, 0, 0};
  std::shared_ptr<Shape> pec;
  double dmFrac = 0.0;
  bool literalUpperPeriodicE = false;
  std::unique_ptr<Field> B, E, D, Psi;
  std::vector<Dielectric> diels;

  Sim(I3 n, D3 o, D3 l) : grid(n, o, l) {}

  void setup() {
    B.reset(new Field(&grid, FIELD_B));
    B->dmFrac = dmFrac;
    B->setBCs(lower, upper);
    E.reset(new Field(&grid, FIELD_E, B.get()));
    E->setBCs(lower, upper);
    Psi.reset(new Field(&grid, FIELD_PSI, B.get()));
    Psi->setBCs(lower, upper);
    std::vector<Field*> fields = {B.get(), E.get()};
    if (hasDielectric()) {                      // MxEMSim.cpp:90-95: the D field exists only with dielectrics
      D.reset(new Field(&grid, FIELD_D));
      D->setBCs(lower, upper);
      fields.push_back(D.get());
    }
    fields.push_back(Psi.get());
    for (Field* f : fields) {
      f->setPhaseShifts(phaseShifts);
      f->literalUpperPeriodicE = literalUpperPeriodicE;
    }
    if (pec) {
      for (Field* f : fields) f->addShapeRep(*pec, "pec", true);  // B first, then E, [D], psi
    }
    // MxEMSim.cpp:134-148: dielectric fractions on every field except B (and H); no region
    for (Field* f : fields)
      if (f->kind != FIELD_B)
        for (const Dielectric& d : diels) f->addShapeRep(*d.shape, d.name, false);
    if (!pec)
      for (Field* f : fields)
        if (f->kind != FIELD_D) f->setMap();
    if (D) D->shareMap(*E);                      // MxEMSim.cpp:186-190
  }
  bool hasDielectric() const { return !diels.empty(); }
  bool hasPEC() const { return bool(pec); }
  bool isComplex() const { return phaseShifts[0] != 0 || phaseShifts[1] != 0 || phaseShifts[2] != 0; }
};

// ---- MxYeeFitInvEps (second-order anisotropic inverse permittivity) ---------------------------
struct M3 {
  cplx a[3][3];
};
inline M3 m3Zero() { M3 m; for (auto& r : m.a) for (auto& v : r) v = 0.0; return m; }
inline M3 m3Eye() { M3 m = m3Zero(); for (int i = 0; i < 3; ++i) m.a[i][i] = 1.0; return m; }
inline M3 m3Mul(const M3& x, const M3& y) {
  M3 r = m3Zero();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      cplx s = 0.0;
      for (int k = 0; k < 3; ++k) s += x.a[i][k] * y.a[k][j];
      r.a[i][j] = s;
    }
  return r;
}
inline M3 m3Inv(const M3& m) {
  const cplx (*a)[3] = m.a;
  const cplx det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                   a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
  M3 r;
  r.a[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) / det;
  r.a[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) / det;
  r.a[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) / det;
  r.a[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) / det;
  r.a[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) / det;
  r.a[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) / det;
  r.a[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) / det;
  r.a[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) / det;
  r.a[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) / det;
  return r;
}

struct EpsStencil {
  int comps[9];
  I3 cells[9];
};
// MxYeeFitInvEps.cpp:33-72: E_c0(cell) couples to D_c0(cell), four D_c1 and four D_c2 neighbours
inline EpsStencil epsStencil(int c0, const I3& cell) {
  const int c1 = (c0 + 1) % 3, c2 = (c1 + 1) % 3;
  EpsStencil st;
  st.comps[0] = c0;
  for (int i = 1; i < 5; ++i) st.comps[i] = c1;
  for (int i = 5; i < 9; ++i) st.comps[i] = c2;
  for (int i = 0; i < 9; ++i) st.cells[i] = cell;
  st.cells[1][c1]--;
  st.cells[3][c0]++; st.cells[3][c1]--;
  st.cells[4][c0]++;
  st.cells[5][c2]--;
  st.cells[7][c0]++; st.cells[7][c2]--;
  st.cells[8][c0]++;
  return st;
}

// gamma / pi averaging of one (E_c0, D_c1, D_c2) triplet (MxYeeFitInvEps.cpp:327-416)
inline M3 tupleUpdate(const Sim& s, const int comps[3], const I3 cells[3], const D3& n) {
  const Field &E = *s.E, &D = *s.D;
  const cplx nc[3] = {n[0], n[1], n[2]};
  M3 aveGamma = m3Zero(), avePi = m3Zero();
  double lsum[3] = {0, 0, 0}, asum[3] = {0, 0, 0};
  auto accumulate = [&](const cplx eps[3][3], const double lfr[3], const double afr[3]) {
    M3 e, nn, eyeMinusEps;
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) {
        e.a[j][k] = eps[j][k];
        nn.a[j][k] = nc[j] * nc[k];
        eyeMinusEps.a[j][k] = (j == k ? cplx(1.0) : cplx(0.0)) - eps[j][k];
      }
    cplx nEn = 0.0;
    for (int j = 0; j < 3; ++j) {
      cplx t = 0.0;
      for (int k = 0; k < 3; ++k) t += e.a[j][k] * nc[k];
      nEn += nc[j] * t;
    }
    M3 gamma = m3Mul(nn, eyeMinusEps);
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) gamma.a[j][k] = (j == k ? cplx(1.0) : cplx(0.0)) + gamma.a[j][k] / nEn;
    const M3 pi = m3Mul(e, gamma);
    for (int j = 0; j < 3; ++j)      // cartProjs[j] * frac[j] selects row j
      for (int k = 0; k < 3; ++k) {
        aveGamma.a[j][k] += lfr[j] * gamma.a[j][k];
        avePi.a[j][k] += afr[j] * pi.a[j][k];
      }
  };
  for (const Dielectric& d : s.diels) {
    double lfr[3], afr[3];
    for (int j = 0; j < 3; ++j) {
      const int c = comps[j];
      lfr[c] = E.frac(c, cells[j], d.name);
      afr[c] = D.frac(c, cells[j], d.name);
      lsum[c] += lfr[c];
      asum[c] += afr[c];
    }
    accumulate(d.eps, lfr, afr);
  }
  cplx bg[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};    // background dielectric: vacuum
  double lfr[3], afr[3];
  for (int j = 0; j < 3; ++j) { lfr[j] = 1.0 - lsum[j]; afr[j] = 1.0 - asum[j]; }
  accumulate(bg, lfr, afr);
  return m3Mul(aveGamma, m3Inv(avePi));
}

// MxYeeFitInvEps.cpp:420-596 (no PML). Entries whose D column is not in the map are dropped (the
// reference would hand Epetra a column outside the domain map); explicit zeros in the map are kept.
template <class S>
Csr<S> invEps(const Sim& s) {
  const Field &E = *s.E, &D = *s.D;
  Csr<S> m;
  startMatrix(m, E, D);
  RowBuf<S> row;
  for (size_t r = 0; r < E.gids.size(); ++r) {
    I3 cell; int c0;
    cellCompOf(E, E.gids[r], cell, c0);
    bool inDiel = false, epsIsDiag = false;
    const Dielectric* diel = nullptr;
    for (const Dielectric& d : s.diels) {
      const double l0 = E.frac(c0, cell, d.name), a0 = D.frac(c0, cell, d.name);
      if (l0 == 1 && a0 == 1) { inDiel = true; epsIsDiag = d.isDiag(); diel = &d; break; }
      else if (l0 == 0 && a0 == 0) continue;
      else { inDiel = true; diel = &d; break; }
    }
    if (!inDiel) epsIsDiag = true;   // background (vacuum) is diagonal
    if (epsIsDiag) {
      const cplx e00 = inDiel ? diel->eps[c0][c0] : cplx(1.0);
      const cplx v = D.factor(c0, cell) / e00;
      row.add(int32_t(r), fromFactor<S>(v));
    } else {
      const EpsStencil st = epsStencil(c0, cell);
      // interface normal (MxYeeFitInvEps.cpp:271-324)
      D3 nsum{0, 0, 0};
      std::vector<D3> norms;
      for (const Dielectric& d : s.diels) {
        bool cut = false;
        for (int j = 0; j < 9; ++j) {
          const double lf = E.frac(st.comps[j], st.cells[j], d.name), af = D.frac(st.comps[j], st.cells[j], d.name);
          if ((lf != 0 && lf != 1) || (af != 0 && af != 1)) cut = true;
        }
        if (cut) {
          const D3 g = d.shape->grad(E.g->nodeCoord(st.cells[0]) + E.xi[st.comps[0]]);
          norms.push_back(g / norm(g));
        }
      }
      D3 n{1, 0, 0};
      if (!norms.empty()) {
        for (size_t i = 0; i < norms.size(); ++i) {
          const double sg = (i > 0 && dot(norms[0], norms[i]) < 0) ? -1.0 : 1.0;
          nsum = nsum + sg * norms[i];
        }
        n = nsum / norm(nsum);
      }
      // triplets (MxYeeFitInvEps.cpp:126-171)
      static const int T[8][3] = {{0, 1, 5}, {0, 1, 6}, {0, 2, 5}, {0, 2, 6}, {0, 3, 7}, {0, 3, 8}, {0, 4, 7}, {0, 4, 8}};
      std::vector<int> used;
      for (int t = 0; t < 8; ++t) {
        bool use = true;
        if (s.hasPEC())
          for (int j = 0; j < 3; ++j)
            if (E.frac(st.comps[T[t][j]], st.cells[T[t][j]], "pec") == 0) use = false;
        if (use) used.push_back(t);
      }
      cplx vals[9];
      for (auto& v : vals) v = 0.0;
      for (int t : used) {
        int comps[3]; I3 cells[3];
        for (int j = 0; j < 3; ++j) { comps[j] = st.comps[T[t][j]]; cells[j] = st.cells[T[t][j]]; }
        const M3 ie = tupleUpdate(s, comps, cells, n);
        for (int j = 0; j < 3; ++j) vals[T[t][j]] += ie.a[c0][comps[j]] / double(used.size());
      }
      for (int i = 0; i < 9; ++i) {
        const cplx v = vals[i] * D.factor(st.comps[i], st.cells[i]);
        const int32_t l = D.lid(D.gid(st.comps[i], st.cells[i]));
        if (l >= 0) row.add(l, fromFactor<S>(v));
      }
    }
    row.flushInto(m);
  }
  return m;
}

// MxYeeFitInvEps.cpp:650-725: cell-averaged scalar 1/eps on the psi field (3 / trace(eps))
template <class S>
Csr<S> invEpsVolAve(const Sim& s) {
  const Field& Psi = *s.Psi;
  Csr<S> m;
  startMatrix(m, Psi, Psi);
  for (size_t i = 0; i < Psi.gids.size(); ++i) {
    I3 cell; int c;
    cellCompOf(Psi, Psi.gids[i], cell, c);
    double sum = 0;
    cplx ave = 0.0;
    for (const Dielectric& d : s.diels) {
      const double f = Psi.frac(c, cell, d.name);
      sum += f;
      ave += f * (3.0 / (d.eps[0][0] + d.eps[1][1] + d.eps[2][2]));
    }
    ave += (1.0 - sum) * cplx(1.0);
    m.col.push_back(int32_t(i));
    m.val.push_back(fromFactor<S>(ave));
    m.rowptr.push_back(int64_t(i + 1));
  }
  return m;
}

template <class S>
Csr<S> buildOp(const Sim& s, const std::string& name) {
  const Field &B = *s.B, &E = *s.E, &Psi = *s.Psi;
  if (name == "curlE") return curlE<S>(B, E);
  if (name == "curlB") return curlB<S>(B, E);
  if (name == "divB") return divB<S>(B, Psi);
  if (name == "gradPsi") return gradPsi<S>(B, Psi);
  if (name == "dmA") return fracs<S>(B, false, 0.e-12);        // MxEMOps.cpp:56-58
  if (name == "dmL") return fracs<S>(E, false, 0.e-6);         // MxEMOps.cpp:61-63
  if (name == "dmVInv") return fracs<S>(Psi, true, 0.e-6);     // MxEMOps.cpp:125-127
  if (name == "invEps") return invEps<S>(s);                    // MxEMOps.cpp:72-81
  if (name == "invEpsVolAve") return invEpsVolAve<S>(s);
  if (name == "curlCurl") {                                     // MxMagWaveOp.cpp:144-153
    Csr<S> m = curlB<S>(B, E);
    if (s.hasDielectric()) m = multiply(invEps<S>(s), m);
    if (s.hasPEC()) m = multiply(fracs<S>(E, false, 0.e-6), m);
    return multiply(curlE<S>(B, E), m);
  }
  if (name == "gradDiv") {                                      // MxMagWaveOp.cpp:156-179
    Csr<S> m = s.hasPEC() ? multiply(divB<S>(B, Psi), fracs<S>(B, false, 0.e-12)) : divB<S>(B, Psi);
    if (s.hasPEC()) m = multiply(fracs<S>(Psi, true, 0.e-6), m);
    if (s.hasDielectric()) m = multiply(invEpsVolAve<S>(s), m);
    m = multiply(gradPsi<S>(B, Psi), m);
    if (s.hasPEC()) m = multiply(fracs<S>(B, false, 0.e-12), m);
    return m;
  }
  if (name == "vecLapl") {                                      // MxMagWaveOp.cpp:183-205
    return purgeZeros(add(buildOp<S>(s, "curlCurl"), S(1.0), buildOp<S>(s, "gradDiv"), S(-1.0)));
  }
  if (name == "scaLapl") {                                      // MxMagWaveOp.cpp:208-223
    Csr<S> m = gradPsi<S>(B, Psi);
    if (s.hasPEC()) m = multiply(fracs<S>(B, false, 0.e-12), m);
    m = multiply(divB<S>(B, Psi), m);
    scaleInPlace(m, S(-1.0));
    return m;
  }
  if (name == "mRhs") {                                         // MxMagWaveOp.cpp:227-241
    return s.hasPEC() ? fracs<S>(B, false, 0.e-12) : diagonal<S>(B, S(1.0));
  }
  throw std::runtime_error("mxo: unknown operator '" + name + "'");
}

}  // namespace mxo
