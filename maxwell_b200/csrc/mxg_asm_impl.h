// Operator assembly on an executor: the simulation set-up of MxEMSim (fields, boundary conditions, DOF
// maps), the named operators of MxEMOps / MxMagWaveOp::initMatrices and the CRS algebra they are chained with. The
// executor X supplies memory, a parallel for-each over rows, a prefix scan and a maximum; mxg_asm.cu instantiates this
// with CUDA kernels (the product), tests/cpp/asm_replay.cpp with plain loops (CPU replay of the same row functions).
//
// Reference: MxEMSim.cpp:54-199 (set-up order), MxYeeElecFieldBase.cpp:89-138 / MxYeeMagFieldBase.cpp:119-168 /
// MxYeePsiField.cpp:74-97 (wall translation), MxGridField.cpp:34-38 (phase factors), MxEMOps.cpp:39-168 and
// MxMagWaveOp.cpp:137-245 (operator chains), MxCrsMatrix.cpp:358-430 (multiply / add).
#pragma once
#include <cmath>
#include <complex>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "mxg_yee.h"
#include "mxg_shape.h"
#include "mxg_eps.h"

namespace mxa {

using mxy::Cx;

template <class S>
struct Csr {
  int64_t nrows = 0, ncols = 0, nnz = 0;
  int64_t* rowptr = nullptr;
  int32_t* col = nullptr;
  S* val = nullptr;
  int rowField = -1, colField = -1;   // which field maps the rows / columns live on
  mxy::CsrView<S> view() const { return {nrows, ncols, rowptr, col, val}; }
};

template <class S> inline S scalarFrom(double re, double im);
template <> inline double scalarFrom<double>(double re, double) { return re; }
template <> inline Cx scalarFrom<Cx>(double re, double im) { return {re, im}; }

template <class S>
struct ScaleVals {
  S* val;
  S s;
  MXY_HD void operator()(int64_t i) const { val[i] = mxy::mulS(val[i], s); }
};
struct RowLen2 {      // |row i of A| + |row i of B|
  const int64_t *a, *b;
  MXY_HD int operator()(int64_t i) const { return int((a[i + 1] - a[i]) + (b ? b[i + 1] - b[i] : 0)); }
};
template <class Fn>
struct StoreInt {
  Fn fn;
  int32_t* out;
  MXY_HD void operator()(int64_t i) const { out[i] = fn(i); }
};

// executor memory released at scope exit (temporaries of the passes below; an exception must not leak them)
template <class X, class T>
struct Scoped {
  X& x;
  T* p;
  Scoped(X& exec, int64_t n) : x(exec), p(exec.template alloc<T>(n)) {}
  ~Scoped() { x.free(p); }
  Scoped(const Scoped&) = delete;
  Scoped& operator=(const Scoped&) = delete;
  T* release() { T* q = p; p = nullptr; return q; }
};

template <class X>
class Assembler {
 public:
  explicit Assembler(X& x) : x_(x) { std::memset(&h_, 0, sizeof(h_)); }
  ~Assembler() {
    for (int k = 0; k < mxy::NUM_FIELDS; ++k) {
      x_.free(const_cast<double*>(h_.f[k].region));
      x_.free(const_cast<int32_t*>(h_.f[k].lidOf));
      x_.free(const_cast<int64_t*>(h_.f[k].gids));
    }
    x_.free(const_cast<double*>(h_.f[mxy::FIELD_D].region));      // D borrows E's map
    for (int d = 0; d < h_.numDiel; ++d) {
      x_.free(const_cast<double*>(h_.diel[d].fracE));
      x_.free(const_cast<double*>(h_.diel[d].fracD));
      x_.free(const_cast<double*>(h_.diel[d].fracPsi));
      x_.free(const_cast<void*>(h_.diel[d].shape));
    }
    x_.free(h_.err);
    x_.free(dSim_);
  }
  Assembler(const Assembler&) = delete;
  Assembler& operator=(const Assembler&) = delete;

  X& exec() { return x_; }
  const mxy::Sim& sim() const { return h_; }
  bool hasPEC() const { return hasPEC_; }
  bool isSetUp() const { return setUp_; }

  // MxEMSim.cpp:54-120: grid, the three Yee fields, wall translation, Bloch phases
  void init(const int n[3], const double origin[3], const double size[3], const int lower[3], const int upper[3],
            const double phaseShifts[3], double dmFrac, int literalUpperPeriodicE) {
    for (int i = 0; i < 3; ++i) {
      if (n[i] < 1) throw std::runtime_error("grid needs at least one cell per direction");
      h_.g.N[i] = n[i];
      h_.g.origin[i] = origin[i];
      h_.g.d[i] = size[i] / double(n[i]);
    }
    if (mxy::numNodes(h_.g) * 3 >= (int64_t(1) << 31)) throw std::runtime_error("grid too large for 32-bit DOF ids");
    const double dx = h_.g.d[0], dy = h_.g.d[1], dz = h_.g.d[2];
    std::complex<double> phase[3];
    for (int i = 0; i < 3; ++i) phase[i] = std::exp(std::complex<double>(0.0, 1.0) * phaseShifts[i]);   // MxGridField.cpp:34-38
    complex_ = phaseShifts[0] != 0 || phaseShifts[1] != 0 || phaseShifts[2] != 0;
    for (int k = 0; k < mxy::NUM_ALL_FIELDS; ++k) {
      mxy::Field& f = h_.f[k];
      f.kind = k;
      f.ncomp = k == mxy::FIELD_PSI ? 1 : 3;
      for (auto& r : f.xi) for (double& v : r) v = 0.0;
      if (k == mxy::FIELD_E || k == mxy::FIELD_D) {  // MxYeeElecFieldBase.cpp:64-82 (D: same positions, MxYeeFitDField)
        f.xi[0][0] = 0.5 * dx; f.xi[1][1] = 0.5 * dy; f.xi[2][2] = 0.5 * dz;
      } else if (k == mxy::FIELD_B) {                // MxYeeMagFieldBase.cpp:85-116
        f.xi[0][1] = 0.5 * dy; f.xi[0][2] = 0.5 * dz;
        f.xi[1][2] = 0.5 * dz; f.xi[1][0] = 0.5 * dx;
        f.xi[2][0] = 0.5 * dx; f.xi[2][1] = 0.5 * dy;
      } else {                                       // MxYeePsiField.cpp:51-72
        f.xi[0][0] = 0.5 * dx; f.xi[0][1] = 0.5 * dy; f.xi[0][2] = 0.5 * dz;
      }
      for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 3; ++i) {
          f.lbc[c][i] = translateBC(k, lower[i], c, i);
          f.ubc[c][i] = translateBC(k, upper[i], c, i);
        }
      f.dmFrac = k == mxy::FIELD_B ? dmFrac : 0.0;
      f.regionSet = 0;
      f.literalUpperPeriodicE = literalUpperPeriodicE;
      for (int code = 0; code < 125; ++code) {       // MxGridField.cpp:145-190, direction 0 first
        const int act[3] = {code / 25, (code / 5) % 5, code % 5};
        std::complex<double> res(1.0, 0.0);
        for (int i = 0; i < 3; ++i) {
          if (act[i] == mxy::ACT_DIV) res /= phase[i];
          else if (act[i] == mxy::ACT_MUL) res *= phase[i];
          else if (act[i] == mxy::ACT_NEG) res *= -1.0;
          else if (act[i] == mxy::ACT_ZERO) res *= 0.0;
        }
        f.facRe[code] = res.real();
        f.facIm[code] = res.imag();
      }
    }
    h_.err = x_.template alloc<int>(1);
    const int zero = 0;
    x_.toExec(h_.err, &zero, sizeof(int));
    dSim_ = x_.template alloc<mxy::Sim>(1);
    pushSim();
  }

  // PEC fractions of one field on the guarded block ((N+3)^3 cells x ncomp), host array as MxGridField keeps it
  // (MxGridField.hpp:259-279). All three fields must be given before setup() when a PEC shape exists.
  void setRegionFromHost(int kind, const double* fracs) {
    mxy::Field& f = h_.f[kind];
    const int64_t n = mxy::numFullCells(h_.g) * f.ncomp;
    double* d = regionBuffer(kind);
    x_.toExec(d, fracs, size_t(n) * sizeof(double));
    pushSim();
  }
  // the executor-side fraction array of a field (filled by the fraction kernels)
  double* regionBuffer(int kind) {
    mxy::Field& f = h_.f[kind];
    if (!f.region) f.region = x_.template alloc<double>(mxy::numFullCells(h_.g) * f.ncomp);
    f.regionSet = 1;
    hasPEC_ = true;
    setUp_ = false;
    return const_cast<double*>(f.region);
  }
  void regionChanged() { pushSim(); }

  // PEC region from a shape: edge / face / cell fractions of B, E and psi computed on the executor
  // (MxEMSim.cpp:122-129 -> MxGridField::addShapeRep, MxGridField.cpp:193-226)
  void setPecShape(const Shape& shape) {
    checkShape(shape);
    ShapeNode* nodes = uploadShape(shape);
    for (int k = 0; k < mxy::NUM_FIELDS; ++k) {
      double* out = regionBuffer(k);
      pushSim();
      x_.forEach(mxy::numFullCells(h_.g), FractionCells{dSim_, nodes, k, out});
    }
    x_.sync();
    x_.free(nodes);
    pec_ = shape;
    // fractions of the dual faces are only needed next to dielectrics; drop stale ones
    x_.free(const_cast<double*>(h_.f[mxy::FIELD_D].region));
    h_.f[mxy::FIELD_D].region = nullptr;
    pushSim();
  }

  // MxEMSim::addDielectric + MxEMSim.cpp:134-148: fractions of the dielectric's shape on E, D and psi; eps = 3x3 complex
  // tensor, row-major (re, im) pairs
  void addDielectric(const Shape& shape, const double eps[18]) {
    checkShape(shape);
    if (h_.numDiel >= mxy::kMaxDielectrics) throw std::runtime_error("at most 4 dielectric objects");
    mxy::DielectricRep& D = h_.diel[h_.numDiel];
    ShapeNode* nodes = uploadShape(shape);
    const int64_t cells = mxy::numFullCells(h_.g);
    double* fe = x_.template alloc<double>(cells * 3);
    double* fd = x_.template alloc<double>(cells * 3);
    double* fp = x_.template alloc<double>(cells);
    x_.forEach(cells, FractionCells{dSim_, nodes, mxy::FIELD_E, fe});
    x_.forEach(cells, FractionCells{dSim_, nodes, mxy::FIELD_D, fd});
    x_.forEach(cells, FractionCells{dSim_, nodes, mxy::FIELD_PSI, fp});
    x_.sync();
    D.fracE = fe; D.fracD = fd; D.fracPsi = fp;
    D.shape = nodes;
    bool offDiag = false;
    for (int i = 0; i < 9; ++i) { D.epsRe[i] = eps[2 * i]; D.epsIm[i] = eps[2 * i + 1]; }
    for (int i : {1, 2, 5}) offDiag = offDiag || D.epsRe[i] != 0.0 || D.epsIm[i] != 0.0;    // MxDielectric.cpp:27-36
    D.isDiag = offDiag ? 0 : 1;
    h_.numDiel++;
    pushSim();
  }
  int numDielectrics() const { return h_.numDiel; }

  // MxGridField.cpp:256-297 for B, E and psi (B first: the others consult its rules)
  void setup() {
    if (hasPEC_)
      for (int k = 0; k < mxy::NUM_FIELDS; ++k)
        if (!h_.f[k].region) throw std::runtime_error("a PEC region needs fractions for the B, E and psi fields");
    if (hasPEC_ && h_.numDiel > 0 && !h_.f[mxy::FIELD_D].region) {     // MxEMSim.cpp:122-129: the D field joins with dielectrics
      if (pec_.nodes.empty()) throw std::runtime_error("dielectrics next to host-supplied PEC fractions need those of the D field too (dfield)");
      ShapeNode* nodes = uploadShape(pec_);
      double* out = x_.template alloc<double>(mxy::numFullCells(h_.g) * 3);
      x_.forEach(mxy::numFullCells(h_.g), FractionCells{dSim_, nodes, mxy::FIELD_D, out});
      x_.sync();
      x_.free(nodes);
      h_.f[mxy::FIELD_D].region = out;
    }
    h_.f[mxy::FIELD_D].regionSet = hasPEC_ ? 1 : 0;
    for (int k = 0; k < mxy::NUM_FIELDS; ++k) {
      mxy::Field& f = h_.f[k];
      x_.free(const_cast<int32_t*>(f.lidOf));
      x_.free(const_cast<int64_t*>(f.gids));
      f.lidOf = nullptr; f.gids = nullptr; f.nLoc = 0;
      const int64_t n = mxy::numNodes(h_.g) * f.ncomp;
      Scoped<X, int32_t> flag(x_, n);
      Scoped<X, int64_t> off(x_, n + 1);
      x_.forEach(n, mxy::MapFlags{dSim_, k, flag.p});
      const int64_t total = x_.scan(flag.p, off.p, n);
      Scoped<X, int32_t> lid(x_, n);
      Scoped<X, int64_t> gids(x_, total > 0 ? total : 1);
      x_.forEach(n, mxy::MapFill{flag.p, off.p, lid.p, gids.p});
      x_.sync();
      f.lidOf = lid.release();
      f.gids = gids.release();
      f.nLoc = total;
      pushSim();
    }
    h_.f[mxy::FIELD_D].lidOf = h_.f[mxy::FIELD_E].lidOf;      // MxEMSim.cpp:186-190: D shares E's map
    h_.f[mxy::FIELD_D].gids = h_.f[mxy::FIELD_E].gids;
    h_.f[mxy::FIELD_D].nLoc = h_.f[mxy::FIELD_E].nLoc;
    pushSim();
    checkErr("map set-up");
    setUp_ = true;
  }

  int64_t mapSize(int kind) const { return h_.f[kind].nLoc; }
  int64_t numGlobal(int kind) const { return mxy::numNodes(h_.g) * h_.f[kind].ncomp; }
  void copyMap(int kind, int64_t* out) { x_.toHost(out, h_.f[kind].gids, size_t(h_.f[kind].nLoc) * sizeof(int64_t)); }
  void copyFractions(int kind, double* out) {
    const mxy::Field& f = h_.f[kind];
    if (!f.region) throw std::runtime_error("field has no PEC fractions");
    x_.toHost(out, f.region, size_t(mxy::numFullCells(h_.g) * f.ncomp) * sizeof(double));
  }

  // ---- CRS algebra ------------------------------------------------------------------------------------------
  template <class S>
  void destroy(Csr<S>& m) {
    x_.free(m.rowptr); x_.free(m.col); x_.free(m.val);
    m = Csr<S>();
  }

  template <class S, class RowFn>
  Csr<S> buildRows(int64_t nrows, int64_t ncols, const RowFn& fn) {
    Csr<S> m;
    m.nrows = nrows; m.ncols = ncols;
    Scoped<X, int64_t> rowptr(x_, nrows + 1);
    {
      Scoped<X, int32_t> cnt(x_, nrows > 0 ? nrows : 1);
      x_.forEach(nrows, mxy::CountRows<S, RowFn>{fn, cnt.p});
      m.nnz = x_.scan(cnt.p, rowptr.p, nrows);
    }
    Scoped<X, int32_t> col(x_, m.nnz > 0 ? m.nnz : 1);
    Scoped<X, S> val(x_, m.nnz > 0 ? m.nnz : 1);
    x_.forEach(nrows, mxy::FillRows<S, RowFn>{fn, rowptr.p, col.p, val.p});
    m.rowptr = rowptr.release();
    m.col = col.release();
    m.val = val.release();
    return m;
  }

  template <class Fn>
  int maxOverRows(int64_t nrows, const Fn& fn) {
    if (nrows == 0) return 0;
    Scoped<X, int32_t> tmp(x_, nrows);
    x_.forEach(nrows, StoreInt<Fn>{fn, tmp.p});
    return x_.maxOf(tmp.p, nrows);
  }

  template <class S>
  Csr<S> multiply(const Csr<S>& A, const Csr<S>& B) {       // MxCrsMatrix.cpp:358-382
    if (A.ncols != B.nrows) throw std::runtime_error("multiply: shapes do not match");
    const int bound = maxOverRows(A.nrows, mxy::ProductBound<S>{A.view(), B.view()});
    Csr<S> C;
    if (bound <= 16) C = buildRows<S>(A.nrows, B.ncols, mxy::ProductRow<S, 16>{A.view(), B.view()});
    else if (bound <= 32) C = buildRows<S>(A.nrows, B.ncols, mxy::ProductRow<S, 32>{A.view(), B.view()});
    else if (bound <= 64) C = buildRows<S>(A.nrows, B.ncols, mxy::ProductRow<S, 64>{A.view(), B.view()});
    else if (bound <= 160) C = buildRows<S>(A.nrows, B.ncols, mxy::ProductRow<S, 160>{A.view(), B.view()});
    else if (bound <= 400) C = buildRows<S>(A.nrows, B.ncols, mxy::ProductRow<S, 400>{A.view(), B.view()});
    else throw std::runtime_error("multiply: a product row would exceed 400 entries");
    C.rowField = A.rowField;
    C.colField = B.colField;
    return C;
  }

  template <class S>
  Csr<S> add(const Csr<S>& A, S sa, const Csr<S>& B, S sb, bool purge) {   // MxCrsMatrix.cpp:401-430 (+ :84-117)
    if (A.nrows != B.nrows || A.ncols != B.ncols) throw std::runtime_error("add: shapes do not match");
    const int bound = maxOverRows(A.nrows, RowLen2{A.rowptr, B.rowptr});
    Csr<S> C;
    const int p = purge ? 1 : 0;
    if (bound <= 32) C = buildRows<S>(A.nrows, A.ncols, mxy::SumRow<S, 32>{A.view(), B.view(), sa, sb, p});
    else if (bound <= 160) C = buildRows<S>(A.nrows, A.ncols, mxy::SumRow<S, 160>{A.view(), B.view(), sa, sb, p});
    else if (bound <= 800) C = buildRows<S>(A.nrows, A.ncols, mxy::SumRow<S, 800>{A.view(), B.view(), sa, sb, p});
    else throw std::runtime_error("add: a row would exceed 800 entries");
    C.rowField = A.rowField;
    C.colField = A.colField;
    return C;
  }

  template <class S>
  Csr<S> purgeZeros(const Csr<S>& A) {                      // MxCrsMatrix.cpp:84-117
    const int bound = maxOverRows(A.nrows, RowLen2{A.rowptr, nullptr});
    Csr<S> C;
    if (bound <= 32) C = buildRows<S>(A.nrows, A.ncols, mxy::PurgeRow<S, 32>{A.view()});
    else if (bound <= 400) C = buildRows<S>(A.nrows, A.ncols, mxy::PurgeRow<S, 400>{A.view()});
    else throw std::runtime_error("purge: a row exceeds 400 entries");
    C.rowField = A.rowField;
    C.colField = A.colField;
    return C;
  }

  template <class S>
  void scale(Csr<S>& A, S s) { x_.forEach(A.nnz, ScaleVals<S>{A.val, s}); }

  // host CSR (local column indices) -> executor; used for operators that are still generated on the host
  // (the dielectric invEps of MxYeeFitInvEps) so that the chains around them run here
  template <class S>
  Csr<S> upload(int64_t nrows, int64_t ncols, const int64_t* rowptr, const int32_t* col, const S* val, int rowField, int colField) {
    Csr<S> m;
    m.nrows = nrows; m.ncols = ncols; m.nnz = rowptr[nrows];
    m.rowField = rowField; m.colField = colField;
    m.rowptr = x_.template alloc<int64_t>(nrows + 1);
    m.col = x_.template alloc<int32_t>(m.nnz > 0 ? m.nnz : 1);
    m.val = x_.template alloc<S>(m.nnz > 0 ? m.nnz : 1);
    x_.toExec(m.rowptr, rowptr, size_t(nrows + 1) * sizeof(int64_t));
    x_.toExec(m.col, col, size_t(m.nnz) * sizeof(int32_t));
    x_.toExec(m.val, val, size_t(m.nnz) * sizeof(S));
    return m;
  }

  // (conjugate) transpose, rows sorted by column
  template <class S>
  Csr<S> transpose(const Csr<S>& A) {
    Csr<S> T;
    T.nrows = A.ncols; T.ncols = A.nrows; T.nnz = A.nnz;
    T.rowField = A.colField; T.colField = A.rowField;
    Scoped<X, int32_t> cnt(x_, T.nrows > 0 ? T.nrows : 1);
    Scoped<X, int64_t> rowptr(x_, T.nrows + 1);
    Scoped<X, int32_t> col(x_, T.nnz > 0 ? T.nnz : 1);
    Scoped<X, S> val(x_, T.nnz > 0 ? T.nnz : 1);
    x_.zero(cnt.p, size_t(T.nrows) * sizeof(int32_t));
    x_.forEach(A.nnz, mxy::TransposeCount{A.col, cnt.p});
    x_.scan(cnt.p, rowptr.p, T.nrows);
    x_.zero(cnt.p, size_t(T.nrows) * sizeof(int32_t));
    x_.forEach(A.nrows, mxy::TransposeFill<S>{A.view(), rowptr.p, cnt.p, col.p, val.p});
    x_.forEach(T.nrows, mxy::SortRowInPlace<S>{rowptr.p, col.p, val.p});
    x_.sync();                       // cnt is released on return
    T.rowptr = rowptr.release();
    T.col = col.release();
    T.val = val.release();
    return T;
  }

  // MxGridFieldInterpolator.cpp:28-122: `field` of the simulation `from` interpolated at the DOFs of this simulation
  // (rows: this simulation's map, columns: from's map). Both live on the same executor.
  template <class S>
  Csr<S> interpolatorFrom(Assembler& from, int field) {
    requireSetUp();
    from.requireSetUp();
    Csr<S> m = buildRows<S>(h_.f[field].nLoc, from.h_.f[field].nLoc, mxy::InterpRow<S>{from.dSim_, dSim_, field});
    m.rowField = field;
    m.colField = field;
    return m;
  }

  // ---- generators and chains ---------------------------------------------------------------------------------
  template <class S>
  Csr<S> generate(int op, int fracField = 0, bool inverse = false, double minFrac = 0.0) {
    requireSetUp();
    static const int rowF[5] = {mxy::FIELD_B, mxy::FIELD_E, mxy::FIELD_PSI, mxy::FIELD_B, -1};
    static const int colF[5] = {mxy::FIELD_E, mxy::FIELD_B, mxy::FIELD_B, mxy::FIELD_PSI, -1};
    const int rf = op == mxy::GEN_FRACS ? fracField : rowF[op];
    const int cf = op == mxy::GEN_FRACS ? fracField : colF[op];
    mxy::GenRow<S> fn{dSim_, op, fracField, inverse ? 1 : 0, minFrac};
    Csr<S> m = buildRows<S>(h_.f[rf].nLoc, h_.f[cf].nLoc, fn);
    m.rowField = rf;
    m.colField = cf;
    checkErr("operator generation");
    return m;
  }

  // a diagonal of ones when there is no PEC shape (MxCrsMatrix.cpp:31-45 via MxMagWaveOp.cpp:227-241)
  template <class S>
  Csr<S> fracsOrIdentity(int field, bool inverse, double minFrac) {
    return generate<S>(mxy::GEN_FRACS, field, inverse, minFrac);   // without a region every fraction reads 1
  }

  // MxEMOps.cpp:39-168 and MxMagWaveOp.cpp:137-245. invEps / invEpsVolAve (dielectrics) are optional host-generated
  // factors handed in by the caller.
  template <class S>
  Csr<S> buildOp(const std::string& name, const Csr<S>* invEps = nullptr, const Csr<S>* invEpsVolAve = nullptr) {
    using namespace mxy;
    if (name == "curlE") return generate<S>(GEN_CURL_E);
    if (name == "curlB") return generate<S>(GEN_CURL_B);
    if (name == "divB") return generate<S>(GEN_DIV_B);
    if (name == "gradPsi") return generate<S>(GEN_GRAD_PSI);
    if (name == "dmA") return generate<S>(GEN_FRACS, FIELD_B, false, 0.e-12);     // MxEMOps.cpp:56-58
    if (name == "dmL") return generate<S>(GEN_FRACS, FIELD_E, false, 0.e-6);      // MxEMOps.cpp:61-63
    if (name == "dmVInv") return generate<S>(GEN_FRACS, FIELD_PSI, true, 0.e-6);  // MxEMOps.cpp:125-127
    if (name == "mRhs") return generate<S>(GEN_FRACS, FIELD_B, false, 0.e-12);    // MxMagWaveOp.cpp:227-241
    if (name == "invEps" || name == "invEpsVolAve") {                              // MxEMOps.cpp:72-81
      requireSetUp();
      if (h_.numDiel == 0) throw std::runtime_error("the simulation has no dielectric objects");
      Csr<S> m;
      if (name == "invEps") {
        m = buildRows<S>(h_.f[FIELD_E].nLoc, h_.f[FIELD_E].nLoc, InvEpsRow<S>{dSim_, hasPEC_ ? 1 : 0});
        m.rowField = m.colField = FIELD_E;
      } else {
        m = buildRows<S>(h_.f[FIELD_PSI].nLoc, h_.f[FIELD_PSI].nLoc, InvEpsVolAveRow<S>{dSim_});
        m.rowField = m.colField = FIELD_PSI;
      }
      checkErr("inverse permittivity");
      return m;
    }
    // dielectric objects of the simulation supply the factors the caller did not hand in
    Csr<S> ownEps, ownVol;
    struct Owned {
      Assembler* a; Csr<S>*x, *y;
      ~Owned() { a->destroy(*x); a->destroy(*y); }
    } owned{this, &ownEps, &ownVol};
    if (h_.numDiel > 0 && (name == "curlCurl" || name == "vecLapl") && !invEps) { ownEps = buildOp<S>("invEps"); invEps = &ownEps; }
    if (h_.numDiel > 0 && (name == "gradDiv" || name == "vecLapl") && !invEpsVolAve) { ownVol = buildOp<S>("invEpsVolAve"); invEpsVolAve = &ownVol; }
    if (name == "curlCurl") {                                                      // MxMagWaveOp.cpp:144-153
      Csr<S> m = generate<S>(GEN_CURL_B);
      if (invEps) replaceWith(m, multiply(*invEps, m));
      if (hasPEC_) {
        Csr<S> d = generate<S>(GEN_FRACS, FIELD_E, false, 0.e-6);
        replaceWith(m, multiply(d, m));
        destroy(d);
      }
      Csr<S> ce = generate<S>(GEN_CURL_E);
      replaceWith(m, multiply(ce, m));
      destroy(ce);
      return m;
    }
    if (name == "gradDiv") {                                                       // MxMagWaveOp.cpp:156-179
      Csr<S> m = generate<S>(GEN_DIV_B);
      if (hasPEC_) {
        Csr<S> a = generate<S>(GEN_FRACS, FIELD_B, false, 0.e-12);
        replaceWith(m, multiply(m, a));
        Csr<S> v = generate<S>(GEN_FRACS, FIELD_PSI, true, 0.e-6);
        replaceWith(m, multiply(v, m));
        destroy(v);
        destroy(a);
      }
      if (invEpsVolAve) replaceWith(m, multiply(*invEpsVolAve, m));
      Csr<S> gp = generate<S>(GEN_GRAD_PSI);
      replaceWith(m, multiply(gp, m));
      destroy(gp);
      if (hasPEC_) {
        Csr<S> a = generate<S>(GEN_FRACS, FIELD_B, false, 0.e-12);
        replaceWith(m, multiply(a, m));
        destroy(a);
      }
      return m;
    }
    if (name == "vecLapl") {                                                       // MxMagWaveOp.cpp:183-205
      Csr<S> cc = buildOp<S>("curlCurl", invEps, invEpsVolAve);
      Csr<S> gd = buildOp<S>("gradDiv", invEps, invEpsVolAve);
      Csr<S> m = add(cc, scalarFrom<S>(1.0, 0.0), gd, scalarFrom<S>(-1.0, 0.0), true);
      destroy(cc);
      destroy(gd);
      return m;
    }
    if (name == "scaLapl") {                                                       // MxMagWaveOp.cpp:208-223
      Csr<S> m = generate<S>(GEN_GRAD_PSI);
      if (hasPEC_) {
        Csr<S> a = generate<S>(GEN_FRACS, FIELD_B, false, 0.e-12);
        replaceWith(m, multiply(a, m));
        destroy(a);
      }
      Csr<S> db = generate<S>(GEN_DIV_B);
      replaceWith(m, multiply(db, m));
      destroy(db);
      scale(m, scalarFrom<S>(-1.0, 0.0));
      return m;
    }
    throw std::runtime_error("unknown operator '" + name + "'");
  }

  bool complexByDefault() const { return complex_; }
  mxy::Sim* simOnExec() { return dSim_; }
  void checkErr(const char* what) {
    int e = 0;
    x_.toHost(&e, h_.err, sizeof(int));
    if (e) {
      const int zero = 0;
      x_.toExec(h_.err, &zero, sizeof(int));
      throw std::runtime_error(std::string(what) + (e == 1 ? ": a stencil left the guarded block" : ": a column is not in the domain map"));
    }
  }

 private:
  static int translateBC(int kind, int bc, int comp, int dir) {
    using namespace mxy;
    if (bc != PEC && bc != PMC) return bc;
    const bool normal = (comp == dir);
    if (kind == FIELD_PSI) return bc == PEC ? CONSTANT : ZERO;                                    // MxYeePsiField.cpp:84-93
    if (kind == FIELD_B) return bc == PEC ? (normal ? ZERO : CONSTANT) : (normal ? CONSTANT : ZERO);   // MxYeeMagFieldBase.cpp:134-163
    return bc == PEC ? (normal ? CONSTANT : ZERO) : (normal ? ZERO : CONSTANT);                   // MxYeeElecFieldBase.cpp:104-133
  }
  void pushSim() { x_.toExec(dSim_, &h_, sizeof(mxy::Sim)); }
  static void checkShape(const Shape& shape) {
    if (shape.nodes.empty()) throw std::runtime_error("empty shape");
    if (shape.depth() > kShapeMaxDepth) throw std::runtime_error("shape tree deeper than 8 levels");
  }
  ShapeNode* uploadShape(const Shape& shape) {
    ShapeNode* d = x_.template alloc<ShapeNode>(int64_t(shape.nodes.size()));
    x_.toExec(d, shape.nodes.data(), shape.nodes.size() * sizeof(ShapeNode));
    return d;
  }
  void requireSetUp() const {
    if (!setUp_) throw std::runtime_error("the simulation has not been set up (mxg_sim_setup)");
  }
  template <class S>
  void replaceWith(Csr<S>& m, Csr<S> next) {
    destroy(m);
    m = next;
  }

  X& x_;
  mxy::Sim h_;            // host copy; its pointers are executor pointers
  mxy::Sim* dSim_ = nullptr;
  Shape pec_;             // kept for the D-field fractions a later dielectric needs
  bool hasPEC_ = false, setUp_ = false, complex_ = false;
};

}  // namespace mxa
