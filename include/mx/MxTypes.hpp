// C++ shims over the libmxgpu C ABI with the reference's class and method names, so host code
// written against bauerca/maxwell's linear-algebra layer (src/MxMap.hpp, MxMultiVector.hpp,
// MxAnasaziMV.hpp, MxCrsMatrix.hpp, MxOperator.hpp) compiles against the B200 path unchanged
// in shape. Header-only; link with -lmxgpu.
//
// Where the reference derives from Anasazi::MultiVec / Anasazi::Operator and passes
// Teuchos::SerialDenseMatrix, these headers use the minimal look-alikes below (Trilinos is not
// available in this build). Define MX_HAVE_ANASAZI to derive from the real Anasazi classes
// instead.
#pragma once
#include <complex>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "mxgpu.h"

typedef std::complex<double> MxComplex;   // reference: src/MxTypes.h
typedef int64_t MxIndex;

namespace mx {

// error convention: the reference prints and exit()s or throws 1 (MxAnasaziMV.cpp:20-23,
// MxCrsMatrix.cpp:89-91); the shims throw std::runtime_error carrying mxg_last_error().
inline void check(int rc) {
  if (rc != 0) throw std::runtime_error(mxg_last_error());
}

template <class Scalar> struct ScalarTraits;
template <> struct ScalarTraits<double> {
  static constexpr bool isComplex = false;
  static void pack(double s, double out[2]) { out[0] = s; out[1] = 0.0; }
  static double unpack(const double* p) { return p[0]; }
  static double conj(double s) { return s; }
  static double one() { return 1.0; }
  static double zero() { return 0.0; }
};
template <> struct ScalarTraits<MxComplex> {
  static constexpr bool isComplex = true;
  static void pack(MxComplex s, double out[2]) { out[0] = s.real(); out[1] = s.imag(); }
  static MxComplex unpack(const double* p) { return MxComplex(p[0], p[1]); }
  static MxComplex conj(MxComplex s) { return std::conj(s); }
  static MxComplex one() { return MxComplex(1.0, 0.0); }
  static MxComplex zero() { return MxComplex(0.0, 0.0); }
};

#ifdef MX_HAVE_ANASAZI
}  // namespace mx
// With Trilinos present the shims derive from the real interfaces: MxAnasaziMV IS an Anasazi::MultiVec and
// can be handed to BlockKrylovSchurSolMgr / BlockDavidsonSolMgr exactly as in src/MxSolver.cpp:62-94. (Trilinos is absent from
// this build container, so this branch is not compiled here.)
#include "AnasaziMultiVec.hpp"
#include "AnasaziOperator.hpp"
#include "Teuchos_SerialDenseMatrix.hpp"
namespace mx {
template <class Ordinal, class Scalar> using SerialDenseMatrix = Teuchos::SerialDenseMatrix<Ordinal, Scalar>;
template <class Scalar> using MultiVec = Anasazi::MultiVec<Scalar>;
template <class Scalar> using Operator = Anasazi::Operator<Scalar>;
#else
// Look-alike of Teuchos::SerialDenseMatrix<int, Scalar> (column-major host matrix): the subset
// the reference touches -- numRows/numCols/values/stride/operator() (MxAnasaziMV.cpp:12-14,45-52).
template <class Ordinal, class Scalar>
class SerialDenseMatrix {
 public:
  SerialDenseMatrix() : r_(0), c_(0) {}
  SerialDenseMatrix(Ordinal rows, Ordinal cols) : r_(rows), c_(cols), v_(size_t(rows) * cols, Scalar(0)) {}
  void shape(Ordinal rows, Ordinal cols) { r_ = rows; c_ = cols; v_.assign(size_t(rows) * cols, Scalar(0)); }
  Ordinal numRows() const { return r_; }
  Ordinal numCols() const { return c_; }
  Ordinal stride() const { return r_; }
  Scalar* values() { return v_.data(); }
  const Scalar* values() const { return v_.data(); }
  Scalar& operator()(Ordinal i, Ordinal j) { return v_[size_t(j) * r_ + i]; }
  const Scalar& operator()(Ordinal i, Ordinal j) const { return v_[size_t(j) * r_ + i]; }
  void putScalar(Scalar s) { std::fill(v_.begin(), v_.end(), s); }

 private:
  Ordinal r_, c_;
  std::vector<Scalar> v_;
};

// Look-alike of Anasazi::MultiVec<Scalar> (the virtuals MxAnasaziMV implements,
// src/MxAnasaziMV.hpp:23-136).
template <class Scalar>
class MultiVec {
 public:
  virtual ~MultiVec() {}
  virtual MultiVec<Scalar>* Clone(const int numVecs) const = 0;
  virtual MultiVec<Scalar>* CloneCopy() const = 0;
  virtual MultiVec<Scalar>* CloneCopy(const std::vector<int>& index) const = 0;
  virtual const MultiVec<Scalar>* CloneView(const std::vector<int>& index) const = 0;
  virtual MultiVec<Scalar>* CloneViewNonConst(const std::vector<int>& index) = 0;
  virtual int GetVecLength() const = 0;
  virtual int GetNumberVecs() const = 0;
  virtual void MvTimesMatAddMv(Scalar alpha, const MultiVec<Scalar>& A, const SerialDenseMatrix<int, Scalar>& B, Scalar beta) = 0;
  virtual void MvAddMv(Scalar alpha, const MultiVec<Scalar>& A, Scalar beta, const MultiVec<Scalar>& B) = 0;
  virtual void MvTransMv(Scalar alpha, const MultiVec<Scalar>& A, SerialDenseMatrix<int, Scalar>& B) const = 0;
  virtual void MvDot(const MultiVec<Scalar>& A, std::vector<Scalar>& b) const = 0;
  virtual void MvNorm(std::vector<double>& normvec) const = 0;
  virtual void SetBlock(const MultiVec<Scalar>& A, const std::vector<int>& index) = 0;
  virtual void MvScale(Scalar alpha) = 0;
  virtual void MvScale(const std::vector<Scalar>& alpha) = 0;
  virtual void MvRandom() = 0;
  virtual void MvInit(Scalar alpha) = 0;
  virtual void MvPrint(std::ostream& os) const = 0;
};

// Look-alike of Anasazi::Operator<Scalar> (src/MxMagWaveOp.h:79-80).
template <class Scalar>
class Operator {
 public:
  virtual ~Operator() {}
  virtual void Apply(const MultiVec<Scalar>& x, MultiVec<Scalar>& y) const = 0;
};

#endif  // MX_HAVE_ANASAZI

// A constraint the eigensolver keeps its search space in (MxSolverT::setConstraint): the divergence-free subspace.
template <class S>
class Constraint {
 public:
  virtual ~Constraint() {}
  // b <- P b in place; the inner solve reduces its residual by `tol` or stops after maxIters iterations (0 = no cap)
  virtual void project(MultiVec<S>& b, double tol, int maxIters = 0) const = 0;
  // out[j] = |D M b_j|_2 given Mb = M b (the reference's checkDivergences, MxMagWaveOp.cpp:1211-1234)
  virtual void violation(const MultiVec<S>& Mb, std::vector<double>& out) const = 0;
};

}  // namespace mx
