// Host-side eigensolver driver for the B200 path: the counterpart of src/MxSolver.{h,cpp} +
// src/MxMagWaveOp.{h,cpp} + src/MxGeoMultigridPrec.{h,cpp} of the reference.
//
// The reference hands an Anasazi::Operator (MxMagWaveOp) and Anasazi::MultiVec objects
// (MxAnasaziMV) to a Trilinos solver manager (MxSolver.cpp:62-103). Anasazi is not available
// here, so MxSolver below is a small block eigensolver (LOBPCG with soft locking) written ONLY
// against that same MultiVecTraits / OperatorTraits surface: Clone / CloneView / MvTransMv /
// MvTimesMatAddMv / MvAddMv / MvNorm / MvScale / SetBlock and Operator::Apply. Everything
// O(n) runs on the GPU through libmxgpu; the host only sees k x b dense matrices.
#pragma once
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <functional>
#include <numeric>

#include "MxLinAlg.hpp"

// ---- MxGeoMultigridPrec (src/MxGeoMultigridPrec.h): V-cycle / FMG preconditioner -------------------
template <class Scalar>
class MxGeoMultigridPrec : public mx::Operator<Scalar> {
 public:
  MxGeoMultigridPrec(std::shared_ptr<MxComm> comm, const std::vector<mxg_crs*>& ops, const std::vector<mxg_crs*>& restrictors,
                     const std::vector<mxg_crs*>& prolongators, const mxg_gmg_params* params = nullptr) {
    mx::check(mxg_gmg_create(comm->raw(), int(ops.size()), ops.data(), restrictors.data(), prolongators.data(), params, &gmg_));
  }
  MxGeoMultigridPrec(mxg_gmg* raw, bool own) : gmg_(raw), own_(own) {}
  ~MxGeoMultigridPrec() { if (own_ && gmg_) mxg_gmg_destroy(gmg_); }
  // Epetra_Operator::ApplyInverse of the reference (MxGeoMultigridPrec.cpp:496-542)
  void ApplyInverse(const MxMultiVector<Scalar>& b, MxMultiVector<Scalar>& x) const { mx::check(mxg_gmg_apply(gmg_, b.getRawMV(), x.getRawMV())); }
  void Apply(const mx::MultiVec<Scalar>& x, mx::MultiVec<Scalar>& y) const override {
    ApplyInverse(dynamic_cast<const MxAnasaziMV<Scalar>&>(x), dynamic_cast<MxAnasaziMV<Scalar>&>(y));
  }
  mxg_gmg* raw() const { return gmg_; }

 private:
  mxg_gmg* gmg_ = nullptr;
  bool own_ = true;
};

// An assembled CRS operator as an Anasazi-style Operator
template <class Scalar>
class MxCrsOperator : public mx::Operator<Scalar> {
 public:
  explicit MxCrsOperator(mxg_crs* A) : A_(A) {}
  void Apply(const mx::MultiVec<Scalar>& x, mx::MultiVec<Scalar>& y) const override {
    mx::check(mxg_crs_apply(A_, dynamic_cast<const MxAnasaziMV<Scalar>&>(x).getRawMV(), dynamic_cast<MxAnasaziMV<Scalar>&>(y).getRawMV()));
  }

 private:
  mxg_crs* A_;
};

// A diagonal operator (mRhs = dmA, MxMagWaveOp.cpp:227-241)
template <class Scalar>
class MxDiagOperator : public mx::Operator<Scalar> {
 public:
  explicit MxDiagOperator(mxg_mv* d) : d_(d) {}
  void Apply(const mx::MultiVec<Scalar>& x, mx::MultiVec<Scalar>& y) const override {
    mx::check(mxg_mv_diag_mult(dynamic_cast<MxAnasaziMV<Scalar>&>(y).getRawMV(), d_, dynamic_cast<const MxAnasaziMV<Scalar>&>(x).getRawMV()));
  }

 private:
  mxg_mv* d_;
};

namespace mx {
namespace dense {
// Small dense symmetric helpers (column-major, double). The reference delegates these to
// Teuchos::LAPACK inside Anasazi; sizes here are at most 3 x block (<= ~120).
inline bool cholesky(std::vector<double>& a, int n) {  // lower factor in place; false if not SPD
  for (int j = 0; j < n; ++j) {
    double d = a[j + size_t(j) * n];
    for (int k = 0; k < j; ++k) d -= a[j + size_t(k) * n] * a[j + size_t(k) * n];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    a[j + size_t(j) * n] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = a[i + size_t(j) * n];
      for (int k = 0; k < j; ++k) s -= a[i + size_t(k) * n] * a[j + size_t(k) * n];
      a[i + size_t(j) * n] = s / d;
    }
    for (int i = 0; i < j; ++i) a[i + size_t(j) * n] = 0.0;
  }
  return true;
}
// inverse of a lower-triangular matrix
inline std::vector<double> invLower(const std::vector<double>& L, int n) {
  std::vector<double> X(size_t(n) * n, 0.0);
  for (int j = 0; j < n; ++j) {
    X[j + size_t(j) * n] = 1.0 / L[j + size_t(j) * n];
    for (int i = j + 1; i < n; ++i) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s += L[i + size_t(k) * n] * X[k + size_t(j) * n];
      X[i + size_t(j) * n] = -s / L[i + size_t(i) * n];
    }
  }
  return X;
}
// cyclic Jacobi for a symmetric matrix: a -> eigenvalues (ascending) in w, eigenvectors in v
inline void symEig(std::vector<double> a, int n, std::vector<double>& w, std::vector<double>& v) {
  v.assign(size_t(n) * n, 0.0);
  for (int i = 0; i < n; ++i) v[i + size_t(i) * n] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) (i == j ? diag : off) += a[i + size_t(j) * n] * a[i + size_t(j) * n];
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = a[p + size_t(q) * n];
        if (apq == 0.0) continue;
        const double app = a[p + size_t(p) * n], aqq = a[q + size_t(q) * n];
        const double tau = (aqq - app) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < n; ++k) {  // columns p, q
          const double akp = a[k + size_t(p) * n], akq = a[k + size_t(q) * n];
          a[k + size_t(p) * n] = c * akp - s * akq;
          a[k + size_t(q) * n] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // rows p, q
          const double apk = a[p + size_t(k) * n], aqk = a[q + size_t(k) * n];
          a[p + size_t(k) * n] = c * apk - s * aqk;
          a[q + size_t(k) * n] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = v[k + size_t(p) * n], vkq = v[k + size_t(q) * n];
          v[k + size_t(p) * n] = c * vkp - s * vkq;
          v[k + size_t(q) * n] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(), [&](int i, int j) { return a[i + size_t(i) * n] < a[j + size_t(j) * n]; });
  w.resize(n);
  std::vector<double> vs(size_t(n) * n);
  for (int j = 0; j < n; ++j) {
    w[j] = a[idx[j] + size_t(idx[j]) * n];
    for (int i = 0; i < n; ++i) vs[i + size_t(j) * n] = v[i + size_t(idx[j]) * n];
  }
  v.swap(vs);
}
// C = op(A) * B, A is n x n (transA: use A^T), B n x m
inline std::vector<double> mul(const std::vector<double>& A, bool transA, const std::vector<double>& B, int n, int m) {
  std::vector<double> C(size_t(n) * m, 0.0);
  for (int j = 0; j < m; ++j)
    for (int k = 0; k < n; ++k) {
      const double b = B[k + size_t(j) * n];
      if (b == 0.0) continue;
      for (int i = 0; i < n; ++i) C[i + size_t(j) * n] += (transA ? A[k + size_t(i) * n] : A[i + size_t(k) * n]) * b;
    }
  return C;
}
// generalized symmetric-definite problem A z = w B z; returns false if B is not numerically SPD
inline bool genSymEig(const std::vector<double>& A, const std::vector<double>& B, int n, std::vector<double>& w, std::vector<double>& Z) {
  std::vector<double> L = B;
  if (!cholesky(L, n)) return false;
  const std::vector<double> Li = invLower(L, n);
  // H = Li * A * Li^T
  std::vector<double> T = mul(Li, false, A, n, n);
  std::vector<double> H(size_t(n) * n, 0.0);
  for (int j = 0; j < n; ++j)
    for (int k = 0; k < n; ++k) {
      const double l = Li[j + size_t(k) * n];
      if (l == 0.0) continue;
      for (int i = 0; i < n; ++i) H[i + size_t(j) * n] += T[i + size_t(k) * n] * l;
    }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < j; ++i) {
      const double s = 0.5 * (H[i + size_t(j) * n] + H[j + size_t(i) * n]);
      H[i + size_t(j) * n] = H[j + size_t(i) * n] = s;
    }
  std::vector<double> V;
  symEig(H, n, w, V);
  Z = mul(Li, true, V, n, n);  // Z = Li^T V
  return true;
}
// Same problem, robust to a nearly dependent basis: B = D^1/2 (V L V^T) D^1/2, directions with
// L_i <= eps * L_max are discarded, A is projected on the rest. Z has n rows and `kept` columns.
inline int genSymEigRobust(const std::vector<double>& A, const std::vector<double>& B, int n, double eps,
                           std::vector<double>& w, std::vector<double>& Z) {
  std::vector<double> d(n), G(size_t(n) * n);
  for (int j = 0; j < n; ++j) d[j] = B[j + size_t(j) * n] > 0 ? 1.0 / std::sqrt(B[j + size_t(j) * n]) : 0.0;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) G[i + size_t(j) * n] = B[i + size_t(j) * n] * d[i] * d[j];
  std::vector<double> lam, V;
  symEig(G, n, lam, V);
  std::vector<int> keep;
  for (int j = 0; j < n; ++j)
    if (lam[j] > eps * lam[n - 1]) keep.push_back(j);
  const int k = int(keep.size());
  if (k == 0) return 0;
  std::vector<double> Q(size_t(n) * k);   // B-orthonormal basis of the kept subspace
  for (int j = 0; j < k; ++j) {
    const double s = 1.0 / std::sqrt(lam[keep[j]]);
    for (int i = 0; i < n; ++i) Q[i + size_t(j) * n] = d[i] * V[i + size_t(keep[j]) * n] * s;
  }
  std::vector<double> AQ = mul(A, false, Q, n, k), H(size_t(k) * k, 0.0);
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < k; ++i) {
      double s = 0.0;
      for (int r = 0; r < n; ++r) s += Q[r + size_t(i) * n] * AQ[r + size_t(j) * n];
      H[i + size_t(j) * k] = s;
    }
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < j; ++i) H[i + size_t(j) * k] = H[j + size_t(i) * k] = 0.5 * (H[i + size_t(j) * k] + H[j + size_t(i) * k]);
  std::vector<double> Y;
  symEig(H, k, w, Y);
  Z.assign(size_t(n) * k, 0.0);
  for (int j = 0; j < k; ++j)
    for (int c = 0; c < k; ++c) {
      const double y = Y[c + size_t(j) * k];
      for (int i = 0; i < n; ++i) Z[i + size_t(j) * n] += Q[i + size_t(c) * n] * y;
    }
  return k;
}
// ---- Hermitian (complex) counterparts, used by the complex instantiation of the eigensolver driver --------------
typedef std::complex<double> cplx;
// cyclic Jacobi with complex rotations: a (Hermitian) -> real eigenvalues (ascending) in w, unitary eigenvectors in v.
// Rotation in the (p,q) plane: U_pp = U_qq = c, U_pq = s e, U_qp = -s conj(e), e = a_pq / |a_pq|; a <- U^H a U.
inline void hermEig(std::vector<cplx> a, int n, std::vector<double>& w, std::vector<cplx>& v) {
  v.assign(size_t(n) * n, cplx(0.0));
  for (int i = 0; i < n; ++i) v[i + size_t(i) * n] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) (i == j ? diag : off) += std::norm(a[i + size_t(j) * n]);
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const cplx apq = a[p + size_t(q) * n];
        const double mag = std::abs(apq);
        if (mag == 0.0) continue;
        const cplx e = apq / mag;
        const double app = a[p + size_t(p) * n].real(), aqq = a[q + size_t(q) * n].real();
        const double tau = (aqq - app) / (2.0 * mag);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), sn = t * c;
        const cplx se = sn * e, sec = sn * std::conj(e);
        for (int k = 0; k < n; ++k) {  // columns p, q:  a <- a U
          const cplx akp = a[k + size_t(p) * n], akq = a[k + size_t(q) * n];
          a[k + size_t(p) * n] = c * akp - sec * akq;
          a[k + size_t(q) * n] = se * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // rows p, q:  a <- U^H a
          const cplx apk = a[p + size_t(k) * n], aqk = a[q + size_t(k) * n];
          a[p + size_t(k) * n] = c * apk - se * aqk;
          a[q + size_t(k) * n] = sec * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {  // v <- v U
          const cplx vkp = v[k + size_t(p) * n], vkq = v[k + size_t(q) * n];
          v[k + size_t(p) * n] = c * vkp - sec * vkq;
          v[k + size_t(q) * n] = se * vkp + c * vkq;
        }
      }
  }
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(), [&](int i, int j) { return a[i + size_t(i) * n].real() < a[j + size_t(j) * n].real(); });
  w.resize(n);
  std::vector<cplx> vs(size_t(n) * n);
  for (int j = 0; j < n; ++j) {
    w[j] = a[idx[j] + size_t(idx[j]) * n].real();
    for (int i = 0; i < n; ++i) vs[i + size_t(j) * n] = v[i + size_t(idx[j]) * n];
  }
  v.swap(vs);
}
// Hermitian-definite generalized problem on a possibly nearly dependent basis (see genSymEigRobust)
inline int genHermEigRobust(const std::vector<cplx>& A, const std::vector<cplx>& B, int n, double eps, std::vector<double>& w,
                            std::vector<cplx>& Z) {
  std::vector<double> d(n);
  std::vector<cplx> G(size_t(n) * n);
  for (int j = 0; j < n; ++j) d[j] = B[j + size_t(j) * n].real() > 0 ? 1.0 / std::sqrt(B[j + size_t(j) * n].real()) : 0.0;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) G[i + size_t(j) * n] = B[i + size_t(j) * n] * (d[i] * d[j]);
  std::vector<double> lam;
  std::vector<cplx> V;
  hermEig(G, n, lam, V);
  std::vector<int> keep;
  for (int j = 0; j < n; ++j)
    if (lam[j] > eps * lam[n - 1]) keep.push_back(j);
  const int k = int(keep.size());
  if (k == 0) return 0;
  std::vector<cplx> Q(size_t(n) * k);   // B-orthonormal basis of the kept subspace
  for (int j = 0; j < k; ++j) {
    const double s = 1.0 / std::sqrt(lam[keep[j]]);
    for (int i = 0; i < n; ++i) Q[i + size_t(j) * n] = (d[i] * s) * V[i + size_t(keep[j]) * n];
  }
  std::vector<cplx> AQ(size_t(n) * k, cplx(0.0)), H(size_t(k) * k, cplx(0.0));
  for (int j = 0; j < k; ++j)
    for (int c = 0; c < n; ++c) {
      const cplx q = Q[c + size_t(j) * n];
      if (q == cplx(0.0)) continue;
      for (int i = 0; i < n; ++i) AQ[i + size_t(j) * n] += A[i + size_t(c) * n] * q;
    }
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < k; ++i) {
      cplx s = 0.0;
      for (int r = 0; r < n; ++r) s += std::conj(Q[r + size_t(i) * n]) * AQ[r + size_t(j) * n];
      H[i + size_t(j) * k] = s;
    }
  for (int j = 0; j < k; ++j) {
    H[j + size_t(j) * k] = H[j + size_t(j) * k].real();
    for (int i = 0; i < j; ++i) {
      const cplx s = 0.5 * (H[i + size_t(j) * k] + std::conj(H[j + size_t(i) * k]));
      H[i + size_t(j) * k] = s;
      H[j + size_t(i) * k] = std::conj(s);
    }
  }
  std::vector<cplx> Y;
  hermEig(H, k, w, Y);
  Z.assign(size_t(n) * k, cplx(0.0));
  for (int j = 0; j < k; ++j)
    for (int c = 0; c < k; ++c) {
      const cplx y = Y[c + size_t(j) * k];
      for (int i = 0; i < n; ++i) Z[i + size_t(j) * n] += Q[i + size_t(c) * n] * y;
    }
  return k;
}

// scalar-generic front ends used by the driver: the real instantiation calls exactly the real routines above
template <class S> struct Ops;
template <> struct Ops<double> {
  static double conj(double x) { return x; }
  static void eig(const std::vector<double>& a, int n, std::vector<double>& w, std::vector<double>& v) { symEig(a, n, w, v); }
  static int genEigRobust(const std::vector<double>& A, const std::vector<double>& B, int n, double eps, std::vector<double>& w,
                          std::vector<double>& Z) { return genSymEigRobust(A, B, n, eps, w, Z); }
};
template <> struct Ops<cplx> {
  static cplx conj(cplx x) { return std::conj(x); }
  static void eig(const std::vector<cplx>& a, int n, std::vector<double>& w, std::vector<cplx>& v) { hermEig(a, n, w, v); }
  static int genEigRobust(const std::vector<cplx>& A, const std::vector<cplx>& B, int n, double eps, std::vector<double>& w,
                          std::vector<cplx>& Z) { return genHermEigRobust(A, B, n, eps, w, Z); }
};
}  // namespace dense
}  // namespace mx

// ---- MxSolver (src/MxSolver.{h,cpp}): the loop the benchmark times ------------------------------------
struct MxSolverParams {
  int nev = 10;            // "eigensolver : nev" (MxSolver.cpp:37)
  int blockSize = 0;       // 0 -> nev + max(4, nev/2)
  int maxIters = 300;
  double tol = 1e-8;       // relative residual |A x - theta M x| / (|theta| |M x|)
  int verbose = 0;
  uint64_t seed = 12345;
  bool randomInit = true;  // MxSolver.cpp:62-64 starts from MvRandom
  bool profile = false;    // per-phase wall times (synchronises around each phase)
};

struct MxSolverResult {
  std::vector<double> eigenvalues, residuals;   // blockSize entries, ascending
  int iterations = 0, converged = 0;
  long applyA = 0, applyPrec = 0;
  double tApplyA = 0, tPrec = 0, tGram = 0, tUpdate = 0;   // filled when params.profile is set (adds syncs)
  double seconds = 0.0;
};

// Symmetric / Hermitian generalized problem A x = theta M x (M diagonal / SPD, may be null = identity),
// optional preconditioner T ~ A^-1. Scalar = double (default) or std::complex<double> (Bloch-periodic operators). Written against the MultiVec / Operator surface only, so the multivector type is
// a template parameter: MxSolver = MxSolverT<MxAnasaziMV<double>> is the GPU instantiation; tests/cpp/solver_host_check.cpp
// runs the same driver on a plain host multivector (no GPU) to cover its logic in the CPU suite.
template <class MV, class Scalar = double>
class MxSolverT {
  typedef Scalar S;
  typedef mx::SerialDenseMatrix<int, S> Dense;
  typedef mx::dense::Ops<S> Ops;

 public:
  MxSolverT(const mx::Operator<S>* A, const mx::Operator<S>* M, const mx::Operator<S>* prec, MxSolverParams p)
      : A_(A), M_(M), T_(prec), p_(p) {
    if (p_.blockSize <= 0) p_.blockSize = p_.nev + std::max(4, p_.nev / 2);
    if (p_.blockSize < p_.nev) p_.blockSize = p_.nev;
    if (3 * p_.blockSize > MXG_MAX_COLS) throw std::runtime_error("MxSolver: block size too large (3*block must be <= 128)");
  }

  // X: n x blockSize. On return holds the Ritz vectors (M-orthonormal), ascending eigenvalues.
  MxSolverResult solve(MV& X) {
    using clock = std::chrono::steady_clock;
    const int m = p_.blockSize;
    if (X.GetNumberVecs() != m) throw std::runtime_error("MxSolver::solve: X must have blockSize columns");
    MxSolverResult res;
    auto map = X.getMap();
    // S = [X | W | P] and its images under A and M live in three 3m-column allocations; views pick blocks
    MV Sb(map, 3 * m), ASb(map, 3 * m), MSb(map, 3 * m), tmp(map, 2 * m);
    // second set: the Rayleigh-Ritz update writes X_new / P_new straight into it and the sets swap roles (no copy-back)
    MV Sb2(map, 3 * m), ASb2(map, 3 * m), MSb2(map, 3 * m);
    auto range = [](int b, int n) { std::vector<int> v(n); std::iota(v.begin(), v.end(), b); return v; };
    auto view = [&](MV& base, const std::vector<int>& cols) { return std::unique_ptr<MV>(static_cast<MV*>(base.CloneViewNonConst(cols))); };
    auto timeit = [&](double& acc, auto&& f) {
      if (!p_.profile) { f(); return; }
      map->getComm()->sync();
      const auto a = clock::now();
      f();
      map->getComm()->sync();
      acc += std::chrono::duration<double>(clock::now() - a).count();
    };
    auto applyM = [&](const MV& in, MV& out) { if (M_) M_->Apply(in, out); else out = in; };
    auto rightMul = [&](MV& base, const std::vector<int>& srcCols, const Dense& C, const std::vector<int>& dstCols) {
      // base(:, dstCols) = base(:, srcCols) * C  through the temp block (src and dst may overlap)
      timeit(res.tUpdate, [&] {
        auto src = view(base, srcCols);
        auto t = view(tmp, range(0, C.numCols()));
        t->MvTimesMatAddMv(S(1.0), *src, C, S(0.0));
        auto dst = view(base, dstCols);
        *dst = *t;
      });
    };
    auto gram = [&](MV& left, const std::vector<int>& lc, MV& right, const std::vector<int>& rc) {
      auto l = view(left, lc);
      auto r = view(right, rc);
      Dense G(int(lc.size()), int(rc.size()));
      timeit(res.tGram, [&] { r->MvTransMv(S(1.0), *l, G); });
      return G;
    };
    auto toVec = [](const Dense& G) { return std::vector<S>(G.values(), G.values() + size_t(G.numRows()) * G.numCols()); };

    const auto t0 = clock::now();
    const std::vector<int> xc = range(0, m);
    {
      auto x = view(Sb, xc);
      if (p_.randomInit) { X.setSeed(p_.seed); X.MvRandom(); }
      *x = X;
      auto mx_ = view(MSb, xc);
      applyM(*x, *mx_);
    }
    // M-orthonormalise a block in place (and carry its A/M images along): B <- B L^-T
    // SVQB (Stathopoulos & Wu): G = D^-1/2 (B^T M B) D^-1/2 = V L V^T, B <- B D^-1/2 V L^-1/2, dropping
    // directions with L_i below a relative threshold. Returns the surviving column count; `cols` is
    // truncated to it (the block is compacted into its leading columns).
    auto orthonormalize = [&](std::vector<int>& cols, bool haveA) -> int {
      const int k = int(cols.size());
      if (k == 0) return 0;
      Dense G = gram(Sb, cols, MSb, cols);
      std::vector<S> g = toVec(G);
      std::vector<double> d(k);
      double dmax = 0.0;
      for (int j = 0; j < k; ++j) dmax = std::max(dmax, std::real(g[j + size_t(j) * k]));
      if (!(dmax > 0.0) || !std::isfinite(dmax)) { cols.clear(); return 0; }
      for (int j = 0; j < k; ++j) {
        const double gj = std::real(g[j + size_t(j) * k]);
        d[j] = gj > 1e-28 * dmax ? 1.0 / std::sqrt(gj) : 0.0;   // columns that vanished are dropped
      }
      for (int j = 0; j < k; ++j)
        for (int i = 0; i <= j; ++i) {
          const S s = 0.5 * (g[i + size_t(j) * k] + Ops::conj(g[j + size_t(i) * k])) * d[i] * d[j];
          g[i + size_t(j) * k] = s;
          g[j + size_t(i) * k] = Ops::conj(s);
        }
      std::vector<double> lam;
      std::vector<S> V;
      Ops::eig(g, k, lam, V);
      std::vector<int> keep;
      for (int j = 0; j < k; ++j)
        if (lam[j] > 1e-10 * std::max(lam[k - 1], 1.0)) keep.push_back(j);
      const int kk = int(keep.size());
      if (kk == 0) { cols.clear(); return 0; }
      Dense C(k, kk);
      for (int j = 0; j < kk; ++j) {
        const double s = 1.0 / std::sqrt(lam[keep[j]]);
        for (int i = 0; i < k; ++i) C(i, j) = d[i] * V[i + size_t(keep[j]) * k] * s;
      }
      std::vector<int> dst(cols.begin(), cols.begin() + kk);
      rightMul(Sb, cols, C, dst);
      rightMul(MSb, cols, C, dst);
      if (haveA) rightMul(ASb, cols, C, dst);
      cols = dst;
      return kk;
    };
    {
      std::vector<int> x0 = xc;
      if (orthonormalize(x0, false) != m) throw std::runtime_error("MxSolver: initial block is rank deficient");
    }
    {
      auto x = view(Sb, xc);
      auto ax = view(ASb, xc);
      timeit(res.tApplyA, [&] { A_->Apply(*x, *ax); });
      res.applyA += m;
    }
    std::vector<double> theta(m, 0.0);
    {  // initial Rayleigh-Ritz on X
      Dense H = gram(Sb, xc, ASb, xc);
      std::vector<double> w;
      std::vector<S> V;
      Ops::eig(toVec(H), m, w, V);
      Dense C(m, m);
      std::copy(V.begin(), V.end(), C.values());
      rightMul(Sb, xc, C, xc);
      rightMul(ASb, xc, C, xc);
      rightMul(MSb, xc, C, xc);
      theta = w;
    }
    int np = 0;  // columns currently in the P block
    std::vector<double> relres(m, 1.0);
    int it = 0;
    for (; it < p_.maxIters; ++it) {
      // residuals R = A X - M X Theta (formed in the temp block)
      auto R = view(tmp, range(0, m));
      {
        auto ax = view(ASb, xc);
        auto mxv = view(MSb, xc);
        auto scaled = view(tmp, range(m, m));
        *scaled = *mxv;
        scaled->MvScale(std::vector<S>(theta.begin(), theta.end()));
        R->MvAddMv(S(1.0), *ax, S(-1.0), *scaled);
        std::vector<double> rn, mn;
        R->MvNorm(rn);
        mxv->MvNorm(mn);
        // |A x - theta M x| / (|theta| |M x|); for (near-)null modes theta is replaced by a floor
        // relative to the largest Ritz value of the block so that exact zeros can converge
        double thetaRef = 0.0;
        for (int j = 0; j < p_.nev; ++j) thetaRef = std::max(thetaRef, std::fabs(theta[j]));
        for (int j = 0; j < m; ++j) {
          const double den = std::max(std::fabs(theta[j]), 0.1 * thetaRef) * mn[j];
          relres[j] = den > 0 ? rn[j] / den : rn[j];
        }
      }
      std::vector<int> active;
      int nconv = 0;
      for (int j = 0; j < m; ++j) {
        if (relres[j] >= p_.tol) active.push_back(j);
        else if (j < p_.nev) ++nconv;
      }
      bool done = true;
      for (int j = 0; j < p_.nev; ++j) done = done && relres[j] < p_.tol;
      if (p_.verbose) {
        double worst = 0;
        for (int j = 0; j < p_.nev; ++j) worst = std::max(worst, relres[j]);
        std::printf("MxSolver iter %3d  conv %2d/%d  theta[0]=%.10g theta[nev-1]=%.10g  max res %.3e  active %d\n", it, nconv, p_.nev,
                    theta[0], theta[p_.nev - 1], worst, int(active.size()));
      }
      res.converged = nconv;
      if (done) break;
      const int na = int(active.size());
      std::vector<int> wc = range(m, na), pc = range(2 * m, np);
      {  // W = T R(:, active)
        auto Ra = view(tmp, active);
        auto W = view(Sb, wc);
        if (T_) { timeit(res.tPrec, [&] { T_->Apply(*Ra, *W); }); res.applyPrec += na; }
        else *W = *Ra;
        auto MW = view(MSb, wc);
        applyM(*W, *MW);
        // W <- W - X (X^T M W)
        Dense C = gram(Sb, xc, MSb, wc);
        auto Xv = view(Sb, xc);
        auto MXv = view(MSb, xc);
        W->MvTimesMatAddMv(S(-1.0), *Xv, C, S(1.0));
        MW->MvTimesMatAddMv(S(-1.0), *MXv, C, S(1.0));
      }
      orthonormalize(wc, false);                 // may drop directions that vanished after the projection
      if (!wc.empty()) {
        auto W = view(Sb, wc);
        auto AW = view(ASb, wc);
        timeit(res.tApplyA, [&] { A_->Apply(*W, *AW); });
        res.applyA += long(wc.size());
      }
      np = orthonormalize(pc, true);             // a degenerate search block shrinks or disappears
      if (wc.empty() && np == 0) {
        if (p_.verbose) std::printf("MxSolver: search space exhausted at iteration %d; stopping\n", it);
        break;
      }
      // Rayleigh-Ritz on span[X W P]
      std::vector<double> w;
      std::vector<S> Z;
      int ns = 0;
      for (int attempt = 0; attempt < 2; ++attempt) {
        std::vector<int> sc = xc;
        sc.insert(sc.end(), wc.begin(), wc.end());
        if (np > 0) { const std::vector<int> pcc = range(2 * m, np); sc.insert(sc.end(), pcc.begin(), pcc.end()); }
        ns = int(sc.size());
        std::vector<S> ga, gm;
        const int nwb = int(wc.size());
        if (it % 10 == 0) {
          // explicit Gram matrices of the whole basis (also resets the round-off drift of the implicit ones)
          Dense GA = gram(Sb, sc, ASb, sc), GM = gram(Sb, sc, MSb, sc);
          ga = toVec(GA);
          gm = toVec(GM);
        } else {
          // implicit blocks: X^T A X = Theta, X^T M X = W^T M W = P^T M P = I, X^T M W = 0 by construction;
          // only the columns belonging to W and P are computed (8 m^2 instead of 18 m^2 inner products)
          std::vector<int> wpc2 = wc, xw = xc;
          xw.insert(xw.end(), wc.begin(), wc.end());
          if (np > 0) { const std::vector<int> pcc = range(2 * m, np); wpc2.insert(wpc2.end(), pcc.begin(), pcc.end()); }
          const int nwp = int(wpc2.size());
          Dense GA1 = gram(Sb, sc, ASb, wpc2);                    // ns x (nw + np)
          ga.assign(size_t(ns) * ns, S(0.0));
          gm.assign(size_t(ns) * ns, S(0.0));
          for (int j = 0; j < m; ++j) ga[j + size_t(j) * ns] = theta[j];
          for (int j = 0; j < ns; ++j) gm[j + size_t(j) * ns] = 1.0;
          for (int j = 0; j < nwp; ++j)
            for (int i = 0; i < ns; ++i) {
              ga[i + size_t(m + j) * ns] = GA1(i, j);
              if (i < m) ga[(m + j) + size_t(i) * ns] = Ops::conj(GA1(i, j));
            }
          if (np > 0) {
            const std::vector<int> pcc = range(2 * m, np);
            Dense GM1 = gram(Sb, xw, MSb, pcc);                   // (m + nw) x np
            for (int j = 0; j < np; ++j)
              for (int i = 0; i < m + nwb; ++i) {
                gm[i + size_t(m + nwb + j) * ns] = GM1(i, j);
                gm[(m + nwb + j) + size_t(i) * ns] = Ops::conj(GM1(i, j));
              }
          }
        }
        for (int j = 0; j < ns; ++j)
          for (int i = 0; i < j; ++i) {
            const S a = 0.5 * (ga[i + size_t(j) * ns] + Ops::conj(ga[j + size_t(i) * ns]));
            const S b = 0.5 * (gm[i + size_t(j) * ns] + Ops::conj(gm[j + size_t(i) * ns]));
            ga[i + size_t(j) * ns] = a;
            ga[j + size_t(i) * ns] = Ops::conj(a);
            gm[i + size_t(j) * ns] = b;
            gm[j + size_t(i) * ns] = Ops::conj(b);
          }
        // truncated generalized eigenproblem: nearly dependent directions of [X W P] are discarded
        if (Ops::genEigRobust(ga, gm, ns, 1e-12, w, Z) >= m) break;
        if (np == 0) throw std::runtime_error("MxSolver: projected basis has rank below the block size");
        np = 0;  // retry without P
      }
      // X_new = S Z(:, 0:m);  P_new = [W P] Z(m:, 0:m)
      const int nw = ns - m;   // rows of Z belonging to W and P
      Dense Cx(ns, m), Cp(nw, m);
      for (int j = 0; j < m; ++j) {
        for (int i = 0; i < ns; ++i) Cx(i, j) = Z[i + size_t(j) * ns];
        for (int i = 0; i < nw; ++i) Cp(i, j) = Z[(m + i) + size_t(j) * ns];
      }
      std::vector<int> sc = xc, wpc = wc;
      sc.insert(sc.end(), wc.begin(), wc.end());
      if (np > 0) { const std::vector<int> pcc = range(2 * m, np); sc.insert(sc.end(), pcc.begin(), pcc.end()); wpc.insert(wpc.end(), pcc.begin(), pcc.end()); }
      const std::vector<int> pnew = range(2 * m, m);
      MV* cur[3] = {&Sb, &ASb, &MSb};
      MV* alt[3] = {&Sb2, &ASb2, &MSb2};
      for (int t = 0; t < 3; ++t) timeit(res.tUpdate, [&] {
        auto s = view(*cur[t], sc);
        auto wp = view(*cur[t], wpc);
        auto tx = view(*alt[t], xc);
        auto tp = view(*alt[t], pnew);
        tx->MvTimesMatAddMv(S(1.0), *s, Cx, S(0.0));
        tp->MvTimesMatAddMv(S(1.0), *wp, Cp, S(0.0));
        cur[t]->swap(*alt[t]);
      });
      np = m;
      for (int j = 0; j < m; ++j) theta[j] = w[j];
    }
    {
      auto x = view(Sb, xc);
      X = *x;
    }
    X.getMap()->getComm()->sync();
    res.iterations = it;
    res.eigenvalues = theta;
    res.residuals = relres;
    res.seconds = std::chrono::duration<double>(clock::now() - t0).count();
    return res;
  }

 private:
  const mx::Operator<S>* A_;
  const mx::Operator<S>* M_;
  const mx::Operator<S>* T_;
  MxSolverParams p_;
};
typedef MxSolverT<MxAnasaziMV<double>> MxSolver;

// ---- MxMagWaveOp (src/MxMagWaveOp.{h,cpp}): the shift-invert operator the reference hands to Anasazi ------
//   Apply:  y = P (L - sigma M)^-1 M x            (MxMagWaveOp.cpp:825-943)
//   with L = vecLapl, M = mRhs (diagonal), and the divergence-cleaning projection
//   P b = b + gradPsi * scaLapl^-1 * divB * M * b  (:893-924), scaLapl = -(divB M gradPsi) (:208-223).
// The reference solves both systems with AztecOO GMRES/CG + ML or ILUT (:285-537). Here both are block
// preconditioned CG on the GPU (valid for sigma below the lowest eigenvalue -- the reference's default
// automatic shift 0.05 (2 pi / L)^2, src/mx.py:711-712 -- where L - sigma M is positive definite), with the
// multigrid V-cycle as the vector preconditioner and Jacobi (or a second V-cycle) for the scalar solve.
struct MxMagWaveOpParams {
  double shift = 0.0;          // sigma
  double linTol = 1e-10;       // "linear solver : tol" on |r| / |b|
  int maxLinIters = 1000;
  bool hasCurlNull = true;     // 3-D: project out the gradient fields
};

class MxMagWaveOp : public mx::Operator<double> {
  typedef MxAnasaziMV<double> MV;

 public:
  MxMagWaveOp(mxg_crs* vecLapl, mxg_mv* mDiag, mxg_crs* divB, mxg_crs* gradPsi, mxg_crs* scaLapl,
              const mx::Operator<double>* vecPrec, const mx::Operator<double>* scaPrec, MxMagWaveOpParams p)
      : L_(vecLapl), m_(mDiag), D_(divB), G_(gradPsi), S_(scaLapl), Tv_(vecPrec), Ts_(scaPrec), p_(p) {}

  // counters the reference prints from its destructor (MxMagWaveOp.cpp:96-115)
  mutable long numApplies = 0, numVecLinIters = 0, numScaLinIters = 0;

  void Apply(const mx::MultiVec<double>& x, mx::MultiVec<double>& y) const override {
    const MV& x2 = dynamic_cast<const MV&>(x);
    MV& y2 = dynamic_cast<MV&>(y);
    ++numApplies;
    const int nb = x2.GetNumberVecs();
    std::shared_ptr<MxMap> bmap = x2.getMap();
    MV rhs(bmap, nb), bWork(bmap, nb);
    mx::check(mxg_mv_diag_mult(rhs.getRawMV(), m_, x2.getRawMV()));                       // y = mRhs x (:865)
    numVecLinIters += pcg([&](const MV& in, MV& out) { applyShifted(in, out); }, Tv_, nullptr, rhs, bWork);   // (:869-885)
    if (!p_.hasCurlNull) { y2 = bWork; return; }
    mx::check(mxg_mv_diag_mult(rhs.getRawMV(), m_, bWork.getRawMV()));                    // y = mRhs bWork (:895)
    std::shared_ptr<MxMap> pmap(new MxMap(mxg_crs_row_map(D_), bmap->getComm(), false));
    MV psi1(pmap, nb), psi2(pmap, nb);
    mx::check(mxg_crs_apply(D_, rhs.getRawMV(), psi1.getRawMV()));                         // psi1 = divB y (:896)
    numScaLinIters += pcg([&](const MV& in, MV& out) { mx::check(mxg_crs_apply(S_, in.getRawMV(), out.getRawMV())); },
                          Ts_, S_, psi1, psi2);                                             // (:903-913)
    mx::check(mxg_crs_apply(G_, psi2.getRawMV(), y2.getRawMV()));                          // y = gradPsi psi2 (:919)
    y2.MvAddMv(1.0, y2, 1.0, bWork);                                                        // y += bWork (:921)
  }

  // E = [invEps] curlB B (MxMagWaveOp.cpp:1237-1250). invEps may be NULL (no dielectric).
  static void magToElec(mxg_crs* curlB, mxg_crs* invEps, const MV& mag, MV& elec) {
    if (!invEps) { mx::check(mxg_crs_apply(curlB, mag.getRawMV(), elec.getRawMV())); return; }
    MV d(elec.getMap(), elec.GetNumberVecs());
    mx::check(mxg_crs_apply(curlB, mag.getRawMV(), d.getRawMV()));
    mx::check(mxg_crs_apply(invEps, d.getRawMV(), elec.getRawMV()));
  }
  // eigenvalue of the (possibly shift-inverted) operator -> frequency in Hz (MxMagWaveOp.cpp:1252-1271)
  static void eigValsToFreqs(const std::vector<std::complex<double>>& eigVals, std::vector<std::complex<double>>& freqs,
                             double shift, bool invert) {
    const double lightspeed = 299792458., pi = 3.14159265358979323846;   // MxUtil.hpp:23-24
    freqs.resize(eigVals.size());
    for (size_t i = 0; i < eigVals.size(); ++i) {
      std::complex<double> k2 = invert ? 1.0 / eigVals[i] + shift : eigVals[i] + shift;
      freqs[i] = std::sqrt(k2) * lightspeed / 2.0 / pi;
    }
  }

 private:
  void applyShifted(const MV& in, MV& out) const {   // out = (L - sigma M) in
    mx::check(mxg_crs_apply(L_, in.getRawMV(), out.getRawMV()));
    if (p_.shift != 0.0) {
      MV t(in.getMap(), in.GetNumberVecs());
      mx::check(mxg_mv_diag_mult(t.getRawMV(), m_, in.getRawMV()));
      out.MvAddMv(1.0, out, -p_.shift, t);
    }
  }
  // block preconditioned CG, one independent recurrence per column; returns the iteration count.
  // prec == nullptr and jac != nullptr: Jacobi with the operator's diagonal; both null: unpreconditioned.
  template <class ApplyA>
  long pcg(ApplyA&& A, const mx::Operator<double>* prec, mxg_crs* jac, const MV& b, MV& x) const {
    const int nb = b.GetNumberVecs();
    std::shared_ptr<MxMap> map = b.getMap();
    MV r(b), z(map, nb), pdir(map, nb), q(map, nb), tmp(map, nb);
    x.MvInit(0.0);
    std::vector<double> bn, rn, rz(nb), rzNew(nb), pq(nb), alpha(nb), beta(nb);
    b.MvNorm(bn);
    auto precond = [&](const MV& in, MV& out) {
      if (prec) prec->Apply(in, out);
      else if (jac) mx::check(mxg_crs_jacobi(jac, in.getRawMV(), out.getRawMV()));
      else out = in;
    };
    precond(r, z);
    pdir = z;
    r.MvDot(z, rz);
    long it = 0;
    for (; it < p_.maxLinIters; ++it) {
      r.MvNorm(rn);
      bool done = true;
      for (int j = 0; j < nb; ++j) done = done && (bn[j] == 0.0 || rn[j] <= p_.linTol * bn[j]);
      if (done) break;
      A(pdir, q);
      pdir.MvDot(q, pq);
      for (int j = 0; j < nb; ++j) alpha[j] = pq[j] != 0.0 ? rz[j] / pq[j] : 0.0;
      tmp = pdir; tmp.MvScale(alpha); x.MvAddMv(1.0, x, 1.0, tmp);
      tmp = q; tmp.MvScale(alpha); r.MvAddMv(1.0, r, -1.0, tmp);
      precond(r, z);
      r.MvDot(z, rzNew);
      for (int j = 0; j < nb; ++j) { beta[j] = rz[j] != 0.0 ? rzNew[j] / rz[j] : 0.0; rz[j] = rzNew[j]; }
      pdir.MvScale(beta);
      pdir.MvAddMv(1.0, pdir, 1.0, z);
    }
    return it;
  }

  mxg_crs* L_;
  mxg_mv* m_;
  mxg_crs *D_, *G_, *S_;
  const mx::Operator<double>* Tv_;
  const mx::Operator<double>* Ts_;
  MxMagWaveOpParams p_;
};
