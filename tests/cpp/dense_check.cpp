// Host-side dense kernels of the eigensolver driver (include/mx/MxSolver.hpp, namespace mx::dense): the pieces the
// reference gets from Teuchos::LAPACK inside Anasazi. No GPU needed.
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

#include "mx/MxSolver.hpp"

static int failures = 0;
#define CHECK(cond, ...)                 \
  do {                                   \
    if (!(cond)) {                       \
      std::printf("FAILED %s:%d: ", __FILE__, __LINE__); \
      std::printf(__VA_ARGS__);          \
      std::printf("\n");                 \
      ++failures;                        \
    }                                    \
  } while (0)

typedef std::vector<double> Mat;
static double at(const Mat& m, int n, int i, int j) { return m[i + size_t(j) * n]; }

static Mat randomSym(int n, std::mt19937_64& rng, bool spd) {
  std::uniform_real_distribution<double> u(-1.0, 1.0);
  Mat g(size_t(n) * n), a(size_t(n) * n, 0.0);
  for (auto& v : g) v = u(rng);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      double s = 0;
      if (spd) for (int k = 0; k < n; ++k) s += at(g, n, k, i) * at(g, n, k, j);
      else s = 0.5 * (at(g, n, i, j) + at(g, n, j, i));
      a[i + size_t(j) * n] = s + (spd && i == j ? 0.1 : 0.0);
    }
  return a;
}

int main() {
  std::mt19937_64 rng(7);
  using namespace mx::dense;
  for (int n : {1, 2, 5, 17, 48}) {
    // Cholesky + triangular inverse
    Mat B = randomSym(n, rng, true), L = B;
    CHECK(cholesky(L, n), "cholesky rejected an SPD matrix, n=%d", n);
    double err = 0;
    for (int j = 0; j < n; ++j)
      for (int i = j; i < n; ++i) {
        double s = 0;
        for (int k = 0; k <= j; ++k) s += at(L, n, i, k) * at(L, n, j, k);
        err = std::fmax(err, std::fabs(s - at(B, n, i, j)));
      }
    CHECK(err < 1e-12 * n, "L L^T != B (%.3e), n=%d", err, n);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < j; ++i) L[i + size_t(j) * n] = 0.0;
    Mat Li = invLower(L, n), I = mul(Li, false, L, n, n);
    err = 0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) err = std::fmax(err, std::fabs(at(I, n, i, j) - (i == j ? 1.0 : 0.0)));
    CHECK(err < 1e-10, "Li L != I (%.3e), n=%d", err, n);
    // Jacobi eigen-decomposition
    Mat A = randomSym(n, rng, false), w, V;
    symEig(A, n, w, V);
    Mat AV = mul(A, false, V, n, n), VtV = mul(V, true, V, n, n);
    err = 0;
    double orth = 0;
    for (int j = 0; j < n; ++j) {
      if (j) CHECK(w[j] >= w[j - 1], "eigenvalues not ascending");
      for (int i = 0; i < n; ++i) {
        err = std::fmax(err, std::fabs(at(AV, n, i, j) - w[j] * at(V, n, i, j)));
        orth = std::fmax(orth, std::fabs(at(VtV, n, i, j) - (i == j ? 1.0 : 0.0)));
      }
    }
    CHECK(err < 1e-12 * n && orth < 1e-12 * n, "symEig residual %.3e orth %.3e n=%d", err, orth, n);
    // generalized problem A z = w B z, Z^T B Z = I
    Mat Z;
    CHECK(genSymEig(A, B, n, w, Z), "genSymEig failed, n=%d", n);
    Mat AZ = mul(A, false, Z, n, n), BZ = mul(B, false, Z, n, n), ZtBZ = mul(Z, true, BZ, n, n);
    err = orth = 0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        err = std::fmax(err, std::fabs(at(AZ, n, i, j) - w[j] * at(BZ, n, i, j)));
        orth = std::fmax(orth, std::fabs(at(ZtBZ, n, i, j) - (i == j ? 1.0 : 0.0)));
      }
    CHECK(err < 1e-9 && orth < 1e-10, "genSymEig residual %.3e orth %.3e n=%d", err, orth, n);
    // robust variant agrees on a well-conditioned pencil
    Mat w2, Z2;
    const int kept = genSymEigRobust(A, B, n, 1e-14, w2, Z2);
    CHECK(kept == n, "robust variant dropped %d directions of a full-rank basis", n - kept);
    for (int j = 0; j < kept && j < n; ++j) CHECK(std::fabs(w2[j] - w[j]) < 1e-8 * (1 + std::fabs(w[j])), "robust eigenvalue %d", j);
  }
  // Not SPD -> Cholesky refuses
  {
    Mat B = {1.0, 2.0, 2.0, 1.0};
    CHECK(!cholesky(B, 2), "cholesky accepted an indefinite matrix");
  }
  // Dependent basis: S = [s1 s2 s1+s2] (Gram of rank 2). The robust solver keeps 2 directions and returns the
  // Ritz values of the 2-dimensional pencil.
  {
    const int n = 3;
    // basis vectors e1, e2, e1+e2 in R^2 with A = diag(1, 3): Gram B = S^T S, A_s = S^T A S
    const double S[2][3] = {{1, 0, 1}, {0, 1, 1}}, D[2] = {1.0, 3.0};
    Mat A(9), B(9);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        A[i + 3 * j] = S[0][i] * D[0] * S[0][j] + S[1][i] * D[1] * S[1][j];
        B[i + 3 * j] = S[0][i] * S[0][j] + S[1][i] * S[1][j];
      }
    Mat w, Z;
    const int kept = genSymEigRobust(A, B, n, 1e-12, w, Z);
    CHECK(kept == 2, "kept %d directions of a rank-2 basis", kept);
    if (kept == 2) CHECK(std::fabs(w[0] - 1.0) < 1e-12 && std::fabs(w[1] - 3.0) < 1e-12, "Ritz values %.15g %.15g", w[0], w[1]);
    Mat L = B;
    CHECK(!cholesky(L, n) || true, "unreachable");   // Cholesky may or may not notice; the robust path must not depend on it
  }
  // Hermitian counterparts (complex Jacobi, rank-revealing generalized problem)
  {
    typedef std::complex<double> Z;
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    for (int n : {1, 2, 6, 23, 40}) {
      std::vector<Z> g(size_t(n) * n), A(size_t(n) * n), B(size_t(n) * n, Z(0.0));
      for (auto& v : g) v = Z(u(rng), u(rng));
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
          A[i + size_t(j) * n] = 0.5 * (g[i + size_t(j) * n] + std::conj(g[j + size_t(i) * n]));
          Z s = 0;
          for (int k = 0; k < n; ++k) s += std::conj(g[k + size_t(i) * n]) * g[k + size_t(j) * n];
          B[i + size_t(j) * n] = s + (i == j ? 0.1 : 0.0);
        }
      std::vector<double> w;
      std::vector<Z> V;
      hermEig(A, n, w, V);
      double err = 0, orth = 0;
      for (int j = 0; j < n; ++j) {
        if (j) CHECK(w[j] >= w[j - 1], "Hermitian eigenvalues not ascending");
        for (int i = 0; i < n; ++i) {
          Z av = 0, vv = 0;
          for (int k = 0; k < n; ++k) { av += A[i + size_t(k) * n] * V[k + size_t(j) * n]; vv += std::conj(V[k + size_t(i) * n]) * V[k + size_t(j) * n]; }
          err = std::fmax(err, std::abs(av - w[j] * V[i + size_t(j) * n]));
          orth = std::fmax(orth, std::abs(vv - (i == j ? 1.0 : 0.0)));
        }
      }
      CHECK(err < 1e-12 * n && orth < 1e-12 * n, "hermEig residual %.3e orth %.3e n=%d", err, orth, n);
      std::vector<Z> Zv;
      const int kept = genHermEigRobust(A, B, n, 1e-14, w, Zv);
      CHECK(kept == n, "genHermEigRobust dropped %d directions of a full-rank basis", n - kept);
      err = orth = 0;
      for (int j = 0; j < kept; ++j)
        for (int i = 0; i < n; ++i) {
          Z az = 0, bz = 0;
          for (int k = 0; k < n; ++k) { az += A[i + size_t(k) * n] * Zv[k + size_t(j) * n]; bz += B[i + size_t(k) * n] * Zv[k + size_t(j) * n]; }
          err = std::fmax(err, std::abs(az - w[j] * bz));
        }
      for (int j = 0; j < kept; ++j)
        for (int i = 0; i < kept; ++i) {
          Z s = 0;
          for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) s += std::conj(Zv[r + size_t(i) * n]) * B[r + size_t(c) * n] * Zv[c + size_t(j) * n];
          orth = std::fmax(orth, std::abs(s - (i == j ? 1.0 : 0.0)));
        }
      CHECK(err < 1e-8 && orth < 1e-9, "genHermEigRobust residual %.3e orth %.3e n=%d", err, orth, n);
    }
    // dependent complex basis: s3 = i s1 + s2 in C^2, A = diag(1, 3)
    const Z S[2][3] = {{1, 0, Z(0, 1)}, {0, 1, 1}};
    const double D[2] = {1.0, 3.0};
    std::vector<Z> A(9), B(9);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        A[i + 3 * j] = std::conj(S[0][i]) * D[0] * S[0][j] + std::conj(S[1][i]) * D[1] * S[1][j];
        B[i + 3 * j] = std::conj(S[0][i]) * S[0][j] + std::conj(S[1][i]) * S[1][j];
      }
    std::vector<double> w;
    std::vector<Z> Zv;
    const int kept = genHermEigRobust(A, B, 3, 1e-12, w, Zv);
    CHECK(kept == 2, "kept %d directions of a rank-2 complex basis", kept);
    if (kept == 2) CHECK(std::fabs(w[0] - 1.0) < 1e-12 && std::fabs(w[1] - 3.0) < 1e-12, "complex Ritz values %.15g %.15g", w[0], w[1]);
  }
  if (failures == 0) std::printf("PASSED\n");
  return failures ? 1 : 0;
}
