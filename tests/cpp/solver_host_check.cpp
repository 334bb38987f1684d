// The LOBPCG driver of include/mx/MxSolver.hpp (MxSolverT) on a plain host multivector: covers the solver's host
// logic -- SVQB orthonormalisation, soft locking, implicit Gram blocks, ping-pong basis buffers, the rank-revealing
// Rayleigh-Ritz -- in the CPU suite. The GPU instantiation MxSolverT<MxAnasaziMV<double>> is the same code.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "host_mv.hpp"
#include "mx/MxSolver.hpp"

using hostmv::HostMV;

static int failures = 0;
#define CHECK(cond, ...)                                   \
  do {                                                     \
    if (!(cond)) {                                         \
      std::printf("FAILED %s:%d: ", __FILE__, __LINE__);   \
      std::printf(__VA_ARGS__);                            \
      std::printf("\n");                                   \
      ++failures;                                          \
    }                                                      \
  } while (0)

// 7-point Dirichlet Laplacian on an nx x ny x nz grid with spacings hx, hy, hz
struct Laplace : mx::Operator<double> {
  int nx, ny, nz;
  double cx, cy, cz;
  Laplace(int nx, int ny, int nz, double hx, double hy, double hz) : nx(nx), ny(ny), nz(nz), cx(1 / (hx * hx)), cy(1 / (hy * hy)), cz(1 / (hz * hz)) {}
  int64_t n() const { return int64_t(nx) * ny * nz; }
  void Apply(const mx::MultiVec<double>& x_, mx::MultiVec<double>& y_) const override {
    const HostMV& x = dynamic_cast<const HostMV&>(x_);
    HostMV& y = dynamic_cast<HostMV&>(y_);
    for (int v = 0; v < x.GetNumberVecs(); ++v) {
      const double* a = x.col(v);
      double* b = y.col(v);
      for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
          for (int k = 0; k < nz; ++k) {
            const int64_t p = (int64_t(i) * ny + j) * nz + k;
            double s = 2 * (cx + cy + cz) * a[p];
            if (i > 0) s -= cx * a[p - int64_t(ny) * nz];
            if (i < nx - 1) s -= cx * a[p + int64_t(ny) * nz];
            if (j > 0) s -= cy * a[p - nz];
            if (j < ny - 1) s -= cy * a[p + nz];
            if (k > 0) s -= cz * a[p - 1];
            if (k < nz - 1) s -= cz * a[p + 1];
            b[p] = s;
          }
    }
  }
  std::vector<double> spectrum() const {
    std::vector<double> w;
    auto lam = [](int m, int n, double c) { const double s = std::sin(0.5 * M_PI * m / (n + 1)); return 4 * c * s * s; };
    for (int i = 1; i <= nx; ++i)
      for (int j = 1; j <= ny; ++j)
        for (int k = 1; k <= nz; ++k) w.push_back(lam(i, nx, cx) + lam(j, ny, cy) + lam(k, nz, cz));
    std::sort(w.begin(), w.end());
    return w;
  }
};
struct Diag : mx::Operator<double> {
  std::vector<double> d;
  void Apply(const mx::MultiVec<double>& x_, mx::MultiVec<double>& y_) const override {
    const HostMV& x = dynamic_cast<const HostMV&>(x_);
    HostMV& y = dynamic_cast<HostMV&>(y_);
    for (int v = 0; v < x.GetNumberVecs(); ++v)
      for (size_t i = 0; i < d.size(); ++i) y.col(v)[i] = d[i] * x.col(v)[i];
  }
};

int main() {
  typedef MxSolverT<HostMV> Solver;
  // 1. standard problem with degenerate eigenvalues (cube: the second eigenvalue is triple), no preconditioner
  {
    Laplace A(9, 9, 9, 0.1, 0.1, 0.1);
    const std::vector<double> want = A.spectrum();
    MxSolverParams p;
    p.nev = 6; p.blockSize = 10; p.tol = 1e-9; p.maxIters = 400;
    auto map = std::make_shared<hostmv::Map>(A.n());
    HostMV X(map, size_t(p.blockSize));
    Solver solver(&A, nullptr, nullptr, p);
    MxSolverResult r = solver.solve(X);
    CHECK(r.converged == p.nev, "converged %d of %d in %d iterations", r.converged, p.nev, r.iterations);
    for (int j = 0; j < p.nev; ++j) CHECK(std::fabs(r.eigenvalues[j] - want[j]) < 1e-9 * want[j], "eigenvalue %d: %.12g vs %.12g", j, r.eigenvalues[j], want[j]);
    CHECK(std::fabs(want[1] - want[3]) < 1e-12 && std::fabs(r.eigenvalues[1] - r.eigenvalues[3]) < 1e-7 * want[1], "triple eigenvalue");
    // Ritz vectors are orthonormal and satisfy the eigen-equation
    mx::SerialDenseMatrix<int, double> G(p.blockSize, p.blockSize);
    X.MvTransMv(1.0, X, G);
    for (int j = 0; j < p.nev; ++j)
      for (int i = 0; i < p.nev; ++i) CHECK(std::fabs(G(i, j) - (i == j ? 1.0 : 0.0)) < 1e-8, "X^T X (%d,%d) = %.3e", i, j, G(i, j));
    HostMV AX(map, size_t(p.blockSize));
    A.Apply(X, AX);
    std::vector<double> th(r.eigenvalues.begin(), r.eigenvalues.end());
    HostMV XT(X);
    XT.MvScale(th);
    AX.MvAddMv(1.0, AX, -1.0, XT);
    std::vector<double> rn;
    AX.MvNorm(rn);
    for (int j = 0; j < p.nev; ++j) CHECK(rn[j] < 2e-9 * want[j], "residual %d = %.3e", j, rn[j]);
    std::printf("standard:    %d iterations, %ld operator columns\n", r.iterations, r.applyA);
  }
  // 2. generalized problem with a diagonal mass matrix and a Jacobi-like preconditioner (anisotropic box)
  {
    Laplace A(10, 8, 6, 0.1, 0.125, 0.15);
    Diag M, T;
    M.d.resize(size_t(A.n()));
    T.d.resize(size_t(A.n()));
    for (size_t i = 0; i < M.d.size(); ++i) { M.d[i] = 1.0 + 0.5 * std::sin(0.37 * double(i)); T.d[i] = 1.0 / (2 * (A.cx + A.cy + A.cz)); }
    MxSolverParams p;
    p.nev = 4; p.blockSize = 8; p.tol = 1e-9; p.maxIters = 500;
    auto map = std::make_shared<hostmv::Map>(A.n());
    HostMV X(map, size_t(p.blockSize));
    Solver solver(&A, &M, &T, p);
    MxSolverResult r = solver.solve(X);
    CHECK(r.converged == p.nev, "generalized: converged %d of %d in %d iterations", r.converged, p.nev, r.iterations);
    CHECK(r.applyPrec > 0, "preconditioner never applied");
    // check A x = theta M x and M-orthonormality
    HostMV AX(map, size_t(p.blockSize)), MX(map, size_t(p.blockSize));
    A.Apply(X, AX);
    M.Apply(X, MX);
    mx::SerialDenseMatrix<int, double> G(p.blockSize, p.blockSize);
    MX.MvTransMv(1.0, X, G);
    for (int j = 0; j < p.nev; ++j)
      for (int i = 0; i < p.nev; ++i) CHECK(std::fabs(G(i, j) - (i == j ? 1.0 : 0.0)) < 1e-8, "X^T M X (%d,%d) = %.3e", i, j, G(i, j));
    std::vector<double> th(r.eigenvalues.begin(), r.eigenvalues.end()), rn, mn;
    HostMV MXT(MX);
    MXT.MvScale(th);
    AX.MvAddMv(1.0, AX, -1.0, MXT);
    AX.MvNorm(rn);
    MX.MvNorm(mn);
    for (int j = 0; j < p.nev; ++j) CHECK(rn[j] < 2e-9 * th[j] * mn[j], "generalized residual %d = %.3e", j, rn[j]);
    for (int j = 1; j < p.nev; ++j) CHECK(th[j] >= th[j - 1], "eigenvalues ascending");
    // Rayleigh quotients bracket: the generalized eigenvalues lie between lambda/max(M) and lambda/min(M)
    const std::vector<double> lam = A.spectrum();
    CHECK(th[0] > lam[0] / 1.5 - 1e-9 && th[0] < lam[0] / 0.5 + 1e-9, "lowest generalized eigenvalue %.6g outside its bracket", th[0]);
    std::printf("generalized: %d iterations, %ld operator columns, %ld preconditioner columns\n", r.iterations, r.applyA, r.applyPrec);
  }
  // 3. user-supplied start block (randomInit off), rank-deficient start is rejected
  {
    Laplace A(5, 5, 5, 0.2, 0.2, 0.2);
    MxSolverParams p;
    p.nev = 2; p.blockSize = 4; p.randomInit = false; p.maxIters = 50;
    auto map = std::make_shared<hostmv::Map>(A.n());
    HostMV X(map, size_t(p.blockSize));
    X.MvInit(1.0);   // four identical columns
    bool threw = false;
    try { Solver(&A, nullptr, nullptr, p).solve(X); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw, "rank-deficient start block must be rejected");
    bool threw2 = false;
    try { HostMV Y(map, 3); Solver(&A, nullptr, nullptr, p).solve(Y); } catch (const std::runtime_error&) { threw2 = true; }
    CHECK(threw2, "wrong block width must be rejected");
  }
  // 4. the whole spectrum of a tiny operator: block as large as the space allows, search space gets exhausted
  {
    Laplace A(2, 2, 2, 1.0, 1.0, 1.0);   // n = 8
    const std::vector<double> want = A.spectrum();
    MxSolverParams p;
    p.nev = 3; p.blockSize = 4; p.tol = 1e-10; p.maxIters = 100;
    auto map = std::make_shared<hostmv::Map>(A.n());
    HostMV X(map, size_t(p.blockSize));
    MxSolverResult r = Solver(&A, nullptr, nullptr, p).solve(X);
    CHECK(r.converged == p.nev, "tiny: converged %d", r.converged);
    for (int j = 0; j < p.nev; ++j) CHECK(std::fabs(r.eigenvalues[j] - want[j]) < 1e-9, "tiny eigenvalue %d: %.12g vs %.12g", j, r.eigenvalues[j], want[j]);
  }
  // 5. complex Hermitian instantiation: Bloch-periodic Laplacian (phase picked up on the wrap-around links only),
  //    analytic spectrum sum_d 4 c_d sin^2((2 pi m_d + phi_d) / (2 n_d)); standard and generalized (diagonal M) problems
  {
    typedef std::complex<double> Z;
    typedef hostmv::HostMVC MVC;
    struct Bloch : mx::Operator<Z> {
      int n[3];
      double c[3], phi[3];
      int64_t size() const { return int64_t(n[0]) * n[1] * n[2]; }
      void Apply(const mx::MultiVec<Z>& x_, mx::MultiVec<Z>& y_) const override {
        const MVC& x = dynamic_cast<const MVC&>(x_);
        MVC& y = dynamic_cast<MVC&>(y_);
        const int64_t stride[3] = {int64_t(n[1]) * n[2], n[2], 1};
        for (int v = 0; v < x.GetNumberVecs(); ++v) {
          const Z* a = x.col(size_t(v));
          Z* b = y.col(size_t(v));
          for (int i = 0; i < n[0]; ++i)
            for (int j = 0; j < n[1]; ++j)
              for (int k = 0; k < n[2]; ++k) {
                const int idx[3] = {i, j, k};
                const int64_t p = i * stride[0] + j * stride[1] + k;
                Z s = 2 * (c[0] + c[1] + c[2]) * a[p];
                for (int d = 0; d < 3; ++d) {
                  const bool wrapUp = idx[d] == n[d] - 1, wrapDown = idx[d] == 0;
                  const int64_t up = wrapUp ? p - (n[d] - 1) * stride[d] : p + stride[d];
                  const int64_t dn = wrapDown ? p + (n[d] - 1) * stride[d] : p - stride[d];
                  s -= c[d] * (wrapUp ? std::polar(1.0, phi[d]) : Z(1.0)) * a[up];
                  s -= c[d] * (wrapDown ? std::polar(1.0, -phi[d]) : Z(1.0)) * a[dn];
                }
                b[p] = s;
              }
        }
      }
      std::vector<double> spectrum() const {
        std::vector<double> w;
        for (int i = 0; i < n[0]; ++i)
          for (int j = 0; j < n[1]; ++j)
            for (int k = 0; k < n[2]; ++k) {
              const int m[3] = {i, j, k};
              double s = 0;
              for (int d = 0; d < 3; ++d) { const double t = std::sin((2 * M_PI * m[d] + phi[d]) / (2.0 * n[d])); s += 4 * c[d] * t * t; }
              w.push_back(s);
            }
        std::sort(w.begin(), w.end());
        return w;
      }
    };
    struct DiagZ : mx::Operator<Z> {
      std::vector<double> d;
      void Apply(const mx::MultiVec<Z>& x_, mx::MultiVec<Z>& y_) const override {
        const MVC& x = dynamic_cast<const MVC&>(x_);
        MVC& y = dynamic_cast<MVC&>(y_);
        for (int v = 0; v < x.GetNumberVecs(); ++v)
          for (size_t i = 0; i < d.size(); ++i) y.col(size_t(v))[i] = d[i] * x.col(size_t(v))[i];
      }
    };
    Bloch A;
    A.n[0] = 8; A.n[1] = 7; A.n[2] = 6;
    A.c[0] = 64; A.c[1] = 49; A.c[2] = 36;
    A.phi[0] = 0.7; A.phi[1] = -0.4; A.phi[2] = 1.1;
    {  // the operator really is Hermitian and complex
      auto map = std::make_shared<hostmv::Map>(A.size());
      MVC U(map, 2), AU(map, 2);
      U.MvRandom();
      A.Apply(U, AU);
      mx::SerialDenseMatrix<int, Z> G(2, 2);
      AU.MvTransMv(Z(1.0), U, G);
      CHECK(std::abs(G(0, 1) - std::conj(G(1, 0))) < 1e-9 * std::abs(G(0, 0)) && std::fabs(G(0, 1).imag()) > 1e-6, "Bloch operator not Hermitian/complex");
    }
    typedef MxSolverT<MVC, Z> SolverC;
    const std::vector<double> want = A.spectrum();
    MxSolverParams p;
    p.nev = 5; p.blockSize = 9; p.tol = 1e-9; p.maxIters = 500;
    auto map = std::make_shared<hostmv::Map>(A.size());
    MVC X(map, size_t(p.blockSize));
    MxSolverResult r = SolverC(&A, nullptr, nullptr, p).solve(X);
    CHECK(r.converged == p.nev, "complex: converged %d of %d in %d iterations", r.converged, p.nev, r.iterations);
    for (int j = 0; j < p.nev; ++j) CHECK(std::fabs(r.eigenvalues[j] - want[j]) < 1e-8 * want[p.nev], "complex eigenvalue %d: %.12g vs %.12g", j, r.eigenvalues[j], want[j]);
    MVC AX(map, size_t(p.blockSize)), XT(X);
    A.Apply(X, AX);
    XT.MvScale(std::vector<Z>(r.eigenvalues.begin(), r.eigenvalues.end()));
    AX.MvAddMv(Z(1.0), AX, Z(-1.0), XT);
    std::vector<double> rn;
    AX.MvNorm(rn);
    for (int j = 0; j < p.nev; ++j) CHECK(rn[j] < 2e-9 * want[p.nev], "complex residual %d = %.3e", j, rn[j]);
    mx::SerialDenseMatrix<int, Z> G(p.blockSize, p.blockSize);
    X.MvTransMv(Z(1.0), X, G);
    for (int j = 0; j < p.nev; ++j)
      for (int i = 0; i < p.nev; ++i) CHECK(std::abs(G(i, j) - (i == j ? 1.0 : 0.0)) < 1e-8, "X^H X (%d,%d)", i, j);
    std::printf("complex:     %d iterations, %ld operator columns\n", r.iterations, r.applyA);
    // generalized Hermitian problem with a diagonal mass matrix and a Jacobi preconditioner
    DiagZ M, T;
    M.d.resize(size_t(A.size()));
    T.d.resize(size_t(A.size()));
    for (size_t i = 0; i < M.d.size(); ++i) { M.d[i] = 1.0 + 0.4 * std::cos(0.23 * double(i)); T.d[i] = 1.0 / (2 * (A.c[0] + A.c[1] + A.c[2])); }
    MVC Xg(map, size_t(p.blockSize));
    p.seed = 99;
    MxSolverResult rg = SolverC(&A, &M, &T, p).solve(Xg);
    CHECK(rg.converged == p.nev, "complex generalized: converged %d of %d in %d iterations", rg.converged, p.nev, rg.iterations);
    MVC AXg(map, size_t(p.blockSize)), MXg(map, size_t(p.blockSize));
    A.Apply(Xg, AXg);
    M.Apply(Xg, MXg);
    MVC MXT(MXg);
    MXT.MvScale(std::vector<Z>(rg.eigenvalues.begin(), rg.eigenvalues.end()));
    AXg.MvAddMv(Z(1.0), AXg, Z(-1.0), MXT);
    std::vector<double> mn;
    AXg.MvNorm(rn);
    MXg.MvNorm(mn);
    for (int j = 0; j < p.nev; ++j) CHECK(rn[j] < 2e-9 * std::max(rg.eigenvalues[j], 0.1 * rg.eigenvalues[p.nev - 1]) * mn[j] * 1.01, "complex generalized residual %d = %.3e", j, rn[j]);
    std::printf("complex gen: %d iterations, %ld operator columns\n", rg.iterations, rg.applyA);
  }
  if (failures == 0) std::printf("PASSED\n");
  return failures ? 1 : 0;
}
