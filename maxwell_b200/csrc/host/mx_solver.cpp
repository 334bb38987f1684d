// libmxsolver: C entry points around include/mx/MxSolver.hpp (host C++, no CUDA in this file).
#include "mxsolver.h"

#include <cstring>
#include <string>

#include "mx/MxSolver.hpp"

namespace {
thread_local std::string g_err;
thread_local double g_profile[8] = {0, 0, 0, 0, 0, 0, 0, 0};
struct Wrapped {
  std::shared_ptr<MxComm> comm;
  std::shared_ptr<MxMap> map;
};
Wrapped wrap(mxg_ctx* ctx, mxg_mv* X) {
  Wrapped w;
  w.comm = std::make_shared<MxComm>(ctx, false);
  w.map = std::make_shared<MxMap>(mxg_mv_get_map(X), w.comm, false);
  return w;
}

// Scalar = double: real symmetric pencil; Scalar = MxComplex: Hermitian pencil of a Bloch-periodic simulation (complex
// operator, complex multivectors, m_diag a complex one-column multivector with real entries).
template <class Scalar>
void runLobpcg(mxg_ctx* ctx, mxg_crs* A, mxg_mv* m_diag, mxg_gmg* prec, const mxs_projection* proj, mxg_mv* X, const mxs_params* p,
               double* evals, double* resnorms, int64_t* info, int ninfo, double* seconds, double* violation) {
  Wrapped w = wrap(ctx, X);
  MxAnasaziMV<Scalar> Xmv(X, w.map, false);
  MxCrsOperator<Scalar> Aop(A);
  std::unique_ptr<MxDiagOperator<Scalar>> Mop;
  if (m_diag) Mop.reset(new MxDiagOperator<Scalar>(m_diag));
  std::unique_ptr<MxGeoMultigridPrec<Scalar>> Top;
  if (prec) Top.reset(new MxGeoMultigridPrec<Scalar>(prec, false));
  MxSolverParams sp;
  sp.nev = p->nev;
  sp.blockSize = p->block_size > 0 ? p->block_size : mxg_mv_num_cols(X);
  sp.maxIters = p->max_iters;
  sp.tol = p->tol;
  sp.verbose = p->verbose;
  sp.seed = p->seed;
  sp.randomInit = p->random_init != 0;
  sp.profile = p->verbose >= 2;
  MxSolverT<MxAnasaziMV<Scalar>, Scalar> solver(&Aop, Mop.get(), Top.get(), sp);
  std::unique_ptr<MxGeoMultigridPrec<Scalar>> Tsca;
  std::unique_ptr<MxDivProjector<Scalar>> P;
  if (proj) {
    if (proj->sca_prec) Tsca.reset(new MxGeoMultigridPrec<Scalar>(proj->sca_prec, false));
    P.reset(new MxDivProjector<Scalar>(proj->divB, proj->gradPsi, proj->scaLapl, m_diag, Tsca.get(), w.comm,
                                       proj->tol_init > 0 ? proj->tol_init : 1e-6, proj->max_iters > 0 ? proj->max_iters : 500));
    solver.setConstraint(P.get());
  }
  if (proj) {
    MxSolverParams& q = solver.params();
    if (proj->tol_init > 0) q.projTolInit = proj->tol_init;
    if (proj->tol_w > 0) q.projTolW = proj->tol_w;
    if (proj->tol_x > 0) q.projTolX = proj->tol_x;
    if (proj->reproject_ratio > 0) q.reprojectRatio = proj->reproject_ratio;
    if (proj->max_iters_w > 0) q.projMaxItersW = proj->max_iters_w;
  }
  MxSolverResult r = solver.solve(Xmv);
  const int m = sp.blockSize;
  if (evals) std::memcpy(evals, r.eigenvalues.data(), sizeof(double) * m);
  if (resnorms) std::memcpy(resnorms, r.residuals.data(), sizeof(double) * m);
  if (info) {
    const int64_t all[8] = {r.iterations, r.converged, r.applyA, r.applyPrec, r.projections, r.reprojections,
                            P ? int64_t(P->numLinIters) : 0, P ? int64_t(P->numApplies) : 0};
    for (int i = 0; i < ninfo && i < 8; ++i) info[i] = all[i];
  }
  if (seconds) *seconds = r.seconds;
  if (violation)
    for (int j = 0; j < m; ++j) violation[j] = j < int(r.violation.size()) ? r.violation[j] : 0.0;
  g_profile[0] = r.tApplyA; g_profile[1] = r.tPrec; g_profile[2] = r.tGram; g_profile[3] = r.tUpdate; g_profile[4] = r.tProj;
}

template <class Scalar>
void runCheck(mxg_ctx* ctx, mxg_crs* A, mxg_mv* m_diag, mxg_crs* divB, mxg_mv* X, const double* evals, double* res, double* div) {
  Wrapped w = wrap(ctx, X);
  MxAnasaziMV<Scalar> Xmv(X, w.map, false);
  const int m = Xmv.GetNumberVecs();
  MxAnasaziMV<Scalar> AX(w.map, m), MX(w.map, m);
  MxCrsOperator<Scalar>(A).Apply(Xmv, AX);
  if (m_diag) MxDiagOperator<Scalar>(m_diag).Apply(Xmv, MX); else MX = Xmv;
  std::vector<Scalar> th(evals, evals + m);
  std::vector<double> rn, mn;
  MxAnasaziMV<Scalar> scaled(MX);
  scaled.MvScale(th);
  AX.MvAddMv(Scalar(1.0), AX, Scalar(-1.0), scaled);
  AX.MvNorm(rn);
  MX.MvNorm(mn);
  for (int j = 0; j < m; ++j) res[j] = rn[j] / std::fabs(evals[j]);   // MxMagWaveOp.cpp:1195-1203
  if (divB && div) {
    mxg_mv* d = nullptr;
    mx::check(mxg_mv_create(mxg_crs_row_map(divB), m, mxg_mv_is_complex(X), &d));
    mx::check(mxg_crs_apply(divB, MX.getRawMV(), d));
    std::vector<double> dn(m);
    mx::check(mxg_mv_norm2(d, dn.data()));
    mxg_mv_destroy(d);
    for (int j = 0; j < m; ++j) div[j] = mn[j] > 0 ? dn[j] / mn[j] : 0.0;
  }
}
}  // namespace

extern "C" {

void mxs_default_params(mxs_params* p) {
  if (!p) return;
  MxSolverParams d;
  p->nev = d.nev;
  p->block_size = 0;
  p->max_iters = d.maxIters;
  p->tol = d.tol;
  p->verbose = 0;
  p->seed = d.seed;
  p->random_init = 1;
}

const char* mxs_last_error(void) { return g_err.c_str(); }
void mxs_last_profile(double out[4]) { for (int i = 0; i < 4; ++i) out[i] = g_profile[i]; }
void mxs_last_profile_ex(double out[8]) { for (int i = 0; i < 8; ++i) out[i] = g_profile[i]; }

int mxs_lobpcg(mxg_ctx* ctx, mxg_crs* A, mxg_mv* m_diag, mxg_gmg* prec, mxg_mv* X, const mxs_params* p,
               double* evals, double* resnorms, int64_t info[4], double* seconds) {
  try {
    if (!ctx || !A || !X || !p) throw std::runtime_error("mxs_lobpcg: NULL argument");
    if (mxg_mv_is_complex(X)) runLobpcg<MxComplex>(ctx, A, m_diag, prec, nullptr, X, p, evals, resnorms, info, 4, seconds, nullptr);
    else runLobpcg<double>(ctx, A, m_diag, prec, nullptr, X, p, evals, resnorms, info, 4, seconds, nullptr);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

int mxs_lobpcg_projected(mxg_ctx* ctx, mxg_crs* A, mxg_mv* m_diag, mxg_gmg* prec, const mxs_projection* proj, mxg_mv* X,
                         const mxs_params* p, double* evals, double* resnorms, double* violation, int64_t info[8], double* seconds) {
  try {
    if (!ctx || !A || !X || !p || !proj) throw std::runtime_error("mxs_lobpcg_projected: NULL argument");
    if (!proj->divB || !proj->gradPsi || !proj->scaLapl) throw std::runtime_error("mxs_lobpcg_projected: projection operators missing");
    if (mxg_mv_is_complex(X)) runLobpcg<MxComplex>(ctx, A, m_diag, prec, proj, X, p, evals, resnorms, info, 8, seconds, violation);
    else runLobpcg<double>(ctx, A, m_diag, prec, proj, X, p, evals, resnorms, info, 8, seconds, violation);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

int mxs_check_eigensolution(mxg_ctx* ctx, mxg_crs* A, mxg_mv* m_diag, mxg_crs* divB, mxg_mv* X, const double* evals,
                            double* res, double* div) {
  try {
    if (!ctx || !A || !X || !evals) throw std::runtime_error("mxs_check_eigensolution: NULL argument");
    if (mxg_mv_is_complex(X)) runCheck<MxComplex>(ctx, A, m_diag, divB, X, evals, res, div);
    else runCheck<double>(ctx, A, m_diag, divB, X, evals, res, div);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

}  // extern "C"

// MxMagWaveOp::Apply (reference src/MxMagWaveOp.cpp:825-943): Y = P (L - sigma M)^-1 M X.
// info[0] = vector-solve Krylov iterations, info[1] = scalar-solve CG iterations.
namespace {
template <class Scalar>
void runMagWave(mxg_ctx* ctx, mxg_crs* vecLapl, mxg_mv* m_diag, mxg_crs* divB, mxg_crs* gradPsi, mxg_crs* scaLapl, mxg_gmg* vecPrec,
                mxg_gmg* scaPrec, double shift, double linTol, int hasCurlNull, int linSolver, int linBasis, int maxIters, mxg_mv* X,
                mxg_mv* Y, int64_t info[2]) {
  Wrapped w = wrap(ctx, X);
  MxAnasaziMV<Scalar> Xmv(X, w.map, false), Ymv(Y, w.map, false);
  std::unique_ptr<MxGeoMultigridPrec<Scalar>> Tv, Ts;
  if (vecPrec) Tv.reset(new MxGeoMultigridPrec<Scalar>(vecPrec, false));
  if (scaPrec) Ts.reset(new MxGeoMultigridPrec<Scalar>(scaPrec, false));
  MxMagWaveOpParams p;
  p.shift = shift;
  p.linTol = linTol;
  p.hasCurlNull = hasCurlNull != 0;
  p.linSolver = linSolver == 1 ? mx::LIN_BICGSTAB : (linSolver == 2 ? mx::LIN_GMRES : mx::LIN_CG);
  if (linBasis > 0) p.linBasis = linBasis;
  if (maxIters > 0) p.maxLinIters = maxIters;
  MxMagWaveOpT<Scalar> op(vecLapl, m_diag, divB, gradPsi, scaLapl, Tv.get(), Ts.get(), w.comm, p);
  op.Apply(Xmv, Ymv);
  if (info) { info[0] = op.numVecLinIters; info[1] = op.numScaLinIters(); }
}
}  // namespace

extern "C" int mxs_magwave_apply_ex(mxg_ctx* ctx, mxg_crs* vecLapl, mxg_mv* m_diag, mxg_crs* divB, mxg_crs* gradPsi, mxg_crs* scaLapl,
                                    mxg_gmg* vecPrec, mxg_gmg* scaPrec, double shift, double linTol, int hasCurlNull, int linSolver,
                                    int linBasis, int maxIters, mxg_mv* X, mxg_mv* Y, int64_t info[2]) {
  try {
    if (!ctx || !vecLapl || !m_diag || !X || !Y) throw std::runtime_error("mxs_magwave_apply: NULL argument");
    if (hasCurlNull && (!divB || !gradPsi || !scaLapl)) throw std::runtime_error("mxs_magwave_apply: projection operators missing");
    if (mxg_mv_is_complex(X))
      runMagWave<MxComplex>(ctx, vecLapl, m_diag, divB, gradPsi, scaLapl, vecPrec, scaPrec, shift, linTol, hasCurlNull, linSolver, linBasis, maxIters, X, Y, info);
    else
      runMagWave<double>(ctx, vecLapl, m_diag, divB, gradPsi, scaLapl, vecPrec, scaPrec, shift, linTol, hasCurlNull, linSolver, linBasis, maxIters, X, Y, info);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
extern "C" int mxs_magwave_apply(mxg_ctx* ctx, mxg_crs* vecLapl, mxg_mv* m_diag, mxg_crs* divB, mxg_crs* gradPsi, mxg_crs* scaLapl,
                                 mxg_gmg* vecPrec, mxg_gmg* scaPrec, double shift, double linTol, int hasCurlNull,
                                 mxg_mv* X, mxg_mv* Y, int64_t info[2]) {
  return mxs_magwave_apply_ex(ctx, vecLapl, m_diag, divB, gradPsi, scaLapl, vecPrec, scaPrec, shift, linTol, hasCurlNull, 0, 0, 0, X, Y, info);
}

// the projection alone: X <- P X = X + gradPsi scaLapl^-1 divB M X (src/MxMagWaveOp.cpp:893-924); info[0] = CG iterations
extern "C" int mxs_div_project(mxg_ctx* ctx, mxg_mv* m_diag, const mxs_projection* proj, double tol, mxg_mv* X, int64_t info[1]) {
  try {
    if (!ctx || !proj || !X || !proj->divB || !proj->gradPsi || !proj->scaLapl) throw std::runtime_error("mxs_div_project: NULL argument");
    Wrapped w = wrap(ctx, X);
    auto run = [&](auto tag) {
      typedef decltype(tag) Scalar;
      MxAnasaziMV<Scalar> Xmv(X, w.map, false);
      std::unique_ptr<MxGeoMultigridPrec<Scalar>> Ts;
      if (proj->sca_prec) Ts.reset(new MxGeoMultigridPrec<Scalar>(proj->sca_prec, false));
      MxDivProjector<Scalar> P(proj->divB, proj->gradPsi, proj->scaLapl, m_diag, Ts.get(), w.comm, tol, proj->max_iters > 0 ? proj->max_iters : 500);
      P.project(Xmv, tol);
      if (info) info[0] = P.numLinIters;
    };
    if (mxg_mv_is_complex(X)) run(MxComplex()); else run(double());
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// MxMagWaveOp::magToElec (reference src/MxMagWaveOp.cpp:1237-1250): E = [invEps] curlB B. invEps NULL = vacuum.
extern "C" int mxs_mag_to_elec(mxg_ctx* ctx, mxg_crs* curlB, mxg_crs* invEps, mxg_mv* mag, mxg_mv* elec) {
  try {
    if (!ctx || !curlB || !mag || !elec) throw std::runtime_error("mxs_mag_to_elec: NULL argument");
    Wrapped wm = wrap(ctx, mag), we = wrap(ctx, elec);
    MxAnasaziMV<double> B(mag, wm.map, false), E(elec, we.map, false);
    MxMagWaveOpT<double>::magToElec(curlB, invEps, B, E);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// MxMagWaveOp::eigValsToFreqs (reference src/MxMagWaveOp.cpp:1252-1271); host arithmetic only.
extern "C" int mxs_eigvals_to_freqs(const double* re, const double* im, int n, double shift, int invert, double* fre, double* fim) {
  try {
    if (n < 0 || (n > 0 && (!re || !fre || !fim))) throw std::runtime_error("mxs_eigvals_to_freqs: bad argument");
    std::vector<std::complex<double>> ev(n), fr;
    for (int i = 0; i < n; ++i) ev[i] = std::complex<double>(re[i], im ? im[i] : 0.0);
    MxMagWaveOpT<double>::eigValsToFreqs(ev, fr, shift, invert != 0);
    for (int i = 0; i < n; ++i) { fre[i] = fr[i].real(); fim[i] = fr[i].imag(); }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
