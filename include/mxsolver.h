/* libmxsolver -- C entry points of the host-side eigensolver driver (C++: include/mx/MxSolver.hpp).
 * Replaces MxSolver::solve + the Anasazi solver manager it drives (reference src/MxSolver.cpp:100-239):
 * lowest eigenpairs of A x = theta M x on the GPU through libmxgpu. Plain handles only. */
#ifndef MXSOLVER_H
#define MXSOLVER_H
#include "mxgpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mxs_params {
  int nev;          /* wanted eigenpairs ("eigensolver : nev", MxSolver.cpp:37) */
  int block_size;   /* 0 = nev + max(4, nev/2) */
  int max_iters;
  double tol;       /* |A x - theta M x| / (|theta| |M x|) */
  int verbose;
  uint64_t seed;
  int random_init;  /* 1: start from MvRandom (MxSolver.cpp:62-64), 0: use X as given */
} mxs_params;

/* The divergence-cleaning projection of MxMagWaveOp::Apply (MxMagWaveOp.cpp:893-924), P b = b + gradPsi scaLapl^-1 divB M b,
 * as a constraint of the eigensolver: the iteration stays in { b : divB M b = 0 }, so only Maxwell modes are returned
 * (the reference gets the same from applying P inside its shift-invert operator). */
typedef struct mxs_projection {
  mxg_crs* divB;          /* psi <- B  (MxYeeDeyMittraDivB) */
  mxg_crs* gradPsi;       /* B <- psi  (MxYeeDeyMittraGradPsi) */
  mxg_crs* scaLapl;       /* -(divB M gradPsi) (MxMagWaveOp.cpp:208-223) */
  mxg_gmg* sca_prec;      /* multigrid on the scalar hierarchy, or NULL = Jacobi */
  double tol_init;        /* relative accuracy of the inner CG: initial block (0 = 1e-6) */
  double tol_w;           /* ... preconditioned residuals, every iteration (0 = 0.1) */
  double tol_x;           /* ... re-projection of the iterate (0 = 1e-2) */
  double reproject_ratio; /* re-project X when |D M x|/|M x| > ratio * max(relative residual, tol) (0 = 0.05) */
  int max_iters;          /* inner CG iteration cap (0 = 500) */
  int max_iters_w;        /* cap for the per-iteration projections of the preconditioned residuals (0 = none) */
} mxs_projection;

void mxs_default_params(mxs_params* p);
const char* mxs_last_error(void);
/* per-phase seconds of the last mxs_lobpcg call made with verbose >= 2 (which synchronises around phases):
 * out[0]=operator applies, [1]=preconditioner, [2]=Gram products, [3]=basis updates */
void mxs_last_profile(double out[4]);
/* A: assembled operator (curlCurl or vecLapl); m_diag: one-column multivector holding the diagonal
 * of the mass matrix mRhs (NULL = identity); prec: multigrid preconditioner (NULL = none).
 * X: n x block multivector, receives the M-orthonormal Ritz vectors. Real symmetric pencil, or -- when X is complex --
 * the Hermitian pencil of a Bloch-periodic simulation (A complex, m_diag a complex one-column multivector).
 * evals/resnorms: block entries. info[0]=iterations, [1]=converged among nev, [2]=A applies (columns),
 * [3]=preconditioner applies (columns). */
int mxs_lobpcg(mxg_ctx* ctx, mxg_crs* A, mxg_mv* m_diag, mxg_gmg* prec, mxg_mv* X, const mxs_params* p,
               double* evals, double* resnorms, int64_t info[4], double* seconds);
/* Same solver constrained to the divergence-free fields (mxs_projection above): what MxSolver::solve returns in the reference
 * (MxSolver.cpp:85-103 on MxMagWaveOp::Apply). violation[j] = |divB M x_j| / |M x_j| of the returned vectors (may be NULL).
 * info[0..3] as mxs_lobpcg, [4] = projected columns, [5] = re-projections of the iterate, [6] = inner CG iterations,
 * [7] = projection calls. */
int mxs_lobpcg_projected(mxg_ctx* ctx, mxg_crs* A, mxg_mv* m_diag, mxg_gmg* prec, const mxs_projection* proj, mxg_mv* X,
                         const mxs_params* p, double* evals, double* resnorms, double* violation, int64_t info[8], double* seconds);
/* X <- P X alone; info[0] = inner CG iterations */
int mxs_div_project(mxg_ctx* ctx, mxg_mv* m_diag, const mxs_projection* proj, double tol, mxg_mv* X, int64_t info[1]);
/* out[0..3] as mxs_last_profile, out[4] = constraint projections */
void mxs_last_profile_ex(double out[8]);
/* acceptance metrics of MxMagWaveOp::checkEigensolution / checkDivergences (MxMagWaveOp.cpp:1118-1234):
 * res[j] = |A x_j - theta_j M x_j|_2 / |theta_j| ; div[j] = |D M x_j|_2 / |M x_j|_2 (D may be NULL) */
int mxs_check_eigensolution(mxg_ctx* ctx, mxg_crs* A, mxg_mv* m_diag, mxg_crs* divB, mxg_mv* X, const double* evals,
                            double* res, double* div);
/* MxMagWaveOp::Apply (MxMagWaveOp.cpp:825-943): Y = P (L - sigma M)^-1 M X with L = vec_lapl, M = diag(m_diag) and the
 * divergence-cleaning projection P b = b + gradPsi scaLapl^-1 divB M b (has_curl_null != 0). Both inner solves are
 * block preconditioned CG on the GPU (vec_prec / sca_prec: multigrid handles or NULL; NULL scalar preconditioner = Jacobi).
 * Valid for sigma below the lowest eigenvalue (the reference's automatic shift). info[0]/[1]: CG iterations. */
int mxs_magwave_apply(mxg_ctx* ctx, mxg_crs* vec_lapl, mxg_mv* m_diag, mxg_crs* divB, mxg_crs* gradPsi, mxg_crs* scaLapl,
                      mxg_gmg* vec_prec, mxg_gmg* sca_prec, double shift, double lin_tol, int has_curl_null,
                      mxg_mv* X, mxg_mv* Y, int64_t info[2]);
/* Same with the inner solver of the vector system chosen as in "linear solver : type" (MxMagWaveOp.cpp:326-338):
 * lin_solver 0 = CG, 1 = BiCGStab, 2 = restarted GMRES(lin_basis) -- the last two accept shifts inside the spectrum.
 * Real or complex (Bloch-periodic) operands. lin_basis / max_iters 0 = defaults (20 / 1000). */
int mxs_magwave_apply_ex(mxg_ctx* ctx, mxg_crs* vec_lapl, mxg_mv* m_diag, mxg_crs* divB, mxg_crs* gradPsi, mxg_crs* scaLapl,
                         mxg_gmg* vec_prec, mxg_gmg* sca_prec, double shift, double lin_tol, int has_curl_null, int lin_solver,
                         int lin_basis, int max_iters, mxg_mv* X, mxg_mv* Y, int64_t info[2]);
/* MxMagWaveOp::magToElec (MxMagWaveOp.cpp:1237-1250): elec = [invEps] curlB mag; invEps NULL when there is no dielectric. */
int mxs_mag_to_elec(mxg_ctx* ctx, mxg_crs* curlB, mxg_crs* invEps, mxg_mv* mag, mxg_mv* elec);
/* MxMagWaveOp::eigValsToFreqs (MxMagWaveOp.cpp:1252-1271): f = sqrt(k2) c / 2 pi [Hz], k2 = ev + shift, or 1/ev + shift for the
 * shift-inverted operator. im may be NULL (real eigenvalues). Host arithmetic. */
int mxs_eigvals_to_freqs(const double* re, const double* im, int n, double shift, int invert, double* fre, double* fim);
#ifdef __cplusplus
}
#endif
#endif
