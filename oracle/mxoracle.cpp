// TEST INFRASTRUCTURE ONLY -- CPU oracle for maxwell_b200 (see mxo_geom.hpp header).
// C entry points so tests/ and bench.py's cpu_baseline leg can drive the oracle via ctypes.
#include <cstring>
#include <omp.h>

#include "mxo_ops.hpp"

using namespace mxo;

namespace {
thread_local std::string g_err;
struct Mat {
  bool cplx_ = false;
  Csr<double> r;
  Csr<cplx> c;
};
template <class F> int guard(F&& f) {
  try { f(); return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
}  // namespace

extern "C" {

const char* mxo_last_error() { return g_err.c_str(); }

// ---- shapes -----------------------------------------------------------------------------
void* mxo_shape_cylinder(double r, const double axis[3], const double loc[3]) {
  return new std::shared_ptr<Shape>(new Cylinder(r, {axis[0], axis[1], axis[2]}, {loc[0], loc[1], loc[2]}));
}
void* mxo_shape_sphere(double r, const double loc[3]) {
  return new std::shared_ptr<Shape>(new Sphere(r, {loc[0], loc[1], loc[2]}));
}
void* mxo_shape_slab(double thickness, const double n[3], const double loc[3]) {
  return new std::shared_ptr<Shape>(makeSlab(thickness, {n[0], n[1], n[2]}, {loc[0], loc[1], loc[2]}));
}
void* mxo_shape_halfspace(const double point[3], const double n[3]) {
  return new std::shared_ptr<Shape>(new HalfSpace({point[0], point[1], point[2]}, {n[0], n[1], n[2]}));
}
void* mxo_shape_intersection(void* const* shapes, int n) {
  auto is = std::make_shared<Intersection>();
  for (int i = 0; i < n; ++i) is->subs.push_back(*static_cast<std::shared_ptr<Shape>*>(shapes[i]));
  return new std::shared_ptr<Shape>(is);
}
static std::shared_ptr<Shape>& shp(void* h) { return *static_cast<std::shared_ptr<Shape>*>(h); }
void* mxo_shape_torus(double majorRadius, double minorRadius, const double axis[3], const double loc[3]) {
  return new std::shared_ptr<Shape>(new Torus(majorRadius, minorRadius, {axis[0], axis[1], axis[2]}, {loc[0], loc[1], loc[2]}));
}
void* mxo_shape_cone(double angle, const double axis[3], const double vertex[3]) {
  return new std::shared_ptr<Shape>(new Cone(angle, {axis[0], axis[1], axis[2]}, {vertex[0], vertex[1], vertex[2]}));
}
void* mxo_shape_union(void* const* shapes, int n) {
  auto u = std::make_shared<Union>();
  for (int i = 0; i < n; ++i) u->subs.push_back(shp(shapes[i]));
  return new std::shared_ptr<Shape>(u);
}
void* mxo_shape_subtract(void* base, void* const* removed, int n) {
  auto d = std::make_shared<Subtract>();
  d->base = shp(base);
  for (int i = 0; i < n; ++i) d->rm.subs.push_back(shp(removed[i]));
  return new std::shared_ptr<Shape>(d);
}
void* mxo_shape_mirror(void* shape, const double normal[3], const double point[3]) {
  return new std::shared_ptr<Shape>(new Mirror(shp(shape), {normal[0], normal[1], normal[2]}, {point[0], point[1], point[2]}));
}
void* mxo_shape_repeat(void* shape, const double origin[3], const double dir[3], double step, int numPos, int numNeg) {
  return new std::shared_ptr<Shape>(
      new Repeat(shp(shape), {origin[0], origin[1], origin[2]}, {dir[0], dir[1], dir[2]}, step, numPos, numNeg));
}
void* mxo_shape_ellipsoid(const double loc[3], const double axes[3]) {
  return new std::shared_ptr<Shape>(new Ellipsoid({loc[0], loc[1], loc[2]}, {axes[0], axes[1], axes[2]}));
}
void mxo_shape_rotate(void* sh, const double axis[3], double angle, const double* pivot) {
  if (pivot) shp(sh)->rotate({axis[0], axis[1], axis[2]}, angle, {pivot[0], pivot[1], pivot[2]});
  else shp(sh)->rotate({axis[0], axis[1], axis[2]}, angle);
}
void mxo_shape_scale(void* sh, const double mags[3], const double origin[3]) {
  shp(sh)->scale({mags[0], mags[1], mags[2]}, {origin[0], origin[1], origin[2]});
}
void mxo_shape_translate(void* sh, const double v[3]) { shp(sh)->translate({v[0], v[1], v[2]}); }
void mxo_shape_reflect(void* sh, const double normal[3], const double point[3]) {
  shp(sh)->reflect({normal[0], normal[1], normal[2]}, {point[0], point[1], point[2]});
}
void mxo_shape_grad(void* sh, const double p[3], double g[3]) {
  const D3 r = shp(sh)->grad({p[0], p[1], p[2]});
  g[0] = r[0]; g[1] = r[1]; g[2] = r[2];
}
void mxo_shape_invert(void* sh) { (*static_cast<std::shared_ptr<Shape>*>(sh))->sign *= -1.0; }
double mxo_shape_func(void* sh, const double p[3]) { return (*static_cast<std::shared_ptr<Shape>*>(sh))->func({p[0], p[1], p[2]}); }
void mxo_shape_destroy(void* sh) { delete static_cast<std::shared_ptr<Shape>*>(sh); }

// cut-cell fractions exposed for unit tests: kind 0 = edge (axis), 1 = face (type), 2 = box
double mxo_fraction(void* sh, int kind, int axis, const double len[3], const double p[3]) {
  const Shape& s = **static_cast<std::shared_ptr<Shape>*>(sh);
  const D3 pp{p[0], p[1], p[2]};
  if (kind == 0) return segmentFraction(s, axis, len[0], pp);
  if (kind == 1) return rectFraction(s, axis, len[0], len[1], pp);
  return boxFraction(s, len[0], len[1], len[2], pp);
}

// ---- simulation -------------------------------------------------------------------------
void* mxo_sim_create(const int n[3], const double origin[3], const double size[3]) {
  return new Sim({n[0], n[1], n[2]}, {origin[0], origin[1], origin[2]}, {size[0], size[1], size[2]});
}
void mxo_sim_destroy(void* s) { delete static_cast<Sim*>(s); }
void mxo_sim_set_bcs(void* s, const int lower[3], const int upper[3]) {
  Sim* sim = static_cast<Sim*>(s);
  for (int i = 0; i < 3; ++i) { sim->lower[i] = BCType(lower[i]); sim->upper[i] = BCType(upper[i]); }
}
void mxo_sim_set_phase_shifts(void* s, const double ph[3]) {
  Sim* sim = static_cast<Sim*>(s);
  for (int i = 0; i < 3; ++i) sim->phaseShifts[i] = ph[i];
}
void mxo_sim_set_pec(void* s, void* shape) { static_cast<Sim*>(s)->pec = *static_cast<std::shared_ptr<Shape>*>(shape); }
// eps: 9 complex entries (18 doubles), row-major
void mxo_sim_add_dielectric(void* s, void* shape, const double* eps, const char* name) {
  Dielectric d;
  d.name = name;
  d.shape = *static_cast<std::shared_ptr<Shape>*>(shape);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) d.eps[i][j] = cplx(eps[2 * (3 * i + j)], eps[2 * (3 * i + j) + 1]);
  static_cast<Sim*>(s)->diels.push_back(d);
}
void mxo_sim_set_literal_upper_periodic_e(void* s, int on) { static_cast<Sim*>(s)->literalUpperPeriodicE = on != 0; }
int mxo_sim_setup(void* s) { return guard([&] { static_cast<Sim*>(s)->setup(); }); }

static const Field* fieldOf(const Sim* sim, const char* name) {
  const std::string n(name);
  if (n == "bfield") return sim->B.get();
  if (n == "efield") return sim->E.get();
  if (n == "dfield" && sim->D) return sim->D.get();
  if (n == "psifield") return sim->Psi.get();
  throw std::runtime_error("mxo: unknown field '" + n + "'");
}
int64_t mxo_sim_map_size(void* s, const char* field) {
  int64_t n = -1;
  guard([&] { n = int64_t(fieldOf(static_cast<Sim*>(s), field)->gids.size()); });
  return n;
}
int mxo_sim_map_copy(void* s, const char* field, int64_t* out) {
  return guard([&] {
    const Field* f = fieldOf(static_cast<Sim*>(s), field);
    std::memcpy(out, f->gids.data(), f->gids.size() * sizeof(int64_t));
  });
}
int64_t mxo_sim_map_num_global(void* s, const char* field) {
  int64_t n = -1;
  guard([&] { n = fieldOf(static_cast<Sim*>(s), field)->numGlobal(); });
  return n;
}
// PEC fraction of every DOF in a field's map (1.0 when no PEC)
int mxo_sim_map_fracs(void* s, const char* field, double* out) {
  return guard([&] {
    const Field* f = fieldOf(static_cast<Sim*>(s), field);
    for (size_t i = 0; i < f->gids.size(); ++i) {
      I3 cell; int c;
      cellCompOf(*f, f->gids[i], cell, c);
      out[i] = f->frac(c, cell, "pec");
    }
  });
}

// fractions of a named shape representation ("pec", "diel0", ...) on the guarded block: (N+3)^3 cells x components.
// Returns the value count (0 when the field has no such representation); `out` may be NULL to query it.
int64_t mxo_sim_rep_copy(void* s, const char* field, const char* rep, double* out) {
  int64_t n = -1;
  guard([&] {
    const Field* f = fieldOf(static_cast<Sim*>(s), field);
    auto it = f->reps.find(rep);
    if (it == f->reps.end()) { n = 0; return; }
    n = int64_t(it->second.size());
    if (out) std::memcpy(out, it->second.data(), it->second.size() * sizeof(double));
  });
  return n;
}

// ---- operators --------------------------------------------------------------------------
void* mxo_build_op(void* s, const char* name, int is_complex) {
  Mat* m = new Mat;
  m->cplx_ = is_complex != 0;
  const int rc = guard([&] {
    if (m->cplx_) m->c = buildOp<cplx>(*static_cast<Sim*>(s), name);
    else m->r = buildOp<double>(*static_cast<Sim*>(s), name);
  });
  if (rc != 0) { delete m; return nullptr; }
  return m;
}
// the reference's real 2n x 2n storage of a complex matrix
void* mxo_mat_kform(void* h) {
  Mat* src = static_cast<Mat*>(h);
  if (!src->cplx_) { g_err = "mxo: kform needs a complex matrix"; return nullptr; }
  Mat* m = new Mat;
  m->r = toKForm(src->c);
  return m;
}
void* mxo_mat_multiply(void* a, void* b) {
  Mat *A = static_cast<Mat*>(a), *B = static_cast<Mat*>(b);
  Mat* m = new Mat;
  m->cplx_ = A->cplx_;
  const int rc = guard([&] {
    if (A->cplx_ != B->cplx_) throw std::runtime_error("mxo: mixed scalar types");
    if (A->cplx_) m->c = multiply(A->c, B->c); else m->r = multiply(A->r, B->r);
  });
  if (rc != 0) { delete m; return nullptr; }
  return m;
}
// interpolation of field `field` of simFrom at the component positions of simTo
// (MxGridFieldInterpolator; the refiners / coarseners of MxGeoMultigridPrec.cpp:438-452)
void* mxo_build_interp(void* simFrom, void* simTo, const char* field, int is_complex) {
  Mat* m = new Mat;
  m->cplx_ = is_complex != 0;
  const int rc = guard([&] {
    const Field* f = fieldOf(static_cast<Sim*>(simFrom), field);
    const Field* t = fieldOf(static_cast<Sim*>(simTo), field);
    if (m->cplx_) m->c = interpolator<cplx>(*f, *t); else m->r = interpolator<double>(*f, *t);
  });
  if (rc != 0) { delete m; return nullptr; }
  return m;
}
void* mxo_mat_transpose(void* h, int normalize_rows) {
  Mat* src = static_cast<Mat*>(h);
  Mat* m = new Mat;
  m->cplx_ = src->cplx_;
  if (src->cplx_) { m->c = transpose(src->c); if (normalize_rows) normalizeRows(m->c); }
  else { m->r = transpose(src->r); if (normalize_rows) normalizeRows(m->r); }
  return m;
}
// C = sa*A + sb*B (EpetraExt::MatrixMatrix::Add semantics; shifted operators, MxMagWaveOp.cpp:247-256)
void* mxo_mat_add(void* a, const double sa[2], void* b, const double sb[2]) {
  Mat *A = static_cast<Mat*>(a), *B = static_cast<Mat*>(b);
  Mat* m = new Mat;
  m->cplx_ = A->cplx_;
  const int rc = guard([&] {
    if (A->cplx_ != B->cplx_) throw std::runtime_error("mxo: mixed scalar types");
    if (A->cplx_) m->c = add(A->c, cplx(sa[0], sa[1]), B->c, cplx(sb[0], sb[1]));
    else m->r = add(A->r, sa[0], B->r, sb[0]);
  });
  if (rc != 0) { delete m; return nullptr; }
  return m;
}
void mxo_mat_scale(void* h, const double s[2]) {
  Mat* m = static_cast<Mat*>(h);
  if (m->cplx_) scaleInPlace(m->c, cplx(s[0], s[1])); else scaleInPlace(m->r, s[0]);
}
void mxo_mat_destroy(void* h) { delete static_cast<Mat*>(h); }
int mxo_mat_is_complex(void* h) { return static_cast<Mat*>(h)->cplx_ ? 1 : 0; }
void mxo_mat_shape(void* h, int64_t* nrows, int64_t* ncols, int64_t* nnz) {
  Mat* m = static_cast<Mat*>(h);
  if (m->cplx_) { *nrows = m->c.nrows; *ncols = m->c.ncols; *nnz = m->c.nnz(); }
  else { *nrows = m->r.nrows; *ncols = m->r.ncols; *nnz = m->r.nnz(); }
}
// vals: nnz doubles (real) or 2*nnz doubles (complex, interleaved)
void mxo_mat_copy(void* h, int64_t* rowptr, int32_t* col, double* vals) {
  Mat* m = static_cast<Mat*>(h);
  if (m->cplx_) {
    std::memcpy(rowptr, m->c.rowptr.data(), m->c.rowptr.size() * sizeof(int64_t));
    std::memcpy(col, m->c.col.data(), m->c.col.size() * sizeof(int32_t));
    std::memcpy(vals, m->c.val.data(), m->c.val.size() * sizeof(cplx));
  } else {
    std::memcpy(rowptr, m->r.rowptr.data(), m->r.rowptr.size() * sizeof(int64_t));
    std::memcpy(col, m->r.col.data(), m->r.col.size() * sizeof(int32_t));
    std::memcpy(vals, m->r.val.data(), m->r.val.size() * sizeof(double));
  }
}
void mxo_mat_maps(void* h, int64_t* rowGid, int64_t* colGid) {
  Mat* m = static_cast<Mat*>(h);
  const auto& rg = m->cplx_ ? m->c.rowGid : m->r.rowGid;
  const auto& cg = m->cplx_ ? m->c.colGid : m->r.colGid;
  if (rowGid) std::memcpy(rowGid, rg.data(), rg.size() * sizeof(int64_t));
  if (colGid) std::memcpy(colGid, cg.data(), cg.size() * sizeof(int64_t));
}
// Y = A X (column-major, ld in scalars). nthreads <= 0 keeps the OpenMP default.
int mxo_mat_apply(void* h, const double* X, int64_t ldx, double* Y, int64_t ldy, int nvec, int nthreads) {
  Mat* m = static_cast<Mat*>(h);
  return guard([&] {
    const int prev = omp_get_max_threads();
    if (nthreads > 0) omp_set_num_threads(nthreads);
    if (m->cplx_) spmm(m->c, reinterpret_cast<const cplx*>(X), ldx, reinterpret_cast<cplx*>(Y), ldy, nvec);
    else spmm(m->r, X, ldx, Y, ldy, nvec);
    omp_set_num_threads(prev);
  });
}
int mxo_num_threads() { return omp_get_max_threads(); }

// Generic CSR apply on caller-provided arrays (used for the CPU baseline on sub-blocks)
int mxo_csr_apply(int64_t nrows, const int64_t* rowptr, const int32_t* col, const double* val, int is_complex,
                  const double* X, int64_t ldx, double* Y, int64_t ldy, int nvec, int nthreads) {
  return guard([&] {
    const int prev = omp_get_max_threads();
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nrows; ++i) {
      for (int v = 0; v < nvec; ++v) {
        if (is_complex) {
          const cplx* x = reinterpret_cast<const cplx*>(X) + v * ldx;
          const cplx* a = reinterpret_cast<const cplx*>(val);
          cplx sum(0.0);
          for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) sum += a[p] * x[col[p]];
          reinterpret_cast<cplx*>(Y)[i + v * ldy] = sum;
        } else {
          const double* x = X + v * ldx;
          double sum = 0.0;
          for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) sum += val[p] * x[col[p]];
          Y[i + v * ldy] = sum;
        }
      }
    }
    omp_set_num_threads(prev);
  });
}

}  // extern "C"
