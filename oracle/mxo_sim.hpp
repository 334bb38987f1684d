// TEST INFRASTRUCTURE ONLY -- CPU oracle for maxwell_b200 (see mxo_geom.hpp header).
//
// Grid indexing, Yee field layouts, boundary-condition / Bloch factors and DOF maps,
// restated from the reference (paths relative to /root/reference/src).
#pragma once
#include <complex>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>

#include "mxo_geom.hpp"

namespace mxo {

using cplx = std::complex<double>;

enum BCType { PERIODIC = 0, ZERO = 1, CONSTANT = 2, PEC = 3, PMC = 4 };

// MxGrid.h:96-128, MxGrid.cpp:11-24,268-273
struct Grid {
  I3 N;
  D3 origin, size, d;
  Grid(I3 n, D3 o, D3 l) : N(n), origin(o), size(l) {
    for (int i = 0; i < 3; ++i) d[i] = size[i] / double(N[i]);
  }
  // (N+1)^3 node indexing, z fastest, out-of-range indices wrap (MxGrid.h:96-114)
  int64_t cellToGlobal(const I3& c) const {
    int64_t res = 0, factor = 1;
    for (int i = 2; i >= 0; --i) {
      const int ni = N[i] + 1;
      int g = c[i];
      if (g >= ni) g -= ni; else if (g < 0) g += ni;
      res += int64_t(g) * factor;
      factor *= ni;
    }
    return res;
  }
  I3 globalToCell(int64_t idx) const {
    I3 c;
    int64_t factor = 1;
    for (int i = 2; i >= 0; --i) {
      const int ni = N[i] + 1;
      c[i] = int((idx / factor) % ni);
      factor *= ni;
    }
    return c;
  }
  D3 nodeCoord(const I3& c) const {
    return {origin[0] + double(c[0]) * d[0], origin[1] + double(c[1]) * d[1], origin[2] + double(c[2]) * d[2]};
  }
  int64_t numNodes() const { return int64_t(N[0] + 1) * (N[1] + 1) * (N[2] + 1); }
};

enum FieldKind { FIELD_B, FIELD_E, FIELD_D, FIELD_PSI };

// MxGridField.{hpp,cpp} + the Yee specialisations. One guard cell for shape fractions
// (MxEMSim.cpp:122-129: addShapeRep(*pec, "pec", 1, true)).
struct Field {
  const Grid* g;
  FieldKind kind;
  int ncomp;
  D3 xi[3];         // cell-relative component positions
  int compDir[3];   // vector direction of each component
  BCType lbc[3][3], ubc[3][3];
  cplx phase[3];
  const Field* bfield = nullptr;  // E and psi consult the B map (Dey-Mittra)
  double dmFrac = 0.0;            // MxYeeFitBField.h: mDMFrac
  bool regionSet = false;
  std::string regionName;
  std::map<std::string, std::vector<double>> reps;  // shape fractions on the guarded block
  std::vector<int64_t> gids;                        // the map (ascending on one rank)
  std::vector<int32_t> lidOf;                       // dense GID -> LID (-1 if absent)
  // When true, reproduce MxYeeFitEField::getCompFactor literally: an E component whose
  // un-wrapped cell sits on a PERIODIC upper boundary gets factor 0.
  bool literalUpperPeriodicE = false;

  Field(const Grid* grid, FieldKind k, const Field* b = nullptr) : g(grid), kind(k), bfield(b) {
    const double dx = g->d[0], dy = g->d[1], dz = g->d[2];
    for (int c = 0; c < 3; ++c) { xi[c] = {0, 0, 0}; compDir[c] = c; }
    if (k == FIELD_E || k == FIELD_D) {          // MxYeeElecFieldBase.cpp:64-82
      ncomp = 3;
      xi[0][0] = 0.5 * dx; xi[1][1] = 0.5 * dy; xi[2][2] = 0.5 * dz;
    } else if (k == FIELD_B) {                   // MxYeeMagFieldBase.cpp:85-116
      ncomp = 3;
      xi[0][1] = 0.5 * dy; xi[0][2] = 0.5 * dz;
      xi[1][2] = 0.5 * dz; xi[1][0] = 0.5 * dx;
      xi[2][0] = 0.5 * dx; xi[2][1] = 0.5 * dy;
    } else {                                     // MxYeePsiField.cpp:51-72
      ncomp = 1;
      xi[0] = {0.5 * dx, 0.5 * dy, 0.5 * dz};
    }
    for (int c = 0; c < 3; ++c)
      for (int i = 0; i < 3; ++i) { lbc[c][i] = PERIODIC; ubc[c][i] = PERIODIC; }
    for (int i = 0; i < 3; ++i) phase[i] = cplx(1.0, 0.0);
  }

  // MxYeeElecFieldBase.cpp:89-138, MxYeeMagFieldBase.cpp:119-168, MxYeePsiField.cpp:74-97
  void setBCs(const BCType lower[3], const BCType upper[3]) {
    for (int c = 0; c < ncomp; ++c)
      for (int i = 0; i < 3; ++i) {
        lbc[c][i] = translate(lower[i], c, i);
        ubc[c][i] = translate(upper[i], c, i);
      }
  }
  BCType translate(BCType bc, int comp, int dir) const {
    if (bc != PEC && bc != PMC) return bc;
    const bool normal = (compDir[comp] == dir);
    if (kind == FIELD_PSI) return bc == PEC ? CONSTANT : ZERO;
    if (kind == FIELD_B) return bc == PEC ? (normal ? ZERO : CONSTANT) : (normal ? CONSTANT : ZERO);
    return bc == PEC ? (normal ? CONSTANT : ZERO) : (normal ? ZERO : CONSTANT);
  }
  // MxGridField.cpp:34-38
  void setPhaseShifts(const double ph[3]) {
    for (int i = 0; i < 3; ++i) phase[i] = std::exp(cplx(0.0, 1.0) * ph[i]);
  }

  // guarded block: cells -1 .. N+1 in each direction (MxGrid.cpp:171-193 on one rank,
  // MxGridDomain.cpp:60-89)
  int64_t fullIndex(const I3& c) const {
    int64_t res = 0, factor = 1;
    for (int i = 2; i >= 0; --i) {
      const int lo = -1, hi = g->N[i] + 2;
      if (c[i] < lo || c[i] >= hi) throw std::runtime_error("mxo: cell outside guarded block");
      res += int64_t(c[i] - lo) * factor;
      factor *= (g->N[i] + 3);
    }
    return res;
  }
  int64_t numFullCells() const { return int64_t(g->N[0] + 3) * (g->N[1] + 3) * (g->N[2] + 3); }

  // MxGridField.hpp:259-279
  double frac(int comp, const I3& cell, const std::string& name) const {
    auto it = reps.find(name);
    if (it == reps.end()) return 1.0;
    return it->second[comp + ncomp * fullIndex(cell)];
  }
  double regionFrac(int comp, const I3& cell) const { return regionVec ? (*regionVec)[comp + ncomp * fullIndex(cell)] : 1.0; }
  const std::vector<double>* regionVec = nullptr;

  // MxGridField.cpp:80-142 (note: "case 1" there is always overwritten by "case 2")
  I3 interior(int comp, const I3& cell) const {
    I3 nc = cell;
    for (int i = 0; i < 3; ++i) {
      const int n = g->N[i];
      const bool onLower = (xi[comp][i] == 0.0);
      if (cell[i] < 0) {
        nc[i] = (lbc[comp][i] == PERIODIC) ? n + cell[i] : -cell[i] - 1;
      } else if (cell[i] == n && onLower) {
        if (ubc[comp][i] == PERIODIC) nc[i] = 0;
      } else if (cell[i] >= n && onLower) {
        nc[i] = (ubc[comp][i] == PERIODIC) ? cell[i] - n : n - (cell[i] - n);
      } else if (cell[i] >= n) {
        nc[i] = (ubc[comp][i] == PERIODIC) ? cell[i] - n : n - (cell[i] - n + 1);
      }
    }
    return nc;
  }
  // MxGridField.hpp:142-145
  int64_t gid(int comp, const I3& cell) const { return comp + ncomp * g->cellToGlobal(interior(comp, cell)); }

  // MxGridField.cpp:41-77
  bool baseUse(int comp, const I3& cell) const {
    if (regionSet && regionFrac(comp, cell) == 0.0) return false;
    for (int i = 0; i < 3; ++i) {
      const int n = g->N[i];
      const double x = xi[comp][i];
      if (cell[i] == 0 && x == 0.0) {
        if (lbc[comp][i] == ZERO) return false;
      } else if (cell[i] == n && x > 0.0) {
        return false;
      } else if (cell[i] == n && x == 0.0) {
        if (ubc[comp][i] == ZERO || ubc[comp][i] == PERIODIC) return false;
      }
    }
    return true;
  }
  // MxYeeFitBField.cpp:63-80, MxYeeFitEField.cpp:53-88, MxYeePsiField.cpp:100-114
  bool use(int comp, I3 cell) const {
    switch (kind) {
      case FIELD_B: {
        for (int c = 0; c < ncomp; ++c)
          if (baseUse(c, cell)) return true;
        bool res = baseUse(comp, cell);
        if (regionSet && regionFrac(comp, cell) < dmFrac) res = false;
        return res;
      }
      case FIELD_E: {
        if (!baseUse(comp, cell)) return false;
        const int c2 = (comp + 1) % 3, c3 = (comp + 2) % 3;
        if (!bfield->use(c2, cell)) return false;
        cell[c3]--;
        if (!bfield->use(c2, cell)) return false;
        cell[c3]++;
        if (!bfield->use(c3, cell)) return false;
        cell[c2]--;
        if (!bfield->use(c3, cell)) return false;
        return true;
      }
      case FIELD_PSI: {
        if (!baseUse(comp, cell)) return false;
        for (int i = 0; i < 3; ++i) {
          if (bfield->use(i, cell)) return true;
          cell[i]++;
          if (bfield->use(i, cell)) return true;
          cell[i]--;
        }
        return false;
      }
      default:
        return baseUse(comp, cell);
    }
  }

  // MxGridField.cpp:145-190
  cplx baseFactor(int comp, const I3& cell) const {
    if (regionSet && regionFrac(comp, cell) == 0.0) return 0.0;
    cplx res(1.0, 0.0);
    for (int i = 0; i < 3; ++i) {
      const int n = g->N[i];
      const double x = xi[comp][i];
      const BCType lo = lbc[comp][i], up = ubc[comp][i];
      if (cell[i] < 0) {
        if (lo == PERIODIC) res /= phase[i];
        else if (lo == ZERO) res *= -1.0;
      } else if (cell[i] == 0 && x == 0.0) {
        if (lo == ZERO) res *= 0.0;
      } else if (cell[i] == n && x == 0.0) {
        if (up == PERIODIC) res *= phase[i];
        else if (up == ZERO) res *= 0.0;
      } else if (cell[i] >= n) {
        if (up == PERIODIC) res *= phase[i];
        else if (up == ZERO) res *= -1.0;
      }
    }
    return res;
  }
  // MxYeeFitBField.cpp:82-90, MxYeeFitEField.cpp:90-98, MxYeePsiField.cpp:116-124
  cplx factor(int comp, const I3& cell) const {
    switch (kind) {
      case FIELD_B:
        if (regionSet && regionFrac(comp, cell) < dmFrac) return 0.0;
        return baseFactor(comp, cell);
      case FIELD_E:
      case FIELD_PSI: {
        const cplx res = baseFactor(comp, cell);
        // The reference tests useCompInMap on the *un-wrapped* cell, which is false on a
        // PERIODIC upper boundary; by default we test the wrapped component instead so the
        // wrap-around entry survives. literalUpperPeriodicE restores the
        // as-written behaviour.
        const I3 c = literalUpperPeriodicE ? cell : periodicWrap(comp, cell);
        if (!use(comp, c)) return 0.0;
        return res;
      }
      default:
        return baseFactor(comp, cell);
    }
  }
  // wrap only across PERIODIC upper boundaries; everything else is left as the caller gave it
  I3 periodicWrap(int comp, const I3& cell) const {
    I3 c = cell;
    for (int i = 0; i < 3; ++i)
      if (c[i] >= g->N[i] && ubc[comp][i] == PERIODIC && lbc[comp][i] == PERIODIC) c[i] -= g->N[i];
    return c;
  }

  // MxGridField.cpp:193-226 with calcCompFrac (MxGridField.hpp:186-191): fractions for
  // every component of every cell in the guarded block.
  void addShapeRep(const Shape& sh, const std::string& name, bool setRegionFlag);
  // MxGridField.cpp:256-297: interior cells x slow .. z fast, components inner.
  void setMap();
  void shareMap(const Field& partner) { gids = partner.gids; lidOf = partner.lidOf; }
  int32_t lid(int64_t gidv) const { return lidOf[gidv]; }
  int64_t numGlobal() const { return int64_t(ncomp) * g->numNodes(); }
};

inline void Field::addShapeRep(const Shape& sh, const std::string& name, bool setRegionFlag) {
  const int64_t nfull = numFullCells();
  std::vector<double> rep(size_t(ncomp) * nfull);
  const int n1 = g->N[1] + 3, n2 = g->N[2] + 3;
  const double dx = g->d[0], dy = g->d[1], dz = g->d[2];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nfull; ++i) {
    I3 cell;
    cell[2] = int(i % n2) - 1;
    cell[1] = int((i / n2) % n1) - 1;
    cell[0] = int(i / (int64_t(n2) * n1)) - 1;
    for (int comp = 0; comp < ncomp; ++comp) {
      const I3 nc = interior(comp, cell);
      const D3 p = g->nodeCoord(nc) + xi[comp];
      double f;
      switch (kind) {
        case FIELD_E:  // MxYeeFitEField.cpp:40-48: edges
          f = segmentFraction(sh, comp, g->d[comp], p);
          break;
        case FIELD_B:  // MxYeeFitBField.cpp:52-58: faces (x: dy,dz; y: dz,dx; z: dx,dy)
        case FIELD_D:  // MxYeeFitDField.cpp:53-58: dual faces, same rectangles
          f = comp == 0 ? rectFraction(sh, 0, dy, dz, p)
            : comp == 1 ? rectFraction(sh, 1, dz, dx, p)
                        : rectFraction(sh, 2, dx, dy, p);
          break;
        default:       // MxYeePsiField.cpp:64: cell box
          f = boxFraction(sh, dx, dy, dz, p);
      }
      rep[comp + ncomp * i] = f;
    }
  }
  reps[name] = std::move(rep);
  if (setRegionFlag) {
    regionName = name;
    regionSet = true;
    regionVec = &reps[name];
    setMap();
  }
}

inline void Field::setMap() {
  const int nx = g->N[0] + 1, ny = g->N[1] + 1, nz = g->N[2] + 1;
  const int64_t ncell = int64_t(nx) * ny * nz;
  std::vector<uint8_t> flag(size_t(ncell) * ncomp);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < ncell; ++i) {
    I3 cell{int(i / (int64_t(ny) * nz)), int((i / nz) % ny), int(i % nz)};
    for (int comp = 0; comp < ncomp; ++comp) flag[i * ncomp + comp] = use(comp, cell) ? 1 : 0;
  }
  gids.clear();
  lidOf.assign(size_t(ncell) * ncomp, -1);
  for (int64_t i = 0; i < ncell * ncomp; ++i)
    if (flag[i]) {
      lidOf[i] = int32_t(gids.size());
      gids.push_back(i);  // comp + ncomp * cellToGlobal(cell) == i for in-range cells
    }
}

}  // namespace mxo
