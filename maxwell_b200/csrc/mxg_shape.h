// Implicit-function solids and the cut-cell fractions of the Dey-Mittra fields: edge length, face area
// and cell volume fractions inside a CSG shape. Like mxg_yee.h this is host/device code: the fraction kernels of
// mxg_asm.cu run one thread per cell of the guarded block, the CPU replay of the tests runs the same functions in a loop.
//
// A shape is a flat array of nodes (root = node 0; children linked by firstChild / nextSibling). f > 0 inside;
// func(p) = sign * f0(Ainv p - Ainv b) with the affine placement x -> A x + b (MxShape.hpp:143-165, MxShape.cpp:89-210).
// Reference: MxCylinder.hpp:34-84, MxHalfSpace.hpp:31-33, MxSphere.hpp:31-33, MxEllipsoid.hpp:31-50, MxTorus.hpp:31-65,
// MxCone.hpp:31-62, MxSlab.hpp:33-100, MxShapeIntersection.hpp:120-206, MxShapeUnion.hpp:62-154,
// MxShapeSubtract.hpp:69-176, MxShapeMirror.hpp:85-165, MxShapeRepeat.hpp:84-148 (solids); MxUtil.hpp:295-362 (root
// finder); MxSegment.cpp:22-43, MxCartRect.hpp:34-88, MxConvexPolygon.cpp:25-292, MxPolytope.cpp:7-97, MxCartBox.cpp:21-62,
// MxConvexPolyhedron.cpp:130-344 (fractions); MxGridField.cpp:193-226 (addShapeRep).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "mxg_yee.h"

#if defined(__CUDACC__)
#define MXS_HD __host__ __device__
#define MXS_NOINLINE __noinline__
#else
#define MXS_HD
#define MXS_NOINLINE __attribute__((noinline))
#endif

namespace mxa {

enum ShapeType {
  SHAPE_CYLINDER = 0, SHAPE_HALFSPACE = 1, SHAPE_SPHERE = 2, SHAPE_ELLIPSOID = 3, SHAPE_TORUS = 4, SHAPE_CONE = 5,
  SHAPE_INTERSECTION = 6, SHAPE_UNION = 7, SHAPE_SUBTRACT = 8, SHAPE_MIRROR = 9, SHAPE_REPEAT = 10
};
constexpr int kShapeMaxDepth = 8;

struct ShapeNode {
  int32_t type, firstChild, nextSibling, pad;
  double sign;
  double Ainv[9], Ainvb[3];
  double par[16];
  double A[9], b[3];      // host-side bookkeeping of the placement
};

struct V3 {
  double v[3];
};
MXS_HD inline V3 vAdd(V3 a, V3 b) { return {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; }
MXS_HD inline V3 vSub(V3 a, V3 b) { return {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; }
MXS_HD inline V3 vScale(double s, V3 a) { return {{s * a.v[0], s * a.v[1], s * a.v[2]}}; }
MXS_HD inline V3 vTimes(V3 a, double s) { return {{a.v[0] * s, a.v[1] * s, a.v[2] * s}}; }
MXS_HD inline V3 vDiv(V3 a, double s) { return {{a.v[0] / s, a.v[1] / s, a.v[2] / s}}; }
MXS_HD inline double vDot(V3 a, V3 b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
MXS_HD inline double vNorm(V3 a) { return ::sqrt(vDot(a, a)); }
MXS_HD inline V3 vCross(V3 a, V3 b) {
  return {{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}};
}
MXS_HD inline V3 mTimes(const double* m, V3 a) {      // M v, M row-major 3x3
  V3 r;
  for (int i = 0; i < 3; ++i) r.v[i] = m[3 * i] * a.v[0] + m[3 * i + 1] * a.v[1] + m[3 * i + 2] * a.v[2];
  return r;
}
MXS_HD inline V3 rowTimes(V3 a, const double* m) {    // v^T M
  V3 r;
  for (int j = 0; j < 3; ++j) r.v[j] = a.v[0] * m[j] + a.v[1] * m[3 + j] + a.v[2] * m[6 + j];
  return r;
}
MXS_HD inline int sgn(double v) { return v < 0.0 ? -1 : (v > 0.0 ? 1 : 0); }   // MxUtil.hpp:26,44-49 with dEps = 0

MXS_HD inline V3 cylPp(const double* P, V3 p) {       // accumulates from zero like the reference's loop
  V3 r;
  for (int i = 0; i < 3; ++i) {
    double s = 0;
    for (int j = 0; j < 3; ++j) s += P[3 * i + j] * p.v[j];
    r.v[i] = s;
  }
  return r;
}
MXS_HD inline V3 repeatFold(const ShapeNode& n, V3 p) {
  const V3 o = {{n.par[0], n.par[1], n.par[2]}}, dir = {{n.par[3], n.par[4], n.par[5]}};
  const double s = n.par[6], np = n.par[7], nn = n.par[8];
  const double slabPt = vDot(vSub(p, o), dir) / s + 0.5;
  const double fl = ::floor(slabPt);
  if (slabPt >= 0.0) return vSub(p, vScale(s * (fl < np ? fl : np), dir));
  return vSub(p, vScale(s * (fl > nn ? fl : nn), dir));
}
MXS_HD inline double mirrorPlane(const ShapeNode& n, V3 p) {   // the half-space of MxShapeMirror: n . (p - point)
  const V3 nrm = {{n.par[0], n.par[1], n.par[2]}}, pt = {{n.par[3], n.par[4], n.par[5]}};
  return vDot(nrm, vSub(p, pt));
}

// Depth-bounded static recursion: the call graph is acyclic, so the device stack frame is known at compile time.
template <int D>
struct ShapeEval {
  static MXS_HD MXS_NOINLINE double func(const ShapeNode* nodes, int idx, V3 p) {
    const ShapeNode& n = nodes[idx];
    const V3 q = vSub(mTimes(n.Ainv, p), V3{{n.Ainvb[0], n.Ainvb[1], n.Ainvb[2]}});
    return n.sign * f0(nodes, n, q);
  }
  static MXS_HD MXS_NOINLINE V3 grad(const ShapeNode* nodes, int idx, V3 p) {
    const ShapeNode& n = nodes[idx];
    const V3 q = vSub(mTimes(n.Ainv, p), V3{{n.Ainvb[0], n.Ainvb[1], n.Ainvb[2]}});
    return vScale(n.sign, rowTimes(g0(nodes, n, q), n.Ainv));
  }
  static MXS_HD double childFunc(const ShapeNode* nodes, int c, V3 q) { return ShapeEval<D - 1>::func(nodes, c, q); }
  static MXS_HD V3 childGrad(const ShapeNode* nodes, int c, V3 q) { return ShapeEval<D - 1>::grad(nodes, c, q); }

  // the (inverted) union of removal shapes of MxShapeSubtract: children after the first
  static MXS_HD double removedFunc(const ShapeNode* nodes, const ShapeNode& n, V3 q, int* arg) {
    double fmax = -1.7976931348623157e308;
    int best = -1;
    for (int c = nodes[n.firstChild].nextSibling; c >= 0; c = nodes[c].nextSibling) {
      const double f = childFunc(nodes, c, q);
      if (f > fmax) { fmax = f; best = c; }
    }
    if (arg) *arg = best;
    return -1.0 * fmax;
  }

  static MXS_HD double f0(const ShapeNode* nodes, const ShapeNode& n, V3 p) {
    switch (n.type) {
      case SHAPE_CYLINDER: return n.par[0] - vDot(p, cylPp(n.par + 1, p));
      case SHAPE_HALFSPACE: return vDot(V3{{n.par[0], n.par[1], n.par[2]}}, p);
      case SHAPE_SPHERE: return 1.0 - vDot(p, p) / n.par[0];
      case SHAPE_ELLIPSOID:
        return 1.0 - (p.v[0] * (n.par[0] * p.v[0]) + p.v[1] * (n.par[1] * p.v[1]) + p.v[2] * (n.par[2] * p.v[2]));
      case SHAPE_TORUS: {
        const V3 axis = {{n.par[0], n.par[1], n.par[2]}};
        const double d = n.par[3] - vNorm(mTimes(n.par + 5, p)), h = vDot(axis, p);
        return n.par[4] - d * d - h * h;
      }
      case SHAPE_CONE: {
        const V3 axis = {{n.par[0], n.par[1], n.par[2]}};
        const double t = n.par[3] * vDot(axis, p);
        return t * t - vDot(p, mTimes(n.par + 4, p));
      }
      case SHAPE_INTERSECTION: {
        double fmin = 1.7976931348623157e308;
        for (int c = n.firstChild; c >= 0; c = nodes[c].nextSibling) {
          const double f = childFunc(nodes, c, p);
          if (f < fmin) fmin = f;
        }
        return fmin;
      }
      case SHAPE_UNION: {
        double fmax = -1.7976931348623157e308;
        for (int c = n.firstChild; c >= 0; c = nodes[c].nextSibling) {
          const double f = childFunc(nodes, c, p);
          if (f > fmax) fmax = f;
        }
        return fmax;
      }
      case SHAPE_SUBTRACT: {
        const double fBase = childFunc(nodes, n.firstChild, p), fRm = removedFunc(nodes, n, p, nullptr);
        const bool inBase = fBase > 0, inRm = fRm < 0;
        if (inRm && inBase) return fRm;
        if (!inRm && !inBase) return fBase;
        return fBase < fRm ? fBase : fRm;
      }
      case SHAPE_MIRROR: {
        const int c = mirrorPlane(n, p) > 0.0 ? n.firstChild : nodes[n.firstChild].nextSibling;
        return childFunc(nodes, c, p);
      }
      case SHAPE_REPEAT: return childFunc(nodes, n.firstChild, repeatFold(n, p));
    }
    return 0.0;
  }

  static MXS_HD V3 g0(const ShapeNode* nodes, const ShapeNode& n, V3 p) {
    switch (n.type) {
      case SHAPE_CYLINDER: return vScale(-2.0, cylPp(n.par + 1, p));
      case SHAPE_HALFSPACE: return V3{{n.par[0], n.par[1], n.par[2]}};
      case SHAPE_SPHERE: return vDiv(vScale(-2.0, p), n.par[0]);
      case SHAPE_ELLIPSOID: return V3{{-2.0 * n.par[0] * p.v[0], -2.0 * n.par[1] * p.v[1], -2.0 * n.par[2] * p.v[2]}};
      case SHAPE_TORUS: {   // restated literally: the in-plane term has the opposite sign of the true derivative (DESIGN R14)
        const V3 axis = {{n.par[0], n.par[1], n.par[2]}};
        const double c = 1.0 - n.par[3] / vNorm(mTimes(n.par + 5, p));
        double cP[9];
        for (int i = 0; i < 9; ++i) cP[i] = c * n.par[5 + i];
        return vScale(2.0, vSub(mTimes(cP, p), vScale(vDot(axis, p), axis)));
      }
      case SHAPE_CONE: {
        const V3 axis = {{n.par[0], n.par[1], n.par[2]}};
        return vScale(2.0, vSub(vScale(n.par[3] * n.par[3] * vDot(axis, p), axis), mTimes(n.par + 4, p)));
      }
      case SHAPE_INTERSECTION: {
        double fmin = 1.7976931348623157e308;
        int arg = n.firstChild;
        for (int c = n.firstChild; c >= 0; c = nodes[c].nextSibling) {
          const double f = childFunc(nodes, c, p);
          if (f < fmin) { fmin = f; arg = c; }
        }
        return childGrad(nodes, arg, p);
      }
      case SHAPE_UNION: {
        double fmax = -1.7976931348623157e308;
        int arg = n.firstChild;
        for (int c = n.firstChild; c >= 0; c = nodes[c].nextSibling) {
          const double f = childFunc(nodes, c, p);
          if (f > fmax) { fmax = f; arg = c; }
        }
        return childGrad(nodes, arg, p);
      }
      case SHAPE_SUBTRACT: {
        int arg = -1;
        const double fBase = childFunc(nodes, n.firstChild, p), fRm = removedFunc(nodes, n, p, &arg);
        const bool inBase = fBase > 0, inRm = fRm < 0;
        const bool useRm = (inRm && inBase) ? true : ((!inRm && !inBase) ? false : !(fBase < fRm));
        if (!useRm || arg < 0) return childGrad(nodes, n.firstChild, p);
        return vScale(-1.0, childGrad(nodes, arg, p));
      }
      case SHAPE_MIRROR: {
        const int c = mirrorPlane(n, p) > 0.0 ? n.firstChild : nodes[n.firstChild].nextSibling;
        return childGrad(nodes, c, p);
      }
      case SHAPE_REPEAT: return childGrad(nodes, n.firstChild, repeatFold(n, p));
    }
    return V3{{0, 0, 0}};
  }
};
template <>
struct ShapeEval<0> {    // deeper than kShapeMaxDepth: rejected when the shape is imported
  static MXS_HD double func(const ShapeNode*, int, V3) { return 0.0; }
  static MXS_HD V3 grad(const ShapeNode*, int, V3) { return V3{{0, 0, 0}}; }
};

MXS_HD inline double shapeFunc(const ShapeNode* nodes, V3 p) { return ShapeEval<kShapeMaxDepth>::func(nodes, 0, p); }
MXS_HD inline V3 shapeGrad(const ShapeNode* nodes, V3 p) { return ShapeEval<kShapeMaxDepth>::grad(nodes, 0, p); }

// Safeguarded Newton / bisection on the segment p1 -> p2 (MxUtil.hpp:295-362; 500 iterations whatever the caller asks,
// tolerance tol * |p2 - p1|).
MXS_HD inline V3 rootFind(const ShapeNode* sh, V3 p1, V3 p2, double tol) {
  const int maxiter = 500;
  const double len = vNorm(vSub(p2, p1));
  const double stol = tol * len;
  const V3 dir = vDiv(vSub(p2, p1), len);
  const double f1 = shapeFunc(sh, p1), f2 = shapeFunc(sh, p2);
  if (f1 == 0) return p1;
  if (f2 == 0) return p2;
  double lo, hi;   // lo: f < 0 side, hi: f >= 0 side
  if (f1 < 0) { lo = 0.0; hi = len; } else { lo = len; hi = 0.0; }
  double t = 0.5 * (lo + hi);
  double stepPrev = ::fabs(hi - lo), step = stepPrev;
  double f = shapeFunc(sh, vAdd(p1, vTimes(dir, t)));
  double df = vDot(dir, shapeGrad(sh, vAdd(p1, vTimes(dir, t))));
  for (int it = 0; it < maxiter; ++it) {
    const bool outOfBracket = ((t - hi) * df - f) * ((t - lo) * df - f) >= 0;
    const bool slow = ::fabs(2.0 * f) > ::fabs(stepPrev * df);
    if (outOfBracket || slow) {
      stepPrev = step;
      step = 0.5 * (hi - lo);
      t = lo + step;
      if (lo == t) return vAdd(p1, vTimes(dir, t));
    } else {
      stepPrev = step;
      step = f / df;
      const double told = t;
      t -= step;
      if (told == t) return vAdd(p1, vTimes(dir, t));
    }
    if (::fabs(step) < stol) return vAdd(p1, vTimes(dir, t));
    f = shapeFunc(sh, vAdd(p1, vTimes(dir, t)));
    df = vDot(dir, shapeGrad(sh, vAdd(p1, vTimes(dir, t))));
    if (f < 0) lo = t; else hi = t;
  }
  return vAdd(p1, vTimes(dir, t));
}

// MxSegment.cpp:22-43 with MxCartSeg.hpp: axis-aligned edge of length len about `mid`
MXS_HD inline double segmentFraction(const ShapeNode* sh, int axis, double len, V3 mid) {
  V3 d = {{0, 0, 0}};
  d.v[axis] = 1.0;
  const V3 p1 = vAdd(mid, vScale(0.5 * len, d));
  const V3 p2 = vSub(mid, vScale(0.5 * len, d));
  const double f1 = shapeFunc(sh, p1), f2 = shapeFunc(sh, p2);
  const int s1 = sgn(f1), s2 = sgn(f2);
  if ((s1 == 1 && s2 != -1) || (s2 == 1 && s1 != -1)) return 1;
  if ((s1 == -1 && s2 != 1) || (s2 == -1 && s1 != 1)) return 0;
  if (s1 == 0 && s2 == 0) return sgn(shapeFunc(sh, mid)) == 1 ? 1 : 0;
  const V3 p = rootFind(sh, p1, p2, 1.e-12);
  return (s1 == 1 ? vNorm(vSub(p, p1)) : vNorm(vSub(p, p2))) / len;
}

// MxCartRect.hpp:34-88 + MxConvexPolygon.cpp:25-76,166-292 + MxPolytope.cpp:7-97: fan of triangles from the first edge
// crossing. type 0 = x-face (d1 = y, d2 = z), 1 = y-face (z, x), 2 = z-face (x, y).
MXS_HD inline double rectFraction(const ShapeNode* sh, int type, double l1, double l2, V3 p) {
  V3 d1 = {{0, 0, 0}}, d2 = {{0, 0, 0}};
  d1.v[(type + 1) % 3] = 1.0;
  d2.v[(type + 2) % 3] = 1.0;
  V3 v[4];
  v[0] = vAdd(p, vScale(0.5, vSub(vScale(-l1, d1), vScale(l2, d2))));
  v[1] = vAdd(p, vScale(0.5, vSub(vScale(l1, d1), vScale(l2, d2))));
  v[2] = vAdd(p, vScale(0.5, vAdd(vScale(-l1, d1), vScale(l2, d2))));
  v[3] = vAdd(p, vScale(0.5, vAdd(vScale(l1, d1), vScale(l2, d2))));
  const int E[4][2] = {{0, 1}, {0, 2}, {1, 3}, {2, 3}};
  double fv[4];
  bool noneOut = true, noneIn = true, allOn = true;
  for (int i = 0; i < 4; ++i) {
    fv[i] = shapeFunc(sh, v[i]);
    const int s = sgn(fv[i]);
    if (s == -1) { noneOut = false; allOn = false; }
    else if (s == 1) { noneIn = false; allOn = false; }
  }
  if (allOn) return sgn(shapeFunc(sh, p)) == 1 ? 1 : 0;
  if (noneOut) return 1;
  if (noneIn) return 0;
  bool has[4] = {false, false, false, false};
  V3 ex[4];
  int first = -1;
  for (int e = 0; e < 4; ++e) {
    const int a = E[e][0], c = E[e][1];
    const int sa = sgn(fv[a]), sc = sgn(fv[c]);
    if ((sa == -1 && sc == 1) || (sa == 1 && sc == -1)) { ex[e] = rootFind(sh, v[a], v[c], 1.e-12); has[e] = true; }
    else if (sa == 0 && sc == 1) { ex[e] = v[a]; has[e] = true; }
    else if (sc == 0 && sa == 1) { ex[e] = v[c]; has[e] = true; }
    if (has[e] && first < 0) first = e;
  }
  if (first < 0) return 0;   // cannot happen for a sign change; keeps the index in range
  const V3 v0 = ex[first];
  double area = 0.0;
  for (int e = 0; e < 4; ++e) {
    const int a = E[e][0], c = E[e][1];
    const int sa = sgn(fv[a]), sc = sgn(fv[c]);
    V3 q1, q2;
    if (has[e]) { q1 = sa != -1 ? v[a] : ex[e]; q2 = sc != -1 ? v[c] : ex[e]; }
    else if (sa == 1 || sc == 1) { q1 = v[a]; q2 = v[c]; }
    else continue;
    area += 0.5 * vNorm(vCross(vSub(q1, v0), vSub(q2, v0)));
  }
  return area / (l1 * l2);
}

// MxCartBox.cpp:21-62 (connectivity) + MxConvexPolyhedron.cpp:130-344: pyramids over the inside faces from every cut-edge
// point, weighted by the cut lengths of the adjacent faces.
MXS_HD inline double boxFraction(const ShapeNode* sh, double lx, double ly, double lz, V3 p) {
  V3 v[8];
  for (int i = 0; i < 8; ++i) {
    v[i] = p;
    v[i].v[0] += (i & 4 ? 0.5 : -0.5) * lx;
    v[i].v[1] += (i & 2 ? 0.5 : -0.5) * ly;
    v[i].v[2] += (i & 1 ? 0.5 : -0.5) * lz;
  }
  const int EV[12][2] = {{0, 4}, {1, 5}, {2, 6}, {3, 7}, {0, 2}, {1, 3}, {4, 6}, {5, 7}, {0, 1}, {2, 3}, {4, 5}, {6, 7}};
  const int EF[12][2] = {{0, 2}, {1, 2}, {0, 3}, {1, 3}, {0, 4}, {1, 4}, {0, 5}, {1, 5}, {2, 4}, {3, 4}, {2, 5}, {3, 5}};
  // face -> edges and face -> vertices, both ascending (derived from EF / EV)
  const int FE[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 8, 10}, {2, 3, 9, 11}, {4, 5, 8, 9}, {6, 7, 10, 11}};
  const int FV[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
  double fv[8];
  bool noneOut = true, noneIn = true, allOn = true;
  for (int i = 0; i < 8; ++i) {
    fv[i] = shapeFunc(sh, v[i]);
    const int s = sgn(fv[i]);
    if (s == -1) { noneOut = false; allOn = false; }
    else if (s == 1) { noneIn = false; allOn = false; }
  }
  if (allOn) return sgn(shapeFunc(sh, p)) == 1 ? 1 : 0;
  if (noneOut) return 1;
  if (noneIn) return 0;
  bool has[12];
  V3 ex[12];
  for (int e = 0; e < 12; ++e) {
    has[e] = false;
    const int a = EV[e][0], c = EV[e][1];
    const int sa = sgn(fv[a]), sc = sgn(fv[c]);
    if ((sa == -1 && sc == 1) || (sa == 1 && sc == -1)) { ex[e] = rootFind(sh, v[a], v[c], 1.e-12); has[e] = true; }
    else if (sa == 0 && sc == 1) { ex[e] = v[a]; has[e] = true; }
    else if (sc == 0 && sa == 1) { ex[e] = v[c]; has[e] = true; }
  }
  bool faceUsed[6];
  V3 faceArea[6], faceVert[6];
  double cutLen[6];
  for (int f = 0; f < 6; ++f) {
    faceUsed[f] = false;
    cutLen[f] = 0.0;
    bool inside = false;
    for (int k = 0; k < 4; ++k)
      if (sgn(fv[FV[f][k]]) == 1) { inside = true; break; }
    if (!inside) continue;
    int e0 = -1;
    for (int k = 0; k < 4; ++k)
      if (has[FE[f][k]]) { e0 = FE[f][k]; break; }
    V3 av = {{0, 0, 0}};
    faceUsed[f] = true;
    if (e0 >= 0) {
      const V3 v0 = ex[e0];
      faceVert[f] = v0;
      for (int k = 0; k < 4; ++k) {
        const int e = FE[f][k];
        const int a = EV[e][0], c = EV[e][1];
        const int sa = sgn(fv[a]), sc = sgn(fv[c]);
        V3 q1, q2;
        if (has[e] && e != e0) {
          q1 = ex[e];
          q2 = sa != -1 ? v[a] : v[c];
          cutLen[f] = vNorm(vSub(v0, q1));
        } else if (sa == 1 || sc == 1) { q1 = v[a]; q2 = v[c]; }
        else continue;
        const V3 tri = vCross(vSub(q1, v0), vSub(q2, v0));
        av = vDot(av, tri) > 0 ? vAdd(av, tri) : vSub(av, tri);
      }
    } else {
      const int base = FV[f][0];
      const V3 v0 = v[base];
      faceVert[f] = v0;
      for (int k = 0; k < 4; ++k) {
        const int e = FE[f][k];
        const int a = EV[e][0], c = EV[e][1];
        if (a == base || c == base) continue;
        const V3 tri = vCross(vSub(v[a], v0), vSub(v[c], v0));
        av = vDot(av, tri) > 0 ? vAdd(av, tri) : vSub(av, tri);
      }
    }
    faceArea[f] = vScale(0.5, av);
  }
  double vol = 0, wtSum = 0;
  for (int e = 0; e < 12; ++e) {
    if (!has[e]) continue;
    const double wt = cutLen[EF[e][0]] + cutLen[EF[e][1]];
    wtSum += wt;
    for (int f = 0; f < 6; ++f)
      if (faceUsed[f]) vol += wt * ::fabs(vDot(vSub(faceVert[f], ex[e]), faceArea[f])) / 3.;
  }
  vol /= wtSum;
  return vol / (lx * ly * lz);
}

// MxGridField.cpp:193-226 with calcCompFrac (MxGridField.hpp:186-191): fractions of every component of one cell of the
// guarded block. E: edges (MxYeeFitEField.cpp:40-48), B: faces (MxYeeFitBField.cpp:52-58), D: the dual faces at the E
// positions (MxYeeFitDField.cpp:53-58), psi: the cell box (MxYeePsiField.cpp:64).
struct FractionCells {
  const mxy::Sim* sim;
  const ShapeNode* shape;
  int kind;
  double* out;
  MXS_HD void operator()(int64_t i) const {
    const mxy::Sim& s = *sim;
    const mxy::Field& f = s.f[kind];
    const mxy::Grid& g = s.g;
    int cell[3];
    mxy::fullToCell(g, i, cell);
    for (int comp = 0; comp < f.ncomp; ++comp) {
      int nc[3];
      mxy::interior(g, f, comp, cell, nc);
      V3 p;
      for (int k = 0; k < 3; ++k) p.v[k] = (g.origin[k] + double(nc[k]) * g.d[k]) + f.xi[comp][k];
      double fr;
      if (kind == mxy::FIELD_E) fr = segmentFraction(shape, comp, g.d[comp], p);
      else if (kind == mxy::FIELD_B || kind == mxy::FIELD_D)
        fr = comp == 0 ? rectFraction(shape, 0, g.d[1], g.d[2], p)
           : comp == 1 ? rectFraction(shape, 1, g.d[2], g.d[0], p)
                       : rectFraction(shape, 2, g.d[0], g.d[1], p);
      else fr = boxFraction(shape, g.d[0], g.d[1], g.d[2], p);
      out[comp + f.ncomp * i] = fr;
    }
  }
};

// ---- host side: building shapes --------------------------------------------------------------------------------
inline void m33Mul(const double* x, const double* y, double* r) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = x[3 * i] * y[j] + x[3 * i + 1] * y[3 + j] + x[3 * i + 2] * y[6 + j];
  for (int i = 0; i < 9; ++i) r[i] = t[i];
}
// The reference inverts with LAPACK GESV (MxDimMatrix.hpp:242-256); the cofactor form gives the same exact result for
// the signed permutation / reflection matrices its examples produce.
inline void m33Inv(const double* a, double* r) {
  const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
  double t[9];
  t[0] = (a[4] * a[8] - a[5] * a[7]) / det;
  t[1] = (a[2] * a[7] - a[1] * a[8]) / det;
  t[2] = (a[1] * a[5] - a[2] * a[4]) / det;
  t[3] = (a[5] * a[6] - a[3] * a[8]) / det;
  t[4] = (a[0] * a[8] - a[2] * a[6]) / det;
  t[5] = (a[2] * a[3] - a[0] * a[5]) / det;
  t[6] = (a[3] * a[7] - a[4] * a[6]) / det;
  t[7] = (a[1] * a[6] - a[0] * a[7]) / det;
  t[8] = (a[0] * a[4] - a[1] * a[3]) / det;
  for (int i = 0; i < 9; ++i) r[i] = t[i];
}

// A shape under construction: node 0 is the root.
struct Shape {
  std::vector<ShapeNode> nodes;
  ShapeNode& root() { return nodes[0]; }
  const ShapeNode& root() const { return nodes[0]; }

  static ShapeNode blank(int type) {
    ShapeNode n;
    std::memset(&n, 0, sizeof(n));
    n.type = type;
    n.firstChild = n.nextSibling = -1;
    n.sign = 1.0;
    n.A[0] = n.A[4] = n.A[8] = 1.0;
    n.Ainv[0] = n.Ainv[4] = n.Ainv[8] = 1.0;
    return n;
  }
  static V3 unit(const double a[3]) {
    const V3 v = {{a[0], a[1], a[2]}};
    return vDiv(v, vNorm(v));
  }
  static void complAxisProj(V3 a, double* P) {           // I - a a^T
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) P[3 * i + j] = (i == j ? 1.0 : 0.0) - a.v[i] * a.v[j];
  }
  void refreshInverse() {
    ShapeNode& n = root();
    m33Inv(n.A, n.Ainv);
    const V3 ab = mTimes(n.Ainv, V3{{n.b[0], n.b[1], n.b[2]}});
    for (int i = 0; i < 3; ++i) n.Ainvb[i] = ab.v[i];
  }
  void translate(const double v[3]) {                     // MxShape.cpp:179-185
    ShapeNode& n = root();
    for (int i = 0; i < 3; ++i) n.b[i] = n.b[i] + v[i];
    const V3 ab = mTimes(n.Ainv, V3{{n.b[0], n.b[1], n.b[2]}});
    for (int i = 0; i < 3; ++i) n.Ainvb[i] = ab.v[i];
  }
  void applyLinear(const double* M, const double pivot[3]) {   // A <- M A, b <- M (b - pivot) + pivot
    ShapeNode& n = root();
    m33Mul(M, n.A, n.A);
    const V3 t = vAdd(mTimes(M, vSub(V3{{n.b[0], n.b[1], n.b[2]}}, V3{{pivot[0], pivot[1], pivot[2]}})), V3{{pivot[0], pivot[1], pivot[2]}});
    for (int i = 0; i < 3; ++i) n.b[i] = t.v[i];
    refreshInverse();
  }
  void reflect(const double normal[3], const double pointInPlane[3]) {   // MxShape.cpp:196-210
    const V3 nn = unit(normal);
    double M[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) M[3 * i + j] = (i == j ? 1.0 : 0.0) - 2.0 * (nn.v[i] * nn.v[j]);
    applyLinear(M, pointInPlane);
  }
  // MxShape.cpp:89-125. The reference fills the cross-product matrix with the opposite sign of the usual convention, so
  // R = I + M sin + M^2 (1 - cos) turns by -angle about the axis; restated literally. pivot NULL = the translation point.
  void rotate(const double axis[3], double angle, const double* pivot) {
    const V3 a = unit(axis);
    const double M[9] = {0, a.v[2], -a.v[1], -a.v[2], 0, a.v[0], a.v[1], -a.v[0], 0};
    double M2[9], R[9];
    m33Mul(M, M, M2);
    for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + (M[i] * std::sin(angle) + M2[i] * (1.0 - std::cos(angle)));
    const double own[3] = {root().b[0], root().b[1], root().b[2]};
    applyLinear(R, pivot ? pivot : own);
  }
  void scale(const double mags[3], const double origin[3]) {   // MxShape.cpp:158-168
    const double S[9] = {mags[0], 0, 0, 0, mags[1], 0, 0, 0, mags[2]};
    applyLinear(S, origin);
  }
  void invert() { root().sign *= -1.0; }

  // composite: a new root over copies of the parts
  static Shape compose(int type, const std::vector<const Shape*>& parts) {
    Shape s;
    s.nodes.push_back(blank(type));
    int prev = -1;
    for (const Shape* p : parts) {
      const int off = int(s.nodes.size());
      for (ShapeNode n : p->nodes) {
        if (n.firstChild >= 0) n.firstChild += off;
        if (n.nextSibling >= 0) n.nextSibling += off;
        s.nodes.push_back(n);
      }
      if (prev < 0) s.nodes[0].firstChild = off; else s.nodes[prev].nextSibling = off;
      prev = off;
    }
    return s;
  }
  int depth(int idx = 0) const {
    int d = 0;
    for (int c = nodes[idx].firstChild; c >= 0; c = nodes[c].nextSibling) d = std::max(d, depth(c));
    return d + 1;
  }
};

inline Shape makeCylinder(double r, const double axis[3], const double loc[3]) {   // MxCylinder.hpp:34-40,74-84
  Shape s;
  s.nodes.push_back(Shape::blank(SHAPE_CYLINDER));
  s.root().par[0] = r * r;
  Shape::complAxisProj(Shape::unit(axis), s.root().par + 1);
  s.translate(loc);
  return s;
}
inline Shape makeHalfSpace(const double pointInPlane[3], const double normal[3]) {   // MxHalfSpace.hpp:31-33
  Shape s;
  s.nodes.push_back(Shape::blank(SHAPE_HALFSPACE));
  const V3 n = Shape::unit(normal);
  for (int i = 0; i < 3; ++i) s.root().par[i] = n.v[i];
  s.translate(pointInPlane);
  return s;
}
inline Shape makeSphere(double r, const double loc[3]) {   // MxSphere.hpp:31-33
  Shape s;
  s.nodes.push_back(Shape::blank(SHAPE_SPHERE));
  s.root().par[0] = r * r;
  s.translate(loc);
  return s;
}
inline Shape makeEllipsoid(const double loc[3], const double axes[3]) {   // MxEllipsoid.hpp:31-33,44-50
  Shape s;
  s.nodes.push_back(Shape::blank(SHAPE_ELLIPSOID));
  for (int i = 0; i < 3; ++i) s.root().par[i] = 1.0 / (axes[i] * axes[i]);
  s.translate(loc);
  return s;
}
inline Shape makeTorus(double majorRadius, double minorRadius, const double axis[3], const double loc[3]) {   // MxTorus.hpp:31-37
  Shape s;
  s.nodes.push_back(Shape::blank(SHAPE_TORUS));
  const V3 a = Shape::unit(axis);
  for (int i = 0; i < 3; ++i) s.root().par[i] = a.v[i];
  s.root().par[3] = majorRadius;
  s.root().par[4] = minorRadius * minorRadius;
  Shape::complAxisProj(a, s.root().par + 5);
  s.translate(loc);
  return s;
}
inline Shape makeCone(double angle, const double axis[3], const double vertex[3]) {   // MxCone.hpp:31-37
  Shape s;
  s.nodes.push_back(Shape::blank(SHAPE_CONE));
  const V3 a = Shape::unit(axis);
  for (int i = 0; i < 3; ++i) s.root().par[i] = a.v[i];
  s.root().par[3] = std::tan(angle);
  Shape::complAxisProj(a, s.root().par + 4);
  s.translate(vertex);
  return s;
}
inline Shape makeSlab(double thickness, const double normal[3], const double loc[3]) {   // MxSlab.hpp:33-39,95-100
  const V3 n = Shape::unit(normal);
  const V3 lo = vScale(-0.5 * thickness, n), hi = vScale(0.5 * thickness, n), minusN = vScale(-1.0, n);
  const Shape a = makeHalfSpace(lo.v, n.v), b = makeHalfSpace(hi.v, minusN.v);
  Shape s = Shape::compose(SHAPE_INTERSECTION, {&a, &b});
  s.translate(loc);
  return s;
}
inline Shape makeSubtract(const Shape& base, const std::vector<const Shape*>& removed) {   // MxShapeSubtract.hpp:14-18
  std::vector<const Shape*> parts{&base};
  parts.insert(parts.end(), removed.begin(), removed.end());
  return Shape::compose(SHAPE_SUBTRACT, parts);
}
inline Shape makeMirror(const Shape& shape, const double normal[3], const double pointInPlane[3]) {   // MxShapeMirror.hpp:85-116
  const V3 n1 = Shape::unit(normal);
  Shape mirrored = shape;
  mirrored.reflect(n1.v, pointInPlane);
  Shape s = Shape::compose(SHAPE_MIRROR, {&shape, &mirrored});
  const V3 n2 = Shape::unit(n1.v);       // the plane's own constructor normalises once more
  for (int i = 0; i < 3; ++i) { s.root().par[i] = n2.v[i]; s.root().par[3 + i] = pointInPlane[i]; }
  return s;
}
inline Shape makeRepeat(const Shape& shape, const double origin[3], const double direction[3], double step, int numPos, int numNeg) {
  Shape s = Shape::compose(SHAPE_REPEAT, {&shape});     // MxShapeRepeat.hpp:84-118
  const V3 d = Shape::unit(direction);
  for (int i = 0; i < 3; ++i) { s.root().par[i] = origin[i]; s.root().par[3 + i] = d.v[i]; }
  s.root().par[6] = step;
  s.root().par[7] = double(numPos);
  s.root().par[8] = double(-numNeg);
  return s;
}

}  // namespace mxa
