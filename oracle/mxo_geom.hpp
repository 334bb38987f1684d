// TEST INFRASTRUCTURE ONLY -- CPU oracle for maxwell_b200.
// Nothing under oracle/ may be imported, linked or executed by the product path
// (maxwell_b200/, include/). Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, and only as the checker.
//
// Geometry restatement: implicit-function solids and the cut-cell fraction rules that
// feed the Dey-Mittra operators. Each routine cites the reference file:line it follows
// (paths relative to /root/reference/src).
#pragma once
#include <array>
#include <cmath>
#include <limits>
#include <memory>
#include <vector>

namespace mxo {

using D3 = std::array<double, 3>;
using I3 = std::array<int, 3>;

inline D3 operator+(const D3& a, const D3& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
inline D3 operator-(const D3& a, const D3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline D3 operator*(double s, const D3& a) { return {s * a[0], s * a[1], s * a[2]}; }
inline D3 operator*(const D3& a, double s) { return {a[0] * s, a[1] * s, a[2] * s}; }
inline D3 operator/(const D3& a, double s) { return {a[0] / s, a[1] / s, a[2] / s}; }
inline double dot(const D3& a, const D3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double norm(const D3& a) { return std::sqrt(dot(a, a)); }
inline D3 cross(const D3& a, const D3& b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}

// MxUtil.hpp:26,44-49 -- sign with dEps = 0: -1 / 0 / +1.
inline int sgn(double v) { return v < 0.0 ? -1 : (v > 0.0 ? 1 : 0); }

// ---------------------------------------------------------------------------------------
// Shapes: f > 0 inside. MxShape.hpp:143-165: func(p) = sign * f0(Ainv p - Ainv b) with the
// affine placement x -> A x + b built up by translate / reflect / rotate / scale (MxShape.cpp:89-210).
// ---------------------------------------------------------------------------------------
using R33 = std::array<std::array<double, 3>, 3>;
inline R33 r33Eye() { return {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }
inline D3 operator*(const R33& m, const D3& v) {
  D3 r;
  for (int i = 0; i < 3; ++i) r[i] = m[i][0] * v[0] + m[i][1] * v[1] + m[i][2] * v[2];
  return r;
}
inline D3 rowTimes(const D3& v, const R33& m) {   // v^T M (MxDimVector * MxDimMatrix)
  D3 r;
  for (int j = 0; j < 3; ++j) r[j] = v[0] * m[0][j] + v[1] * m[1][j] + v[2] * m[2][j];
  return r;
}
inline R33 operator*(const R33& x, const R33& y) {
  R33 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[i][j] = x[i][0] * y[0][j] + x[i][1] * y[1][j] + x[i][2] * y[2][j];
  return r;
}
inline R33 operator*(double s, const R33& m) {
  R33 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[i][j] = s * m[i][j];
  return r;
}
// The reference inverts with LAPACK GESV (MxDimMatrix.hpp:242-256); the cofactor form gives the same
// exact result for the signed permutation / reflection matrices the configs produce.
inline R33 r33Inv(const R33& a) {
  const double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                     a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
  R33 r;
  r[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) / det;
  r[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) / det;
  r[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) / det;
  r[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) / det;
  r[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) / det;
  r[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) / det;
  r[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) / det;
  r[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) / det;
  r[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) / det;
  return r;
}
// P = I - a a^T for a unit axis (the "complementary axis projection" of cylinder / torus / cone)
inline R33 complAxisProj(const D3& a) {
  R33 P = r33Eye();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) P[i][j] -= a[i] * a[j];
  return P;
}

struct Shape {
  double sign = 1.0;
  D3 b{0, 0, 0}, Ainvb{0, 0, 0};
  R33 A = r33Eye(), Ainv = r33Eye();
  virtual ~Shape() {}
  virtual double f0(const D3& p) const = 0;
  virtual D3 g0(const D3& p) const = 0;
  virtual std::shared_ptr<Shape> clone() const = 0;   // copy; composite shapes keep sharing their parts
  double func(const D3& p) const { return sign * f0(Ainv * p - Ainvb); }
  D3 grad(const D3& p) const { return sign * rowTimes(g0(Ainv * p - Ainvb), Ainv); }
  void translate(const D3& v) {   // MxShape.cpp:179-185
    b = b + v;
    Ainvb = Ainv * b;
  }
  void reflect(const D3& normal, const D3& pointInPlane) {   // MxShape.cpp:196-210
    const D3 n = normal / norm(normal);
    R33 M = r33Eye();
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) M[i][j] -= 2.0 * (n[i] * n[j]);
    A = M * A;
    b = M * (b - pointInPlane) + pointInPlane;
    Ainv = r33Inv(A);
    Ainvb = Ainv * b;
  }
  // MxShape.cpp:89-125. The reference fills the cross-product matrix with the opposite sign of the usual convention
  // (M = -[a]x), so R = I + M sin + M^2 (1 - cos) turns by -angle about the axis; restated literally.
  void rotate(const D3& axis, double angle, const D3& pivot) {
    const D3 a = axis / norm(axis);
    R33 M{{{0, a[2], -a[1]}, {-a[2], 0, a[0]}, {a[1], -a[0], 0}}};
    const R33 M2 = M * M;
    R33 R = r33Eye();
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R[i][j] += M[i][j] * std::sin(angle) + M2[i][j] * (1.0 - std::cos(angle));
    A = R * A;
    b = R * (b - pivot) + pivot;
    Ainv = r33Inv(A);
    Ainvb = Ainv * b;
  }
  void rotate(const D3& axis, double angle) { rotate(axis, angle, b); }   // about the current translation point
  // MxShape.cpp:158-168
  void scale(const D3& magnitudes, const D3& origin = D3{0, 0, 0}) {
    R33 S{{{magnitudes[0], 0, 0}, {0, magnitudes[1], 0}, {0, 0, magnitudes[2]}}};
    A = S * A;
    b = S * (b - origin) + origin;
    Ainv = r33Inv(A);
    Ainvb = Ainv * b;
  }
  void invert() { sign *= -1.0; }
};
#define MXO_CLONE(T) std::shared_ptr<Shape> clone() const override { return std::make_shared<T>(*this); }

// MxCylinder.hpp:34-40,74-84: f = r^2 - p.(P p), P = I - a a^T.
struct Cylinder : Shape {
  double r2;
  double P[3][3];
  Cylinder(double r, D3 axis, D3 loc) : r2(r * r) {
    axis = axis / norm(axis);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) P[i][j] = (i == j ? 1.0 : 0.0) - axis[i] * axis[j];
    translate(loc);
  }
  D3 Pp(const D3& p) const {
    D3 r{0, 0, 0};
    for (int i = 0; i < 3; ++i) {
      double s = 0;
      for (int j = 0; j < 3; ++j) s += P[i][j] * p[j];
      r[i] = s;
    }
    return r;
  }
  double f0(const D3& p) const override { return r2 - dot(p, Pp(p)); }
  D3 g0(const D3& p) const override { return -2.0 * Pp(p); }
  MXO_CLONE(Cylinder)
};

// MxHalfSpace.hpp:31-33: f = n.p
struct HalfSpace : Shape {
  D3 n;
  HalfSpace(D3 pointInPlane, D3 normal) : n(normal / norm(normal)) { translate(pointInPlane); }
  double f0(const D3& p) const override { return dot(n, p); }
  D3 g0(const D3&) const override { return n; }
  MXO_CLONE(HalfSpace)
};

// MxSphere.hpp:31-33: f = 1 - p.p / r^2
struct Sphere : Shape {
  double r2;
  Sphere(double r, D3 loc) : r2(r * r) { translate(loc); }
  double f0(const D3& p) const override { return 1.0 - dot(p, p) / r2; }
  D3 g0(const D3& p) const override { return -2.0 * p / r2; }
  MXO_CLONE(Sphere)
};

// MxEllipsoid.hpp:31-33,44-50: f = 1 - sum_i p_i^2 / a_i^2
struct Ellipsoid : Shape {
  D3 invAxes2;
  Ellipsoid(D3 loc, D3 axes) : invAxes2{1.0 / (axes[0] * axes[0]), 1.0 / (axes[1] * axes[1]), 1.0 / (axes[2] * axes[2])} { translate(loc); }
  double f0(const D3& p) const override { return 1.0 - (p[0] * (invAxes2[0] * p[0]) + p[1] * (invAxes2[1] * p[1]) + p[2] * (invAxes2[2] * p[2])); }
  D3 g0(const D3& p) const override { return {-2.0 * invAxes2[0] * p[0], -2.0 * invAxes2[1] * p[1], -2.0 * invAxes2[2] * p[2]}; }
  MXO_CLONE(Ellipsoid)
};

// MxShapeIntersection.hpp:120-136 (min of sub-shape funcs), :186-206 (gradient of the
// first strict minimiser).
struct Intersection : Shape {
  std::vector<std::shared_ptr<Shape>> subs;
  double f0(const D3& p) const override {
    double fmin = std::numeric_limits<double>::max();
    for (auto& s : subs) {
      double f = s->func(p);
      if (f < fmin) fmin = f;
    }
    return fmin;
  }
  D3 g0(const D3& p) const override {
    double fmin = std::numeric_limits<double>::max();
    const Shape* arg = nullptr;
    for (auto& s : subs) {
      double f = s->func(p);
      if (f < fmin) { fmin = f; arg = s.get(); }
    }
    return arg->grad(p);
  }
  MXO_CLONE(Intersection)
};

// MxSlab.hpp:33-39,95-100: two half-spaces at -/+ l/2 along n, then translated.
inline std::shared_ptr<Shape> makeSlab(double thickness, D3 n, D3 loc) {
  n = n / norm(n);
  auto s = std::make_shared<Intersection>();
  s->subs.push_back(std::make_shared<HalfSpace>(-0.5 * thickness * n, n));
  s->subs.push_back(std::make_shared<HalfSpace>(0.5 * thickness * n, -1.0 * n));
  s->translate(loc);
  return s;
}

// MxTorus.hpp:31-37,51-65: f = r^2 - (R - |P p|)^2 - (a.p)^2. The gradient is restated literally: its in-plane
// term carries the opposite sign of the true derivative (quirk R14); it only steers the Newton proposals of
// rootFind, whose bisection safeguard still brackets the root.
struct Torus : Shape {
  D3 axis;
  double R, r2;
  R33 P;
  Torus(double majorRadius, double minorRadius, D3 torusAxis, D3 loc)
      : axis(torusAxis / norm(torusAxis)), R(majorRadius), r2(minorRadius * minorRadius), P(complAxisProj(axis)) {
    translate(loc);
  }
  double f0(const D3& p) const override {
    const double d = R - norm(P * p), h = dot(axis, p);
    return r2 - d * d - h * h;
  }
  D3 g0(const D3& p) const override {
    const double c = 1.0 - R / norm(P * p);
    return 2.0 * (((c * P) * p) - dot(axis, p) * axis);
  }
  MXO_CLONE(Torus)
};

// MxCone.hpp:31-37,52-62: double cone about `axis` with half-angle theta, vertex at `vertex`.
struct Cone : Shape {
  D3 axis;
  double tanTheta;
  R33 P;
  Cone(double angle, D3 coneAxis, D3 vertex) : axis(coneAxis / norm(coneAxis)), tanTheta(std::tan(angle)), P(complAxisProj(axis)) {
    translate(vertex);
  }
  double f0(const D3& p) const override {
    const double t = tanTheta * dot(axis, p);
    return t * t - dot(p, P * p);
  }
  D3 g0(const D3& p) const override { return 2.0 * ((tanTheta * tanTheta * dot(axis, p)) * axis - P * p); }
  MXO_CLONE(Cone)
};

// MxShapeUnion.hpp:62-78 (max of the sub-shape funcs), :135-154 (gradient of the first strict maximiser).
struct Union : Shape {
  std::vector<std::shared_ptr<Shape>> subs;
  double f0(const D3& p) const override {
    double fmax = -std::numeric_limits<double>::max();
    for (auto& s : subs) {
      const double f = s->func(p);
      if (f > fmax) fmax = f;
    }
    return fmax;
  }
  D3 g0(const D3& p) const override {
    double fmax = -std::numeric_limits<double>::max();
    const Shape* arg = nullptr;
    for (auto& s : subs) {
      const double f = s->func(p);
      if (f > fmax) { fmax = f; arg = s.get(); }
    }
    return arg->grad(p);
  }
  MXO_CLONE(Union)
};

// MxShapeSubtract.hpp:14-18,69-92,158-176: base minus an (inverted) union of removal shapes.
struct Subtract : Shape {
  std::shared_ptr<Shape> base;
  Union rm;
  Subtract() { rm.invert(); }
  double f0(const D3& p) const override {
    const double fBase = base->func(p), fRm = rm.func(p);
    const bool inBase = fBase > 0, inRm = fRm < 0;
    if (inRm && inBase) return fRm;
    if (!inRm && !inBase) return fBase;
    return fBase < fRm ? fBase : fRm;
  }
  D3 g0(const D3& p) const override {
    const double fBase = base->func(p), fRm = rm.func(p);
    const bool inBase = fBase > 0, inRm = fRm < 0;
    if (inRm && inBase) return rm.grad(p);
    if (!inRm && !inBase) return base->grad(p);
    return fBase < fRm ? base->grad(p) : rm.grad(p);
  }
  MXO_CLONE(Subtract)
};

// MxShapeMirror.hpp:85-116,155-165: the shape on the positive side of the plane, its reflected clone elsewhere.
struct Mirror : Shape {
  std::shared_ptr<Shape> shape, mirrored;
  HalfSpace plane;
  Mirror(std::shared_ptr<Shape> s, D3 normal, D3 pointInPlane) : shape(s), plane(pointInPlane, normal / norm(normal)) {
    mirrored = shape->clone();
    mirrored->reflect(normal / norm(normal), pointInPlane);
  }
  double f0(const D3& p) const override { return plane.func(p) > 0.0 ? shape->func(p) : mirrored->func(p); }
  D3 g0(const D3& p) const override { return plane.func(p) > 0.0 ? shape->grad(p) : mirrored->grad(p); }
  MXO_CLONE(Mirror)
};

// MxShapeRepeat.hpp:84-118,138-148: fold p back into the base period along `dir`, clamped to [-numNeg, numPos] periods.
struct Repeat : Shape {
  std::shared_ptr<Shape> shape;
  D3 o, dir;
  double s, np, nn;
  Repeat(std::shared_ptr<Shape> sh, D3 origin, D3 direction, double step, int numPos, int numNeg)
      : shape(sh), o(origin), dir(direction / norm(direction)), s(step), np(numPos), nn(-numNeg) {}
  D3 fold(const D3& p) const {
    const double slabPt = dot(p - o, dir) / s + 0.5;
    if (slabPt >= 0.0) return p - (s * std::min(std::floor(slabPt), np)) * dir;
    return p - (s * std::max(std::floor(slabPt), nn)) * dir;
  }
  double f0(const D3& p) const override { return shape->func(fold(p)); }
  D3 g0(const D3& p) const override { return shape->grad(fold(p)); }
  MXO_CLONE(Repeat)
};

// ---------------------------------------------------------------------------------------
// Safeguarded Newton / bisection on the segment p1->p2 (MxUtil.hpp:295-362). The
// reference ignores the caller's maxiter and uses 500; tolerance is tol * |p2 - p1|.
// ---------------------------------------------------------------------------------------
inline D3 rootFind(const Shape& sh, const D3& p1, const D3& p2, double tol) {
  const int maxiter = 500;
  const double len = norm(p2 - p1);
  const double stol = tol * len;
  const D3 dir = (p2 - p1) / len;
  const double f1 = sh.func(p1), f2 = sh.func(p2);
  if (f1 == 0) return p1;
  if (f2 == 0) return p2;
  double lo, hi;  // lo: f < 0 side, hi: f >= 0 side
  if (f1 < 0) { lo = 0.0; hi = len; } else { lo = len; hi = 0.0; }
  double t = 0.5 * (lo + hi);
  double stepPrev = std::fabs(hi - lo), step = stepPrev;
  double f = sh.func(p1 + dir * t);
  double df = dot(dir, sh.grad(p1 + dir * t));
  for (int it = 0; it < maxiter; ++it) {
    const bool outOfBracket = ((t - hi) * df - f) * ((t - lo) * df - f) >= 0;
    const bool slow = std::fabs(2.0 * f) > std::fabs(stepPrev * df);
    if (outOfBracket || slow) {
      stepPrev = step;
      step = 0.5 * (hi - lo);
      t = lo + step;
      if (lo == t) return p1 + dir * t;
    } else {
      stepPrev = step;
      step = f / df;
      const double told = t;
      t -= step;
      if (told == t) return p1 + dir * t;
    }
    if (std::fabs(step) < stol) return p1 + dir * t;
    f = sh.func(p1 + dir * t);
    df = dot(dir, sh.grad(p1 + dir * t));
    if (f < 0) lo = t; else hi = t;
  }
  return p1 + dir * t;
}

// ---------------------------------------------------------------------------------------
// Polytopes (cell-relative primitives whose inside-fraction is the Dey-Mittra weight)
// ---------------------------------------------------------------------------------------

// MxSegment.cpp:22-43 with MxCartSeg.hpp: axis-aligned edge of length len, midpoint mid.
inline double segmentFraction(const Shape& sh, int axis, double len, const D3& mid) {
  D3 d{0, 0, 0};
  d[axis] = 1.0;
  const D3 p1 = mid + 0.5 * len * d;
  const D3 p2 = mid - 0.5 * len * d;
  const double f1 = sh.func(p1), f2 = sh.func(p2);
  const int s1 = sgn(f1), s2 = sgn(f2);
  if ((s1 == 1 && s2 != -1) || (s2 == 1 && s1 != -1)) return 1;
  if ((s1 == -1 && s2 != 1) || (s2 == -1 && s1 != 1)) return 0;
  if (s1 == 0 && s2 == 0) return sgn(sh.func(mid)) == 1 ? 1 : 0;
  const D3 p = rootFind(sh, p1, p2, 1.e-12);
  return (s1 == 1 ? norm(p - p1) : norm(p - p2)) / len;
}

// MxCartRect.hpp:34-88 (vertex/edge numbering) + MxConvexPolygon.cpp:25-76,257-292 (fan
// of triangles from the first edge crossing) + MxPolytope.cpp:7-36,40-97 (state, edge
// crossings). type: 0 = x-face (d1=y,d2=z), 1 = y-face (d1=z,d2=x), 2 = z-face (d1=x,d2=y).
inline double rectFraction(const Shape& sh, int type, double l1, double l2, const D3& p) {
  D3 d1{0, 0, 0}, d2{0, 0, 0};
  d1[(type + 1) % 3] = 1.0;
  d2[(type + 2) % 3] = 1.0;
  D3 v[4];
  v[0] = p + 0.5 * (-l1 * d1 - l2 * d2);
  v[1] = p + 0.5 * (l1 * d1 - l2 * d2);
  v[2] = p + 0.5 * (-l1 * d1 + l2 * d2);
  v[3] = p + 0.5 * (l1 * d1 + l2 * d2);
  static const int E[4][2] = {{0, 1}, {0, 2}, {1, 3}, {2, 3}};
  double fv[4];
  bool noneOut = true, noneIn = true, allOn = true;
  for (int i = 0; i < 4; ++i) {
    fv[i] = sh.func(v[i]);
    const int s = sgn(fv[i]);
    if (s == -1) { noneOut = false; allOn = false; }
    else if (s == 1) { noneIn = false; allOn = false; }
  }
  if (allOn) return sgn(sh.func(p)) == 1 ? 1 : 0;
  if (noneOut) return 1;
  if (noneIn) return 0;
  // edge crossings
  bool has[4] = {false, false, false, false};
  D3 ex[4];
  int first = -1;
  for (int e = 0; e < 4; ++e) {
    const int a = E[e][0], c = E[e][1];
    const int sa = sgn(fv[a]), sc = sgn(fv[c]);
    if ((sa == -1 && sc == 1) || (sa == 1 && sc == -1)) { ex[e] = rootFind(sh, v[a], v[c], 1.e-12); has[e] = true; }
    else if (sa == 0 && sc == 1) { ex[e] = v[a]; has[e] = true; }
    else if (sc == 0 && sa == 1) { ex[e] = v[c]; has[e] = true; }
    if (has[e] && first < 0) first = e;
  }
  const D3 v0 = ex[first];  // MxConvexPolygon.cpp:166-253: 3-D getCorner returns the first crossing
  double area = 0.0;
  for (int e = 0; e < 4; ++e) {
    const int a = E[e][0], c = E[e][1];
    const int sa = sgn(fv[a]), sc = sgn(fv[c]);
    D3 q1, q2;
    if (has[e]) { q1 = sa != -1 ? v[a] : ex[e]; q2 = sc != -1 ? v[c] : ex[e]; }
    else if (sa == 1 || sc == 1) { q1 = v[a]; q2 = v[c]; }
    else continue;
    area += 0.5 * norm(cross(q1 - v0, q2 - v0));
  }
  return area / (l1 * l2);
}

// MxCartBox.cpp:21-62 (connectivity) + MxConvexPolyhedron.cpp:130-344: pyramid sum over
// inside faces from each cut-edge point, weighted by the cut lengths of adjacent faces.
inline double boxFraction(const Shape& sh, double lx, double ly, double lz, const D3& p) {
  D3 v[8];
  for (int i = 0; i < 8; ++i) {
    v[i] = p;
    v[i][0] += (i & 4 ? 0.5 : -0.5) * lx;
    v[i][1] += (i & 2 ? 0.5 : -0.5) * ly;
    v[i][2] += (i & 1 ? 0.5 : -0.5) * lz;
  }
  static const int EV[12][2] = {{0, 4}, {1, 5}, {2, 6}, {3, 7}, {0, 2}, {1, 3},
                                {4, 6}, {5, 7}, {0, 1}, {2, 3}, {4, 5}, {6, 7}};
  static const int EF[12][2] = {{0, 2}, {1, 2}, {0, 3}, {1, 3}, {0, 4}, {1, 4},
                                {0, 5}, {1, 5}, {2, 4}, {3, 4}, {2, 5}, {3, 5}};
  // face -> edges (ascending edge index) and face -> verts (ascending vertex index)
  int FE[6][4], nFE[6] = {0, 0, 0, 0, 0, 0};
  for (int e = 0; e < 12; ++e)
    for (int k = 0; k < 2; ++k) FE[EF[e][k]][nFE[EF[e][k]]++] = e;
  int FV[6][4], nFV[6] = {0, 0, 0, 0, 0, 0};
  for (int f = 0; f < 6; ++f)
    for (int vi = 0; vi < 8; ++vi) {
      bool on = false;
      for (int k = 0; k < 4 && !on; ++k)
        if (EV[FE[f][k]][0] == vi || EV[FE[f][k]][1] == vi) on = true;
      if (on) FV[f][nFV[f]++] = vi;
    }
  double fv[8];
  bool noneOut = true, noneIn = true, allOn = true;
  for (int i = 0; i < 8; ++i) {
    fv[i] = sh.func(v[i]);
    const int s = sgn(fv[i]);
    if (s == -1) { noneOut = false; allOn = false; }
    else if (s == 1) { noneIn = false; allOn = false; }
  }
  if (allOn) return sgn(sh.func(p)) == 1 ? 1 : 0;
  if (noneOut) return 1;
  if (noneIn) return 0;
  bool has[12];
  D3 ex[12];
  for (int e = 0; e < 12; ++e) {
    has[e] = false;
    const int a = EV[e][0], c = EV[e][1];
    const int sa = sgn(fv[a]), sc = sgn(fv[c]);
    if ((sa == -1 && sc == 1) || (sa == 1 && sc == -1)) { ex[e] = rootFind(sh, v[a], v[c], 1.e-12); has[e] = true; }
    else if (sa == 0 && sc == 1) { ex[e] = v[a]; has[e] = true; }
    else if (sc == 0 && sa == 1) { ex[e] = v[c]; has[e] = true; }
  }
  bool faceUsed[6];
  D3 faceArea[6], faceVert[6];
  double cutLen[6];
  for (int f = 0; f < 6; ++f) {
    faceUsed[f] = false;
    cutLen[f] = 0.0;
    bool inside = false;
    for (int k = 0; k < nFV[f]; ++k)
      if (sgn(fv[FV[f][k]]) == 1) { inside = true; break; }
    if (!inside) continue;
    int e0 = -1;
    for (int k = 0; k < 4; ++k)
      if (has[FE[f][k]]) { e0 = FE[f][k]; break; }
    D3 av{0, 0, 0};
    faceUsed[f] = true;
    if (e0 >= 0) {
      const D3 v0 = ex[e0];
      faceVert[f] = v0;
      for (int k = 0; k < 4; ++k) {
        const int e = FE[f][k];
        const int a = EV[e][0], c = EV[e][1];
        const int sa = sgn(fv[a]), sc = sgn(fv[c]);
        D3 q1, q2;
        if (has[e] && e != e0) {
          q1 = ex[e];
          q2 = sa != -1 ? v[a] : v[c];
          cutLen[f] = norm(v0 - q1);
        } else if (sa == 1 || sc == 1) { q1 = v[a]; q2 = v[c]; }
        else continue;
        const D3 tri = cross(q1 - v0, q2 - v0);
        av = dot(av, tri) > 0 ? av + tri : av - tri;
      }
    } else {
      const int base = FV[f][0];
      const D3 v0 = v[base];
      faceVert[f] = v0;
      for (int k = 0; k < 4; ++k) {
        const int e = FE[f][k];
        const int a = EV[e][0], c = EV[e][1];
        if (a == base || c == base) continue;
        const D3 tri = cross(v[a] - v0, v[c] - v0);
        av = dot(av, tri) > 0 ? av + tri : av - tri;
      }
    }
    faceArea[f] = 0.5 * av;
  }
  double vol = 0, wtSum = 0;
  for (int e = 0; e < 12; ++e) {
    if (!has[e]) continue;
    const double wt = cutLen[EF[e][0]] + cutLen[EF[e][1]];
    wtSum += wt;
    for (int f = 0; f < 6; ++f)
      if (faceUsed[f]) vol += wt * std::fabs(dot(faceVert[f] - ex[e], faceArea[f])) / 3.;
  }
  vol /= wtSum;
  return vol / (lx * ly * lz);
}

}  // namespace mxo
