// Drop-in counterpart of the data type of common/Ray3D.h:20-29 of zivy/LSQRRecipes (re-authored).
// r(t) = p + t*n, t in [0, inf).  Layout contract: p[3] then n[3] (48 bytes).  The distance /
// intersection helpers and the OpenInventor writer of the reference are not on the RANSAC path
//  and are not provided.
#ifndef LSQR_B200_RAY3D_H
#define LSQR_B200_RAY3D_H
#include <ostream>

#include "Frame.h"
#include "Point3D.h"
#include "Vector3D.h"

namespace lsqrRecipes {

class Ray3D {
 public:
  Point3D p;
  Vector3D n;
  Ray3D() {}
  Ray3D(const Ray3D& other) : p(other.p), n(other.n) {}
  Ray3D& operator=(const Ray3D& other) { p = other.p; n = other.n; return *this; }
  void transform(Frame& transformation) { transformation.apply(p); transformation.apply(n); }
  friend std::ostream& operator<<(std::ostream& out, const Ray3D& r) { return out << "p: " << r.p << " n: " << r.n; }
};

}  // namespace lsqrRecipes
#endif
