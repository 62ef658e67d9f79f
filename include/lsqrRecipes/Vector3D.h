// Drop-in counterpart of common/Vector3D.h + Vector3D.cxx:5-12 of zivy/LSQRRecipes (re-authored, header-only).
#ifndef LSQR_B200_VECTOR3D_H
#define LSQR_B200_VECTOR3D_H
#include "Vector.h"
namespace lsqrRecipes {
typedef Vector<double, 3> Vector3D;
inline Vector3D crossProduct(const Vector3D& a, const Vector3D& b) {
  Vector3D r;
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
  return r;
}
}  // namespace lsqrRecipes
#endif
