// Drop-in counterpart of common/Point3D.h of zivy/LSQRRecipes (re-authored).
#ifndef LSQR_B200_POINT3D_H
#define LSQR_B200_POINT3D_H
#include "Point.h"
namespace lsqrRecipes {
typedef Point<double, 3> Point3D;
}  // namespace lsqrRecipes
#endif
