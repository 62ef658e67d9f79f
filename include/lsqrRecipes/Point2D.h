// Drop-in counterpart of common/Point2D.h of zivy/LSQRRecipes (re-authored).
#ifndef LSQR_B200_POINT2D_H
#define LSQR_B200_POINT2D_H
#include "Point.h"
namespace lsqrRecipes {
typedef Point<double, 2> Point2D;
}  // namespace lsqrRecipes
#endif
