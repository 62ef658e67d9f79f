// Drop-in counterpart of common/Epsilon.h:19 of zivy/LSQRRecipes (re-authored).
#ifndef LSQR_B200_EPSILON_H
#define LSQR_B200_EPSILON_H
namespace lsqrRecipes {
// Same value as the reference (DBL_EPSILON); the device code carries its own copy (models.cuh kEps).
const double EPS = 2.220446049250313e-016;
}  // namespace lsqrRecipes
#endif
