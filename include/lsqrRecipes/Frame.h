// Drop-in counterpart of the part of common/Frame.{h,cxx} that lies on the RANSAC path
// (SURVEY.md section 2 row 11): construction from a matrix or a quaternion (Frame.cxx:55-76,
// :174-199), setRotationQuaternion (:750-772), setRotationMatrix (:674-698), setTranslation,
// get*, apply (:208-248) and getRotationQuaternion (:952-988).  Euler angles, axis-angle, slerp,
// lerp, invert and mul are pose utilities unrelated to consensus scoring and are out of scope.
// Re-authored, VNL-free, header-only.
//
// Layout contract: rotation[3][3] row-major, then translation[3], first in the
// object -- the 12 leading doubles of a Frame are the datum that crosses the C ABI (stride
// sizeof(Frame) = 104).
#ifndef LSQR_B200_FRAME_H
#define LSQR_B200_FRAME_H
#include <cmath>
#include <cstring>
#include <ostream>

#include "Point3D.h"
#include "Vector3D.h"

namespace lsqrRecipes {

class Frame {
 private:
  double rotation[3][3];
  double translation[3];

 public:
  enum { MATRIX = 0, EULER_ANGLES, AXIS_ANGLE, QUATERNION };
  int outputFormat;

  Frame(double R[3][3] = NULL, double t[3] = NULL) : outputFormat(MATRIX) {
    setIdentity();
    if (t) std::memcpy(translation, t, sizeof(translation));
    if (R) std::memcpy(rotation, R, sizeof(rotation));
  }
  Frame(const Frame& f) : outputFormat(MATRIX) { set(f); }
  Frame& operator=(const Frame& f) { set(f); return *this; }
  Frame(double x, double y, double z, double s, double qx, double qy, double qz, bool normalizeQuaternion = false) : outputFormat(MATRIX) {
    setTranslation(x, y, z);
    setRotationQuaternion(s, qx, qy, qz, normalizeQuaternion);
  }

  void set(const Frame& f) {
    std::memcpy(rotation, f.rotation, sizeof(rotation));
    std::memcpy(translation, f.translation, sizeof(translation));
  }
  void setIdentity() {
    std::memset(rotation, 0, sizeof(rotation));
    std::memset(translation, 0, sizeof(translation));
    rotation[0][0] = rotation[1][1] = rotation[2][2] = 1.0;
  }
  void setTranslation(double x, double y, double z) { translation[0] = x; translation[1] = y; translation[2] = z; }
  void setTranslation(double t[3]) { std::memcpy(translation, t, sizeof(translation)); }
  void setRotationMatrix(double R[3][3]) { std::memcpy(rotation, R, sizeof(rotation)); }
  void setRotationMatrix(double m00, double m01, double m02, double m10, double m11, double m12, double m20, double m21, double m22) {
    rotation[0][0] = m00; rotation[0][1] = m01; rotation[0][2] = m02;
    rotation[1][0] = m10; rotation[1][1] = m11; rotation[1][2] = m12;
    rotation[2][0] = m20; rotation[2][1] = m21; rotation[2][2] = m22;
  }
  void setRotationQuaternion(double s, double qx, double qy, double qz, bool normalizeQuaternion = false) {
    if (normalizeQuaternion) {
      const double norm = std::sqrt(s * s + qx * qx + qy * qy + qz * qz);
      s /= norm; qx /= norm; qy /= norm; qz /= norm;
    }
    rotation[0][0] = 1 - 2 * (qy * qy + qz * qz); rotation[0][1] = 2 * (qx * qy - s * qz);     rotation[0][2] = 2 * (qx * qz + s * qy);
    rotation[1][0] = 2 * (qx * qy + s * qz);     rotation[1][1] = 1 - 2 * (qx * qx + qz * qz); rotation[1][2] = 2 * (qy * qz - s * qx);
    rotation[2][0] = 2 * (qx * qz - s * qy);     rotation[2][1] = 2 * (qy * qz + s * qx);     rotation[2][2] = 1 - 2 * (qx * qx + qy * qy);
  }
  void setRotationQuaternion(double q[4], bool normalizeQuaternion = false) { setRotationQuaternion(q[0], q[1], q[2], q[3], normalizeQuaternion); }

  void getTranslation(double t[3]) const { std::memcpy(t, translation, sizeof(translation)); }
  void getTranslation(double& x, double& y, double& z) const { x = translation[0]; y = translation[1]; z = translation[2]; }
  void getRotationMatrix(double R[3][3]) const { std::memcpy(R, rotation, sizeof(rotation)); }
  void getRotationQuaternion(double q[4]) const {
    const double halfPi = 3.14159265358979323846 / 2.0, smallAngle = 0.008726535498373935;
    q[0] = 0.5 * std::sqrt(rotation[0][0] + rotation[1][1] + rotation[2][2] + 1);
    const double halfTheta = std::acos(q[0]);
    if (!(halfTheta > halfPi - smallAngle && halfTheta < halfPi + smallAngle)) {
      const double denom = 4 * q[0];
      q[1] = (rotation[2][1] - rotation[1][2]) / denom;
      q[2] = (rotation[0][2] - rotation[2][0]) / denom;
      q[3] = (rotation[1][0] - rotation[0][1]) / denom;
    } else {  // rotation by about 180 degrees: recover the vector part from the diagonal
      int i = 0;
      if (rotation[1][1] > rotation[i][i]) i = 1;
      if (rotation[2][2] > rotation[i][i]) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      const double w = std::sqrt(rotation[i][i] - rotation[j][j] - rotation[k][k] + 1);
      q[i + 1] = w / 2.0;
      q[j + 1] = (rotation[i][j] + rotation[j][i]) / (2 * w);
      q[k + 1] = (rotation[i][k] + rotation[k][i]) / (2 * w);
    }
  }

  void apply(const Point3D& p, Point3D& out) const {
    const double x = rotation[0][0] * p[0] + rotation[0][1] * p[1] + rotation[0][2] * p[2] + translation[0];
    const double y = rotation[1][0] * p[0] + rotation[1][1] * p[1] + rotation[1][2] * p[2] + translation[1];
    const double z = rotation[2][0] * p[0] + rotation[2][1] * p[1] + rotation[2][2] * p[2] + translation[2];
    out[0] = x; out[1] = y; out[2] = z;
  }
  void apply(Point3D& p) const { apply(p, p); }
  void apply(Vector3D& v) const {
    const double x = rotation[0][0] * v[0] + rotation[0][1] * v[1] + rotation[0][2] * v[2];
    const double y = rotation[1][0] * v[0] + rotation[1][1] * v[1] + rotation[1][2] * v[2];
    const double z = rotation[2][0] * v[0] + rotation[2][1] * v[1] + rotation[2][2] * v[2];
    v[0] = x; v[1] = y; v[2] = z;
  }

  // the 12 doubles that cross the C ABI
  const double* packed() const { return &rotation[0][0]; }

  friend std::ostream& operator<<(std::ostream& out, const Frame& f) {
    out << "translation:\n\t[ " << f.translation[0] << ", " << f.translation[1] << ", " << f.translation[2] << "]\nrotation:\n";
    for (int r = 0; r < 3; r++) out << "\t[" << f.rotation[r][0] << ", " << f.rotation[r][1] << ", " << f.rotation[r][2] << "]\n";
    return out;
  }
};

}  // namespace lsqrRecipes
#endif
