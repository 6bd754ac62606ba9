// Small multiview-geometry helpers used by the host-side scoring (role of the reference's `util` class,
// magmaHC/util.hpp:21-252).  Row-major 3x3 matrices, float arithmetic like the reference.
#ifndef HCB200_HOST_MVG_HPP
#define HCB200_HOST_MVG_HPP
#include <array>
#include <cmath>

namespace hcb200 { namespace mvg {

using Mat3 = std::array<float, 9>;
using Vec3 = std::array<float, 3>;

inline float norm3(const Vec3& v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
inline Vec3 normalized(Vec3 v) { const float n = norm3(v); return {v[0] / n, v[1] / n, v[2] / n}; }
inline float dot3(const Vec3& a, const Vec3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline Vec3 mul(const Mat3& M, const Vec3& v)
{ return {M[0] * v[0] + M[1] * v[1] + M[2] * v[2], M[3] * v[0] + M[4] * v[1] + M[5] * v[2], M[6] * v[0] + M[7] * v[1] + M[8] * v[2]}; }
inline Vec3 mul_transposed(const Mat3& M, const Vec3& v)
{ return {M[0] * v[0] + M[3] * v[1] + M[6] * v[2], M[1] * v[0] + M[4] * v[1] + M[7] * v[2], M[2] * v[0] + M[5] * v[1] + M[8] * v[2]}; }
inline Mat3 matmul(const Mat3& A, const Mat3& B)
{
  Mat3 C{};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { float s = 0; for (int k = 0; k < 3; k++) s += A[i * 3 + k] * B[k * 3 + j]; C[i * 3 + j] = s; }
  return C;
}
inline Mat3 transposed(const Mat3& A) { return {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]}; }
inline float det3(const Mat3& R)
{ return R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] - R[1] * R[3] * R[8] - R[0] * R[5] * R[7]; }

// Cayley parameters -> rotation (util.hpp:32-68): un-normalised Cayley matrix, then every COLUMN scaled to unit length.
inline Mat3 cayley_to_rotation(const Vec3& r)
{
  Mat3 R = {1 + r[0] * r[0] - (r[1] * r[1] + r[2] * r[2]), 2 * (r[0] * r[1] - r[2]), 2 * (r[0] * r[2] + r[1]),
            2 * (r[0] * r[1] + r[2]), 1 + r[1] * r[1] - (r[0] * r[0] + r[2] * r[2]), 2 * (r[1] * r[2] - r[0]),
            2 * (r[0] * r[2] - r[1]), 2 * (r[1] * r[2] + r[0]), 1 + r[2] * r[2] - (r[0] * r[0] + r[1] * r[1])};
  for (int c = 0; c < 3; c++) {
    const float n = std::sqrt(R[c] * R[c] + R[3 + c] * R[3 + c] + R[6 + c] * R[6 + c]);
    R[c] /= n; R[3 + c] /= n; R[6 + c] /= n;
  }
  return R;
}

// acos((trace(Rgt' R) - 1) / 2)   (Evaluations.cpp:360-374)
inline float rotation_residual(const Mat3& Rgt, const Mat3& R)
{
  const Mat3 P = matmul(transposed(Rgt), R);
  return std::acos(0.5 * ((P[0] + P[4] + P[8]) - 1.0));
}
// | <t_gt, t> - 1 | for unit vectors   (Evaluations.cpp:376-380)
inline float translation_residual(const Vec3& tgt, const Vec3& t) { return std::fabs(dot3(tgt, t) - 1.0); }

// depth of gamma1 in view 1 from a correspondence (gamma1, gamma2) and relative pose (R, T)   (util.hpp:169-186)
inline float depth_rho(const Vec3& g1, const Vec3& g2, const Mat3& R, const Vec3& T)
{
  const Vec3 Rtg2 = mul_transposed(R, g2);
  float rho = T[2] * Rtg2[2] - mul_transposed(R, T)[2];
  rho /= (float)(1 - mul(R, g1)[2] * Rtg2[2]);
  return rho;
}
// pixel distance between K*(rho R g1 + T)/z and K*g2   (util.hpp:188-209)
inline float reprojection_error_pixels(const Vec3& g1, const Vec3& g2, const Mat3& R, const Vec3& T, const float K[9], float rho)
{
  Vec3 q = mul(R, g1);
  for (int i = 0; i < 3; i++) q[i] = q[i] * rho + T[i];
  const float u = (q[0] / q[2]) * K[0] + K[2], v = (q[1] / q[2]) * K[4] + K[5];
  const float du = u - (g2[0] * K[0] + K[2]), dv = v - (g2[1] * K[4] + K[5]);
  return std::sqrt(du * du + dv * dv);
}
inline Mat3 skew(const Vec3& t) { return {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0}; }
inline Mat3 inverse3(const Mat3& M)
{
  const float id = 1.0f / det3(M);
  return {(M[4] * M[8] - M[7] * M[5]) * id, (M[2] * M[7] - M[1] * M[8]) * id, (M[1] * M[5] - M[2] * M[4]) * id,
          (M[5] * M[6] - M[3] * M[8]) * id, (M[0] * M[8] - M[2] * M[6]) * id, (M[3] * M[2] - M[0] * M[5]) * id,
          (M[3] * M[7] - M[6] * M[4]) * id, (M[6] * M[1] - M[0] * M[7]) * id, (M[0] * M[4] - M[3] * M[1]) * id};
}
// F = K^-T [T]x R K^-1   (util.hpp:217-229)
inline Mat3 fundamental_matrix(const float K[9], const Mat3& R, const Vec3& T)
{
  Mat3 Km; for (int i = 0; i < 9; i++) Km[i] = K[i];
  const Mat3 Ki = inverse3(Km);
  return matmul(matmul(transposed(Ki), matmul(skew(T), R)), Ki);
}
}}  // namespace hcb200::mvg
#endif
