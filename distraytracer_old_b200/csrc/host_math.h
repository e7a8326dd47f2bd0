// Host-side double-precision 3-vector / 4x4 helpers for the scene flattener.
// Arithmetic order follows the reference so that flattened CTMs, boxes and centroids carry the same bits
// the Java objects would: myVector.java:27-43 (vector), :76-90 (mat x mat, mat x vert),
// :111-196 (cofactor inverse, |det| > 1e-7 guard), DistRayTracer.java:336-349 (axis rotation).
#pragma once
#include <cmath>
#include <cstring>
#include <limits>

namespace drt {

static const double kEps = .0000001;                         // DistRayTracer.java:53
static const double kDMax = std::numeric_limits<double>::max();
static const double kPiF = (double)3.14159265358979323846f;  // PConstants.PI (float)
static const double kTwoPiF = (double)6.28318530717958647693f;
static const double kDegToRadF = (double)0.017453292519943295f;

struct V3 { double x, y, z; };
inline V3 v3(double x, double y, double z) { V3 r = {x, y, z}; return r; }
inline double vdot(V3 a, V3 b) { return ((a.x * b.x) + (a.y * b.y) + (a.z * b.z)); }
inline V3 vcross(V3 a, V3 b) { return v3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)); }
inline double vmag(V3 a) { return std::sqrt(((a.x * a.x) + (a.y * a.y) + (a.z * a.z))); }
inline V3 vnorm(V3 a) { double m = vmag(a); if (m == 0) return a; return v3(a.x / m, a.y / m, a.z / m); }
inline V3 vnormOrZero(V3 a) { double m = vmag(a); if (m == 0) return v3(0, 0, 0); return v3(a.x / m, a.y / m, a.z / m); }
inline V3 vscale(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 vadd(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }

struct M4 {
  double a[16];   // row major
  static M4 ident() { M4 r; for (int i = 0; i < 16; ++i) r.a[i] = (i % 5 == 0) ? 1.0 : 0.0; return r; }
  double& at(int r, int c) { return a[4 * r + c]; }
  double at(int r, int c) const { return a[4 * r + c]; }
  bool same(const M4& o) const { return std::memcmp(a, o.a, sizeof(a)) == 0; }
};
inline M4 mmul(const M4& A, const M4& B) {
  M4 R;
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { double acc = 0; for (int k = 0; k < 4; ++k) acc += A.at(r, k) * B.at(k, c); R.at(r, c) = acc; }
  return R;
}
inline M4 mtranspose(const M4& A) { M4 R; for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) R.at(c, r) = A.at(r, c); return R; }
inline V3 mpoint(const M4& A, V3 p) {
  double b[4] = {p.x, p.y, p.z, 1}, o[3];
  for (int r = 0; r < 3; ++r) { double acc = 0; for (int c = 0; c < 4; ++c) acc += A.at(r, c) * b[c]; o[r] = acc; }
  return v3(o[0], o[1], o[2]);
}
inline V3 mvector(const M4& A, V3 p) {
  double b[4] = {p.x, p.y, p.z, 0}, o[3];
  for (int r = 0; r < 3; ++r) { double acc = 0; for (int c = 0; c < 4; ++c) acc += A.at(r, c) * b[c]; o[r] = acc; }
  return v3(o[0], o[1], o[2]);
}
// 4x4 inverse by cofactors of the transposed source; identity when |det| <= 1e-7 (the reference prints and keeps identity)
inline M4 minverse(const M4& A) {
  double s[16], p[12], d[16];
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) s[4 * c + r] = A.at(r, c);
  p[0] = s[10] * s[15]; p[1] = s[11] * s[14]; p[2] = s[9] * s[15]; p[3] = s[11] * s[13]; p[4] = s[9] * s[14]; p[5] = s[10] * s[13];
  p[6] = s[8] * s[15]; p[7] = s[11] * s[12]; p[8] = s[8] * s[14]; p[9] = s[10] * s[12]; p[10] = s[8] * s[13]; p[11] = s[9] * s[12];
  d[0] = p[0] * s[5] + p[3] * s[6] + p[4] * s[7]; d[0] -= p[1] * s[5] + p[2] * s[6] + p[5] * s[7];
  d[1] = p[1] * s[4] + p[6] * s[6] + p[9] * s[7]; d[1] -= p[0] * s[4] + p[7] * s[6] + p[8] * s[7];
  d[2] = p[2] * s[4] + p[7] * s[5] + p[10] * s[7]; d[2] -= p[3] * s[4] + p[6] * s[5] + p[11] * s[7];
  d[3] = p[5] * s[4] + p[8] * s[5] + p[11] * s[6]; d[3] -= p[4] * s[4] + p[9] * s[5] + p[10] * s[6];
  d[4] = p[1] * s[1] + p[2] * s[2] + p[5] * s[3]; d[4] -= p[0] * s[1] + p[3] * s[2] + p[4] * s[3];
  d[5] = p[0] * s[0] + p[7] * s[2] + p[8] * s[3]; d[5] -= p[1] * s[0] + p[6] * s[2] + p[9] * s[3];
  d[6] = p[3] * s[0] + p[6] * s[1] + p[11] * s[3]; d[6] -= p[2] * s[0] + p[7] * s[1] + p[10] * s[3];
  d[7] = p[4] * s[0] + p[9] * s[1] + p[10] * s[2]; d[7] -= p[5] * s[0] + p[8] * s[1] + p[11] * s[2];
  p[0] = s[2] * s[7]; p[1] = s[3] * s[6]; p[2] = s[1] * s[7]; p[3] = s[3] * s[5]; p[4] = s[1] * s[6]; p[5] = s[2] * s[5];
  p[6] = s[0] * s[7]; p[7] = s[3] * s[4]; p[8] = s[0] * s[6]; p[9] = s[2] * s[4]; p[10] = s[0] * s[5]; p[11] = s[1] * s[4];
  d[8] = p[0] * s[13] + p[3] * s[14] + p[4] * s[15]; d[8] -= p[1] * s[13] + p[2] * s[14] + p[5] * s[15];
  d[9] = p[1] * s[12] + p[6] * s[14] + p[9] * s[15]; d[9] -= p[0] * s[12] + p[7] * s[14] + p[8] * s[15];
  d[10] = p[2] * s[12] + p[7] * s[13] + p[10] * s[15]; d[10] -= p[3] * s[12] + p[6] * s[13] + p[11] * s[15];
  d[11] = p[5] * s[12] + p[8] * s[13] + p[11] * s[14]; d[11] -= p[4] * s[12] + p[9] * s[13] + p[10] * s[14];
  d[12] = p[2] * s[10] + p[5] * s[11] + p[1] * s[9]; d[12] -= p[4] * s[11] + p[0] * s[9] + p[3] * s[10];
  d[13] = p[8] * s[11] + p[0] * s[8] + p[7] * s[10]; d[13] -= p[6] * s[10] + p[9] * s[11] + p[1] * s[8];
  d[14] = p[6] * s[9] + p[11] * s[11] + p[3] * s[8]; d[14] -= p[10] * s[11] + p[2] * s[8] + p[7] * s[9];
  d[15] = p[10] * s[10] + p[4] * s[8] + p[9] * s[9]; d[15] -= p[8] * s[9] + p[11] * s[10] + p[5] * s[8];
  double det = s[0] * d[0] + s[1] * d[1] + s[2] * d[2] + s[3] * d[3];
  M4 R = M4::ident();
  if (std::fabs(det) > .0000001) for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) R.at(r, c) = d[4 * r + c] / det;
  return R;
}
inline V3 rotateAboutAxis(V3 v, V3 u, double th) {           // DistRayTracer.java:336-349
  double c = std::cos(th), s = std::sin(th), omc = 1 - c;
  double ux2 = u.x * u.x, uy2 = u.y * u.y, uz2 = u.z * u.z, uxy = u.x * u.y, uxz = u.x * u.z, uyz = u.y * u.z;
  double uzS = u.z * s, uyS = u.y * s, uxS = u.x * s, uxzC = uxz * omc, uxyC = uxy * omc, uyzC = uyz * omc;
  return v3((ux2 * omc + c) * v.x + (uxyC - uzS) * v.y + (uxzC + uyS) * v.z,
            (uxyC + uzS) * v.x + (uy2 * omc + c) * v.y + (uyzC - uxS) * v.z,
            (uxzC - uyS) * v.x + (uyzC + uxS) * v.y + (uz2 * omc + c) * v.z);
}
inline V3 orthoVec(V3 v) {                                     // DistRayTracer.java:455-462
  V3 t = vnorm(v3(1, 1, 0));
  if (std::fabs(vdot(t, v) - 1) < kEps) t = v3(0, 0, 1);
  return vnorm(vcross(v, t));
}
// Java Math.min/max (NaN propagating) and Double.compare ordering
inline double jmin(double a, double b) { if (a != a) return a; if (b != b) return b; return a < b ? a : b; }
inline double jmax(double a, double b) { if (a != a) return a; if (b != b) return b; return a > b ? a : b; }
inline int javaDoubleCompare(double a, double b) {
  if (a < b) return -1; if (a > b) return 1;
  long long x, y; std::memcpy(&x, &a, 8); std::memcpy(&y, &b, 8);
  return x == y ? 0 : (x < y ? -1 : 1);
}

}  // namespace drt
