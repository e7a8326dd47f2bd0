// ORACLE (test infrastructure only -- never linked into the product path).
// CPU restatement of the reference's math layer, in plain C++17 / IEEE double.
// PARITY UNPINNED: the reference ships no tests or golden vectors and no JVM
// exists in the build container, so this restatement is pinned only by the
// reference source it cites (see DESIGN.md "Oracle pinning").
//
// Follows (reference file:line, relative to /root/reference/src/rayTracerDistAccelShdPhtnMap/):
//   myVector.java:7-63     3-vector
//   myVector.java:65-223   4x4 matrix, cofactor inverse with |det| > 1e-7 guard
//   myVector.java:225-256  matrix stack (only 10 of 20 slots exist)
//   DistRayTracer.java:336-349 rotVecAroundAxis, :387-405 CTM array builders, :455-462 getOrthoVec
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

static const double EPS = .0000001;                       // DistRayTracer.java:53
static const double DMAX = std::numeric_limits<double>::max();
// Processing float constants widened to double (PConstants.PI / TWO_PI / DEG_TO_RAD)
static const double PI_F = (double)3.14159265358979323846f;
static const double TWO_PI_F = (double)6.28318530717958647693f;
static const double DEG_TO_RAD_F = (double)0.017453292519943295f;

// Java Math.min/max propagate NaN; (int) casts saturate.
inline double jmin(double a, double b) { if (a != a) return a; if (b != b) return b; return a < b ? a : b; }
inline double jmax(double a, double b) { if (a != a) return a; if (b != b) return b; return a > b ? a : b; }
inline int j2i(double d) {
  if (d != d) return 0;
  if (d >= 2147483647.0) return 2147483647;
  if (d <= -2147483648.0) return (int)0x80000000;
  return (int)d;
}
inline int fastfloor(double x) { return x > 0 ? j2i(x) : j2i(x) - 1; }   // DistRayTracer.java:307
inline int fastfloorf(float x) { return x > 0 ? j2i((double)x) : j2i((double)x) - 1; }  // :306
// Double.compare total order (-0.0 < +0.0), used by TreeMap<Double,..> keys.
inline int dcompare(double a, double b) {
  if (a < b) return -1;
  if (a > b) return 1;
  uint64_t ba, bb; std::memcpy(&ba, &a, 8); std::memcpy(&bb, &b, 8);
  int64_t sa = (int64_t)ba, sb = (int64_t)bb;
  return sa == sb ? 0 : (sa < sb ? -1 : 1);
}

struct Vec3 {
  double x = 0, y = 0, z = 0;
  Vec3() {}
  Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
  Vec3(const Vec3& p, const Vec3& q) : x(q.x - p.x), y(q.y - p.y), z(q.z - p.z) {}  // vector p -> q
  double sqMag() const { return ((x * x) + (y * y) + (z * z)); }
  double mag() const { return std::sqrt(sqMag()); }
  void set(double a, double b, double c) { x = a; y = b; z = c; }
  void mult(double n) { x *= n; y *= n; z *= n; }
  void add(double a, double b, double c) { x += a; y += b; z += c; }
  void add(const Vec3& v) { x += v.x; y += v.y; z += v.z; }
  void sub(const Vec3& v) { x -= v.x; y -= v.y; z -= v.z; }
  void div(double q) { x /= q; y /= q; z /= q; }
  void normalize() { double m = mag(); if (m == 0) return; div(m); }     // myVector.java:30
  Vec3 normalized() const { double m = mag(); return m == 0 ? Vec3(0, 0, 0) : Vec3(x / m, y / m, z / m); }
  Vec3 cross(const Vec3& b) const { return Vec3((y * b.z) - (z * b.y), (z * b.x) - (x * b.z), (x * b.y) - (y * b.x)); }
  double dot(const Vec3& b) const { return ((x * b.x) + (y * b.y) + (z * b.z)); }
  double dist(const Vec3& q) const { return std::sqrt(((x - q.x) * (x - q.x)) + ((y - q.y) * (y - q.y)) + ((z - q.z) * (z - q.z))); }
  double L1Dist(const Vec3& q) const { return std::fabs(x - q.x) + std::fabs(y - q.y) + std::fabs(z - q.z); }
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 vsub(const Vec3& p, const Vec3& q) { return Vec3(p.x - q.x, p.y - q.y, p.z - q.z); }
inline Vec3 velemMult(const Vec3& p, const Vec3& q) { return Vec3(p.x * q.x, p.y * q.y, p.z * q.z); }

struct Mat4 {
  double m[4][4];
  Mat4() { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m[r][c] = (r == c) ? 1 : 0; }
  Mat4 multMat(const Mat4& b) const {                                   // myVector.java:76-81
    Mat4 res;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
      double v = 0; for (int k = 0; k < 4; ++k) v += m[r][k] * b.m[k][c];
      res.m[r][c] = v;
    }
    return res;
  }
  void multVert(const double b[4], double out[4]) const {               // :85-90
    for (int r = 0; r < 4; ++r) { double v = 0; for (int c = 0; c < 4; ++c) v += m[r][c] * b[c]; out[r] = v; }
  }
  Vec3 xfPt(const Vec3& p) const { double b[4] = {p.x, p.y, p.z, 1}, o[4]; multVert(b, o); return Vec3(o[0], o[1], o[2]); }
  Vec3 xfVec(const Vec3& p) const { double b[4] = {p.x, p.y, p.z, 0}, o[4]; multVert(b, o); return Vec3(o[0], o[1], o[2]); }
  Mat4 transpose() const { Mat4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[j][i] = m[i][j]; return r; }
  // cofactor expansion, same pairing/summation order as myVector.java:111-196
  Mat4 inverse() const {
    double t[12], s[16], d[16];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) s[(4 * c) + r] = m[r][c];
    t[0] = s[10] * s[15]; t[1] = s[11] * s[14]; t[2] = s[9] * s[15]; t[3] = s[11] * s[13];
    t[4] = s[9] * s[14]; t[5] = s[10] * s[13]; t[6] = s[8] * s[15]; t[7] = s[11] * s[12];
    t[8] = s[8] * s[14]; t[9] = s[10] * s[12]; t[10] = s[8] * s[13]; t[11] = s[9] * s[12];
    d[0] = t[0] * s[5] + t[3] * s[6] + t[4] * s[7];   d[0] -= t[1] * s[5] + t[2] * s[6] + t[5] * s[7];
    d[1] = t[1] * s[4] + t[6] * s[6] + t[9] * s[7];   d[1] -= t[0] * s[4] + t[7] * s[6] + t[8] * s[7];
    d[2] = t[2] * s[4] + t[7] * s[5] + t[10] * s[7];  d[2] -= t[3] * s[4] + t[6] * s[5] + t[11] * s[7];
    d[3] = t[5] * s[4] + t[8] * s[5] + t[11] * s[6];  d[3] -= t[4] * s[4] + t[9] * s[5] + t[10] * s[6];
    d[4] = t[1] * s[1] + t[2] * s[2] + t[5] * s[3];   d[4] -= t[0] * s[1] + t[3] * s[2] + t[4] * s[3];
    d[5] = t[0] * s[0] + t[7] * s[2] + t[8] * s[3];   d[5] -= t[1] * s[0] + t[6] * s[2] + t[9] * s[3];
    d[6] = t[3] * s[0] + t[6] * s[1] + t[11] * s[3];  d[6] -= t[2] * s[0] + t[7] * s[1] + t[10] * s[3];
    d[7] = t[4] * s[0] + t[9] * s[1] + t[10] * s[2];  d[7] -= t[5] * s[0] + t[8] * s[1] + t[11] * s[2];
    t[0] = s[2] * s[7]; t[1] = s[3] * s[6]; t[2] = s[1] * s[7]; t[3] = s[3] * s[5];
    t[4] = s[1] * s[6]; t[5] = s[2] * s[5]; t[6] = s[0] * s[7]; t[7] = s[3] * s[4];
    t[8] = s[0] * s[6]; t[9] = s[2] * s[4]; t[10] = s[0] * s[5]; t[11] = s[1] * s[4];
    d[8] = t[0] * s[13] + t[3] * s[14] + t[4] * s[15];    d[8] -= t[1] * s[13] + t[2] * s[14] + t[5] * s[15];
    d[9] = t[1] * s[12] + t[6] * s[14] + t[9] * s[15];    d[9] -= t[0] * s[12] + t[7] * s[14] + t[8] * s[15];
    d[10] = t[2] * s[12] + t[7] * s[13] + t[10] * s[15];  d[10] -= t[3] * s[12] + t[6] * s[13] + t[11] * s[15];
    d[11] = t[5] * s[12] + t[8] * s[13] + t[11] * s[14];  d[11] -= t[4] * s[12] + t[9] * s[13] + t[10] * s[14];
    d[12] = t[2] * s[10] + t[5] * s[11] + t[1] * s[9];    d[12] -= t[4] * s[11] + t[0] * s[9] + t[3] * s[10];
    d[13] = t[8] * s[11] + t[0] * s[8] + t[7] * s[10];    d[13] -= t[6] * s[10] + t[9] * s[11] + t[1] * s[8];
    d[14] = t[6] * s[9] + t[11] * s[11] + t[3] * s[8];    d[14] -= t[10] * s[11] + t[2] * s[8] + t[7] * s[9];
    d[15] = t[10] * s[10] + t[4] * s[8] + t[9] * s[9];    d[15] -= t[8] * s[9] + t[11] * s[10] + t[5] * s[8];
    double det = s[0] * d[0] + s[1] * d[1] + s[2] * d[2] + s[3] * d[3];
    Mat4 out;   // identity when singular (reference prints and returns identity, :192-195)
    if (std::fabs(det) > .0000001) {
      for (int j = 0; j < 16; ++j) d[j] /= det;
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out.m[r][c] = d[(4 * r) + c];
    }
    return out;
  }
};

// {M, M^-1, M^T, M^-T}  (DistRayTracer.java:399)
struct CTM {
  Mat4 glbl, inv, trans, adj;
  static CTM* build(const Mat4& m) { CTM* c = new CTM; c->glbl = m; c->inv = m.inverse(); c->trans = m.transpose(); c->adj = c->inv.transpose(); return c; }
};

// DistRayTracer.java:336-349
inline Vec3 rotVecAroundAxis(const Vec3& v1, const Vec3& u, double thet) {
  double cT = std::cos(thet), sT = std::sin(thet), oneMC = 1 - cT,
         ux2 = u.x * u.x, uy2 = u.y * u.y, uz2 = u.z * u.z,
         uxy = u.x * u.y, uxz = u.x * u.z, uyz = u.y * u.z,
         uzS = u.z * sT, uyS = u.y * sT, uxS = u.x * sT,
         uxzC1 = uxz * oneMC, uxyC1 = uxy * oneMC, uyzC1 = uyz * oneMC;
  return Vec3((ux2 * oneMC + cT) * v1.x + (uxyC1 - uzS) * v1.y + (uxzC1 + uyS) * v1.z,
              (uxyC1 + uzS) * v1.x + (uy2 * oneMC + cT) * v1.y + (uyzC1 - uxS) * v1.z,
              (uxzC1 - uyS) * v1.x + (uyzC1 + uxS) * v1.y + (uz2 * oneMC + cT) * v1.z);
}
// DistRayTracer.java:455-462
inline Vec3 getOrthoVec(const Vec3& vec) {
  Vec3 tmp(1, 1, 0); tmp.normalize();
  if (std::fabs(tmp.dot(vec) - 1) < EPS) tmp.set(0, 0, 1);
  Vec3 r = vec.cross(tmp); r.normalize(); return r;
}

// ---------------------------------------------------------------------------
// Seeded counter-based sampler replacing the reference's unseeded
// ThreadLocalRandom.  Philox4x32-10; the product's device code
// implements the same function independently and tests assert bit equality.
// counter = (a, b, c, d), key = (seed_lo, seed_hi ^ stream)
// ---------------------------------------------------------------------------
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1,
             n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
enum : uint32_t { STREAM_PIXEL = 0x0u, STREAM_PHOTON = 0x50484F54u /*'PHOT'*/ };
// uniform double in [0,1) with 53 random bits
inline double u01(uint64_t seed, uint32_t stream, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  uint32_t ctr[4] = {a, b, c, d};
  philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32) ^ stream);
  uint64_t hi = ctr[0] >> 5, lo = ctr[1] >> 6;          // 27 + 26 bits
  return (double)((hi << 26) | lo) * (1.0 / 9007199254740992.0);
}
// ThreadLocalRandom.nextDouble(origin, bound) semantics
inline double urange(double u, double origin, double bound) {
  double r = u * (bound - origin) + origin;
  if (r >= bound) r = std::nextafter(bound, -DMAX);
  return r;
}
// sample dimensions (pixel stream): counter = (pixel, sample, path, dim)
enum : uint32_t {
  DIM_AA_Y = 0, DIM_AA_X = 1, DIM_TIME = 2, DIM_LENS_ANGLE = 3, DIM_LENS_RADIUS = 4,
  DIM_LIGHT_BASE = 16, DIM_LIGHT_STRIDE = 8   // +0,1: disk sample A (angle, radius); +2,3: sample B; +4: shadow-ray time
};

// java.util.Random (48-bit LCG), bit exact -- used by the Worley texture (myTextureHandler.java:435-479)
struct JavaRandom {
  uint64_t seed = 0;
  void setSeed(int64_t s) { seed = ((uint64_t)s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
  int next(int bits) { seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1); return (int)(int64_t)(seed >> (48 - bits)); }
  double nextDouble() { int64_t a = next(26); int64_t b = next(27); return (double)((a << 27) + b) * (1.0 / 9007199254740992.0); }
};

}  // namespace orc
