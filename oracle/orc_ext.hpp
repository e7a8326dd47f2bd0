// ORACLE (test infrastructure only -- never linked into the product path): extension primitives.
//
// BASELINE's north star lists "quadrics, tori" among the intersection kernels, but the reference has neither: no reader command, a myTorus class
// whose intersectCheck always misses (myImpObject.java:330-390) and a page of unfinished quadric algebra (src/tmpQuadricEQ.txt).  The product
// therefore ignores `torus` / `quadric` lines exactly as the reference does, unless the scene says `extensions on`; the semantics of the two
// primitives are THIS PROJECT'S OWN and there is nothing in the reference to pin them to -- PARITY UNPINNED.  This file is the CPU twin the GPU
// path is compared with (same operations in the same order, FP64, no contraction), plus what tests/test_oracle.py checks analytically:
//   quadric a..j [clip box]: a x^2 + b y^2 + c z^2 + d xy + e xz + f yz + g x + h y + i z + j = 0, nearest root > eps whose point lies in the box
//   torus R r [facets] x y z: axis = object-space y; smallest root > eps of the quartic inside the ray's span of the bounding sphere (radius 1.001 (R + r): a root ON the tight sphere would sit on the scan's first sample), isolated by
//   a fixed 64-step scan for the first sign change and 64 bisection steps (a grazing double root between two samples is missed -- stated)
#pragma once

namespace orc {

struct Quadric : ImpObject {
  double q[10], bx[6];
  Quadric(Scene* s, const double* coef, const double* box) : ImpObject(s, (box[0] + box[3]) * .5, (box[1] + box[4]) * .5, (box[2] + box[5]) * .5) {
    for (int i = 0; i < 10; ++i) q[i] = coef[i]; for (int i = 0; i < 6; ++i) bx[i] = box[i];
    type = G_QUADRIC; minVals = getMinVec(); maxVals = getMaxVec(); postProcBBox();
  }
  Vec3 getMaxVec() override { return Vec3(bx[3], bx[4], bx[5]); }
  Vec3 getMinVec() override { return Vec3(bx[0], bx[1], bx[2]); }
  Vec3 getNormalAtPoint(const Vec3& pt, const int* args) override {
    Vec3 n((((2 * q[0]) * pt.x) + (q[3] * pt.y)) + ((q[4] * pt.z) + q[6]), (((2 * q[1]) * pt.y) + (q[3] * pt.x)) + ((q[5] * pt.z) + q[7]), (((2 * q[2]) * pt.z) + (q[4] * pt.x)) + ((q[5] * pt.y) + q[8]));
    n.normalize(); if (args && args[0] == 1) n.mult(-1); if (inverted) n.mult(-1); return n;
  }
  RayHit intersectCheck(Ray& _ray, Ray& tr, CTM* ct) override {
    ++scene->stats.primTests;
    const double a = q[0], b = q[1], c = q[2], d = q[3], e = q[4], f = q[5], g = q[6], hh = q[7], ii = q[8], j = q[9];
    const double ox = tr.origin.x, oy = tr.origin.y, oz = tr.origin.z, dx = tr.direction.x, dy = tr.direction.y, dz = tr.direction.z;
    const double A = (((((a * dx) * dx) + ((b * dy) * dy)) + ((c * dz) * dz)) + ((d * dx) * dy)) + (((e * dx) * dz) + ((f * dy) * dz));
    const double B = ((2 * ((((a * ox) * dx) + ((b * oy) * dy)) + ((c * oz) * dz))) + (((d * ((ox * dy) + (oy * dx))) + (e * ((ox * dz) + (oz * dx)))) + (f * ((oy * dz) + (oz * dy))))) + (((g * dx) + (hh * dy)) + (ii * dz));
    const double C = ((((((a * ox) * ox) + ((b * oy) * oy)) + ((c * oz) * oz)) + ((d * ox) * oy)) + (((e * ox) * oz) + ((f * oy) * oz))) + ((((g * ox) + (hh * oy)) + (ii * oz)) + j);
    double t0, t1; int nr;
    if (A == 0) { if (B == 0) return RayHit(); t0 = -C / B; t1 = t0; nr = 1; }
    else {
      const double disc = (B * B) - ((4 * A) * C); if (disc < 0) return RayHit();
      const double sq = std::sqrt(disc), ta = (-B - sq) / (2 * A), tb = (-B + sq) / (2 * A); t0 = jmin(ta, tb); t1 = jmax(ta, tb); nr = 2;
    }
    for (int k = 0; k < nr; ++k) {
      const double t = k == 0 ? t0 : t1; if (!(t > EPS)) continue;
      const double px = (dx * t) + ox, py = (dy * t) + oy, pz = (dz * t) + oz;
      if (px < bx[0] || py < bx[1] || pz < bx[2] || px > bx[3] || py > bx[4] || pz > bx[5]) continue;
      int args[2] = {k, 0}; return objHit(tr, _ray.direction, ct, tr.pointOnRay(t), args, t);
    }
    return RayHit();
  }
};

struct Torus : ImpObject {
  double R, rr;
  Torus(Scene* s, double bodyRad, double ringRad, double x, double y, double z) : ImpObject(s, x, y, z), R(bodyRad), rr(ringRad) { type = G_TORUS; minVals = getMinVec(); maxVals = getMaxVec(); postProcBBox(); }
  Vec3 getMaxVec() override { Vec3 r(origin); const double e = R + rr; r.add(e, rr, e); return r; }
  Vec3 getMinVec() override { Vec3 r(origin); const double e = R + rr; r.add(-e, -rr, -e); return r; }
  Vec3 getNormalAtPoint(const Vec3& pt, const int*) override {
    const double x = pt.x - origin.x, y = pt.y - origin.y, z = pt.z - origin.z, m = std::sqrt((x * x) + (z * z));
    Vec3 n = (m > 0) ? Vec3(x - ((R * x) / m), y, z - ((R * z) / m)) : Vec3(0, y, 0);
    n.normalize(); if (inverted) n.mult(-1); return n;
  }
  RayHit intersectCheck(Ray& _ray, Ray& tr, CTM* ct) override {
    ++scene->stats.primTests;
    const double ox = tr.origin.x - origin.x, oy = tr.origin.y - origin.y, oz = tr.origin.z - origin.z, dx = tr.direction.x, dy = tr.direction.y, dz = tr.direction.z;
    const double al = ((dx * dx) + (dy * dy)) + (dz * dz), be = ((ox * dx) + (oy * dy)) + (oz * dz), oo = ((ox * ox) + (oy * oy)) + (oz * oz);
    if (!(al > 0)) return RayHit();
    const double Rb = (R + rr) * 1.001, dsc = (be * be) - (al * (oo - (Rb * Rb))); if (dsc < 0) return RayHit();
    const double sq = std::sqrt(dsc); double ta = (-be - sq) / al, tb = (-be + sq) / al;
    if (!(tb > EPS)) return RayHit(); if (ta < EPS) ta = EPS;
    const double kk = oo + ((R * R) - (rr * rr)), R4 = 4 * (R * R);
    const double c4 = al * al, c3 = 4 * (al * be), c2 = ((2 * (al * kk)) + (4 * (be * be))) - (R4 * ((dx * dx) + (dz * dz))), c1 = (4 * (be * kk)) - ((2 * R4) * ((ox * dx) + (oz * dz))), c0 = (kk * kk) - (R4 * ((ox * ox) + (oz * oz)));
    auto F = [&](double t) { return ((((((c4 * t) + c3) * t) + c2) * t + c1) * t) + c0; };
    const double step = (tb - ta) / 64; double lo = ta, flo = F(lo); bool found = false; double hi = ta;
    for (int k = 1; k <= 64; ++k) { hi = (k == 64) ? tb : ta + (step * k); const double fhi = F(hi); if ((flo > 0) != (fhi > 0)) { found = true; break; } lo = hi; flo = fhi; }
    if (!found) return RayHit();
    for (int k = 0; k < 64; ++k) { const double mid = 0.5 * (lo + hi), fm = F(mid); if ((fm > 0) == (flo > 0)) { lo = mid; flo = fm; } else hi = mid; }
    const double t = 0.5 * (lo + hi); if (!(t > EPS)) return RayHit();
    int args[2] = {0, 0}; return objHit(tr, _ray.direction, ct, tr.pointOnRay(t), args, t);
  }
};

}  // namespace orc
