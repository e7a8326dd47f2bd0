// ORACLE (test infrastructure only -- never linked into the product path).
// CPU restatement of the reference's ray / hit records, geometry, acceleration
// structures and lights.  PARITY UNPINNED (no reference tests / no JVM), see orc_math.hpp.
//
// Follows (relative to /root/reference/src/rayTracerDistAccelShdPhtnMap/):
//   myRay.java:7-127 (myRay), :130-186 (rayHit)
//   myGeomBase.java:10-87 (base), :90-197 (myBBox), :200-249 (myAccelStruct),
//                   :251-306 (myGeomList), :309-423 (myBVH)
//   mySceneObject.java:5-57, :59-92 (myRndrdBox), :95-145 (myInstance)
//   myPlanarObject.java:7-289 ; myImpObject.java:7-327 ; myLight.java:12-275
//   DistRayTracer.java:353-380 (box growth), :409-418 (max centroid span)
#pragma once
#include "orc_math.hpp"
#include <algorithm>
#include <map>
#include <memory>

namespace orc {

struct Scene; struct Geom; struct Shader; struct Image;

struct Stats {
  uint64_t primary = 0, shadow = 0, reflect = 0, refract = 0, photonSeg = 0;  // logical rays
  uint64_t boxTests = 0, primTests = 0;
  uint64_t boxTestsPrimary = 0, primTestsPrimary = 0;
  uint64_t photonsStored = 0;
  void add(const Stats& o) {
    primary += o.primary; shadow += o.shadow; reflect += o.reflect; refract += o.refract; photonSeg += o.photonSeg;
    boxTests += o.boxTests; primTests += o.primTests; boxTestsPrimary += o.boxTestsPrimary; primTestsPrimary += o.primTestsPrimary;
    photonsStored += o.photonsStored;
  }
};

// sampler key carried by every ray: counter = (a, b, c, dim)
struct SKey { uint32_t stream = STREAM_PIXEL, a = 0, b = 0, c = 1, timeDim = DIM_TIME; };

struct Ray {                                                      // myRay.java:7-127
  Scene* scn;
  double currKTrans[5];
  Vec3 origin, direction;
  double originAra[3], dirAra[3];
  int gen;
  SKey key;
  bool dirNormalized;       // bookkeeping for the canonical (non-literal) re-normalisation mode
  Ray() : scn(nullptr), gen(0), dirNormalized(false) {}
  Ray(Scene* s, const Vec3& o, const Vec3& d, int g, const SKey& k) : scn(s), origin(o), direction(d), gen(g), key(k) {
    for (int i = 0; i < 5; ++i) currKTrans[i] = 1;
    direction.normalize();
    originAra[0] = origin.x; originAra[1] = origin.y; originAra[2] = origin.z;
    dirAra[0] = direction.x; dirAra[1] = direction.y; dirAra[2] = direction.z;
    dirNormalized = true;
  }
  double getTime() const;                                          // :49-52 (keyed, see orc_math.hpp)
  void setCurrKTrans(double kt, double perm, const Vec3& pc) { currKTrans[0] = kt; currKTrans[1] = perm; currKTrans[2] = pc.x; currKTrans[3] = pc.y; currKTrans[4] = pc.z; }
  Vec3 pointOnRay(double t) const { Vec3 r(direction); r.mult(t); r.add(origin); return r; }   // :82-87
  Ray getTransformedRay(Ray& src, const Mat4& trans) const;        // :91-102
};

struct RayHit {                                                   // myRay.java:130-186
  bool isHit = false;
  double t = DMAX, ltMult = 1;
  Geom* obj = nullptr;
  Shader* shdr = nullptr;
  CTM* ctm = nullptr;
  Vec3 objNorm, hitLoc, fwdTransHitLoc, fwdTransRayDir;
  int args[2] = {0, 0};
  // copy of the owning (transformed) ray's state used by shading
  int gen = 0; double rayKTrans[5] = {1, 1, 1, 1, 1}; SKey key;
  double phtnPwr[3] = {0, 0, 0};
  int instSerial = -1;     // AOV: innermost instance the hit came through (oracle bookkeeping only)
  void reCalcCTMHitNorm(CTM* c);                                   // :168-175
};

struct BBox {                                                     // myGeomBase.java:90-197 (arithmetic only)
  Vec3 minVals{100000, 100000, 100000}, maxVals{-100000, -100000, -100000};
  void calcMinMax(const Vec3& mn, const Vec3& mx) {                // :102-117
    minVals.set(jmin(mn.x, minVals.x), jmin(mn.y, minVals.y), jmin(mn.z, minVals.z));
    maxVals.set(jmax(mx.x, maxVals.x), jmax(mx.y, maxVals.y), jmax(mx.z, maxVals.z));
  }
  void expandPt(const Vec3& p) {                                   // DistRayTracer.java:353-361
    minVals.x = (minVals.x < p.x) ? minVals.x : p.x; minVals.y = (minVals.y < p.y) ? minVals.y : p.y; minVals.z = (minVals.z < p.z) ? minVals.z : p.z;
    maxVals.x = (maxVals.x > p.x) ? maxVals.x : p.x; maxVals.y = (maxVals.y > p.y) ? maxVals.y : p.y; maxVals.z = (maxVals.z > p.z) ? maxVals.z : p.z;
  }
  void expandByBox(const BBox& s, const Mat4& fwd) { expandPt(fwd.xfPt(s.minVals)); expandPt(fwd.xfPt(s.maxVals)); }  // :364-367
  void expandByBox(const BBox& s) { expandPt(s.minVals); expandPt(s.maxVals); }                                        // :368-371
  // slab test with the "entry t must be > 0" rule, myGeomBase.java:132-162
  bool test(const Ray& tr, double& tOut, int& idxOut, Stats& st) const {
    ++st.boxTests;
    const double* rayO = tr.originAra; const double* rayD = tr.dirAra;
    double mn[3] = {minVals.x, minVals.y, minVals.z}, mx[3] = {maxVals.x, maxVals.y, maxVals.z};
    double v1[3], v2[3], tMin[3], tMax[3];
    double biggestMin = -DMAX; int idx = -1;
    for (int i = 0; i < 3; ++i) { v1[i] = (mn[i] - rayO[i]) / rayD[i]; v2[i] = (mx[i] - rayO[i]) / rayD[i]; }
    for (int i = 0; i < 3; ++i) {
      if (v1[i] < v2[i]) { tMin[i] = v1[i]; tMax[i] = v2[i]; if (biggestMin < v1[i]) { idx = i; biggestMin = v1[i]; } }
      else { tMin[i] = v2[i]; tMax[i] = v1[i]; if (biggestMin < v2[i]) { idx = i + 3; biggestMin = v2[i]; } }
    }
    // DistRayTracer.min/max: running comparison, NaN never replaces (:424-425)
    double minMax = DMAX; for (int i = 0; i < 3; ++i) if (tMax[i] < minMax) minMax = tMax[i];
    double maxMin = -DMAX; for (int i = 0; i < 3; ++i) if (tMin[i] > maxMin) maxMin = tMin[i];
    if ((minMax > maxMin) && biggestMin > 0) { tOut = biggestMin; idxOut = idx; return true; }
    return false;
  }
  static Vec3 faceNormal(int idx) {                                // :175-186
    switch (idx) { case 0: return Vec3(-1, 0, 0); case 1: return Vec3(0, -1, 0); case 2: return Vec3(0, 0, -1);
      case 3: return Vec3(1, 0, 0); case 4: return Vec3(0, 1, 0); case 5: return Vec3(0, 0, 1); default: return Vec3(0, 0, -1); }
  }
};

enum GType { G_NONE, G_SPHERE, G_MOVSPHERE, G_HCYL, G_CYL, G_TRI, G_QUAD, G_PLANE, G_BOX, G_INSTANCE, G_LIST, G_BVH, G_POINTLIGHT, G_SPOTLIGHT, G_DISKLIGHT, G_TORUS, G_QUADRIC };

struct Geom {                                                     // myGeomBase.java:10-87
  Scene* scene; int ID; GType type = G_NONE;
  Vec3 origin; double trans_origin[3];
  Shader* shdr = nullptr;
  CTM* ctm;
  BBox bbox;                  // "_bbox" (every object has one)
  Vec3 minVals{100000, 100000, 100000}, maxVals{-100000, -100000, -100000};
  int primSerial = -1;        // AOV id (parse order over renderable primitives)
  bool inverted = false;      // rFlags[invertedIDX]
  Geom(Scene* s, double x, double y, double z);
  virtual ~Geom() {}
  void postProcBBox() { bbox = BBox(); bbox.calcMinMax(minVals, maxVals); }                  // :42-45
  void setTransOrigin() { Vec3 t = ctm->glbl.xfPt(origin); trans_origin[0] = t.x; trans_origin[1] = t.y; trans_origin[2] = t.z; }
  virtual Vec3 getOrigin(double) { return origin; }
  virtual Vec3 getMaxVec() = 0;
  virtual Vec3 getMinVec() = 0;
  virtual int calcShadowHit(Ray& _ray, Ray& trans, CTM* ct, double distToLight);             // mySceneObject.java:33-38
  virtual RayHit intersectCheck(Ray& _ray, Ray& trans, CTM* ct) = 0;
  virtual Vec3 getNormalAtPoint(const Vec3& pt, const int* args) = 0;
  virtual void findTxtrCoords(const Vec3& pt, const Image* tex, double time, double uv[2]) { (void)pt; (void)tex; (void)time; uv[0] = 0; uv[1] = 0; }
  virtual bool isAccel() const { return false; }
  virtual bool isLight() const { return false; }
  // myRay.objHit (myRay.java:119-125) + rayHit ctor (:147-161)
  RayHit objHit(const Ray& tr, const Vec3& rawRayDir, CTM* ct, const Vec3& pt, const int* args, double t);
};

// ---- renderable box (mySceneObject.java:59-92; myGeomBase.java:132-162)
struct RndrdBox : Geom {
  RndrdBox(Scene* s, double x, double y, double z, const Vec3& mn, const Vec3& mx) : Geom(s, x, y, z) { minVals = mn; maxVals = mx; type = G_BOX; postProcBBox(); }
  Vec3 getMaxVec() override { return bbox.maxVals; }
  Vec3 getMinVec() override { return bbox.minVals; }
  RayHit intersectCheck(Ray& _ray, Ray& tr, CTM* ct) override;
  Vec3 getNormalAtPoint(const Vec3&, const int* args) override { return BBox::faceNormal(args[1]); }
};

// ---- planar objects (myPlanarObject.java)
struct PolyState {             // one winding order of the polygon (reference mutates in place)
  double vx[4], vy[4], vz[4], vu[4], vv[4];
  Vec3 P[4], P2P[4], P2P0, N;
  double dotVals[5], baryIDenomTxtr, peqA, peqB, peqC, peqD;
};
struct Planar : Geom {
  int vCount;
  PolyState st[2]; int cur = 0;       // st[0] = as loaded, st[1] = reversed
  bool isInfinitePlane = false;
  Planar(Scene* s, int n, GType ty) : Geom(s, 0, 0, 0), vCount(n) {
    type = ty; postProcBBox();
    for (int k = 0; k < 2; ++k) { PolyState& p = st[k]; for (int i = 0; i < 4; ++i) { p.vx[i] = p.vy[i] = p.vz[i] = p.vu[i] = p.vv[i] = 0; } p.N = Vec3(0, 0, 1); p.peqA = p.peqB = p.peqC = p.peqD = 0; for (int i = 0; i < 5; ++i) p.dotVals[i] = 0; p.baryIDenomTxtr = 0; }
  }
  void setPointsAndNormal(PolyState& p) {                           // :44-69
    double tx = 0, ty = 0, tz = 0;
    for (int i = 0; i < vCount; ++i) {
      tx += p.vx[i]; ty += p.vy[i]; tz += p.vz[i];
      p.P[i].set(p.vx[i], p.vy[i], p.vz[i]);
      int idx = (i != 0 ? i - 1 : vCount - 1);
      p.P2P[idx].set(p.vx[i] - p.vx[idx], p.vy[i] - p.vy[idx], p.vz[i] - p.vz[idx]);
      p.dotVals[idx] = p.P2P[idx].dot(p.P2P[idx]);
    }
    p.P2P0 = p.P2P[2]; p.P2P0.mult(-1.0);
    p.dotVals[vCount] = -p.P2P[0].dot(p.P2P[2]);
    p.baryIDenomTxtr = 1.0 / ((p.dotVals[0] * p.dotVals[2]) - (p.dotVals[vCount] * p.dotVals[vCount]));
    p.N = p.P2P[1].cross(p.P2P[0]); p.N.normalize();
    origin.set(tx / vCount, ty / vCount, tz / vCount);
    setTransOrigin();
  }
  void setEQ(PolyState& p) { p.peqA = p.N.x; p.peqB = p.N.y; p.peqC = p.N.z; p.peqD = -((p.peqA * p.vx[0]) + (p.peqB * p.vy[0]) + (p.peqC * p.vz[0])); }  // :90
  void invertNormal() {                                             // :71-88 (literal: reverse the live state, recompute)
    PolyState& a = st[cur]; PolyState b = a;
    for (int i = 0; i < vCount; ++i) { b.vx[vCount - 1 - i] = a.vx[i]; b.vy[vCount - 1 - i] = a.vy[i]; b.vz[vCount - 1 - i] = a.vz[i]; b.vu[vCount - 1 - i] = a.vu[i]; b.vv[vCount - 1 - i] = a.vv[i]; }
    Vec3 o = origin; double to[3] = {trans_origin[0], trans_origin[1], trans_origin[2]};
    setPointsAndNormal(b); setEQ(b);
    origin = o; trans_origin[0] = to[0]; trans_origin[1] = to[1]; trans_origin[2] = to[2];   // centroid is order independent up to rounding; BVH already built
    cur ^= 1; st[cur] = b;
  }
  void setVert(double x, double y, double z, int i) { st[0].vx[i] = x; st[0].vy[i] = y; st[0].vz[i] = z; }
  void setTxtrCoord(double u, double v, int i) { st[0].vu[i] = u; st[0].vv[i] = v; }
  void finalizePoly() {                                             // :93-100
    setPointsAndNormal(st[0]); setEQ(st[0]);
    double mnx = DMAX, mny = DMAX, mnz = DMAX, mxx = -DMAX, mxy = -DMAX, mxz = -DMAX;
    for (int i = 0; i < vCount; ++i) { const PolyState& p = st[0];
      if (p.vx[i] < mnx) mnx = p.vx[i]; if (p.vy[i] < mny) mny = p.vy[i]; if (p.vz[i] < mnz) mnz = p.vz[i];
      if (p.vx[i] > mxx) mxx = p.vx[i]; if (p.vy[i] > mxy) mxy = p.vy[i]; if (p.vz[i] > mxz) mxz = p.vz[i]; }
    minVals.set(mnx, mny, mnz); maxVals.set(mxx, mxy, mxz);
    bbox.calcMinMax(minVals, maxVals);
  }
  virtual bool checkInside(const Vec3& rp) {                         // :165-175 / :200-211
    const PolyState& p = st[cur];
    for (int i = 0; i < vCount; ++i) {
      int pIdx = (i == 0 ? vCount - 1 : i - 1);
      Vec3 intRay(rp.x - p.vx[i], rp.y - p.vy[i], rp.z - p.vz[i]);
      Vec3 tmp = intRay.cross(p.P2P[pIdx]);
      if (tmp.dot(p.N) < -EPS) return false;
    }
    return true;
  }
  RayHit intersectCheck(Ray& _ray, Ray& tr, CTM* ct) override;       // :104-115
  Vec3 getNormalAtPoint(const Vec3&, const int*) override;           // :130-136
  Vec3 getNormalAtPointImpl(bool literal) {
    // the reference normalises N itself (`res = N; res._normalize()` aliases), so N drifts by an ulp after the first hit;
    // canonical mode normalises a copy so that hit decisions do not depend on the order rays were traced in
    Vec3 res = st[cur].N; res.normalize();
    if (literal) st[cur].N = res;
    if (inverted) { invertNormal(); res = st[cur].N; res.normalize(); }
    return res;
  }
  Vec3 getMaxVec() override { const PolyState& p = st[cur]; double a = -DMAX, b = -DMAX, c = -DMAX; for (int i = 0; i < vCount; ++i) { if (p.vx[i] > a) a = p.vx[i]; if (p.vy[i] > b) b = p.vy[i]; if (p.vz[i] > c) c = p.vz[i]; } return Vec3(a, b, c); }
  Vec3 getMinVec() override { const PolyState& p = st[cur]; double a = DMAX, b = DMAX, c = DMAX; for (int i = 0; i < vCount; ++i) { if (p.vx[i] < a) a = p.vx[i]; if (p.vy[i] < b) b = p.vy[i]; if (p.vz[i] < c) c = p.vz[i]; } return Vec3(a, b, c); }
  void findTxtrCoords(const Vec3& pt, const Image* tex, double time, double uv[2]) override;   // :178-186
};
struct Plane : Planar {                                              // :227-289
  Plane(Scene* s) : Planar(s, 4, G_PLANE) { isInfinitePlane = true; }
  void setPlaneVals(double a, double b, double c, double d) {
    PolyState& p = st[cur];
    p.N.set(a, b, c); double mag = p.N.mag(); p.N.normalize();
    p.peqA = p.N.x; p.peqB = p.N.y; p.peqC = p.N.z; p.peqD = d / mag;
    Vec3 rotVec(p.peqB, p.peqC, p.peqA);
    if ((p.peqA == p.peqB) && (p.peqA == p.peqC)) rotVec.add(1, 0, 0);
    rotVec.normalize();
    int idx = 7; double sum = p.peqA + p.peqB + p.peqC;
    if (sum == 0) { sum = p.peqA + p.peqB; idx = 6; if (sum == 0) { sum = p.peqA + p.peqC; idx = 5; if (sum == 0) { sum = p.peqB + p.peqC; idx = 3; } } }
    Vec3 planePt(((idx & 4) == 4 ? -p.peqD / sum : 0), ((idx & 2) == 2 ? -p.peqD / sum : 0), ((idx & 1) == 1 ? -p.peqD / sum : 0));
    Vec3 inU = p.N.cross(rotVec), inV = p.N.cross(inU);
    p.vx[0] = planePt.x; p.vy[0] = planePt.y; p.vz[0] = planePt.z;
    Vec3 nx(planePt); nx.add(inU); p.vx[1] = nx.x; p.vy[1] = nx.y; p.vz[1] = nx.z;
    nx.add(inV); p.vx[2] = nx.x; p.vy[2] = nx.y; p.vz[2] = nx.z;
    nx = planePt; nx.add(inV); p.vx[3] = nx.x; p.vy[3] = nx.y; p.vz[3] = nx.z;
    origin.set(p.vx[0], p.vy[0], p.vz[0]);
  }
  bool checkInside(const Vec3&) override { return true; }
  Vec3 getMaxVec() override { return Vec3(DMAX, DMAX, DMAX); }
  Vec3 getMinVec() override { return Vec3(-DMAX, -DMAX, -DMAX); }
};

// ---- implicit objects (myImpObject.java)
struct ImpObject : Geom {
  double radX = 0, radY = 0, radZ = 0;
  ImpObject(Scene* s, double x, double y, double z) : Geom(s, x, y, z) {}
  Vec3 originRadCalc(const Ray& ray) {                                // :19-23
    Vec3 o = getOrigin(ray.getTime());
    return Vec3((ray.origin.x - o.x) / radX, (ray.origin.y - o.y) / radY, (ray.origin.z - o.z) / radZ);
  }
};
struct Sphere : ImpObject {                                           // :35-141
  Sphere(Scene* s, double rx, double ry, double rz, double x, double y, double z) : ImpObject(s, x, y, z) {
    type = G_SPHERE; radX = rx; radY = ry; radZ = rz; minVals = getMinVec(); maxVals = getMaxVec(); postProcBBox();
  }
  double getAVal(const Ray& r) { return ((r.direction.x / radX) * (r.direction.x / radX)) + ((r.direction.y / radY) * (r.direction.y / radY)) + ((r.direction.z / radZ) * (r.direction.z / radZ)); }
  double getBVal(const Ray& r) { Vec3 pC = originRadCalc(r); return 2 * (((r.direction.x / radX) * pC.x) + ((r.direction.y / radY) * pC.y) + ((r.direction.z / radZ) * pC.z)); }
  double getCVal(const Ray& r) { Vec3 pC = originRadCalc(r); return (pC.x * pC.x) + (pC.y * pC.y) + (pC.z * pC.z) - 1; }
  Vec3 getNormalAtPoint(const Vec3& pt, const int*) override { Vec3 r(pt); r.sub(origin); r.normalize(); if (inverted) r.mult(-1.0); return r; }
  RayHit intersectCheck(Ray& _ray, Ray& tr, CTM* ct) override;
  void findTxtrCoords(const Vec3& pt, const Image* tex, double time, double uv[2]) override;
  Vec3 getMaxVec() override { Vec3 r(origin); double v = radX + radY + radZ; r.add(v, v, v); return r; }
  Vec3 getMinVec() override { Vec3 r(origin); double v = radX + radY + radZ; r.add(-v, -v, -v); return r; }
};
struct MovingSphere : Sphere {                                        // :144-155
  Vec3 origin0, origin1;
  MovingSphere(Scene* s, double r, double x0, double y0, double z0, double x1, double y1, double z1) : Sphere(s, r, r, r, x0, y0, z0) { type = G_MOVSPHERE; origin0 = origin; origin1 = Vec3(x1, y1, z1); }
  Vec3 getOrigin(double t) override { Vec3 bMa(origin0, origin1); return Vec3(origin0.x + t * bMa.x, origin0.y + t * bMa.y, origin0.z + t * bMa.z); }   // DistRayTracer.java:429-432
};
struct HollowCylinder : ImpObject {                                   // :157-237
  double myHeight, yTop, yBottom;
  HollowCylinder(Scene* s, double rad, double h, double x, double y, double z) : ImpObject(s, x, y, z) {
    type = G_HCYL; radX = rad; radZ = rad; myHeight = h; yTop = origin.y + myHeight; yBottom = origin.y;
    minVals = getMinVec(); maxVals = getMaxVec(); postProcBBox();
  }
  double getAVal(const Ray& r) { return ((r.direction.x / radX) * (r.direction.x / radX)) + ((r.direction.z / radZ) * (r.direction.z / radZ)); }
  // originRadCalc divides y by radY (= 0 here): x/z components are unaffected, y is unused
  double getBVal(const Ray& r) { Vec3 pC = originRadCalc(r); return 2 * (((r.direction.x / radX) * pC.x) + ((r.direction.z / radZ) * pC.z)); }
  double getCVal(const Ray& r) { Vec3 pC = originRadCalc(r); return (pC.x * pC.x) + (pC.z * pC.z) - 1; }
  RayHit intersectCheck(Ray& _ray, Ray& tr, CTM* ct) override;
  Vec3 getNormalAtPoint(const Vec3& pt, const int* args) override {
    Vec3 r = (args[0] == 1) ? Vec3((origin.x - pt.x), 0, (origin.z - pt.z)) : Vec3((pt.x - origin.x), 0, (pt.z - origin.z));
    r.normalize(); if (inverted) r.mult(-1); return r;
  }
  Vec3 getMaxVec() override { Vec3 r(origin); double v = radX + radZ; r.add(v, myHeight, v); return r; }
  Vec3 getMinVec() override { Vec3 r(origin); double v = radX + radZ; r.add(-v, 0, -v); return r; }
};
struct Cylinder : HollowCylinder {                                    // :239-327
  double capEqs[2][4];
  Cylinder(Scene* s, double rad, double h, double x, double y, double z, double xO, double yO, double zO) : HollowCylinder(s, rad, h, x, y, z) {
    type = G_CYL;
    capEqs[0][0] = xO; capEqs[0][1] = yO; capEqs[0][2] = zO; capEqs[0][3] = -yTop;
    capEqs[1][0] = xO; capEqs[1][1] = -yO; capEqs[1][2] = zO; capEqs[1][3] = yBottom;
    minVals = getMinVec(); maxVals = getMaxVec(); postProcBBox();
  }
  RayHit intersectCheck(Ray& _ray, Ray& tr, CTM* ct) override;
  Vec3 getNormalAtPoint(const Vec3& pt, const int* args) override {
    Vec3 r; if (args[0] >= 2) r = Vec3((pt.x - origin.x), 0, (pt.z - origin.z)); else r = Vec3(capEqs[args[0]][0], capEqs[args[0]][1], capEqs[args[0]][2]);
    r.normalize(); if (inverted) r.mult(-1); return r;
  }
};

// ---- instance (mySceneObject.java:95-145)
struct Instance : Geom {
  Geom* obj; bool useShader = false; int instSerial = -1;
  Instance(Scene* s, Geom* base);
  Vec3 getMaxVec() override { return obj->ctm->glbl.xfPt(obj->getMaxVec()); }
  Vec3 getMinVec() override { return obj->ctm->glbl.xfPt(obj->getMinVec()); }
  int calcShadowHit(Ray&, Ray& trans, CTM* ct, double d) override { return obj->calcShadowHit(trans, trans, ct, d); }
  RayHit intersectCheck(Ray&, Ray& trans, CTM* ct) override {
    RayHit h = obj->intersectCheck(trans, trans, ct);
    if (useShader) h.shdr = shdr;
    if (h.isHit && h.instSerial < 0) h.instSerial = instSerial;
    return h;
  }
  Vec3 getNormalAtPoint(const Vec3& pt, const int* args) override { return obj->getNormalAtPoint(pt, args); }
  void findTxtrCoords(const Vec3& pt, const Image* tex, double time, double uv[2]) override { obj->findTxtrCoords(ctm->inv.xfVec(pt), tex, time, uv); }
};

// ---- acceleration structures (myGeomBase.java:200-423)
struct AccelStruct : Geom {
  AccelStruct(Scene* s) : Geom(s, 0, 0, 0) {}
  bool isAccel() const override { return true; }
  virtual RayHit traverseStruct(Ray& _ray, Ray& trans, CTM* ct) = 0;
  RayHit intersectCheck(Ray& _ray, Ray& tr, CTM* ct) override;       // :216-222 root gate
  Vec3 getNormalAtPoint(const Vec3&, const int* args) override { return BBox::faceNormal(args[1]); }
  Vec3 getMaxVec() override { return bbox.maxVals; }
  Vec3 getMinVec() override { return bbox.minVals; }
};
struct GeomList : AccelStruct {                                       // :251-306
  std::vector<Geom*> objList;
  std::vector<CTM*> hitCtm;      // cache of reBuildCTMara(child, list) (the reference rebuilds it on every leaf hit, :298)
  GeomList(Scene* s) : AccelStruct(s) { type = G_LIST; postProcBBox(); }
  void addObj(Geom* o) { objList.push_back(o); hitCtm.push_back(nullptr); Mat4 tmp = ctm->inv.multMat(o->ctm->glbl); bbox.expandByBox(o->bbox, tmp); }
  int calcShadowHit(Ray& _ray, Ray& trans, CTM* ct, double d) override;
  RayHit traverseStruct(Ray& _ray, Ray& trans, CTM* ct) override;
};
struct BVH : AccelStruct {                                            // :309-423
  bool isLeaf = true; GeomList* leafVals; BVH *leftChild = nullptr, *rightChild = nullptr; int maxSpanSplitIDX = -1;
  BVH(Scene* s) : AccelStruct(s) { type = G_BVH; leafVals = new GeomList(s); postProcBBox(); }
  typedef std::vector<Geom*> GL;
  static void buildSortedObjAras(const GL& sorted, int idxToSkip, GL res[3]);   // :338-357
  void addObjList(GL lists[3], int stIDX, int endIDX);                          // :360-386
  int calcShadowHit(Ray& _ray, Ray& trans, CTM* ct, double d) override;         // :397-404
  RayHit traverseStruct(Ray& _ray, Ray& trans, CTM* ct) override;               // :407-421
};

// ---- lights (myLight.java:12-275)
struct SampleCtx { uint32_t stream, a, b, c, dim; };   // where the next random numbers come from
struct Light : Geom {
  Vec3 lightColor; int lightID; Vec3 orientation;
  Light(Scene* s, int id, double r, double g, double b, double x, double y, double z, double dx, double dy, double dz);
  bool isLight() const override { return true; }
  RayHit intersectCheck(Ray& _ray, Ray& tr, CTM* ct) override { (void)_ray; (void)tr; (void)ct; return RayHit(); }   // shading uses lightHit() below
  // light.intersectCheck as used by calcShadowColor (myLight.java:33-41,159-163): distance to the (sampled) origin + penumbra factor
  virtual void lightHit(const Ray& shadowRay, const SampleCtx& sc, double& t, double& ltMult);
  virtual Vec3 sampleOrigin(const SampleCtx& sc) { (void)sc; return origin; }       // getOrigin(time); disk light draws 2 randoms (:251-266)
  Vec3 getNormalAtPoint(const Vec3&, const int*) override { return Vec3(0, 1, 0); }
  Vec3 getMaxVec() override { Vec3 r(origin); r.add(EPS, EPS, EPS); return r; }
  Vec3 getMinVec() override { Vec3 r(origin); r.add(-EPS, -EPS, -EPS); return r; }
  // photon emission (:85); draws start at counter dim `draw`, which is advanced
  virtual void genRndPhtnRay(uint32_t photon, uint32_t lightIdx, uint32_t& draw, Vec3& org, Vec3& dir) = 0;
  static double getAngleProb(double angle, double inner, double outer, double diff) { return (angle < inner) ? 1 : (angle > outer) ? 0 : (outer - angle) / diff; }   // :77
  Vec3 getRandDir(uint32_t photon, uint32_t lightIdx, uint32_t& draw);                                   // :59-74
};
struct PointLight : Light {
  PointLight(Scene* s, int id, double r, double g, double b, double x, double y, double z) : Light(s, id, r, g, b, x, y, z, 0, 0, 0) { type = G_POINTLIGHT; }
  void genRndPhtnRay(uint32_t photon, uint32_t lightIdx, uint32_t& draw, Vec3& org, Vec3& dir) override { dir = getRandDir(photon, lightIdx, draw); org = ctm->glbl.xfPt(origin); }
};
struct SpotLight : Light {                                              // :134-206
  double innerThet, outerThet, innerThetRad, outerThetRad, radDiff; Vec3 oPhAxis;
  SpotLight(Scene* s, int id, double r, double g, double b, double x, double y, double z, double dx, double dy, double dz, double inT, double outT) : Light(s, id, r, g, b, x, y, z, dx, dy, dz) {
    type = G_SPOTLIGHT; innerThet = inT; innerThetRad = innerThet * DEG_TO_RAD_F; outerThet = outT; outerThetRad = outerThet * DEG_TO_RAD_F;
    radDiff = outerThetRad - innerThetRad; oPhAxis = getOrthoVec(orientation);
  }
  void lightHit(const Ray& shadowRay, const SampleCtx& sc, double& t, double& ltMult) override {
    Light::lightHit(shadowRay, sc, t, ltMult);
    double angle = std::acos(-1 * shadowRay.direction.dot(orientation));
    ltMult = getAngleProb(angle, innerThetRad, outerThetRad, radDiff);
  }
  void genRndPhtnRay(uint32_t photon, uint32_t lightIdx, uint32_t& draw, Vec3& org, Vec3& dir) override;
};
struct DiskLight : Light {                                              // :212-275
  double radius; Vec3 surfTangent;
  DiskLight(Scene* s, int id, double r, double g, double b, double x, double y, double z, double dx, double dy, double dz, double rad) : Light(s, id, r, g, b, x, y, z, dx, dy, dz) {
    type = G_DISKLIGHT; radius = rad; surfTangent = getOrthoVec(orientation);
  }
  Vec3 diskPos(double uAngle, double uRad) {                           // getRandomDiskPos :251-258
    Vec3 tmp = rotVecAroundAxis(surfTangent, orientation, urange(uAngle, 0, TWO_PI_F));
    tmp.normalize(); tmp.mult(urange(uRad, 0, radius)); tmp.add(origin); return tmp;
  }
  Vec3 sampleOrigin(const SampleCtx& sc) override;
  Vec3 getMaxVec() override { Vec3 r(origin); r.add(radius, radius, radius); return r; }
  Vec3 getMinVec() override { Vec3 r(origin); r.add(-radius, -radius, -radius); return r; }
  void genRndPhtnRay(uint32_t photon, uint32_t lightIdx, uint32_t& draw, Vec3& org, Vec3& dir) override;
};

}  // namespace orc
