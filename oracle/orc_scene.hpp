// ORACLE (test infrastructure only -- never linked into the product path).
// CPU restatement of the reference's scene state, .cli interpreter, render
// drivers (cameras) and photon emission.  PARITY UNPINNED, see orc_math.hpp.
//
// Follows (relative to /root/reference/src/rayTracerDistAccelShdPhtnMap/):
//   myRTFileReader.java:15-378 (interpreter)
//   myScene.java:167-324 (state, list/BVH finalisation), :328-410 (Sierpinski, instances),
//                :413-565 (lights, primitives, shader factory), :571-777 (texture params),
//                :805-857 (setters), :868-914 (queries), :919-1099 (photons),
//                :1104-1149 (skydome), :1235-1323 (matrix stack), :1348-1755 (cameras)
#pragma once
#include "orc_shade.hpp"
#include <fstream>
#include <functional>
#include <sstream>

namespace orc {

struct Options {
  int cols = 300, rows = 300;      // reference hard-codes 300x300 (DistRayTracer.java:15-16)
  int sppOverride = -1;            // <0: use the file's rays_per_pixel
  uint64_t seed = 0x5EED;
  bool literalRenorm = false;      // true: re-normalise the source ray on every getTransformedRay (myRay.java:93) even if already unit length
  long long photonOverride = -1;   // <0: use the file's photon count
  std::string dataDir;             // directory of .cli files
  std::string texDir;              // directory of decoded textures (<name>.argb)
};

enum SceneKind { SC_FOV, SC_FISHEYE, SC_ORTHO };

struct Scene {
  Options opt;
  int sceneCols, sceneRows, numRays = 8, numPhotonRays = 4, maxPrimsPerLeaf = 5;
  int objCnt = 0;
  Vec3 eyeOrigin;
  double rayYOffset, rayXOffset, maxDim, yStart, xStart, fishMult;
  std::string saveName;
  const Image *currTextureTop = nullptr, *currTextureBottom = nullptr, *currBkgTexture = nullptr;
  std::vector<Geom*> allObjsToFind, objList, lightList, tmpObjList;
  std::map<std::string, Geom*> namedObjs;
  Sphere* mySkyDome = nullptr; const Image* skyTex = nullptr;
  int numLights = 0, objCount = 0, numNonLights = 0;
  // flags (myScene.java:72-98)
  bool extensions = false;        // `extensions on` (orc_ext.hpp)
  bool simpleRefr = false, hasDpthOfFld = false, addToTmpList = false, glblTxtrdBkg = false, glblRefine = false,
       glblTxtrdTop = false, glblTxtrdBtm = false, usePhotonMap = false, isCausticPhtn = false, isPhtnMapRndrd = false;
  KDTree* photonTree = nullptr; int numPhotons = 0, kNhood = 0; float ph_max_near_dist = 0;
  double causticsLightPwrMult = 40.0, diffuseLightPwrMult = 8.0;
  // procedural texture state (:117-139)
  int txtrType = 0; double noiseScale = 1;
  std::vector<Vec3> noiseColors{mkColor(.7, .7, .7), mkColor(.2, .2, .2)};
  int numOctaves = 8; double turbMult = 1.0, colorScale = 10.0, colorMult = .2; Vec3 pdMult{10, 10, 10};
  bool rndColors = false, useCustClrs = false, useFwdTrans = false; int numOverlays = 1;
  double avgNumPerCell = 1, mortarThresh = .04; int numPtsDist = 2, distFunc = 1, roiFunc = 1;
  // current material (:144-145)
  Vec3 currDiffuseColor, currAmbientColor, currSpecularColor, globCurPermClr, currKReflClr, backgroundColor;
  double currPhongExp = 0, currKRefl = 0, globRfrIdx = 0, currKTrans = 0, currDepth = 0, lens_radius = 0, lens_focal_distance = 0;
  Plane* focalPlane = nullptr;
  int numRaysPerPixel = 0;
  // matrix stack (myVector.java:225-256): 20 slots declared, 10 allocated
  Mat4 stack[10]; int top = 0; int currMatrixDepthIDX = 0;
  // camera
  SceneKind kind = SC_FOV; double fov = 60, fovRad = 0, viewZ = -1, fishEye = 0, fishEyeRad = 0, aperatureHlf = 0, orthoWidth = 0, orthoHeight = 0, orthPerRow = 0, orthPerCol = 0;
  // bookkeeping
  int primSerialCnt = 0, instSerialCnt = 0, shaderSerialCnt = 0;
  Stats stats; bool countingPrimary = false;
  std::map<std::string, Image*> imageCache;
  std::vector<std::string> warnings;
  std::vector<Shader*> allShaders;

  explicit Scene(const Options& o) : opt(o) {
    setImageSize(o.cols, o.rows);
    focalPlane = new Plane(this);
    setSceneParamsFOV(60);
  }
  void setImageSize(int c, int r) {                                      // myScene.java:780-793
    sceneCols = c; sceneRows = r; rayYOffset = sceneRows / 2.0; rayXOffset = sceneCols / 2.0;
    maxDim = std::max(sceneRows, sceneCols);
    yStart = ((maxDim - sceneRows) / 2.0) - rayYOffset; xStart = ((maxDim - sceneCols) / 2.0) - rayXOffset; fishMult = 2.0 / maxDim;
  }
  // ---- cameras (:1367-1381, :1556-1560, :1678-1684)
  void setSceneParamsFOV(double f) {
    kind = SC_FOV; fov = f; fovRad = M_PI * fov / 180.0;
    if (std::fabs(fov - 180) < .001) { fov -= .001; fovRad -= .0001; }
    viewZ = -1 * (std::max(sceneRows, sceneCols) / 2.0) / std::tan(fovRad / 2);
    if (hasDpthOfFld) focalPlane->setPlaneVals(0, 0, 1, lens_focal_distance);
  }
  void setSceneParamsFish(double f) { kind = SC_FISHEYE; fishEye = f; fishEyeRad = M_PI * fishEye / 180.0; aperatureHlf = fishEyeRad / 2.0; }
  void setSceneParamsOrtho(double w, double h) { kind = SC_ORTHO; orthoWidth = w; orthoHeight = h; double div = std::min(sceneCols, sceneRows); orthPerRow = orthoHeight / div; orthPerCol = orthoWidth / div; }
  // ---- matrix stack (:1235-1323)
  Mat4 peek() const { return stack[top]; }
  void gtPushMatrix() { if (currMatrixDepthIDX < 20) { if (top + 1 >= 10) throw std::runtime_error("matrix stack overflow (reference NPE, SURVEY Q18)"); ++top; stack[top] = stack[top - 1]; currMatrixDepthIDX++; } }
  void gtPopMatrix() { if (top == 0) return; top--; currMatrixDepthIDX--; }
  void updateCTM(const Mat4& m) { stack[top] = stack[top].multMat(m); }
  void gtTranslate(double tx, double ty, double tz) { Mat4 t; t.m[0][3] = tx; t.m[1][3] = ty; t.m[2][3] = tz; updateCTM(t); }
  void gtScale(double sx, double sy, double sz) { Mat4 s; s.m[0][0] = sx; s.m[1][1] = sy; s.m[2][2] = sz; updateCTM(s); }
  void gtRotate(double angle, double ax, double ay, double az) {            // :1280-1318
    double ar = (double)(angle * M_PI) / 180.0;
    Mat4 R1, R2;
    Vec3 axis(ax, ay, az), an = axis.normalized();
    Vec3 nv = (ax == 0) ? Vec3(1, 0, 0) : Vec3(0, 1, 0);
    Vec3 b = an.cross(nv), bn = b.normalized(); Vec3 c = an.cross(bn), cn = c.normalized();
    R1.m[0][0] = an.x; R1.m[0][1] = an.y; R1.m[0][2] = an.z;
    R1.m[1][0] = bn.x; R1.m[1][1] = bn.y; R1.m[1][2] = bn.z;
    R1.m[2][0] = cn.x; R1.m[2][1] = cn.y; R1.m[2][2] = cn.z;
    Mat4 R1T = R1.transpose();
    R2.m[1][1] = std::cos(ar); R2.m[1][2] = -std::sin(ar); R2.m[2][1] = std::sin(ar); R2.m[2][2] = std::cos(ar);
    Mat4 tmp = R2.multMat(R1); updateCTM(R1T.multMat(tmp));
  }
  // ---- material setters (:817-857)
  void setKRefl(double k, double r, double g, double b) { currKRefl = k; currKReflClr = mkColor(r, g, b); }
  void setRfrIdx(double i, double r, double g, double b) { globRfrIdx = i; globCurPermClr = mkColor(r, g, b); }
  void setSurface(const Vec3& d, const Vec3& a, const Vec3& s, double ph, double kr) {
    txtrType = 0; currDiffuseColor = mkColor(d.x, d.y, d.z); currAmbientColor = mkColor(a.x, a.y, a.z); currSpecularColor = mkColor(s.x, s.y, s.z);
    currPhongExp = ph; setKRefl(kr, kr, kr, kr); setRfrIdx(0, 0, 0, 0); currKTrans = 0; globRfrIdx = 0; globCurPermClr = mkColor(0, 0, 0);
  }
  void setSurface(const Vec3& d, const Vec3& a, const Vec3& s, double ph, double kr, double kt) { setSurface(d, a, s, ph, kr); currKTrans = kt; setRfrIdx(0, 0, 0, 0); }
  void setSurface(const Vec3& d, const Vec3& a, const Vec3& s, double ph, double kr, double kt, double ri) { setSurface(d, a, s, ph, kr, kt); setRfrIdx(ri, ri, ri, ri); }
  void setDpthOfFld(double lRad, double lFD) { lens_radius = lRad; lens_focal_distance = lFD; hasDpthOfFld = true; focalPlane->setPlaneVals(0, 0, 1, lens_focal_distance); }
  // ---- object lists (:300-324, :394-410, :558-565)
  void addObjectToScene(Geom* o) { addObjectToScene(o, o); }
  void addObjectToScene(Geom* o, Geom* cmp) {
    if (addToTmpList) { tmpObjList.push_back(o); return; }
    if (cmp->isLight()) { lightList.push_back(o); numLights++; } else { objList.push_back(o); numNonLights++; }
    allObjsToFind.push_back(o); objCount++;
  }
  void startTmpObjList() { tmpObjList.clear(); addToTmpList = true; }
  void endTmpObjList(int lstType) {
    addToTmpList = false; AccelStruct* acc = nullptr;
    if (lstType == 0) { GeomList* l = new GeomList(this); for (Geom* g : tmpObjList) l->addObj(g); acc = l; }
    else {
      BVH* b = new BVH(this); BVH::GL lists[3];
      BVH::buildSortedObjAras(tmpObjList, -1, lists);
      b->addObjList(lists, 0, (int)lists[0].size() - 1);          // endIDX = size-1: SURVEY Q2
      acc = b;
    }
    addObjectToScene(acc);
  }
  void setObjectAsNamedObject(const std::string& name) {
    if (allObjsToFind.empty()) throw std::runtime_error("named_object with empty scene");
    Geom* o = allObjsToFind.back(); allObjsToFind.pop_back(); --objCount;
    if (o->isLight()) { lightList.pop_back(); numLights--; } else { objList.pop_back(); numNonLights--; }
    namedObjs[name] = o;
  }
  void addInstance(const std::string& name, bool addShdr);
  // ---- Sierpinski (:328-392)
  void setSierpShdr(int level, int maxLevel) {
    float bVal = 1.0f - std::min(1.0f, (1.5f * level / maxLevel)), rVal = 1.0f - bVal,
          tmp = std::min((1.2f * (level - (maxLevel / 2))) / (1.0f * maxLevel), 1.0f), gVal = (tmp * tmp);
    Vec3 cDiff = mkColor(std::min(1.0f, rVal + .5f), std::min(1.0f, gVal + .5f), std::min(1.0f, bVal + .5f));
    glblTxtrdTop = false; glblTxtrdBtm = false;
    setSurface(cDiff, Vec3(0, 0, 0), Vec3(0, 0, 0), 0, 0);
  }
  void sierpShiftObj(float nt) { gtRotate(120, 1, 0, 0); gtTranslate(0, nt, 0); gtRotate(-120, 1, 0, 0); }
  void buildSierpSubTri(float dim, float scVal, const std::string& inst, int level, int maxLevel, bool addShader) {
    if (level >= maxLevel) return;
    float newDim = scVal * dim;
    gtPushMatrix(); gtTranslate(0, .1f * dim, 0); gtRotate(70, 0, 1, 0);
    if (addShader) setSierpShdr(level, maxLevel);
    addInstance(inst, addShader); gtPopMatrix();
    static const float sqrt66 = std::sqrt(6.0f) / 6.0f;
    float newTrans = sqrt66 * dim;
    gtPushMatrix(); gtTranslate(0, newTrans, 0); gtScale(scVal, scVal, scVal); buildSierpSubTri(newDim, scVal, inst, level + 1, maxLevel, addShader); gtPopMatrix();
    gtPushMatrix(); sierpShiftObj(newTrans); gtScale(scVal, scVal, scVal); buildSierpSubTri(newDim, scVal, inst, level + 1, maxLevel, addShader); gtPopMatrix();
    gtPushMatrix(); gtRotate(120, 0, 1, 0); sierpShiftObj(newTrans); gtRotate(-120, 0, 1, 0); gtScale(scVal, scVal, scVal); buildSierpSubTri(newDim, scVal, inst, level + 1, maxLevel, addShader); gtPopMatrix();
    gtPushMatrix(); gtRotate(-120, 0, 1, 0); sierpShiftObj(newTrans); gtRotate(120, 0, 1, 0); gtScale(scVal, scVal, scVal); buildSierpSubTri(newDim, scVal, inst, level + 1, maxLevel, addShader); gtPopMatrix();
  }
  void buildSierpinski(const std::string& name, float scVal, int depth, bool useShdr) { startTmpObjList(); buildSierpSubTri(8, scVal, name, 0, depth, useShdr); endTmpObjList(1); }
  // ---- shader / texture factory (:524-542)
  Shader* getCurShader();
  Texture* getCurTexture(Shader* sh);
  // ---- procedural texture parameters (:571-777)
  void resetDfltTxtrVals() {
    txtrType = 0; numOctaves = 4; numOverlays = 1; numPtsDist = 2; distFunc = 1; roiFunc = 1; rndColors = false; useCustClrs = false; useFwdTrans = false;
    noiseScale = 1.0; turbMult = 1.0; colorScale = 5.0; colorMult = .1; avgNumPerCell = 1.0; mortarThresh = 0.05; pdMult = Vec3(1.0, 1.0, 1.0);
    noiseColors = {getClr("clr_nearblack"), getClr("clr_white")};
  }
  void setProcTxtrVals(int tt, int oct, int ovl, int npd, int df, int rf, bool rc, bool ucc, bool uft, double ns, double tm, double cs, double cm, double anpc, double mt, const Vec3& pd) {
    txtrType = tt; numOctaves = oct; numOverlays = ovl; numPtsDist = npd; distFunc = df; roiFunc = rf; rndColors = rc; useCustClrs = ucc; useFwdTrans = uft;
    noiseScale = ns; turbMult = tm; colorScale = cs; colorMult = cm; avgNumPerCell = anpc; mortarThresh = mt; pdMult = pd;
  }
  // ---- images
  std::function<Image*(const std::string&)> imageLoader;
  const Image* loadImage(const std::string& name);
  // ---- queries (:879-914)
  int calcShadow(Ray& ray, double distToLight) {
    for (Geom* o : objList) { Ray tr = ray.getTransformedRay(ray, o->ctm->inv); if (o->calcShadowHit(ray, tr, o->ctm, distToLight) == 1) return 1; }
    return 0;
  }
  RayHit findClosestRayHit(Ray& ray) {
    RayHit best;                                  // rayHit(false), t = MAX ; TreeMap keeps the first-inserted among equal t
    for (Geom* o : objList) { Ray tr = ray.getTransformedRay(ray, o->ctm->inv); RayHit h = o->intersectCheck(ray, tr, o->ctm); if (h.isHit && dcompare(h.t, best.t) < 0) best = h; }
    return best;
  }
  Vec3 reflectRay(Ray& ray);
  Vec3 getBackgroundTextureColor(const Ray& ray);                          // :1104-1149
  // ---- photons (:919-1099)
  void sendCausticPhotons(); void sendDiffusePhotons();
  void initRender() { if (usePhotonMap && !isPhtnMapRndrd) { if (isCausticPhtn) sendCausticPhotons(); else { sendDiffusePhotons(); } isPhtnMapRndrd = true; } }
  // ---- render (cameras :1386-1462, :1481-1531, :1563-1643, :1687-1753)
  struct PixelOut { Vec3 rgb; int32_t argb; int hitPrim = -1, hitInst = -1; double t = 0; };
  PixelOut renderPixel(int row, int col);
  Vec3 tracePrimary(Ray& ray, PixelOut* aov);
  // ---- parser
  void readRTFile(const std::string& fileName, bool isMain);
  void readPrimData(const std::vector<std::string>& tk);
  void setTexture(const std::vector<std::string>& tk);
  void setTxtrColor(const std::vector<std::string>& tk);
};

// ===========================================================================
// implementations
// ===========================================================================
inline double Ray::getTime() const { return u01(scn->opt.seed, key.stream, key.a, key.b, key.c, key.timeDim); }

inline Ray Ray::getTransformedRay(Ray& src, const Mat4& trans) const {      // myRay.java:91-102
  if (scn->opt.literalRenorm || !src.dirNormalized) { src.direction.normalize(); src.dirNormalized = true; }
  Ray nr; nr.scn = scn; nr.gen = src.gen; nr.key = src.key;
  nr.origin = trans.xfPt(src.origin); nr.direction = trans.xfVec(src.direction);
  nr.originAra[0] = nr.origin.x; nr.originAra[1] = nr.origin.y; nr.originAra[2] = nr.origin.z;
  nr.dirAra[0] = nr.direction.x; nr.dirAra[1] = nr.direction.y; nr.dirAra[2] = nr.direction.z;
  nr.dirNormalized = false;
  for (int i = 0; i < 5; ++i) nr.currKTrans[i] = src.currKTrans[i];
  return nr;
}

inline void RayHit::reCalcCTMHitNorm(CTM* c) {
  ctm = c; fwdTransHitLoc = ctm->glbl.xfPt(hitLoc);
  Vec3 n = obj->getNormalAtPoint(hitLoc, args);
  objNorm = ctm->adj.xfVec(n); objNorm.normalize();
}

inline Geom::Geom(Scene* s, double x, double y, double z) : scene(s), origin(x, y, z) {
  ID = scene->objCnt++; ctm = CTM::build(scene->peek()); setTransOrigin();
}
inline int Geom::calcShadowHit(Ray& _ray, Ray& trans, CTM* ct, double distToLight) {
  RayHit h = intersectCheck(_ray, trans, ct);
  if (h.isHit && (distToLight - h.t) > EPS) return 1;
  return 0;
}
inline RayHit Geom::objHit(const Ray& tr, const Vec3& rawRayDir, CTM* ct, const Vec3& pt, const int* args, double t) {
  RayHit h; h.isHit = true; h.obj = this; h.shdr = shdr; h.ctm = ct; h.t = t; h.hitLoc = pt;
  if (args) { h.args[0] = args[0]; h.args[1] = args[1]; }
  Vec3 nn = ct->adj.xfVec(getNormalAtPoint(pt, h.args)); nn.normalize(); h.objNorm = nn;
  h.fwdTransHitLoc = ct->glbl.xfPt(pt); h.fwdTransRayDir = rawRayDir; h.ltMult = 1;
  h.gen = tr.gen; for (int i = 0; i < 5; ++i) h.rayKTrans[i] = tr.currKTrans[i]; h.key = tr.key;
  return h;
}

inline Vec3 Planar::getNormalAtPoint(const Vec3&, const int*) { return getNormalAtPointImpl(scene->opt.literalRenorm); }
inline RayHit RndrdBox::intersectCheck(Ray&, Ray& tr, CTM* ct) {
  double t; int idx; if (!bbox.test(tr, t, idx, scene->stats)) return RayHit();
  int args[2] = {0, idx};
  return objHit(tr, ct->glbl.xfVec(tr.direction), ct, tr.pointOnRay(t), args, t);
}

inline RayHit Planar::intersectCheck(Ray& _ray, Ray& tr, CTM* ct) {         // myPlanarObject.java:104-115
  ++scene->stats.primTests;
  for (int guard = 0; guard < 3; ++guard) {
    const PolyState& p = st[cur];
    double planeRes = p.N.dot(tr.direction);
    if (std::fabs(planeRes) > 0) {
      if (planeRes > 0) { invertNormal(); continue; }      // reference recurses; a third flip would never terminate there
      double t = -(p.N.dot(tr.origin) + p.peqD) / planeRes;
      if ((t > EPS) && checkInside(tr.pointOnRay(t))) return objHit(tr, _ray.direction, ct, tr.pointOnRay(t), nullptr, t);
    }
    return RayHit();
  }
  return RayHit();
}
inline void Planar::findTxtrCoords(const Vec3& pt, const Image* tex, double, double uv[2]) {
  const PolyState& p = st[cur];
  Vec3 v2 = vsub(pt, p.P[0]);
  double dot20 = v2.dot(p.P2P[0]), dot21 = v2.dot(p.P2P0),
         c_u = ((p.dotVals[2] * dot20) - (p.dotVals[vCount] * dot21)) * p.baryIDenomTxtr,
         c_v = ((p.dotVals[0] * dot21) - (p.dotVals[vCount] * dot20)) * p.baryIDenomTxtr, c_w = 1 - c_u - c_v;
  double u = p.vu[0] * c_w + p.vu[1] * c_u + p.vu[2] * c_v, v = p.vv[0] * c_w + p.vv[1] * c_u + p.vv[2] * c_v;
  uv[0] = u * (tex->width - 1); uv[1] = (1 - v) * (tex->height - 1);
}

inline RayHit Sphere::intersectCheck(Ray& _ray, Ray& tr, CTM* ct) {          // myImpObject.java:76-94
  ++scene->stats.primTests;
  double a = getAVal(tr), ta = 2 * a, b = getBVal(tr), c = getCVal(tr), discr = ((b * b) - (2 * ta * c));
  if (!(discr < 0)) {
    double d1 = std::sqrt(discr), t1 = (-1 * b + d1) / (ta), t2 = (-1 * b - d1) / (ta);
    double tVal = jmin(t1, t2);
    if (tVal < EPS) { tVal = jmax(t1, t2); if (tVal < EPS) return RayHit(); }
    return objHit(tr, _ray.direction, ct, tr.pointOnRay(tVal), nullptr, tVal);
  }
  return RayHit();
}
inline void Sphere::findTxtrCoords(const Vec3& pt, const Image* tex, double time, double uv[2]) {   // :97-122
  Vec3 to = getOrigin(time);
  double a1v = (pt.y - to.y) / radY; a1v = (a1v > 1) ? 1 : (a1v < -1) ? -1 : a1v;
  double v = (tex->height - 1) * std::acos(a1v) / M_PI;
  double shWm1 = tex->width - 1, z1 = (pt.z - to.z);
  double q = v / (tex->height - 1);
  double a0 = (pt.x - to.x) / radX; a0 = (a0 > 1) ? 1 : (a0 < -1) ? -1 : a0;
  double a1 = std::sin(q * M_PI);
  double a2 = (std::fabs(a1) < EPS) ? 1 : a0 / a1;
  double u = (z1 <= EPS) ? ((shWm1 * (std::acos(a2)) / (TWO_PI_F)) + shWm1 / 2.0f) : shWm1 - ((shWm1 * (std::acos(a2)) / (TWO_PI_F)) + shWm1 / 2.0f);
  u = (u < 0) ? 0 : (u > shWm1) ? shWm1 : u;
  uv[0] = u; uv[1] = v;
}
inline RayHit HollowCylinder::intersectCheck(Ray& _ray, Ray& tr, CTM* ct) {   // :174-192
  ++scene->stats.primTests;
  double a = getAVal(tr), b = getBVal(tr), c = getCVal(tr), discr = ((b * b) - (4 * a * c));
  if (!(discr < 0)) {
    double d1 = std::sqrt(discr), t1 = (-b + d1) / (2 * a), t2 = (-b - d1) / (2 * a);
    double cyltVal = jmin(t1, t2), cyltOtr = jmax(t1, t2);
    if (cyltVal < -EPS) { double tmp = cyltOtr; cyltOtr = cyltVal; cyltVal = tmp; if (cyltVal < -EPS) return RayHit(); }
    double yInt1 = tr.origin.y + (cyltVal * tr.direction.y);
    if ((cyltVal > EPS) && (yInt1 > yBottom) && (yInt1 < yTop)) { int args[2] = {0, 0}; return objHit(tr, _ray.direction, ct, tr.pointOnRay(cyltVal), args, cyltVal); }
    double yInt2 = tr.origin.y + (cyltOtr * tr.direction.y);
    if ((cyltOtr > EPS) && (yInt2 > yBottom) && (yInt2 < yTop)) { int args[2] = {1, 0}; return objHit(tr, _ray.direction, ct, tr.pointOnRay(cyltOtr), args, cyltOtr); }
  }
  return RayHit();
}
inline RayHit Cylinder::intersectCheck(Ray& _ray, Ray& tr, CTM* ct) {         // :259-302
  ++scene->stats.primTests;
  double a = getAVal(tr), b = getBVal(tr), c = getCVal(tr);
  double discr = ((b * b) - (4 * a * c));
  if (!(discr < 0)) {
    double d1 = std::sqrt(discr), t1 = (-b + d1) / (2 * a), t2 = (-b - d1) / (2 * a);
    double cyltVal = jmin(t1, t2), cyltOtr = jmax(t1, t2);
    if (cyltVal < EPS) { cyltOtr = cyltVal; cyltVal = jmax(t1, t2); if (cyltVal < EPS) return RayHit(); }
    bool planeRes = true; double num[2] = {0, 0}, denom[2] = {1, 1}, pl[2] = {0, 0};
    for (int i = 0; i < 2; ++i) {
      denom[i] = capEqs[i][0] * tr.direction.x + capEqs[i][1] * tr.direction.y + capEqs[i][2] * tr.direction.z;
      if (std::fabs(denom[i]) > EPS) { num[i] = capEqs[i][0] * tr.origin.x + capEqs[i][1] * tr.origin.y + capEqs[i][2] * tr.origin.z + capEqs[i][3]; pl[i] = -num[i] / denom[i]; }
      else pl[i] = 10000;
    }
    double pltVal = jmin(pl[0], pl[1]); int idxVis = (pltVal == pl[0] ? 0 : 1);
    if (pltVal < 0) { pltVal = pl[idxVis]; if (pltVal < EPS) planeRes = false; }
    double tVal = 0, maxCylT = jmax(cyltVal, cyltOtr), minCylT = jmin(cyltVal, cyltOtr);
    if (planeRes && (((minCylT <= 0) && (pltVal >= -EPS) && (pltVal <= maxCylT)) || ((pltVal > minCylT) && (pltVal <= maxCylT)))) tVal = pltVal;
    else { tVal = cyltVal; idxVis = 2; }
    double yInt1 = tr.origin.y + (tVal * tr.direction.y);
    if ((yInt1 + EPS >= yBottom) && (yInt1 - EPS <= yTop)) { int args[2] = {idxVis, 0}; return objHit(tr, _ray.direction, ct, tr.pointOnRay(tVal), args, tVal); }
  }
  return RayHit();
}

inline Instance::Instance(Scene* s, Geom* base) : Geom(s, 0, 0, 0), obj(base) {   // mySceneObject.java:98-110
  ctm = CTM::build(obj->ctm->glbl.multMat(scene->peek()));     // buildCTMara(scene, obj.CTM): obj.CTM x stackTop
  type = G_INSTANCE; minVals = getMinVec(); maxVals = getMaxVec(); postProcBBox(); shdr = nullptr;
}
inline void Scene::addInstance(const std::string& name, bool addShdr) {
  auto it = namedObjs.find(name); if (it == namedObjs.end()) throw std::runtime_error("instance of unknown object " + name);
  Instance* in = new Instance(this, it->second); in->instSerial = instSerialCnt++;
  if (addShdr) { in->useShader = true; in->shdr = getCurShader(); }
  addObjectToScene(in, it->second);
}

inline RayHit AccelStruct::intersectCheck(Ray& _ray, Ray& tr, CTM* ct) {
  double t; int idx; if (!bbox.test(tr, t, idx, scene->stats)) return RayHit();
  return traverseStruct(_ray, tr, ct);
}
inline int GeomList::calcShadowHit(Ray& _ray, Ray& trans, CTM* ct, double d) {   // myGeomBase.java:268-277
  double t; int idx;
  if (!(bbox.test(trans, t, idx, scene->stats) && (d - t) > EPS)) return 0;
  for (Geom* o : objList) { Ray r = _ray.getTransformedRay(_ray, o->ctm->inv); if (o->calcShadowHit(_ray, r, ct, d) == 1) return 1; }
  return 0;
}
inline RayHit GeomList::traverseStruct(Ray& _ray, Ray&, CTM*) {                   // :281-302
  double clsT = DMAX; RayHit clsHit; Geom* clsObj = nullptr; Ray clsRay; bool have = false; size_t clsIdx = 0;
  for (size_t i = 0; i < objList.size(); ++i) {
    Geom* o = objList[i];
    Ray r = _ray.getTransformedRay(_ray, o->ctm->inv);
    RayHit h = o->intersectCheck(_ray, r, o->ctm);
    if (h.t < clsT) { clsObj = o; clsHit = h; clsT = h.t; clsRay = r; have = true; clsIdx = i; }
  }
  if (!have) return RayHit();
  if (!hitCtm[clsIdx]) hitCtm[clsIdx] = CTM::build(ctm->glbl.multMat(clsObj->ctm->glbl));   // reBuildCTMara(child, list): list.CTM x child.CTM
  clsHit.reCalcCTMHitNorm(hitCtm[clsIdx]);
  if (!clsObj->isAccel()) return clsHit;
  return static_cast<AccelStruct*>(clsObj)->traverseStruct(_ray, clsRay, clsHit.ctm);
}
inline void BVH::buildSortedObjAras(const GL& sorted, int skip, GL res[3]) {
  if (skip != -1) res[skip] = sorted;
  for (int i = 0; i < 3; ++i) {
    if (i == skip) continue;
    res[i] = sorted;       // TreeMap<Double, List>: ascending key, insertion order among equal keys == stable sort with Double.compare
    std::stable_sort(res[i].begin(), res[i].end(), [i](const Geom* a, const Geom* b) { return dcompare(a->trans_origin[i], b->trans_origin[i]) < 0; });
  }
}
inline void BVH::addObjList(GL lists[3], int stIDX, int endIDX) {
  int objListSize = endIDX - stIDX;
  if (objListSize <= scene->maxPrimsPerLeaf) {
    isLeaf = true; leafVals = new GeomList(scene); leafVals->ctm = ctm;
    for (Geom* o : lists[0]) leafVals->addObj(o);
    bbox.expandByBox(leafVals->bbox);
  } else {
    isLeaf = false;
    int split = (int)(.5 * objListSize);
    double maxSpan = -1; int ax = -1; int n = (int)lists[0].size();              // DistRayTracer.java:409-418
    for (int i = 0; i < 3; ++i) { double diff = lists[i][n - 1]->trans_origin[i] - lists[i][0]->trans_origin[i]; if (maxSpan < diff) { maxSpan = diff; ax = i; } }
    if (ax < 0) throw std::runtime_error("BVH split axis undefined (NaN centroids)");
    maxSpanSplitIDX = ax;
    leftChild = new BVH(scene); leftChild->ctm = ctm; rightChild = new BVH(scene); rightChild->ctm = ctm;
    GL sub(lists[ax].begin(), lists[ax].begin() + split), l3[3];
    buildSortedObjAras(sub, ax, l3); leftChild->addObjList(l3, stIDX, stIDX + split);
    GL sub2(lists[ax].begin() + split, lists[ax].begin() + objListSize), r3[3];
    buildSortedObjAras(sub2, ax, r3); rightChild->addObjList(r3, stIDX + split, endIDX);
    bbox.expandByBox(leftChild->bbox); bbox.expandByBox(rightChild->bbox);
  }
}
inline int BVH::calcShadowHit(Ray& _ray, Ray& trans, CTM* ct, double d) {
  if (isLeaf) return leafVals->calcShadowHit(_ray, trans, ct, d);
  double t; int idx;
  bool l = leftChild->bbox.test(trans, t, idx, scene->stats) && (d - t) > EPS;
  if (l && leftChild->calcShadowHit(_ray, trans, ct, d) == 1) return 1;
  bool r = rightChild->bbox.test(trans, t, idx, scene->stats) && (d - t) > EPS;
  if (r && rightChild->calcShadowHit(_ray, trans, ct, d) == 1) return 1;
  return 0;
}
inline RayHit BVH::traverseStruct(Ray& _ray, Ray& trans, CTM* ct) {
  if (isLeaf) return leafVals->traverseStruct(_ray, trans, ct);
  double tl, tr2; int idx;
  RayHit hit;  // miss
  if (leftChild->bbox.test(trans, tl, idx, scene->stats)) hit = leftChild->traverseStruct(_ray, trans, ct);
  RayHit hit2;
  if (rightChild->bbox.test(trans, tr2, idx, scene->stats)) {
    if (!hit.isHit || (tr2 < hit.t)) hit2 = rightChild->traverseStruct(_ray, trans, ct);
    else { hit2.isHit = false; hit2.t = tr2; }     // un-traversed box "hit": can never win the <= below (tr2 >= hit.t)
  }
  if (hit.t <= hit2.t) return hit;
  return hit2.obj ? hit2 : RayHit();
}

// ---- lights
inline Light::Light(Scene* s, int id, double r, double g, double b, double x, double y, double z, double dx, double dy, double dz) : Geom(s, x, y, z) {
  minVals = getMinVec(); maxVals = getMaxVec(); postProcBBox();
  lightColor = mkColor(r, g, b); origin.set(x, y, z); lightID = id; orientation = Vec3(dx, dy, dz); orientation.normalize();
}
inline void Light::lightHit(const Ray& sr, const SampleCtx& sc, double& t, double& ltMult) {
  SampleCtx s2 = sc; s2.dim += 2;                       // second, independent origin sample
  t = sr.origin.dist(sampleOrigin(s2)); ltMult = 1;
}
inline Vec3 DiskLight::sampleOrigin(const SampleCtx& sc) {
  double ua = u01(scene->opt.seed, sc.stream, sc.a, sc.b, sc.c, sc.dim), ur = u01(scene->opt.seed, sc.stream, sc.a, sc.b, sc.c, sc.dim + 1);
  return diskPos(ua, ur);
}
inline Vec3 Light::getRandDir(uint32_t photon, uint32_t li, uint32_t& draw) {
  double x, y, z, sq;
  do {
    x = urange(u01(scene->opt.seed, STREAM_PHOTON, photon, li, 0, draw), -1.0, 1.0);
    y = urange(u01(scene->opt.seed, STREAM_PHOTON, photon, li, 0, draw + 1), -1.0, 1.0);
    z = urange(u01(scene->opt.seed, STREAM_PHOTON, photon, li, 0, draw + 2), -1.0, 1.0); draw += 3;
    sq = (x * x) + (y * y) + (z * z);
  } while ((sq > 1.0) || (sq < EPS));
  double mag = std::sqrt(sq); return Vec3(x / mag, y / mag, z / mag);
}
inline void SpotLight::genRndPhtnRay(uint32_t photon, uint32_t li, uint32_t& draw, Vec3& org, Vec3& dir) {   // myLight.java:165-185
  double prob, angle; double checkProb = urange(u01(scene->opt.seed, STREAM_PHOTON, photon, li, 0, draw++), 0, 1);
  do { angle = urange(u01(scene->opt.seed, STREAM_PHOTON, photon, li, 0, draw++), 0, outerThetRad); prob = getAngleProb(angle, innerThetRad, outerThetRad, radDiff); } while (prob > checkProb);
  Vec3 tmp = rotVecAroundAxis(orientation, oPhAxis, angle); tmp.normalize();
  tmp = rotVecAroundAxis(tmp, orientation, urange(u01(scene->opt.seed, STREAM_PHOTON, photon, li, 0, draw++), 0, TWO_PI_F));
  dir = tmp; org = ctm->glbl.xfPt(origin);
}
inline void DiskLight::genRndPhtnRay(uint32_t photon, uint32_t li, uint32_t& draw, Vec3& org, Vec3& dir) {   // :229-242
  double prob, angle;
  do { angle = urange(u01(scene->opt.seed, STREAM_PHOTON, photon, li, 0, draw), 0, M_PI); prob = getAngleProb(angle, 0, M_PI, M_PI);
       double chk = urange(u01(scene->opt.seed, STREAM_PHOTON, photon, li, 0, draw + 1), 0, 1); draw += 2; if (!(prob > chk)) break; } while (true);
  Vec3 d = rotVecAroundAxis(orientation, surfTangent, angle); d.normalize();
  d = rotVecAroundAxis(d, orientation, urange(u01(scene->opt.seed, STREAM_PHOTON, photon, li, 0, draw), 0, TWO_PI_F));
  double ua = u01(scene->opt.seed, STREAM_PHOTON, photon, li, 0, draw + 1), ur = u01(scene->opt.seed, STREAM_PHOTON, photon, li, 0, draw + 2); draw += 3;
  dir = d; org = ctm->glbl.xfPt(diskPos(ua, ur));
}

// ---- textures
inline ImageTexture::ImageTexture(Scene* s, Shader* sh) : Texture(s, sh) {
  if (scene->glblTxtrdTop) { txtrdTop = true; top = scene->currTextureTop; }
  if (scene->glblTxtrdBtm) { txtrdBtm = true; bottom = scene->currTextureBottom; }
}
inline void ImageTexture::getTextureColor(const RayHit& hit, const Image* tex, double out[3]) {    // myTextureHandler.java:84-103
  double uv[2]; Ray tmp; tmp.scn = scene; tmp.key = hit.key;
  hit.obj->findTxtrCoords(hit.hitLoc, tex, tmp.getTime(), uv);
  double u = uv[0], v = uv[1];
  int uInt = j2i(u), vInt = j2i(v);
  long long n = (long long)tex->width * tex->height;
  long long i00 = (long long)vInt * tex->width + uInt, i10 = i00 + tex->width, i01 = i00 + 1, i11 = i10 + 1;
  auto px = [&](long long i) { if (i < 0) i = 0; if (i >= n) i = n - 1; return colorFromInt(tex->pixels[(size_t)i]); };   // Java throws out of range; clamp (documented)
  Vec3 c00 = px(i00), c10 = px(i10), c01 = px(i01), c11 = px(i11);
  double fu = u - uInt, fv = v - vInt;
  Vec3 c0 = interpColor(c00, fu, c01), c1 = interpColor(c10, fu, c11), c = interpColor(c0, fv, c1);
  out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
inline NoiseTexture::NoiseTexture(Scene* s, Shader* sh, double scl) : Texture(s, sh), scale(scl) {
  colors = s->noiseColors; numOctaves = s->numOctaves; turbMult = s->turbMult; periodMult = s->pdMult;
  colorScale = s->colorScale; colorMult = s->colorMult; rndColors = s->rndColors; useFwdTrans = s->useFwdTrans;
}
inline CellularTexture::CellularTexture(Scene* s, Shader* sh, double scl) : NoiseTexture(s, sh, scl) {
  avgNumPerCell = s->avgNumPerCell; mortarThresh = s->mortarThresh; numPtsDist = s->numPtsDist; roiFunc = s->roiFunc; distFunc = s->distFunc;
  double lastDist = 1.0 / std::pow(M_E, avgNumPerCell), cumProb = lastDist;
  for (int i = 1; i < 15; ++i) { lastDist *= (avgNumPerCell / (1.0 * i)); cumProb += lastDist; pdfs.push_back({cumProb, i}); }
}
inline Texture* Scene::getCurTexture(Shader* sh) {
  switch (txtrType) {
    case 1: return new ImageTexture(this, sh);
    case 2: return new NoiseTexture(this, sh, noiseScale);
    case 3: return new BaseWoodTexture(this, sh, noiseScale);
    case 4: return new MarbleTexture(this, sh, noiseScale);
    case 5: return new CellularTexture(this, sh, noiseScale);
    case 6: return new WoodTexture(this, sh, noiseScale);
    default: return new NonTexture(this, sh);
  }
}
inline Shader* Scene::getCurShader() { Shader* sh = new Shader(this, simpleRefr); sh->txtr = getCurTexture(sh); sh->serial = shaderSerialCnt++; allShaders.push_back(sh); return sh; }

// ---- shader
inline Shader::Shader(Scene* s, bool simple_) : scene(s), simple(simple_) {      // setCurrColors, myObjShader.java:51-75
  diffuseColor = s->currDiffuseColor; avgDiffClr = (1.0 / 3.0) * (diffuseColor.x + diffuseColor.y + diffuseColor.z);
  if (avgDiffClr != 0) phtnDiffScl = Vec3(diffuseColor.x / avgDiffClr, diffuseColor.y / avgDiffClr, diffuseColor.z / avgDiffClr);
  ambientColor = s->currAmbientColor; specularColor = s->currSpecularColor;
  avgSpecClr = (1.0 / 3.0) * (specularColor.x + specularColor.y + specularColor.z);
  if (avgSpecClr != 0) phtnSpecScl = Vec3(specularColor.x / avgSpecClr, specularColor.y / avgSpecClr, specularColor.z / avgSpecClr);
  KRefl = s->currKRefl; KReflClr = s->currKReflClr; KTrans = s->currKTrans; curPermClr = s->globCurPermClr;
  avgPermClr = (1.0 / 3.0) * (curPermClr.x + curPermClr.y + curPermClr.z);
  if (avgPermClr != 0) phtnPermClr = Vec3(curPermClr.x / avgPermClr, curPermClr.y / avgPermClr, curPermClr.z / avgPermClr);
  currPerm = s->globRfrIdx;
  hasCaustic = ((KRefl > 0.0) || (currPerm > 0.0) || (KTrans > 0.0));
  usePhotonMap = s->usePhotonMap; isCausticPhtn = s->isCausticPhtn;
  diffConst = 1 - currPerm; phongExp = s->currPhongExp;
}
inline void Shader::calcShadowColor(const RayHit& hit, const double tex[3], double out[3]) {
  double r = 0, g = 0, b = 0; const Vec3& hitLoc = hit.fwdTransHitLoc;
  uint32_t li = 0;
  for (Geom* lo : scene->lightList) {
    Light* light = lo->isLight() ? static_cast<Light*>(lo) : static_cast<Light*>(static_cast<Instance*>(lo)->obj);
    SampleCtx sc{hit.key.stream, hit.key.a, hit.key.b, hit.key.c, DIM_LIGHT_BASE + DIM_LIGHT_STRIDE * li};
    SKey sk = hit.key; sk.timeDim = sc.dim + 4; ++li;
    Vec3 lightNorm = light->ctm->glbl.xfPt(light->sampleOrigin(sc));
    lightNorm.sub(hitLoc); lightNorm.normalize();
    Ray shadowRay(scene, hitLoc, lightNorm, hit.gen + 1, sk);
    double t, ltMult; light->lightHit(shadowRay, sc, t, ltMult);
    if (ltMult == 0) continue;
    ++scene->stats.shadow;
    int blocked = scene->calcShadow(shadowRay, t);
    if (blocked == 0) {
      shadowRay.direction.normalize();
      double ld = shadowRay.direction.dot(hit.objNorm) * ltMult;
      if (ld > EPS) { r += tex[0] * light->lightColor.x * ld; g += tex[1] * light->lightColor.y * ld; b += tex[2] * light->lightColor.z * ld; }
      if (phongExp == 0) continue;
      Vec3 hN(shadowRay.direction); hN.sub(hit.fwdTransRayDir); hN.normalize();
      double hd = hN.dot(hit.objNorm) * ltMult;
      if (hd > EPS) { double ph = std::pow(hd * hd, phongExp); r += specularColor.x * light->lightColor.x * ph; g += specularColor.y * light->lightColor.y * ph; b += specularColor.z * light->lightColor.z * ph; }
    }
  }
  out[0] = r; out[1] = g; out[2] = b;
}
inline Shader::Fres Shader::fresnel(const RayHit& hit, double matIdx, double rayIdx) const {
  Fres f; f.backToEye = hit.fwdTransRayDir; f.backToEye.mult(-1);
  double n1 = 0, n2 = 0; f.n = 1; f.transReflRatio = 0; f.oneM = 1; f.refractNormMult = 1.0; f.TIR = false; f.cosTheta2 = 0;
  f.N = hit.objNorm; f.cosTheta1 = f.backToEye.dot(f.N);
  if (f.cosTheta1 < EPS) { f.refractNormMult = -1.0; f.N.mult(-1); }
  f.cosTheta1 = f.backToEye.dot(f.N);
  double thetaIncident = std::acos(f.backToEye.dot(f.N) / (f.backToEye.mag() * f.N.mag()));     // DistRayTracer._angleBetween :445-452
  if (f.refractNormMult < 0) {
    double thetaCrit = std::asin(1 / matIdx);
    if (thetaIncident < thetaCrit) { n1 = matIdx; n2 = 1; f.n = (n1 / n2); f.cosTheta2 = std::sqrt(1.0 - (f.n * f.n) * (1.0 - (f.cosTheta1 * f.cosTheta1))); }
    else { f.transReflRatio = 1; f.oneM = 1 - f.transReflRatio; f.TIR = true; f.cosTheta2 = 0; }
  } else { n1 = rayIdx; n2 = matIdx; f.n = (n1 / n2); f.cosTheta2 = std::sqrt(1.0 - (f.n * f.n) * (1.0 - (f.cosTheta1 * f.cosTheta1))); }
  if (!f.TIR) {
    double sinAcos = std::sin(std::acos(f.cosTheta1)), resCosThetT = std::sqrt(1.0 - ((n1 / n2) * sinAcos * sinAcos));
    double rPerp = fresPerp(n1, n2, f.cosTheta1, resCosThetT), rPar = fresPlel(n1, n2, f.cosTheta1, resCosThetT);
    f.transReflRatio = (rPerp + rPar) / 2.0; f.oneM = 1 - f.transReflRatio;
  }
  return f;
}
inline void Shader::calcTransClr(const RayHit& hit, double out[3]) {
  double r = 0, g = 0, b = 0; Fres f = fresnel(hit, KTrans, hit.rayKTrans[0]); const Vec3& permClr = curPermClr;
  if (f.oneM > EPS) {
    SKey k = hit.key; k.c = hit.key.c * 2 + 1; k.timeDim = DIM_TIME;
    Ray rr(scene, hit.fwdTransHitLoc, refractDir(f), hit.gen + 1, k); rr.setCurrKTrans(KTrans, currPerm, curPermClr);
    ++scene->stats.refract; Vec3 c = scene->reflectRay(rr);
    r += (f.oneM) * permClr.x * (c.x); g += (f.oneM) * permClr.y * (c.y); b += (f.oneM) * permClr.z * (c.z);
  }
  if (f.transReflRatio > EPS) {
    Vec3 rd = compReflDir(f.backToEye, f.N); rd.mult(f.refractNormMult);
    SKey k = hit.key; k.c = hit.key.c * 2; k.timeDim = DIM_TIME;
    Ray rr(scene, hit.fwdTransHitLoc, rd, hit.gen + 1, k); rr.setCurrKTrans(KTrans, currPerm, curPermClr);
    ++scene->stats.reflect; Vec3 c = scene->reflectRay(rr);
    r += (f.transReflRatio) * permClr.x * (c.x); g += (f.transReflRatio) * permClr.y * (c.y); b += (f.transReflRatio) * permClr.z * (c.z);
  }
  out[0] = r; out[1] = g; out[2] = b;
}
inline void Shader::calcSimpleTransClr(const RayHit& hit, double out[3]) {
  double r = 0, g = 0, b = 0; Fres f = fresnel(hit, currPerm, hit.rayKTrans[1]);
  Vec3 reflDir = compReflDir(f.backToEye, f.N);
  if (f.oneM > 0) {
    SKey k = hit.key; k.c = hit.key.c * 2 + 1; k.timeDim = DIM_TIME;
    Ray rr(scene, hit.fwdTransHitLoc, refractDir(f), hit.gen + 1, k); rr.setCurrKTrans(KTrans, currPerm, curPermClr);
    ++scene->stats.refract; Vec3 c = scene->reflectRay(rr); double m = f.oneM * KTrans;
    r += m * (c.x); g += m * (c.y); b += m * (c.z);
  }
  if (f.transReflRatio > 0) {
    reflDir.mult(f.refractNormMult);
    SKey k = hit.key; k.c = hit.key.c * 2; k.timeDim = DIM_TIME;
    Ray rr(scene, hit.fwdTransHitLoc, reflDir, hit.gen + 1, k);
    ++scene->stats.reflect; Vec3 c = scene->reflectRay(rr); double m = f.transReflRatio * KRefl;
    r += m * (c.x); g += m * (c.y); b += m * (c.z);
  }
  out[0] = r; out[1] = g; out[2] = b;
}
inline void Shader::calcReflClr(const RayHit& hit, double out[3]) {
  double r = 0, g = 0, b = 0; Vec3 back(hit.fwdTransRayDir); back.mult(-1);
  Vec3 rd = compReflDir(back, hit.objNorm);
  if (rd.dot(hit.objNorm) >= 0) {
    SKey k = hit.key; k.c = hit.key.c * 2; k.timeDim = DIM_TIME;
    Ray rr(scene, hit.fwdTransHitLoc, rd, hit.gen + 1, k);
    ++scene->stats.reflect; Vec3 c = scene->reflectRay(rr);
    r += (KReflClr.x * c.x); g += (KReflClr.y * c.y); b += (KReflClr.z * c.z);
  }
  out[0] = r; out[1] = g; out[2] = b;
}
inline void Shader::getIrradianceFromPhtnTree(const RayHit& hit, double res[3]) {
  res[0] = res[1] = res[2] = 0; if (!scene->photonTree) return;
  std::vector<KDTree::Near> hood; scene->photonTree->find_near(hit.fwdTransHitLoc.x, hit.fwdTransHitLoc.y, hit.fwdTransHitLoc.z, hood);
  if (hood.empty()) return;
  double rSq = hood[0].d2, area = PI_F * rSq;
  for (auto& n : hood) { res[0] += n.p->pwr[0]; res[1] += n.p->pwr[1]; res[2] += n.p->pwr[2]; }
  res[0] /= area; res[1] /= area; res[2] /= area;
}
inline Vec3 Shader::getColorAtPos(const RayHit& hit) {
  double r = ambientColor.x, g = ambientColor.y, b = ambientColor.z;
  if (!simple && (KRefl == 0.0) && usePhotonMap) {
    double irr[3]; getIrradianceFromPhtnTree(hit, irr);
    if (isCausticPhtn) { r += irr[0]; g += irr[1]; b += irr[2]; }
    else { r += diffuseColor.x * irr[0]; g += diffuseColor.y * irr[1]; b += diffuseColor.z * irr[2]; }
  }
  double tex[3], sh[3]; txtr->getDiffTxtrColor(hit, diffuseColor, simple ? 1.0 : diffConst, tex);
  calcShadowColor(hit, tex, sh); r += sh[0]; g += sh[1]; b += sh[2];
  if ((hit.gen < scene->numRays - 2) && hasCaustic) {
    double res[3] = {0, 0, 0};
    if (!simple) { if ((KTrans > 0) || (currPerm > 0.0)) calcTransClr(hit, res); else if (KRefl > 0.0) calcReflClr(hit, res); }
    else { if (KTrans > 0) calcSimpleTransClr(hit, res); else if (KRefl > 0.0) calcReflClr(hit, res); }
    r += res[0]; g += res[1]; b += res[2];
  }
  return mkColor(r, g, b);
}
// photon path: next ray off a specular surface; scales hit.phtnPwr in place (:461-478, :297-406)
inline bool Shader::findCausticRayHit(RayHit& hit, Ray& out) {
  if (!((hit.gen < scene->numPhotonRays) && hasCaustic)) return false;
  double pm[3] = {1.0, 1.0, 1.0}; bool have = false;
  SKey k = hit.key; k.c = hit.key.c + 1;                  // photon stream: c = segment index
  if ((KTrans > 0.0) || (currPerm > 0.0)) {
    pm[0] = phtnPermClr.x; pm[1] = phtnPermClr.y; pm[2] = phtnPermClr.z;
    Fres f = fresnel(hit, KTrans, hit.rayKTrans[0]);
    if (f.oneM > EPS) { out = Ray(scene, hit.fwdTransHitLoc, refractDir(f), hit.gen + 1, k); }
    else { Vec3 rd = compReflDir(f.backToEye, f.N); rd.mult(f.refractNormMult); out = Ray(scene, hit.fwdTransHitLoc, rd, hit.gen + 1, k); }
    out.setCurrKTrans(KTrans, currPerm, curPermClr); have = true;
  } else if (KRefl > 0.0) {
    pm[0] = pm[1] = pm[2] = KRefl;
    Vec3 back(hit.fwdTransRayDir); back.mult(-1);
    out = Ray(scene, hit.fwdTransHitLoc, compReflDir(back, hit.objNorm), hit.gen + 1, k); have = true;
  }
  for (int i = 0; i < 3; ++i) hit.phtnPwr[i] = hit.phtnPwr[i] * pm[i];
  return have;
}

// ---- scene queries
inline Vec3 Scene::reflectRay(Ray& ray) {                                      // myScene.java:907-914
  RayHit h = findClosestRayHit(ray);
  if (h.isHit) return h.shdr->getColorAtPos(h);
  if (glblTxtrdBkg) return getBackgroundTextureColor(ray);
  return backgroundColor;
}
inline Vec3 Scene::getBackgroundTextureColor(const Ray& ray) {
  Sphere* sd = mySkyDome; double t = -DMAX;
  double a = sd->getAVal(ray), b = sd->getBVal(ray), c = sd->getCVal(ray), discr = ((b * b) - (4 * a * c));
  if (discr > 0) { double d1 = std::sqrt(discr), t1 = (-1 * b + d1) / (2 * a), t2 = (-1 * b - d1) / (2 * a), tv = jmin(t1, t2); if (tv < EPS) tv = jmax(t1, t2); t = tv; }
  Vec3 p = ray.pointOnRay(t); const Image* tex = skyTex;
  double a0 = p.y - sd->origin.y, a1 = a0 / (sd->radY); a1 = (a1 > 1) ? 1 : (a1 < -1) ? -1 : a1;
  double v = (tex->height - 1) * std::acos(a1) / M_PI;
  double shWm1 = tex->width - 1, z1 = (p.z - sd->origin.z), q = v / (tex->height - 1);
  double b0 = (p.x - sd->origin.x) / (sd->radX); b0 = (b0 > 1) ? 1 : (b0 < -1) ? -1 : b0;
  double b1 = std::sin(q * M_PI), b2 = (std::fabs(b1) < EPS) ? 1 : b0 / b1;
  double u = (z1 <= EPS) ? ((shWm1 * (std::acos(b2)) / (TWO_PI_F)) + shWm1 / 2.0f) : shWm1 - ((shWm1 * (std::acos(b2)) / (TWO_PI_F)) + shWm1 / 2.0f);
  u = (u < 0) ? 0 : (u > shWm1) ? shWm1 : u;
  long long idx = (long long)j2i(v) * tex->width + j2i(u), n = (long long)tex->width * tex->height;
  if (idx < 0) idx = 0; if (idx >= n) idx = n - 1;
  return colorFromInt(tex->pixels[(size_t)idx]);
}

// ---- photon emission
inline void Scene::sendCausticPhotons() {                                     // myScene.java:952-998
  double pwrMult = causticsLightPwrMult / photonTree->num_Cast; uint32_t li = 0;
  for (Geom* lg : lightList) {
    Light* L = lg->isLight() ? static_cast<Light*>(lg) : static_cast<Light*>(static_cast<Instance*>(lg)->obj);
    for (int i = 0; i < photonTree->num_Cast; ++i) {
      double pw[3] = {L->lightColor.x * pwrMult, L->lightColor.y * pwrMult, L->lightColor.z * pwrMult};
      uint32_t draw = 0; Vec3 o, d; L->genRndPhtnRay((uint32_t)i, li, draw, o, d);
      SKey k; k.stream = STREAM_PHOTON; k.a = (uint32_t)i; k.b = li; k.c = 0; k.timeDim = 0xFFFF;
      Ray ray(this, o, d, 0, k); ++stats.photonSeg;
      RayHit h = findClosestRayHit(ray);
      if (!h.isHit || !h.shdr->hasCaustic) continue;
      for (int q = 0; q < 3; ++q) h.phtnPwr[q] = pw[q];
      Ray rr; bool have;
      do {
        have = h.shdr->findCausticRayHit(h, rr);
        if (have) { double tp[3] = {h.phtnPwr[0], h.phtnPwr[1], h.phtnPwr[2]}; ++stats.photonSeg; h = findClosestRayHit(rr); for (int q = 0; q < 3; ++q) h.phtnPwr[q] = tp[q]; }
        else h.isHit = false;
      } while (h.isHit && h.shdr->hasCaustic && (rr.gen <= numPhotonRays));
      if (!h.isHit || (rr.gen > numPhotonRays)) continue;
      Photon* p = new Photon{{h.phtnPwr[0], h.phtnPwr[1], h.phtnPwr[2]}, {h.fwdTransHitLoc.x, h.fwdTransHitLoc.y, h.fwdTransHitLoc.z, 0}};
      photonTree->add_photon(p); ++stats.photonsStored;
    }
    ++li;
  }
  photonTree->build_tree();
}
inline void Scene::sendDiffusePhotons() {                                     // :1000-1091
  double pwrMult = diffuseLightPwrMult / photonTree->num_Cast; uint32_t li = 0;
  for (Geom* lg : lightList) {
    Light* L = lg->isLight() ? static_cast<Light*>(lg) : static_cast<Light*>(static_cast<Instance*>(lg)->obj);
    for (int i = 0; i < photonTree->num_Cast; ++i) {
      double pw[3] = {L->lightColor.x * pwrMult, L->lightColor.y * pwrMult, L->lightColor.z * pwrMult};
      uint32_t draw = 0; Vec3 o, d; L->genRndPhtnRay((uint32_t)i, li, draw, o, d);
      SKey k; k.stream = STREAM_PHOTON; k.a = (uint32_t)i; k.b = li; k.c = 0; k.timeDim = 0xFFFF;
      Ray ray(this, o, d, 0, k); ++stats.photonSeg;
      RayHit h = findClosestRayHit(ray);
      if (!h.isHit) continue;
      for (int q = 0; q < 3; ++q) h.phtnPwr[q] = pw[q];
      bool done = false, firstDiff = true;
      do {
        if (h.shdr->KRefl == 0) {
          double prob = 0; uint32_t seg = h.key.c + 1, dr = 0;   // draws for this bounce live in segment (c+1)
          if (!firstDiff) {
            Photon* p = new Photon{{h.phtnPwr[0], h.phtnPwr[1], h.phtnPwr[2]}, {h.fwdTransHitLoc.x, h.fwdTransHitLoc.y, h.fwdTransHitLoc.z, 0}};
            photonTree->add_photon(p); ++stats.photonsStored;
            prob = urange(u01(opt.seed, STREAM_PHOTON, (uint32_t)i, li, seg, dr++), 0, 1.0);
          }
          firstDiff = false;
          if (prob < h.shdr->avgDiffClr) {
            double x = 0, y = 0, z = 0, sq;
            do { x = urange(u01(opt.seed, STREAM_PHOTON, (uint32_t)i, li, seg, dr), -1.0, 1.0); y = urange(u01(opt.seed, STREAM_PHOTON, (uint32_t)i, li, seg, dr + 1), -1.0, 1.0); dr += 2; sq = (x * x) + (y * y); } while ((sq >= 1.0) || (sq < EPS));
            z = std::sqrt(1 - (sq));
            Vec3 n(h.objNorm); double nxSq = n.x * n.x, nySq = n.y * n.y, nzSq = n.z * n.z;
            Vec3 tmpV = ((nxSq > nySq) && (nxSq > nzSq)) ? Vec3(0, 0, 1) : Vec3(1, 0, 0);
            Vec3 _p = n.cross(tmpV), _q = _p.cross(n);
            n.mult(z); _p.mult(x); _q.mult(y);
            Vec3 bd(n.x + _p.x + _q.x, n.y + _p.y + _q.y, n.z + _p.z + _q.z); bd.normalize();
            double tp[3] = {h.phtnPwr[0] * h.shdr->phtnDiffScl.x, h.phtnPwr[1] * h.shdr->phtnDiffScl.y, h.phtnPwr[2] * h.shdr->phtnDiffScl.z};
            SKey nk = h.key; nk.c = seg;
            Ray rr(this, h.fwdTransHitLoc, bd, h.gen + 1, nk); ++stats.photonSeg;
            h = findClosestRayHit(rr); for (int q = 0; q < 3; ++q) h.phtnPwr[q] = tp[q];
          } else done = true;
        } else {
          Ray rr; bool have = h.shdr->findCausticRayHit(h, rr);
          if (have) { double tp[3] = {h.phtnPwr[0], h.phtnPwr[1], h.phtnPwr[2]}; ++stats.photonSeg; h = findClosestRayHit(rr); for (int q = 0; q < 3; ++q) h.phtnPwr[q] = tp[q]; }
          else h.isHit = false;
        }
      } while (h.isHit && !done && (h.gen <= numPhotonRays));
      // (the reference stores nothing after the loop, :1075-1079)
    }
    ++li;
  }
  photonTree->build_tree();
}

// ---- render drivers
inline Vec3 Scene::tracePrimary(Ray& ray, PixelOut* aov) {
  ++stats.primary;
  uint64_t b0 = stats.boxTests, p0 = stats.primTests;
  RayHit h = findClosestRayHit(ray);
  stats.boxTestsPrimary += stats.boxTests - b0; stats.primTestsPrimary += stats.primTests - p0;
  if (aov) { aov->hitPrim = h.isHit ? h.obj->primSerial : -1; aov->hitInst = h.isHit ? h.instSerial : -1; aov->t = h.isHit ? h.t : 0; }
  if (h.isHit) return h.shdr->getColorAtPos(h);
  if (glblTxtrdBkg) return getBackgroundTextureColor(ray);
  return backgroundColor;
}
inline Scene::PixelOut Scene::renderPixel(int row, int col) {
  PixelOut po; uint32_t pix = (uint32_t)(row * sceneCols + col);
  auto U = [&](uint32_t smp, uint32_t dim) { return u01(opt.seed, STREAM_PIXEL, pix, smp, 1, dim); };
  auto key = [&](uint32_t smp) { SKey k; k.stream = STREAM_PIXEL; k.a = pix; k.b = smp; k.c = 1; k.timeDim = DIM_TIME; return k; };
  int n = numRaysPerPixel;
  if (kind == SC_FOV) {
    double rayY = (-1 * (row - rayYOffset)), rayX = col - rayXOffset;
    if (hasDpthOfFld) {                                                    // :1386-1406, :868-875
      double rv = 0, gv = 0, bv = 0;
      Vec3 lensCtr(rayX, rayY, viewZ); lensCtr.normalize();
      Ray ray(this, eyeOrigin, lensCtr, 0, key(0));
      Ray tr = ray.getTransformedRay(ray, focalPlane->ctm->inv);
      RayHit fh = focalPlane->intersectCheck(ray, tr, focalPlane->ctm);
      Vec3 focalPt = fh.hitLoc;
      for (int s = 0; s < n; ++s) {
        Vec3 tmp = rotVecAroundAxis(Vec3(0, 1, 0), Vec3(0, 0, -1), urange(U(s, DIM_LENS_ANGLE), 0, TWO_PI_F)); tmp.normalize();
        tmp.mult(urange(U(s, DIM_LENS_RADIUS), 0, lens_radius)); tmp.add(lensCtr);
        Ray r2(this, tmp, Vec3(tmp, focalPt), 0, key(s));
        Vec3 c = tracePrimary(r2, s == 0 ? &po : nullptr); rv += c.x; gv += c.y; bv += c.z;
      }
      po.rgb = mkColor(rv / n, gv / n, bv / n);
    } else if (n == 1) {
      Ray ray(this, eyeOrigin, Vec3(rayX, rayY, viewZ), 0, key(0)); po.rgb = tracePrimary(ray, &po);
    } else {
      double rv = 0, gv = 0, bv = 0;
      for (int s = 0; s < n; ++s) {
        double ry = rayY + urange(U(s, DIM_AA_Y), -.5, .5), rx = rayX + urange(U(s, DIM_AA_X), -.5, .5);
        Ray ray(this, eyeOrigin, Vec3(rx, ry, viewZ), 0, key(s));
        Vec3 c = tracePrimary(ray, s == 0 ? &po : nullptr); rv += c.x; gv += c.y; bv += c.z;
      }
      po.rgb = mkColor(rv / n, gv / n, bv / n);
    }
  } else if (kind == SC_FISHEYE) {                                          // :1563-1643
    if (n == 1) {
      double yVal = (row + yStart) * fishMult, ySq = yVal * yVal, xVal = (col + xStart) * fishMult, rTmp = xVal * xVal + ySq;
      if (rTmp > 1) po.rgb = mkColor(0, 0, 0);
      else { double r = std::sqrt(rTmp), theta = r * aperatureHlf, phi = std::atan2(-yVal, xVal), sTh = std::sin(theta);
        Ray ray(this, eyeOrigin, Vec3(sTh * std::cos(phi), sTh * std::sin(phi), -std::cos(theta)), 0, key(0)); po.rgb = tracePrimary(ray, &po); }
    } else {
      double yB = (row + yStart), xB = (col + xStart), rv = 0, gv = 0, bv = 0;
      for (int s = 0; s < n; ++s) {
        double yVal = (yB + urange(U(s, DIM_AA_Y), -.5, .5)) * fishMult, xVal = (xB + urange(U(s, DIM_AA_X), -.5, .5)) * fishMult, rSq = yVal * yVal + xVal * xVal;
        if (rSq <= 1) { double r = std::sqrt(rSq), theta = r * aperatureHlf, phi = std::atan2(-yVal, xVal), sTh = std::sin(theta);
          Ray ray(this, eyeOrigin, Vec3(sTh * std::cos(phi), sTh * std::sin(phi), -std::cos(theta)), 0, key(s));
          Vec3 c = tracePrimary(ray, s == 0 ? &po : nullptr); rv += c.x; gv += c.y; bv += c.z; }
      }
      po.rgb = mkColor(rv / n, gv / n, bv / n);
    }
  } else {                                                                 // ortho :1687-1753
    double ryo = sceneRows / 2.0, rxo = sceneCols / 2.0;
    if (n == 1) {
      double rayY = orthPerRow * (-1 * (row - ryo)), rayX = orthPerCol * (col - rxo);
      Ray ray(this, Vec3(rayX, rayY, 0), Vec3(0, 0, -1), 0, key(0)); po.rgb = tracePrimary(ray, &po);
    } else {
      double yB = orthPerRow * ((-1 * (row - ryo)) - .5), xB = orthPerCol * (col - rxo - .5), rv = 0, gv = 0, bv = 0;
      for (int s = 0; s < n; ++s) {
        double ry = yB + (orthPerRow * urange(U(s, DIM_AA_Y), -.5, .5)), rx = xB + (orthPerCol * urange(U(s, DIM_AA_X), -.5, .5));
        Ray ray(this, Vec3(rx, ry, 0), Vec3(0, 0, -1), 0, key(s));
        Vec3 c = tracePrimary(ray, s == 0 ? &po : nullptr); rv += c.x; gv += c.y; bv += c.z;
      }
      po.rgb = mkColor(rv / n, gv / n, bv / n);
    }
  }
  po.argb = colorGetInt(po.rgb);
  return po;
}

// ---- images
inline const Image* Scene::loadImage(const std::string& name) {
  auto it = imageCache.find(name); if (it != imageCache.end()) return it->second;
  Image* im = nullptr;
  if (imageLoader) im = imageLoader(name);
  if (!im) {   // decoded texture cache: <texDir>/<name>.argb = "ARGB" int32 w, int32 h, w*h int32 (little endian)
    std::string p = opt.texDir + "/" + name + ".argb"; FILE* f = fopen(p.c_str(), "rb");
    if (!f) throw std::runtime_error("texture not found: " + p);
    char mg[4]; int32_t wh[2];
    if (fread(mg, 1, 4, f) != 4 || memcmp(mg, "ARGB", 4) != 0 || fread(wh, 4, 2, f) != 2) { fclose(f); throw std::runtime_error("bad texture file " + p); }
    im = new Image; im->width = wh[0]; im->height = wh[1]; im->pixels.resize((size_t)wh[0] * wh[1]);
    if (fread(im->pixels.data(), 4, im->pixels.size(), f) != im->pixels.size()) { fclose(f); throw std::runtime_error("short texture file " + p); }
    fclose(f);
  }
  imageCache[name] = im; return im;
}

// ===========================================================================
// .cli interpreter (myRTFileReader.java:15-378)
// ===========================================================================
struct TokErr {};                          // stands for ArrayIndexOutOfBounds / NumberFormatException
struct Tok {
  std::vector<std::string> t;
  const std::string& s(size_t i) const { if (i >= t.size()) throw TokErr(); return t[i]; }
  double d(size_t i) const { const std::string& x = s(i); char* e = nullptr; double v = strtod(x.c_str(), &e); if (e == x.c_str() || (*e && !(((*e == 'f') || (*e == 'F') || (*e == 'd') || (*e == 'D')) && !e[1]))) throw TokErr(); return v; }
  float f(size_t i) const { return (float)d(i); }     // Float.parseFloat rounds the decimal directly; double rounding differs only in rare ties
  int i(size_t k) const { const std::string& x = s(k); size_t p = 0; if (x.empty()) throw TokErr(); if (x[0] == '-' || x[0] == '+') p = 1; if (p >= x.size()) throw TokErr(); for (size_t q = p; q < x.size(); ++q) if (x[q] < '0' || x[q] > '9') throw TokErr(); return (int)strtol(x.c_str(), nullptr, 10); }
  Vec3 color(size_t st) const { return mkColor(d(st), d(st + 1), d(st + 2)); }
};
inline std::string lower(std::string s) { for (auto& c : s) c = (char)tolower(c); return s; }

}  // namespace orc
#include "orc_ext.hpp"
namespace orc {
inline void Scene::readPrimData(const std::vector<std::string>& tv) {           // myScene.java:447-521
  Tok k{tv}; Geom* tmp = nullptr; const std::string& c = tv[0];
  if (c == "box") {
    double minX = jmin(k.d(1), k.d(4)), maxX = jmax(k.d(1), k.d(4)), ctrX = (minX + maxX) * .5, minY = jmin(k.d(2), k.d(5)), maxY = jmax(k.d(2), k.d(5)), ctrY = (minY + maxY) * .5,
           minZ = jmin(k.d(3), k.d(6)), maxZ = jmax(k.d(3), k.d(6)), ctrZ = (minZ + maxZ) * .5;
    tmp = new RndrdBox(this, ctrX, ctrY, ctrZ, Vec3(minX, minY, minZ), Vec3(maxX, maxY, maxZ));
  } else if (c == "plane") { Plane* p = new Plane(this); p->setPlaneVals(k.d(1), k.d(2), k.d(3), k.d(4)); tmp = p; }
  else if (c == "cyl") {
    double rad = k.d(1), h = k.d(2), xC = k.d(3), yC = k.d(4), zC = k.d(5), xO = 0, yO = 1, zO = 0;
    try { xO = k.d(6); yO = k.d(7); zO = k.d(8); } catch (TokErr&) {}
    tmp = new Cylinder(this, rad, h, xC, yC, zC, xO, yO, zO);
  } else if (c == "cylinder") { double rad = k.d(1), xC = k.d(2), zC = k.d(3), yMin = k.d(4), yMax = k.d(5); tmp = new Cylinder(this, rad, yMax - yMin, xC, yMin, zC, 0, 1, 0); }
  else if (c == "hollow_cylinder") { double rad = k.d(1), xC = k.d(2), zC = k.d(3), yMin = k.d(4), yMax = k.d(5); tmp = new HollowCylinder(this, rad, yMax - yMin, xC, yMin, zC); }
  else if (c == "sphere") { double r = k.d(1); tmp = new Sphere(this, r, r, r, k.d(2), k.d(3), k.d(4)); }
  else if (c == "moving_sphere") tmp = new MovingSphere(this, k.d(1), k.d(2), k.d(3), k.d(4), k.d(5), k.d(6), k.d(7));
  else if (c == "sphereIn") { double r = k.d(1); tmp = new Sphere(this, r, r, r, k.d(2), k.d(3), k.d(4)); tmp->inverted = true; }
  else if (c == "ellipsoid") tmp = new Sphere(this, k.d(1), k.d(2), k.d(3), k.d(4), k.d(5), k.d(6));
  else if (c == "torus") { const bool six = tv.size() >= 7; tmp = new Torus(this, k.d(1), k.d(2), k.d(six ? 4 : 3), k.d(six ? 5 : 4), k.d(six ? 6 : 5)); }      // extension, see orc_ext.hpp
  else if (c == "quadric") {
    double coef[10], bx[6] = {-100000, -100000, -100000, 100000, 100000, 100000};
    for (int i = 0; i < 10; ++i) coef[i] = k.d(1 + i);
    if (tv.size() >= 17) for (int i = 0; i < 6; ++i) bx[i] = k.d(11 + i);
    for (int i = 0; i < 3; ++i) if (bx[i] > bx[3 + i]) std::swap(bx[i], bx[3 + i]);
    tmp = new Quadric(this, coef, bx);
  }
  else return;
  tmp->primSerial = primSerialCnt++;
  tmp->shdr = getCurShader(); addObjectToScene(tmp);
}
inline void Scene::setTxtrColor(const std::vector<std::string>& tv) {            // :604-639
  Tok k{tv};
  if (!useCustClrs) { noiseColors.clear(); useCustClrs = true; }
  try { Vec3 c; if (k.s(1) == "named") c = getClr(k.s(2)); else c = k.color(1); noiseColors.push_back(c); } catch (TokErr&) {}
}
inline void Scene::setTexture(const std::vector<std::string>& tv) {              // :642-777
  Tok k{tv}; resetDfltTxtrVals(); const std::string& typ = tv[0];
  auto readPerlin = [&]() -> bool {
    try {
      noiseScale = k.d(1); numOctaves = k.i(2); turbMult = k.d(3); pdMult = Vec3(k.d(4), k.d(5), k.d(6));
      Vec3 py(k.d(7), k.d(8), k.d(9));
      if (py.sqMag() > 0) { py.mult(TWO_PI_F - 1.0); py.add(1.0, 1.0, 1.0); pdMult = velemMult(pdMult, py); }
      useFwdTrans = (k.d(10) == 1.0);
      try { colorScale = k.d(11); colorMult = k.d(12); rndColors = true; try { numOverlays = k.i(13); } catch (TokErr&) { numOverlays = 1; } }
      catch (TokErr&) { rndColors = false; colorScale = 25.0; colorMult = .1; numOverlays = 1; }
      return false;
    } catch (TokErr&) { return true; }
  };
  auto readWorley = [&]() -> bool {
    try {
      int p = 1; noiseScale = k.d(p++); distFunc = k.i(p++); roiFunc = k.i(p++); numPtsDist = k.i(p++); avgNumPerCell = k.d(p++); mortarThresh = k.d(p++); useFwdTrans = (k.d(p++) == 1.0);
      try { colorScale = k.d(p++); colorMult = k.d(p++); rndColors = true; try { numOverlays = k.i(p++); } catch (TokErr&) { numOverlays = 1; } }
      catch (TokErr&) { rndColors = false; colorScale = 25.0; colorMult = .1; numOverlays = 1; }
      return false;
    } catch (TokErr&) { return true; }
  };
  if (typ == "wood") { txtrType = 3; bool dflt = readPerlin(); if (!useCustClrs) noiseColors = {getClr("clr_dkwood1"), getClr("clr_ltwood1")};
    if (dflt) setProcTxtrVals(txtrType, 4, 1, 2, 1, 1, true, useCustClrs, false, 2.0, .4, 25.0, .2, 1.0, 0.05, Vec3(TWO_PI_F * 2.7, 3.6, 4.3)); }
  else if (typ == "wood2") { txtrType = 6; bool dflt = readPerlin(); if (!useCustClrs) noiseColors = {getClr("clr_dkwood2"), getClr("clr_ltwood2")};
    if (dflt) setProcTxtrVals(txtrType, 8, 1, 2, 1, 1, true, useCustClrs, false, 1.0, .4, 25.0, .3, 1.0, 0.05, Vec3(TWO_PI_F * 3.5, 7.9, 6.2)); }
  else if (typ == "marble") { txtrType = 4; bool dflt = readPerlin(); if (!useCustClrs) noiseColors = {getClr("clr_nearblack"), getClr("clr_offwhite")};
    if (dflt) setProcTxtrVals(txtrType, 16, 1, 2, 1, 1, true, useCustClrs, false, 1.0, 15.0, 24.0, .1, 1.0, 0.05, Vec3(TWO_PI_F * 0.1, TWO_PI_F * 31.4, TWO_PI_F * 4.1)); }
  else if (typ == "stone") { txtrType = 5; bool dflt = readWorley();
    if (!useCustClrs) noiseColors = {getClr("clr_mortar1"), getClr("clr_mortar2"), getClr("clr_brick1_1"), getClr("clr_brick1_2"), getClr("clr_brick2_1"), getClr("clr_brick2_2"), getClr("clr_brick3_1"), getClr("clr_brick3_2"), getClr("clr_brick4_1"), getClr("clr_brick4_2")};
    if (dflt) setProcTxtrVals(txtrType, 8, 1, 2, 1, 1, true, useCustClrs, false, 4.0, 1.0, 12.0, .2, 1.0, 0.05, Vec3(10.0, 10.0, 10.0)); }
  else { txtrType = 0; }
}

inline void Scene::readRTFile(const std::string& fileName, bool isMain) {
  std::ifstream in(opt.dataDir + "/" + fileName);
  if (!in) { warnings.push_back("File Read Error : " + fileName); if (isMain) throw std::runtime_error("cannot read scene file " + opt.dataDir + "/" + fileName); return; }
  std::string vertType = "triangle"; int myVertCount = 0; int curNumRaysPerPxl = numRaysPerPixel; Planar* myPoly = nullptr;
  std::string line;
  while (std::getline(in, line)) {
    while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
    Tok k; { size_t p = 0; while (p < line.size()) { while (p < line.size() && line[p] == ' ') ++p; size_t q = p; while (q < line.size() && line[q] != ' ') ++q; if (q > p) k.t.push_back(line.substr(p, q - p)); p = q; } }
    if (k.t.empty() || k.t[0][0] == '#') continue;
    const std::string& c = k.t[0];
    try {
    if (c == "fov") { if (!isMain) continue; numRaysPerPixel = (curNumRaysPerPxl != 0) ? curNumRaysPerPxl : 1; setSceneParamsFOV(k.d(1)); }
    else if (c == "lens") { setDpthOfFld(k.d(1), k.d(2)); }
    else if (c == "fishEye" || c == "fisheye") { if (!isMain) continue; numRaysPerPixel = (curNumRaysPerPxl != 0) ? curNumRaysPerPxl : 1; setSceneParamsFish(k.d(1)); }
    else if (c == "ortho" || c == "orthographic") { if (!isMain) continue; numRaysPerPixel = (curNumRaysPerPxl != 0) ? curNumRaysPerPxl : 1; setSceneParamsOrtho(k.d(1), k.d(2)); }
    else if (c == "write") { saveName = k.s(1); }      // rendering is driven by the caller, not the parser
    else if (c == "read") { readRTFile(k.s(1), false); }
    else if (c == "reset_timer" || c == "print_timer") {}
    else if (c == "refine") { glblRefine = lower(k.s(1)) == "on"; }
    else if (c == "rays_per_pixel") { int r = k.i(1); curNumRaysPerPxl = r; numRaysPerPixel = r; }
    else if (c == "antialias") { int prod = k.i(1) * k.i(2); curNumRaysPerPxl = prod; numRaysPerPixel = prod; }
    else if (c == "background") {
      if (k.s(1) == "texture") {
        currBkgTexture = loadImage(k.s(2)); glblTxtrdBkg = true;
        double rad = k.d(3), xC = k.d(4), yC = k.d(5), zC = k.d(6);
        mySkyDome = new Sphere(this, rad, rad, rad, xC, yC, zC); skyTex = currBkgTexture;
      } else { backgroundColor = mkColor(k.d(1), k.d(2), k.d(3)); txtrType = 0; }
    }
    else if (c == "point_light") { addObjectToScene(new PointLight(this, numLights, k.d(4), k.d(5), k.d(6), k.d(1), k.d(2), k.d(3))); }
    else if (c == "spotlight") { double inT = k.d(7), outT = k.d(8); addObjectToScene(new SpotLight(this, numLights, k.d(9), k.d(10), k.d(11), k.d(1), k.d(2), k.d(3), k.d(4), k.d(5), k.d(6), inT, outT)); }
    else if (c == "disk_light") { double rad = k.d(4); addObjectToScene(new DiskLight(this, numLights, k.d(8), k.d(9), k.d(10), k.d(1), k.d(2), k.d(3), k.d(5), k.d(6), k.d(7), rad)); }
    else if (c == "caustic_photons" || c == "diffuse_photons") {               // myScene.java:919-931
      usePhotonMap = true; isPhtnMapRndrd = false; isCausticPhtn = c.find("caustic") != std::string::npos;
      numPhotons = k.i(1); kNhood = k.i(2); ph_max_near_dist = k.f(3);
      if (opt.photonOverride >= 0) numPhotons = (int)opt.photonOverride;
      photonTree = new KDTree(numPhotons, kNhood, (double)ph_max_near_dist);
    }
    else if (c == "final_gather") { (void)k.i(1); }
    else if (c == "diffuse") { Vec3 d = k.color(1), a = k.color(4); glblTxtrdTop = false; glblTxtrdBtm = false; setSurface(d, a, Vec3(0, 0, 0), 0, 0); }
    else if (c == "shiny" || c == "surface") {                                 // setSurfaceShiny :358-378
      bool useSimple = (c == "shiny");
      Vec3 d = k.color(1), a = k.color(4), s = k.color(7); double ph = k.d(10), kr = k.d(11), kt = 0, ri = 0;
      glblTxtrdTop = false; glblTxtrdBtm = false; setSurface(d, a, s, ph, kr);
      try { kt = k.d(12); setSurface(d, a, s, ph, kr, kt); ri = k.d(13); setSurface(d, a, s, ph, kr, kt, ri); setRfrIdx(ri, ri, ri, ri); setRfrIdx(ri, k.d(14), k.d(15), k.d(16)); } catch (TokErr&) {}
      if (useSimple && ((kt > 0) || (ri > 0))) simpleRefr = true;
    }
    else if (c == "reflective") { Vec3 d = k.color(1), a = k.color(4); glblTxtrdTop = false; glblTxtrdBtm = false; double kr = k.d(7); setSurface(d, a, Vec3(0, 0, 0), 0, kr); }
    else if (c == "perm") { double v = k.d(1); setRfrIdx(v, v, v, v); try { setRfrIdx(k.d(1), k.d(2), k.d(3), k.d(4)); } catch (TokErr&) {} }
    else if (c == "phong") currPhongExp = k.d(1);
    else if (c == "krefl") { double v = k.d(1); setKRefl(v, v, v, v); }
    else if (c == "depth") currDepth = k.d(1);
    else if (c == "ktrans") currKTrans = k.d(1);
    else if (c == "begin_list") startTmpObjList();
    else if (c == "end_list") endTmpObjList(0);
    else if (c == "end_accel") endTmpObjList(1);
    else if (c == "sierpinski") {
      std::string name = k.s(1); float scale = .5f; int depth = 5; bool useShdr = false;
      try { depth = k.i(2); scale = k.f(3); (void)k.s(4); useShdr = true; } catch (TokErr&) {}
      buildSierpinski(name, scale, depth, useShdr);
    }
    else if (c == "named_object") setObjectAsNamedObject(k.s(1));
    else if (c == "instance") { std::string name = k.s(1); bool use = k.t.size() > 2; addInstance(name, use); }
    else if (c == "image_texture" || c == "texture") {
      std::string side = lower(k.s(1));
      if (side == "top" || side != "bottom") { std::string nm = (side == "top") ? k.s(2) : k.s(1); currTextureTop = loadImage(nm); glblTxtrdTop = true; }
      else { currTextureBottom = loadImage(k.s(1)); glblTxtrdBtm = true; }
      txtrType = 1;
    }
    else if (c == "noise") { double sc = k.d(1); resetDfltTxtrVals(); txtrType = 2; noiseScale = sc; }
    else if (c == "noise_color") setTxtrColor(k.t);
    else if (c == "marble" || c == "stone" || c == "wood" || c == "wood2") setTexture(k.t);
    else if (c == "begin") {
      try { vertType = k.s(1); } catch (TokErr&) {}
      myVertCount = 0;
      if (vertType == "quad") myPoly = new Planar(this, 4, G_QUAD); else myPoly = new Planar(this, 3, G_TRI);
    }
    else if (c == "texture_coord") { if (myPoly && myVertCount < myPoly->vCount) myPoly->setTxtrCoord(k.d(1), k.d(2), myVertCount); }
    else if (c == "vertex") { if (myPoly && myVertCount < myPoly->vCount) myPoly->setVert(k.d(1), k.d(2), k.d(3), myVertCount); myVertCount++; }
    else if (c == "end") {
      if (myPoly) { myPoly->finalizePoly(); myPoly->shdr = getCurShader(); myPoly->primSerial = primSerialCnt++; addObjectToScene(myPoly); }
      vertType = "triangle"; myVertCount = 0;
    }
    else if (c == "box" || c == "plane" || c == "cyl" || c == "cylinder" || c == "hollow_cylinder" || c == "sphere" || c == "moving_sphere" || c == "sphereIn" || c == "ellipsoid") readPrimData(k.t);
    else if (c == "extensions") { bool on = true; try { std::string v = k.s(1); for (auto& ch : v) ch = (char)tolower(ch); on = v != "off"; } catch (TokErr&) {} extensions = on; }
    else if ((c == "torus" || c == "quadric") && extensions) readPrimData(k.t);
    else if (c == "push") gtPushMatrix();
    else if (c == "pop") gtPopMatrix();
    else if (c == "rotate") gtRotate(k.d(1), k.d(2), k.d(3), k.d(4));
    else if (c == "scale") gtScale(k.d(1), k.d(2), k.d(3));
    else if (c == "translate") gtTranslate(k.d(1), k.d(2), k.d(3));
    else warnings.push_back("unknown command '" + c + "' in " + fileName);
    } catch (TokErr&) { throw std::runtime_error("malformed line in " + fileName + ": " + line); }
  }
}

}  // namespace orc
