// `.cli` interpreter + scene flattener (see host_scene.h for the reference map).
#include "host_scene.h"
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <chrono>

namespace drt {
static double nowMs() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }


// ---------------------------------------------------------------------------------------------
// tokens: PApplet.splitTokens(line, " ") + Java number parsing (missing token / bad number -> Missing)
// ---------------------------------------------------------------------------------------------
struct Missing {};
struct HostScene::Tokens {
  std::vector<std::string> t;
  explicit Tokens(const std::string& line) {
    size_t p = 0, n = line.size(); t.reserve(8);
    while (p < n) { while (p < n && line[p] == ' ') ++p; size_t q = p; while (q < n && line[q] != ' ') ++q; if (q > p) t.push_back(line.substr(p, q - p)); p = q; }
  }
  size_t size() const { return t.size(); }
  const std::string& str(size_t i) const { if (i >= t.size()) throw Missing(); return t[i]; }
  double num(size_t i) const {
    const std::string& s = str(i); char* e = nullptr; double v = std::strtod(s.c_str(), &e);
    if (e == s.c_str()) throw Missing();
    if (*e) { bool suffix = (*e == 'f' || *e == 'F' || *e == 'd' || *e == 'D') && e[1] == 0; if (!suffix) throw Missing(); }
    return v;
  }
  int integer(size_t i) const {
    const std::string& s = str(i); size_t p = (s.size() && (s[0] == '-' || s[0] == '+')) ? 1 : 0;
    if (p >= s.size()) throw Missing();
    for (size_t q = p; q < s.size(); ++q) if (s[q] < '0' || s[q] > '9') throw Missing();
    return (int)std::strtol(s.c_str(), nullptr, 10);
  }
  V3 rgb(size_t i) const { return v3(jmin(1, num(i)), jmin(1, num(i + 1)), jmin(1, num(i + 2))); }   // myColor clamps to <= 1
};
static std::string lowered(std::string s) { for (auto& c : s) c = (char)tolower(c); return s; }
static V3 clampColor(V3 c) { return v3(jmin(1, c.x), jmin(1, c.y), jmin(1, c.z)); }

static V3 namedColor(const std::string& nameIn) {               // DistRayTracer.java:467-530
  struct E { const char* k; double r, g, b; };
  static const E tab[] = {{"clr_gray",0.47,0.47,0.47},{"clr_white",1,1,1},{"clr_yellow",1,1,0},{"clr_cyan",0,1,1},{"clr_magenta",1,0,1},{"clr_red",1,0,0},{"clr_blue",0,0,1},{"clr_purple",0.6,0.2,1},{"clr_green",0,1,0},
    {"clr_ltwood1",0.94,0.47,0.12},{"clr_ltwood2",0.94,0.8,0.4},{"clr_dkwood1",0.2,0.08,0.08},{"clr_dkwood2",0.3,0.20,0.16},{"clr_mortar1",0.2,0.2,0.2},{"clr_mortar2",0.7,0.7,0.7},
    {"clr_brick1_1",0.6,0.18,0.22},{"clr_brick1_2",0.8,0.26,0.33},{"clr_brick2_1",0.6,0.32,0.16},{"clr_brick2_2",0.8,0.45,0.25},{"clr_brick3_1",0.3,0.01,0.07},{"clr_brick3_2",0.6,0.02,0.13},{"clr_brick4_1",0.4,0.1,0.17},{"clr_brick4_2",0.6,0.3,0.13},
    {"clr_darkgray",0.31,0.31,0.31},{"clr_darkred",0.47,0,0},{"clr_darkblue",0,0,0.47},{"clr_darkpurple",0.4,0.2,0.6},{"clr_darkgreen",0,0.47,0},{"clr_darkyellow",0.47,0.47,0},{"clr_darkmagenta",0.47,0,0.47},{"clr_darkcyan",0,0.47,0.47},
    {"clr_lightgray",0.78,0.78,0.78},{"clr_lightred",1,.43,.43},{"clr_lightblue",0.43,0.43,1},{"clr_lightgreen",0.43,1,0.43},{"clr_lightyellow",1,1,.43},{"clr_lightmagenta",1,.43,1},{"clr_lightcyan",0.43,1,1},
    {"clr_black",0,0,0},{"clr_nearblack",0.05,0.05,0.05},{"clr_faintgray",0.43,0.43,0.43},{"clr_faintred",0.43,0,0},{"clr_faintblue",0,0,0.43},{"clr_faintgreen",0,0.43,0},{"clr_faintyellow",0.43,0.43,0},{"clr_faintcyan",0,0.43,0.43},{"clr_faintmagenta",0.43,0,0.43},{"clr_offwhite",0.95,0.98,0.92}};
  std::string n = lowered(nameIn);
  for (const E& e : tab) if (n == e.k) return v3(e.r, e.g, e.b);
  return v3(1, 1, 1);   // unknown (and the irreproducible clr_rnd) -> white
}

// ---------------------------------------------------------------------------------------------
HostScene::HostScene(int cols, int rows) {
  std::memset(&g, 0, sizeof(g));
  g.cols = cols; g.rows = rows; g.spp = 0; g.numRays = 8; g.numPhotonRays = 4; g.seed = 0x5EED;
  g.rayYOffset = rows / 2.0; g.rayXOffset = cols / 2.0;
  double maxDim = std::max(rows, cols);
  g.yStart = ((maxDim - rows) / 2.0) - g.rayYOffset; g.xStart = ((maxDim - cols) / 2.0) - g.rayXOffset; g.fishMult = 2.0 / maxDim;
  g.causticPwrMult = 40.0; g.diffusePwrMult = 8.0; g.skyImage = -1;
  for (int i = 0; i < 10; ++i) stack_[i] = M4::ident();
  mat_.diff = mat_.amb = mat_.spec = mat_.perm = mat_.kreflClr = v3(0, 0, 0);
  noiseColors_ = {v3(.7, .7, .7), v3(.2, .2, .2)}; pdMult_ = v3(10, 10, 10);
  // default camera: FOV 60 (myRTFileReader.java:27-28)
  g.camKind = CAM_FOV;
  double fovRad = M_PI * 60.0 / 180.0; g.viewZ = -1 * (std::max(rows, cols) / 2.0) / std::tan(fovRad / 2);
}

// ---- matrix stack (myScene.java:1235-1323, myVector.java:225-256)
void HostScene::push() { if (top_ + 1 >= 10) throw std::runtime_error("matrix stack deeper than the reference's 10 allocated slots"); stack_[top_ + 1] = stack_[top_]; ++top_; }
void HostScene::pop() { if (top_ > 0) --top_; }
void HostScene::mulTop(const M4& m) { stack_[top_] = mmul(stack_[top_], m); }
void HostScene::translate(double x, double y, double z) { M4 t = M4::ident(); t.at(0, 3) = x; t.at(1, 3) = y; t.at(2, 3) = z; mulTop(t); }
void HostScene::scale(double x, double y, double z) { M4 s = M4::ident(); s.at(0, 0) = x; s.at(1, 1) = y; s.at(2, 2) = z; mulTop(s); }
void HostScene::rotate(double deg, double ax, double ay, double az) {
  double rad = (double)(deg * M_PI) / 180.0;
  V3 a = vnormOrZero(v3(ax, ay, az));
  V3 ref = (ax == 0) ? v3(1, 0, 0) : v3(0, 1, 0);
  V3 b = vnormOrZero(vcross(a, ref)), c = vnormOrZero(vcross(a, b));
  M4 toX = M4::ident();
  toX.at(0, 0) = a.x; toX.at(0, 1) = a.y; toX.at(0, 2) = a.z;
  toX.at(1, 0) = b.x; toX.at(1, 1) = b.y; toX.at(1, 2) = b.z;
  toX.at(2, 0) = c.x; toX.at(2, 1) = c.y; toX.at(2, 2) = c.z;
  M4 aboutX = M4::ident();
  aboutX.at(1, 1) = std::cos(rad); aboutX.at(1, 2) = -std::sin(rad); aboutX.at(2, 1) = std::sin(rad); aboutX.at(2, 2) = std::cos(rad);
  mulTop(mmul(mtranspose(toX), mmul(aboutX, toX)));
}
int HostScene::xformOf(const M4& m) {
  if (lastXform_ >= 0 && std::memcmp(lastXformM_.a, m.a, sizeof(m.a)) == 0) return lastXform_;     // every polygon of a mesh asks for the same CTM
  const int r = xformOfSlow(m); lastXform_ = r; lastXformM_ = m; return r;
}
int HostScene::xformOfSlow(const M4& m) {
  std::string key((const char*)m.a, sizeof(m.a));
  auto it = xformCache_.find(key); if (it != xformCache_.end()) return it->second;
  FXform x; M4 inv = minverse(m), adj = mtranspose(inv);
  std::memcpy(x.m, m.a, sizeof(x.m)); std::memcpy(x.inv, inv.a, sizeof(x.inv)); std::memcpy(x.adj, adj.a, sizeof(x.adj));
  int idx = (int)xforms.size(); xforms.push_back(x); xformCache_[key] = idx; return idx;
}
static M4 xm(const FXform& x) { M4 r; std::memcpy(r.a, x.m, sizeof(r.a)); return r; }
static M4 xinv(const FXform& x) { M4 r; std::memcpy(r.a, x.inv, sizeof(r.a)); return r; }

// ---- material state
void HostScene::setSurface(V3 d, V3 a, V3 s, double ph, double kr) {         // myScene.java:817-828
  txtrType_ = 0; mat_.diff = clampColor(d); mat_.amb = clampColor(a); mat_.spec = clampColor(s); mat_.phong = ph;
  mat_.krefl = kr; mat_.kreflClr = clampColor(v3(kr, kr, kr)); mat_.ktrans = 0; mat_.rfrIdx = 0; mat_.perm = v3(0, 0, 0);
}
void HostScene::resetTxtrDefaults() {                                         // :579-586
  txtrType_ = 0; numOctaves_ = 4; numPtsDist_ = 2; distFunc_ = 1; roiFunc_ = 1; rndColors_ = false; useCustClrs_ = false; useFwdTrans_ = false;
  noiseScale_ = 1.0; turbMult_ = 1.0; colorScale_ = 5.0; colorMult_ = .1; avgNumPerCell_ = 1.0; mortarThresh_ = 0.05; pdMult_ = v3(1, 1, 1);
  noiseColors_ = {namedColor("clr_nearblack"), namedColor("clr_white")};
}
void HostScene::setNoiseColor(const Tokens& k) {                              // :604-639
  if (!useCustClrs_) { noiseColors_.clear(); useCustClrs_ = true; }
  try { V3 c = (k.str(1) == "named") ? namedColor(k.str(2)) : k.rgb(1); noiseColors_.push_back(clampColor(c)); } catch (Missing&) {}
}
void HostScene::setTexture(const Tokens& k) {                                 // :642-777
  resetTxtrDefaults();
  const std::string& typ = k.str(0);
  bool worley = (typ == "stone"), useDefaults;
  try {
    if (!worley) {
      noiseScale_ = k.num(1); numOctaves_ = k.integer(2); turbMult_ = k.num(3); pdMult_ = v3(k.num(4), k.num(5), k.num(6));
      V3 py = v3(k.num(7), k.num(8), k.num(9));
      if (((py.x * py.x) + (py.y * py.y) + (py.z * py.z)) > 0) {
        py = vscale(py, kTwoPiF - 1.0); py = vadd(py, v3(1.0, 1.0, 1.0)); pdMult_ = v3(pdMult_.x * py.x, pdMult_.y * py.y, pdMult_.z * py.z);
      }
      useFwdTrans_ = (k.num(10) == 1.0);
      try { colorScale_ = k.num(11); colorMult_ = k.num(12); rndColors_ = true; } catch (Missing&) { rndColors_ = false; colorScale_ = 25.0; colorMult_ = .1; }
    } else {
      noiseScale_ = k.num(1); distFunc_ = k.integer(2); roiFunc_ = k.integer(3); numPtsDist_ = k.integer(4); avgNumPerCell_ = k.num(5); mortarThresh_ = k.num(6);
      useFwdTrans_ = (k.num(7) == 1.0);
      try { colorScale_ = k.num(8); colorMult_ = k.num(9); rndColors_ = true; } catch (Missing&) { rndColors_ = false; colorScale_ = 25.0; colorMult_ = .1; }
    }
    useDefaults = false;
  } catch (Missing&) { useDefaults = true; }
  auto dflt = [&](int oct, double ns, double tm, double cs, double cm, V3 pd) {
    numOctaves_ = oct; numPtsDist_ = 2; distFunc_ = 1; roiFunc_ = 1; rndColors_ = true; useFwdTrans_ = false;
    noiseScale_ = ns; turbMult_ = tm; colorScale_ = cs; colorMult_ = cm; avgNumPerCell_ = 1.0; mortarThresh_ = 0.05; pdMult_ = pd;
  };
  if (typ == "wood") { txtrType_ = 3; if (!useCustClrs_) noiseColors_ = {namedColor("clr_dkwood1"), namedColor("clr_ltwood1")}; if (useDefaults) dflt(4, 2.0, .4, 25.0, .2, v3(kTwoPiF * 2.7, 3.6, 4.3)); }
  else if (typ == "wood2") { txtrType_ = 6; if (!useCustClrs_) noiseColors_ = {namedColor("clr_dkwood2"), namedColor("clr_ltwood2")}; if (useDefaults) dflt(8, 1.0, .4, 25.0, .3, v3(kTwoPiF * 3.5, 7.9, 6.2)); }
  else if (typ == "marble") { txtrType_ = 4; if (!useCustClrs_) noiseColors_ = {namedColor("clr_nearblack"), namedColor("clr_offwhite")}; if (useDefaults) dflt(16, 1.0, 15.0, 24.0, .1, v3(kTwoPiF * 0.1, kTwoPiF * 31.4, kTwoPiF * 4.1)); }
  else if (typ == "stone") {
    txtrType_ = 5;
    if (!useCustClrs_) noiseColors_ = {namedColor("clr_mortar1"), namedColor("clr_mortar2"), namedColor("clr_brick1_1"), namedColor("clr_brick1_2"), namedColor("clr_brick2_1"), namedColor("clr_brick2_2"), namedColor("clr_brick3_1"), namedColor("clr_brick3_2"), namedColor("clr_brick4_1"), namedColor("clr_brick4_2")};
    if (useDefaults) dflt(8, 4.0, 1.0, 12.0, .2, v3(10.0, 10.0, 10.0));
  }
}

// getCurShader (myScene.java:524-542) + setCurrColors (myObjShader.java:51-75) + texture constructors
int HostScene::currentShader() {
  FShader s; std::memset(&s, 0, sizeof(s));
  FTexture t; std::memset(&t, 0, sizeof(t));
  auto put = [](double* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
  put(s.diff, mat_.diff); put(s.amb, mat_.amb); put(s.spec, mat_.spec); put(s.perm, mat_.perm); put(s.kreflClr, mat_.kreflClr);
  s.phong = mat_.phong; s.KRefl = mat_.krefl; s.KTrans = mat_.ktrans; s.currPerm = mat_.rfrIdx; s.diffConst = 1 - s.currPerm;
  s.avgDiff = (1.0 / 3.0) * (mat_.diff.x + mat_.diff.y + mat_.diff.z);
  if (s.avgDiff != 0) put(s.phtnDiffScl, v3(mat_.diff.x / s.avgDiff, mat_.diff.y / s.avgDiff, mat_.diff.z / s.avgDiff));
  double avgPerm = (1.0 / 3.0) * (mat_.perm.x + mat_.perm.y + mat_.perm.z);
  if (avgPerm != 0) put(s.phtnPermClr, v3(mat_.perm.x / avgPerm, mat_.perm.y / avgPerm, mat_.perm.z / avgPerm));
  if ((s.KRefl > 0.0) || (s.currPerm > 0.0) || (s.KTrans > 0.0)) s.flags |= SF_HAS_CAUSTIC;
  if (usePhotonMap_) s.flags |= SF_USE_PHOTON;
  if (isCausticPhtn_) s.flags |= SF_IS_CAUSTIC_PHTN;
  if (simpleRefr_) s.flags |= SF_SIMPLE;
  t.kind = txtrType_; t.imgTop = -1;
  if (txtrType_ == 1) { if (txtrdTop_) t.imgTop = curTopImage_; else t.kind = TK_IMAGE; }
  std::vector<double> cols;
  if (txtrType_ >= 2) {
    t.numOctaves = numOctaves_; t.numPtsDist = numPtsDist_; t.distFunc = distFunc_; t.roiFunc = roiFunc_; t.rndColors = rndColors_; t.useFwdTrans = useFwdTrans_;
    t.scale = noiseScale_; t.turbMult = turbMult_; t.colorScale = colorScale_; t.colorMult = colorMult_; t.avgNumPerCell = avgNumPerCell_; t.mortarThresh = mortarThresh_;
    t.periodMult[0] = pdMult_.x; t.periodMult[1] = pdMult_.y; t.periodMult[2] = pdMult_.z; t.periodMag = vmag(pdMult_);
    double last = 1.0 / std::pow(M_E, avgNumPerCell_), cum = last;                 // myTextureHandler.java:426-432
    for (int i = 1; i < 15; ++i) { last *= (avgNumPerCell_ / (1.0 * i)); cum += last; t.pdf[i - 1] = cum; }
    for (V3 c : noiseColors_) { cols.push_back(c.x); cols.push_back(c.y); cols.push_back(c.z); }
    t.colorCount = (int)noiseColors_.size();
  }
  // dedupe on the full snapshot
  std::string key((const char*)&s, sizeof(s)); key.append((const char*)&t, sizeof(t)); key.append((const char*)cols.data(), cols.size() * sizeof(double));
  auto& cache = shaderCache_;
  if (lastShader_ >= 0 && key == lastShaderKey_) { shaderOfSerial.push_back(lastShader_); return lastShader_; }     // consecutive polygons of one mesh
  int idx; auto it = cache.find(key);
  if (it != cache.end()) idx = it->second;
  else {
    t.colorStart = (int)(texColors.size() / 3); texColors.insert(texColors.end(), cols.begin(), cols.end());
    s.tex = (int)textures.size(); textures.push_back(t);
    s.serial = (int)shaders.size(); idx = (int)shaders.size(); shaders.push_back(s); cache[key] = idx;
  }
  shaderOfSerial.push_back(idx); lastShader_ = idx; lastShaderKey_ = key;
  return idx;
}

int HostScene::loadImage(const std::string& name) {
  auto it = imageIdx_.find(name); if (it != imageIdx_.end()) return it->second;
  HostImage im; bool ok = loader_ ? loader_(name, im) : false;
  if (!ok) {   // decoded-texture cache written by tools/decode_textures.py
    std::string p = texDir_ + "/" + name + ".argb"; FILE* f = std::fopen(p.c_str(), "rb");
    if (!f) throw std::runtime_error("texture not found: " + name + " (no loader result, no " + p + ")");
    char mg[4]; int32_t wh[2];
    if (std::fread(mg, 1, 4, f) != 4 || std::memcmp(mg, "ARGB", 4) != 0 || std::fread(wh, 4, 2, f) != 2) { std::fclose(f); throw std::runtime_error("bad texture file " + p); }
    im.w = wh[0]; im.h = wh[1]; im.px.resize((size_t)im.w * im.h);
    size_t got = std::fread(im.px.data(), 4, im.px.size(), f); std::fclose(f);
    if (got != im.px.size()) throw std::runtime_error("short texture file " + p);
  }
  FImage fi; fi.w = im.w; fi.h = im.h; fi.offset = (int64_t)texels.size();
  texels.insert(texels.end(), im.px.begin(), im.px.end());
  int idx = (int)images.size(); images.push_back(fi); imageIdx_[name] = idx; return idx;
}

// ---- objects
void HostScene::addGeom(const HGeom& gm, bool cmpIsLight) {                    // addObjectToScene, myScene.java:558-565
  if (toTmp_) { tmpList_.push_back(gm); return; }
  if (!cmpIsLight) topGeoms_.push_back(gm);
  allIsLight_.push_back(cmpIsLight);
}
HostScene::HGeom HostScene::makePrim(int type, int flags, const std::vector<double>& data, V3 origin, V3 bmin, V3 bmax) {
  FPrim p; std::memset(&p, 0, sizeof(p));
  p.type = type; p.flags = flags; p.xform = xformOf(ctm()); p.shader = currentShader(); p.data = (int)pdata.size(); p.serial = primSerial_++;
  pdata.insert(pdata.end(), data.begin(), data.end());
  HGeom gm; gm.kind = OK_PRIM; gm.idx = (int)prims.size(); gm.xform = p.xform;
  // postProcBBox: new myBBox(min,max) clips against its own +-100000 sentinels (myGeomBase.java:102-104)
  gm.bmin = v3(jmin(bmin.x, 100000), jmin(bmin.y, 100000), jmin(bmin.z, 100000)); gm.bmax = v3(jmax(bmax.x, -100000), jmax(bmax.y, -100000), jmax(bmax.z, -100000));
  V3 k = mpoint(ctm(), origin); gm.key[0] = k.x; gm.key[1] = k.y; gm.key[2] = k.z;
  prims.push_back(p); return gm;
}
void HostScene::addPrimitive(const Tokens& k) {                                // readPrimData, myScene.java:447-521
  const std::string& c = k.str(0);
  if (c == "box") {
    double x0 = jmin(k.num(1), k.num(4)), x1 = jmax(k.num(1), k.num(4)), y0 = jmin(k.num(2), k.num(5)), y1 = jmax(k.num(2), k.num(5)), z0 = jmin(k.num(3), k.num(6)), z1 = jmax(k.num(3), k.num(6));
    addGeom(makePrim(PT_BOX, 0, {x0, y0, z0, x1, y1, z1}, v3((x0 + x1) * .5, (y0 + y1) * .5, (z0 + z1) * .5), v3(x0, y0, z0), v3(x1, y1, z1)), false);
  } else if (c == "sphere" || c == "sphereIn" || c == "ellipsoid" || c == "moving_sphere") {
    double rx, ry, rz, x, y, z; std::vector<double> d;
    if (c == "ellipsoid") { rx = k.num(1); ry = k.num(2); rz = k.num(3); x = k.num(4); y = k.num(5); z = k.num(6); }
    else { rx = ry = rz = k.num(1); x = k.num(2); y = k.num(3); z = k.num(4); }
    d = {x, y, z, rx, ry, rz};
    int type = PT_SPHERE;
    if (c == "moving_sphere") { type = PT_MOVSPHERE; d.push_back(k.num(5)); d.push_back(k.num(6)); d.push_back(k.num(7)); }
    double ext = rx + ry + rz;                                                 // L1 "radius" box, myImpObject.java:127-139
    addGeom(makePrim(type, c == "sphereIn" ? PF_INVERTED : 0, d, v3(x, y, z), v3(x + -ext, y + -ext, z + -ext), v3(x + ext, y + ext, z + ext)), false);
  } else if (c == "cyl" || c == "cylinder" || c == "hollow_cylinder") {
    double rad, h, x, y, z, ox = 0, oy = 1, oz = 0;
    if (c == "cyl") { rad = k.num(1); h = k.num(2); x = k.num(3); y = k.num(4); z = k.num(5); try { ox = k.num(6); oy = k.num(7); oz = k.num(8); } catch (Missing&) {} }
    else { rad = k.num(1); x = k.num(2); z = k.num(3); double y0 = k.num(4), y1 = k.num(5); h = y1 - y0; y = y0; }
    double yTop = y + h, yBot = y, ext = rad + rad;
    std::vector<double> d = {x, y, z, rad, rad, yTop, yBot};
    int type = PT_HCYL;
    if (c != "hollow_cylinder") { type = PT_CYL; double caps[8] = {ox, oy, oz, -yTop, ox, -oy, oz, yBot}; d.insert(d.end(), caps, caps + 8); }
    addGeom(makePrim(type, 0, d, v3(x, y, z), v3(x + -ext, y + 0, z + -ext), v3(x + ext, y + h, z + ext)), false);
  } else if (c == "torus") {
    // torus R r [facets] x y z : centre (x,y,z), axis = object-space y, major radius R (myTorus.bodyRad), tube radius r (ringRad); the optional third
    // number of data/c2torus.cli (the facet count of the reference's abandoned polygonal torus) is accepted and ignored
    const bool six = k.size() >= 7; const double R = k.num(1), r = k.num(2), x = k.num(six ? 4 : 3), y = k.num(six ? 5 : 4), z = k.num(six ? 6 : 5), ext = R + r;
    addGeom(makePrim(PT_TORUS, 0, {x, y, z, R, r}, v3(x, y, z), v3(x + -ext, y + -r, z + -ext), v3(x + ext, y + r, z + ext)), false);
  } else if (c == "quadric") {
    // quadric a b c d e f g h i j [xmin ymin zmin xmax ymax zmax] : a x^2 + b y^2 + c z^2 + d xy + e xz + f yz + g x + h y + i z + j = 0 (the ten
    // coefficients of src/tmpQuadricEQ.txt), kept inside the optional clip box
    std::vector<double> d; for (int i = 1; i <= 10; ++i) d.push_back(k.num(i));
    double bx[6] = {-100000, -100000, -100000, 100000, 100000, 100000};
    if (k.size() >= 17) for (int i = 0; i < 6; ++i) bx[i] = k.num(11 + i);
    for (int i = 0; i < 3; ++i) if (bx[i] > bx[3 + i]) std::swap(bx[i], bx[3 + i]);
    d.insert(d.end(), bx, bx + 6);
    addGeom(makePrim(PT_QUADRIC, 0, d, v3((bx[0] + bx[3]) * .5, (bx[1] + bx[4]) * .5, (bx[2] + bx[5]) * .5), v3(bx[0], bx[1], bx[2]), v3(bx[3], bx[4], bx[5])), false);
  } else if (c == "plane") {
    // myPlane.setPlaneVals (myPlanarObject.java:236-270) + the two states invertNormal() would produce (:71-88)
    double a = k.num(1), b = k.num(2), cc = k.num(3), dd = k.num(4);
    V3 N0 = v3(a, b, cc); double mag = vmag(N0); N0 = vnorm(N0); double D0 = dd / mag;
    V3 rot = v3(N0.y, N0.z, N0.x); if ((N0.x == N0.y) && (N0.x == N0.z)) rot = vadd(rot, v3(1, 0, 0)); rot = vnorm(rot);
    int idx = 7; double sum = N0.x + N0.y + N0.z;
    if (sum == 0) { sum = N0.x + N0.y; idx = 6; if (sum == 0) { sum = N0.x + N0.z; idx = 5; if (sum == 0) { sum = N0.y + N0.z; idx = 3; } } }
    V3 P = v3(((idx & 4) == 4 ? -D0 / sum : 0), ((idx & 2) == 2 ? -D0 / sum : 0), ((idx & 1) == 1 ? -D0 / sum : 0));
    V3 U = vcross(N0, rot), V = vcross(N0, U);
    V3 vA[4] = {P, vadd(P, U), vadd(vadd(P, U), V), vadd(P, V)};
    auto polyND = [](const V3 v[4], V3& N, double& D) { V3 e0 = vsub(v[1], v[0]), e1 = vsub(v[2], v[1]); N = vnorm(vcross(e1, e0)); D = -((N.x * v[0].x) + (N.y * v[0].y) + (N.z * v[0].z)); };
    V3 vB[4] = {vA[3], vA[2], vA[1], vA[0]}, NB, NA; double DB, DA; polyND(vB, NB, DB); polyND(vA, NA, DA);
    std::vector<double> d = {N0.x, N0.y, N0.z, D0, NB.x, NB.y, NB.z, DB, NA.x, NA.y, NA.z, DA};
    addGeom(makePrim(PT_PLANE, 0, d, P, v3(100000, 100000, 100000), v3(-100000, -100000, -100000)), false);
  }
}
void HostScene::endPoly() {                                                    // myPlanarObject.java:44-100 for both winding orders
  if (!poly_.active) return;
  int n = poly_.n; std::vector<double> d; d.reserve(48);
  auto state = [&](const double v[4][3]) {
    V3 e0 = v3(v[1][0] - v[0][0], v[1][1] - v[0][1], v[1][2] - v[0][2]), e1 = v3(v[2][0] - v[1][0], v[2][1] - v[1][1], v[2][2] - v[1][2]);
    V3 N = vnorm(vcross(e1, e0)); double D = -((N.x * v[0][0]) + (N.y * v[0][1]) + (N.z * v[0][2]));
    for (int i = 0; i < n; ++i) { d.push_back(v[i][0]); d.push_back(v[i][1]); d.push_back(v[i][2]); }
    d.push_back(N.x); d.push_back(N.y); d.push_back(N.z); d.push_back(D);
  };
  double rv[4][3], ruv[4][2];
  for (int i = 0; i < n; ++i) { for (int c = 0; c < 3; ++c) rv[n - 1 - i][c] = poly_.v[i][c]; ruv[n - 1 - i][0] = poly_.uv[i][0]; ruv[n - 1 - i][1] = poly_.uv[i][1]; }
  state(poly_.v); state(rv);
  for (int i = 0; i < n; ++i) { d.push_back(poly_.uv[i][0]); d.push_back(poly_.uv[i][1]); }
  for (int i = 0; i < n; ++i) { d.push_back(ruv[i][0]); d.push_back(ruv[i][1]); }
  double sx = 0, sy = 0, sz = 0; V3 mn = v3(kDMax, kDMax, kDMax), mx = v3(-kDMax, -kDMax, -kDMax);
  for (int i = 0; i < n; ++i) {
    sx += poly_.v[i][0]; sy += poly_.v[i][1]; sz += poly_.v[i][2];
    if (poly_.v[i][0] < mn.x) mn.x = poly_.v[i][0]; if (poly_.v[i][1] < mn.y) mn.y = poly_.v[i][1]; if (poly_.v[i][2] < mn.z) mn.z = poly_.v[i][2];
    if (poly_.v[i][0] > mx.x) mx.x = poly_.v[i][0]; if (poly_.v[i][1] > mx.y) mx.y = poly_.v[i][1]; if (poly_.v[i][2] > mx.z) mx.z = poly_.v[i][2];
  }
  // the polygon keeps the CTM that was current at `begin` (myGeomBase ctor), the shader current at `end`
  M4 saved = stack_[top_]; stack_[top_] = poly_.m;
  HGeom gm = makePrim(n == 4 ? PT_QUAD : PT_TRI, 0, d, v3(sx / n, sy / n, sz / n), mn, mx);
  stack_[top_] = saved;
  addGeom(gm, false); poly_.active = false;
}

// myGeomList: children in order, box grown from the children's transformed min/max corners only
int HostScene::buildList(const std::vector<HGeom>& objs, int listXform, const M4& listM, V3& bmin, V3& bmax) {
  FList L; std::memset(&L, 0, sizeof(L)); L.xform = listXform; L.childStart = (int)children.size(); L.childCount = (int)objs.size();
  V3 mn = v3(100000, 100000, 100000), mx = v3(-100000, -100000, -100000);
  M4 linv = xinv(xforms[listXform]);
  auto grow = [&](V3 p) { mn.x = (mn.x < p.x) ? mn.x : p.x; mn.y = (mn.y < p.y) ? mn.y : p.y; mn.z = (mn.z < p.z) ? mn.z : p.z; mx.x = (mx.x > p.x) ? mx.x : p.x; mx.y = (mx.y > p.y) ? mx.y : p.y; mx.z = (mx.z > p.z) ? mx.z : p.z; };
  for (const HGeom& o : objs) {
    if (o.kind == OK_LIST || o.kind == OK_BVH) throw std::runtime_error("nested acceleration structures are not reachable through the .cli grammar");
    M4 cm = xm(xforms[o.xform]);
    FObjRef r; r.kind = o.kind; r.idx = o.idx; r.xform = o.xform; r.hitXform = xformOf(mmul(listM, cm));
    children.push_back(r);
    M4 rel = mmul(linv, cm); grow(mpoint(rel, o.bmin)); grow(mpoint(rel, o.bmax));
  }
  L.bmin[0] = mn.x; L.bmin[1] = mn.y; L.bmin[2] = mn.z; L.bmax[0] = mx.x; L.bmax[1] = mx.y; L.bmax[2] = mx.z;
  bmin = mn; bmax = mx; lists.push_back(L); return (int)lists.size() - 1;
}

// Object order of the reference's median-split tree, from the centroid keys alone (myGeomBase.java:338-386, DistRayTracer.java:409-418).
// A node receives its objects in the order its parent's split left them (the root: insertion order).  The reference keeps three lists per
// node, list i = that order stable-sorted by centroid coordinate i (TreeMap<Double, List> groups equal keys in arrival order,
// buildSortedObjAras); only two of them are ever read: the list of the split axis (first axis of strictly largest span, last - first of the
// sorted list = max - min) which is cut at (int)(.5 * count), and list 0, whose order a leaf keeps.  So: an inner node stable-sorts its range
// by the split-axis key and hands the two halves down; a leaf stable-sorts its range by x.  `count` is the reference's objListSize, which at
// the root is one less than the m objects present: the last object of the root's split-axis order is dropped (it stays behind
// at ord[count]); a root that is a leaf keeps all m.  keys = [3][n].
static void refOrderRange(const double* keys, int n, int32_t* ord, int s, int m, int count) {
  auto byAxis = [&](int ax) { const double* k = keys + (size_t)ax * n; std::stable_sort(ord + s, ord + s + m, [k](int32_t a, int32_t b) { return javaDoubleCompare(k[a], k[b]) < 0; }); };
  if (count <= 5) { byAxis(0); return; }                // DistRayTracer.maxPrimsPerLeaf
  double widest = -1; int axis = -1;
  for (int i = 0; i < 3; ++i) {
    const double* k = keys + (size_t)i * n; double lo = k[ord[s]], hi = lo;
    for (int j = 1; j < m; ++j) { const double v = k[ord[s + j]]; if (javaDoubleCompare(v, lo) < 0) lo = v; if (javaDoubleCompare(v, hi) > 0) hi = v; }
    const double span = hi - lo; if (widest < span) { widest = span; axis = i; }
  }
  if (axis < 0) throw std::runtime_error("BVH split axis undefined (NaN centroids)");
  byAxis(axis);
  const int split = (int)(.5 * count);
  refOrderRange(keys, n, ord, s, split, split);
  refOrderRange(keys, n, ord, s + split, count - split, count - split);
}
void HostScene::refOrderHost(int n, const double* keys, int32_t* ord) { for (int i = 0; i < n; ++i) ord[i] = i; if (n > 0) refOrderRange(keys, n, ord, 0, n, n - 1); }

// myBVH.addObjList (myGeomBase.java:360-386) over the order computed above: the tree's SHAPE depends on the counts only.  Returns a child
// reference (>= 0 inner node, < 0 ~list) and the node's box.
int32_t HostScene::buildBvhNode(const std::vector<HGeom>& objs, const int32_t* ord, int s, int m, int count, int bvhXform, const M4& bvhM, V3& bmin, V3& bmax) {
  if (count <= 5) {
    std::vector<HGeom> leaf; leaf.reserve(m); for (int i = 0; i < m; ++i) leaf.push_back(objs[ord[s + i]]);
    return ~buildList(leaf, bvhXform, bvhM, bmin, bmax);      // leaf keeps every object of the x-sorted list, in that order
  }
  const int split = (int)(.5 * count);
  int32_t me = (int32_t)nodes.size(); nodes.push_back(FNode()); std::memset(&nodes[me], 0, sizeof(FNode));
  V3 lmn, lmx, rmn, rmx;
  int32_t l = buildBvhNode(objs, ord, s, split, split, bvhXform, bvhM, lmn, lmx);
  int32_t r = buildBvhNode(objs, ord, s + split, count - split, count - split, bvhXform, bvhM, rmn, rmx);
  FNode& nd = nodes[me];
  nd.left = l; nd.right = r;
  nd.lmin[0] = lmn.x; nd.lmin[1] = lmn.y; nd.lmin[2] = lmn.z; nd.lmax[0] = lmx.x; nd.lmax[1] = lmx.y; nd.lmax[2] = lmx.z;
  nd.rmin[0] = rmn.x; nd.rmin[1] = rmn.y; nd.rmin[2] = rmn.z; nd.rmax[0] = rmx.x; nd.rmax[1] = rmx.y; nd.rmax[2] = rmx.z;
  // node box = sentinel box grown by both child boxes (min and max corners)
  bmin = v3(100000, 100000, 100000); bmax = v3(-100000, -100000, -100000);
  auto grow = [&](V3 p) { bmin.x = (bmin.x < p.x) ? bmin.x : p.x; bmin.y = (bmin.y < p.y) ? bmin.y : p.y; bmin.z = (bmin.z < p.z) ? bmin.z : p.z; bmax.x = (bmax.x > p.x) ? bmax.x : p.x; bmax.y = (bmax.y > p.y) ? bmax.y : p.y; bmax.z = (bmax.z > p.z) ? bmax.z : p.z; };
  grow(lmn); grow(lmx); grow(rmn); grow(rmx);
  return me;
}
void HostScene::endList(int type) {                                           // endTmpObjList, myScene.java:305-324
  toTmp_ = false;
  HGeom gm; gm.xform = xformOf(ctm()); M4 M = ctm();
  V3 k = mpoint(M, v3(0, 0, 0)); gm.key[0] = k.x; gm.key[1] = k.y; gm.key[2] = k.z;
  if (type == 0) { gm.kind = OK_LIST; gm.idx = buildList(tmpList_, gm.xform, M, gm.bmin, gm.bmax); }
  else {
    FBvh B; std::memset(&B, 0, sizeof(B)); B.xform = gm.xform; B.dropped = -1;
    int first = (int)nodes.size();
    if (tmpList_.empty()) { V3 a, b; B.root = ~buildList(tmpList_, gm.xform, M, a, b); gm.bmin = a; gm.bmax = b; }
    else {
      // the order is a function of the centroid keys alone: computed by the device (segmented radix sorts, csrc/refbvh.cuh) when the context has
      // one and the list is large, else by the host recursion -- both give the same array, bit for bit (tests/test_gpu_refbvh.py)
      const int n = (int)tmpList_.size(); std::vector<double> keys((size_t)3 * n); std::vector<int32_t> ord(n);
      for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) keys[(size_t)k * n + i] = tmpList_[i].key[k];
      bool done = false; const double t0 = nowMs();
      if (orderer_ && n >= ordererMin_) { done = orderer_(n, keys.data(), ord.data()); if (done) ++bvhDeviceBuilds; }
      if (!done) refOrderHost(n, keys.data(), ord.data());
      const double t1 = nowMs(); msBvhOrder += t1 - t0; bvhObjects += n;
      B.root = buildBvhNode(tmpList_, ord.data(), 0, n, n - 1, gm.xform, M, gm.bmin, gm.bmax);
      msBvhShape += nowMs() - t1;
    }
    B.nodeCount = (int)nodes.size() - first;
    B.bmin[0] = gm.bmin.x; B.bmin[1] = gm.bmin.y; B.bmin[2] = gm.bmin.z; B.bmax[0] = gm.bmax.x; B.bmax[1] = gm.bmax.y; B.bmax[2] = gm.bmax.z;
    gm.kind = OK_BVH; gm.idx = (int)bvhs.size(); bvhs.push_back(B);
  }
  addGeom(gm, false);
}
void HostScene::addInstance(const std::string& name, bool useShader) {         // myScene.java:404-410, mySceneObject.java:98-110
  auto it = named_.find(name); if (it == named_.end()) throw std::runtime_error("instance of unknown named object '" + name + "'");
  const HGeom& base = it->second;
  if (base.kind == OK_INSTANCE) throw std::runtime_error("instance of an instance is not supported");
  if (base.isLight) throw std::runtime_error("instanced lights are not supported");
  M4 baseM = xm(xforms[base.xform]);
  FInstance I; std::memset(&I, 0, sizeof(I)); I.baseKind = base.kind; I.baseIdx = base.idx; I.xform = xformOf(mmul(baseM, ctm()));
  I.shader = useShader ? currentShader() : -1; I.serial = instSerial_++;
  HGeom gm; gm.kind = OK_INSTANCE; gm.idx = (int)instances.size(); gm.xform = I.xform;
  V3 mn = mpoint(baseM, base.bmin), mx = mpoint(baseM, base.bmax);
  gm.bmin = v3(jmin(mn.x, 100000), jmin(mn.y, 100000), jmin(mn.z, 100000)); gm.bmax = v3(jmax(mx.x, -100000), jmax(mx.y, -100000), jmax(mx.z, -100000));
  V3 k = mpoint(ctm(), v3(0, 0, 0)); gm.key[0] = k.x; gm.key[1] = k.y; gm.key[2] = k.z;    // trans_origin is taken before the CTM is replaced (myGeomBase.java:37-39)
  instances.push_back(I); addGeom(gm, false);
}
// ---- Sierpinski layout (myScene.java:328-392); dim / shifts are float arithmetic widened to double
void HostScene::sierpShader(int level, int maxLevel) {
  float b = 1.0f - std::min(1.0f, (1.5f * level / maxLevel)), r = 1.0f - b, tmp = std::min((1.2f * (level - (maxLevel / 2))) / (1.0f * maxLevel), 1.0f), gg = (tmp * tmp);
  txtrdTop_ = false; txtrdBtm_ = false;
  setSurface(v3(std::min(1.0f, r + .5f), std::min(1.0f, gg + .5f), std::min(1.0f, b + .5f)), v3(0, 0, 0), v3(0, 0, 0), 0, 0);
}
void HostScene::sierpShift(float t) { rotate(120, 1, 0, 0); translate(0, t, 0); rotate(-120, 1, 0, 0); }
void HostScene::sierpSub(float dim, float sc, const std::string& name, int level, int maxLevel, bool shdr) {
  if (level >= maxLevel) return;
  float newDim = sc * dim;
  push(); translate(0, .1f * dim, 0); rotate(70, 0, 1, 0); if (shdr) sierpShader(level, maxLevel); addInstance(name, shdr); pop();
  static const float sqrt66 = std::sqrt(6.0f) / 6.0f;
  float shift = sqrt66 * dim;
  push(); translate(0, shift, 0); scale(sc, sc, sc); sierpSub(newDim, sc, name, level + 1, maxLevel, shdr); pop();
  push(); sierpShift(shift); scale(sc, sc, sc); sierpSub(newDim, sc, name, level + 1, maxLevel, shdr); pop();
  push(); rotate(120, 0, 1, 0); sierpShift(shift); rotate(-120, 0, 1, 0); scale(sc, sc, sc); sierpSub(newDim, sc, name, level + 1, maxLevel, shdr); pop();
  push(); rotate(-120, 0, 1, 0); sierpShift(shift); rotate(120, 0, 1, 0); scale(sc, sc, sc); sierpSub(newDim, sc, name, level + 1, maxLevel, shdr); pop();
}
void HostScene::sierpinski(const std::string& name, float sc, int depth, bool shdr) { tmpList_.clear(); toTmp_ = true; sierpSub(8, sc, name, 0, depth, shdr); endList(1); }

void HostScene::addLight(int type, const Tokens& k) {                          // myScene.java:413-444, myLight.java:20-30,150-157,244-247
  FLight L; std::memset(&L, 0, sizeof(L)); L.type = type; L.xform = xformOf(ctm());
  auto put = [](double* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
  V3 pos = v3(k.num(1), k.num(2), k.num(3)), orient = v3(0, 0, 0), col;
  if (type == LT_POINT) col = k.rgb(4);
  else if (type == LT_SPOT) {
    orient = vnorm(v3(k.num(4), k.num(5), k.num(6))); double inT = k.num(7), outT = k.num(8); col = k.rgb(9);
    L.innerRad = inT * kDegToRadF; L.outerRad = outT * kDegToRadF; L.radDiff = L.outerRad - L.innerRad; put(L.tangent, orthoVec(orient));
  } else { L.radius = k.num(4); orient = vnorm(v3(k.num(5), k.num(6), k.num(7))); col = k.rgb(8); put(L.tangent, orthoVec(orient)); }
  put(L.color, col); put(L.origin, pos); put(L.orient, orient);
  HGeom gm; gm.isLight = true; gm.lightIdx = (int)lights.size(); lights.push_back(L);
  if (toTmp_) throw std::runtime_error("lights inside begin_list are not supported");
  allIsLight_.push_back(true);
}

// ---------------------------------------------------------------------------------------------
// interpreter (myRTFileReader.java:45-347)
// ---------------------------------------------------------------------------------------------
void HostScene::loadFile(const std::string& file, const std::string& dataDir) { const double t0 = nowMs(); dataDir_ = dataDir; if (texDir_.empty()) texDir_ = dataDir + "/txtrs_argb"; readFile(file, true); msParse += nowMs() - t0; }
void HostScene::readFile(const std::string& file, bool isMain) {
  std::ifstream in(dataDir_ + "/" + file);
  if (!in) { if (isMain) throw std::runtime_error("cannot read scene file " + dataDir_ + "/" + file); warnings.push_back("File Read Error : " + file); return; }
  bool savedMain = isMain_; int savedSpp = curSpp_; std::string savedVert = vertType_;
  isMain_ = isMain; curSpp_ = g.spp; vertType_ = "triangle";     // locals of readRTFile
  std::string line;
  while (std::getline(in, line)) { while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back(); command(line); }
  isMain_ = savedMain; curSpp_ = savedSpp; vertType_ = savedVert;
}
// next space-separated token of [p, end) as a number, with Tokens::num's rules; false = absent or not a number (the caller then takes the general path,
// which reports the line exactly as before)
static inline bool fastNum(const char*& p, const char* end, double& v) {
  while (p < end && *p == ' ') ++p;
  if (p == end) return false;
  const char* q = p; while (q < end && *q != ' ') ++q;
  char* e = nullptr; v = std::strtod(p, &e);
  if (e == p || e > q) return false;
  if (e != q && !((*e == 'f' || *e == 'F' || *e == 'd' || *e == 'D') && e + 1 == q)) return false;
  p = q; return true;
}
void HostScene::command(const std::string& line) {
  {   // meshes are three `vertex` lines out of five: parsed in place, without building the token list (1.3 -> 0.5 us per line on a 262 k-triangle soup)
    const char* p = line.c_str(); const char* const end = p + line.size();
    while (p < end && *p == ' ') ++p;
    if (end - p > 7 && std::memcmp(p, "vertex ", 7) == 0) {
      const char* q = p + 6; double x, y, z;
      if (fastNum(q, end, x) && fastNum(q, end, y) && fastNum(q, end, z)) {
        if (poly_.active && poly_.cnt < poly_.n) { poly_.v[poly_.cnt][0] = x; poly_.v[poly_.cnt][1] = y; poly_.v[poly_.cnt][2] = z; }
        poly_.cnt++; return;
      }
    }
  }
  Tokens k(line);
  if (k.size() == 0 || k.t[0][0] == '#') return;
  const std::string& c = k.t[0];
  try {
    // meshes are >99 % `vertex` / `begin` / `end` lines: dispatch them before the long chain of command names (same handlers as below)
    if (c == "vertex") { if (poly_.active && poly_.cnt < poly_.n) { poly_.v[poly_.cnt][0] = k.num(1); poly_.v[poly_.cnt][1] = k.num(2); poly_.v[poly_.cnt][2] = k.num(3); } poly_.cnt++; return; }
    if (c == "end") { endPoly(); vertType_ = "triangle"; return; }
    if (c == "begin") {
      if (k.size() > 1) vertType_ = k.t[1];            // (a bare `begin` keeps the current type; no exception on the per-triangle path)
      poly_ = Poly(); poly_.active = true; poly_.n = (vertType_ == "quad") ? 4 : 3; poly_.m = ctm();
      for (int i = 0; i < 4; ++i) { poly_.v[i][0] = poly_.v[i][1] = poly_.v[i][2] = 0; poly_.uv[i][0] = poly_.uv[i][1] = 0; }
      return;
    }
    if (c == "fov" || c == "fishEye" || c == "fisheye" || c == "ortho" || c == "orthographic") {
      if (!isMain_) { warnings.push_back("scene type in child file ignored"); return; }
      g.spp = (curSpp_ != 0) ? curSpp_ : 1;
      if (c == "fov") {                                                        // myScene.java:1367-1381
        double fov = k.num(1), rad = M_PI * fov / 180.0; if (std::fabs(fov - 180) < .001) { fov -= .001; rad -= .0001; }
        g.camKind = CAM_FOV; g.viewZ = -1 * (std::max(g.rows, g.cols) / 2.0) / std::tan(rad / 2);
      } else if (c == "ortho" || c == "orthographic") { double w = k.num(1), h = k.num(2), div = std::min(g.cols, g.rows); g.camKind = CAM_ORTHO; g.orthPerRow = h / div; g.orthPerCol = w / div; }
      else { double rad = M_PI * k.num(1) / 180.0; g.camKind = CAM_FISHEYE; g.aperatureHlf = rad / 2.0; }
    }
    else if (c == "lens") { g.lensRadius = k.num(1); g.focalD = k.num(2); g.hasDof = 1; }
    else if (c == "write") { saveName = k.str(1); sawWrite = true; }
    else if (c == "read") readFile(k.str(1), false);
    else if (c == "reset_timer" || c == "print_timer") {}
    else if (c == "refine") refine = lowered(k.str(1)) == "on";                 // setRefine, myScene.java:796-803 (the previews are made by drt_refine_pass)
    else if (c == "rays_per_pixel") { int r = k.integer(1); curSpp_ = r; g.spp = r; }
    else if (c == "antialias") { int r = k.integer(1) * k.integer(2); curSpp_ = r; g.spp = r; }
    else if (c == "background") {
      if (k.str(1) == "texture") {
        g.skyImage = loadImage(k.str(2)); g.hasSky = 1; double rad = k.num(3);
        g.skyOrigin[0] = k.num(4); g.skyOrigin[1] = k.num(5); g.skyOrigin[2] = k.num(6); g.skyRad[0] = g.skyRad[1] = g.skyRad[2] = rad;
      } else { V3 b = k.rgb(1); g.bg[0] = b.x; g.bg[1] = b.y; g.bg[2] = b.z; txtrType_ = 0; }
    }
    else if (c == "point_light") addLight(LT_POINT, k);
    else if (c == "spotlight") addLight(LT_SPOT, k);
    else if (c == "disk_light") addLight(LT_DISK, k);
    else if (c == "caustic_photons" || c == "diffuse_photons") {               // myScene.java:919-931
      usePhotonMap_ = true; isCausticPhtn_ = c.find("caustic") != std::string::npos;
      g.photonKind = isCausticPhtn_ ? 1 : 2; g.numPhotonsCast = k.integer(1); g.kNhood = k.integer(2);
      double md = (double)(float)k.num(3); g.phMaxDist2 = md * md;
    }
    else if (c == "final_gather") (void)k.integer(1);
    else if (c == "diffuse") { V3 d = k.rgb(1), a = k.rgb(4); txtrdTop_ = txtrdBtm_ = false; setSurface(d, a, v3(0, 0, 0), 0, 0); }
    else if (c == "reflective") { V3 d = k.rgb(1), a = k.rgb(4); txtrdTop_ = txtrdBtm_ = false; double kr = k.num(7); setSurface(d, a, v3(0, 0, 0), 0, kr); }
    else if (c == "shiny" || c == "surface") {                                 // setSurfaceShiny, myRTFileReader.java:358-378
      V3 d = k.rgb(1), a = k.rgb(4), s = k.rgb(7); double ph = k.num(10), kr = k.num(11), kt = 0, ri = 0;
      txtrdTop_ = txtrdBtm_ = false; setSurface(d, a, s, ph, kr);
      try {
        kt = k.num(12); setSurface(d, a, s, ph, kr); mat_.ktrans = kt;
        ri = k.num(13); mat_.rfrIdx = ri; mat_.perm = clampColor(v3(ri, ri, ri));
        mat_.perm = clampColor(v3(k.num(14), k.num(15), k.num(16)));
      } catch (Missing&) {}
      if (c == "shiny" && ((kt > 0) || (ri > 0))) simpleRefr_ = true;
    }
    else if (c == "perm") { double v = k.num(1); mat_.rfrIdx = v; mat_.perm = clampColor(v3(v, v, v)); try { V3 p = v3(k.num(2), k.num(3), k.num(4)); mat_.perm = clampColor(p); } catch (Missing&) {} }
    else if (c == "phong") mat_.phong = k.num(1);
    else if (c == "krefl") { double v = k.num(1); mat_.krefl = v; mat_.kreflClr = clampColor(v3(v, v, v)); }
    else if (c == "depth") (void)k.num(1);
    else if (c == "ktrans") mat_.ktrans = k.num(1);
    else if (c == "begin_list") { tmpList_.clear(); toTmp_ = true; }
    else if (c == "end_list") endList(0);
    else if (c == "end_accel") endList(1);
    else if (c == "sierpinski") {
      std::string name = k.str(1); float sc = .5f; int depth = 5; bool shdr = false;
      try { depth = k.integer(2); sc = (float)k.num(3); (void)k.str(4); shdr = true; } catch (Missing&) {}
      sierpinski(name, sc, depth, shdr);
    }
    else if (c == "named_object") {                                            // myScene.java:394-402
      std::string name = k.str(1);
      if (allIsLight_.empty()) throw std::runtime_error("named_object with no object to name");
      bool wasLight = allIsLight_.back(); allIsLight_.pop_back();
      if (wasLight) throw std::runtime_error("named lights are not supported");
      named_[name] = topGeoms_.back(); topGeoms_.pop_back();
    }
    else if (c == "instance") addInstance(k.str(1), k.size() > 2);
    else if (c == "image_texture" || c == "texture") {                         // myRTFileReader.java:257-273
      std::string side = lowered(k.str(1));
      if (side == "top" || side != "bottom") { curTopImage_ = loadImage(side == "top" ? k.str(2) : k.str(1)); txtrdTop_ = true; }
      else { (void)loadImage(k.str(1)); txtrdBtm_ = true; }
      txtrType_ = 1;
    }
    else if (c == "noise") { double sc = k.num(1); resetTxtrDefaults(); txtrType_ = 2; noiseScale_ = sc; }
    else if (c == "noise_color") setNoiseColor(k);
    else if (c == "marble" || c == "stone" || c == "wood" || c == "wood2") setTexture(k);
    else if (c == "texture_coord") { if (poly_.active && poly_.cnt < poly_.n) { poly_.uv[poly_.cnt][0] = k.num(1); poly_.uv[poly_.cnt][1] = k.num(2); } }
    else if (c == "vertex") { if (poly_.active && poly_.cnt < poly_.n) { poly_.v[poly_.cnt][0] = k.num(1); poly_.v[poly_.cnt][1] = k.num(2); poly_.v[poly_.cnt][2] = k.num(3); } poly_.cnt++; }
    else if (c == "end") { endPoly(); vertType_ = "triangle"; }
    else if (c == "box" || c == "plane" || c == "cyl" || c == "cylinder" || c == "hollow_cylinder" || c == "sphere" || c == "moving_sphere" || c == "sphereIn" || c == "ellipsoid") addPrimitive(k);
    // ---- extension primitives (north star: "quadrics, tori").  The reference has no reader command for either and its myTorus never reports a hit
    // (myImpObject.java:330-390; notes in src/tmpQuadricEQ.txt), so a drop-in must ignore these lines -- which is what happens unless the scene
    // itself says `extensions on` (a line the reference ignores as well).  Semantics are this project's own: parity unpinned.
    else if (c == "extensions") { extensions_ = (k.size() < 2 || lowered(k.str(1)) != "off"); }
    else if ((c == "torus" || c == "quadric") && extensions_) addPrimitive(k);
    else if (c == "push") push();
    else if (c == "pop") pop();
    else if (c == "rotate") rotate(k.num(1), k.num(2), k.num(3), k.num(4));
    else if (c == "scale") scale(k.num(1), k.num(2), k.num(3));
    else if (c == "translate") translate(k.num(1), k.num(2), k.num(3));
    else warnings.push_back("unknown command '" + c + "'");
  } catch (Missing&) { throw std::runtime_error("malformed .cli line: " + line); }
}

// A leaf qualifies when it holds 1..7 plain triangles that all use (triXform, triHitXform) and whose reversed winding state carries
// exactly the negated normal. Returns the tri-leaf code or -1.
int32_t HostScene::packLeaf(const FList& L, int triXform, int triHitXform) {
  if (L.childCount < 1 || L.childCount > 7) return -1;
  const size_t start = tris.size();
  for (int i = 0; i < L.childCount; ++i) {
    const FObjRef& c = children[L.childStart + i];
    if (c.kind != OK_PRIM || c.xform != triXform || c.hitXform != triHitXform || prims[c.idx].type != PT_TRI) { tris.resize(start); return -1; }
    const double* q = pdata.data() + prims[c.idx].data; const double* r = q + DRT_TRI_STATE;
    FTri T; std::memset(&T, 0, sizeof(T));
    for (int k = 0; k < 9; ++k) T.v[k] = q[k];
    for (int k = 0; k < 3; ++k) T.N[k] = q[9 + k];
    T.D = q[12]; T.Drev = r[12]; T.prim = c.idx; T.pad[0] = (int32_t)tris.size();   // pad[0] = rank in the reference's visiting order
    bool ok = true;
    for (int k = 0; k < 3; ++k) { double neg = -q[9 + k]; if (std::memcmp(&neg, &r[9 + k], 8) != 0 && !(neg == 0 && r[9 + k] == 0)) ok = false; }
    for (int k = 0; k < 9; ++k) if (r[k] != q[3 * (2 - k / 3) + k % 3]) ok = false;
    if (!ok) { tris.resize(start); return -1; }
    tris.push_back(T);
  }
  return (int32_t)((start << 3) | (size_t)L.childCount);
}
void HostScene::buildFastBvh(FBvh& B) {
  B.fast = 0; B.triXform = B.triHitXform = -1; B.fastRoot = B.root; B.triStart = (int)tris.size(); B.triCount = 0;
  if (B.root < 0) return;                                   // a single leaf: nothing to accelerate
  // first leaf decides the shared CTMs
  int32_t n = B.root; while (n >= 0) n = nodes[n].left;
  const FList& L0 = lists[~n]; if (L0.childCount < 1) return;
  const FObjRef& c0 = children[L0.childStart]; const int tx = c0.xform, thx = c0.hitXform;
  const size_t triMark = tris.size();
  // DFS, left first: FTri index order == the reference's visiting order (used as the tie-break of the fast traversal)
  // (simple recursive walk: depth is O(log n) for the median-split tree)
  struct Walk { HostScene* h; int tx, thx; bool ok = true;
    void go(int32_t r) { if (!ok) return; FNode& nd = h->nodes[r]; nd.triL = nd.triR = -1;
      if (nd.left >= 0) go(nd.left); else { nd.triL = h->packLeaf(h->lists[~nd.left], tx, thx); if (nd.triL < 0) ok = false; }
      if (!ok) return;
      if (nd.right >= 0) go(nd.right); else { nd.triR = h->packLeaf(h->lists[~nd.right], tx, thx); if (nd.triR < 0) ok = false; } } } w{this, tx, thx};
  w.go(B.root);
  if (!w.ok) { tris.resize(triMark); struct Clr { HostScene* h; void go(int32_t r) { FNode& nd = h->nodes[r]; nd.triL = nd.triR = -1; if (nd.left >= 0) go(nd.left); if (nd.right >= 0) go(nd.right); } } c{this}; c.go(B.root); return; }
  B.fast = 1; B.triXform = tx; B.triHitXform = thx; B.triStart = (int)triMark; B.triCount = (int)(tris.size() - triMark);
}

void HostScene::finalize() {
  const double tF0 = nowMs();
  tris.clear();
  for (FBvh& B : bvhs) buildFastBvh(B);
  top.clear();
  for (const HGeom& gm : topGeoms_) { FObjRef r; r.kind = gm.kind; r.idx = gm.idx; r.xform = gm.xform; r.hitXform = gm.xform; top.push_back(r); }
  // plain triangles of the top-level list get a packed record too (FPrim::pad0 = FTri index, -1 otherwise): the fast modes test them with the
  // same lean routine as BVH leaves instead of the generic primitive switch
  for (FPrim& P : prims) P.pad0 = -1;
  for (const FObjRef& r : top) if (r.kind == OK_PRIM && prims[r.idx].type == PT_TRI && prims[r.idx].pad0 < 0) {
    FList one; std::memset(&one, 0, sizeof(one)); one.childStart = (int)children.size(); one.childCount = 1;
    FObjRef c = r; children.push_back(c);                       // temporary child entry so that packLeaf() can be reused verbatim
    const int32_t code = packLeaf(one, r.xform, r.hitXform);
    children.pop_back();
    if (code >= 0) prims[r.idx].pad0 = code >> 3;
  }
  g.numTop = (int)top.size(); g.numLights = (int)lights.size();
  // FP32 mirror of every node (round to nearest) and, per fast BVH, the largest |coordinate| of its node boxes per axis
  nodes32.resize(nodes.size());
  for (size_t i = 0; i < nodes.size(); ++i) { const FNode& n = nodes[i]; FNode32& m = nodes32[i];
    for (int k = 0; k < 3; ++k) { m.lmin[k] = (float)n.lmin[k]; m.lmax[k] = (float)n.lmax[k]; m.rmin[k] = (float)n.rmin[k]; m.rmax[k] = (float)n.rmax[k]; }
    m.left = n.left; m.right = n.right; m.triL = n.triL; m.triR = n.triR; }
  for (FBvh& B : bvhs) { B.absMax[0] = B.absMax[1] = B.absMax[2] = 0; B.pad2 = 0; if (!B.fast || B.root < 0) continue;
    std::vector<int32_t> st{B.root};
    while (!st.empty()) { const FNode& n = nodes[st.back()]; st.pop_back();
      for (int k = 0; k < 3; ++k) { const double m = std::max(std::max(std::fabs(n.lmin[k]), std::fabs(n.lmax[k])), std::max(std::fabs(n.rmin[k]), std::fabs(n.rmax[k])));
        const float f = (float)m; B.absMax[k] = std::max(B.absMax[k], f >= m ? f : std::nextafter(f, INFINITY)); }
      if (n.left >= 0) st.push_back(n.left); if (n.right >= 0) st.push_back(n.right); } }
  msFinalize = nowMs() - tF0;
}

void HostScene::dumpNode(int32_t ref, std::vector<int32_t>& out) const {
  if (ref < 0) { const FList& L = lists[~ref]; out.push_back(-2); out.push_back(L.childCount);
    for (int i = 0; i < L.childCount; ++i) { const FObjRef& c = children[L.childStart + i]; out.push_back(c.kind == OK_INSTANCE ? (0x40000000 | instances[c.idx].serial) : prims[c.idx].serial); } }
  else { out.push_back(-1); dumpNode(nodes[ref].left, out); dumpNode(nodes[ref].right, out); }
}
void HostScene::dumpBvh(int topIdx, std::vector<int32_t>& out, double box[6]) const {
  out.clear(); if (topIdx < 0 || topIdx >= (int)topGeoms_.size() || topGeoms_[topIdx].kind != OK_BVH) return;
  const FBvh& B = bvhs[topGeoms_[topIdx].idx]; dumpNode(B.root, out);
  for (int i = 0; i < 3; ++i) { box[i] = B.bmin[i]; box[3 + i] = B.bmax[i]; }
}

}  // namespace drt
