// ORACLE (test infrastructure only -- never linked into the product path).
// CPU restatement of the reference's shaders, textures and photon kd-tree.
// PARITY UNPINNED (no reference tests / no JVM), see orc_math.hpp.
//
// Follows (relative to /root/reference/src/rayTracerDistAccelShdPhtnMap/):
//   myObjShader.java:7-493 (myObjShader), :496-656 (mySimpleReflObjShdr), :659-674 (myColor)
//   myTextureHandler.java:44-59,61-134,178-306,309-386,388-498,501-680
//   DistRayTracer.java:232-310 (Perlin noise, float), :21-29 (Worley constants), :467-530 (named colours)
//   myLight.java:278-446 (myPhoton, myKD_Tree)
#pragma once
#include "orc_core.hpp"

namespace orc {

struct Image { int width = 0, height = 0; std::vector<int32_t> pixels; };   // PImage.pixels, ARGB

// myColor clamps every channel to <= 1 at construction (myObjShader.java:661)
inline Vec3 mkColor(double r, double g, double b) { return Vec3(jmin(1, r), jmin(1, g), jmin(1, b)); }
inline Vec3 colorFromInt(int32_t c) { return mkColor(((c >> 16) & 0xFF) / 255.0, ((c >> 8) & 0xFF) / 255.0, (c & 0xFF) / 255.0); }
inline int32_t colorGetInt(const Vec3& c) { return (int32_t)(((uint32_t)255 << 24) + ((uint32_t)j2i(c.x * 255) << 16) + ((uint32_t)j2i(c.y * 255) << 8) + (uint32_t)j2i(c.z * 255)); }
inline Vec3 interpColor(const Vec3& A, double t, const Vec3& B) { return mkColor((A.x + t * (B.x - A.x)), (A.y + t * (B.y - A.y)), (A.z + t * (B.z - A.z))); }

// ---- Perlin noise, float arithmetic (DistRayTracer.java:234-310). Build with -ffp-contract=off.
struct Perlin {
  int perm[512];
  Perlin() {
    static const int p[256] = {151,160,137,91,90,15,131,13,201,95,96,53,194,233,7,225,140,36,103,30,69,142,8,99,37,240,21,10,23,190,6,148,247,120,234,75,0,26,197,62,94,252,219,203,117,35,11,32,57,177,33,88,237,149,56,87,174,20,125,136,171,168,68,175,74,165,71,134,139,48,27,166,77,146,158,231,83,111,229,122,60,211,133,230,220,105,92,41,55,46,245,40,244,102,143,54,65,25,63,161,1,216,80,73,209,76,132,187,208,89,18,169,200,196,135,130,116,188,159,86,164,100,109,198,173,186,3,64,52,217,226,250,124,123,5,202,38,147,118,126,255,82,85,212,207,206,59,227,47,16,58,17,182,189,28,42,223,183,170,213,119,248,152,2,44,154,163,70,221,153,101,155,167,43,172,9,129,22,39,253,19,98,108,110,79,113,224,232,178,185,112,104,218,246,97,228,251,34,242,193,238,210,144,12,191,179,162,241,81,51,145,235,249,14,239,107,49,192,214,31,181,199,106,157,184,84,204,176,115,121,50,45,127,4,150,254,138,236,205,93,222,114,67,29,24,72,243,141,128,195,78,66,215,61,156,180};
    for (int i = 0; i < 512; ++i) perm[i] = p[i & 255];
  }
  static float dotg(const int g[3], float x, float y, float z) { return g[0] * x + g[1] * y + g[2] * z; }
  static float mix(float a, float b, float t) { return (1 - t) * a + t * b; }
  static float fade(float t) { return t * t * t * (t * (t * 6 - 15) + 10); }
  float noise(float x, float y, float z) const {
    static const int grad3[12][3] = {{1,1,0},{-1,1,0},{1,-1,0},{-1,-1,0},{1,0,1},{-1,0,1},{1,0,-1},{-1,0,-1},{0,1,1},{0,-1,1},{0,1,-1},{0,-1,-1}};
    int X = fastfloorf(x), Y = fastfloorf(y), Z = fastfloorf(z);
    x = x - X; y = y - Y; z = z - Z;
    X &= 255; Y &= 255; Z &= 255;
    int gi000 = perm[X + perm[Y + perm[Z]]] % 12, gi001 = perm[X + perm[Y + perm[Z + 1]]] % 12,
        gi010 = perm[X + perm[Y + 1 + perm[Z]]] % 12, gi011 = perm[X + perm[Y + 1 + perm[Z + 1]]] % 12,
        gi100 = perm[X + 1 + perm[Y + perm[Z]]] % 12, gi101 = perm[X + 1 + perm[Y + perm[Z + 1]]] % 12,
        gi110 = perm[X + 1 + perm[Y + 1 + perm[Z]]] % 12, gi111 = perm[X + 1 + perm[Y + 1 + perm[Z + 1]]] % 12;
    float n000 = dotg(grad3[gi000], x, y, z), n100 = dotg(grad3[gi100], x - 1, y, z),
          n010 = dotg(grad3[gi010], x, y - 1, z), n110 = dotg(grad3[gi110], x - 1, y - 1, z),
          n001 = dotg(grad3[gi001], x, y, z - 1), n101 = dotg(grad3[gi101], x - 1, y, z - 1),
          n011 = dotg(grad3[gi011], x, y - 1, z - 1), n111 = dotg(grad3[gi111], x - 1, y - 1, z - 1);
    float u = fade(x), v = fade(y), w = fade(z);
    return mix(mix(mix(n000, n100, u), mix(n010, n110, u), v), mix(mix(n001, n101, u), mix(n011, n111, u), v), w);
  }
  double noise3(const Vec3& p) const { return noise((float)p.x, (float)p.y, (float)p.z); }
};
inline const Perlin& perlin() { static Perlin P; return P; }

// named colours (DistRayTracer.java:467-530); clr_rnd is not reproducible (Processing random) -> white
inline Vec3 getClr(std::string n) {
  for (auto& c : n) c = (char)tolower(c);
  struct E { const char* k; double r, g, b; };
  static const E tab[] = {{"clr_gray",0.47,0.47,0.47},{"clr_white",1,1,1},{"clr_yellow",1,1,0},{"clr_cyan",0,1,1},{"clr_magenta",1,0,1},{"clr_red",1,0,0},{"clr_blue",0,0,1},{"clr_purple",0.6,0.2,1},{"clr_green",0,1,0},
    {"clr_ltwood1",0.94,0.47,0.12},{"clr_ltwood2",0.94,0.8,0.4},{"clr_dkwood1",0.2,0.08,0.08},{"clr_dkwood2",0.3,0.20,0.16},{"clr_mortar1",0.2,0.2,0.2},{"clr_mortar2",0.7,0.7,0.7},
    {"clr_brick1_1",0.6,0.18,0.22},{"clr_brick1_2",0.8,0.26,0.33},{"clr_brick2_1",0.6,0.32,0.16},{"clr_brick2_2",0.8,0.45,0.25},{"clr_brick3_1",0.3,0.01,0.07},{"clr_brick3_2",0.6,0.02,0.13},{"clr_brick4_1",0.4,0.1,0.17},{"clr_brick4_2",0.6,0.3,0.13},
    {"clr_darkgray",0.31,0.31,0.31},{"clr_darkred",0.47,0,0},{"clr_darkblue",0,0,0.47},{"clr_darkpurple",0.4,0.2,0.6},{"clr_darkgreen",0,0.47,0},{"clr_darkyellow",0.47,0.47,0},{"clr_darkmagenta",0.47,0,0.47},{"clr_darkcyan",0,0.47,0.47},
    {"clr_lightgray",0.78,0.78,0.78},{"clr_lightred",1,.43,.43},{"clr_lightblue",0.43,0.43,1},{"clr_lightgreen",0.43,1,0.43},{"clr_lightyellow",1,1,.43},{"clr_lightmagenta",1,.43,1},{"clr_lightcyan",0.43,1,1},
    {"clr_black",0,0,0},{"clr_nearblack",0.05,0.05,0.05},{"clr_faintgray",0.43,0.43,0.43},{"clr_faintred",0.43,0,0},{"clr_faintblue",0,0,0.43},{"clr_faintgreen",0,0.43,0},{"clr_faintyellow",0.43,0.43,0},{"clr_faintcyan",0,0.43,0.43},{"clr_faintmagenta",0.43,0,0.43},{"clr_offwhite",0.95,0.98,0.92}};
  for (const E& e : tab) if (n == e.k) return mkColor(e.r, e.g, e.b);
  return mkColor(1, 1, 1);
}

// ---- textures (myTextureHandler.java)
struct Texture {
  Scene* scene; Shader* shdr; bool txtrdTop = false, txtrdBtm = false;
  Texture(Scene* s, Shader* sh) : scene(s), shdr(sh) {}
  virtual ~Texture() {}
  virtual void getDiffTxtrColor(const RayHit& hit, const Vec3& diffuse, double diffConst, double out[3]) = 0;
  virtual const Image* topImage() const { return nullptr; }
};
struct NonTexture : Texture {                                          // :44-59
  using Texture::Texture;
  void getDiffTxtrColor(const RayHit&, const Vec3& d, double k, double out[3]) override { out[0] = d.x * k; out[1] = d.y * k; out[2] = d.z * k; }
};
struct ImageTexture : Texture {                                        // :61-134
  const Image *top = nullptr, *bottom = nullptr;
  ImageTexture(Scene* s, Shader* sh);
  const Image* topImage() const override { return top; }
  void getTextureColor(const RayHit& hit, const Image* tex, double out[3]);
  void getDiffTxtrColor(const RayHit& hit, const Vec3& d, double k, double out[3]) override {
    if (txtrdTop) getTextureColor(hit, top, out); else { out[0] = d.x; out[1] = d.y; out[2] = d.z; }
    out[0] *= k; out[1] *= k; out[2] *= k;
    // (bottom texture is fetched and discarded by the reference, :112-116)
  }
};
struct NoiseTexture : Texture {                                        // :178-306
  double scale; std::vector<Vec3> colors; int numOctaves; double turbMult, colorScale, colorMult; bool rndColors, useFwdTrans; Vec3 periodMult;
  NoiseTexture(Scene* s, Shader* sh, double scl);
  double getNoiseVal(Vec3& hl) { hl.mult(scale); return perlin().noise3(hl); }
  double getTurbVal(Vec3& t) { t.mult(scale); double res = 0, f = 1.0, a = 1.0; for (int i = 0; i < numOctaves; ++i) { res += perlin().noise((float)(t.x * f), (float)(t.y * f), (float)(t.z * f)) * a; a *= .5; f *= 1.92; } return res; }
  double getAbsTurbVal(Vec3& t) { t.mult(scale); double res = 0, f = 1.0, a = 1.0; for (int i = 0; i < numOctaves; ++i) { res += std::fabs(perlin().noise((float)(t.x * f), (float)(t.y * f), (float)(t.z * f))) * a; a *= .5; f *= 1.92; } return res; }
  Vec3 getHitLoc(const RayHit& h) { return useFwdTrans ? h.fwdTransHitLoc : h.hitLoc; }
  double linPtVal(const Vec3& v) { return (v.x * periodMult.x + v.y * periodMult.y + v.z * periodMult.z); }
  double sqPtVal(const Vec3& v) { return std::sqrt((v.x * v.x) * periodMult.x + (v.y * v.y) * periodMult.y + (v.z * v.z) * periodMult.z); }
  void getClrAra(double distVal, const Vec3& rawPt, int i0, int i1, double out[3]) {    // :277-294
    Vec3 pt(rawPt); pt.mult(colorScale); double mult = colorMult;
    double rm[3] = {1.0, 1.0, 1.0};
    if (rndColors) {
      rm[0] = 1.0 + (mult * perlin().noise((float)pt.x, (float)pt.z, (float)pt.y));
      rm[1] = 1.0 + (mult * perlin().noise((float)pt.y, (float)pt.x, (float)pt.z));
      rm[2] = 1.0 + (mult * perlin().noise((float)pt.z, (float)pt.y, (float)pt.x));
    }
    int n = (int)colors.size(); if (i0 >= n) i0 = n - 1; if (i1 >= n) i1 = n - 1;   // Java would throw; clamp
    const Vec3 &c0 = colors[i0], &c1 = colors[i1];
    out[0] = jmax(0, jmin(1.0, (c0.x) + rm[0] * distVal * ((c1.x) - (c0.x))));
    out[1] = jmax(0, jmin(1.0, (c0.y) + rm[1] * distVal * ((c1.y) - (c0.y))));
    out[2] = jmax(0, jmin(1.0, (c0.z) + rm[2] * distVal * ((c1.z) - (c0.z))));
  }
  void applyDiffConst(double k, double out[3]) { if (std::fabs(k - 1.0) > EPS) { out[0] *= k; out[1] *= k; out[2] *= k; } }
  void getDiffTxtrColor(const RayHit& hit, const Vec3&, double k, double out[3]) override {   // :257-265
    Vec3 hl = getHitLoc(hit); double res = turbMult * getNoiseVal(hl); double val = .5 * res + .5;
    out[0] = out[1] = out[2] = val; applyDiffConst(k, out);
  }
};
struct BaseWoodTexture : NoiseTexture {                                // :309-334
  using NoiseTexture::NoiseTexture;
  void getDiffTxtrColor(const RayHit& hit, const Vec3&, double k, double out[3]) override {
    Vec3 hv = getHitLoc(hit); double res = getNoiseVal(hv);
    double sq = sqPtVal(hv) + turbMult * res;
    double dv = std::sin(sq * periodMult.mag()); dv *= 1.1; dv += .5; dv = (dv < 0 ? 0 : (dv > 1 ? 1 : dv));
    getClrAra(dv, hit.hitLoc, 0, 1, out); applyDiffConst(k, out);
  }
};
struct WoodTexture : NoiseTexture {                                    // :337-359
  using NoiseTexture::NoiseTexture;
  void getDiffTxtrColor(const RayHit& hit, const Vec3&, double k, double out[3]) override {
    Vec3 hv = getHitLoc(hit); double res = getTurbVal(hv);
    double sq = sqPtVal(hv) + turbMult * res;
    double dv = std::sin(sq * periodMult.mag()); dv = 1 - (dv < 0 ? 0 : dv);
    getClrAra(dv, hv, 0, 1, out); applyDiffConst(k, out);
  }
};
struct MarbleTexture : NoiseTexture {                                  // :362-386
  using NoiseTexture::NoiseTexture;
  void getDiffTxtrColor(const RayHit& hit, const Vec3&, double k, double out[3]) override {
    Vec3 hv = getHitLoc(hit); double res = getAbsTurbVal(hv);
    double spt = linPtVal(hv) / periodMult.mag() + turbMult * res;
    double dv = .5 * std::sin(spt) + .5;
    getClrAra(dv, hv, 0, 1, out); applyDiffConst(k, out);
  }
};
struct CellularTexture : NoiseTexture {                                // :388-498, ROI functors :515-680
  double avgNumPerCell, mortarThresh; int numPtsDist, roiFunc, distFunc;
  std::vector<std::pair<double, int>> pdfs;       // cumulative Poisson table (key, value)
  CellularTexture(Scene* s, Shader* sh, double scl);
  static int hashInts(int x, int y, int z) { return (int)((uint32_t)x * 1572869u + (uint32_t)y * 6291469u + (uint32_t)z); }
  int lookupNumPoints(double prob) const {        // pdfs.get(lowerKey(prob) or firstKey)
    int best = -1; for (size_t i = 0; i < pdfs.size(); ++i) if (pdfs[i].first < prob) best = (int)i;   // keys are ascending
    return pdfs[best < 0 ? 0 : best].second;
  }
  static double fixDist(double d) { if (d < 0) d *= -1; if (d > 1.0) d = 1.0 / d; return d; }
  double calcROI(const std::vector<double>& k) const {
    int i = 0, modVal = -1; double dist = 0; int n = numPtsDist;
    switch (roiFunc) {
      case 0: for (double d : k) { dist += d; i++; if (i >= n) break; } return fixDist(dist);
      case 2: for (double d : k) { dist += 1.0 / (modVal * d); i++; if (i >= n) break; modVal *= -1; } return fixDist(dist);
      case 3: for (double d : k) { dist += (modVal * std::pow(d, ++i)); if (i >= n) break; modVal *= -1; } return fixDist(dist);
      case 4: for (double d : k) { dist += (modVal * std::log(1 + d)); i++; if (i >= n) break; modVal *= -1; } return fixDist(dist);
      case 5: for (double d : k) { dist += std::pow(d, ++i); if (i >= n) break; } return fixDist(dist);
      case 6: for (double d : k) { dist += std::log(1 + d); i++; if (i >= n) break; } return dist;
      case 7: for (double d : k) { dist += std::pow(d, -(++i)); if (i >= n) break; } return fixDist(dist);
      case 8: for (double d : k) { dist += 1.0 / std::log(1 + d); i++; if (i >= n) break; } return fixDist(dist);
      case 1: default: for (double d : k) { dist += (modVal * d); i++; if (i >= n) break; modVal *= -1; } return fixDist(dist);
    }
  }
  void getDiffTxtrColor(const RayHit& hit, const Vec3&, double k, double out[3]) override {
    static const int nb[27][3] = {{0,0,0},{0,0,1},{0,0,-1},{0,1,0},{0,1,1},{0,1,-1},{0,-1,0},{0,-1,1},{0,-1,-1},{1,0,0},{1,0,1},{1,0,-1},{1,1,0},{1,1,1},{1,1,-1},{1,-1,0},{1,-1,1},{1,-1,-1},{-1,0,0},{-1,0,1},{-1,0,-1},{-1,1,0},{-1,1,1},{-1,1,-1},{-1,-1,0},{-1,-1,1},{-1,-1,-1}};
    Vec3 hv = getHitLoc(hit); hv.mult(scale);
    int hl[3] = {fastfloor(hv.x), fastfloor(hv.y), fastfloor(hv.z)};
    // ordered map keyed by distance with Double.compare ordering; equal keys overwrite the cell (last writer wins)
    struct Cmp { bool operator()(double a, double b) const { return dcompare(a, b) < 0; } };
    std::map<double, int, Cmp> distToPts;     // value = seed of owning cell
    JavaRandom gen;
    for (int i = 0; i < 27; ++i) {
      int cx = hl[0] + nb[i][0], cy = hl[1] + nb[i][1], cz = hl[2] + nb[i][2];
      int seed = hashInts(cx, cy, cz);
      gen.setSeed((int64_t)seed);
      double prob = gen.nextDouble();
      int numPoints = lookupNumPoints(prob);
      for (int j = 0; j < numPoints; ++j) {
        double px = cx + gen.nextDouble(), py = cy + gen.nextDouble(), pz = cz + gen.nextDouble();
        Vec3 pt(px, py, pz);
        double d = (distFunc == 0) ? hv.L1Dist(pt) : hv.dist(pt);
        distToPts[d] = seed;
      }
    }
    std::vector<double> keys; keys.reserve(distToPts.size());
    for (auto& e : distToPts) keys.push_back(e.first);
    double dist = calcROI(keys); dist = (dist < 0 ? 0 : dist > 1 ? 1 : dist);
    int brick = 2;
    if (dist < mortarThresh) brick = 0;
    else {
      gen.setSeed((int64_t)distToPts.begin()->second);
      double res = gen.nextDouble();
      brick = 2 * (1 + (fastfloor(((int)(colors.size() / 2) - 1) * res)));
    }
    getClrAra(.65, hv, brick, brick + 1, out); applyDiffConst(k, out);
  }
};

// ---- photon map (myLight.java:278-446)
struct Photon { double pwr[3]; double pos[4]; };
struct KDNode { Photon* photon; int split_axis; KDNode *left, *right; };
struct KDTree {
  std::vector<Photon*> photon_list; KDNode* root = nullptr;
  int num_Cast, maxNumNeighbors; double baseMaxDist2;
  KDTree(int numCast, int numNear, double maxDist) : num_Cast(numCast), maxNumNeighbors(numNear), baseMaxDist2(maxDist * maxDist) {}
  void add_photon(Photon* p) { photon_list.push_back(p); }
  void build_tree() { root = photon_list.empty() ? nullptr : build(photon_list.data(), (int)photon_list.size()); }
  KDNode* build(Photon** pl, int n) {                                   // :332-381 (stable sort == Collections.sort)
    KDNode* node = new KDNode();
    if (n == 1) { node->photon = pl[0]; node->split_axis = -1; node->left = node->right = nullptr; return node; }
    double mins[3] = {1e20, 1e20, 1e20}, maxs[3] = {-1e20, -1e20, -1e20};
    for (int i = 0; i < n; ++i) for (int j = 0; j < 3; ++j) { if (pl[i]->pos[j] < mins[j]) mins[j] = pl[i]->pos[j]; if (pl[i]->pos[j] > maxs[j]) maxs[j] = pl[i]->pos[j]; }
    double dx = maxs[0] - mins[0], dy = maxs[1] - mins[1], dz = maxs[2] - mins[2];
    int ax = -1;
    if (dx >= dy && dx >= dz) ax = 0; else if (dy >= dx && dy >= dz) ax = 1; else ax = 2;
    std::stable_sort(pl, pl + n, [ax](const Photon* a, const Photon* b) { return a->pos[ax] < b->pos[ax]; });
    int sp = n / 2;
    node->photon = pl[sp]; node->split_axis = ax;
    node->left = (sp == 0) ? nullptr : build(pl, sp);
    node->right = (sp == n - 1) ? nullptr : build(pl + sp + 1, n - sp - 1);
    return node;
  }
  // k nearest within radius, farthest first (:389-445). Own max-heap on squared distance.
  struct Near { double d2; const Photon* p; };
  void find_near(double x, double y, double z, std::vector<Near>& out) const {
    out.clear(); if (!root) return;
    double max_dist2 = baseMaxDist2; double pos[3] = {x, y, z};
    std::vector<Near> heap; heap.reserve(maxNumNeighbors + 1);
    visit(pos, root, heap, max_dist2);
    auto cmp = [](const Near& a, const Near& b) { return a.d2 < b.d2; };
    while (!heap.empty()) { std::pop_heap(heap.begin(), heap.end(), cmp); out.push_back(heap.back()); heap.pop_back(); }
  }
  void visit(const double pos[3], const KDNode* node, std::vector<Near>& heap, double& max_dist2) const {
    auto cmp = [](const Near& a, const Near& b) { return a.d2 < b.d2; };
    const Photon* ph = node->photon; int axis = node->split_axis;
    if (axis != -1) {
      double delta = pos[axis] - ph->pos[axis], delta2 = delta * delta;
      if (delta < 0) { if (node->left) visit(pos, node->left, heap, max_dist2); if (node->right && delta2 < max_dist2) visit(pos, node->right, heap, max_dist2); }
      else { if (node->right) visit(pos, node->right, heap, max_dist2); if (node->left && delta2 < max_dist2) visit(pos, node->left, heap, max_dist2); }
    }
    double dx = pos[0] - ph->pos[0], dy = pos[1] - ph->pos[1], dz = pos[2] - ph->pos[2];
    double len2 = dx * dx + dy * dy + dz * dz;
    if (len2 < max_dist2) {
      heap.push_back({len2, ph}); std::push_heap(heap.begin(), heap.end(), cmp);
      if ((int)heap.size() > maxNumNeighbors) { std::pop_heap(heap.begin(), heap.end(), cmp); heap.pop_back(); }
      if ((int)heap.size() == maxNumNeighbors) { if (heap.front().d2 < max_dist2) max_dist2 = heap.front().d2; }
    }
  }
};

// ---- shaders (myObjShader.java)
struct Shader {
  Scene* scene; Texture* txtr = nullptr; bool simple = false;
  double phongExp, KRefl, KTrans, currPerm, diffConst;
  Vec3 diffuseColor, ambientColor, specularColor, curPermClr, KReflClr;
  Vec3 phtnDiffScl, phtnSpecScl, phtnPermClr; double avgDiffClr, avgSpecClr, avgPermClr;
  bool hasCaustic, usePhotonMap, isCausticPhtn;
  int serial = -1;
  Shader(Scene* s, bool simple_);
  static double fresPerp(double n1, double n2, double cI, double cT) { double a = n1 * cI, b = n2 * cT, nd = (a - b) / (a + b); return nd * nd; }   // :78-81
  static double fresPlel(double n1, double n2, double cI, double cT) { double a = n1 * cT, b = n2 * cI, nd = (a - b) / (a + b); return nd * nd; }   // :83-86
  static Vec3 compReflDir(const Vec3& eye, const Vec3& n) { double dp = 2 * (eye.dot(n)); Vec3 t(n.x * dp, n.y * dp, n.z * dp); Vec3 r = vsub(t, eye); r.normalize(); return r; }  // :89-96
  void calcShadowColor(const RayHit& hit, const double tex[3], double out[3]);        // :98-153
  // Fresnel split shared by calcTransClr (:157-276), calcTransRay (:297-397) and calcSimpleTransClr (:503-631)
  struct Fres { Vec3 N, backToEye; double n, cosTheta1, cosTheta2, transReflRatio, oneM, refractNormMult; bool TIR; };
  Fres fresnel(const RayHit& hit, double matIdx, double rayIdx) const;
  Vec3 refractDir(const Fres& f) const { Vec3 u(f.backToEye); u.mult(f.n * -1); Vec3 nv(f.N); nv.mult((f.n * f.cosTheta1) - f.cosTheta2); u.add(nv); u.normalize(); return u; }
  void calcTransClr(const RayHit& hit, double out[3]);
  void calcSimpleTransClr(const RayHit& hit, double out[3]);
  void calcReflClr(const RayHit& hit, double out[3]);                                   // :278-294
  Vec3 getColorAtPos(const RayHit& hit);                                                // :409-438 / :635-651
  void getIrradianceFromPhtnTree(const RayHit& hit, double res[3]);                     // :441-458
  bool findCausticRayHit(RayHit& hit, Ray& out);                                        // :461-478 (+ calcTransRay / calcReflRay)
};

}  // namespace orc
