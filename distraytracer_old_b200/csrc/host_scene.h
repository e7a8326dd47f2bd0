// Host side of the drop-in boundary: the `.cli` scene interpreter and the flattener that turns the
// reference's scene-building calls into the POD arrays of scene_flat.h.
//
// Mirrors (reference file:line, /root/reference/src/rayTracerDistAccelShdPhtnMap/):
//   myRTFileReader.java:15-378   command interpreter, one call of command() per line
//   myScene.java:300-324         begin_list / end_list / end_accel
//   myScene.java:328-410         sierpinski, named_object, instance
//   myScene.java:413-542         lights, primitives, shader+texture snapshot
//   myScene.java:571-777         procedural texture parameters
//   myScene.java:805-857         lens / material setters
//   myScene.java:1235-1323       matrix stack
//   myGeomBase.java:338-386      median-split BVH (kept topology-exact: SURVEY Q2,Q3,Q5)
#pragma once
#include "host_math.h"
#include "scene_flat.h"
#include <functional>
#include <map>
#include <string>
#include <vector>

namespace drt {

struct HostImage { int w = 0, h = 0; std::vector<int32_t> px; };
// returns false when the image cannot be provided
typedef std::function<bool(const std::string& name, HostImage& out)> ImageLoader;

// fills ord[n] with the object order of the reference's median-split tree (HostScene::refOrderHost); false = not done, the host does it
typedef std::function<bool(int n, const double* keys, int32_t* ord)> BvhOrderer;

class HostScene {
 public:
  HostScene(int cols, int rows);
  // ---- interpreter
  void loadFile(const std::string& file, const std::string& dataDir);      // readRTFile(file, null)
  void command(const std::string& line);                                     // one line of a .cli file
  void setImageLoader(ImageLoader l) { loader_ = l; }
  void setTexDir(const std::string& d) { texDir_ = d; }
  void overrideSpp(int spp) { if (spp > 0) g.spp = spp; }
  void overridePhotons(long long n) { if (n >= 0) g.numPhotonsCast = (int)n; }
  void setSeed(uint64_t s) { g.seed = s; }
  // object order of the reference's median-split tree from the [3][n] centroid keys (see host_scene.cpp); the device version is installed by
  // the context when it owns a GPU and is used for lists of at least minObjects objects
  static void refOrderHost(int n, const double* keys, int32_t* ord);
  void setBvhOrderer(BvhOrderer o, int minObjects) { orderer_ = o; ordererMin_ = minObjects; }
  int bvhDeviceBuilds = 0;              // BVHs whose order came from the device
  double msBvhOrder = 0, msBvhOrderDevice = 0, msBvhShape = 0, msFinalize = 0, msParse = 0;   // host wall clock of the build phases (drt_build_info); msBvhOrderDevice = CUDA-event time inside the device orderer
  long long bvhObjects = 0;             // objects handed to median-split builds
  // ---- flat result (valid after parsing; finalize() fills FGlobals counts)
  void finalize();
  FGlobals g;
  std::vector<FXform> xforms;
  std::vector<FPrim> prims;
  std::vector<double> pdata;
  std::vector<FObjRef> top;           // myScene.objList
  std::vector<FObjRef> children;      // list children, all lists concatenated
  std::vector<FInstance> instances;
  std::vector<FList> lists;
  std::vector<FBvh> bvhs;
  std::vector<FNode> nodes;
  std::vector<FNode32> nodes32;       // FP32 mirror of `nodes` (built by finalize)
  std::vector<FTri> tris;             // packed triangles of the fast BVHs (see scene_flat.h)
  std::vector<FLight> lights;
  std::vector<FShader> shaders;
  std::vector<FTexture> textures;
  std::vector<double> texColors;      // rgb triples
  std::vector<FImage> images;
  std::vector<int32_t> texels;
  std::vector<int32_t> shaderOfSerial;   // reference creates one shader object per getCurShader(); map creation order -> deduped index
  std::vector<std::string> warnings;
  std::string saveName;
  bool sawWrite = false;
  bool refine = false;                // `refine on` (myScene.setRefine): the host shows the progressive previews of drt_refine_pass
  // ---- introspection used by tests (BVH order KAT)
  void dumpBvh(int topIdx, std::vector<int32_t>& out, double box[6]) const;

 private:
  // build-time view of one myGeomBase
  struct HGeom { int32_t kind = OK_PRIM, idx = -1, xform = -1; V3 bmin, bmax; double key[3]; bool isLight = false; int lightIdx = -1; };
  struct Tokens;
  // matrix stack
  M4 stack_[10]; int top_ = 0;
  const M4& ctm() const { return stack_[top_]; }
  void push(); void pop(); void mulTop(const M4& m);
  void translate(double x, double y, double z); void scale(double x, double y, double z); void rotate(double deg, double ax, double ay, double az);
  int xformOf(const M4& m);
  int xformOfSlow(const M4& m);
  int lastXform_ = -1; M4 lastXformM_;
  std::map<std::string, int> xformCache_;
  // material / texture state (myScene.java:117-145)
  struct Mat { V3 diff, amb, spec, perm, kreflClr; double phong = 0, krefl = 0, ktrans = 0, rfrIdx = 0; } mat_;
  int txtrType_ = 0; double noiseScale_ = 1; std::vector<V3> noiseColors_; int numOctaves_ = 8; double turbMult_ = 1, colorScale_ = 10, colorMult_ = .2; V3 pdMult_;
  bool rndColors_ = false, useCustClrs_ = false, useFwdTrans_ = false; double avgNumPerCell_ = 1, mortarThresh_ = .04; int numPtsDist_ = 2, distFunc_ = 1, roiFunc_ = 1;
  bool extensions_ = false;           // `extensions on`: torus / quadric commands are honoured (the reference ignores both lines)
  bool simpleRefr_ = false, txtrdTop_ = false, txtrdBtm_ = false, usePhotonMap_ = false, isCausticPhtn_ = false;
  int curTopImage_ = -1;
  void setSurface(V3 d, V3 a, V3 s, double ph, double kr);
  void resetTxtrDefaults();
  void setTexture(const Tokens& k); void setNoiseColor(const Tokens& k);
  int currentShader();                 // getCurShader(): snapshot, deduped against the previous snapshot
  std::map<std::string, int> shaderCache_; int lastShader_ = -1; std::string lastShaderKey_;
  // objects
  std::vector<HGeom> topGeoms_, tmpList_; std::vector<bool> allIsLight_; bool toTmp_ = false;
  std::map<std::string, HGeom> named_;
  void addGeom(const HGeom& gm, bool cmpIsLight);
  void endList(int type);
  int buildList(const std::vector<HGeom>& objs, int listXform, const M4& listM, V3& bmin, V3& bmax);
  int32_t buildBvhNode(const std::vector<HGeom>& objs, const int32_t* ord, int s, int m, int count, int bvhXform, const M4& bvhM, V3& bmin, V3& bmax);
  BvhOrderer orderer_; int ordererMin_ = 1 << 30;
  void addPrimitive(const Tokens& k);
  HGeom makePrim(int type, int flags, const std::vector<double>& data, V3 origin, V3 bmin, V3 bmax);
  void addInstance(const std::string& name, bool useShader);
  void sierpinski(const std::string& name, float sc, int depth, bool shdr);
  void sierpSub(float dim, float sc, const std::string& name, int level, int maxLevel, bool shdr);
  void sierpShader(int level, int maxLevel);
  void sierpShift(float t);
  void addLight(int type, const Tokens& k);
  int loadImage(const std::string& name);
  std::map<std::string, int> imageIdx_;
  // polygon being read
  struct Poly { bool active = false; int n = 3; double v[4][3]; double uv[4][2]; int cnt = 0; int xform = -1; M4 m; } poly_;
  std::string vertType_ = "triangle";
  void endPoly();
  // reader state
  int curSpp_ = 0; bool isMain_ = true; std::string dataDir_, texDir_; ImageLoader loader_;
  void readFile(const std::string& file, bool isMain);
  int primSerial_ = 0, instSerial_ = 0;
  void dumpNode(int32_t ref, std::vector<int32_t>& out) const;
  void buildFastBvh(FBvh& B);          // packed triangle records + tri-leaf codes when the BVH qualifies
  int32_t packLeaf(const FList& L, int triXform, int triHitXform);
};

}  // namespace drt
