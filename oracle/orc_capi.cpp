// ORACLE (test infrastructure only -- never linked into the product path).
// C entry points over the CPU restatement, consumed via ctypes by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
// PARITY UNPINNED (no reference tests / no JVM), see orc_math.hpp.
#include "orc_scene.hpp"
#include <chrono>
#include <thread>
#include <atomic>

using namespace orc;

struct OrcHandle {
  Options opt; std::string file;
  std::vector<Scene*> copies;     // one scene per worker thread (reference objects mutate while rendering)
  std::string err;
  Scene* get(size_t i) {
    while (copies.size() <= i) {
      Scene* s = new Scene(opt); s->readRTFile(file, true);
      if (opt.sppOverride > 0) s->numRaysPerPixel = opt.sppOverride;
      if (!copies.empty() && copies[0]->isPhtnMapRndrd) { s->photonTree = copies[0]->photonTree; s->isPhtnMapRndrd = true; }   // seeded: identical anyway
      copies.push_back(s);
    }
    return copies[i];
  }
};

extern "C" {

struct orc_opts { int cols, rows, spp, literal_renorm; uint64_t seed; long long photons; const char* data_dir; const char* tex_dir; };

void* orc_load(const char* scene_file, const orc_opts* o, char* err, int errlen) {
  OrcHandle* h = new OrcHandle;
  try {
    h->opt.cols = o->cols; h->opt.rows = o->rows; h->opt.sppOverride = o->spp; h->opt.literalRenorm = o->literal_renorm != 0; h->opt.seed = o->seed;
    h->opt.photonOverride = o->photons; h->opt.dataDir = o->data_dir ? o->data_dir : "."; h->opt.texDir = o->tex_dir ? o->tex_dir : ".";
    h->file = scene_file; h->get(0);
    return h;
  } catch (std::exception& e) { if (err && errlen > 0) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; } delete h; return nullptr; }
}
void orc_free(void* hv) { delete (OrcHandle*)hv; }   // scenes are intentionally leaked with the process (test tool)

int orc_info(void* hv, int* out /*8*/) {
  Scene* s = ((OrcHandle*)hv)->get(0);
  out[0] = s->sceneCols; out[1] = s->sceneRows; out[2] = s->numRaysPerPixel; out[3] = (int)s->objList.size(); out[4] = (int)s->lightList.size();
  out[5] = s->primSerialCnt; out[6] = s->instSerialCnt; out[7] = s->usePhotonMap ? (s->isCausticPhtn ? 1 : 2) : 0;
  return 0;
}

// Emit the photon map (if the scene asks for one). Returns stored photon count.
long long orc_emit_photons(void* hv) {
  Scene* s = ((OrcHandle*)hv)->get(0); s->initRender();
  return s->photonTree ? (long long)s->photonTree->photon_list.size() : 0;
}
// copy photons (x,y,z,r,g,b) in emission order
long long orc_get_photons(void* hv, double* out, long long cap) {
  Scene* s = ((OrcHandle*)hv)->get(0); if (!s->photonTree) return 0;
  // photon_list is re-ordered by the kd build; order is not meaningful -> callers sort
  long long n = 0; for (Photon* p : s->photonTree->photon_list) { if (n >= cap) break; double* o = out + 6 * n; o[0] = p->pos[0]; o[1] = p->pos[1]; o[2] = p->pos[2]; o[3] = p->pwr[0]; o[4] = p->pwr[1]; o[5] = p->pwr[2]; ++n; }
  return n;
}

// myKD_Tree.find_near + getIrradianceFromPhtnTree at explicit world points: out4 = {sum r, sum g, sum b, d^2 of the farthest neighbour}
long long orc_photon_probe(void* hv, long long n, const double* pts, double* out4) {
  Scene* s = ((OrcHandle*)hv)->get(0); s->initRender(); if (!s->photonTree) return 0;
  std::vector<KDTree::Near> hood;
  for (long long i = 0; i < n; ++i) {
    s->photonTree->find_near(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], hood);
    double* o = out4 + 4 * i; o[0] = o[1] = o[2] = o[3] = 0;
    if (hood.empty()) continue;
    o[3] = hood[0].d2; for (auto& nb : hood) { o[0] += nb.p->pwr[0]; o[1] += nb.p->pwr[1]; o[2] += nb.p->pwr[2]; }
  }
  return n;
}

// Render the pixel rectangle [x0,x1) x [y0,y1). Output arrays are (y1-y0)*(x1-x0), row-major; any may be NULL.
// stats: 10 x uint64 {primary, shadow, reflect, refract, photonSeg, boxTests, primTests, boxTestsPrimary, primTestsPrimary, photonsStored}
// returns wall seconds of the pixel loop (photon emission excluded; it is reported by orc_emit_photons' caller)
double orc_render(void* hv, int x0, int y0, int x1, int y1, int threads, int32_t* argb, int32_t* hitPrim, int32_t* hitInst, double* rgb, double* tOut, uint64_t* stats) {
  OrcHandle* h = (OrcHandle*)hv;
  if (threads < 1) threads = 1;
  for (int i = 0; i < threads; ++i) { Scene* s = h->get(i); s->stats = Stats(); }
  h->get(0)->initRender();
  for (int i = 1; i < threads; ++i) { Scene* s = h->copies[i]; if (h->copies[0]->isPhtnMapRndrd) { s->photonTree = h->copies[0]->photonTree; s->isPhtnMapRndrd = true; } }
  int W = x1 - x0, H = y1 - y0;
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int tid) {
    Scene* s = h->copies[tid];
    for (int row = y0 + tid; row < y1; row += threads) for (int col = x0; col < x1; ++col) {
      Scene::PixelOut p = s->renderPixel(row, col); size_t i = (size_t)(row - y0) * W + (col - x0);
      if (argb) argb[i] = p.argb; if (hitPrim) hitPrim[i] = p.hitPrim; if (hitInst) hitInst[i] = p.hitInst;
      if (rgb) { rgb[3 * i] = p.rgb.x; rgb[3 * i + 1] = p.rgb.y; rgb[3 * i + 2] = p.rgb.z; } if (tOut) tOut[i] = p.t;
    }
  };
  (void)H;
  if (threads == 1) work(0);
  else { std::vector<std::thread> th; for (int i = 0; i < threads; ++i) th.emplace_back(work, i); for (auto& t : th) t.join(); }
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (stats) { Stats a; for (int i = 0; i < threads; ++i) a.add(h->copies[i]->stats);
    stats[0] = a.primary; stats[1] = a.shadow; stats[2] = a.reflect; stats[3] = a.refract; stats[4] = a.photonSeg; stats[5] = a.boxTests; stats[6] = a.primTests; stats[7] = a.boxTestsPrimary; stats[8] = a.primTestsPrimary; stats[9] = a.photonsStored; }
  return secs;
}

// Render the interleaved row chunks {c : c % chunkStride == chunkPhase} (chunkRows rows each) of the frame -- the pixel set a GPU rank with
// world = chunkStride, rank = chunkPhase renders (bench.py uses it as the bounded whole-frame sample of the reference arm).  argb: compact,
// chunks back to back, cols ints per row (may be NULL).  Rows are handed out dynamically to `threads` workers.  Returns wall seconds.
double orc_render_chunks(void* hv, int chunkRows, int chunkStride, int chunkPhase, int threads, int32_t* argb, uint64_t* stats) {
  OrcHandle* h = (OrcHandle*)hv;
  if (threads < 1) threads = 1; if (chunkRows < 1) chunkRows = 1; if (chunkStride < 1) chunkStride = 1;
  for (int i = 0; i < threads; ++i) { Scene* s = h->get(i); s->stats = Stats(); }
  h->get(0)->initRender();
  for (int i = 1; i < threads; ++i) { Scene* s = h->copies[i]; if (h->copies[0]->isPhtnMapRndrd) { s->photonTree = h->copies[0]->photonTree; s->isPhtnMapRndrd = true; } }
  const int rows = h->copies[0]->sceneRows, cols = h->copies[0]->sceneCols;
  std::vector<std::pair<int, int>> work;          // (absolute row, compact row)
  { int compact = 0; for (int c = chunkPhase; c * chunkRows < rows; c += chunkStride) for (int r = 0; r < chunkRows; ++r, ++compact) if (c * chunkRows + r < rows) work.push_back({c * chunkRows + r, compact}); }
  std::atomic<size_t> next{0};
  auto t0 = std::chrono::steady_clock::now();
  auto run = [&](int tid) { Scene* s = h->copies[tid];
    for (size_t k = next.fetch_add(1); k < work.size(); k = next.fetch_add(1)) for (int col = 0; col < cols; ++col) {
      Scene::PixelOut p = s->renderPixel(work[k].first, col); if (argb) argb[(size_t)work[k].second * cols + col] = p.argb; } };
  if (threads == 1) run(0);
  else { std::vector<std::thread> th; for (int i = 0; i < threads; ++i) th.emplace_back(run, i); for (auto& t : th) t.join(); }
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (stats) { Stats a; for (int i = 0; i < threads; ++i) a.add(h->copies[i]->stats);
    stats[0] = a.primary; stats[1] = a.shadow; stats[2] = a.reflect; stats[3] = a.refract; stats[4] = a.photonSeg; stats[5] = a.boxTests; stats[6] = a.primTests; stats[7] = a.boxTestsPrimary; stats[8] = a.primTestsPrimary; stats[9] = a.photonsStored; }
  return secs;
}

// Trace explicit world-space rays (closest hit only): out per ray = {hitPrim, hitInst}, t
int orc_trace_rays(void* hv, long long n, const double* org, const double* dir, int32_t* ids, double* tOut) {
  Scene* s = ((OrcHandle*)hv)->get(0);
  for (long long i = 0; i < n; ++i) {
    SKey k; Ray r(s, Vec3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), Vec3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), 0, k);
    RayHit h = s->findClosestRayHit(r);
    ids[2 * i] = h.isHit ? h.obj->primSerial : -1; ids[2 * i + 1] = h.isHit ? h.instSerial : -1; if (tOut) tOut[i] = h.isHit ? h.t : 0;
  }
  return 0;
}

// Pre-order dump of top-level object `objIdx` if it is a BVH: inner = -1; leaf = -2, count, ids...
// id = primSerial, or 0x40000000|instSerial for instances. Returns words written (or needed if > cap), -1 if not a BVH.
static void dumpNode(BVH* b, std::vector<int32_t>& o) {
  if (b->isLeaf) { o.push_back(-2); o.push_back((int32_t)b->leafVals->objList.size());
    for (Geom* g : b->leafVals->objList) o.push_back(g->type == G_INSTANCE ? (0x40000000 | static_cast<Instance*>(g)->instSerial) : g->primSerial); }
  else { o.push_back(-1); dumpNode(b->leftChild, o); dumpNode(b->rightChild, o); }
}
long long orc_dump_bvh(void* hv, int objIdx, int32_t* out, long long cap, double* rootBox /*6*/) {
  Scene* s = ((OrcHandle*)hv)->get(0);
  if (objIdx < 0 || objIdx >= (int)s->objList.size() || s->objList[objIdx]->type != G_BVH) return -1;
  BVH* b = static_cast<BVH*>(s->objList[objIdx]); std::vector<int32_t> o; dumpNode(b, o);
  if (rootBox) { rootBox[0] = b->bbox.minVals.x; rootBox[1] = b->bbox.minVals.y; rootBox[2] = b->bbox.minVals.z; rootBox[3] = b->bbox.maxVals.x; rootBox[4] = b->bbox.maxVals.y; rootBox[5] = b->bbox.maxVals.z; }
  for (size_t i = 0; i < o.size() && (long long)i < cap; ++i) out[i] = o[i];
  return (long long)o.size();
}

// Evaluate the diffuse texture of shader `serial` at (local hit, world hit) points: out = rgb * diffConst
int orc_eval_texture(void* hv, int serial, long long n, const double* hitLoc, const double* fwdLoc, double* out) {
  Scene* s = ((OrcHandle*)hv)->get(0); if (serial < 0 || serial >= (int)s->allShaders.size()) return -1;
  Shader* sh = s->allShaders[serial];
  if (sh->txtr->topImage()) return -2;   // image textures need a primitive for (u,v); probe procedural textures only
  for (long long i = 0; i < n; ++i) {
    RayHit h; h.isHit = true; h.hitLoc = Vec3(hitLoc[3 * i], hitLoc[3 * i + 1], hitLoc[3 * i + 2]); h.fwdTransHitLoc = Vec3(fwdLoc[3 * i], fwdLoc[3 * i + 1], fwdLoc[3 * i + 2]);
    sh->txtr->getDiffTxtrColor(h, sh->diffuseColor, sh->simple ? 1.0 : sh->diffConst, out + 3 * i);
  }
  return 0;
}

double orc_u01(uint64_t seed, uint32_t stream, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return u01(seed, stream, a, b, c, d); }
void orc_java_random(long long seed, int n, double* out) { JavaRandom r; r.setSeed(seed); for (int i = 0; i < n; ++i) out[i] = r.nextDouble(); }
float orc_perlin(float x, float y, float z) { return perlin().noise(x, y, z); }
// current matrix-stack product after interpreting a scene (KAT for the transform builders): object `objIdx` CTM (16 doubles, row-major)
int orc_obj_ctm(void* hv, int objIdx, double* out16) {
  Scene* s = ((OrcHandle*)hv)->get(0); if (objIdx < 0 || objIdx >= (int)s->objList.size()) return -1;
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out16[4 * r + c] = s->objList[objIdx]->ctm->glbl.m[r][c];
  return 0;
}
int orc_warnings(void* hv, char* buf, int len) {
  Scene* s = ((OrcHandle*)hv)->get(0); std::string a; for (auto& w : s->warnings) { a += w; a += "\n"; }
  if (buf && len > 0) { strncpy(buf, a.c_str(), len - 1); buf[len - 1] = 0; } return (int)s->warnings.size();
}

}  // extern "C"
