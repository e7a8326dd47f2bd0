// Host-side handle of the wavefront renderer (device memory owner + kernel launcher).
#pragma once
#include "host_scene.h"
#include <string>

namespace drt {

struct RenderStats {
  unsigned long long primary, shadow, reflect, refract, photonSeg;   // logical rays
  unsigned long long boxTests, primTests, boxTestsClosest, primTestsClosest;   // only when counters are enabled
  unsigned long long photonsStored;
  unsigned long long kernelLaunches;
  unsigned long long deferred, retries, hostSyncs;                    // rays / light records the lean kernels left to the generic pass; re-rendered frames; host syncs of the call
  double msTrace, msShade, msLight, msOther, msTotal;               // CUDA-event times of the last render call
};

struct RenderOutputs {           // device pointers, any may be null; sized cols*rows (rgb: 3x)
  int32_t* argb; int32_t* hitPrim; int32_t* hitInst; double* rgb; double* t;
};

class Renderer {
 public:
  Renderer(int device);
  ~Renderer();
  void upload(const HostScene& hs, bool sameScene = false);   // flat scene -> HBM (sameScene: a re-upload of what is already resident, keeps the depth hint)
  void setBatchRays(long long n) { batchRays_ = n; }
  void setCounters(bool on) { counters_ = on; }
  void setTraceMode(int m) { traceMode_ = m; }
  // render pixels [pix0, pix1) (row-major pixel indices) into the device buffers of `out` (indexed by absolute pixel)
  void renderRange(long long pix0, long long pix1, const RenderOutputs& out, RenderStats* stats);
  void renderChunks(long long pix0, long long pix1, int world, int rank, int chunkRows, const RenderOutputs& out, RenderStats* stats);
  // full frame to host memory (D2H inside)
  void renderToHost(int32_t* argbHost, int32_t* hitPrimHost, int32_t* hitInstHost, double* rgbHost, double* tHost, RenderStats* stats);
  // explicit world-space rays (closest hit only) for parity tests: ids = {primSerial, instSerial} per ray
  void traceRays(long long n, const double* orgHost, const double* dirHost, int32_t* idsHost, double* tHost);
  void evalTexture(int shaderIdx, long long n, const double* hitLocHost, const double* fwdLocHost, double* outHost);
  void emitPhotons(RenderStats* stats);               // builds the photon map if the scene asks for one
  long long getPhotons(double* out6Host, long long cap);
  void emitPhotonsRange(long long i0, long long i1, RenderStats* stats);          // photon indices [i0,i1) of every light, no grid build
  long long exportPhotonsDevice(double* dst6Dev, long long cap);                  // canonical-order records -> caller's device buffer; returns count
  void buildPhotonsFromDevice(const double* src6Dev, long long n, RenderStats* stats);   // replace the record set (e.g. all-gathered) and build the grid
  void probePhotons(long long n, const double* ptsHost, double* out5Host);       // per point: sum r,g,b of the k nearest, d^2 of the farthest, candidates visited
  void accelInfo(double out[4]) const;                 // LBVH build ms (CUDA events), triangles and nodes it covers, scene bytes in HBM
  int cols() const { return g_.cols; }
  int rows() const { return g_.rows; }
  int spp() const { return g_.spp; }
  void* stream() const { return stream_; }
  const FGlobals& globals() const { return g_; }
  struct Impl;

 private:
  Impl* impl_;
  FGlobals g_;
  int device_;
  void* stream_;
  long long batchRays_ = 8ll << 20;
  bool counters_ = false;
  int traceMode_ = 0;
  bool keepDepthHint_ = false;
};

}  // namespace drt
