// Flat (pointer-free) scene description shared by the host flattener and the sm_100a kernels.
// Everything the reference keeps as an object graph (myScene.java:50-56 lists of myGeomBase, each with
// a CTM 4-array, myGeomBase.java:19-24) is laid out here as index-linked POD arrays that are uploaded
// to HBM verbatim.  All geometry is IEEE double, as in the reference.
#pragma once
#include <stdint.h>

namespace drt {

// {M, M^-1, (M^-1)^T} of a reference CTM array (DistRayTracer.java:399); M^T is never read on the hot path.
struct FXform { double m[16]; double inv[16]; double adj[16]; };

enum PrimType : int32_t { PT_SPHERE = 1, PT_MOVSPHERE = 2, PT_HCYL = 3, PT_CYL = 4, PT_TRI = 5, PT_QUAD = 6, PT_PLANE = 7, PT_BOX = 8,
                           PT_TORUS = 9, PT_QUADRIC = 10 };      // extension primitives (`extensions on`): the reference only has a torus stub that never hits (myImpObject.java:330-390)
enum : int32_t { PF_INVERTED = 1 };

// One renderable primitive (mySceneObject subclasses). `data` indexes the double pool `pdata`:
//  SPHERE     : o[3] r[3]                                  (myImpObject.java:35-46)
//  MOVSPHERE  : o0[3] r[3] o1[3]                           (:144-155)
//  HCYL       : o[3] radX radZ yTop yBottom                (:157-171)
//  CYL        : o[3] radX radZ yTop yBottom capTop[4] capBtm[4]   (:239-256)
//  TRI / QUAD : two winding states, each: verts[3n] N[3] D ; then uvA[2n] uvB[2n]   (myPlanarObject.java:44-100)
//  PLANE      : three states (as given / after 1 flip / after 2 flips), each N[3] D            (:236-270, :71-88)
//  BOX        : min[3] max[3]                              (mySceneObject.java:59-67)
struct FPrim { int32_t type, flags, xform, shader, data, serial, pad0, pad1; };

#define DRT_TRI_STATE 13   // doubles per triangle winding state
#define DRT_QUAD_STATE 16

enum ObjKind : int32_t { OK_PRIM = 0, OK_INSTANCE = 1, OK_LIST = 2, OK_BVH = 3 };

// entry of the top-level object list (myScene.objList) or of a flat list / BVH leaf (myGeomList.objList).
// xform    = the object's own CTM array (obj.CTMara)
// hitXform = CTM the hit record ends up with: own CTM at top level, list.CTM x child.CTM inside a list
struct FObjRef { int32_t kind, idx, xform, hitXform; };

// myInstance (mySceneObject.java:95-145): base object + own CTM + optional shader override
struct FInstance { int32_t baseKind, baseIdx, xform, shader, serial, pad0, pad1, pad2; };

// myGeomList (myGeomBase.java:251-306): own CTM, own box (children's min/max corners only), children in file order
struct FList { int32_t xform, childStart, childCount, pad; double bmin[3], bmax[3]; };

// myBVH root (myGeomBase.java:309-423). root >= 0: inner node index; root < 0: ~list index (single leaf)
// "fast" BVHs (every leaf holds only triangles that share one CTM) additionally own a contiguous run of packed triangle records
// (FTri, DFS leaf order) and tri-leaf codes in their nodes; fastRoot indexes DScene::fnodes (the same nodes in reference-topology
// modes, the GPU-built LBVH nodes in DRT_ACCEL_LBVH mode).
// absMax: per-axis bound of |coordinate| over every node box the fast descent can load (error budget of the FP32 pre-test, dev_isect.cuh)
struct FBvh { int32_t xform, root, nodeCount, dropped; double bmin[3], bmax[3]; int32_t fast, triXform, triHitXform, fastRoot, triStart, triCount, pad0, pad1; float absMax[3]; int32_t pad2; };

// inner node with both child boxes; child >= 0 inner node, child < 0 -> ~list index. 128 B, 16 B aligned (128-bit loads).
// triL / triR: when the child is a pure-triangle leaf of a fast BVH, (first FTri index << 3) | count ; else -1.
struct alignas(16) FNode { double lmin[3], lmax[3], rmin[3], rmax[3]; int32_t left, right, triL, triR; int32_t pad2[4]; };

// FP32 mirror of an FNode (same index): both child boxes rounded to nearest float + the links.  64 B = 4 x 128-bit loads.  The lean descent
// decides a box on these when the decision is certain under a rigorous error bound and reads the FP64 node only when it is not.
struct alignas(16) FNode32 { float lmin[3], lmax[3], rmin[3], rmax[3]; int32_t left, right, triL, triR; };

// packed triangle of a fast BVH, 128 B, read with 8 x 128-bit loads. Winding state 0 as given; state 1 (myPlanarObject.invertNormal,
// :71-88) = vertices in reverse order, normal -N (bitwise), plane offset Drev (formed from the reversed first vertex, so stored).
struct alignas(16) FTri { double v[9]; double N[3]; double D, Drev; int32_t prim, pad[3]; };   // pad[0] = rank in the reference's DFS order (tie-break)

enum LightType : int32_t { LT_POINT = 0, LT_SPOT = 1, LT_DISK = 2 };
struct FLight {
  int32_t type, xform, pad0, pad1;
  double color[3], origin[3], orient[3];
  double innerRad, outerRad, radDiff, radius;
  double tangent[3];           // spot: oPhAxis, disk: surfTangent (myLight.java:156,246)
};

enum TexKind : int32_t { TK_NONE = 0, TK_IMAGE = 1, TK_NOISE = 2, TK_BASEWOOD = 3, TK_MARBLE = 4, TK_CELL = 5, TK_WOOD = 6 };
struct FTexture {
  int32_t kind, imgTop, numOctaves, numPtsDist, distFunc, roiFunc, rndColors, useFwdTrans, colorStart, colorCount, pad0, pad1;
  double scale, turbMult, colorScale, colorMult, avgNumPerCell, mortarThresh;
  double periodMult[3], periodMag;
  double pdf[14];              // cumulative Poisson table (myTextureHandler.java:426-432)
};
struct FImage { int32_t w, h; int64_t offset; };

enum : int32_t { SF_HAS_CAUSTIC = 1, SF_USE_PHOTON = 2, SF_IS_CAUSTIC_PHTN = 4, SF_SIMPLE = 8 };
struct FShader {
  int32_t flags, tex, serial, pad;
  double diff[3], amb[3], spec[3], perm[3], kreflClr[3];
  double phong, KRefl, KTrans, currPerm, diffConst, avgDiff;
  double phtnDiffScl[3], phtnPermClr[3];
};

enum CamKind : int32_t { CAM_FOV = 0, CAM_FISHEYE = 1, CAM_ORTHO = 2 };
struct FGlobals {
  int32_t cols, rows, spp, camKind, hasDof, hasSky, numRays, numPhotonRays;
  int32_t numTop, numLights, skyImage, photonKind;        // photonKind: 0 none, 1 caustic, 2 diffuse
  int32_t numPhotonsCast, kNhood, pad0, pad1;
  double eye[3], viewZ, rayXOffset, rayYOffset, xStart, yStart, fishMult, aperatureHlf, orthPerRow, orthPerCol;
  double lensRadius, focalD;                              // focal plane z = -focalD (myScene.java:805-810)
  double bg[3];
  double skyOrigin[3], skyRad[3];
  double phMaxDist2, causticPwrMult, diffusePwrMult;
  uint64_t seed;
};

}  // namespace drt
