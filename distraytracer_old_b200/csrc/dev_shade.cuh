// Surface evaluation: normals, texture lookups (image, Perlin family, Worley), Fresnel split.
//
// Replaces (reference file:line, /root/reference/src/rayTracerDistAccelShdPhtnMap/):
//   myRay.java:119-125,168-175        hit point / normal through the hit CTM (M and M^-T)
//   myPlanarObject.java:130-136,178-186 ; myImpObject.java:68-74,97-122,196-202,305-312 ; myGeomBase.java:175-186
//   myTextureHandler.java:84-117 (bilinear image), :223-294 (noise, turbulence, colour ramp), :312-377 (wood, wood2, marble),
//                         :445-487 + :515-680 (Worley "stone" with java.util.Random cell seeding)
//   DistRayTracer.java:234-310 (Perlin noise in float), :21-29 (Worley constants)
//   myObjShader.java:78-96 (Fresnel terms, reflection direction), :157-276 / :503-631 (dielectric split)
#pragma once
#include "dev_isect.cuh"

namespace drt {

__device__ __forceinline__ D3 clampColor1(D3 c) { return d3(jminD(1, c.x), jminD(1, c.y), jminD(1, c.z)); }     // myColor ctor
__device__ __forceinline__ D3 colorOfArgb(int32_t c) { return clampColor1(d3(((c >> 16) & 0xFF) / 255.0, ((c >> 8) & 0xFF) / 255.0, (c & 0xFF) / 255.0)); }
__device__ __forceinline__ D3 lerpColor(D3 A, double t, D3 B) { return clampColor1(d3((A.x + t * (B.x - A.x)), (A.y + t * (B.y - A.y)), (A.z + t * (B.z - A.z)))); }
__device__ __forceinline__ int32_t packArgb(D3 c) { return (int32_t)(((uint32_t)255 << 24) + ((uint32_t)j2iD(c.x * 255) << 16) + ((uint32_t)j2iD(c.y * 255) << 8) + (uint32_t)j2iD(c.z * 255)); }

// object-space normal (getNormalAtPoint of each primitive class)
__device__ inline D3 primNormal(const DScene& S, const FPrim& P, D3 pt, int arg0, int arg1, int state) {
  const double* __restrict__ q = S.pdata + P.data; D3 n;
  switch (P.type) {
    case PT_SPHERE: case PT_MOVSPHERE: n = norm3(d3(pt.x - q[0], pt.y - q[1], pt.z - q[2])); if (P.flags & PF_INVERTED) n = scale3(n, -1.0); return n;   // (moving spheres use origin0 too, :68-74)
    case PT_TRI: { const double* s = q + state * DRT_TRI_STATE; return norm3(d3(s[9], s[10], s[11])); }
    case PT_QUAD: { const double* s = q + state * DRT_QUAD_STATE; return norm3(d3(s[12], s[13], s[14])); }
    case PT_PLANE: return norm3(d3(q[4 * state], q[4 * state + 1], q[4 * state + 2]));
    case PT_HCYL: n = (arg0 == 1) ? d3((q[0] - pt.x), 0, (q[2] - pt.z)) : d3((pt.x - q[0]), 0, (pt.z - q[2])); n = norm3(n); if (P.flags & PF_INVERTED) n = scale3(n, -1); return n;
    case PT_CYL: if (arg0 >= 2) n = d3((pt.x - q[0]), 0, (pt.z - q[2])); else n = d3(q[7 + 4 * arg0], q[8 + 4 * arg0], q[9 + 4 * arg0]); n = norm3(n); if (P.flags & PF_INVERTED) n = scale3(n, -1); return n;
    case PT_QUADRIC: {        // gradient; the far root (arg0 = 1) is seen from inside: flipped like the inside of a hollow cylinder
      n = d3((((2 * q[0]) * pt.x) + (q[3] * pt.y)) + ((q[4] * pt.z) + q[6]), (((2 * q[1]) * pt.y) + (q[3] * pt.x)) + ((q[5] * pt.z) + q[7]), (((2 * q[2]) * pt.z) + (q[4] * pt.x)) + ((q[5] * pt.y) + q[8]));
      n = norm3(n); if (arg0 == 1) n = scale3(n, -1); if (P.flags & PF_INVERTED) n = scale3(n, -1); return n; }
    case PT_TORUS: {
      const double x = pt.x - q[0], y = pt.y - q[1], z = pt.z - q[2], m = sqrt((x * x) + (z * z));
      n = (m > 0) ? d3(x - ((q[3] * x) / m), y, z - ((q[3] * z) / m)) : d3(0, y, 0);
      n = norm3(n); if (P.flags & PF_INVERTED) n = scale3(n, -1); return n; }
    case PT_BOX: switch (arg1) { case 0: return d3(-1, 0, 0); case 1: return d3(0, -1, 0); case 2: return d3(0, 0, -1); case 3: return d3(1, 0, 0); case 4: return d3(0, 1, 0); case 5: return d3(0, 0, 1); default: return d3(0, 0, -1); }
  }
  return d3(0, 0, 1);
}

// ---- Perlin noise, float arithmetic
__device__ __constant__ unsigned char c_perm[512];
// gradient table {1,1,0},{-1,1,0},{1,-1,0},{-1,-1,0},{1,0,1},{-1,0,1},{1,0,-1},{-1,0,-1},{0,1,1},{0,-1,1},{0,1,-1},{0,-1,-1} (DistRayTracer.java:286)
// packed 2 bits per component (value + 1) so that the lookup is a shift, not a dynamically indexed local array
__device__ __forceinline__ float pgrad(int gi, float x, float y, float z) {
  const uint32_t GX = 0x552222u, GY = 0x22550Au, GZ = 0x0A0A55u; const int sh = 2 * gi;
  const float gx = (float)((int)((GX >> sh) & 3u) - 1), gy = (float)((int)((GY >> sh) & 3u) - 1), gz = (float)((int)((GZ >> sh) & 3u) - 1);
  return gx * x + gy * y + gz * z;
}
__device__ __forceinline__ float pmix(float a, float b, float t) { return (1 - t) * a + t * b; }
__device__ __forceinline__ float pfade(float t) { return t * t * t * (t * (t * 6 - 15) + 10); }
__device__ inline float perlin3(float x, float y, float z) {
  int X = fastfloorF(x), Y = fastfloorF(y), Z = fastfloorF(z);
  x = x - X; y = y - Y; z = z - Z; X &= 255; Y &= 255; Z &= 255;
  int gi000 = c_perm[X + c_perm[Y + c_perm[Z]]] % 12, gi001 = c_perm[X + c_perm[Y + c_perm[Z + 1]]] % 12,
      gi010 = c_perm[X + c_perm[Y + 1 + c_perm[Z]]] % 12, gi011 = c_perm[X + c_perm[Y + 1 + c_perm[Z + 1]]] % 12,
      gi100 = c_perm[X + 1 + c_perm[Y + c_perm[Z]]] % 12, gi101 = c_perm[X + 1 + c_perm[Y + c_perm[Z + 1]]] % 12,
      gi110 = c_perm[X + 1 + c_perm[Y + 1 + c_perm[Z]]] % 12, gi111 = c_perm[X + 1 + c_perm[Y + 1 + c_perm[Z + 1]]] % 12;
  float n000 = pgrad(gi000, x, y, z), n100 = pgrad(gi100, x - 1, y, z), n010 = pgrad(gi010, x, y - 1, z), n110 = pgrad(gi110, x - 1, y - 1, z),
        n001 = pgrad(gi001, x, y, z - 1), n101 = pgrad(gi101, x - 1, y, z - 1), n011 = pgrad(gi011, x, y - 1, z - 1), n111 = pgrad(gi111, x - 1, y - 1, z - 1);
  float u = pfade(x), v = pfade(y), w = pfade(z);
  return pmix(pmix(pmix(n000, n100, u), pmix(n010, n110, u), v), pmix(pmix(n001, n101, u), pmix(n011, n111, u), v), w);
}

// java.util.Random
struct JRand {
  uint64_t s;
  __device__ void setSeed(int32_t seed) { s = ((uint64_t)(int64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
  __device__ int next(int bits) { s = (s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1); return (int)(int64_t)(s >> (48 - bits)); }
  __device__ double nextDouble() { int64_t a = next(26); int64_t b = next(27); return (double)((a << 27) + b) * (1.0 / 9007199254740992.0); }
};

// colour ramp with noise-perturbed multipliers (getClrAra, myTextureHandler.java:277-294)
__device__ inline D3 rampColor(const DScene& S, const FTexture& T, double distVal, D3 rawPt, int i0, int i1) {
  D3 pt = scale3(rawPt, T.colorScale); double rm0 = 1.0, rm1 = 1.0, rm2 = 1.0;
  if (T.rndColors) {
    rm0 = 1.0 + (T.colorMult * (double)perlin3((float)pt.x, (float)pt.z, (float)pt.y));
    rm1 = 1.0 + (T.colorMult * (double)perlin3((float)pt.y, (float)pt.x, (float)pt.z));
    rm2 = 1.0 + (T.colorMult * (double)perlin3((float)pt.z, (float)pt.y, (float)pt.x));
  }
  if (i0 >= T.colorCount) i0 = T.colorCount - 1; if (i1 >= T.colorCount) i1 = T.colorCount - 1;
  const double* c0 = S.texColors + 3 * (T.colorStart + i0); const double* c1 = S.texColors + 3 * (T.colorStart + i1);
  return d3(jmaxD(0, jminD(1.0, (c0[0]) + rm0 * distVal * ((c1[0]) - (c0[0])))), jmaxD(0, jminD(1.0, (c0[1]) + rm1 * distVal * ((c1[1]) - (c0[1])))), jmaxD(0, jminD(1.0, (c0[2]) + rm2 * distVal * ((c1[2]) - (c0[2])))));
}
__device__ __forceinline__ double turbulence(D3 t, int oct, bool absolute) {
  double res = 0, f = 1.0, a = 1.0;
  for (int i = 0; i < oct; ++i) { float n = perlin3((float)(t.x * f), (float)(t.y * f), (float)(t.z * f)); res += (absolute ? fabs((double)n) : (double)n) * a; a *= .5; f *= 1.92; }
  return res;
}
__device__ __forceinline__ double fixDist(double d) { if (d < 0) d *= -1; if (d > 1.0) d = 1.0 / d; return d; }

// texture coordinates in texel units (findTxtrCoords)
__device__ inline void primTexCoords(const DScene& S, const FPrim& P, D3 pt, int state, double time, int w, int h, double& u, double& v) {
  const double* __restrict__ q = S.pdata + P.data; u = 0; v = 0;
  if (P.type == PT_SPHERE || P.type == PT_MOVSPHERE) {
    D3 c = d3(q[0], q[1], q[2]);
    if (P.type == PT_MOVSPHERE) { D3 bMa = d3(q[6] - c.x, q[7] - c.y, q[8] - c.z); c = d3(c.x + time * bMa.x, c.y + time * bMa.y, c.z + time * bMa.z); }
    double a1v = (pt.y - c.y) / q[4]; a1v = (a1v > 1) ? 1 : (a1v < -1) ? -1 : a1v;
    v = (h - 1) * acos(a1v) / DRT_PI;
    double shWm1 = w - 1, z1 = (pt.z - c.z), qq = v / (h - 1);
    double a0 = (pt.x - c.x) / q[3]; a0 = (a0 > 1) ? 1 : (a0 < -1) ? -1 : a0;
    double a1 = sin(qq * DRT_PI), a2 = (fabs(a1) < DRT_EPS) ? 1 : a0 / a1;
    u = (z1 <= DRT_EPS) ? ((shWm1 * (acos(a2)) / (DRT_TWO_PI_F)) + shWm1 / 2.0f) : shWm1 - ((shWm1 * (acos(a2)) / (DRT_TWO_PI_F)) + shWm1 / 2.0f);
    u = (u < 0) ? 0 : (u > shWm1) ? shWm1 : u;
  } else if (P.type == PT_TRI || P.type == PT_QUAD) {
    const int n = (P.type == PT_TRI) ? 3 : 4, stride = (P.type == PT_TRI) ? DRT_TRI_STATE : DRT_QUAD_STATE;
    const double* s = q + state * stride; const double* uv = q + 2 * stride + state * 2 * n;
    D3 P0 = d3(s[0], s[1], s[2]), P1 = d3(s[3], s[4], s[5]), P2 = d3(s[6], s[7], s[8]);
    D3 e0 = d3(P1.x - P0.x, P1.y - P0.y, P1.z - P0.z);                       // P2P[0]
    D3 l2 = (n == 3) ? P0 : d3(s[9], s[10], s[11]);                           // vertex after index 2
    D3 e2 = d3(l2.x - P2.x, l2.y - P2.y, l2.z - P2.z);                        // P2P[2]
    D3 e20 = scale3(e2, -1.0);                                                // P2P0
    double d0 = dot3(e0, e0), d2 = dot3(e2, e2), dn = -dot3(e0, e2);
    double inv = 1.0 / ((d0 * d2) - (dn * dn));
    D3 v2 = sub3(pt, P0);
    double dot20 = dot3(v2, e0), dot21 = dot3(v2, e20);
    double cu = ((d2 * dot20) - (dn * dot21)) * inv, cv = ((d0 * dot21) - (dn * dot20)) * inv, cw = 1 - cu - cv;
    double uu = uv[0] * cw + uv[2] * cu + uv[4] * cv, vv = uv[1] * cw + uv[3] * cu + uv[5] * cv;
    u = uu * (w - 1); v = (1 - vv) * (h - 1);
  }
}

// diffuse texture colour x diffConst (getDiffTxtrColor of every texture class)
__device__ inline D3 evalTexture(const DScene& S, const FShader& sh, const FPrim& P, D3 hitLoc, D3 fwdLoc, int state, double time) {
  const FTexture T = S.textures[sh.tex];
  const double k = (sh.flags & SF_SIMPLE) ? 1.0 : sh.diffConst;
  D3 diffuse = d3(sh.diff[0], sh.diff[1], sh.diff[2]);
  if (T.kind == TK_NONE || (T.kind == TK_IMAGE && T.imgTop < 0)) return d3(diffuse.x * k, diffuse.y * k, diffuse.z * k);
  if (T.kind == TK_IMAGE) {
    const FImage im = S.images[T.imgTop]; const int32_t* __restrict__ px = S.texels + im.offset;
    double u, v; primTexCoords(S, P, hitLoc, state, time, im.w, im.h, u, v);
    int ui = j2iD(u), vi = j2iD(v);
    long long n = (long long)im.w * im.h, i00 = (long long)vi * im.w + ui, i10 = i00 + im.w, i01 = i00 + 1, i11 = i10 + 1;
    // Java would throw past the last texel; clamp (documented deviation)
    i00 = i00 < 0 ? 0 : (i00 >= n ? n - 1 : i00); i10 = i10 < 0 ? 0 : (i10 >= n ? n - 1 : i10); i01 = i01 < 0 ? 0 : (i01 >= n ? n - 1 : i01); i11 = i11 < 0 ? 0 : (i11 >= n ? n - 1 : i11);
    D3 c00 = colorOfArgb(px[i00]), c10 = colorOfArgb(px[i10]), c01 = colorOfArgb(px[i01]), c11 = colorOfArgb(px[i11]);
    double fu = u - ui, fv = v - vi;
    D3 c0 = lerpColor(c00, fu, c01), c1 = lerpColor(c10, fu, c11), c = lerpColor(c0, fv, c1);
    return d3(c.x * k, c.y * k, c.z * k);
  }
  D3 hl = T.useFwdTrans ? fwdLoc : hitLoc; D3 out;
  if (T.kind == TK_NOISE) {
    D3 s = scale3(hl, T.scale); double res = T.turbMult * (double)perlin3((float)s.x, (float)s.y, (float)s.z), val = .5 * res + .5; out = d3(val, val, val);
  } else if (T.kind == TK_BASEWOOD) {
    D3 s = scale3(hl, T.scale); double res = (double)perlin3((float)s.x, (float)s.y, (float)s.z);
    double sq = sqrt((s.x * s.x) * T.periodMult[0] + (s.y * s.y) * T.periodMult[1] + (s.z * s.z) * T.periodMult[2]) + T.turbMult * res;
    double dv = sin(sq * T.periodMag); dv *= 1.1; dv += .5; dv = (dv < 0 ? 0 : (dv > 1 ? 1 : dv));
    out = rampColor(S, T, dv, hitLoc, 0, 1);
  } else if (T.kind == TK_WOOD) {
    D3 s = scale3(hl, T.scale); double res = turbulence(s, T.numOctaves, false);
    double sq = sqrt((s.x * s.x) * T.periodMult[0] + (s.y * s.y) * T.periodMult[1] + (s.z * s.z) * T.periodMult[2]) + T.turbMult * res;
    double dv = sin(sq * T.periodMag); dv = 1 - (dv < 0 ? 0 : dv);
    out = rampColor(S, T, dv, s, 0, 1);
  } else if (T.kind == TK_MARBLE) {
    D3 s = scale3(hl, T.scale); double res = turbulence(s, T.numOctaves, true);
    double spt = (s.x * T.periodMult[0] + s.y * T.periodMult[1] + s.z * T.periodMult[2]) / T.periodMag + T.turbMult * res;
    double dv = .5 * sin(spt) + .5;
    out = rampColor(S, T, dv, s, 0, 1);
  } else {   // TK_CELL
    const int nb[27][3] = {{0,0,0},{0,0,1},{0,0,-1},{0,1,0},{0,1,1},{0,1,-1},{0,-1,0},{0,-1,1},{0,-1,-1},{1,0,0},{1,0,1},{1,0,-1},{1,1,0},{1,1,1},{1,1,-1},{1,-1,0},{1,-1,1},{1,-1,-1},{-1,0,0},{-1,0,1},{-1,0,-1},{-1,1,0},{-1,1,1},{-1,1,-1},{-1,-1,0},{-1,-1,1},{-1,-1,-1}};
    D3 hv = scale3(hl, T.scale);
    int cx0 = fastfloorD(hv.x), cy0 = fastfloorD(hv.y), cz0 = fastfloorD(hv.z);
    // the reference sorts every distance in a map (equal keys collapse, last writer keeps the cell) and reads only the
    // first numPtsDist keys: keep the K smallest distinct keys while streaming
    const int KMAX = 8; int K = T.numPtsDist < 1 ? 1 : (T.numPtsDist > KMAX ? KMAX : T.numPtsDist);
    double keys[KMAX]; int32_t seeds[KMAX]; int cnt = 0;
    JRand gen;
    for (int i = 0; i < 27; ++i) {
      int cx = cx0 + nb[i][0], cy = cy0 + nb[i][1], cz = cz0 + nb[i][2];
      int32_t seed = (int32_t)((uint32_t)cx * 1572869u + (uint32_t)cy * 6291469u + (uint32_t)cz);
      gen.setSeed(seed);
      double prob = gen.nextDouble(); int best = -1;
      for (int j = 0; j < 14; ++j) if (T.pdf[j] < prob) best = j;
      int numPoints = (best < 0 ? 0 : best) + 1;
      for (int j = 0; j < numPoints; ++j) {
        double px = cx + gen.nextDouble(), py = cy + gen.nextDouble(), pz = cz + gen.nextDouble();
        double d = (T.distFunc == 0) ? (fabs(hv.x - px) + fabs(hv.y - py) + fabs(hv.z - pz)) : sqrt(((hv.x - px) * (hv.x - px)) + ((hv.y - py) * (hv.y - py)) + ((hv.z - pz) * (hv.z - pz)));
        int pos = cnt; bool dup = false;
        for (int m = 0; m < cnt; ++m) { if (d == keys[m]) { seeds[m] = seed; dup = true; break; } if (d < keys[m]) { pos = m; break; } }
        if (dup || pos >= K) continue;
        int last = (cnt < K) ? cnt : K - 1;
        for (int m = last; m > pos; --m) { keys[m] = keys[m - 1]; seeds[m] = seeds[m - 1]; }
        keys[pos] = d; seeds[pos] = seed; if (cnt < K) ++cnt;
      }
    }
    double dist = 0; int i = 0, modVal = -1; const int n = T.numPtsDist;
    switch (T.roiFunc) {
      case 0: for (int m = 0; m < cnt; ++m) { dist += keys[m]; i++; if (i >= n) break; } dist = fixDist(dist); break;
      case 2: for (int m = 0; m < cnt; ++m) { dist += 1.0 / (modVal * keys[m]); i++; if (i >= n) break; modVal *= -1; } dist = fixDist(dist); break;
      case 3: for (int m = 0; m < cnt; ++m) { dist += (modVal * pow(keys[m], (double)(++i))); if (i >= n) break; modVal *= -1; } dist = fixDist(dist); break;
      case 4: for (int m = 0; m < cnt; ++m) { dist += (modVal * log(1 + keys[m])); i++; if (i >= n) break; modVal *= -1; } dist = fixDist(dist); break;
      case 5: for (int m = 0; m < cnt; ++m) { dist += pow(keys[m], (double)(++i)); if (i >= n) break; } dist = fixDist(dist); break;
      case 6: for (int m = 0; m < cnt; ++m) { dist += log(1 + keys[m]); i++; if (i >= n) break; } break;
      case 7: for (int m = 0; m < cnt; ++m) { dist += pow(keys[m], (double)(-(++i))); if (i >= n) break; } dist = fixDist(dist); break;
      case 8: for (int m = 0; m < cnt; ++m) { dist += 1.0 / log(1 + keys[m]); i++; if (i >= n) break; } dist = fixDist(dist); break;
      default: for (int m = 0; m < cnt; ++m) { dist += (modVal * keys[m]); i++; if (i >= n) break; modVal *= -1; } dist = fixDist(dist); break;
    }
    dist = (dist < 0 ? 0 : dist > 1 ? 1 : dist);
    int brick = 2;
    if (dist < T.mortarThresh) brick = 0;
    else { gen.setSeed(seeds[0]); double res = gen.nextDouble(); brick = 2 * (1 + (fastfloorD(((T.colorCount / 2) - 1) * res))); }
    out = rampColor(S, T, .65, hv, brick, brick + 1);
  }
  if (fabs(k - 1.0) > DRT_EPS) out = d3(out.x * k, out.y * k, out.z * k);
  return out;
}

// ---- Fresnel split shared by the complex and "simple" shaders
struct Fres { D3 N, back; double n, c1, c2, ratio, oneM, mult; };
__device__ inline Fres fresnel(D3 rawDir, D3 objNorm, double matIdx, double rayIdx) {
  Fres f; f.back = scale3(rawDir, -1); f.N = objNorm; f.n = 1; f.ratio = 0; f.oneM = 1; f.mult = 1.0; f.c2 = 0;
  double n1 = 0, n2 = 0; bool TIR = false;
  f.c1 = dot3(f.back, f.N);
  if (f.c1 < DRT_EPS) { f.mult = -1.0; f.N = scale3(f.N, -1); }
  f.c1 = dot3(f.back, f.N);
  double thetaInc = acos(dot3(f.back, f.N) / (mag3(f.back) * mag3(f.N)));
  if (f.mult < 0) {
    double thetaCrit = asin(1 / matIdx);
    if (thetaInc < thetaCrit) { n1 = matIdx; n2 = 1; f.n = (n1 / n2); f.c2 = sqrt(1.0 - (f.n * f.n) * (1.0 - (f.c1 * f.c1))); }
    else { f.ratio = 1; f.oneM = 1 - f.ratio; TIR = true; f.c2 = 0; }
  } else { n1 = rayIdx; n2 = matIdx; f.n = (n1 / n2); f.c2 = sqrt(1.0 - (f.n * f.n) * (1.0 - (f.c1 * f.c1))); }
  if (!TIR) {
    double sA = sin(acos(f.c1)), cT = sqrt(1.0 - ((n1 / n2) * sA * sA));
    double a = n1 * f.c1, b = n2 * cT, nd = (a - b) / (a + b), rPerp = nd * nd;
    double a2 = n1 * cT, b2 = n2 * f.c1, nd2 = (a2 - b2) / (a2 + b2), rPar = nd2 * nd2;
    f.ratio = (rPerp + rPar) / 2.0; f.oneM = 1 - f.ratio;
  }
  return f;
}
__device__ __forceinline__ D3 reflDir(D3 eye, D3 n) { double dp = 2 * (dot3(eye, n)); D3 t = d3(n.x * dp, n.y * dp, n.z * dp); return norm3(sub3(t, eye)); }
__device__ __forceinline__ D3 refractDir(const Fres& f) { D3 u = scale3(f.back, f.n * -1); D3 nv = scale3(f.N, (f.n * f.c1) - f.c2); return norm3(add3(u, nv)); }

}  // namespace drt
