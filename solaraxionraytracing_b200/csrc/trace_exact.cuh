// trace_exact.cuh — the FP64 per-ray pipeline ("exact" precision mode).
//
// Device restatement of traceAxion (src/raytracer.nim:1736-2221) organised for a GPU thread: per-shell and
// per-run constants come pre-derived (device_params.h), the duplicated getMirrorAngle work of rt:2003-2010 /
// 2030-2037 is shared with getVectoraAfterMirror, the identity rotations of lineIntersectsCylinder (the bore
// axis is the z axis, rt:286-294) are dropped, and the geometry is separated from the axion-mass dependent
// weight so a mass scan can reuse one traced ray. Every value that feeds a hit/miss decision is still produced
// by the same IEEE-754 operation sequence as the reference (this TU is compiled with -fmad=false), which is
// what makes the exit code of every ray reproducible bit for bit against the CPU oracle.
#pragma once
#include <cfloat>
#include <cmath>

#include "device_params.h"
#include "philox.cuh"

namespace sart {

constexpr double kPi = 3.141592653589793;
constexpr double kRadPerDeg = kPi / 180.0;

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ V3 normalize(V3 a) { return a * (1.0 / sqrt(dot(a, a))); }

// std/algorithm.lowerBound on a device array.
__device__ __forceinline__ int lower_bound(const double* __restrict__ a, int lo, int hi, double key) {
  int count = hi - lo;
  while (count > 0) {
    const int step = count >> 1;
    const int pos = lo + step;
    if (__ldg(a + pos) < key) { lo = pos + 1; count -= step + 1; }
    else count = step;
  }
  return lo;
}

// std/math.almostEqual, 4 ulp (rt:2055).
__device__ __forceinline__ bool almost_equal(double x, double y) {
  if (x == y) return true;
  const double diff = fabs(x - y);
  return diff <= DBL_EPSILON * fabs(x + y) * 4.0 || diff < DBL_MIN;
}

// numericalnim newLinear1D.eval on an irregular sorted grid; clamps (and flags) outside the grid.
__device__ __forceinline__ double eval_linear1d(const double* __restrict__ X, const double* __restrict__ Y, int n,
                                                double x, int& clamped) {
  if (n < 2) { clamped = 1; return n == 1 ? __ldg(Y) : 0.0; }
  const double x0 = __ldg(X), xn = __ldg(X + n - 1);
  if (!(x >= x0)) { clamped = 1; x = x0; }
  if (!(x <= xn)) { clamped = 1; x = xn; }
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(X + mid) <= x) lo = mid; else hi = mid;
  }
  const double xl = __ldg(X + lo), yl = __ldg(Y + lo);
  const double slope = (__ldg(Y + lo + 1) - yl) / (__ldg(X + lo + 1) - xl);
  return yl + (x - xl) * slope;
}

// numericalnim newBilinearSpline.eval on the regular (angle, energy) grid of one coating (rt:1567-1578).
__device__ __forceinline__ double eval_bilinear(const Params& P, const double* __restrict__ z, double x, double y,
                                                int& clamped) {
  const int nx = P.nAngles, ny = P.nReflEnergies;
  if (!(x >= P.angleMin)) { clamped = 1; x = P.angleMin; }
  if (!(x <= P.angleMax)) { clamped = 1; x = P.angleMax; }
  if (!(y >= P.reflEMin)) { clamped = 1; y = P.reflEMin; }
  if (!(y <= P.reflEMax)) { clamped = 1; y = P.reflEMax; }
  const double dx = P.reflDx, dy = P.reflDy;
  int i = int(floor((x - P.angleMin) / dx));
  int j = int(floor((y - P.reflEMin) / dy));
  if (i > nx - 2) i = nx - 2;
  if (j > ny - 2) j = ny - 2;
  const double xc = x - (P.angleMin + dx * double(i));
  const double yc = y - (P.reflEMin + dy * double(j));
  const double* r0 = z + size_t(i) * ny + j;
  const double z00 = __ldg(r0), z01 = __ldg(r0 + 1), z10 = __ldg(r0 + ny), z11 = __ldg(r0 + ny + 1);
  const double beta = (z10 - z00) / dx;
  const double gamma = (z01 - z00) / dy;
  const double delta = (z11 + z00 - z10 - z01) / (dx * dy);
  return z00 + beta * xc + gamma * yc + delta * xc * yc;
}

// ray ∩ plane z = zc from p1 along v: the point used by lineIntersectsCircle / getIntersectlineIntersectsCircle
// (rt:481-492, 529-534) and by the pointExit* expressions (rt:1850-1872), which are the same arithmetic.
__device__ __forceinline__ V3 plane_point(V3 p1, V3 v, double zc) {
  const double lambda = (zc - p1.z) / v.z;
  return p1 + lambda * v;
}
__device__ __forceinline__ bool inside(V3 q, double R) { return sqrt(q.x * q.x + q.y * q.y) < R; }

// rotateInX / rotateInY with pre-evaluated cos/sin (rt:338-354).
__device__ __forceinline__ V3 rot_in_x(V3 v, double c, double s, double off) {
  const double zt = v.z + (-off);
  V3 r = {v.x * c + zt * s, v.y, zt * c - v.x * s};
  r.z += off;
  return r;
}
__device__ __forceinline__ V3 rot_in_y(V3 v, double c, double s, double off) {
  const double zt = v.z + (-off);
  V3 r = {v.x, v.y * c - zt * s, zt * c + v.y * s};
  r.z += off;
  return r;
}

// Root choice shared by findPosCone / findPosParabolic / findPosHyperbolic (rt:646-658).
__device__ __forceinline__ V3 pick_root(V3 p, V3 d, double a, double half_b, double c, double zmin, double zmax) {
  const double sq = sqrt(half_b * half_b - a * c);
  const double root1 = (-half_b - sq) / a;
  const double root2 = (-half_b + sq) / a;
  double s;
  const double zr1 = p.z + root1 * d.z;
  const double zr2 = p.z + root2 * d.z;
  if (zr1 > zmin && zr1 < zmax) s = root1;
  else if (zr2 > zmin && zr2 < zmax) s = root2;
  else s = 0.0;
  return p + s * d;
}

// Reflection off a mirror whose (unnormalised) surface normal is n: returns the reflected unit vector and the
// grazing angle in rad (getVectoraAfterMirror rt:762-780 + getMirrorAngle rt:782-795).
__device__ __forceinline__ V3 reflect(V3 n, V3 from, V3 to, double& alpha) {
  const V3 vbm = normalize(to - from);
  const V3 axis = normalize(cross(n, vbm));
  alpha = asin(fabs(dot(n, vbm) / sqrt(dot(n, n))));
  const V3 vba = cross(vbm, axis);
  double s2, c2;
  sincos(2.0 * alpha, &s2, &c2);
  return vbm * c2 - vba * s2;
}

// ------------------------------------------------------------------------------------------------------------
// Geometry result: everything traceAxion knows about a ray before the axion-mass dependent weight.
struct Geo {
  int code;        // exit code of a geometric early return, or -1 if the ray reached the weight stage
  int shell;       // hitLayer
  int clamped;
  int windowMiss;  // rt:2139-2147 (decided here, reported after the weight like the reference)
  int strongback;  // window strip hit (rt:2167-2177); 2 = no strip loop ran (numberOfStrips == 0)
  int eIdx;        // index of the ray's energy in the tabulated energies (Monte Carlo rays), -1 if it is not a table value
  double energy, pathCB, ya, cosya, alpha1, alpha2, pitch, distancePipe;
  double xw, yw;   // pointDetectorWindow after the shifts (window frame)
  double deviationDet;
};

template <bool kFull>
__device__ __forceinline__ void trace_geometry(const Params& P, const Tables& T, V3 O, V3 E, double energy, Geo& g) {
  g.code = -1; g.shell = -1; g.clamped = 0; g.windowMiss = 0; g.strongback = 0; g.eIdx = -1;
  g.energy = energy; g.pathCB = 0.0; g.ya = 0.0; g.cosya = 0.0; g.alpha1 = 0.0; g.alpha2 = 0.0; g.pitch = 0.0;
  g.distancePipe = 0.0; g.xw = 0.0; g.yw = 0.0; g.deviationDet = 0.0;

  // ---- bore entry rt:1813-1843
  const V3 v = E - O;
  V3 entry = plane_point(O, v, 0.0);
  if (!inside(entry, P.radiusCB)) {
    // lineIntersectsCylinder rt:538-588 with the bore on the z axis (rotations are identities)
    const double lambda_dummy = (-1000.0 - O.z) / v.z;
    const V3 dummy = O + lambda_dummy * v;
    const V3 vd = E - dummy;
    const double factor = vd.x * vd.x + vd.y * vd.y;
    const double p = 2.0 * (dummy.x * vd.x + dummy.y * vd.y) / factor;
    const double q = (dummy.x * dummy.x + dummy.y * dummy.y - P.radiusCB * P.radiusCB) / factor;
    const double sq = sqrt(p * p / 4.0 - q);
    const double lambda_1 = -p / 2.0 + sq;
    const double lambda_2 = -p / 2.0 - sq;
    const V3 i1 = dummy + lambda_1 * vd;
    const V3 i2 = dummy + lambda_2 * vd;
    const bool valid1 = (i1.z > 0.0) && (i1.z < P.zExitCB);
    const bool valid2 = (i2.z > 0.0) && (i2.z < P.zExitCB);
    if (valid1 == valid2) { g.code = SART_EXIT_MISSED_BORE; return; }  // rt:598-600, 1825
    entry = valid1 ? i1 : i2;
  }
  {
    const V3 d = E - entry;
    g.pathCB = sqrt(dot(d, d));
  }
  // ---- exit of the cold bore, the two pipes rt:1846-1872
  V3 pExitCB = plane_point(O, v, P.zExitCB);
  if (!inside(pExitCB, P.radiusCB)) { g.code = SART_EXIT_CLIP_EXIT_CB; return; }
  const V3 v1 = pExitCB - E;
  const V3 pPipe1 = plane_point(E, v1, P.zPipe1);
  if (!inside(pPipe1, P.rPipe1)) { g.code = SART_EXIT_CLIP_PIPE_VT3; return; }
  const V3 v2 = pPipe1 - pExitCB;
  V3 pPipe2 = plane_point(pExitCB, v2, P.zPipe2);
  if (!inside(pPipe2, P.rPipe1)) { g.code = SART_EXIT_CLIP_PIPE_XRT; return; }  // quirk Q2: same radius

  // ---- telescope frame rt:1888-1905
  pExitCB.z -= P.zPipe2;
  pExitCB = rot_in_y(rot_in_x(pExitCB, P.cosTX, P.sinTX, P.halfLenTel), P.cosTY, P.sinTY, P.halfLenTel);
  pExitCB.x -= P.oeX; pExitCB.y -= P.oeY; pExitCB.z -= 0.0;
  pPipe2.z -= P.zPipe2;
  pPipe2 = rot_in_y(rot_in_x(pPipe2, P.cosTX, P.sinTX, P.halfLenTel), P.cosTY, P.sinTY, P.halfLenTel);
  pPipe2.x -= P.oeX; pPipe2.y -= P.oeY; pPipe2.z -= 0.0;
  const V3 vXRT = pPipe2 - pExitCB;
  const double factor = (0.0 - pExitCB.z) / vXRT.z;
  const V3 pEnt = pExitCB + factor * vXRT;
  const double radialDist = sqrt(pEnt.x * pEnt.x + pEnt.y * pEnt.y);

  // ---- opaque structures rt:1635-1704
  if (P.telKind == SART_TK_XMM || P.telKind == SART_TK_ABRIXAS) {
    const bool xmm = P.telKind == SART_TK_XMM;
    const double zs = xmm ? -85.0 : -35.0;
    const double factorSpider = (zs - pExitCB.z) / vXRT.z;
    const V3 pSp = pExitCB + factorSpider * vXRT;
    const double phiFlat = acos(pEnt.x / radialDist) / kRadPerDeg;
    const double rSp = sqrt(pSp.x * pSp.x + pSp.y * pSp.y);
    const double phiSp = acos(pSp.x / rSp) / kRadPerDeg;
    bool hit = false;
    if (xmm) {
      if (radialDist <= 64.7) {
        // rt:1674-1688: with the reference's htNone hole the centre is always opaque; other hole types follow
        // lineIntersectsObject (rt:494-527)
        hit = true;
        if (P.holeType != SART_HT_NONE) {
          const int nH = P.numberOfHoles;
          const int half = nH - int(ceil(double(nH) / 2.0));
          const double R = P.holeInOptics;
          for (int l = -half; l <= half; ++l) {
            double cx = 0.0, cy = 0.0;
            if (l != 0) {
              if ((abs(l) & 1) == 0) cy += 2.0 * double(l) * R;
              else cx += 2.0 * (double(l) + (double(l) / double(abs(l)))) * R;
            }
            // lineIntersectsObject(holeType, pointExitCB, pointEntranceXRT, centerHole, holeInOptics) rt:1683
            const V3 hv = pEnt - pExitCB;
            const double hl = (0.0 - pExitCB.z) / hv.z;
            const V3 hp = pExitCB + hl * hv;
            const double ix = hp.x - cx, iy = hp.y - cy;
            const double tx = ix / sqrt(2.0) - iy / sqrt(2.0), ty = ix / sqrt(2.0) + iy / sqrt(2.0);
            bool in = false;
            switch (P.holeType) {
              case SART_HT_CIRCLE: in = sqrt(ix * ix + iy * iy) < R; break;
              case SART_HT_CROSS:
                in = (fabs(ix) < R && fabs(iy) < R * 16.0) || (fabs(iy) < R && fabs(ix) < R * 16.0); break;
              case SART_HT_STAR:
                in = (fabs(ix) < R && fabs(iy) < R * 16.0) || (fabs(iy) < R && fabs(ix) < R * 16.0) ||
                     (fabs(tx) < R && fabs(ty) < R * 16.0) || (fabs(ty) < R && fabs(tx) < R * 16.0); break;
              case SART_HT_SQUARE: in = fabs(ix) < R && fabs(iy) < R; break;
              case SART_HT_DIAMOND: in = fabs(tx) < R && fabs(ty) < R; break;
              default: break;
            }
            if (in) { hit = false; break; } else hit = true;
          }
        }
      } else if (radialDist < 151.6 && radialDist > (151.6 - 20.9)) {
        hit = true;
      } else if (radialDist > 64.7) {
        for (int i = 0; i <= 16; ++i) {
          const double fi = double(i);
          if ((phiFlat >= (-1.145 + 22.5 * fi) && phiFlat <= (1.145 + 22.5 * fi)) ||
              (phiSp >= (-1.145 + 22.5 * fi) && phiSp <= (1.145 + 22.5 * fi))) { hit = true; break; }
        }
      }
    } else {
      if (radialDist < 37.5) hit = true;
      else {
        for (int i = 0; i <= 6; ++i) {
          const double fi = double(i);
          if ((phiFlat >= (-3.75 + 60.0 * fi) && phiFlat <= (3.75 + 60.0 * fi)) ||
              (phiSp >= (-3.75 + 60.0 * fi) && phiSp <= (3.75 + 60.0 * fi))) { hit = true; break; }
        }
      }
    }
    if (hit) { g.code = SART_EXIT_OPAQUE; return; }
  }
  // LLNL: the graphite-block test never vetoes (bare `return`, rt:1646, quirk Q1)

  // ---- which shell rt:1932-1957
  const ShellF64* __restrict__ S = T.shells;
  const int nS = P.nShells;
  if (radialDist > __ldg(&S[nS - 1].R1)) { g.code = SART_EXIT_OUTSIDE_SHELLS; return; }
  int hitLayer = -1;
  double minDist = INFINITY;
  if (P.shellsMonotonic) {
    // the reference's scan over every shell (no break, rt:1937-1950), for radii that increase shell by shell with no glass
    // reaching the next one: the hit shell is the first with R1 > radialDist, and only the shell below it can show its
    // glass front — the same outcome from a binary search and one test
    int lo = 0, hi = nS;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(&S[mid].R1) > radialDist) hi = mid; else lo = mid + 1;
    }
    if (lo > 0 && radialDist > __ldg(&S[lo - 1].R1) && radialDist < __ldg(&S[lo - 1].R1pT)) { g.code = SART_EXIT_GLASS_FRONT; return; }
    if (lo < nS) { hitLayer = lo; minDist = __ldg(&S[lo].R1) - radialDist; }
  } else {
    for (int j = 0; j < nS; ++j) {
      const double R1 = __ldg(&S[j].R1);
      if (radialDist > R1 && radialDist < __ldg(&S[j].R1pT)) { g.code = SART_EXIT_GLASS_FRONT; return; }
      const double dist = R1 - radialDist;
      if (dist > 0.0 && dist < minDist) { minDist = dist; hitLayer = j; }
    }
  }
  if (hitLayer < 0) {
    // radialDist == allR1[last] exactly: the reference goes on with r1 = beta = 0 and falls out at the
    // degenerate-hit test (rt:2055); measure-zero, reported as that exit.
    g.code = SART_EXIT_NO_MIRROR_HIT; g.shell = 0; return;
  }
  g.shell = hitLayer;
  const ShellF64& sh = S[hitLayer];

  // ---- two reflections rt:1983-2037
  V3 pm1, pm2, vAfter2;
  double a1, a2;
  const double lM = P.lMirror;
  if (P.telKind == SART_TK_XMM || P.telKind == SART_TK_ABRIXAS) {
    // findPosParabolic rt:660-690 (point = pExitCB, direc = pEnt - pExitCB)
    const V3 p = pExitCB, d = pEnt - pExitCB;
    {
      const double e = sh.p_e;
      const double a = d.x * d.x + d.y * d.y;
      const double b = 2.0 * (p.x * d.x + p.y * d.y) + e * d.z;
      const double c = p.x * p.x + p.y * p.y - sh.p_r3sq - sh.p_el + e * p.z;
      pm1 = pick_root(p, d, a, b / 2.0, c, 0.0, sh.zmax1);
    }
    V3 n;  // calcNormalVec msParabolic rt:740-746
    {
      const double m = 1.0 / (sh.p_r3tan / sqrt(sh.p_r3sq + sh.p_r3_2tan * (lM - pm1.z)));
      const double nn = sqrt(pm1.x * pm1.x + pm1.y * pm1.y) - m * pm1.z;
      n = {pm1.x, pm1.y, pm1.z - (-nn / m)};
    }
    const V3 vA1 = reflect(n, pExitCB, pEnt, a1);
    const V3 pam1 = pm1 + 200.0 * vA1;
    {  // findPosHyperbolic rt:692-729
      const V3 p2 = pm1, d2 = pam1 - pm1;
      const double e = sh.h_e, gg = sh.h_g;
      const double a = d2.x * d2.x + d2.y * d2.y - gg * d2.z * d2.z;
      const double b = 2.0 * (p2.x * d2.x + p2.y * d2.y + gg * d2.z * lM - gg * d2.z * p2.z) + e * d2.z;
      const double c = p2.x * p2.x + p2.y * p2.y - sh.h_r3sq - sh.h_el + e * p2.z - sh.h_gll + sh.h_2g * p2.z * lM -
                       gg * p2.z * p2.z;
      pm2 = pick_root(p2, d2, a, b / 2.0, c, sh.distanceMirrors, sh.zmax2);
    }
    {  // calcNormalVec msHyperbolic rt:747-758
      const double z = pm2.z;
      const double m = 1.0 / (sh.h_r3tan * (1.0 + 2.0 * (lM - z) / sh.h_nden) /
                              sqrt(sh.h_r3sq + sh.h_r3_2tan * (lM - z) * (1.0 + (lM - z) / sh.h_nden)));
      const double nn = sqrt(pm2.x * pm2.x + pm2.y * pm2.y) - m * z;
      n = {pm2.x, pm2.y, pm2.z - (-nn / m)};
    }
    vAfter2 = reflect(n, pm1, pam1, a2);
  } else {
    // findPosCone rt:628-658, mirror 1: angle beta, radius r1, distMirr 0
    const V3 p = pExitCB, d = pEnt - pExitCB;
    {
      const double k = sh.k1;
      const double a = d.x * d.x + d.y * d.y - k * d.z * d.z;
      const double b = 2.0 * (p.x * d.x + p.y * d.y + sh.r1tan1 * d.z - k * (p.z - 0.0) * d.z);
      const double c = p.x * p.x + p.y * p.y - sh.r1sq + sh.two_r1_tan1 * (p.z - 0.0) - k * (p.z - 0.0) * (p.z - 0.0);
      pm1 = pick_root(p, d, a, b / 2.0, c, 0.0, sh.zmax1);
    }
    V3 n = {pm1.x, pm1.y, sh.tan1 * sqrt(pm1.x * pm1.x + pm1.y * pm1.y)};  // calcNormalVec msCone rt:737-739
    const V3 vA1 = reflect(n, pExitCB, pEnt, a1);
    const V3 pam1 = pm1 + 200.0 * vA1;
    {  // mirror 2: angle 3 beta, radius r4, distMirr = distanceMirrors
      const V3 p2 = pm1, d2 = pam1 - pm1;
      const double k = sh.k2, dm = sh.distanceMirrors;
      const double a = d2.x * d2.x + d2.y * d2.y - k * d2.z * d2.z;
      const double b = 2.0 * (p2.x * d2.x + p2.y * d2.y + sh.r4tan2 * d2.z - k * (p2.z - dm) * d2.z);
      const double c = p2.x * p2.x + p2.y * p2.y - sh.r4sq + sh.two_r4_tan2 * (p2.z - dm) - k * (p2.z - dm) * (p2.z - dm);
      pm2 = pick_root(p2, d2, a, b / 2.0, c, dm, sh.zmax2);
    }
    n = {pm2.x, pm2.y, sh.tan2 * sqrt(pm2.x * pm2.x + pm2.y * pm2.y)};
    vAfter2 = reflect(n, pm1, pam1, a2);
  }
  const V3 pam2 = pm2 + 200.0 * vAfter2;
  const double alpha1 = a1 / kRadPerDeg, alpha2 = a2 / kRadPerDeg;  // Radian -> Degree
  g.alpha1 = alpha1; g.alpha2 = alpha2;

  // ---- nickel of the shell below rt:1706-1734, 2040-2046
  if (hitLayer > 0) {
    const double tana = tan(alpha1 * kRadPerDeg);
    const double compVal = (sh.R1 - __ldg(&S[hitLayer - 1].R1pT)) / (lM - pm1.z);
    if (tana > compVal) { g.code = SART_EXIT_NICKEL; return; }
  }
  // ---- a findPos* that found no root returned its start point rt:2051-2057
  if (almost_equal(pm1.z, pm2.z) || almost_equal(pm1.z, pExitCB.z)) { g.code = SART_EXIT_NO_MIRROR_HIT; return; }

  // ---- detector plane rt:797-814, 2064-2088
  V3 pdw;
  {
    V3 a = {pm2.x * P.cosPipe + pm2.z * P.sinPipe, pm2.y, pm2.z * P.cosPipe - pm2.x * P.sinPipe};
    a.x -= P.dShift;
    V3 b = {pam2.x * P.cosPipe + pam2.z * P.sinPipe, pam2.y, pam2.z * P.cosPipe - pam2.x * P.sinPipe};
    b.x -= P.dShift;
    const V3 w = b - a;
    const double n = (sh.ddWin - a.z) / w.z;
    pdw = a + n * w;
    if (kFull) {
      const double n3 = (sh.ddEnd - a.z) / w.z;
      const V3 pe = a + n3 * w;
      const double dx = pe.x - pdw.x, dy = pe.y - pdw.y;
      g.deviationDet = sqrt(dx * dx + dy * dy);
    }
  }
  // ---- pitch / yaw of the incoming ray rt:2101-2116
  {
    const V3 vb = {-vXRT.x, -vXRT.y, -vXRT.z};
    if (P.reflKind == SART_RK_EFFECTIVE_AREA) g.pitch = acos(vb.x / sqrt(dot(vb, vb))) / kRadPerDeg - 90.0;
    g.ya = atan2(vb.z, vb.y) / kRadPerDeg + 90.0;
    g.cosya = cos(g.ya);  // degrees fed to cos as radians, as the reference does (rt:1598, quirk Q3)
    g.distancePipe = (pdw.z - pExitCB.z) * 1e-3;
  }
  if (P.testXray && minDist > 100.0) {  // rt:2130-2132: unreflected test-source rays go straight on
    const V3 dv = pEnt - pExitCB;
    const double n = (sh.distDet - pExitCB.z) / dv.z;
    pdw = pExitCB + n * dv;
  }
  pdw.x -= P.lateralShift;
  pdw.y -= P.transversalShift;
  g.xw = pdw.x; g.yw = pdw.y;
  // ---- window aperture rt:2139-2147
  if (!(P.flags & SART_CF_IGNORE_DET_WINDOW) && sqrt(pdw.x * pdw.x + pdw.y * pdw.y) > P.radiusWindow) {
    g.windowMiss = 1;
  } else if (fabs(pdw.x) > P.chipCX || fabs(pdw.y) > P.chipCY) {
    g.windowMiss = 1;
  }
  // ---- strongback strips rt:2149-2185 (rotateAroundZ by theta; only y is used)
  {
    const double y = pdw.y * P.cosTheta - pdw.x * P.sinTheta;
    int sb = 2;
    for (int i = 0; i <= P.nStripHalf - 1; ++i) {
      const double fi = double(i);
      if (fabs(y) > (1.0 * fi + 0.5) * P.stripDist + fi * P.stripWidth &&
          fabs(y) < (1.0 * fi + 0.5) * P.stripDist + (fi + 1.0) * P.stripWidth) { sb = 1; break; }
      else sb = 0;
    }
    g.strongback = sb;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Mass-independent weight factors of a ray that reached the weight stage.
struct Weights {
  double reflect;      // R(alpha1,E)*R(alpha2,E)                                rt:1533-1580
  double transWindow;  // strongback or window transmission                      rt:2165-2185
  double absGas;       // detector gas absorption                                rt:2190
  // buffer gas (axionMassforMagnet.nim:75-113), everything that does not depend on m_a
  double gamma, m_gamma, L, term1, absorb;
};

__device__ __forceinline__ double log_mass_attenuation(double e) {  // am:70-73
  return -1.5832 + 5.9195 * exp(-0.353808 * e) + 4.03598 * exp(-0.970557 * e);
}
__device__ __forceinline__ double he_density(double p, double temp) {  // am:4-15
  const double pressure = p * 1e2;
  const double r = pressure * 4.002602 / (8.314 * temp * 1000.0);
  return r / 1000.0;
}

__device__ __forceinline__ void ray_weights(const Params& P, const Tables& T, Geo& g, Weights& w) {
  const double E = g.energy;
  // reflectivity
  if (P.flags & SART_CF_IGNORE_REFLECTION) {
    w.reflect = 1.0;
  } else if (P.reflKind == SART_RK_EFFECTIVE_AREA) {
    const double p = g.pitch, ya = g.ya;
    const double tp = (0.0008 * p * p * p * p + 1e-04 * p * p * p - 0.4489 * p * p - 0.3116 * p + 96.787) / 100.0;
    const double ty = (6.0e-7 * pow(ya, 6.0) - 1.0e-5 * pow(ya, 5.0) - 0.0001 * pow(ya, 4.0) + 0.0034 * pow(ya, 3.0) -
                       0.0292 * pow(ya, 2.0) - 0.1534 * ya + 99.959) / 100.0;
    const double tt = eval_linear1d(T.ttX, T.ttY, T.ttN, E, g.clamped);
    w.reflect = tt * tp * ty;
  } else {
    int coat = 0;
    if (P.reflKind == SART_RK_MULTI_COATING) {
      while (coat < P.nCoatings && P.layers[coat] < g.shell) ++coat;  // layers.lowerBound(hitLayer) rt:1573
      if (coat > P.nCoatings - 1) { coat = P.nCoatings - 1; g.clamped = 1; }
    }
    const double* z = T.reflectivity + size_t(coat) * P.nAngles * P.nReflEnergies;
    const double r1 = eval_bilinear(P, z, g.alpha1, E, g.clamped);
    const double r2 = eval_bilinear(P, z, g.alpha2, E, g.clamped);
    w.reflect = r1 * r2;
  }
  // window + detector gas: the reference evaluates these only past the aperture cut (rt:2139-2147)
  if (g.windowMiss) { w.transWindow = 0.0; w.absGas = 0.0; }
  else if (g.eIdx >= 0 && T.energyFactors) {   // the same values from the per-energy table (device_params.h: EnergyFactors)
    const EnergyFactors* f = T.energyFactors + g.eIdx;
    const int cl = __ldg(&f->clamped);
    if (g.strongback == 1) { w.transWindow = __ldg(&f->strongback); g.clamped |= (cl >> 1) & 1; }
    else if (g.strongback == 0) { w.transWindow = __ldg(&f->window); g.clamped |= cl & 1; }
    else w.transWindow = 0.0;
    w.absGas = __ldg(&f->gas);
    g.clamped |= (cl >> 2) & 1;
  } else {
    if (g.strongback == 1) w.transWindow = eval_linear1d(T.sbX, T.sbY, T.sbN, E, g.clamped);
    else if (g.strongback == 0) w.transWindow = eval_linear1d(T.wdX, T.wdY, T.wdN, E, g.clamped);
    else w.transWindow = 0.0;
    w.absGas = eval_linear1d(T.gaX, T.gaY, T.gaN, E, g.clamped);
  }
  // buffer gas
  if (P.stage == SART_SK_GAS) {
    const double pathm = g.pathCB * 1e-3;
    const double massAtt = exp(log_mass_attenuation(E));
    const double rhoMagnet = he_density(P.pGas, P.tGas);
    w.gamma = 1.97e-7 * 100.0 * rhoMagnet * massAtt;
    {  // effPhotonMass2 am:51-61
      const double vol = pathm * (kPi * pow(P.radiusCB_m, 2.0));
      const double pressure = P.pGas * 1e2;
      const double amountMol = pressure * vol / (8.314 * P.tGas);
      const double ne = 2.0 * 6.022e23 * (amountMol / vol);
      w.m_gamma = sqrt(pow(1.97e-7, 3.0) * 4.0 * kPi * (1.0 / 137.0) * ne / 511e3);
    }
    w.L = pathm / 1.97e-7;
    const double t1 = (P.g_agamma * 1e-9) * (P.B * 1e3 / 1.444) / 2.0;
    w.term1 = t1 * t1;
    const double rhoPipe = he_density(P.pGas, P.roomTemp);
    w.absorb = exp(-massAtt * rhoPipe * g.distancePipe * 100.0) * exp(-massAtt * rhoMagnet * pathm * 100.0);
  } else {
    w.gamma = w.m_gamma = w.L = w.term1 = 0.0; w.absorb = 1.0;
  }
}

// transmissionMagnet for one axion mass (computeMagnetTransmission rt:1582-1625).
__device__ __forceinline__ double magnet_transmission(const Params& P, const Geo& g, const Weights& w, double mAxion) {
  if (P.stage == SART_SK_VACUUM) {
    double prob = 1.0;
    if (!(P.flags & SART_CF_IGNORE_CONV_PROB)) {
      const double L = g.pathCB * 1e-3;
      const double x = (P.g_agamma * 1e-9) * (P.B * P.tesla_to_eV2) * (L * P.m_to_inv_eV) / 2.0;
      prob = x * x;  // conversionProb rt:363-365
    }
    return g.cosya * prob;
  }
  double prob = 1.0;
  if (!(P.flags & SART_CF_IGNORE_CONV_PROB)) {  // axionConversionProb2 am:75-100
    const double q = fabs((w.m_gamma * w.m_gamma - mAxion * mAxion) / (2.0 * (g.energy * 1000.0)));
    const double term2 = 1.0 / (q * q + w.gamma * w.gamma / 4.0);
    const double term3 = 1.0 + exp(-w.gamma * w.L) - 2.0 * exp(-w.gamma * w.L / 2.0) * cos(q * w.L);
    prob = w.term1 * term2 * term3;
  }
  return g.cosya * prob * w.absorb;
}

struct Final { int code; double w, x, y, r, transMagnet; };

// The tail of traceAxion (rt:2120-2221) for one axion mass.
__device__ __forceinline__ void ray_finish(const Params& P, const Geo& g, const Weights& w, double mAxion, Final& f) {
  f.transMagnet = magnet_transmission(P, g, w, mAxion);
  double weight = (P.flags & SART_CF_IGNORE_REFLECTION) ? f.transMagnet : w.reflect * f.transMagnet;
  int flags = (weight != 0) ? SART_FLAG_PASSED_TILL_WINDOW : 0;
  if (g.clamped) flags |= SART_FLAG_INTERP_CLAMPED;
  f.x = 0.0; f.y = 0.0; f.r = 0.0; f.w = 0.0;
  if (g.windowMiss) { f.code = SART_EXIT_WINDOW_APERTURE | flags; return; }
  if (!(P.flags & SART_CF_IGNORE_DET_WINDOW)) weight *= w.transWindow;
  if (!(P.flags & SART_CF_IGNORE_GAS_ABS)) weight *= w.absGas;
  f.r = sqrt(g.xw * g.xw + g.yw * g.yw);
  f.x = -g.xw + P.chipCX;
  f.y = g.yw + P.chipCY;
  if (!(P.flags & SART_CF_XRAY_TEST)) weight *= P.exposureFactor;
  f.w = weight;
  f.code = ((weight != 0) ? SART_EXIT_PASSED : SART_EXIT_ZERO_WEIGHT) | flags;
}

// ------------------------------------------------------------------------------------------------------------
// Sampling block of traceAxion (rt:1754-1764) from the six Philox uniforms of the ray.
// Returns false if an X-ray test-source ray is stopped by its collimator (rt:1800-1801).
__device__ __forceinline__ bool sample_ray_words(const Params& P, const Tables& T, const uint32_t w[6], V3& O, V3& E,
                                                 double& energy, int& clamped, int* eIdx = nullptr) {
  if (eIdx) *eIdx = -1;
  if (!P.testXray) {
    // getRandomPointFromSolarModel rt:425-442
    const double angle1 = 360.0 * u01(w[0]);
    const double angle2 = 180.0 * u01(w[1]);
    int rLo = 0, rHi = P.nRadii;
    if (T.radiusGuide) {   // window of the word's guide bucket (device_params.h: Tables::radiusGuide)
      const uint32_t k = w[2] >> (32 - kRadGuideBits);
      rLo = int(__ldg(T.radiusGuide + k));
      if (k + 1 < uint32_t(kRadGuide)) rHi = min(int(__ldg(T.radiusGuide + k + 1)), P.nRadii);
    }
    const int rIdx = lower_bound(T.fluxRadiusCDF, rLo, rHi, u01(w[2]));
    const double r = (0.0015 + double(rIdx) * 0.0005) * P.radiusSun;
    double s1, c1, s2, c2;
    sincos(angle1 * kRadPerDeg, &s1, &c1);
    sincos(angle2 * kRadPerDeg, &s2, &c2);
    const V3 inSun = {c1 * s2 * r, s1 * s2 * r, c2 * r};
    O = {inSun.x + P.sunX, inSun.y + P.sunY, inSun.z + P.sunZ};
    // getRandomPointOnDisk rt:412-422
    const double rd = P.radiusCB * sqrt(u01(w[3]));
    const double ang = 360.0 * u01(w[4]);
    double sd, cd;
    sincos(ang * kRadPerDeg, &sd, &cd);
    E = {cd * rd + 0.0, sd * rd + 0.0, 0.0 + P.lengthB};
    // getRandomEnergyFromSolarModel rt:444-471: radius index recovered from the rounded emission point
    const V3 back = {O.x - P.sunX, O.y - P.sunY, O.z - P.sunZ};
    const double rr = sqrt(dot(back, back)) / P.radiusSun;
    const double indexRad = (rr - 0.0015) / 0.0005;
    int iRad = (indexRad - 0.5 > floor(indexRad)) ? int(ceil(indexRad)) : int(floor(indexRad));
    if (iRad < 0) { iRad = 0; clamped = 1; }
    if (iRad > P.nRadii - 1) { iRad = P.nRadii - 1; clamped = 1; }
    const double* cdf = T.diffFluxCDFs + size_t(iRad) * P.nEnergies;
    int eLo = 0, eHi = P.nEnergies;
    if (T.energyGuide) {
      const uint32_t k = w[5] >> (32 - kEnGuideBits);
      const uint16_t* gRow = T.energyGuide + size_t(iRad) * kEnGuide;
      eLo = int(__ldg(gRow + k));
      if (k + 1 < uint32_t(kEnGuide)) eHi = min(int(__ldg(gRow + k + 1)), P.nEnergies);
    }
    int idx = lower_bound(cdf, eLo, eHi, u01(w[5]));
    if (idx > P.nEnergies - 1) { idx = P.nEnergies - 1; clamped = 1; }
    const double e = __ldg(T.energies + idx);
    energy = e > 0.03 ? e : 0.03;
    if (eIdx) *eIdx = idx;
    return true;
  }
  // X-ray test source rt:1765-1801
  {
    const double rd = P.srcRadius * sqrt(u01(w[0]));
    double sd, cd;
    sincos((360.0 * u01(w[1])) * kRadPerDeg, &sd, &cd);
    O = {cd * rd + P.srcX, sd * rd + P.srcY, 0.0 + P.srcZ};
  }
  energy = P.srcEnergy;
  if (P.parallelSource) {
    E = {O.x + (0.5 * u01(w[2])) - 0.25, O.y + (0.5 * u01(w[3])) - 0.25, P.lengthB};
  } else {
    const double rd = P.radiusCB * sqrt(u01(w[2]));
    double sd, cd;
    sincos((360.0 * u01(w[3])) * kRadPerDeg, &sd, &cd);
    E = {cd * rd + 0.0, sd * rd + 0.0, 0.0 + P.lengthB};
  }
  const V3 q = plane_point(O, E - O, P.colZ);
  const double qx = q.x - P.srcX, qy = q.y - P.srcY;
  return sqrt(qx * qx + qy * qy) < P.srcRadius;
}
__device__ __forceinline__ bool sample_ray(const Params& P, const Tables& T, uint64_t seed, uint64_t ray, V3& O, V3& E,
                                           double& energy, int& clamped, int* eIdx = nullptr) {
  uint32_t w[6];
  ray_words(seed, ray, w);
  return sample_ray_words(P, T, w, O, E, energy, clamped, eIdx);
}

}  // namespace sart
