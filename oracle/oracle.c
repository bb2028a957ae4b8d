/* oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar f64 CPU restatement of the per-ray pipeline of jovoy/SolarAxionRayTracing
 * (src/raytracer.nim `traceAxion` and its callees, axionMass/axionMassforMagnet.nim).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product library (libsart.so) never does.
 *
 * PARITY STATUS: the reference cannot be compiled here (no Nim toolchain; ten un-vendored nimble
 * packages; three missing input tables — SURVEY.md §8c) and ships no per-ray golden outputs.
 * This restatement follows the Nim source expression by expression (same operand order, same
 * associativity, no FMA contraction: build with -ffp-contract=off), and is pinned against every
 * known-answer value the reference holds for this path (tests/test_oracle_known_answers.py:
 * effPhotonMass2 values of axionMassforMagnet.nim:116-119, window-strip geometry of
 * calculateWindowValues.nim, lengthTelescope vs optics_exit, the on-axis alpha1 = alpha2 = beta
 * relation of TestMirrors.nim:80-121, vacuum conversion probability).
 * Arithmetic that lives in third-party modules whose source is NOT under /root/reference is restated
 * from its documented behaviour and is "parity unpinned" at the last-ulp level:
 *   numericalnim >= 0.6.1  newBilinearSpline / newLinear1D  (grid lookup + blend; raises outside grid
 *                          — here: clamp and flag SART_FLAG_INTERP_CLAMPED)
 *   unchained              toNaturalUnit(T) = 195.353 eV^2, toNaturalUnit(m) = 1/1.97327e-7 eV^-1,
 *                          .to(m), .to(Radian), .to(Degree) (taken as x*1e-3, x*(pi/180), x/(pi/180))
 *   glm                    normalize(v) = v * (1/sqrt(dot(v,v))), length, dot, cross (textbook)
 *   std/random             xoroshiro128+ global state — NOT matched by design (not reproducible under
 *                          Weave); replaced by Philox4x32-10(seed, ray index)
 * "rt:" abbreviates src/raytracer.nim:, "am:" abbreviates axionMass/axionMassforMagnet.nim:.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/sart.h"

#define PI 3.141592653589793
#define RAD_PER_DEG (PI / 180.0) /* std/math RadPerDeg */

typedef struct { double x, y, z; } v3;

static inline v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vscale(double s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }
static inline v3 vmuls(v3 a, double s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline double vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline double vlength(v3 a) { return sqrt(vdot(a, a)); }
static inline v3 vnormalize(v3 a) { return vmuls(a, 1.0 / sqrt(vdot(a, a))); }
static inline v3 vcross(v3 a, v3 b) {
  return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline double degToRad(double d) { return d * RAD_PER_DEG; }
static inline double radToDeg(double r) { return r / RAD_PER_DEG; }
static inline double cot_(double x) { return 1.0 / tan(x); } /* std/math cot */

/* std/algorithm lowerBound: first index with a[i] >= key, n if none. */
static int lower_bound(const double* a, int n, double key) {
  int lo = 0, count = n;
  while (count != 0) {
    int step = count >> 1;
    int pos = lo + step;
    if (a[pos] < key) { lo = pos + 1; count -= step + 1; }
    else count = step;
  }
  return lo;
}
static int lower_bound_i(const int32_t* a, int n, int key) {
  int lo = 0;
  while (lo < n && a[lo] < key) ++lo;
  return lo;
}

/* std/math almostEqual(x, y, unitsInLastPlace = 4). */
static int almost_equal(double x, double y) {
  if (x == y) return 1;
  double diff = fabs(x - y);
  return diff <= DBL_EPSILON * fabs(x + y) * 4.0 || diff < DBL_MIN;
}

/* ---------------------------------------------------------------- Philox4x32-10 */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  philox4x32_10(c, key[0], key[1]);
  memcpy(out, c, sizeof c);
}
/* Six uniforms of global ray `ray`: two Philox blocks, counter = (ray_lo, ray_hi, block, 0), key = seed;
 * u = (word + 0.5) * 2^-32, strictly inside (0,1) and exact in f64. */
void oracle_ray_uniforms(uint64_t seed, uint64_t ray, double u[6]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t a[4] = {(uint32_t)ray, (uint32_t)(ray >> 32), 0u, 0u};
  uint32_t b[4] = {(uint32_t)ray, (uint32_t)(ray >> 32), 1u, 0u};
  philox4x32_10(a, k0, k1);
  philox4x32_10(b, k0, k1);
  const double s = 1.0 / 4294967296.0;
  u[0] = ((double)a[0] + 0.5) * s; u[1] = ((double)a[1] + 0.5) * s;
  u[2] = ((double)a[2] + 0.5) * s; u[3] = ((double)a[3] + 0.5) * s;
  u[4] = ((double)b[0] + 0.5) * s; u[5] = ((double)b[1] + 0.5) * s;
}

/* ---------------------------------------------------------------- rotations rt:334-361 */
static v3 translateZ(v3 v, double d) { v.z += d; return v; }
static v3 rotateInX(v3 v, double angle, double off) {
  v = translateZ(v, -off);
  v3 r = V(v.x * cos(angle) + v.z * sin(angle), v.y, v.z * cos(angle) - v.x * sin(angle));
  return translateZ(r, off);
}
static v3 rotateInY(v3 v, double angle, double off) {
  v = translateZ(v, -off);
  v3 r = V(v.x, v.y * cos(angle) - v.z * sin(angle), v.z * cos(angle) + v.y * sin(angle));
  return translateZ(r, off);
}
static v3 rotateAroundZ(v3 v, double angle) {
  return V(v.x * cos(angle) + v.y * sin(angle), v.y * cos(angle) - v.x * sin(angle), v.z);
}

/* ---------------------------------------------------------------- clipping rt:481-616 */
static int lineIntersectsCircle(v3 p1, v3 p2, v3 center, double radius) {
  v3 vector = vsub(p2, p1);
  double lambda1 = (center.z - p1.z) / vector.z;
  v3 intersect = vsub(vadd(p1, vscale(lambda1, vector)), center);
  double r_xy = sqrt(intersect.x * intersect.x + intersect.y * intersect.y);
  return r_xy < radius;
}
static int lineIntersectsObject(int kind, v3 p1, v3 p2, v3 center, double radius) { /* rt:494-527 */
  v3 vector = vsub(p2, p1);
  double lambda1 = (center.z - p1.z) / vector.z;
  v3 is = vsub(vadd(p1, vscale(lambda1, vector)), center);
  double r_xy = sqrt(is.x * is.x + is.y * is.y);
  double tx = is.x / sqrt(2.0) - is.y / sqrt(2.0);
  double ty = is.x / sqrt(2.0) + is.y / sqrt(2.0);
  switch (kind) {
    case SART_HT_CIRCLE: return r_xy < radius;
    case SART_HT_CROSS:
      return (fabs(is.x) < radius && fabs(is.y) < radius * 16.0) ||
             (fabs(is.y) < radius && fabs(is.x) < radius * 16.0);
    case SART_HT_STAR:
      return (fabs(is.x) < radius && fabs(is.y) < radius * 16.0) ||
             (fabs(is.y) < radius && fabs(is.x) < radius * 16.0) ||
             (fabs(tx) < radius && fabs(ty) < radius * 16.0) ||
             (fabs(ty) < radius && fabs(tx) < radius * 16.0);
    case SART_HT_SQUARE: return fabs(is.x) < radius && fabs(is.y) < radius;
    case SART_HT_DIAMOND: return fabs(tx) < radius && fabs(ty) < radius;
    default: return 0;
  }
}
static v3 getIntersectLineIntersectsCircle(v3 p1, v3 p2, v3 center) {
  v3 vector = vsub(p2, p1);
  double lambda1 = (center.z - p1.z) / vector.z;
  return vadd(p1, vscale(lambda1, vector));
}
typedef struct { v3 v1; int valid1; v3 v2; int valid2; } cyl_result;
static cyl_result lineIntersectsCylinder(v3 point_1, v3 point_2, v3 cb, v3 ce, double radius) {
  double alpha_x = asin((ce.x - cb.x) / fabs(cb.z - ce.z));
  double alpha_y = asin((ce.y - cb.y) / fabs(cb.z - ce.z));
  double offset_x, offset_y;
  if (fabs(ce.x) <= fabs(cb.x)) offset_x = ce.x; else offset_x = cb.x;
  if (fabs(ce.y) <= fabs(cb.y)) offset_y = ce.y; else offset_y = cb.y;
  v3 off = V(offset_x, offset_y, 0.0);
  v3 p_1 = vsub(rotateInY(rotateInX(point_1, alpha_x, 0.0), alpha_y, 0.0), off);
  v3 p_2 = vsub(rotateInY(rotateInX(point_2, alpha_x, 0.0), alpha_y, 0.0), off);
  v3 vector = vsub(p_2, p_1);
  double lambda_dummy = (-1000.0 - p_1.z) / vector.z;
  v3 dummy = vadd(p_1, vscale(lambda_dummy, vector));
  v3 vd = vsub(p_2, dummy);
  double factor = vd.x * vd.x + vd.y * vd.y;
  double p = 2.0 * (dummy.x * vd.x + dummy.y * vd.y) / factor;
  double q = (dummy.x * dummy.x + dummy.y * dummy.y - radius * radius) / factor;
  double lambda_1 = -p / 2.0 + sqrt(p * p / 4.0 - q);
  double lambda_2 = -p / 2.0 - sqrt(p * p / 4.0 - q);
  v3 i1 = vadd(dummy, vscale(lambda_1, vd));
  v3 i2 = vadd(dummy, vscale(lambda_2, vd));
  i1 = vadd(i1, off);
  i1 = rotateInY(rotateInX(i1, -alpha_x, 0.0), -alpha_y, 0.0);
  i2 = vadd(i2, off);
  i2 = rotateInY(rotateInX(i2, -alpha_x, 0.0), -alpha_y, 0.0);
  cyl_result r;
  r.v1 = i1; r.v2 = i2;
  r.valid1 = (i1.z > cb.z) && (i1.z < ce.z);
  r.valid2 = (i2.z > cb.z) && (i2.z < ce.z);
  return r;
}
static int lineIntersectsCylinderOnce(v3 p1, v3 p2, v3 cb, v3 ce, double radius) { /* rt:591-604 */
  cyl_result r = lineIntersectsCylinder(p1, p2, cb, ce, radius);
  if ((r.valid1 && r.valid2) || (!r.valid1 && !r.valid2)) return 0;
  return 1;
}
static v3 getIntersectLineIntersectsCylinderOnce(v3 p1, v3 p2, v3 cb, v3 ce, double radius) {
  cyl_result r = lineIntersectsCylinder(p1, p2, cb, ce, radius);
  return r.valid1 ? r.v1 : r.v2;
}

/* ---------------------------------------------------------------- mirrors rt:628-795 */
enum { MS_CONE = 0, MS_PARABOLIC = 1, MS_HYPERBOLIC = 2 };

static v3 pick_root(v3 point, v3 direc, double a, double half_b, double c, double distMirr,
                    double lMirror, double angle) {
  double s;
  double root1 = (-half_b - sqrt(half_b * half_b - a * c)) / a;
  double root2 = (-half_b + sqrt(half_b * half_b - a * c)) / a;
  if (point.z + root1 * direc.z > distMirr && point.z + root1 * direc.z < distMirr + lMirror * cos(angle))
    s = root1;
  else if (point.z + root2 * direc.z > distMirr && point.z + root2 * direc.z < distMirr + lMirror * cos(angle))
    s = root2;
  else
    s = 0.0;
  return vadd(point, vscale(s, direc));
}
v3 findPosCone(v3 pointXRT, v3 pointCB, double r1, double angle, double lMirror, double distMirr) {
  v3 point = pointCB, direc = vsub(pointXRT, pointCB);
  double k = tan(angle) * tan(angle);
  double a = direc.x * direc.x + direc.y * direc.y - k * direc.z * direc.z;
  double b = 2.0 * (point.x * direc.x + point.y * direc.y + r1 * tan(angle) * direc.z -
                    k * (point.z - distMirr) * direc.z);
  double half_b = b / 2.0;
  double c = point.x * point.x + point.y * point.y - r1 * r1 + 2.0 * r1 * tan(angle) * (point.z - distMirr) -
             k * (point.z - distMirr) * (point.z - distMirr);
  return pick_root(point, direc, a, half_b, c, distMirr, lMirror, angle);
}
v3 findPosParabolic(v3 pointXRT, v3 pointCB, double r1, double angle, double lMirror, double distMirr) {
  v3 point = pointCB, direc = vsub(pointXRT, pointCB);
  double r3 = -tan(angle) * lMirror + sqrt(tan(angle) * lMirror * tan(angle) * lMirror + r1 * r1);
  double e = 2.0 * r3 * tan(angle);
  double a = direc.x * direc.x + direc.y * direc.y;
  double b = 2.0 * (point.x * direc.x + point.y * direc.y) + e * direc.z;
  double half_b = b / 2.0;
  double c = point.x * point.x + point.y * point.y - r3 * r3 - e * lMirror + e * point.z;
  return pick_root(point, direc, a, half_b, c, distMirr, lMirror, angle);
}
v3 findPosHyperbolic(v3 pointXRT, v3 pointCB, double r1, double angle, double lMirror, double distMirr,
                     double focalLength) {
  v3 point = pointCB, direc = vsub(pointXRT, pointCB);
  double r3 = -tan(angle / 3.0) * lMirror +
              sqrt(tan(angle / 3.0) * lMirror * tan(angle / 3.0) * lMirror + r1 * r1);
  double f = focalLength;
  double e = 2.0 * r3 * tan(angle);
  double g = 2.0 * r3 * tan(angle) / (f + r3 * cot_(2.0 * angle / 3.0));
  double a = direc.x * direc.x + direc.y * direc.y - g * direc.z * direc.z;
  double b = 2.0 * (point.x * direc.x + point.y * direc.y + g * direc.z * lMirror - g * direc.z * point.z) +
             e * direc.z;
  double half_b = b / 2.0;
  double c = point.x * point.x + point.y * point.y - r3 * r3 - e * lMirror + e * point.z -
             g * lMirror * lMirror + 2.0 * g * point.z * lMirror - g * point.z * point.z;
  return pick_root(point, direc, a, half_b, c, distMirr, lMirror, angle);
}
static v3 calcNormalVec(v3 pm, double angle, double r1, double lMirror, double focalLength, int shape) {
  v3 n = V(pm.x, pm.y, 0.0);
  if (shape == MS_CONE) {
    n.z = tan(angle) * sqrt(pm.x * pm.x + pm.y * pm.y);
  } else if (shape == MS_PARABOLIC) {
    double r3 = -tan(angle) * lMirror + sqrt(tan(angle) * lMirror * tan(angle) * lMirror + r1 * r1);
    double m = 1.0 / (r3 * tan(angle) / sqrt(r3 * r3 + r3 * 2.0 * tan(angle) * (lMirror - pm.z)));
    double nn = sqrt(pm.x * pm.x + pm.y * pm.y) - m * pm.z;
    n.z = pm.z - (-nn / m);
  } else {
    double r3 = -tan(angle / 3.0) * lMirror +
                sqrt(tan(angle / 3.0) * lMirror * tan(angle / 3.0) * lMirror + r1 * r1);
    double f = focalLength;
    double al = angle / 3.0;
    double z = pm.z;
    double m = 1.0 / (r3 * tan(angle) * (1.0 + 2.0 * (lMirror - z) / (f + r3 * cot_(2.0 * al))) /
                      sqrt(r3 * r3 + r3 * 2.0 * tan(angle) * (lMirror - z) *
                                         (1.0 + (lMirror - z) / (f + r3 * cot_(2.0 * al)))));
    double nn = sqrt(pm.x * pm.x + pm.y * pm.y) - m * z;
    n.z = pm.z - (-nn / m);
  }
  return n;
}
v3 getVectoraAfterMirror(v3 pointXRT, v3 pointCB, v3 pointMirror, double angle, double r1, double lMirror,
                         double focalLength, int shape) {
  v3 normalVec = calcNormalVec(pointMirror, angle, r1, lMirror, focalLength, shape);
  v3 vbm = vnormalize(vsub(pointXRT, pointCB));
  v3 axis = vnormalize(vcross(normalVec, vbm));
  double alphaMirror = asin(fabs(vdot(normalVec, vbm) / vlength(normalVec)));
  v3 vecBeforeAxis = vcross(vbm, axis);
  return vsub(vmuls(vbm, cos(2.0 * alphaMirror)), vmuls(vecBeforeAxis, sin(2.0 * alphaMirror)));
}
double getMirrorAngle(v3 pointXRT, v3 pointCB, v3 pointMirror, double angle, double r1, double lMirror,
                      double focalLength, int shape) { /* returns degree */
  v3 normalVec = calcNormalVec(pointMirror, angle, r1, lMirror, focalLength, shape);
  v3 vbm = vnormalize(vsub(pointXRT, pointCB));
  double alphaMirror = asin(fabs(vdot(normalVec, vbm) / vlength(normalVec)));
  return radToDeg(alphaMirror);
}

/* ---------------------------------------------------------------- detector plane rt:797-814 */
static v3 getPointDetectorWindow(v3 pm2, v3 pam2, double distDet, double dCBXray, double pipeAngleDeg) {
  double pipeRad = degToRad(pipeAngleDeg);
  v3 shift = V(dCBXray, 0.0, 0.0);
  v3 a = vsub(rotateInX(pm2, pipeRad, 0.0), shift);
  v3 b = vsub(rotateInX(pam2, pipeRad, 0.0), shift);
  v3 vam2 = vsub(b, a);
  double dd = distDet / cos(pipeRad);
  double n = (dd - a.z) / vam2.z;
  return vadd(a, vscale(n, vam2));
}

/* ---------------------------------------------------------------- interpolation (numericalnim) */
static double eval_linear1d(const sart_interp1d_t* t, double x, int* clamped) {
  int n = t->n;
  if (n < 2) { *clamped = 1; return n == 1 ? t->y[0] : 0.0; }
  if (!(x >= t->x[0])) { *clamped = 1; x = t->x[0]; }
  if (!(x <= t->x[n - 1])) { *clamped = 1; x = t->x[n - 1]; }
  /* interval i with x[i] <= x <= x[i+1] */
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (t->x[mid] <= x) lo = mid; else hi = mid;
  }
  double slope = (t->y[lo + 1] - t->y[lo]) / (t->x[lo + 1] - t->x[lo]);
  return t->y[lo] + (x - t->x[lo]) * slope;
}
static double eval_bilinear(const sart_tables_t* tb, int coat, double x, double y, int* clamped) {
  int nx = tb->nAngles, ny = tb->nReflEnergies;
  const double* z = tb->reflectivity + (size_t)coat * nx * ny;
  if (!(x >= tb->angleMin)) { *clamped = 1; x = tb->angleMin; }
  if (!(x <= tb->angleMax)) { *clamped = 1; x = tb->angleMax; }
  if (!(y >= tb->reflEnergyMin)) { *clamped = 1; y = tb->reflEnergyMin; }
  if (!(y <= tb->reflEnergyMax)) { *clamped = 1; y = tb->reflEnergyMax; }
  double dx = (tb->angleMax - tb->angleMin) / (double)(nx - 1);
  double dy = (tb->reflEnergyMax - tb->reflEnergyMin) / (double)(ny - 1);
  int i = (int)floor((x - tb->angleMin) / dx);
  int j = (int)floor((y - tb->reflEnergyMin) / dy);
  if (i > nx - 2) i = nx - 2;
  if (j > ny - 2) j = ny - 2;
  double xc = x - (tb->angleMin + dx * (double)i);
  double yc = y - (tb->reflEnergyMin + dy * (double)j);
  double z00 = z[(size_t)i * ny + j], z10 = z[(size_t)(i + 1) * ny + j];
  double z01 = z[(size_t)i * ny + j + 1], z11 = z[(size_t)(i + 1) * ny + j + 1];
  double alpha = z00;
  double beta = (z10 - z00) / dx;
  double gamma = (z01 - z00) / dy;
  double delta = (z11 + z00 - z10 - z01) / (dx * dy);
  return alpha + beta * xc + gamma * yc + delta * xc * yc;
}

/* ---------------------------------------------------------------- buffer gas am:4-113 */
double oracle_density(double p, double temp) { /* am:4-15 */
  const double gasConstant = 8.314, M = 4.002602;
  double pressure = p * 1e2;
  double result = pressure * M / (gasConstant * temp * 1000.0);
  return result / 1000.0;
}
static double numDensity(double c) { return 2.0 * 6.022e23 * c; } /* am:30-33 */
double oracle_effPhotonMass(double ne) { /* am:35-40 */
  const double alpha = 1.0 / 137.0, me = 511e3;
  return sqrt(pow(1.97e-7, 3.0) * 4.0 * PI * alpha * ne / me);
}
static double molarAmount(double p, double vol, double temp) { /* am:42-48 */
  const double gasConstant = 8.314;
  double pressure = p * 1e2;
  return pressure * vol / (gasConstant * temp);
}
double oracle_effPhotonMass2(double p, double length, double radBore, double temp) { /* am:51-61 */
  double vol = length * (PI * pow(radBore, 2.0));
  double amountMol = molarAmount(p, vol, temp);
  double numPerMol = numDensity(amountMol / vol);
  return oracle_effPhotonMass(numPerMol);
}
static double momentumTransfer(double m_gamma, double m_a, double E_keV) { /* am:63-68 */
  return fabs((m_gamma * m_gamma - m_a * m_a) / (2.0 * (E_keV * 1000.0)));
}
static double logMassAttenuation(double e) { /* am:70-73 */
  return -1.5832 + 5.9195 * exp(-0.353808 * e) + 4.03598 * exp(-0.970557 * e);
}
double oracle_axionConversionProb2(double m_a, double energyAx, double pressure, double temp, double length,
                                   double radBore, double g_agamma, double B) { /* am:75-100 */
  double gamma = 1.97e-7 * 100.0 * oracle_density(pressure, temp) * exp(logMassAttenuation(energyAx));
  double m_gamma = oracle_effPhotonMass2(pressure, length, radBore, temp);
  double L = length / 1.97e-7;
  double g_agammaEV = g_agamma * 1e-9;
  double beV = B * 1e3 / 1.444;
  double q = momentumTransfer(m_gamma, m_a, energyAx);
  double t1 = g_agammaEV * beV / 2.0;
  double term1 = t1 * t1;
  double term2 = 1.0 / (q * q + gamma * gamma / 4.0);
  double term3 = 1.0 + exp(-gamma * L) - 2.0 * exp(-gamma * L / 2.0) * cos(q * L);
  return term1 * term2 * term3;
}
double oracle_intensitySuppression2(double energy, double distanceMagnet, double distancePipe, double pressure,
                                    double tempMagnet, double tempPipe) { /* am:102-113 */
  double massAtt = exp(logMassAttenuation(energy));
  double rhoMagnet = oracle_density(pressure, tempMagnet);
  double rhoPipe = oracle_density(pressure, tempPipe);
  return exp(-massAtt * rhoPipe * distancePipe * 100.0) * exp(-massAtt * rhoMagnet * distanceMagnet * 100.0);
}
/* conversionProb rt:363-365 (unchained natural units; constants come from the setup). */
double oracle_conversionProb(const sart_setup_t* s, double B, double g, double length_mm) {
  double L = length_mm * 1e-3;
  double x = (g * 1e-9) * (B * s->consts.tesla_to_eV2) * (L * s->consts.m_to_inv_eV) / 2.0;
  return x * x;
}

/* ---------------------------------------------------------------- weights rt:1533-1625 */
static double computeMagnetTransmission(const sart_setup_t* s, double mAxion, double energy, double distancePipe_m,
                                        double pathCB_mm, double ya) {
  if (s->stage == SART_SK_VACUUM) {
    double prob = (s->flags & SART_CF_IGNORE_CONV_PROB) ? 1.0
                  : oracle_conversionProb(s, s->magnet.B, s->consts.g_agamma, pathCB_mm);
    return cos(ya) * prob;
  } else {
    double pGas = s->magnet.pGasRoom / s->consts.roomTemp * s->magnet.tGas;
    double pathm = pathCB_mm * 1e-3, radm = s->magnet.radiusCB * 1e-3;
    double prob = (s->flags & SART_CF_IGNORE_CONV_PROB) ? 1.0
                  : oracle_axionConversionProb2(mAxion, energy, pGas, s->magnet.tGas, pathm, radm,
                                                s->consts.g_agamma, s->magnet.B);
    double absorb = oracle_intensitySuppression2(energy, pathm, distancePipe_m, pGas, s->magnet.tGas,
                                                 s->consts.roomTemp);
    return cos(ya) * prob * absorb;
  }
}
static void computeReflectivity(const sart_setup_t* s, const sart_tables_t* tb, double energy, int hitLayer,
                                double transmissionMagnet, double p, double ya, double alpha1, double alpha2,
                                double* reflect, double* weight, int* clamped) {
  if (s->flags & SART_CF_IGNORE_REFLECTION) { *reflect = 1.0; *weight = transmissionMagnet; return; }
  switch (s->telescope.reflKind) {
    case SART_RK_EFFECTIVE_AREA: {
      double tp = (0.0008 * p * p * p * p + 1e-04 * p * p * p - 0.4489 * p * p - 0.3116 * p + 96.787) / 100.0;
      double ty = (6.0e-7 * pow(ya, 6.0) - 1.0e-5 * pow(ya, 5.0) - 0.0001 * pow(ya, 4.0) +
                   0.0034 * pow(ya, 3.0) - 0.0292 * pow(ya, 2.0) - 0.1534 * ya + 99.959) / 100.0;
      double tt = eval_linear1d(&tb->telescopeTransmission, energy, clamped);
      *reflect = tt * tp * ty;
      *weight = *reflect * transmissionMagnet;
    } break;
    case SART_RK_SINGLE_COATING: {
      double r1 = eval_bilinear(tb, 0, alpha1, energy, clamped);
      double r2 = eval_bilinear(tb, 0, alpha2, energy, clamped);
      *reflect = r1 * r2;
      *weight = *reflect * transmissionMagnet;
    } break;
    default: {
      int layerIdx = lower_bound_i(s->telescope.layers, s->telescope.nCoatings, hitLayer);
      if (layerIdx > s->telescope.nCoatings - 1) { layerIdx = s->telescope.nCoatings - 1; *clamped = 1; }
      double r1 = eval_bilinear(tb, layerIdx, alpha1, energy, clamped);
      double r2 = eval_bilinear(tb, layerIdx, alpha2, energy, clamped);
      *reflect = r1 * r2;
      *weight = *reflect * transmissionMagnet;
    }
  }
}

/* ---------------------------------------------------------------- opaque structures rt:1627-1734 */
static void radiusAndPhi(v3 v, double* radius, double* phi) {
  *radius = sqrt(v.x * v.x + v.y * v.y);
  *phi = radToDeg(acos(v.x / *radius));
}
static int lineIntersectsOpaqueTelescopeStructures(const sart_setup_t* s, double radialDist, v3 testVector,
                                                   v3 vectorXRT, v3 pointExitCB, v3 pointEntranceXRT) {
  int result = 0;
  const sart_telescope_t* tel = &s->telescope;
  switch (tel->kind) {
    case SART_TK_LLNL:
      /* rt:1646 bare `return` ⇒ result stays false either way (quirk Q1) */
      return 0;
    case SART_TK_ABRIXAS: {
      double factorSpider = (-35.0 - pointExitCB.z) / vectorXRT.z;
      v3 pes = vadd(pointExitCB, vscale(factorSpider, vectorXRT));
      double radius, phiFlat, radiusSpider, phiFlatSpider;
      radiusAndPhi(pointEntranceXRT, &radius, &phiFlat);
      radiusAndPhi(pes, &radiusSpider, &phiFlatSpider);
      if (radialDist < 37.5) result = 1;
      else {
        for (int i = 0; i <= 6; ++i) {
          double fi = (double)i;
          if ((phiFlat >= (-3.75 + 60.0 * fi) && phiFlat <= (3.75 + 60.0 * fi)) ||
              (phiFlatSpider >= (-3.75 + 60.0 * fi) && phiFlatSpider <= (3.75 + 60.0 * fi))) {
            result = 1; break;
          }
        }
      }
    } break;
    case SART_TK_XMM: {
      double factorSpider = (-85.0 - pointExitCB.z) / vectorXRT.z;
      v3 pes = vadd(pointExitCB, vscale(factorSpider, vectorXRT));
      double radius, phiFlat, radiusSpider, phiFlatSpider;
      radiusAndPhi(pointEntranceXRT, &radius, &phiFlat);
      radiusAndPhi(pes, &radiusSpider, &phiFlatSpider);
      if (radialDist <= 64.7) {
        int nHoles = tel->numberOfHoles;
        int half = nHoles - (int)ceil((double)nHoles / 2.0);
        for (int l = -half; l <= half; ++l) {
          v3 centerHole = testVector;
          if (l != 0) {
            if (abs(l) % 2 == 0) centerHole.y += 2.0 * (double)l * tel->holeInOptics;
            else centerHole.x += 2.0 * ((double)l + ((double)l / (double)abs(l))) * tel->holeInOptics;
          }
          if (lineIntersectsObject(tel->holeType, pointExitCB, pointEntranceXRT, centerHole, tel->holeInOptics)) {
            result = 0; break;
          } else result = 1;
        }
      } else if (radialDist < 151.6 && radialDist > (151.6 - 20.9)) {
        result = 1;
      } else if (radialDist > 64.7) {
        for (int i = 0; i <= 16; ++i) {
          double fi = (double)i;
          if ((phiFlat >= (-1.145 + 22.5 * fi) && phiFlat <= (1.145 + 22.5 * fi)) ||
              (phiFlatSpider >= (-1.145 + 22.5 * fi) && phiFlatSpider <= (1.145 + 22.5 * fi))) {
            result = 1; break;
          }
        }
      }
    } break;
    default: break; /* reference: doAssert false; sart_create rejects these kinds */
  }
  return result;
}
static int lineHitsNickel(const sart_setup_t* s, double alpha1deg, double r1, int hitLayer, v3 pointMirror1) {
  const sart_telescope_t* tel = &s->telescope;
  if (hitLayer > 0) {
    double tana = tan(degToRad(alpha1deg));
    int hL = hitLayer - 1;
    double compVal = (r1 - (tel->allR1[hL] + tel->allThickness[hL])) / (tel->lMirror - pointMirror1.z);
    return tana > compVal;
  }
  return 0;
}

/* ---------------------------------------------------------------- the Axion record rt:192-221 */
typedef struct {
  int passed, passedTillWindow, hitNickel, code, clamped;
  double pointdataX, pointdataY, pointdataR, weights, transmissionMagnet, yawAngles, energiesAx, energiesPre;
  double transProbArgon, deviationDet, reflect, alpha1, alpha2, pathCB;
  int shellNumber;
} axion_t;

/* CenterVectors rt:278-320 */
typedef struct { v3 entranceCB, exitCB, exitPipeCBVT3, exitPipeVT3XRT, exitCBMagneticField, sun, xraySource, collimator; } centers_t;
static centers_t initCenterVectors(const sart_setup_t* s) {
  centers_t c;
  c.sun = V(0.0, -(0.0 * 1.33e10), -s->consts.distanceSunEarth);
  c.entranceCB = V(0.0, -0.0, 0.0);
  c.exitCBMagneticField = V(0.0, 0.0, s->magnet.lengthB);
  c.exitCB = V(0.0, -0.0, s->magnet.lengthColdbore);
  c.exitPipeCBVT3 = V(0.0, 0.0, s->magnet.lengthColdbore + s->pipes.cb2vt3_length);
  c.exitPipeVT3XRT = V(0.0, 0.0, s->magnet.lengthColdbore + s->pipes.cb2vt3_length + s->pipes.vt3xrt_length);
  c.xraySource = V(s->testSource.offAxisLeft, s->testSource.offAxisUp, -s->testSource.distance);
  c.collimator = V(s->testSource.offAxisLeft, s->testSource.offAxisUp, -s->testSource.distance + s->testSource.lengthCol);
  return c;
}

/* ---------------------------------------------------------------- traceAxion rt:1811-2221
 * (everything after the sampling block; sampling is in sample_solar below). */
static void trace_after_sampling(axion_t* res, const sart_setup_t* s, const sart_tables_t* tb,
                                 const centers_t* cv, double mAxion, v3 rayOrigin, v3 pointExitCBMagneticField,
                                 double energyAx, int testXray) {
  const sart_telescope_t* tel = &s->telescope;
  memset(res, 0, sizeof *res);
  res->shellNumber = -1;
  res->energiesPre = energyAx; /* rt:1819 */

  int intersectsEntranceCB = lineIntersectsCircle(rayOrigin, pointExitCBMagneticField, cv->entranceCB, s->magnet.radiusCB);
  int intersectsCB = 0;
  if (!intersectsEntranceCB)
    intersectsCB = lineIntersectsCylinderOnce(rayOrigin, pointExitCBMagneticField, cv->entranceCB, cv->exitCB, s->magnet.radiusCB);
  if (!intersectsEntranceCB && !intersectsCB) { res->code = SART_EXIT_MISSED_BORE; return; }

  v3 intersect;
  if (!intersectsEntranceCB)
    intersect = getIntersectLineIntersectsCylinderOnce(rayOrigin, pointExitCBMagneticField, cv->entranceCB, cv->exitCB, s->magnet.radiusCB);
  else
    intersect = getIntersectLineIntersectsCircle(rayOrigin, pointExitCBMagneticField, cv->entranceCB);

  double pathCB = vlength(vsub(pointExitCBMagneticField, intersect)); /* rt:1843 */
  res->pathCB = pathCB;

  if (!lineIntersectsCircle(rayOrigin, pointExitCBMagneticField, cv->exitCB, s->magnet.radiusCB)) {
    res->code = SART_EXIT_CLIP_EXIT_CB; return;
  }
  v3 d0 = vsub(pointExitCBMagneticField, rayOrigin);
  v3 pointExitCB = vadd(rayOrigin, vscale((cv->exitCB.z - rayOrigin.z) / d0.z, d0)); /* rt:1850-1853 */

  if (!lineIntersectsCircle(pointExitCBMagneticField, pointExitCB, cv->exitPipeCBVT3, s->pipes.cb2vt3_radius)) {
    res->code = SART_EXIT_CLIP_PIPE_VT3; return;
  }
  v3 d1 = vsub(pointExitCB, pointExitCBMagneticField);
  v3 pointExitPipeCBVT3 = vadd(pointExitCBMagneticField,
                               vscale((cv->exitPipeCBVT3.z - pointExitCBMagneticField.z) / d1.z, d1)); /* rt:1860-1863 */

  /* rt:1866-1868: uses coldBoreToVT3.radius again (quirk Q2) */
  if (!lineIntersectsCircle(pointExitCB, pointExitPipeCBVT3, cv->exitPipeVT3XRT, s->pipes.cb2vt3_radius)) {
    res->code = SART_EXIT_CLIP_PIPE_XRT; return;
  }
  v3 d2 = vsub(pointExitPipeCBVT3, pointExitCB);
  v3 pointExitPipeVT3XRT = vadd(pointExitCB, vscale((cv->exitPipeVT3XRT.z - pointExitCB.z) / d2.z, d2)); /* rt:1870-1872 */

  /* telescope frame rt:1878-1905 */
  double turnedX = degToRad(tel->telescope_turned_x);
  double turnedY = degToRad(tel->telescope_turned_y);
  double lengthTelescope = (tel->lMirror + 0.5 * tel->allXsep[0]) * cos(degToRad(tel->allAngles[0])) +
                           (tel->lMirror + 0.5 * tel->allXsep[0]) * cos(3.0 * degToRad(tel->allAngles[0]));
  v3 oe = V(tel->optics_entrance[0], tel->optics_entrance[1], 0.0);
  pointExitCB.z -= cv->exitPipeVT3XRT.z;
  pointExitCB = vsub(rotateInY(rotateInX(pointExitCB, turnedX, lengthTelescope / 2.0), turnedY, lengthTelescope / 2.0), oe);
  pointExitPipeVT3XRT.z -= cv->exitPipeVT3XRT.z;
  pointExitPipeVT3XRT = vsub(rotateInY(rotateInX(pointExitPipeVT3XRT, turnedX, lengthTelescope / 2.0), turnedY, lengthTelescope / 2.0), oe);
  v3 vectorXRT = vsub(pointExitPipeVT3XRT, pointExitCB);
  double factor = (0.0 - pointExitCB.z) / vectorXRT.z;
  v3 pointEntranceXRT = vadd(pointExitCB, vscale(factor, vectorXRT));
  v3 vectorBeforeXRT = vectorXRT;
  double radialDist, phi_unused;
  radiusAndPhi(pointEntranceXRT, &radialDist, &phi_unused);

  if (lineIntersectsOpaqueTelescopeStructures(s, radialDist, V(0.0, 0.0, 0.0), vectorXRT, pointExitCB, pointEntranceXRT)) {
    res->code = SART_EXIT_OPAQUE; return;
  }

  /* shell search rt:1918-1957 */
  double minDist = INFINITY, r1 = 0.0, r2 = 0.0, r3 = 0.0, r4 = 0.0, r5 = 0.0, beta = 0.0, xSep = 0.0;
  int hitLayer = 0;
  int nS = tel->nShells;
  if (radialDist > tel->allR1[nS - 1]) { res->code = SART_EXIT_OUTSIDE_SHELLS; return; }
  for (int j = 0; j < nS; ++j) {
    if (radialDist > tel->allR1[j] && radialDist < (tel->allR1[j] + tel->allThickness[j])) {
      res->code = SART_EXIT_GLASS_FRONT; return;
    }
    double dist = tel->allR1[j] - radialDist;
    if (dist > 0.0 && dist < minDist) {
      minDist = dist;
      hitLayer = j;
      r1 = tel->allR1[j];
      beta = degToRad(tel->allAngles[j]);
      xSep = tel->allXsep[j];
      r2 = r1 - tel->lMirror * sin(beta);
      r3 = r2 - 0.5 * xSep * tan(beta);
      r4 = r3 - 0.5 * xSep * tan(3.0 * beta);
      r5 = r4 - tel->lMirror * sin(3.0 * beta);
    }
  }
  (void)r5;
  double beta3 = 3.0 * beta;
  double distanceMirrors = cos(beta) * (xSep + tel->lMirror);
  double fL = s->detectorInstall.distanceDetectorXRT;
  v3 pointMirror1, vectorAfterMirror1, pointAfterMirror1, pointMirror2, vectorAfterMirrors, pointAfterMirror2;
  double alpha1, alpha2;
  if (tel->kind == SART_TK_XMM || tel->kind == SART_TK_ABRIXAS) { /* rt:1984-2010 */
    pointMirror1 = findPosParabolic(pointEntranceXRT, pointExitCB, r1, beta, tel->lMirror, 0.0);
    vectorAfterMirror1 = getVectoraAfterMirror(pointEntranceXRT, pointExitCB, pointMirror1, beta, r1, tel->lMirror, fL, MS_PARABOLIC);
    pointAfterMirror1 = vadd(pointMirror1, vscale(200.0, vectorAfterMirror1));
    pointMirror2 = findPosHyperbolic(pointAfterMirror1, pointMirror1, r1, beta3, tel->lMirror, distanceMirrors, fL);
    vectorAfterMirrors = getVectoraAfterMirror(pointAfterMirror1, pointMirror1, pointMirror2, beta3, r1, tel->lMirror, fL, MS_HYPERBOLIC);
    pointAfterMirror2 = vadd(pointMirror2, vscale(200.0, vectorAfterMirrors));
    alpha1 = getMirrorAngle(pointEntranceXRT, pointExitCB, pointMirror1, beta, r1, tel->lMirror, fL, MS_PARABOLIC);
    alpha2 = getMirrorAngle(pointAfterMirror1, pointMirror1, pointMirror2, beta3, r1, tel->lMirror, fL, MS_HYPERBOLIC);
  } else { /* rt:2011-2037 */
    pointMirror1 = findPosCone(pointEntranceXRT, pointExitCB, r1, beta, tel->lMirror, 0.0);
    vectorAfterMirror1 = getVectoraAfterMirror(pointEntranceXRT, pointExitCB, pointMirror1, beta, r1, tel->lMirror, fL, MS_CONE);
    pointAfterMirror1 = vadd(pointMirror1, vscale(200.0, vectorAfterMirror1));
    pointMirror2 = findPosCone(pointAfterMirror1, pointMirror1, r4, beta3, tel->lMirror, distanceMirrors);
    vectorAfterMirrors = getVectoraAfterMirror(pointAfterMirror1, pointMirror1, pointMirror2, beta3, r1, tel->lMirror, fL, MS_CONE);
    pointAfterMirror2 = vadd(pointMirror2, vscale(200.0, vectorAfterMirrors));
    alpha1 = getMirrorAngle(pointEntranceXRT, pointExitCB, pointMirror1, beta, r1, tel->lMirror, fL, MS_CONE);
    alpha2 = getMirrorAngle(pointAfterMirror1, pointMirror1, pointMirror2, beta3, r1, tel->lMirror, fL, MS_CONE);
  }
  res->alpha1 = alpha1; res->alpha2 = alpha2;

  res->hitNickel = lineHitsNickel(s, alpha1, r1, hitLayer, pointMirror1); /* rt:2040-2046 */
  if (res->hitNickel) { res->code = SART_EXIT_NICKEL; return; }

  double z0 = pointExitCB.z, z1 = pointMirror1.z, z2 = pointMirror2.z; /* rt:2051-2057 */
  if (almost_equal(z1, z2) || almost_equal(z1, z0)) { res->code = SART_EXIT_NO_MIRROR_HIT; return; }

  /* detector plane rt:2064-2094 */
  double distDet = distanceMirrors - 0.5 * tel->allXsep[8] * cos(beta) + s->detectorInstall.distanceDetectorXRT -
                   s->detectorInstall.distanceWindowFocalPlane;
  double d = -tel->optics_entrance[0];
  v3 pointDetectorWindow = getPointDetectorWindow(pointMirror2, pointAfterMirror2, distDet, d, s->pipes.pipesTurned);
  v3 pointEndDetector = getPointDetectorWindow(pointMirror2, pointAfterMirror2, distDet + s->detector.depthDet, d, s->pipes.pipesTurned);
  {
    double ddx = pointEndDetector.x - pointDetectorWindow.x, ddy = pointEndDetector.y - pointDetectorWindow.y;
    res->deviationDet = sqrt(ddx * ddx + ddy * ddy);
  }

  /* angles + conversion probability rt:2101-2123 */
  vectorBeforeXRT = vneg(vectorBeforeXRT);
  double vecLength = vlength(vectorBeforeXRT);
  double p = radToDeg(acos(vectorBeforeXRT.x / vecLength)) - 90.0;
  double ya = radToDeg(atan2(vectorBeforeXRT.z, vectorBeforeXRT.y)) + 90.0;
  double distancePipe = (pointDetectorWindow.z - pointExitCB.z) * 1e-3;
  res->transmissionMagnet = computeMagnetTransmission(s, mAxion, energyAx, distancePipe, pathCB, ya);
  res->yawAngles = ya;

  double weight = 1.0;
  computeReflectivity(s, tb, energyAx, hitLayer, res->transmissionMagnet, p, ya, alpha1, alpha2, &res->reflect, &weight, &res->clamped);

  if (testXray && minDist > 100.0) { /* rt:2130-2132 */
    v3 dv = vsub(pointEntranceXRT, pointExitCB);
    double n = (distDet - pointExitCB.z) / dv.z;
    pointDetectorWindow = vadd(pointExitCB, vscale(n, dv));
  }
  pointDetectorWindow.x -= s->detectorInstall.lateralShift;
  pointDetectorWindow.y -= s->detectorInstall.transversalShift;
  if (weight != 0) res->passedTillWindow = 1;

  /* window aperture rt:2139-2147 */
  double chipCX = s->consts.chipXMax / 2.0, chipCY = s->consts.chipYMax / 2.0;
  if (!(s->flags & SART_CF_IGNORE_DET_WINDOW) &&
      sqrt(pointDetectorWindow.x * pointDetectorWindow.x + pointDetectorWindow.y * pointDetectorWindow.y) > s->detector.radiusWindow) {
    res->code = SART_EXIT_WINDOW_APERTURE; return;
  } else {
    if (fabs(pointDetectorWindow.x) > chipCX || fabs(pointDetectorWindow.y) > chipCY) {
      res->code = SART_EXIT_WINDOW_APERTURE; return;
    }
  }

  v3 turned = rotateAroundZ(pointDetectorWindow, s->detector.theta);
  double y = turned.y;
  double stripDist = s->detector.stripDistWindow, stripWidth = s->detector.stripWidthWindow;
  double transWindow = 0.0;
  int nHalf = (int)round((double)s->detector.numberOfStrips / 2.0);
  for (int i = 0; i <= nHalf - 1; ++i) { /* rt:2167-2185 */
    double fi = (double)i;
    if (fabs(y) > (1.0 * fi + 0.5) * stripDist + fi * stripWidth &&
        fabs(y) < (1.0 * fi + 0.5) * stripDist + (fi + 1.0) * stripWidth) {
      transWindow = eval_linear1d(&tb->strongbackTransmission, energyAx, &res->clamped);
      break;
    } else {
      transWindow = eval_linear1d(&tb->windowTransmission, energyAx, &res->clamped);
    }
  }
  if (!(s->flags & SART_CF_IGNORE_DET_WINDOW)) weight *= transWindow;

  double absGasDet = eval_linear1d(&tb->gasAbsorption, energyAx, &res->clamped);
  if (!(s->flags & SART_CF_IGNORE_GAS_ABS)) weight *= absGasDet;
  res->transProbArgon = absGasDet;
  res->energiesAx = energyAx;
  res->shellNumber = hitLayer;

  res->pointdataR = sqrt(pointDetectorWindow.x * pointDetectorWindow.x + pointDetectorWindow.y * pointDetectorWindow.y);
  pointDetectorWindow.x = -pointDetectorWindow.x + chipCX;
  pointDetectorWindow.y = pointDetectorWindow.y + chipCY;

  if (!(s->flags & SART_CF_XRAY_TEST)) weight *= s->consts.exposureFactor; /* rt:2207-2212 */
  res->pointdataX = pointDetectorWindow.x;
  res->pointdataY = pointDetectorWindow.y;
  res->weights = weight;
  if (weight != 0) { res->passed = 1; res->code = SART_EXIT_PASSED; }
  else res->code = SART_EXIT_ZERO_WEIGHT;
}

/* ---------------------------------------------------------------- sampling rt:412-471, 1754-1764 */
static v3 getRandomPointOnDisk(v3 center, double radius, double ua, double ub) {
  double r = radius * sqrt(ua);
  double angle = 360.0 * ub;
  double x = cos(degToRad(angle)) * r;
  double y = sin(degToRad(angle)) * r;
  return vadd(V(x, y, 0.0), center);
}
static v3 getRandomPointFromSolarModel(v3 center, double radius, const double* fluxRadiusCDF, int nR,
                                       double u0, double u1, double u2) {
  double angle1 = 360.0 * u0, angle2 = 180.0 * u1;
  int rIdx = lower_bound(fluxRadiusCDF, nR, u2);
  double r = (0.0015 + (double)rIdx * 0.0005) * radius;
  double x = cos(degToRad(angle1)) * sin(degToRad(angle2)) * r;
  double y = sin(degToRad(angle1)) * sin(degToRad(angle2)) * r;
  double z = cos(degToRad(angle2)) * r;
  return vadd(V(x, y, z), center);
}
static double getRandomEnergyFromSolarModel(v3 vectorInSun, v3 center, double radius, const sart_tables_t* tb,
                                            double u, int* clamped) {
  double rad = vlength(vsub(vectorInSun, center));
  double r = rad / radius;
  int iRad;
  double indexRad = (r - 0.0015) / 0.0005;
  if (indexRad - 0.5 > floor(indexRad)) iRad = (int)ceil(indexRad);
  else iRad = (int)floor(indexRad);
  if (iRad < 0) { iRad = 0; *clamped = 1; }              /* reference: IndexDefect */
  if (iRad > tb->nRadii - 1) { iRad = tb->nRadii - 1; *clamped = 1; }
  const double* cdf = tb->diffFluxCDFs + (size_t)iRad * tb->nEnergies;
  int idx = lower_bound(cdf, tb->nEnergies, u);
  if (idx > tb->nEnergies - 1) { idx = tb->nEnergies - 1; *clamped = 1; }
  double energy = tb->energies[idx];
  return energy > 0.03 ? energy : 0.03; /* max(0.03.keV, energy) */
}
/* Solar-mode sampling block rt:1754-1764 with the six uniforms in draw order. */
static void sample_solar(const sart_setup_t* s, const sart_tables_t* tb, const centers_t* cv, const double u[6],
                         v3* rayOrigin, v3* pointExit, double* energy, int* clamped) {
  *rayOrigin = getRandomPointFromSolarModel(cv->sun, s->consts.radiusSun, tb->fluxRadiusCDF, tb->nRadii, u[0], u[1], u[2]);
  *pointExit = getRandomPointOnDisk(cv->exitCBMagneticField, s->magnet.radiusCB, u[3], u[4]);
  *energy = getRandomEnergyFromSolarModel(*rayOrigin, cv->sun, s->consts.radiusSun, tb, u[5], clamped);
}
/* X-ray test source sampling rt:1765-1806 (uniform order: disk r, disk phi, then two more). Returns 0 if the
 * ray is stopped by the collimator. */
static int sample_xray(const sart_setup_t* s, const centers_t* cv, const double u[6], v3* rayOrigin, v3* pointExit,
                       double* energy) {
  *rayOrigin = getRandomPointOnDisk(cv->xraySource, s->testSource.radius, u[0], u[1]);
  *energy = s->testSource.energy;
  if (s->testSource.parallel) {
    pointExit->x = rayOrigin->x + (0.5 * u[2]) - 0.25;
    pointExit->y = rayOrigin->y + (0.5 * u[3]) - 0.25;
    pointExit->z = s->magnet.lengthB;
  } else {
    *pointExit = getRandomPointOnDisk(cv->exitCBMagneticField, s->magnet.radiusCB, u[2], u[3]);
  }
  return lineIntersectsCircle(*rayOrigin, *pointExit, cv->collimator, s->testSource.radius);
}

/* ---------------------------------------------------------------- public oracle API */
static void store(const sart_ray_out_t* o, size_t i, const axion_t* a) {
  o->x[i] = a->pointdataX; o->y[i] = a->pointdataY; o->w[i] = a->weights;
  o->code[i] = a->code | (a->passedTillWindow ? SART_FLAG_PASSED_TILL_WINDOW : 0) | (a->clamped ? SART_FLAG_INTERP_CLAMPED : 0);
  o->shell[i] = a->shellNumber;
  if (o->energy) o->energy[i] = a->energiesPre;
  if (o->reflect) o->reflect[i] = a->reflect;
  if (o->transMagnet) o->transMagnet[i] = a->transmissionMagnet;
  if (o->yaw) o->yaw[i] = a->yawAngles;
  if (o->alpha1) o->alpha1[i] = a->alpha1;
  if (o->alpha2) o->alpha2[i] = a->alpha2;
  if (o->pathCB) o->pathCB[i] = a->pathCB;
  if (o->r) o->r[i] = a->pointdataR;
  if (o->deviationDet) o->deviationDet[i] = a->deviationDet;
  if (o->transProbArgon) o->transProbArgon[i] = a->transProbArgon;
}

int oracle_trace_presampled(const sart_setup_t* s, const sart_tables_t* tb, size_t n, const double* origin_xyz,
                            const double* exit_xy, const double* energy, const sart_ray_out_t* out) {
  centers_t cv = initCenterVectors(s);
  int testXray = s->testSource.active;
#pragma omp parallel for schedule(dynamic, 4096)
  for (long long i = 0; i < (long long)n; ++i) {
    axion_t a;
    v3 o = V(origin_xyz[i], origin_xyz[n + i], origin_xyz[2 * n + i]);
    v3 e = V(exit_xy[i], exit_xy[n + i], s->magnet.lengthB);
    trace_after_sampling(&a, s, tb, &cv, s->consts.mAxion, o, e, energy[i], testXray);
    store(out, (size_t)i, &a);
  }
  return 0;
}

static void trace_mc_one(axion_t* a, const sart_setup_t* s, const sart_tables_t* tb, const centers_t* cv,
                         double mAxion, uint64_t seed, uint64_t ray) {
  double u[6];
  oracle_ray_uniforms(seed, ray, u);
  v3 o, e; double en; int clamped = 0;
  if (!s->testSource.active) {
    sample_solar(s, tb, cv, u, &o, &e, &en, &clamped);
  } else if (!sample_xray(s, cv, u, &o, &e, &en)) {
    memset(a, 0, sizeof *a); a->shellNumber = -1; a->code = SART_EXIT_COLLIMATOR; a->energiesPre = en; return;
  }
  trace_after_sampling(a, s, tb, cv, mAxion, o, e, en, s->testSource.active);
  a->clamped |= clamped;
}

int oracle_trace_mc_rays(const sart_setup_t* s, const sart_tables_t* tb, uint64_t first_ray, size_t n, uint64_t seed,
                         const sart_ray_out_t* out) {
  centers_t cv = initCenterVectors(s);
#pragma omp parallel for schedule(dynamic, 4096)
  for (long long i = 0; i < (long long)n; ++i) {
    axion_t a;
    trace_mc_one(&a, s, tb, &cv, s->consts.mAxion, seed, first_ray + (uint64_t)i);
    store(out, (size_t)i, &a);
  }
  return 0;
}
/* The sampled inputs of the MC rays (so that tests can feed them to the pre-sampled entry points). */
int oracle_sample_rays(const sart_setup_t* s, const sart_tables_t* tb, uint64_t first_ray, size_t n, uint64_t seed,
                       double* origin_xyz, double* exit_xy, double* energy) {
  centers_t cv = initCenterVectors(s);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; ++i) {
    double u[6]; v3 o, e; double en; int clamped = 0;
    oracle_ray_uniforms(seed, first_ray + (uint64_t)i, u);
    sample_solar(s, tb, &cv, u, &o, &e, &en, &clamped);
    origin_xyz[i] = o.x; origin_xyz[n + i] = o.y; origin_xyz[2 * n + i] = o.z;
    exit_xy[i] = e.x; exit_xy[n + i] = e.y; energy[i] = en;
  }
  return 0;
}

/* The same two entry points driven by caller-supplied random words instead of Philox (words = SoA [6][n] in the draw
 * order of oracle_ray_uniforms; u = (word + 0.5) 2^-32): lets tests reach the corners of the inverse-CDF sampling
 * (rt:437, 464) — the words 0 and 0xffffffff, words on either side of every CDF entry, flat CDF tails — that random
 * rays reach once in 2^32 draws. Solar source only. */
static void words_to_uniforms(const uint32_t* words, size_t n, size_t i, double u[6]) {
  const double sc = 1.0 / 4294967296.0;
  for (int k = 0; k < 6; ++k) u[k] = ((double)words[(size_t)k * n + i] + 0.5) * sc;
}
int oracle_sample_words(const sart_setup_t* s, const sart_tables_t* tb, size_t n, const uint32_t* words,
                        double* origin_xyz, double* exit_xy, double* energy) {
  if (s->testSource.active) return -1;
  centers_t cv = initCenterVectors(s);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; ++i) {
    double u[6]; v3 o, e; double en; int clamped = 0;
    words_to_uniforms(words, n, (size_t)i, u);
    sample_solar(s, tb, &cv, u, &o, &e, &en, &clamped);
    origin_xyz[i] = o.x; origin_xyz[n + i] = o.y; origin_xyz[2 * n + i] = o.z;
    exit_xy[i] = e.x; exit_xy[n + i] = e.y; energy[i] = en;
  }
  return 0;
}
int oracle_trace_words(const sart_setup_t* s, const sart_tables_t* tb, size_t n, const uint32_t* words,
                       const sart_ray_out_t* out) {
  if (s->testSource.active) return -1;
  centers_t cv = initCenterVectors(s);
#pragma omp parallel for schedule(dynamic, 4096)
  for (long long i = 0; i < (long long)n; ++i) {
    double u[6]; v3 o, e; double en; int clamped = 0;
    axion_t a;
    words_to_uniforms(words, n, (size_t)i, u);
    sample_solar(s, tb, &cv, u, &o, &e, &en, &clamped);
    trace_after_sampling(&a, s, tb, &cv, s->consts.mAxion, o, e, en, 0);
    a.clamped |= clamped;
    store(out, (size_t)i, &a);
  }
  return 0;
}

static void count(sart_counters_t* c, const axion_t* a) {
  c->n_rays++;
  c->n_exit[a->code & 15]++;
  c->n_passed += a->passed;
  c->n_passed_till_window += a->passedTillWindow;
  c->n_hit_nickel += a->hitNickel;
  c->n_interp_clamped += a->clamped;
  if (a->passed) {
    c->sum_w += a->weights; c->sum_w2 += a->weights * a->weights;
    c->sum_x += a->pointdataX; c->sum_y += a->pointdataY; c->sum_r += a->pointdataR;
  }
}
static void merge(sart_counters_t* d, const sart_counters_t* s) {
  d->n_rays += s->n_rays;
  for (int i = 0; i < 16; ++i) d->n_exit[i] += s->n_exit[i];
  d->n_passed += s->n_passed; d->n_passed_till_window += s->n_passed_till_window;
  d->n_hit_nickel += s->n_hit_nickel; d->n_interp_clamped += s->n_interp_clamped;
  d->sum_w += s->sum_w; d->sum_w2 += s->sum_w2; d->sum_x += s->sum_x; d->sum_y += s->sum_y; d->sum_r += s->sum_r;
}
/* prepareHeatmap rt:818-842 with (256, 256, 0, 14, 0, 14, norm = 1), one ray. Returns 0 if the bin is out of range
 * (the reference would raise IndexDefect, e.g. x == 14.0 exactly). */
static int heat_bin(const sart_setup_t* s, double x, double y, int* bin) {
  double stepX = (s->consts.chipXMax - 0.0) / (double)SART_IMAGE_BINS;
  double stepY = (s->consts.chipYMax - 0.0) / (double)SART_IMAGE_BINS;
  double cx = floor((x - 0.0) / stepX), cy = floor((y - 0.0) / stepY);
  if (!(cx >= 0.0 && cx < SART_IMAGE_BINS && cy >= 0.0 && cy < SART_IMAGE_BINS)) return 0;
  *bin = (int)cy * SART_IMAGE_BINS + (int)cx;
  return 1;
}
/* Fused MC run: image[m][256][256] += w, image_w2 += w², counters[m]; masses == NULL → setup mAxion. */
int oracle_trace_mc(const sart_setup_t* s, const sart_tables_t* tb, uint64_t first_ray, uint64_t n_rays, uint64_t seed,
                    int n_masses, const double* masses, double* image, double* image_w2, sart_counters_t* counters) {
  centers_t cv = initCenterVectors(s);
  const size_t NB = (size_t)SART_IMAGE_BINS * SART_IMAGE_BINS;
  if (n_masses < 1) n_masses = 1;
  for (int m = 0; m < n_masses; ++m) {
    double mAx = masses ? masses[m] : s->consts.mAxion;
    double* img = image + (size_t)m * NB;
    double* img2 = image_w2 ? image_w2 + (size_t)m * NB : NULL;
#pragma omp parallel
    {
      double* limg = (double*)calloc(NB, sizeof(double));
      double* limg2 = img2 ? (double*)calloc(NB, sizeof(double)) : NULL;
      sart_counters_t lc; memset(&lc, 0, sizeof lc);
#pragma omp for schedule(dynamic, 4096)
      for (long long i = 0; i < (long long)n_rays; ++i) {
        axion_t a;
        trace_mc_one(&a, s, tb, &cv, mAx, seed, first_ray + (uint64_t)i);
        count(&lc, &a);
        if (a.passed) {
          int bin;
          if (heat_bin(s, a.pointdataX, a.pointdataY, &bin)) {
            limg[bin] = limg[bin] + 1 * a.weights / 1.0;
            if (limg2) limg2[bin] += a.weights * a.weights;
          }
        }
      }
#pragma omp critical
      {
        for (size_t b = 0; b < NB; ++b) img[b] += limg[b];
        if (img2) for (size_t b = 0; b < NB; ++b) img2[b] += limg2[b];
        merge(&counters[m], &lc);
      }
      free(limg); free(limg2);
    }
  }
  return 0;
}

/* prepareHeatmap rt:818-842, general form, serial like the reference. Returns the number of out-of-range points. */
int oracle_prepare_heatmap(int numberOfRows, int numberOfColumns, double start_x, double stop_x, double start_y,
                           double stop_y, size_t n, const double* data_X, const double* data_Y, const double* weight1,
                           double norm, double* result) {
  double stepsize_X = (stop_x - start_x) / (double)numberOfRows;
  double stepsize_Y = (stop_y - start_y) / (double)numberOfColumns;
  int bad = 0;
  memset(result, 0, sizeof(double) * (size_t)numberOfRows * numberOfColumns);
  for (size_t i = 0; i < n; ++i) {
    double fx = floor((data_X[i] - start_x) / stepsize_X), fy = floor((data_Y[i] - start_y) / stepsize_Y);
    if (!(fx >= 0 && fx < numberOfColumns && fy >= 0 && fy < numberOfRows)) { ++bad; continue; }
    size_t k = (size_t)fy * numberOfColumns + (size_t)fx;
    result[k] = result[k] + 1 * weight1[i] / norm;
  }
  return bad;
}

/* CDF build rt:2679-2705. */
int oracle_build_cdfs(int nRadii, int nEnergies, const double* radii, const double* energies, const double* emRates,
                      double* fluxRadiusCDF, double* diffFluxCDFs) {
  double diffRadiusSum = 0.0;
  for (int iRad = 0; iRad < nRadii; ++iRad) {
    double radius = radii[iRad];
    const double* emRate = emRates + (size_t)iRad * nEnergies;
    double* row = diffFluxCDFs + (size_t)iRad * nEnergies;
    double diffSum = 0.0;
    for (int iE = 0; iE < nEnergies; ++iE) {
      double energy = energies[iE];
      double diffFlux = emRate[iE] * (energy * energy) * radius * radius;
      diffSum += diffFlux;
      row[iE] = diffSum;
    }
    diffRadiusSum += diffSum;
    fluxRadiusCDF[iRad] = diffRadiusSum;
    double integral = row[nEnergies - 1];
    for (int iE = 0; iE < nEnergies; ++iE) row[iE] = row[iE] / integral;
  }
  double integral = fluxRadiusCDF[nRadii - 1];
  for (int i = 0; i < nRadii; ++i) fluxRadiusCDF[i] = fluxRadiusCDF[i] / integral;
  return 0;
}

/* calcWindowVals rt:1431-1462. */
void oracle_calc_window_vals(double radiusWindow, int numberOfStrips, double openApertureRatio, double* width, double* dist) {
  double totalArea = PI * radiusWindow * radiusWindow;
  double areaOfStrips = totalArea * (1.0 - openApertureRatio);
  double dAndwPerStrip = radiusWindow * 2.0 / ((double)numberOfStrips + 1.0);
  double lengthStrip, lengthAllStrips = 0.0;
  int nHalf = (int)round((double)numberOfStrips / 2.0);
  for (int i = 0; i <= nHalf - 1; ++i) {
    double fi = (double)i;
    lengthStrip = sqrt(radiusWindow * radiusWindow -
                       (fi * dAndwPerStrip + 0.5 * dAndwPerStrip) * (fi * dAndwPerStrip + 0.5 * dAndwPerStrip)) * 2.0;
    lengthAllStrips = lengthAllStrips + lengthStrip;
  }
  lengthAllStrips = lengthAllStrips * 2.0;
  *width = areaOfStrips / lengthAllStrips;
  *dist = dAndwPerStrip - *width;
}

/* lengthTelescope rt:1883-1884, exported for the known-answer test (optics_exit z = 454.0, rt:1260). */
double oracle_length_telescope(const sart_setup_t* s) {
  const sart_telescope_t* tel = &s->telescope;
  return (tel->lMirror + 0.5 * tel->allXsep[0]) * cos(degToRad(tel->allAngles[0])) +
         (tel->lMirror + 0.5 * tel->allXsep[0]) * cos(3.0 * degToRad(tel->allAngles[0]));
}

/* TestMirrors.nim:80-121 scenario: a ray through a single shell; returns alpha1, alpha2 (deg) and the exit
 * direction of the cone path. */
void oracle_test_mirrors(double r1, double xSep, double betaDeg, double lMirror, const double pointCB[3],
                         const double pointXRT[3], double out_alpha[2], double out_dir[3]) {
  double beta = degToRad(betaDeg), beta3 = 3.0 * beta;
  double r2 = r1 - lMirror * sin(beta);
  double r3 = r2 - 0.5 * xSep * tan(beta);
  double r4 = r3 - 0.5 * xSep * tan(3.0 * beta);
  double distanceMirrors = cos(beta) * (xSep + lMirror);
  v3 pCB = V(pointCB[0], pointCB[1], pointCB[2]), pX = V(pointXRT[0], pointXRT[1], pointXRT[2]);
  v3 pm1 = findPosCone(pX, pCB, r1, beta, lMirror, 0.0);
  v3 v1 = getVectoraAfterMirror(pX, pCB, pm1, beta, r1, lMirror, 0.0, MS_CONE);
  v3 pam1 = vadd(pm1, vscale(200.0, v1));
  v3 pm2 = findPosCone(pam1, pm1, r4, beta3, lMirror, distanceMirrors);
  v3 v2 = getVectoraAfterMirror(pam1, pm1, pm2, beta3, r1, lMirror, 0.0, MS_CONE);
  out_alpha[0] = getMirrorAngle(pX, pCB, pm1, beta, r1, lMirror, 0.0, MS_CONE);
  out_alpha[1] = getMirrorAngle(pam1, pm1, pm2, beta3, r1, lMirror, 0.0, MS_CONE);
  out_dir[0] = v2.x; out_dir[1] = v2.y; out_dir[2] = v2.z;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
