// derive_fast.cpp — host derivation of the parameter blocks and look-up tables of the "fast" pipeline.
//
// What the reference evaluates per ray but depends only on the (tabulated, hence discrete — rt:470) axion energy
// is tabulated once per energy index here: window / strongback / detector-gas transmissions (rt:2170-2190, linear
// interpolation on the Henke grids), the energy cell and in-cell offset of the bilinear reflectivity lookup
// (rt:1567-1578) and the He mass attenuation of the buffer-gas stage (axionMassforMagnet.nim:70-73).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fast_params.h"
#include "sart_internal.h"

namespace sart {
namespace fast {

static constexpr double kPi = 3.141592653589793;

static double lin1d(const sart_interp1d_t& t, double x) {  // numericalnim newLinear1D.eval, clamped
  const int n = t.n;
  if (n < 1 || !t.x || !t.y) return 0.0;
  if (n == 1) return t.y[0];
  x = std::min(std::max(x, t.x[0]), t.x[n - 1]);
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (t.x[mid] <= x) lo = mid; else hi = mid;
  }
  return t.y[lo] + (x - t.x[lo]) * ((t.y[lo + 1] - t.y[lo]) / (t.x[lo + 1] - t.x[lo]));
}

static double he_density(double p, double temp) {  // axionMassforMagnet.nim:4-15
  return (p * 1e2) * 4.002602 / (8.314 * temp * 1000.0) / 1000.0;
}

bool supported(const sart_setup_t& s, const char** why) {
  const sart_telescope_t& t = s.telescope;
  if (t.kind == SART_TK_XMM && t.holeType != SART_HT_NONE && (t.numberOfHoles < 1 || t.numberOfHoles > 64)) {
    *why = "XMM hole pattern: the throughput pipelines take 1 to 64 holes";
    return false;
  }
  for (int j = 1; j < t.nShells; ++j)
    if (!(t.allR1[j] > t.allR1[j - 1] + t.allThickness[j - 1])) {
      *why = "shell radii must increase and shells must not overlap for the fast pipeline";
      return false;
    }
  return true;
}

void derive_shells(const sart_setup_t& s, const ShellF64* a, ShellFast* out) {
  const sart_telescope_t& t = s.telescope;
  const double l = t.lMirror;
  for (int j = 0; j < t.nShells; ++j) {
    ShellFast& o = out[j];
    std::memset(&o, 0, sizeof o);
    const double beta = t.allAngles[j] * (kPi / 180.0), beta3 = 3.0 * beta;
    o.R1 = a[j].R1; o.R1pT = a[j].R1pT; o.r1sq = a[j].r1sq;
    o.tan1 = a[j].tan1; o.zmax1 = a[j].zmax1; o.cosb = std::cos(beta); o.sinb = std::sin(beta);
    o.r4 = a[j].r4; o.tan2 = a[j].tan2; o.dm = a[j].distanceMirrors; o.zmax2 = a[j].zmax2;
    o.cos3b = std::cos(beta3); o.sin3b = std::sin(beta3);
    o.ddWin = a[j].ddWin; o.distDet = a[j].distDet;
    o.p_e = a[j].p_e; o.p_c0 = a[j].p_r3sq + a[j].p_e * l; o.p_r3sq = a[j].p_r3sq; o.p_r3tan = a[j].p_r3tan;
    o.h_e = a[j].h_e; o.h_g = a[j].h_g; o.h_r3sq = a[j].h_r3sq; o.h_r3tan = a[j].h_r3tan;
    o.h_inv_nden = a[j].h_nden != 0.0 ? 1.0 / a[j].h_nden : 0.0;
    int coat = 0;
    if (t.reflKind == SART_RK_MULTI_COATING) {
      while (coat < t.nCoatings && t.layers[coat] < j) ++coat;   // layers.lowerBound(hitLayer) rt:1573
      if (coat > t.nCoatings - 1) coat = (t.nCoatings - 1) | kCoatClamped;   // past the last coating: clamped + flagged, as in the exact pipeline
    }
    o.coat = coat;
  }
}

// Error budgets of the FP32 decisions (fast_params.h: Tol32), in mm unless stated. eps = 2^-24 is the unit roundoff of
// FP32. Three sources are budgeted (DESIGN.md section 3b has the derivation):
//  (1) FP32 rounding of the pipeline itself: a coordinate of magnitude R carries ~eps R per operation;
//  (2) Monte Carlo rays only: the fast sampling arithmetic (MUFU sin/cos, absolute error 2^-21.4) moves the exit-disc
//      point by ~1e-6 radiusCB and the emission point by ~1e-6 rs R_sun, i.e. the direction by 1e-6 rs R_sun / D;
//  (3) the reference's own rounding: it intersects planes with the line O + lambda (E - O) through a point D = 1.5e14 mm
//      away (rt:481-492, 529-534), which leaves pointExitCB off the true line by up to ulp(|O.x|) + |slope| ulp(|O.z|)
//      <= 2^-51 D (|sx| + |sy|); every later point is extrapolated from E (exact) through that point, so the error grows
//      with z / (zExitCB - lengthB) and reaches the detector multiplied by focal length / (zExitCB - lengthB).
// `scale` multiplies every budget (sart_set_retrace).
void derive_tolerances(const FastParams& f, const ShellFast* a, int nShells, float scale, Tol32* t) {
  std::memset(t, 0, sizeof *t);
  const double eps = 5.9604644775390625e-08;
  const double rPipe = std::sqrt(f.rPipe12);
  double r1max = 0.0, tanMax = 0.0, focal = 0.0;
  for (int j = 0; j < nShells; ++j) {
    r1max = std::max(r1max, a[j].R1pT);
    tanMax = std::max(tanMax, std::fabs(a[j].tan2));
    focal = std::max(focal, std::fabs(a[j].ddWin));
  }
  const double rmax = std::max({f.radiusCB, rPipe + std::fabs(f.oeX), rPipe + std::fabs(f.oeY), r1max});
  const double L = std::fabs(f.dzPipe2) + 100.0;                 // farthest plane of stage A behind the field exit (spider: 85 mm)
  const double delta = std::max(f.dzExitCB, 1.0);                // zExitCB - lengthB: the base of the reference's extrapolation
  const double lever = std::max(L / delta, 1.0);
  const double D = f.testXray ? std::fabs(f.lengthB - f.srcZ) : f.sunDist;
  const double kRef = 1.5 * 4.440892098500626e-16 * D;           // (3), per unit (|sx| + |sy|), safety 1.5
  const double kSamp = f.testXray ? 0.0 : 2e-6 * f.radiusSun / f.sunDist;   // (2) direction, per unit rs, safety 2
  const double tolE = f.testXray ? 16.0 * eps * (f.radiusCB + f.srcRadius) * (1.0 + L / std::max(D, 1.0))
                                 : 16.0 * eps * f.radiusCB;      // (2) exit-disc point
  // a turned telescope: the frame rotation doubles the rounding steps of every point and direction
  const double rot = f.rotated ? 2.0 : 1.0;
  t->latA = float(scale * rot * (tolE + 8.0 * eps * rmax));
  t->latS = float(scale * kSamp * L);
  t->latT = float(scale * (kRef * lever + 4.0 * eps * L));
  t->latTpre = float(scale * 4.0 * eps * L);
  t->latRef = float(scale * 1.5 * lever);
  // entrance plane: the slope terms act over lengthB instead of L, the reference's noise enters without the lever
  t->entK = float(std::max(1.0, std::fabs(f.lengthB) / L));
  // At the detector the optic maps directions to positions (position = focal length x angle; a lateral shift of the
  // incoming ray does not move its image), so the budget is the direction error times the distance fl to the detector:
  // ~6 rounding steps of eps |v_lateral| through the two reflections (measured against 60-digit arithmetic: 4e-5 mm on
  // CAST+LLNL, 1.1e-4 mm on BabyIAXO+XMM; budgeted 4x), the slope errors (2) and (3), and the aperture-plane budget once.
  const double vlat = 1.4 * tanMax + 0.005;                      // lateral direction components after two reflections (~4 beta)
  const double fl = focal + f.lMirror * 2.0;
  t->detA = float(scale * rot * (24.0 * eps * vlat * fl + tolE + 8.0 * eps * rmax));
  t->detS = float(scale * kSamp * fl);
  t->detT = float(scale * (kRef * fl / delta + 4.0 * eps * fl));
  t->detTpre = float(scale * 4.0 * eps * fl);
  t->detRef = float(scale * 1.5 * fl / delta);
  t->rho = float(scale * 4.0 * eps * r1max);
  t->circ2 = float(scale * 8.0 * eps);
  t->spider = float(scale * 512.0 * eps);
  t->cond = float(scale * 2e-15);
  t->zrel = float(scale * 16.0 * eps);
  t->discRel = float(scale * 2e-3);   // |A| budget(C) / hb^2 stays below 1e-3 for every shell of the three optics
  t->ang = float(scale * 2e-5);
  t->sinA = float(scale * 8.0 * eps * (4.0 * tanMax + 0.01));
  t->circCB = float(t->circ2 * f.radiusCB2);
  t->circPipe = float(t->circ2 * f.rPipe12);
  t->circWin = float(t->circ2 * f.radiusWindow2);
  t->angLo = f.angleMax - t->ang;
  {
    double gap = 0.0;
    for (int j = 1; j < nShells; ++j) gap = std::max(gap, a[j].R1 - a[j - 1].R1pT);
    t->nick = float(t->sinA * f.lMirror + t->zrel * gap);
  }
  t->twoRcb = float(2.0 * f.radiusCB);
  t->twoRpipe = float(2.0 * rPipe);
  t->twoRwin = float(2.0 * std::sqrt(f.radiusWindow2));
  t->chipInside = (std::min(f.chipCX, f.chipCY) < std::sqrt(f.radiusWindow2) * 1.001 + 4.0 * t->detA) ? 1 : 0;
  // Development hook (tools/fuzz_setups.py): SART_TOL_BOOST="name=factor,..." multiplies single budgets, to find out which one
  // a misclassified ray was short of.
  if (const char* e = std::getenv("SART_TOL_BOOST")) {
    struct { const char* n; float* p; } M[] = {{"latS", &t->latS}, {"latT", &t->latT}, {"latA", &t->latA}, {"detS", &t->detS},
        {"detT", &t->detT}, {"detA", &t->detA}, {"rho", &t->rho}, {"discRel", &t->discRel}, {"zrel", &t->zrel}, {"nick", &t->nick},
        {"sinA", &t->sinA}, {"cond", &t->cond}, {"spider", &t->spider}, {"circCB", &t->circCB}, {"circPipe", &t->circPipe},
        {"circWin", &t->circWin}, {"entK", &t->entK}};
    for (auto& m : M) {
      const std::string key = std::string(m.n) + "=";
      const char* at = std::strstr(e, key.c_str());
      if (at && (at == e || at[-1] == ',')) *m.p *= float(std::atof(at + key.size()));
    }
  }
}

// Single-precision blocks of precision mode 2, rounded from the FP64 ones.
void derive_f32(const FastParams& f, const ShellFast* a, int nShells, Geo32* g, ShellF32* out, float tolScale) {
  std::memset(g, 0, sizeof *g);
  derive_tolerances(f, a, nShells, tolScale, &g->tol);
  g->depthOverCos = float(f.depthOverCos);
  g->radiusCB = float(f.radiusCB); g->radiusCB2 = float(f.radiusCB2); g->lengthB = float(f.lengthB);
  g->lengthB2 = float(f.lengthB * f.lengthB); g->lengthBplusSun = float(f.lengthB + f.sunDist); g->radiusSun = float(f.radiusSun);
  g->dzExitCB = float(f.dzExitCB); g->dzPipe1 = float(f.dzPipe1); g->dzPipe2 = float(f.dzPipe2); g->rPipe12 = float(f.rPipe12);
  g->cosTX = float(f.cosTX); g->sinTX = float(f.sinTX); g->cosTY = float(f.cosTY); g->sinTY = float(f.sinTY);
  g->halfLenTel = float(f.halfLenTel); g->oeX = float(f.oeX); g->oeY = float(f.oeY); g->zExitCBtel = float(f.zExitCBtel);
  g->lMirror = float(f.lMirror); g->cosPipe = float(f.cosPipe); g->sinPipe = float(f.sinPipe); g->dShift = float(f.dShift);
  g->lateralShift = float(f.lateralShift); g->transversalShift = float(f.transversalShift);
  g->radiusWindow2 = float(f.radiusWindow2); g->chipCX = float(f.chipCX); g->chipCY = float(f.chipCY);
  g->cosTheta = float(f.cosTheta); g->sinTheta = float(f.sinTheta); g->stripDist = float(f.stripDist);
  g->stripWidth = float(f.stripWidth); g->invStripPitch = float(f.invStripPitch); g->invBinX = float(f.invBinX);
  g->invBinY = float(f.invBinY); g->shellRhoMin = float(f.shellRhoMin); g->shellInvStep = float(f.shellInvStep);
  g->srcX = float(f.srcX); g->srcY = float(f.srcY); g->srcRadius = float(f.srcRadius); g->srcRadius2 = float(f.srcRadius2);
  g->invSrcDz = float(1.0 / (f.lengthB - f.srcZ)); g->colDz = float(f.colDz);
  g->holeR = float(f.holeInOptics);
  for (int j = 0; j < nShells; ++j) {
    const ShellFast& s = a[j];
    ShellF32& o = out[j];
    std::memset(&o, 0, sizeof o);
    o.R1 = float(s.R1); o.R1pT = float(s.R1pT); o.tan1 = float(s.tan1); o.zmax1 = float(s.zmax1);
    o.cosb = float(s.cosb); o.sinb = float(s.sinb); o.r4 = float(s.r4); o.tan2 = float(s.tan2); o.dm = float(s.dm);
    o.zmax2 = float(s.zmax2); o.cos3b = float(s.cos3b); o.sin3b = float(s.sin3b); o.ddWin = float(s.ddWin);
    o.p_e = float(s.p_e); o.p_R0 = float(std::sqrt(std::max(s.p_c0, 0.0))); o.p_r3sq = float(s.p_r3sq); o.p_r3tan = float(s.p_r3tan);
    o.h_e = float(s.h_e); o.h_g = float(s.h_g); o.h_r3sq = float(s.h_r3sq); o.h_r3tan = float(s.h_r3tan);
    o.h_inv_nden = float(s.h_inv_nden); o.coat = s.coat;
    o.zmid1 = float(0.5 * s.zmax1); o.zhalf1 = float(0.5 * s.zmax1);
    o.zmid2 = float(0.5 * (s.dm + s.zmax2)); o.zhalf2 = float(0.5 * (s.zmax2 - s.dm));
    o.twoR = 2.0f * (f.telKind == SART_TK_XMM || f.telKind == SART_TK_ABRIXAS ? o.p_R0 : o.R1);
    o.tan2p = float(std::fabs(s.tan2) + 0.01);
  }
}

// Uniform radial grid over [R1[0] - step, R1[last]]: guide[b] = smallest j with R1[j] > lower edge of bucket b, and the
// step is at most half the smallest shell spacing, so the kernel's forward scan from guide[b] takes 0 or 1 steps.
void build_shell_guide(const sart_setup_t& s, FastParams* f, std::vector<uint8_t>* guide) {
  const sart_telescope_t& t = s.telescope;
  // smallest distance between two boundaries of the radial classification (front radius R1[j], outer glass edge
  // R1[j] + thickness[j]): with a step of half of it no bucket holds two of them (build_shell_table verifies)
  double gap = t.allR1[0];
  for (int j = 0; j < t.nShells; ++j) {
    if (t.allThickness[j] > 0.0) gap = std::min(gap, t.allThickness[j]);
    if (j > 0) {
      gap = std::min(gap, t.allR1[j] - t.allR1[j - 1]);
      const double open = t.allR1[j] - (t.allR1[j - 1] + t.allThickness[j - 1]);
      if (open > 0.0) gap = std::min(gap, open);
    }
  }
  double step = 0.5 * gap;
  const double span = t.allR1[t.nShells - 1] - t.allR1[0];
  if (span / step > 4000.0) step = span / 4000.0;
  if (!(step > 0.0)) step = 1.0;
  const double rmin = t.allR1[0] - 1.5 * step;   // R1[0] sits mid-bucket 1: bucket 0 holds no boundary whatever the rounding
  const int n = int(std::ceil((t.allR1[t.nShells - 1] - rmin) / step)) + 2;
  guide->assign(size_t(n), 0);
  for (int b = 0; b < n; ++b) {
    const double edge = rmin + double(b) * step;
    int j = 0;
    while (j < t.nShells - 1 && !(t.allR1[j] > edge)) ++j;
    (*guide)[b] = uint8_t(j);
  }
  f->shellRhoMin = rmin;
  f->shellInvStep = 1.0 / step;
  f->nShellGuide = n;
}

// Walker/Vose alias table of the discrete distribution the integer thresholds define: index i has the weight
// c[i] = thr[i] - thr[i-1] (thr[-1] = 0, and the last index takes what is left of 2^32), i.e. exactly the number of
// 32-bit words the inverse-CDF search maps to i. Exact integer arithmetic: masses c[i] n against a bucket capacity of
// 2^32. Entry k: bits 31..11 = the share bucket k keeps for itself, floored to units of 2^11 (a bucket that keeps
// everything is its own alias, so the flooring never loses it), bits 10..0 = alias. n <= 2048.
bool build_alias_table(const uint32_t* thr, int n, uint32_t* out) {
  if (n < 1 || n > 2048) return false;
  const uint64_t cap = uint64_t(1) << 32;
  std::vector<uint64_t> mass(size_t(n), 0);
  uint64_t prev = 0;
  for (int i = 0; i < n; ++i) {
    const uint64_t t = i < n - 1 ? uint64_t(thr[i]) : cap;   // the last index ends the range
    const uint64_t hi = t < prev ? prev : t;
    mass[size_t(i)] = (hi - prev) * uint64_t(n);
    prev = hi;
  }
  std::vector<int> small, large;
  small.reserve(size_t(n)); large.reserve(size_t(n));
  for (int i = 0; i < n; ++i) (mass[size_t(i)] < cap ? small : large).push_back(i);
  std::vector<uint64_t> keep(size_t(n), cap);
  std::vector<int> alias(static_cast<size_t>(n), 0);
  for (int i = 0; i < n; ++i) alias[size_t(i)] = i;
  while (!small.empty() && !large.empty()) {
    const int s = small.back(); small.pop_back();
    const int l = large.back();
    keep[size_t(s)] = mass[size_t(s)];
    alias[size_t(s)] = l;
    mass[size_t(l)] -= cap - mass[size_t(s)];   // exact: the sum of all masses is n 2^32
    if (mass[size_t(l)] < cap) { large.pop_back(); small.push_back(l); }
  }
  // what is left (either list) holds exactly one capacity each: such a bucket keeps everything
  for (int i = 0; i < n; ++i) {
    uint32_t share = keep[size_t(i)] >= cap ? 0xfffff800u : uint32_t(keep[size_t(i)]) & 0xfffff800u;
    const int a = keep[size_t(i)] >= cap ? i : alias[size_t(i)];
    out[i] = share | uint32_t(a);
  }
  return true;
}

// Shell search of stage A for one radial distance, as the FP32 kernels decide it (rt:1932-1957 on the f32 shell records):
// the shell number, or kShellCellFail + exit code.
int classify_radius(const ShellF32* sh, int nS, float rho) {
  int hit = 0;
  while (hit < nS - 1 && !(sh[hit].R1 > rho)) ++hit;   // first j with R1[j] > rho
  int code = -1;
  if (!(sh[hit].R1 > rho)) code = SART_EXIT_NO_MIRROR_HIT;   // == R1[last] (or NaN)
  if (hit > 0 && rho < sh[hit - 1].R1pT && rho > sh[hit - 1].R1) code = SART_EXIT_GLASS_FRONT;
  if (rho > sh[nS - 1].R1) code = SART_EXIT_OUTSIDE_SHELLS;
  return code >= 0 ? kShellCellFail + code : hit;
}

// One record per radial bucket of the shell guide (fast_params.h: ShellCell). The bucket of a radius is computed with the
// device's own FP32 expression, which is monotone in the radius, so the radii of one bucket form an interval and the
// classification inside it changes only at the boundaries that map to this bucket. Returns false when two boundaries share
// a bucket (shells closer than the grid can resolve): the throughput pipelines then report the setup as unsupported.
bool build_shell_table(const Geo32& g, const ShellF32* sh, int nS, int nBuckets, std::vector<ShellCell>* out) {
  if (nS < 1 || nS > kShellCellFail || nBuckets < 2) return false;
  auto bucket = [&](float rho) {
    const float d = rho - g.shellRhoMin;
    const float x = d * g.shellInvStep;
    int b = x >= float(nBuckets) ? nBuckets - 1 : int(x);   // F2I.TRUNC saturates; NaN -> 0
    if (!(x == x)) b = 0;
    return std::max(0, std::min(b, nBuckets - 1));
  };
  std::vector<float> bnd;
  for (int j = 0; j < nS; ++j) {
    bnd.push_back(sh[j].R1);
    if (j < nS - 1 && sh[j].R1pT > sh[j].R1) bnd.push_back(sh[j].R1pT);
  }
  std::sort(bnd.begin(), bnd.end());
  bnd.erase(std::unique(bnd.begin(), bnd.end()), bnd.end());
  std::vector<int> owner(size_t(nBuckets), -1);
  for (size_t k = 0; k < bnd.size(); ++k) {
    const int b = bucket(bnd[k]);
    if (b == 0) return false;            // bucket 0 (everything inside the first shell, and NaN) must be boundary-free
    if (owner[size_t(b)] >= 0) return false;
    owner[size_t(b)] = int(k);
  }
  out->assign(size_t(nBuckets), ShellCell{0.f, 0u});
  // bucket 0: radii below R1[0] - step / 2 ... and NaN (neither < B nor > B): no mirror hit, like the scan it replaces
  int cur = classify_radius(sh, nS, 0.0f);   // outcome of the open interval the next bucket starts in
  (*out)[0].B = -INFINITY;   // every radius (a norm: >= 0) is "above"; NaN is "at"
  (*out)[0].sel = uint32_t(cur) | (uint32_t(kShellCellFail + SART_EXIT_NO_MIRROR_HIT) << 8) | (uint32_t(cur) << 16);
  for (int b = 1; b < nBuckets; ++b) {
    ShellCell& c = (*out)[size_t(b)];
    if (owner[size_t(b)] < 0) {
      c.B = -INFINITY;   // every radius of the bucket is "above"
      c.sel = uint32_t(cur) | (uint32_t(cur) << 8) | (uint32_t(cur) << 16);
      continue;
    }
    const float B = bnd[size_t(owner[size_t(b)])];
    const int below = classify_radius(sh, nS, std::nextafterf(B, -INFINITY));
    const int at = classify_radius(sh, nS, B);
    const int above = classify_radius(sh, nS, std::nextafterf(B, INFINITY));
    if (below != cur) return false;   // cannot happen with one boundary per bucket; guards the construction
    c.B = B;
    c.sel = uint32_t(below) | (uint32_t(at) << 8) | (uint32_t(above) << 16);
    cur = above;
  }
  // Boundary-free buckets carry the nearest boundary of a neighbouring bucket as B (all three outcomes are the same, so
  // the lookup does not change): the kernel's margin test |rho - B| <= budget then also sees a boundary that sits just
  // across the bucket's edge. Buckets are at most half the smallest boundary distance wide, so one B per bucket suffices
  // for budgets below a quarter of a bucket.
  for (int b = 1; b < nBuckets; ++b) {
    if (owner[size_t(b)] >= 0) continue;
    const float lo = g.shellRhoMin + float(b) / g.shellInvStep, hi = g.shellRhoMin + float(b + 1) / g.shellInvStep;
    float best = -INFINITY, bestDist = INFINITY;
    for (int nb = std::max(1, b - 2); nb <= std::min(nBuckets - 1, b + 2); ++nb) {
      if (owner[size_t(nb)] < 0) continue;
      const float B = bnd[size_t(owner[size_t(nb)])];
      const float d = B < lo ? lo - B : (B > hi ? B - hi : 0.0f);
      if (d < bestDist) { bestDist = d; best = B; }
    }
    (*out)[size_t(b)].B = best;
  }
  return true;
}

// The device's table lookup (kernels_f32.cu, stage A), restated on the host for sart_shell_lookup.
int shell_table_lookup(const Geo32& g, const std::vector<ShellCell>& tab, float rho) {
  const int n = int(tab.size());
  const float x = (rho - g.shellRhoMin) * g.shellInvStep;
  int b = !(x == x) ? 0 : (x >= float(n) ? n - 1 : (x <= -1.0f ? 0 : int(x)));   // F2I.TRUNC saturates, NaN -> 0
  b = std::max(0, std::min(b, n - 1));
  const ShellCell& c = tab[size_t(b)];
  const int k = rho < c.B ? 0 : (rho > c.B ? 2 : 1);
  return int((c.sel >> (8 * k)) & 0xffu);
}

void derive_params(const sart_setup_t& s, const Params& P, FastParams* f) {
  std::memset(f, 0, sizeof *f);
  f->radiusCB2 = P.radiusCB * P.radiusCB;
  f->radiusCB = P.radiusCB;
  f->lengthB = P.lengthB;
  f->dzExitCB = P.zExitCB - P.lengthB;
  f->dzPipe1 = P.zPipe1 - P.lengthB;
  f->dzPipe2 = P.zPipe2 - P.lengthB;
  f->rPipe12 = P.rPipe1 * P.rPipe1;
  f->cosTX = P.cosTX; f->sinTX = P.sinTX; f->cosTY = P.cosTY; f->sinTY = P.sinTY;
  f->halfLenTel = P.halfLenTel; f->oeX = P.oeX; f->oeY = P.oeY;
  f->zExitCBtel = P.zExitCB - P.zPipe2;
  f->lMirror = P.lMirror;
  f->cosPipe = P.cosPipe; f->sinPipe = P.sinPipe; f->dShift = P.dShift;
  f->lateralShift = P.lateralShift; f->transversalShift = P.transversalShift;
  f->radiusWindow2 = P.radiusWindow * P.radiusWindow;
  f->chipCX = P.chipCX; f->chipCY = P.chipCY;
  f->cosTheta = P.cosTheta; f->sinTheta = P.sinTheta;
  f->stripDist = P.stripDist; f->stripWidth = P.stripWidth;
  f->invStripPitch = 1.0 / (P.stripDist + P.stripWidth);
  f->depthOverCos = s.detector.depthDet / P.cosPipe;
  f->invBinX = double(SART_IMAGE_BINS) / (2.0 * P.chipCX);
  f->invBinY = double(SART_IMAGE_BINS) / (2.0 * P.chipCY);
  f->sunDist = s.consts.distanceSunEarth;
  f->radiusSun = s.consts.radiusSun;
  {
    const double k = (P.g_agamma * 1e-9) * (P.B * P.tesla_to_eV2) * (1e-3 * P.m_to_inv_eV) / 2.0;
    f->convK = float(k * k);
  }
  f->exposure = float(P.exposureFactor);
  f->angleMin = float(P.angleMin); f->angleMax = float(P.angleMax);
  f->invReflDx = float(1.0 / P.reflDx);
  // buffer gas (axionMassforMagnet.nim:51-61, 75-113); note pGas is in bar but consumed as mbar (quirk Q4)
  {
    const double rhoMagnet = he_density(P.pGas, P.tGas), rhoPipe = he_density(P.pGas, P.roomTemp);
    f->gasGamma0 = 1.97e-7 * 100.0 * rhoMagnet;
    const double ne = 2.0 * 6.022e23 * ((P.pGas * 1e2) / (8.314 * P.tGas));  // amountMol / vol
    const double mg = std::sqrt(std::pow(1.97e-7, 3.0) * 4.0 * kPi * (1.0 / 137.0) * ne / 511e3);
    f->gasMgamma2 = mg * mg;
    const double t1 = (P.g_agamma * 1e-9) * (P.B * 1e3 / 1.444) / 2.0;
    f->gasTerm1 = t1 * t1;
    f->gasRhoPipe100 = rhoPipe * 100.0;
    f->gasRhoMagnet100 = rhoMagnet * 100.0;
  }
  f->srcX = P.srcX; f->srcY = P.srcY; f->srcZ = P.srcZ; f->srcRadius = P.srcRadius;
  f->colDz = P.colZ - P.srcZ;
  f->srcRadius2 = P.srcRadius * P.srcRadius;
  f->srcEnergy = float(P.srcEnergy);
  f->telKind = P.telKind; f->nShells = P.nShells; f->reflKind = P.reflKind; f->nCoatings = P.nCoatings;
  f->stage = P.stage; f->nStripHalf = P.nStripHalf; f->testXray = P.testXray; f->parallelSource = P.parallelSource;
  for (int i = 0; i < SART_MAX_COATINGS; ++i) f->layers[i] = P.layers[i];
  f->flags = P.flags;
  f->nRadii = P.nRadii; f->nEnergies = P.nEnergies; f->nAngles = P.nAngles; f->nReflEnergies = P.nReflEnergies;
  f->shellsMonotonic = 1;
  f->rotated = (P.sinTX != 0.0 || P.sinTY != 0.0) ? 1 : 0;
  f->srcEIdx = P.nEnergies;  // the record after the tabulated energies holds the X-ray source energy
  f->holeType = P.telKind == SART_TK_XMM ? P.holeType : SART_HT_NONE;
  f->numberOfHoles = P.numberOfHoles;
  f->holeInOptics = P.holeInOptics;
  {
    // largest slope of a ray from the outermost tabulated solar shell through the field-exit disc, and the largest radius
    // it can have at the second pipe plane if it passed the bore exit; 1 mm is far above every error budget
    const double rsMax = 0.0015 + 0.0005 * double(std::max(P.nRadii, 1) - 1);
    const double sMax = (rsMax * s.consts.radiusSun + P.radiusCB) / (s.consts.distanceSunEarth - s.consts.radiusSun);
    const double reach = P.radiusCB + sMax * std::fabs(P.zPipe2 - P.lengthB) * 1.5 + 1.0;
    f->pipesFree = (!P.testXray && P.nRadii > 0 && reach < P.rPipe1 && P.zPipe2 >= P.zPipe1 && P.zPipe1 >= P.zExitCB) ? 1 : 0;
  }
}

// eval_linear1d (trace_exact.cuh) flags an abscissa outside the grid, or a grid it cannot interpolate on
static bool lin1d_clamps(const sart_interp1d_t& t, double x) {
  if (t.n < 2 || !t.x || !t.y) return true;
  return !(x >= t.x[0]) || !(x <= t.x[t.n - 1]);
}

static void lut_entry(double E, const sart_interp1d_t& sb, const sart_interp1d_t& wd, const sart_interp1d_t& ga,
                      double reflEMin, double reflEMax, EnergyLUT* e, GasLUT* g) {
  // A transmission that is non-zero in f64 must stay non-zero in the f32 table: `passed` means weight != 0
  // (rt:2220), and e.g. 200 um of Si transmit 1e-60 at 0.3 keV.
  auto f32nz = [](double v) { return (v != 0.0 && std::fabs(v) < 1.2e-38) ? float(std::copysign(1.2e-38, v)) : float(v); };
  e->Twindow = f32nz(lin1d(wd, E));
  {
    const double t = lin1d(sb, E);
    int ex = 0;
    if (t != 0.0 && std::fabs(t) < 1e-30) {   // mantissa in [0.5, 1) as the FP32 value, the exponent beside it
      const double m = std::frexp(t, &ex);
      e->Tstrongback = float(m);
    } else {
      e->Tstrongback = float(t);
    }
    int cl = 0;
    if (lin1d_clamps(wd, E)) cl |= kLutClampWindow;
    if (lin1d_clamps(sb, E)) cl |= kLutClampStrongback;
    if (lin1d_clamps(ga, E)) cl |= kLutClampGas;
    if (!(E >= reflEMin) || !(E <= reflEMax)) cl |= kLutClampRefl;
    e->sbExp = (ex & 0xffff) | (cl << 16);
  }
  e->Agas = f32nz(lin1d(ga, E));
  const double lma = -1.5832 + 5.9195 * std::exp(-0.353808 * E) + 4.03598 * std::exp(-0.970557 * E);
  g->massAtt = float(std::exp(lma));
  g->inv2E = float(1.0 / (2.0 * (E * 1000.0)));
}

// nEnergies records (E = max(0.03 keV, energies[i]), rt:470-471) + one for the X-ray test-source energy.
void build_energy_lut(int nE, const double* energies, const sart_interp1d_t& sb, const sart_interp1d_t& wd,
                      const sart_interp1d_t& ga, double srcEnergy, double reflEMin, double reflEMax,
                      std::vector<EnergyLUT>* out, std::vector<GasLUT>* gout) {
  out->resize(size_t(nE) + 1);
  gout->resize(size_t(nE) + 1);
  for (int i = 0; i <= nE; ++i)
    lut_entry(i < nE ? std::max(0.03, energies[i]) : srcEnergy, sb, wd, ga, reflEMin, reflEMax, &(*out)[i], &(*gout)[i]);
}

// rkEffectiveArea: the telescope transmission (eval_linear1d of trace_exact.cuh, rt:1560) at the energies of the LUT.
void build_tel_trans(int nE, const double* energies, const sart_interp1d_t& tt, double srcEnergy, std::vector<float>* out) {
  out->resize(size_t(nE) + 1);
  for (int i = 0; i <= nE; ++i) (*out)[i] = float(lin1d(tt, i < nE ? std::max(0.03, energies[i]) : srcEnergy));
}

// Sampling cells of one threshold row (fast_params.h: SampleCell); thr[0..n) non-decreasing.
void build_sample_cells(const uint32_t* thr, int n, int bits, SampleCell* out) {
  const int K = 1 << bits, shift = 32 - bits;
  int pos = 0;   // thresholds below the current cell
  for (int k = 0; k < K; ++k) {
    const uint64_t end = (uint64_t(k) + 1) << shift;   // first word of the next cell
    int q = pos;
    while (q < n && uint64_t(thr[q]) < end) ++q;
    int nIn = q - pos;
    const uint32_t thr0 = nIn > 0 ? thr[pos] : 0xffffffffu;
    // the all-ones word compares >= every threshold and >= the "none" marker: it must take the slow path (f64 fallback)
    if (k == K - 1 && (nIn == 0 || thr[q - 1] == 0xffffffffu)) nIn = std::max(nIn, 2);
    out[k] = SampleCell{thr0, uint32_t(pos) | uint32_t(std::min(nIn, 0xffff)) << 16};
    pos = q;
  }
}

// Reflectivity of one coating interpolated along the energy axis at energy E, for every angle node:
// out[i] = z[i][j] + yc (z[i][j+1] - z[i][j]) with (j, yc) the energy cell of the bilinear spline (rt:1567-1578).
void refl_at_energy(const Params& P, const float* z, double E, float* out) {
  const int nA = P.nAngles, nEn = P.nReflEnergies;
  const double y = std::min(std::max(E, P.reflEMin), P.reflEMax);
  const double fy = (y - P.reflEMin) / P.reflDy;
  int j = int(std::floor(fy));
  if (j > nEn - 2) j = nEn - 2;
  if (j < 0) j = 0;
  const float yc = float(fy - double(j));
  for (int i = 0; i < nA; ++i) {
    const float z0 = z[size_t(i) * nEn + j], z1 = z[size_t(i) * nEn + j + 1];
    out[i] = z0 + yc * (z1 - z0);
  }
}

}  // namespace fast
}  // namespace sart
