// derive.cpp — host derivation of the kernel parameter blocks from sart_setup_t.
//
// Each value is produced by the SAME expression (operand order, associativity, host libm) the reference
// evaluates per ray inside traceAxion / findPos* / calcNormalVec (src/raytracer.nim line numbers beside each),
// so hoisting it out of the per-ray path cannot change a result bit. Compile with -ffp-contract=off.
#include <cmath>
#include <cstring>

#include "sart_internal.h"

namespace sart {

static constexpr double kPi = 3.141592653589793;
static constexpr double kRadPerDeg = kPi / 180.0;
static inline double deg2rad(double d) { return d * kRadPerDeg; }
static inline double cot(double x) { return 1.0 / std::tan(x); }

void derive_shells(const sart_setup_t& s, ShellF64* out) {
  const sart_telescope_t& t = s.telescope;
  const double l = t.lMirror;
  const double f = s.detectorInstall.distanceDetectorXRT;
  const double pipeRad = deg2rad(s.pipes.pipesTurned);
  std::memset(out, 0, sizeof(ShellF64) * SART_MAX_SHELLS);
  for (int j = 0; j < t.nShells && j < SART_MAX_SHELLS; ++j) {
    ShellF64& o = out[j];
    const double r1 = t.allR1[j];
    const double beta = deg2rad(t.allAngles[j]);  // rt:1952
    const double xSep = t.allXsep[j];
    const double beta3 = 3.0 * beta;               // rt:1972
    o.R1 = r1;
    o.R1pT = t.allR1[j] + t.allThickness[j];       // rt:1943
    o.r1sq = r1 * r1;
    const double r2 = r1 - l * std::sin(beta);     // rt:1954
    const double r3 = r2 - 0.5 * xSep * std::tan(beta);         // rt:1955
    const double r4 = r3 - 0.5 * xSep * std::tan(3.0 * beta);   // rt:1956
    o.r4 = r4;
    o.r4sq = r4 * r4;
    o.distanceMirrors = std::cos(beta) * (xSep + l);            // rt:1973
    o.distDet = o.distanceMirrors - 0.5 * t.allXsep[8] * std::cos(beta) + s.detectorInstall.distanceDetectorXRT -
                s.detectorInstall.distanceWindowFocalPlane;     // rt:2070-2072
    o.ddWin = o.distDet / std::cos(pipeRad);                    // rt:811
    o.ddEnd = (o.distDet + s.detector.depthDet) / std::cos(pipeRad);  // rt:2081, 811
    // cone rt:639-651
    o.tan1 = std::tan(beta);
    o.k1 = std::tan(beta) * std::tan(beta);
    o.r1tan1 = r1 * std::tan(beta);
    o.two_r1_tan1 = 2.0 * r1 * std::tan(beta);
    o.zmax1 = 0.0 + l * std::cos(beta);
    o.tan2 = std::tan(beta3);
    o.k2 = std::tan(beta3) * std::tan(beta3);
    o.r4tan2 = r4 * std::tan(beta3);
    o.two_r4_tan2 = 2.0 * r4 * std::tan(beta3);
    o.zmax2 = o.distanceMirrors + l * std::cos(beta3);
    // paraboloid rt:669-676 (angle = beta)
    {
      const double tn = std::tan(beta);
      o.p_r3 = -tn * l + std::sqrt(tn * l * tn * l + r1 * r1);
      o.p_e = 2.0 * o.p_r3 * tn;
      o.p_r3sq = o.p_r3 * o.p_r3;
      o.p_el = o.p_e * l;
      o.p_r3tan = o.p_r3 * tn;
      o.p_r3_2tan = o.p_r3 * 2.0 * tn;
    }
    // hyperboloid rt:702-715 (angle = 3 beta; uses r1, quirk Q13)
    {
      const double angle = beta3;
      const double t3 = std::tan(angle / 3.0);
      o.h_r3 = -t3 * l + std::sqrt(t3 * l * t3 * l + r1 * r1);
      const double tn = std::tan(angle);
      o.h_e = 2.0 * o.h_r3 * tn;
      o.h_g = 2.0 * o.h_r3 * tn / (f + o.h_r3 * cot(2.0 * angle / 3.0));
      o.h_r3sq = o.h_r3 * o.h_r3;
      o.h_el = o.h_e * l;
      o.h_gll = o.h_g * l * l;
      o.h_2g = 2.0 * o.h_g;
      const double al = angle / 3.0;                 // rt:751
      o.h_nden = f + o.h_r3 * cot(2.0 * al);          // rt:754
      o.h_r3tan = o.h_r3 * tn;
      o.h_r3_2tan = o.h_r3 * 2.0 * tn;
    }
  }
}

void derive_params(const sart_setup_t& s, const sart_tables_t* tb, Params* p) {
  std::memset(p, 0, sizeof *p);
  const sart_telescope_t& t = s.telescope;
  p->telKind = t.kind;
  p->nShells = t.nShells;
  p->reflKind = t.reflKind;
  p->nCoatings = t.nCoatings;
  p->stage = s.stage;
  p->experiment = s.experiment;
  p->nStripHalf = int(std::round(double(s.detector.numberOfStrips) / 2.0));  // rt:2167
  p->testXray = s.testSource.active;
  p->parallelSource = s.testSource.parallel;
  p->flags = s.flags;
  for (int i = 0; i < SART_MAX_COATINGS; ++i) p->layers[i] = t.layers[i];
  p->holeType = t.holeType;
  p->numberOfHoles = t.numberOfHoles;
  // initCenterVectors rt:278-320
  p->sunX = 0.0;
  p->sunY = -(0.0 * 1.33e10);
  p->sunZ = -s.consts.distanceSunEarth;
  p->radiusSun = s.consts.radiusSun;
  p->radiusCB = s.magnet.radiusCB;
  p->lengthB = s.magnet.lengthB;
  p->zExitCB = s.magnet.lengthColdbore;
  p->zPipe1 = s.magnet.lengthColdbore + s.pipes.cb2vt3_length;
  p->zPipe2 = s.magnet.lengthColdbore + s.pipes.cb2vt3_length + s.pipes.vt3xrt_length;
  p->rPipe1 = s.pipes.cb2vt3_radius;  // used for BOTH pipe clips (rt:1857, 1867)
  p->B = s.magnet.B;
  p->g_agamma = s.consts.g_agamma;
  p->tesla_to_eV2 = s.consts.tesla_to_eV2;
  p->m_to_inv_eV = s.consts.m_to_inv_eV;
  p->pGas = s.magnet.pGasRoom / s.consts.roomTemp * s.magnet.tGas;  // rt:1601
  p->tGas = s.magnet.tGas;
  p->roomTemp = s.consts.roomTemp;
  p->radiusCB_m = s.magnet.radiusCB * 1e-3;
  // telescope frame rt:1879-1894
  const double turnedX = deg2rad(t.telescope_turned_x), turnedY = deg2rad(t.telescope_turned_y);
  p->cosTX = std::cos(turnedX); p->sinTX = std::sin(turnedX);
  p->cosTY = std::cos(turnedY); p->sinTY = std::sin(turnedY);
  const double lengthTelescope = (t.lMirror + 0.5 * t.allXsep[0]) * std::cos(deg2rad(t.allAngles[0])) +
                                 (t.lMirror + 0.5 * t.allXsep[0]) * std::cos(3.0 * deg2rad(t.allAngles[0]));
  p->halfLenTel = lengthTelescope / 2.0;
  p->oeX = t.optics_entrance[0];
  p->oeY = t.optics_entrance[1];
  p->lMirror = t.lMirror;
  p->fL = s.detectorInstall.distanceDetectorXRT;
  p->holeInOptics = t.holeInOptics;
  // detector rt:797-814, 2133-2153
  const double pipeRad = deg2rad(s.pipes.pipesTurned);
  p->cosPipe = std::cos(pipeRad); p->sinPipe = std::sin(pipeRad);
  p->dShift = -t.optics_entrance[0];  // rt:2073
  p->lateralShift = s.detectorInstall.lateralShift;
  p->transversalShift = s.detectorInstall.transversalShift;
  p->radiusWindow = s.detector.radiusWindow;
  p->chipCX = s.consts.chipXMax / 2.0;
  p->chipCY = s.consts.chipYMax / 2.0;
  p->cosTheta = std::cos(s.detector.theta); p->sinTheta = std::sin(s.detector.theta);
  p->stripDist = s.detector.stripDistWindow;
  p->stripWidth = s.detector.stripWidthWindow;
  p->exposureFactor = s.consts.exposureFactor;
  // X-ray source rt:305-311
  p->srcX = s.testSource.offAxisLeft;
  p->srcY = s.testSource.offAxisUp;
  p->srcZ = -s.testSource.distance;
  p->srcRadius = s.testSource.radius;
  p->srcEnergy = s.testSource.energy;
  p->colZ = -s.testSource.distance + s.testSource.lengthCol;
  p->shellsMonotonic = 1;
  for (int j = 1; j < t.nShells; ++j)
    if (!(t.allR1[j] > t.allR1[j - 1]) || !(t.allR1[j - 1] + t.allThickness[j - 1] < t.allR1[j])) p->shellsMonotonic = 0;
  if (tb) {
    p->nAngles = tb->nAngles;
    p->nReflEnergies = tb->nReflEnergies;
    p->angleMin = tb->angleMin; p->angleMax = tb->angleMax;
    p->reflEMin = tb->reflEnergyMin; p->reflEMax = tb->reflEnergyMax;
    p->reflDx = tb->nAngles > 1 ? (tb->angleMax - tb->angleMin) / double(tb->nAngles - 1) : 1.0;
    p->reflDy = tb->nReflEnergies > 1 ? (tb->reflEnergyMax - tb->reflEnergyMin) / double(tb->nReflEnergies - 1) : 1.0;
    p->nRadii = tb->nRadii;
    p->nEnergies = tb->nEnergies;
  }
}

}  // namespace sart
