// fast_params.h — POD blocks of the "fast" pipeline (kernels_fast.cu), filled on the host by derive_fast.cpp.
#pragma once
#include <cstdint>

#include <vector_types.h>

#include "device_params.h"

namespace sart {
namespace fast {

// Shell record in shared memory (built from ShellF64 on the host, see derive_fast_shells).
struct ShellFast {
  double R1, R1pT, r1sq;
  double tan1, zmax1, cosb, sinb;
  double r4, tan2, dm, zmax2, cos3b, sin3b;
  double ddWin, distDet;
  // Wolter-I
  double p_e, p_c0, p_r3sq, p_r3tan;                       // paraboloid: rho^2 = r3^2 + e (l - z); c0 = r3^2 + e l
  double h_e, h_g, h_r3sq, h_r3tan, h_inv_nden;            // hyperboloid: rho^2 = r3^2 + e (l-z) + g (l-z)^2
  int32_t coat;   // reflectivity table of this shell: layers.lowerBound(hitLayer) (rt:1573)
  int32_t pad_;
};
static_assert(sizeof(ShellFast) % 16 == 8, "odd number of doubles keeps shared-memory rows off the same banks");

struct FastParams {
  // geometry (FP64)
  double radiusCB2, lengthB, dzExitCB, dzPipe1, dzPipe2, rPipe12;  // dz* = plane z - lengthB
  double cosTX, sinTX, cosTY, sinTY, halfLenTel, oeX, oeY, zExitCBtel;  // zExitCBtel = zExitCB - zPipe2
  double lMirror, cosPipe, sinPipe, dShift, lateralShift, transversalShift;
  double radiusWindow2, chipCX, chipCY, cosTheta, sinTheta, stripDist, stripWidth, invStripPitch, invBinX, invBinY;
  double sunDist, radiusSun, radiusCB;
  double shellRhoMin, shellInvStep;   // uniform radial grid -> first candidate shell (shellGuide)
  // weights (FP32)
  float convK;         // (g*1e-9 * B*T2eV2 * 1e-3*m2eV / 2)^2: conversionProb = convK * pathCB^2 (rt:363-365)
  float exposure;
  float angleMin, angleMax, invReflDx;
  // gas stage (FP64 constants of axionMassforMagnet.nim)
  double gasGamma0, gasMgamma2, gasTerm1, gasRhoPipe100, gasRhoMagnet100;
  // X-ray source
  double srcX, srcY, srcZ, srcRadius, colDz, srcRadius2;
  float srcEnergy;
  double enE0, enInvStep;   // uniform-grid guess of an energy's index in the tabulated energies (pre-sampled rays)
  int32_t telKind, nShells, reflKind, nCoatings, stage, nStripHalf, testXray, parallelSource;
  int32_t layers[SART_MAX_COATINGS];
  uint32_t flags;
  int32_t nRadii, nEnergies, nAngles, nReflEnergies, shellsMonotonic, srcEIdx;
  int32_t nShellGuide, rotated;
};

// ---- single-precision pipeline (kernels_f32.cu): the same blocks rounded to FP32, plus a few derived values that
// keep its arithmetic free of cancellation (radii instead of squared radii: C = (rho - R)(rho + R))
struct ShellF32 {
  float R1, R1pT, tan1, zmax1, cosb, sinb;
  float r4, tan2, dm, zmax2, cos3b, sin3b;
  float ddWin;
  float p_e, p_R0, p_r3sq, p_r3tan;                        // p_R0 = sqrt(r3^2 + e l): paraboloid radius at z = 0
  float h_e, h_g, h_r3sq, h_r3tan, h_inv_nden;
  int32_t coat;
};
static_assert(sizeof(ShellF32) == 92, "23 words: an odd stride keeps shared-memory rows off the same banks");

struct Geo32 {
  float radiusCB, radiusCB2, lengthB, lengthB2, lengthBplusSun, radiusSun;
  float dzExitCB, dzPipe1, dzPipe2, rPipe12;
  float cosTX, sinTX, cosTY, sinTY, halfLenTel, oeX, oeY, zExitCBtel;
  float lMirror, cosPipe, sinPipe, dShift, lateralShift, transversalShift;
  float radiusWindow2, chipCX, chipCY, cosTheta, sinTheta, stripDist, stripWidth, invStripPitch, invBinX, invBinY;
  float shellRhoMin, shellInvStep;
  float srcX, srcY, srcRadius, srcRadius2, invSrcDz, colDz;
};

// Radial lookup of the shell search (rt:1932-1957) for the FP32 kernels: the radial buckets of the shell guide are fine
// enough to hold at most one boundary (a shell's front radius R1 or the outer edge R1 + thickness of its glass) each,
// so one 8-byte record decides the outcome: sel holds the outcome below / at / above the boundary B, one byte each
// (bits 0-7, 8-15, 16-23): a shell number < 64, or 64 + exit code.
struct ShellCell {
  float B;
  uint32_t sel;
};
constexpr int kShellCellFail = 64;

struct EnergyLUT {  // one record per tabulated energy index (16 B, one LDG.128)
  float E, Twindow, Tstrongback, Agas;
};
struct GasLUT {     // buffer-gas stage only
  float massAtt;   // exp(logMassAttenuation(E)) am:70-73
  float inv2E;     // 1 / (2 E[eV])
};

struct FastTables {
  // Inverse-CDF sampling (rt:437, 464) in integers: with u = (w + 0.5) 2^-32, cdf[i] < u  <=>  w >= thr[i], where
  // thr[i] is the smallest such 32-bit word (saturated at 0xffffffff; that word falls back to the f64 tables).
  const double* radiusCDF;       // [nRadii] f64, fallback only
  const uint32_t* radiusThr;     // [thr_pitch(nRadii)]
  const uint16_t* radiusGuide;   // [kRadGuide] g[k] = lowerBound(cdf, k/kRadGuide): the search for a word of bucket k starts there
  const double* energyCDF;       // [nRadii][nEnergies] f64, fallback only
  const uint32_t* energyThr;     // [nRadii][thr_pitch(nEnergies)]
  const uint16_t* energyGuide;   // [nRadii][kEnGuide]
  const double* energies;        // [nEnergies] keV (pre-sampled rays: energy -> index)
  const EnergyLUT* elut;         // [nEnergies + 1]
  const GasLUT* glut;            // [nEnergies + 1]
  // reflectivity pre-interpolated along the energy axis at every tabulated energy: [coat][nEnergies + 1][nAngles];
  // the per-ray lookup is then linear in the grazing angle only (same value as the bilinear form of rt:1567-1578)
  const float* reflE;
  const ShellFast* shells;    // [nShells]
  const ShellF32* shells32;   // [nShells] the same records in single precision (kernels_f32.cu)
  const uint8_t* shellGuide;  // [nShellGuide]: smallest j with R1[j] > lower edge of the radial bucket
  const ShellCell* shellTab;  // [nShellGuide] (kernels_f32.cu)
  // Alias tables (sart_set_sampler(SART_SAMPLER_ALIAS), kernels_f32.cu): the same discrete distributions as the
  // thresholds above — P(i) = (thr[i] - thr[i-1]) 2^-32 — as Walker/Vose tables, one 32-bit entry per index:
  // bits 31..11 = the bucket's own share in units of 2^-21, bits 10..0 = the alias index. A word w selects bucket
  // k = (w n) >> 32 and keeps it when the low 32 bits of w n are below the share, else takes the alias: one lookup, no
  // search. Needs n <= 2048; null when unavailable.
  const uint32_t* radiusAlias;   // [nRadii]
  const uint32_t* energyAlias;   // [nRadii][nEnergies]
  int32_t sampler, pad2_;        // SART_SAMPLER_* for this launch
  RadialHist rad;             // optional (w == nullptr: off)
  // Image replicas: block b adds into replica b % nImgRep of the image / w^2 image it is handed (replica r starts
  // imgRepStride doubles after replica 0); 0 or 1 = the image itself. The focal spot concentrates ~1e9 atomic adds per
  // launch on a few hundred bins; spreading them over replicas removes the same-address serialisation in L2.
  int32_t nImgRep, pad_;
  uint64_t imgRepStride;
};

}  // namespace fast
}  // namespace sart
