// host_setup.cpp — host-side constructors of the read-only ray-tracing setup.
//
// C++ replacements for the reference's setup procs (compiled Nim in the reference, so native here too):
//   newExperimentSetup  src/raytracer.nim:1411-1423  (initMagnet :1098-1123, initPipes :1125-1157,
//                                                      initTelescope :1251-1348, initTestXraySource :1350-1379,
//                                                      initDetectorInstallation :1381-1409)
//   newDetectorSetup    src/raytracer.nim:1464-1496, 1528  (the file I/O of :1498-1527 lives in the table loader)
//   calcWindowVals      src/raytracer.nim:1431-1462
//   module globals      src/raytracer.nim:248-272, exposure factors :2207-2212
// All geometry numbers below are the reference's experiment data and must stay verbatim.
#include <cmath>
#include <cstring>
#include <initializer_list>

#include "sart_internal.h"

namespace {

constexpr double kPi = 3.141592653589793;

void fill(double* dst, std::initializer_list<double> v) {
  int i = 0;
  for (double x : v) dst[i++] = x;
}

// ---- magnets (rt:1103-1123)
int init_magnet(int experiment, sart_magnet_t* m) {
  switch (experiment) {
    case SART_ES_CAST:  // CAST LHC dipole prototype
      *m = {/*lengthColdbore*/ 9756.0, /*B*/ 9.0, /*lengthB*/ 9260.0, /*radiusCB*/ 21.5, /*pGasRoom*/ 1.0, /*tGas*/ 1.7};
      return SART_OK;
    case SART_ES_BABYIAXO:
      *m = {11300.0, 2.0, 11000.0, 500.0, 1.0, 100.0};
      return SART_OK;
  }
  return sart::fail(SART_ERR_CONFIG, "unknown experiment kind %d", experiment);
}

// ---- beam pipes between cold bore, VT3 and the telescope (rt:1125-1157)
int init_pipes(int telescope, sart_pipes_t* p) {
  switch (telescope) {
    case SART_TK_LLNL:
      *p = {127.66, 39.89, 111.7, 23.935, 0.0, 2.75};
      return SART_OK;
    case SART_TK_ABRIXAS:
      *p = {114.3, 66.65, 171.43, 47.62, 0.0, 0.0};
      return SART_OK;
    case SART_TK_CUSTOM_BABYIAXO:
    case SART_TK_XMM:
      *p = {225.0, 370.0, 250.0, 370.0, 0.0, 0.0};
      return SART_OK;
  }
  return sart::fail(SART_ERR_CONFIG, "Invalid telescope!");
}

// ---- telescopes (rt:1251-1348); reflectivity kinds from initReflectivity (rt:1160-1249)
int init_telescope(int kind, sart_telescope_t* t) {
  std::memset(t, 0, sizeof *t);
  t->kind = kind;
  switch (kind) {
    case SART_TK_LLNL: {
      t->nShells = 14;
      fill(t->optics_entrance, {-83.0, 0.0, 0.0});
      fill(t->optics_exit, {-83.0, 0.0, 454.0});
      for (int i = 0; i < 14; ++i) t->allThickness[i] = 0.2;
      fill(t->allR1, {63.006, 65.606, 68.305, 71.105, 74.011, 77.027, 80.157, 83.405, 86.775, 90.272, 93.902,
                      97.668, 101.576, 105.632});
      fill(t->allXsep, {4.171, 4.140, 4.221, 4.190, 4.228, 4.245, 4.288, 4.284, 4.306, 4.324, 4.373, 4.387, 4.403,
                        4.481});
      fill(t->allAngles, {0.579, 0.603, 0.628, 0.654, 0.680, 0.708, 0.737, 0.767, 0.798, 0.830, 0.863, 0.898,
                          0.933, 0.970});
      t->lMirror = 225.0;
      t->holeInOptics = 0.0;
      t->numberOfHoles = 5;
      t->holeType = SART_HT_CROSS;
      t->reflKind = SART_RK_MULTI_COATING;
      t->nCoatings = 4;
      t->layers[0] = 2; t->layers[1] = 2 + 3; t->layers[2] = 2 + 3 + 4; t->layers[3] = 2 + 3 + 4 + 5;
      return SART_OK;
    }
    case SART_TK_XMM: {
      t->nShells = 58;
      fill(t->optics_entrance, {0.0, -0.0, 0.0});
      fill(t->optics_exit, {0.0, -0.0, 600.0});
      fill(t->allThickness,
           {0.468, 0.475, 0.482, 0.490, 0.497, 0.504, 0.511, 0.519, 0.526, 0.534, 0.542, 0.549, 0.557, 0.566, 0.574,
            0.583, 0.591, 0.600, 0.609, 0.618, 0.627, 0.636, 0.646, 0.655, 0.665, 0.675, 0.684, 0.694, 0.704, 0.714,
            0.724, 0.735, 0.745, 0.756, 0.768, 0.779, 0.790, 0.802, 0.814, 0.826, 0.838, 0.850, 0.862, 0.874, 0.887,
            0.900, 0.913, 0.927, 0.941, 0.955, 0.968, 0.983, 0.997, 1.011, 1.026, 1.041, 1.055, 1.070});
      fill(t->allR1,
           {153.118,  155.4105, 157.7235, 160.0565, 162.42,   164.803,  167.217,  169.651,  172.115,  174.5995,
            177.1145, 179.6495, 182.2145, 184.9615, 187.739,  190.5465, 193.3845, 196.253,  199.1515, 202.0805,
            205.0395, 208.0795, 211.1495, 214.25,   217.381,  220.542,  223.7435, 226.9755, 230.2375, 233.54,
            236.873,  240.236,  243.6395, 247.2855, 250.9715, 254.6985, 258.4655, 262.2625, 266.1005, 269.9785,
            273.897,  277.856,  281.8555, 285.9055, 289.9955, 294.178,  298.661,  303.0945, 307.5685, 312.093,
            316.658,  321.2735, 325.939,  330.6555, 335.4225, 340.23,   345.0875, 349.996});
      for (int i = 0; i < 58; ++i) t->allXsep[i] = 0.0;
      fill(t->allAngles,
           {0.29,  0.294, 0.298, 0.303, 0.307, 0.312, 0.316, 0.321, 0.325, 0.33,  0.335, 0.34,  0.345, 0.35,  0.355,
            0.36,  0.366, 0.371, 0.377, 0.382, 0.388, 0.393, 0.399, 0.405, 0.411, 0.417, 0.423, 0.429, 0.435, 0.441,
            0.448, 0.454, 0.461, 0.467, 0.474, 0.481, 0.489, 0.496, 0.503, 0.51,  0.518, 0.525, 0.533, 0.54,  0.548,
            0.556, 0.564, 0.573, 0.581, 0.59,  0.598, 0.607, 0.616, 0.625, 0.634, 0.643, 0.652, 0.661});
      t->lMirror = 300.0;
      t->holeInOptics = 0.2;
      t->numberOfHoles = 1;
      t->holeType = SART_HT_NONE;
      t->reflKind = SART_RK_SINGLE_COATING;
      t->nCoatings = 1;
      return SART_OK;
    }
    case SART_TK_ABRIXAS: {
      t->nShells = 27;
      fill(t->optics_entrance, {0.0, -60.0, 0.0});
      fill(t->optics_exit, {0.0, -60.0, 600.0});
      fill(t->allThickness, {0.2,  0.2,  0.2,  0.2,  0.2,  0.2, 0.2, 0.25, 0.25, 0.25, 0.25, 0.25, 0.25, 0.25,
                             0.3,  0.3,  0.3,  0.3,  0.3,  0.35, 0.35, 0.35, 0.35, 0.35, 0.4,  0.4,  0.4});
      fill(t->allR1, {38.125, 39.353, 40.581, 41.809, 43.036, 44.292, 45.577, 46.894, 48.295, 49.731, 51.201, 52.707,
                      54.249, 55.829, 57.447, 59.157, 60.909, 62.703, 64.540, 66.423, 68.403, 70.431, 72.509, 74.637,
                      76.817, 79.102, 81.443});
      for (int i = 0; i < 27; ++i) t->allXsep[i] = 0.0;
      fill(t->allAngles, {0.3335, 0.3443, 0.3550, 0.3657, 0.3765, 0.3874, 0.3987, 0.4102, 0.4225,
                          0.4350, 0.4479, 0.4610, 0.4745, 0.4883, 0.5024, 0.5174, 0.5327, 0.5484,
                          0.5644, 0.5809, 0.5982, 0.6159, 0.6340, 0.6526, 0.6716, 0.6916, 0.7120});
      t->lMirror = 150.0;
      t->holeInOptics = 0.2;
      t->numberOfHoles = 1;
      t->holeType = SART_HT_NONE;
      t->reflKind = SART_RK_SINGLE_COATING;
      t->nCoatings = 1;
      return SART_OK;
    }
    case SART_TK_CUSTOM_BABYIAXO:
      // rt:1232-1234 / rt:1347-1348: not implemented in the reference either.
      return sart::fail(SART_ERR_CONFIG, "The telescope for kind CustomBabyIAXO has not been implemented yet!");
  }
  return sart::fail(SART_ERR_CONFIG, "The telescope for kind %d has not been implemented yet!", kind);
}

// ---- X-ray test source defaults (rt:1350-1379)
void init_test_source(int experiment, uint32_t flags, sart_test_source_t* s) {
  const int active = (flags & SART_CF_XRAY_TEST) ? 1 : 0;
  if (experiment == SART_ES_CAST) {
    *s = {active, /*parallel*/ 1, /*energy*/ 1.0, /*distance*/ 100.0, /*radius*/ 10.0, /*offAxisUp*/ 200.0,
          /*offAxisLeft*/ 0.0, /*activity*/ 1.0, /*lengthCol*/ 50.0};
  } else {
    *s = {active, 1, 0.021, 2000.0, 350.0, 0.0, 0.0, 0.125, 0.0};
  }
}

// ---- where the readout plane sits (rt:1381-1409)
int init_detector_install(int telescope, sart_detector_install_t* d) {
  switch (telescope) {
    case SART_TK_LLNL: *d = {1485.0, 0.0, 0.0, 0.0}; return SART_OK;
    case SART_TK_ABRIXAS: *d = {1600.0, 0.0, 0.0, 0.0}; return SART_OK;
    case SART_TK_XMM:
    case SART_TK_CUSTOM_BABYIAXO: *d = {7500.0, 0.0, 0.0, std::sin(0.0 * (kPi / 180.0)) * 7500.0}; return SART_OK;
  }
  return sart::fail(SART_ERR_CONFIG, "Invalid telescope!");
}

// ---- detector (rt:1464-1496, 1528; window rotation rt:322-332)
int init_detector(int kind, double chipXMax, sart_detector_t* d) {
  std::memset(d, 0, sizeof *d);
  switch (kind) {
    case SART_DK_INGRID2017: d->windowYear = SART_WY_2017; break;
    case SART_DK_INGRID2018: d->windowYear = SART_WY_2018; break;
    case SART_DK_INGRIDIAXO: d->windowYear = SART_WY_IAXO; break;
    default: return sart::fail(SART_ERR_CONFIG, "unknown detector kind %d", kind);
  }
  // identical numbers for all three kinds in the reference
  d->radiusWindow = 7.0;
  d->numberOfStrips = 4;
  d->openApertureRatio = 0.838;
  d->windowThickness = 0.3;
  d->alThickness = 0.02;
  d->depthDet = 30.0;
  d->detectorWindowAperture = chipXMax;
  int rc = sart_calc_window_vals(d->radiusWindow, d->numberOfStrips, d->openApertureRatio, &d->stripWidthWindow,
                                 &d->stripDistWindow);
  if (rc) return rc;
  d->theta = (d->windowYear == SART_WY_IAXO ? 20.0 : 30.0) * (kPi / 180.0);
  return SART_OK;
}

}  // namespace

extern "C" int sart_calc_window_vals(double radiusWindow, int numberOfStrips, double openApertureRatio, double* width,
                                     double* dist) {
  if (!width || !dist || numberOfStrips < 0) return sart::fail(SART_ERR_ARG, "sart_calc_window_vals: bad argument");
  const double totalArea = kPi * radiusWindow * radiusWindow;
  const double areaOfStrips = totalArea * (1.0 - openApertureRatio);
  const double pitch = radiusWindow * 2.0 / (double(numberOfStrips) + 1.0);  // strip width + gap
  double lengthAll = 0.0;
  const int nHalf = int(std::round(double(numberOfStrips) / 2.0));
  for (int i = 0; i < nHalf; ++i) {
    const double off = double(i) * pitch + 0.5 * pitch;  // chord at the strip centre
    lengthAll = lengthAll + std::sqrt(radiusWindow * radiusWindow - off * off) * 2.0;
  }
  lengthAll = lengthAll * 2.0;
  *width = areaOfStrips / lengthAll;
  *dist = pitch - *width;
  return SART_OK;
}

extern "C" int sart_init_setup(int experiment, int detectorKind, int stage, int telescope, uint32_t flags,
                               sart_setup_t* out) {
  if (!out) return sart::fail(SART_ERR_ARG, "sart_init_setup: out is NULL");
  if (stage != SART_SK_VACUUM && stage != SART_SK_GAS) return sart::fail(SART_ERR_CONFIG, "unknown stage kind %d", stage);
  sart_setup_t s;
  std::memset(&s, 0, sizeof s);
  s.abi_version = SART_ABI_VERSION;
  s.flags = flags;
  s.experiment = experiment;
  s.stage = stage;
  s.detectorKind = detectorKind;
  int rc;
  if ((rc = init_magnet(experiment, &s.magnet))) return rc;
  if ((rc = init_telescope(telescope, &s.telescope))) return rc;
  init_test_source(experiment, flags, &s.testSource);
  if ((rc = init_pipes(telescope, &s.pipes))) return rc;
  if ((rc = init_detector_install(telescope, &s.detectorInstall))) return rc;
  // module-level constants rt:248-272
  s.consts.distanceSunEarth = 1.5e14;
  s.consts.radiusSun = 6.9e11;
  s.consts.roomTemp = 293.15;
  s.consts.mAxion = 0.0853;
  s.consts.g_agamma = 1e-12;
  s.consts.chipXMax = 14.0;
  s.consts.chipYMax = 14.0;
  // unchained's natural-unit factors (CODATA): 1 T = 195.353 eV^2, 1 m = 5.06773e6 eV^-1
  s.consts.tesla_to_eV2 = 195.35277121325237;
  s.consts.m_to_inv_eV = 5067730.716548338;
  // rt:2207-2212: three months of tracking; none with the X-ray test source
  if (flags & SART_CF_XRAY_TEST) s.consts.exposureFactor = 1.0;
  else if (experiment == SART_ES_CAST) s.consts.exposureFactor = 3.585e3 * 3600.0 * 1.5 * 90.0;
  else s.consts.exposureFactor = 9.5e6 * 3600.0 * 12.0 * 90.0;
  if ((rc = init_detector(detectorKind, s.consts.chipXMax, &s.detector))) return rc;
  *out = s;
  return SART_OK;
}
