/* oracle_emission.c — CPU restatement (TEST INFRASTRUCTURE ONLY) of the opacity-free part of the reference's
 * emission-table generator, src/readOpacityFile.nim `calculateOpacities` (:598-860), written in the reference's own
 * operation order. Parity status: PINNED ONLY ON FORMULAS — the Nim program cannot be built here, its output file
 * (resources/solar_model_dataframe.csv) is not shipped, and `fNew` (:312-326) integrates with numericalnim's
 * `adaptiveGauss` (source absent; documented as adaptive Gauss-Kronrod G10K21, tol 1e-8) — restated below as a
 * recursive adaptive G10K21 with a 1e-10 tolerance, so the two bremsstrahlung terms are "parity unpinned" below 1e-8.
 *
 * What is restated: the per-radius plasma state (:659-700, :788-812), primakoff (:384-413), comptonEmrate (:360-362),
 * bremsEmrate (:364-367), freefreeEmrate (:378-381), fNew/outer/inner_integral (:297-326), iron (:454-466),
 * longPlasmon (:421-437) and bfield (:328-351). What is NOT: term1 (FB/BB, :369-371) and transPlasmon (:439-452)
 * need the OPCD monochromatic opacities, which are not in the tree; they are evaluated with absCoef = 0 exactly as the
 * reference would for an element list without tables (term1 = 0, transPlasmon = 0).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define N_ELEM 29
/* readOpacityFile.nim:120-133 */
static const double atomicMass[N_ELEM] = {1.0078, 4.0026, 3.0160, 12.0000, 13.0033, 14.0030, 15.0001, 15.9949, 16.9991,
  17.9991, 20.1797, 22.9897, 24.3055, 26.9815, 28.085, 30.9737, 32.0675, 35.4515, 39.8775, 39.0983, 40.078, 44.9559,
  47.867, 50.9415, 51.9961, 54.9380, 55.845, 58.9331, 58.6934};
static const double charges[N_ELEM] = {1.0, 2.0, 2.0, 6.0, 6.0, 7.0, 7.0, 8.0, 8.0, 8.0, 10.0, 11.0, 12.0, 13.0, 14.0,
  15.0, 16.0, 17.0, 18.0, 19.0, 20.0, 21.0, 22.0, 23.0, 24.0, 25.0, 26.0, 27.0, 28.0};

enum { EM_PRIMAKOFF = 1, EM_COMPTON = 2, EM_EE_BREMS = 4, EM_FREE_FREE = 8, EM_IRON57 = 16, EM_LONG_PLASMON = 32 };

/* ---- adaptive Gauss-Kronrod G10K21 (stand-in for numericalnim adaptiveGauss) ---- */
static const double xgk[11] = {0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
  0.930157491355708226001207180059508, 0.865063366688984510732096688423493, 0.780817726586416897063717578345042,
  0.679409568299024406234327365114874, 0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
  0.294392862701460198131126603103866, 0.148874338981631210884826001129720, 0.0};
static const double wgk[11] = {0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
  0.054755896574351996031381300244580, 0.075039674810919952767043140916190, 0.093125454583697605535065465083366,
  0.109387158802297641899210590325805, 0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
  0.142775938577060080797094273138717, 0.147739104901338491374841515972068, 0.149445554002916905664936468389821};
static const double wg[5] = {0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
  0.219086362515982043995534934228163, 0.269266719309996355091226921569469, 0.295524224714752870173815619188769};

typedef double (*fn1)(double, const double*);
static void gk21(fn1 f, const double* ctx, double a, double b, double* result, double* err) {
  const double c = 0.5 * (a + b), h = 0.5 * (b - a);
  const double fc = f(c, ctx);
  double rk = fc * wgk[10], rg = 0.0;
  for (int j = 0; j < 5; ++j) {   /* Gauss nodes are the odd Kronrod ones */
    const int k = 2 * j + 1;
    const double d = h * xgk[k], s = f(c - d, ctx) + f(c + d, ctx);
    rg += wg[j] * s; rk += wgk[k] * s;
  }
  for (int j = 0; j < 5; ++j) {
    const int k = 2 * j;
    const double d = h * xgk[k];
    rk += wgk[k] * (f(c - d, ctx) + f(c + d, ctx));
  }
  *result = rk * h;
  *err = fabs((rk - rg) * h);
}
static double adapt(fn1 f, const double* ctx, double a, double b, double tol, int depth) {
  double r, e;
  gk21(f, ctx, a, b, &r, &e);
  if (depth >= 40 || e <= tol * fmax(1.0, fabs(r)) * 1e-2 || e <= 1e-300) return r;
  const double m = 0.5 * (a + b);
  return adapt(f, ctx, a, m, tol, depth + 1) + adapt(f, ctx, m, b, tol, depth + 1);
}

/* ---- readOpacityFile.nim:297-326 ---- */
static double inner_integral(double t, double y) { return (1.0 / 2.0) * (((y * y) / (t * t + y * y)) + log(t * t + y * y)); }
static double outer_(double x, double w, double y) {
  const double coeff = x * exp(-x * x);
  const double frm = sqrt(x * x + w) - x, to = sqrt(x * x + w) + x;
  return coeff * (inner_integral(to, y) - inner_integral(frm, y));
}
static double fnToInt(double t, const double* ctx) {
  if (t != 0) return outer_((1 - t) / t, ctx[0], ctx[1]) / (t * t);
  return outer_((1 - t) / (t + 1e-8), ctx[0], ctx[1]) / (t * t);
}
double oracle_fNew(double w, double y) {
  const double ctx[2] = {w, y};
  return adapt(fnToInt, ctx, 0.0, 1.0, 1e-10, 0);
}

/* ---- readOpacityFile.nim:328-351 ---- */
double oracle_bfield(double r) {
  const double radius_cz = 0.712, size_tach = 0.02, radius_outer = 0.96, size_outer = 0.035;
  const double bfield_rad_T = 3.0e3, bfield_tach_T = 50.0, bfield_outer_T = 4.0;
  const double lambda1 = 10.0 * radius_cz + 1.0;
  const double lambda_factor = (1.0 + lambda1) * pow(1.0 + 1.0 / lambda1, lambda1);
  double b = 0.0;
  if (r < (radius_cz + size_tach)) {
    const double x = pow(r / radius_cz, 2.0);
    if (x < 1.0) b = bfield_rad_T * lambda_factor * x * pow(1.0 - x, lambda1);
    const double y = pow(((r - radius_cz) / size_tach), 2.0);
    if (y < 1.0) b = bfield_tach_T * (1.0 - y);
  } else {
    const double z = pow((r - radius_outer) / size_outer, 2.0);
    if (z < 1.0) b = bfield_outer_T * (1.0 - z); else b = 0.0;
  }
  return b / (1.0e6 * 1.4440271 * 1.0e-3 * sqrt(4.0 * M_PI));
}

static double omegaPlasmonSq(double alpha, double ne, double me) { return 4.0 * alpha * M_PI * ne / me; }
static double comptonEmrate(double alpha, double gae, double energy, double ne, double me, double temp) {
  return (alpha * gae * gae * energy * energy * ne) / (3.0 * pow(me, 4) * (exp(energy / temp) - 1.0));
}
static double bremsEmrate(double alpha, double gae, double energy, double ne, double me, double temp, double w, double y) {
  return (alpha * alpha * gae * gae * 4.0 * sqrt(M_PI) * ne * ne * exp(-energy / temp) * oracle_fNew(w, sqrt(2.0) * y)) /
         (3.0 * sqrt(temp) * pow(me, 3.5) * energy);
}
static double freefreeEmrate(double alpha, double gae, double energy, double ne, double me, double temp, double nzZ2,
                             double w, double y) {
  return (oracle_fNew(w, y) * alpha * alpha * gae * gae * 8.0 * sqrt(M_PI) * ne * nzZ2 * exp(-energy / temp)) /
         (3.0 * sqrt(2.0 * temp) * pow(me, 3.5) * energy);
}
static double primakoff_bracket(double t, double u) {
  double a = 0.0;
  if (u > 1.0) a += (u * u - 1.0) * log((u - 1.0) / (u + 1.0));
  const double v = u + t;
  if (v > 1.0) a -= (v * v - 1.0) * log((v - 1.0) / (v + 1.0));
  a *= 0.5 / t;
  a -= 1.0;
  return a;
}
double oracle_primakoff(double temp, double energy, double gagamma, double ks2, double alpha, double ne, double me,
                        double n_Z2, double n_Z1) {
  const double prefactor6 = gagamma * gagamma * 1e-12 * alpha / 8.0;
  const double omPlSq = omegaPlasmonSq(alpha, ne, me);
  const double z = energy / temp, om2 = energy * energy, x = om2 / omPlSq;
  if (x < 1.0 || energy == 0.0) return 0.0;
  const double phase_factor = 2.0 / (sqrt(1.0 - 1.0 / x) * (exp(z) - 1.0));
  const double n_dens = ne + n_Z1 * 7.645e-24 + 4.0 * n_Z2 * 7.645e-24;
  const double s = 2.0 * energy * sqrt(om2 - omPlSq);
  const double t = ks2 / s, u = (2.0 * om2 - omPlSq) / s;
  return prefactor6 * phase_factor * n_dens * primakoff_bracket(t, u);
}
static double longPlasmon(double energy, double ne, double me, double alpha, double bfieldR, double temp, double opacity,
                          double gagamma) {
  const double omPlSq = omegaPlasmonSq(alpha, ne, me), prefactor = gagamma * gagamma * 1e-12, om2 = energy * energy,
               z = energy / temp;
  double gammaL = (1.0 - exp(-z)) * opacity;
  gammaL = fmax(gammaL, 1e-4);
  const double xi2 = gammaL * energy;
  const double fwhm = sqrt(om2 + xi2) - sqrt(om2 - xi2);
  if (fabs(energy - sqrt(omPlSq)) > 18.0 * fwhm) return 0;
  const double average_bfield_sq = bfieldR * bfieldR / 3.0;
  const double fraction = energy * xi2 / (pow(om2 - omPlSq, 2.0) + xi2 * xi2);
  return prefactor * average_bfield_sq * fraction / (exp(z) - 1.0);
}
static double iron(double ganuclei, double temp, double energy, double rho) {
  const double tau_gamma = 1.3e-6 * 1.519e18, n = 3.0e17 * 1.7826e-30, e_gamma = 14.4;
  const double m_Fe = 56.9353928 * 1.6605e-24 * 5.60958616722e29;
  const double u = e_gamma / temp;
  const double w_1 = 4.0 * exp(-u) / (2.0 + 4.0 * exp(-u));
  const double gamma_frac = 1.82 * ganuclei * ganuclei;
  const double sigma = e_gamma * sqrt(temp / m_Fe);
  const double n_a = n * w_1 * gamma_frac / tau_gamma;
  return n_a * exp(-pow(energy - e_gamma, 2.0) / (2.0 * sigma * sigma)) * rho * sqrt(2.0 * M_PI) * M_PI / (sigma * energy * energy);
}

/* Emission rates emRates[nRadii][nEnergies] from the AGSS09 columns temp [K], rho [g/cm^3] and the 29 mass fractions
 * [nRadii][29] in the file's column order (H1, He4, He3, C12 ... Ni). radius(R) = 0.0015 + 0.0005 R (:793). */
int oracle_emission_rates(int nRadii, const double* temp, const double* rho, const double* frac, int nEnergies,
                          const double* energies, uint32_t processes, double g_ae, double gagamma, double ganuclei,
                          double* emRates) {
  const double alpha = 1.0 / 137.0, m_e_keV = 510.998, amu = 1.6605e-24;
  int temperature = 0; /* carried from the previous radius when no table temperature matches (:686-690) */
  for (int R = 0; R < nRadii; ++R) {
    const double* e = frac + (size_t)R * N_ELEM;
    const double nH = (e[0] / atomicMass[0]) * (rho[R] / amu);                      /* n_Z[1] :664 */
    const double nHe = (e[1] + e[2]) / ((atomicMass[1] * e[1] + atomicMass[2] * e[2]) / (e[1] + e[2])) * rho[R] / amu; /* n_Z[2] :667-672 */
    double n_e = 0.0;
    for (int Z = 0; Z < N_ELEM; ++Z) n_e += (rho[R] / amu) * charges[Z] * e[Z] / atomicMass[Z];   /* :683-684 */
    for (int iTemp = 0; iTemp <= 90; ++iTemp) {
      const double distTemp = (log(temp[R]) / log(10.0)) / 0.025 - (double)(140 + 2 * iTemp);
      if (fabs(distTemp) <= 1.0) temperature = 140 + 2 * iTemp;
    }
    const double n_e_keV = n_e * 7.683e-24;                                        /* :790, :797 */
    const double radius = 0.0015 + (double)R * 0.0005;
    const double bfieldR = oracle_bfield(radius);
    const double rho_keV = rho[R] * 7.683e-24 * 5.60958616722e29;
    const double temp_keVTable = pow(10.0, ((double)temperature * 0.025)) * 8.617e-8;
    const double temp_keV = temp[R] * 8.617e-8;
    const double debye_scale_squared = (4.0 * M_PI * alpha / temp_keV) * (n_e_keV + nH * 7.645e-24 + 4.0 * nHe * 7.645e-24);
    const double debye_scale = sqrt(debye_scale_squared);
    const double y = debye_scale / (sqrt(2.0 * m_e_keV * temp_keV));
    const double nZZ2 = (rho[R] / amu) * 7.683e-24;
    for (int iE = 0; iE < nEnergies; ++iE) {
      const double E = energies[iE];
      const double w = E / temp_keVTable;
      double total = 0.0;
      /* the reference's order of summation: compton + term1 + term3 + ffterm + transPlas + primakoff + longPlas + iron57 */
      if (processes & EM_COMPTON) total += comptonEmrate(alpha, g_ae, E, n_e_keV, m_e_keV, temp_keV);
      if (processes & EM_EE_BREMS) total += bremsEmrate(alpha, g_ae, E, n_e_keV, m_e_keV, temp_keV, w, y);
      if (processes & EM_FREE_FREE) total += freefreeEmrate(alpha, g_ae, E, n_e_keV, m_e_keV, temp_keV, nZZ2, w, y);
      if (processes & EM_PRIMAKOFF) total += oracle_primakoff(temp_keV, E, gagamma, debye_scale_squared, alpha, n_e_keV, m_e_keV, nHe, nH);
      if (processes & EM_LONG_PLASMON) total += longPlasmon(E, n_e_keV, m_e_keV, alpha, bfieldR, temp_keV, 0.0, gagamma);
      if (processes & EM_IRON57) total += iron(ganuclei, temp_keV, E, rho_keV);
      emRates[(size_t)R * nEnergies + iE] = total;
    }
  }
  return 0;
}
