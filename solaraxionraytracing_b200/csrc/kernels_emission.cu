// kernels_emission.cu — opacity-free axion emission rates of the solar model on the GPU (sm_100a).
//
// Replaces the per-(radius, energy) loop of `calculateOpacities` (src/readOpacityFile.nim:776-860) for the processes
// that need no OPCD opacity tables: Primakoff (:384-413), Compton (:360-362), electron-electron bremsstrahlung
// (:364-367), free-free bremsstrahlung (:378-381), the 57Fe line (:454-466) and the longitudinal-plasmon resonance in
// its opacity-free limit (:421-437 with absCoef = 0). FB/BB (`term1`, :369-371) and the transverse plasmon (:439-452)
// are proportional to the opacity and are therefore zero here, as in the reference when no table covers an element.
// The output feeds sart_build_cdfs (rt:2679-2705) directly: emRates[nRadii][nEnergies] row-major.
//
// Two kernels: k_plasma_state (one thread per radius: electron / H / He number densities, Debye scale, the
// table-quantised temperature of :686-690 including its carry-over from the previous radius) and k_emission_rates (one
// thread per table cell). `fNew` (:312-326), which the reference integrates adaptively for every cell, is a fixed
// 3 x 32-node Gauss-Legendre rule in x over [0, 7.5] (the integrand carries exp(-x^2)); nodes live in constant memory.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "kernels.h"

namespace sart {

constexpr int kElem = 29;
constexpr int kGL = 32;
__constant__ double c_glx[kGL], c_glw[kGL];           // nodes / weights on [-1, 1]
__constant__ double c_atomicMass[kElem], c_charges[kElem];

struct PlasmaState {
  double ne_keV, nH, nHe, temp_keV, temp_keVTable, ks2, y, nZZ2, rho_keV, bfield;
};

// readOpacityFile.nim:120-133
static const double h_atomicMass[kElem] = {1.0078, 4.0026, 3.0160, 12.0000, 13.0033, 14.0030, 15.0001, 15.9949, 16.9991,
  17.9991, 20.1797, 22.9897, 24.3055, 26.9815, 28.085, 30.9737, 32.0675, 35.4515, 39.8775, 39.0983, 40.078, 44.9559,
  47.867, 50.9415, 51.9961, 54.9380, 55.845, 58.9331, 58.6934};
static const double h_charges[kElem] = {1, 2, 2, 6, 6, 7, 7, 8, 8, 8, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22,
  23, 24, 25, 26, 27, 28};

__device__ double bfield_keV2(double r) {   // :328-351
  const double radius_cz = 0.712, size_tach = 0.02, radius_outer = 0.96, size_outer = 0.035;
  const double lambda1 = 10.0 * radius_cz + 1.0;
  const double lambda_factor = (1.0 + lambda1) * pow(1.0 + 1.0 / lambda1, lambda1);
  double b = 0.0;
  if (r < radius_cz + size_tach) {
    const double x = (r / radius_cz) * (r / radius_cz);
    if (x < 1.0) b = 3.0e3 * lambda_factor * x * pow(1.0 - x, lambda1);
    const double t = (r - radius_cz) / size_tach, y = t * t;
    if (y < 1.0) b = 50.0 * (1.0 - y);
  } else {
    const double t = (r - radius_outer) / size_outer, z = t * t;
    b = z < 1.0 ? 4.0 * (1.0 - z) : 0.0;
  }
  return b / (1.0e6 * 1.4440271 * 1.0e-3 * sqrt(4.0 * 3.141592653589793));
}

// Table temperature index of :686-690 for one radius, or -1 when no grid value is within one step.
__device__ int table_temperature(double T) {
  int found = -1;
  const double l = log(T) / log(10.0) / 0.025;
  for (int i = 0; i <= 90; ++i)
    if (fabs(l - double(140 + 2 * i)) <= 1.0) found = 140 + 2 * i;
  return found;
}

__global__ void k_plasma_state(int nR, const double* __restrict__ temp, const double* __restrict__ rho,
                               const double* __restrict__ frac, PlasmaState* __restrict__ out) {
  const int R = blockIdx.x * blockDim.x + threadIdx.x;
  if (R >= nR) return;
  const double alpha = 1.0 / 137.0, me = 510.998, amu = 1.6605e-24;
  const double* e = frac + size_t(R) * kElem;
  const double rn = rho[R] / amu;
  PlasmaState s;
  s.nH = (e[0] / c_atomicMass[0]) * rn;
  s.nHe = (e[1] + e[2]) / ((c_atomicMass[1] * e[1] + c_atomicMass[2] * e[2]) / (e[1] + e[2])) * rho[R] / amu;
  double ne = 0.0;
  for (int Z = 0; Z < kElem; ++Z) ne += rn * c_charges[Z] * e[Z] / c_atomicMass[Z];
  // the reference keeps the previous radius' table temperature when none matches: walk back to the last match
  int tt = -1;
  for (int r = R; r >= 0 && tt < 0; --r) tt = table_temperature(temp[r]);
  if (tt < 0) tt = 0;
  s.ne_keV = ne * 7.683e-24;
  s.temp_keVTable = pow(10.0, double(tt) * 0.025) * 8.617e-8;
  s.temp_keV = temp[R] * 8.617e-8;
  s.ks2 = (4.0 * 3.141592653589793 * alpha / s.temp_keV) * (s.ne_keV + s.nH * 7.645e-24 + 4.0 * s.nHe * 7.645e-24);
  s.y = sqrt(s.ks2) / sqrt(2.0 * me * s.temp_keV);
  s.nZZ2 = rn * 7.683e-24;
  s.rho_keV = rho[R] * 7.683e-24 * 5.60958616722e29;
  s.bfield = bfield_keV2(0.0015 + double(R) * 0.0005);
  out[R] = s;
}

__device__ __forceinline__ double inner_integral(double t, double y) {   // :297-298
  const double d = t * t + y * y;
  return 0.5 * (y * y / d + log(d));
}
// fNew(w, y) = int_0^inf x exp(-x^2) [I(sqrt(x^2+w)+x) - I(sqrt(x^2+w)-x)] dx  (:300-326)
__device__ double f_new(double w, double y) {
  const double lo[3] = {0.0, 1.0, 3.0}, hi[3] = {1.0, 3.0, 7.5};
  double sum = 0.0;
  for (int p = 0; p < 3; ++p) {
    const double c = 0.5 * (lo[p] + hi[p]), h = 0.5 * (hi[p] - lo[p]);
    double acc = 0.0;
    for (int k = 0; k < kGL; ++k) {
      const double x = fma(h, c_glx[k], c);
      const double r = sqrt(fma(x, x, w));
      const double to = r + x, frm = w / to;   // sqrt(x^2+w) - x without the cancellation
      acc = fma(c_glw[k], x * exp(-x * x) * (inner_integral(to, y) - inner_integral(frm, y)), acc);
    }
    sum = fma(h, acc, sum);
  }
  return sum;
}

__device__ double primakoff_rate(const PlasmaState& s, double E, double gagamma) {   // :384-413
  const double alpha = 1.0 / 137.0, me = 510.998;
  const double omPlSq = 4.0 * alpha * 3.141592653589793 * s.ne_keV / me;
  const double om2 = E * E, x = om2 / omPlSq;
  if (x < 1.0 || E == 0.0) return 0.0;
  const double phase = 2.0 / (sqrt(1.0 - 1.0 / x) * (exp(E / s.temp_keV) - 1.0));
  const double n_dens = s.ne_keV + s.nH * 7.645e-24 + 4.0 * s.nHe * 7.645e-24;
  const double q = 2.0 * E * sqrt(om2 - omPlSq);
  const double t = s.ks2 / q, u = (2.0 * om2 - omPlSq) / q;
  double a = 0.0;
  if (u > 1.0) a += (u * u - 1.0) * log((u - 1.0) / (u + 1.0));
  const double v = u + t;
  if (v > 1.0) a -= (v * v - 1.0) * log((v - 1.0) / (v + 1.0));
  a = a * (0.5 / t) - 1.0;
  return gagamma * gagamma * 1e-12 * alpha / 8.0 * phase * n_dens * a;
}

__global__ void k_emission_rates(int nR, int nE, const PlasmaState* __restrict__ st, const double* __restrict__ energies,
                                 unsigned processes, double gae, double gagamma, double ganuclei,
                                 double* __restrict__ emRates) {
  const size_t cell = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (cell >= size_t(nR) * nE) return;
  const int R = int(cell / nE), iE = int(cell % nE);
  const PlasmaState s = st[R];
  const double E = energies[iE];
  const double alpha = 1.0 / 137.0, me = 510.998, pi = 3.141592653589793;
  const double w = E / s.temp_keVTable;
  double total = 0.0;
  if (processes & SART_EM_COMPTON)
    total += (alpha * gae * gae * E * E * s.ne_keV) / (3.0 * (me * me * me * me) * (exp(E / s.temp_keV) - 1.0));
  if (processes & SART_EM_EE_BREMS)
    total += (alpha * alpha * gae * gae * 4.0 * sqrt(pi) * s.ne_keV * s.ne_keV * exp(-E / s.temp_keV) *
              f_new(w, sqrt(2.0) * s.y)) / (3.0 * sqrt(s.temp_keV) * pow(me, 3.5) * E);
  if (processes & SART_EM_FREE_FREE)
    total += (f_new(w, s.y) * alpha * alpha * gae * gae * 8.0 * sqrt(pi) * s.ne_keV * s.nZZ2 * exp(-E / s.temp_keV)) /
             (3.0 * sqrt(2.0 * s.temp_keV) * pow(me, 3.5) * E);
  if (processes & SART_EM_PRIMAKOFF) total += primakoff_rate(s, E, gagamma);
  if (processes & SART_EM_LONG_PLASMON) {   // :421-437 with opacity 0 => gammaL = 1e-4
    const double omPlSq = 4.0 * alpha * pi * s.ne_keV / me, om2 = E * E, xi2 = 1e-4 * E;
    const double fwhm = sqrt(om2 + xi2) - sqrt(om2 - xi2);
    if (!(fabs(E - sqrt(omPlSq)) > 18.0 * fwhm)) {
      const double d = om2 - omPlSq;
      total += gagamma * gagamma * 1e-12 * (s.bfield * s.bfield / 3.0) * (E * xi2 / (d * d + xi2 * xi2)) /
               (exp(E / s.temp_keV) - 1.0);
    }
  }
  if (processes & SART_EM_IRON57) {   // :454-466
    const double tau_gamma = 1.3e-6 * 1.519e18, n = 3.0e17 * 1.7826e-30, e_gamma = 14.4;
    const double m_Fe = 56.9353928 * 1.6605e-24 * 5.60958616722e29;
    const double eu = exp(-e_gamma / s.temp_keV);
    const double w_1 = 4.0 * eu / (2.0 + 4.0 * eu);
    const double sigma = e_gamma * sqrt(s.temp_keV / m_Fe);
    const double n_a = n * w_1 * (1.82 * ganuclei * ganuclei) / tau_gamma;
    const double d = E - e_gamma;
    total += n_a * exp(-(d * d) / (2.0 * sigma * sigma)) * s.rho_keV * sqrt(2.0 * pi) * pi / (sigma * E * E);
  }
  emRates[cell] = total;
}

// Gauss-Legendre nodes on [-1, 1] by Newton iteration on P_n.
static void gauss_legendre(int n, double* x, double* w) {
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double z = std::cos(3.141592653589793 * (i + 0.75) / (n + 0.5)), pp = 0.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double dz = p1 / pp;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    x[i] = -z; x[n - 1 - i] = z;
    w[i] = w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
}

cudaError_t launch_emission_rates(int nR, int nE, const double* dTemp, const double* dRho, const double* dFrac,
                                  const double* dEnergies, unsigned processes, double gae, double gagamma, double ganuclei,
                                  void* dState, double* dEmRates, cudaStream_t s) {
  static_assert(sizeof(PlasmaState) == 80, "state record");
  double x[kGL], w[kGL];
  gauss_legendre(kGL, x, w);
  cudaError_t e;
  if ((e = cudaMemcpyToSymbolAsync(c_glx, x, sizeof x, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(c_glw, w, sizeof w, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(c_atomicMass, h_atomicMass, sizeof h_atomicMass, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbolAsync(c_charges, h_charges, sizeof h_charges, 0, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
  PlasmaState* st = static_cast<PlasmaState*>(dState);
  k_plasma_state<<<(nR + 127) / 128, 128, 0, s>>>(nR, dTemp, dRho, dFrac, st);
  const size_t cells = size_t(nR) * nE;
  k_emission_rates<<<unsigned((cells + 127) / 128), 128, 0, s>>>(nR, nE, st, dEnergies, processes, gae, gagamma, ganuclei, dEmRates);
  return cudaGetLastError();
}

}  // namespace sart
