/* abi_driver.c — the C-ABI of libsart.so used from plain C99 (no Python, no C++): what a Nim `{.importc.}` binding
 * or any other FFI sees. Builds a CAST + LLNL setup with the library's constructors, a small synthetic table set,
 * and — when a CUDA device is present — runs the fused Monte Carlo pass, the per-ray drop-in of traceAxionWrapper
 * (src/raytracer.nim:2223-2244) and the angular scan, checking ray conservation. Without a device sart_create must
 * fail with SART_ERR_CUDA (there is no CPU fallback). Exit code 0 = all checks passed.
 *   gcc -std=c99 -Wall -Wextra -Werror -pedantic -I include tests/c/abi_driver.c -L solaraxionraytracing_b200 -lsart -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sart.h"

#define CHECK(cond)                                                                   \
  do {                                                                                \
    if (!(cond)) { fprintf(stderr, "FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, sart_last_error()); return 1; } \
  } while (0)

enum { NR = 60, NE = 80, NA = 64, NEN = 48, NT = 32 };

int main(void) {
  CHECK(sart_abi_version() == SART_ABI_VERSION);
  CHECK(sart_sizeof_setup() == sizeof(sart_setup_t));
  CHECK(sart_sizeof_tables() == sizeof(sart_tables_t));
  CHECK(sart_sizeof_counters() == sizeof(sart_counters_t));

  sart_setup_t setup;
  CHECK(sart_init_setup(SART_ES_CAST, SART_DK_INGRID2018, SART_SK_VACUUM, SART_TK_LLNL, 0, &setup) == SART_OK);
  CHECK(setup.telescope.nShells == 14 && setup.magnet.radiusCB == 21.5 && setup.magnet.B == 9.0); /* rt:1105-1112, 1265 */
  CHECK(sart_init_setup(SART_ES_CAST, SART_DK_INGRID2018, SART_SK_VACUUM, SART_TK_OTHER, 0, &setup) == SART_ERR_CONFIG);
  CHECK(sart_init_setup(SART_ES_CAST, SART_DK_INGRID2018, SART_SK_VACUUM, SART_TK_LLNL, 0, &setup) == SART_OK);
  double width, dist;
  CHECK(sart_calc_window_vals(7.0, 4, 0.838, &width, &dist) == SART_OK);
  CHECK(fabs(width - 0.500418) < 1e-5 && fabs(dist - 2.299582) < 1e-5);      /* calculateWindowValues.nim:7-33 */

  /* tables: a smooth emission spectrum per shell -> CDFs by hand (the device build needs a GPU) */
  static double energies[NE], radiusCDF[NR], cdfs[NR * NE], refl[4 * NA * NEN], tx[NT], ty[NT], ga[NT];
  double racc = 0.0;
  for (int i = 0; i < NE; ++i) energies[i] = 0.05 + 10.0 * i / (NE - 1);
  for (int r = 0; r < NR; ++r) {
    const double rad = 0.0015 + 0.0005 * r * 30, T = 1.3 * exp(-rad / 0.3) + 0.1;
    double acc = 0.0;
    for (int i = 0; i < NE; ++i) {
      acc += exp(-energies[i] / T) * energies[i] * energies[i] * rad * rad;
      cdfs[r * NE + i] = acc;
    }
    for (int i = 0; i < NE; ++i) cdfs[r * NE + i] /= acc;
    racc += acc;
    radiusCDF[r] = racc;
  }
  for (int r = 0; r < NR; ++r) radiusCDF[r] /= racc;
  for (int c = 0; c < 4; ++c)
    for (int a = 0; a < NA; ++a)
      for (int e = 0; e < NEN; ++e) refl[(c * NA + a) * NEN + e] = 0.9 / (1.0 + pow(a * 1.5 / (NA - 1) / 0.6, 6.0));
  for (int i = 0; i < NT; ++i) { tx[i] = 15.0 * i / (NT - 1); ty[i] = 0.2 + 0.7 * i / (NT - 1); ga[i] = 1.0 - 0.5 * i / (NT - 1); }
  sart_tables_t tb;
  memset(&tb, 0, sizeof tb);
  tb.nRadii = NR; tb.nEnergies = NE; tb.energies = energies; tb.fluxRadiusCDF = radiusCDF; tb.diffFluxCDFs = cdfs;
  tb.nCoatings = 4; tb.nAngles = NA; tb.nReflEnergies = NEN; tb.angleMin = 0.0; tb.angleMax = 1.5;
  tb.reflEnergyMin = 0.03; tb.reflEnergyMax = 15.0; tb.reflectivity = refl;
  tb.strongbackTransmission.n = NT; tb.strongbackTransmission.x = tx; tb.strongbackTransmission.y = ty;
  tb.windowTransmission = tb.strongbackTransmission;
  tb.gasAbsorption.n = NT; tb.gasAbsorption.x = tx; tb.gasAbsorption.y = ga;

  uint32_t thr[NE];
  sart_cdf_thresholds(cdfs, NE, thr);
  for (int i = 1; i < NE; ++i) CHECK(thr[i] >= thr[i - 1]);

  sart_handle_t* h = NULL;
  const int rc = sart_create(&setup, &tb, 0, &h);
  if (sart_device_count() < 1) {
    CHECK(rc == SART_ERR_CUDA && h == NULL);
    CHECK(strstr(sart_last_error(), "no CPU fallback") != NULL);
    printf("abi_driver: no CUDA device, sart_create refused as it must (%s)\n", sart_last_error());
    return 0;
  }
  CHECK(rc == SART_OK && h != NULL);
  for (int mode = 0; mode <= 2; ++mode) {
    CHECK(sart_has_precision(mode) == 1 && sart_set_precision(h, mode) == SART_OK);
    const uint64_t n = 200000;
    CHECK(sart_reset_image(h) == SART_OK);
    CHECK(sart_trace_mc(h, 0, n / 2, 1) == SART_OK && sart_trace_mc(h, n / 2, n - n / 2, 1) == SART_OK);
    static double image[SART_IMAGE_BINS * SART_IMAGE_BINS];
    sart_counters_t c;
    CHECK(sart_read_image(h, image, NULL, &c) == SART_OK);
    uint64_t total = 0;
    for (int e = 0; e < SART_N_EXIT_CODES; ++e) total += c.n_exit[e];
    double sum = 0.0;
    for (int i = 0; i < SART_IMAGE_BINS * SART_IMAGE_BINS; ++i) sum += image[i];
    CHECK(c.n_rays == n && total == n && c.n_passed == c.n_exit[SART_EXIT_PASSED] && c.n_passed > n / 2);
    CHECK(fabs(sum / c.sum_w - 1.0) < 1e-9);
    /* per-ray records of the same rays: traceAxionWrapper's Axion buffer as a structure of arrays */
    enum { M = 4096 };
    static double x[M], y[M], w[M];
    static int32_t code[M], shell[M];
    sart_ray_out_t out;
    memset(&out, 0, sizeof out);
    out.x = x; out.y = y; out.w = w; out.code = code; out.shell = shell;
    CHECK(sart_trace_mc_rays(h, 0, M, 1, &out) == SART_OK);
    int passed = 0;
    for (int i = 0; i < M; ++i) passed += (code[i] & SART_CODE_MASK) == SART_EXIT_PASSED && w[i] > 0.0 && shell[i] >= 0;
    CHECK(passed > M / 2);
    double ang[3] = {0.0, 0.1, 0.3}, flux[3];
    CHECK(sart_angular_scan(h, 3, ang, 0, 50000, 1, flux, NULL, NULL) == SART_OK);
    CHECK(flux[0] > 0.0 && flux[0] > flux[2]);
    printf("abi_driver: mode %d: %llu rays, %llu passed, total flux %.6e, scan %.3e %.3e %.3e\n", mode,
           (unsigned long long)c.n_rays, (unsigned long long)c.n_passed, c.sum_w, flux[0], flux[1], flux[2]);
  }
  /* ---- the collective: one process, one handle per visible GPU (up to 8), each traces its contiguous shard of the
   * global ray range, sart_allreduce sums image | sum-of-squares image | counters into every handle's merged buffer.
   * The merged result must equal a single-GPU run over the whole range (integer counters exactly). */
  {
    enum { MAXG = 8 };
    const int ng = sart_device_count() < MAXG ? sart_device_count() : MAXG;
    const uint64_t n = 400000, seed = 7;
    static double img1[SART_IMAGE_BINS * SART_IMAGE_BINS], imgN[SART_IMAGE_BINS * SART_IMAGE_BINS];
    sart_counters_t c1, cN;
    CHECK(sart_set_precision(h, 2) == SART_OK && sart_reset_image(h) == SART_OK);
    CHECK(sart_trace_mc(h, 0, n, seed) == SART_OK && sart_read_image(h, img1, NULL, &c1) == SART_OK);
    sart_handle_t* hs[MAXG];
    hs[0] = h;
    for (int g = 1; g < ng; ++g) {
      CHECK(sart_create(&setup, &tb, g, &hs[g]) == SART_OK);
      CHECK(sart_set_precision(hs[g], 2) == SART_OK);
    }
    const int rcc = sart_comm_init_all(hs, ng);
    if (rcc == SART_ERR_CONFIG) {
      printf("abi_driver: NCCL not available (%s): collective skipped\n", sart_last_error());
    } else {
      CHECK(rcc == SART_OK);
      for (int g = 0; g < ng; ++g) {
        const uint64_t lo = n * (uint64_t)g / (uint64_t)ng, hi = n * (uint64_t)(g + 1) / (uint64_t)ng;
        CHECK(sart_reset_image(hs[g]) == SART_OK && sart_trace_mc(hs[g], lo, hi - lo, seed) == SART_OK);
      }
      CHECK(sart_allreduce(hs, ng) == SART_OK);
      for (int g = 0; g < ng; ++g) {
        CHECK(sart_read_merged(hs[g], imgN, NULL, &cN) == SART_OK);
        CHECK(cN.n_rays == n && cN.n_passed == c1.n_passed && cN.n_retraced == c1.n_retraced);
        for (int e = 0; e < SART_N_EXIT_CODES; ++e) CHECK(cN.n_exit[e] == c1.n_exit[e]);
        CHECK(fabs(cN.sum_w / c1.sum_w - 1.0) < 1e-12);
        double d = 0.0, t = 0.0;
        for (int i = 0; i < SART_IMAGE_BINS * SART_IMAGE_BINS; ++i) { d += fabs(imgN[i] - img1[i]); t += img1[i]; }
        CHECK(d <= 1e-12 * t);
      }
      /* repeated without a reset: the handles' own images were not touched by the collective */
      CHECK(sart_allreduce(hs, ng) == SART_OK && sart_read_merged(hs[0], NULL, NULL, &cN) == SART_OK && cN.n_rays == n);
      printf("abi_driver: sart_allreduce over %d GPU(s): %llu rays, %llu passed, merged == single-GPU run\n", ng,
             (unsigned long long)cN.n_rays, (unsigned long long)cN.n_passed);
    }
    for (int g = 1; g < ng; ++g) sart_destroy(hs[g]);
  }
  sart_destroy(h);
  return 0;
}
