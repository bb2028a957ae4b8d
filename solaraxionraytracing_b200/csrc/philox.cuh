// philox.cuh — Philox4x32-10 counter-based generator (Salmon et al., SC'11), host+device.
//
// Replaces std/random's global xoroshiro128+ state used by the reference's sampling procs
// (src/raytracer.nim:276, 418-419, 433-436, 464), which is not reproducible under Weave. Key = run seed,
// counter = (ray index lo, ray index hi, block, 0): a ray's random numbers depend only on (seed, global ray
// index), so any partition of a run over launches, streams or GPUs traces the very same rays.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SART_HD __host__ __device__ __forceinline__
#else
#define SART_HD inline
#endif

namespace sart {

struct Philox4 { uint32_t x, y, z, w; };

SART_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return uint32_t((uint64_t(a) * uint64_t(b)) >> 32);
#endif
}

SART_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    // one 32 x 32 -> 64 bit multiply (IMAD.WIDE.U32) gives both halves of a product; the round keys k + r W are written
    // as such, so that with a kernel-uniform seed they are uniform-datapath values rather than a running per-thread sum
    const uint64_t p0 = uint64_t(0xD2511F53u) * c0, p1 = uint64_t(0xCD9E8D57u) * c2;
    const uint32_t hi0 = uint32_t(p0 >> 32), lo0 = uint32_t(p0);
    const uint32_t hi1 = uint32_t(p1 >> 32), lo1 = uint32_t(p1);
    c0 = hi1 ^ c1 ^ (k0 + uint32_t(r) * 0x9E3779B9u);
    c1 = lo1;
    c2 = hi0 ^ c3 ^ (k1 + uint32_t(r) * 0xBB67AE85u);
    c3 = lo0;
  }
  return Philox4{c0, c1, c2, c3};
}

// The same generator with the ten round keys (k0 + r W0, k1 + r W1) precomputed: passed to a kernel by value they sit in
// the constant bank and enter the XORs as operands, which removes the two key-schedule additions per round (40
// instructions per ray).
struct PhiloxKeys { uint32_t k[10][2]; };
SART_HD PhiloxKeys philox_round_keys(uint64_t seed) {
  PhiloxKeys K;
  uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
  for (int r = 0; r < 10; ++r) { K.k[r][0] = k0; K.k[r][1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  return K;
}
SART_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& K) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    // one 32 x 32 -> 64 bit multiply (IMAD.WIDE.U32) gives both halves of a product: 2 multiplies per round instead of 4
    const uint64_t p0 = uint64_t(0xD2511F53u) * c0, p1 = uint64_t(0xCD9E8D57u) * c2;
    const uint32_t hi0 = uint32_t(p0 >> 32), lo0 = uint32_t(p0);
    const uint32_t hi1 = uint32_t(p1 >> 32), lo1 = uint32_t(p1);
    c0 = hi1 ^ c1 ^ K.k[r][0];
    c1 = lo1;
    c2 = hi0 ^ c3 ^ K.k[r][1];
    c3 = lo0;
  }
  return Philox4{c0, c1, c2, c3};
}
SART_HD void ray_words(const PhiloxKeys& K, uint64_t ray, uint32_t w[6]) {
  const Philox4 a = philox4x32_10(uint32_t(ray), uint32_t(ray >> 32), 0u, 0u, K);
  const Philox4 b = philox4x32_10(uint32_t(ray), uint32_t(ray >> 32), 1u, 0u, K);
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y;
}

// word -> uniform in (0,1): (w + 0.5) * 2^-32, exact in f64.
SART_HD double u01(uint32_t w) { return (double(w) + 0.5) * (1.0 / 4294967296.0); }

// The six uniforms of one ray in the reference's draw order:
// phi_sun, theta_sun, u_radius (rt:433-436), u_disk_r, u_disk_phi (rt:418-419), u_energy (rt:464).
SART_HD void ray_words(uint64_t seed, uint64_t ray, uint32_t w[6]) {
  const uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
  const Philox4 a = philox4x32_10(uint32_t(ray), uint32_t(ray >> 32), 0u, 0u, k0, k1);
  const Philox4 b = philox4x32_10(uint32_t(ray), uint32_t(ray >> 32), 1u, 0u, k0, k1);
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y;
}

}  // namespace sart
