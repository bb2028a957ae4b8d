// trace_f32x2.cuh — the plain run of the single-precision pipeline (solar source, vacuum stage, cone optics, telescope
// not turned, no ignore* flag) with TWO rays per thread, their FP32 arithmetic packed into the two-lane instructions of
// sm_100a (FFMA2 / FMUL2 / FADD2: one issue slot for two operations).
//
// Why: the one-ray kernel (trace_f32.cuh) is bound by instruction issue — 792 warp-instructions per 32 rays at 85 % of
// the issue slots, with the FMA pipe itself only 41 % busy. Of those instructions 330 are FP32 add / multiply / fma, 68
// load a constant from the parameter bank and 60 are control flow. With two rays per thread the first group issues once
// for both rays (the packed instruction's lanes are the two rays), and so do the second and third (constants and branches
// are per thread, not per ray). Integer work (Philox, table addresses), MUFU seeds, compares, selects and min/max have no
// packed form and stay one per ray; they operate on the halves of the packed registers directly, so there is no packing
// or unpacking traffic.
//
// The operations of a ray are the same IEEE operations in the same order as in trace_f32.cuh (stage_a32 / stage_b32 with
// kPlain, !kWolter, !kPre), so the error budgets (fast_params.h: Tol32) and the re-trace protocol carry over unchanged.
// A ray that ends early stays in its lane as a passenger (its arithmetic goes on with harmless values; every table index
// it forms is clamped) until its partner is done, and both outcomes go to the sink at the end of the pair.
#pragma once
#include "trace_f32.cuh"

namespace sart {
namespace fast {

#ifndef SART_F32X2_BLOCK
#define SART_F32X2_BLOCK 512
#endif
constexpr int kBlockX2 = SART_F32X2_BLOCK, kWarpsX2 = kBlockX2 / 32;

// sincos_2pi (fast_common.cuh) of u = w 2^-32 for the two words' float values
__device__ __forceinline__ void sincos_2pi_w2(f2 wf, f2& s, f2& c) {
  const f2 t = f2(6.283185307179586f) * fma2(wf, f2(2.3283064365386963e-10f), f2(-0.5f));
  s = f2(-__sinf(t.v.x), -__sinf(t.v.y));
  c = f2(-__cosf(t.v.x), -__cosf(t.v.y));
}
__device__ __forceinline__ void min_into(float& slack, float v) { slack = fminf(slack, v); }

// Emission shell of a radius word (stage_a32_head_words without the alias sampler).
__device__ __forceinline__ int radius_index32(const FastParams& P, const FastTables& T, const Smem32& S, uint32_t wr) {
  const uint2 c = S.radCells[wr >> (32 - kRadCellBits)];
  bool slow;
  int rIdx = cell_index(c, wr, slow);
  if (slow) rIdx = cell_index_slow(c, wr, T.radiusThr, T.radiusCDF, P.nRadii);
  return min(rIdx, P.nRadii - 1);
}

// pick_root32 for the pair (cone optics: no conditioning term). live0 / live1: the ray still takes part (a passenger's
// margins are not recorded). The returned roots are NaN (kMiss) where no root lies in range.
template <int grp, bool kMargins>
__device__ __forceinline__ f2 pick_root2(const Tol32& Q, f2 A, f2 hb, f2 C, f2 dz, f2 mid, f2 half, f2 tolC, f2 tolEnd,
                                         bool live0, bool live1, float& slack0, float& slack1, f2& tolZ) {
  const f2 disc = fma2(hb, hb, -(A * C));
  const bool neg0 = !(disc.v.x >= 0.0f), neg1 = !(disc.v.y >= 0.0f);
  const f2 rsq = rsqrt_nr2(f2(fmaxf(disc.v.x, 1e-30f), fmaxf(disc.v.y, 1e-30f)));
  const f2 sq = disc * rsq;
  const f2 q = -(hb + f2(copysignf(sq.v.x, hb.v.x), copysignf(sq.v.y, hb.v.y)));
  const f2 reach = fma2(abs2(mid) + half, f2(1.02f), f2(1.0f));   // pick_root32: the far root anywhere near the mirror
  const f2 far = abs2(q * dz) - reach * abs2(A);   // < 0: the far root q/A may lie in range as well
  const bool slow0 = !neg0 && far.v.x < 0.0f, slow1 = !neg1 && far.v.y < 0.0f;
  const f2 ts = C * rcp_nr2(q);
  const f2 zs = ts * dz;
  const f2 d = abs2(zs - mid) - half;
  f2 t(d.v.x < 0.0f ? ts.v.x : kMiss, d.v.y < 0.0f ? ts.v.y : kMiss);
  tolZ = f2(0.0f);
  if (kMargins) {
    tolZ = fma2(f2(Q.zrel), abs2(zs), f2(0.5f) * tolC * rsq);
    const f2 m = abs2(d) - (tolZ + tolEnd);
    const f2 mneg = -fma2(f2(Q.discRel), hb * hb, disc);   // a line that misses the surface by less than the budget of disc
    if (live0 && ((SART_UNC_GROUPS >> grp) & 1u)) min_into(slack0, neg0 ? mneg.v.x : m.v.x);
    if (live1 && ((SART_UNC_GROUPS >> grp) & 1u)) min_into(slack1, neg1 ? mneg.v.y : m.v.y);
    if (neg0) tolZ.v.x = 0.0f;
    if (neg1) tolZ.v.y = 0.0f;
  }
  if (neg0) t.v.x = kMiss;
  if (neg1) t.v.y = kMiss;
  if (slow0 || slow1) {   // rare: the exact order of the reference's root choice, per ray
    if (slow0) {
      if (kMargins && live0) slack0 = -1.0f;
      tolZ.v.x = 0.0f;
      t.v.x = pick_root_slow32(A.v.x, q.v.x, C.v.x, hb.v.x >= 0.0f, dz.v.x, mid.v.x - half.v.x, mid.v.x + half.v.x);
    }
    if (slow1) {
      if (kMargins && live1) slack1 = -1.0f;
      tolZ.v.y = 0.0f;
      t.v.y = pick_root_slow32(A.v.y, q.v.y, C.v.y, hb.v.y >= 0.0f, dz.v.y, mid.v.y - half.v.y, mid.v.y + half.v.y);
    }
  }
  return t;
}

struct F3x2 { f2 x, y, z; };
__device__ __forceinline__ f2 reflect2(const F3x2& n, F3x2& v) {
  const f2 s = fma2(n.x, v.x, fma2(n.y, v.y, n.z * v.z));
  const f2 as = abs2(s);
  const f2 f = fma2(f2(2.0f) * as, s, fma2(f2(-2.0f) * s, s, f2(1.0f)));
  const f2 m2as = f2(-2.0f) * as;
  v.x = fma2(v.x, f, m2as * n.x);
  v.y = fma2(v.y, f, m2as * n.y);
  v.z = fma2(v.z, f, m2as * n.z);
  return as;
}

// The exit of a ray whose line misses mirror 1 (stage_b32: nickel of the shell below, else no mirror hit), one ray.
template <bool kMargins>
__device__ __forceinline__ int miss1_code32(const Geo32& G, const ShellF32& sh, float below, int hitLayer, float x0, float y0,
                                            float tx, float ty, float invLen, float& slack) {
  const Tol32& Q = G.tol;
  int code = SART_EXIT_NO_MIRROR_HIT;
  if (hitLayer > 0) {
    const float lM = G.lMirror;
    const float zc = G.zExitCBtel;
    const float xc = fmaf(zc, tx, x0), yc = fmaf(zc, ty, y0);
    const float rc2 = fmaf(xc, xc, yc * yc);
    const float rc = rc2 * rsqrtf_nr(rc2);
    const float nz = sh.tan1 * rc;
    const float sg = (fmaf(xc, tx, yc * ty) + nz) * invLen * rsqrtf_nr(fmaf(rc, rc, nz * nz));
    const float a = fabsf(sg);
    const float lhs = a * (lM - zc), rhs = sh.R1 - below;
    const float m = fmaf(lhs, lhs, -rhs * rhs * (1.0f - a * a));
    SART_UNC(kUncNickel, fmaf(-fmaf(8.0f * Q.sinA, lM - zc, 4.0f * Q.nick), lhs + rhs, fabsf(m)));   // stage_b32: the budget over this lever arm
    if (m > 0.0f) code = SART_EXIT_NICKEL;
  }
  return code;
}

// Linear reflectivity lookup of the pair (refl_lookup, fast_common.cuh), rows given by their 32-bit element offsets.
__device__ __forceinline__ f2 refl_lookup2(const FastParams& P, const float* __restrict__ row, f2 alphaDeg, uint32_t rowOff0,
                                           uint32_t rowOff1, bool& clamped0, bool& clamped1) {
  const f2 x(fminf(fmaxf(alphaDeg.v.x, P.angleMin), P.angleMax), fminf(fmaxf(alphaDeg.v.y, P.angleMin), P.angleMax));
  clamped0 |= (x.v.x != alphaDeg.v.x);
  clamped1 |= (x.v.y != alphaDeg.v.y);
  const f2 fx = (x - f2(P.angleMin)) * f2(P.invReflDx);
  int i0 = int(fx.v.x), i1 = int(fx.v.y);
  const int iMax = P.nAngles - 2;
  if (i0 > iMax) i0 = iMax;
  if (i1 > iMax) i1 = iMax;
  const float* c0 = row + (rowOff0 + uint32_t(i0));
  const float* c1 = row + (rowOff1 + uint32_t(i1));
  const f2 z0(__ldg(c0), __ldg(c1)), z1(__ldg(c0 + 1), __ldg(c1 + 1));
  return fma2(fx - f2(float(i0), float(i1)), z1 - z0, z0);
}

// One pair of Monte Carlo rays from their Philox words to the sink. has1 = false: the second ray does not exist (tail of
// the launch); it runs as a passenger and reports nothing.
template <bool kMargins, class Sink>
__device__ __forceinline__ void trace_pair32(const FastParams& P, const Geo32& G, const FastTables& T, const Smem32& S,
                                             const PhiloxKeys& K, uint64_t ray0, uint64_t ray1, uint32_t id0, uint32_t id1,
                                             bool has1, Sink& sink) {
  const Tol32& Q = G.tol;
  constexpr float kDeg = 57.29577951308232f;
  uint32_t wa[6], wb[6];
  ray_words(K, ray0, wa);
  ray_words(K, ray1, wb);
  // ---- emission shells, and the energy cells of their rows (used at the end of stage A) rt:437, 464
  const int rI0 = radius_index32(P, T, S, wa[2]), rI1 = radius_index32(P, T, S, wb[2]);
  const uint2 cell0 = __ldg(reinterpret_cast<const uint2*>(cell_row(T, rI0)) + (wa[5] >> (32 - kEnCellBits)));
  const uint2 cell1 = __ldg(reinterpret_cast<const uint2*>(cell_row(T, rI1)) + (wb[5] >> (32 - kEnCellBits)));

  // ---- sampling rt:412-442, 1754-1764
  const f2 rs = fma2(f2(float(rI0), float(rI1)), f2(0.0005f), f2(0.0015f));
  f2 s1, c1, sd, cd;
  sincos_2pi_w2(f2(float(wa[0]), float(wb[0])), s1, c1);
  const f2 a2 = f2(3.14159265358979f) * (f2(float(wa[1]), float(wb[1])) * f2(2.3283064365386963e-10f));
  const f2 s2(__sinf(a2.v.x), __sinf(a2.v.y)), c2(__cosf(a2.v.x), __cosf(a2.v.y));
  const f2 rsun = rs * f2(G.radiusSun);
  const f2 Ox = rsun * (c1 * s2), Oy = rsun * (s1 * s2);
  sincos_2pi_w2(f2(float(wa[4]), float(wb[4])), sd, cd);
  const f2 rd = sqrt_pos2((f2(float(wa[3]), float(wb[3])) + f2(0.5f)) * f2(2.3283064365386963e-10f));
  const f2 ex = f2(G.radiusCB) * (rd * cd), ey = f2(G.radiusCB) * (rd * sd);
  const f2 invD = rcp_nr2(fma2(-c2, rsun, f2(G.lengthBplusSun)));   // 1 / (lengthB - O.z)
  const f2 sx = fma2(ex, invD, -(Ox * invD)), sy = fma2(ey, invD, -(Oy * invD));

  // ---- error budgets of the two rays (Tol32)
  float slack0 = kSlackInf, slack1 = kSlackInf;
  f2 lat(0.0f), det(0.0f), thrCB(0.0f);
  if (kMargins) {
    const f2 s1abs = abs2(sx) + abs2(sy);
    lat = fma2(f2(Q.latS), rs, fma2(f2(Q.latT), s1abs, f2(Q.latA)));
    det = fma2(f2(Q.detS), rs, fma2(f2(Q.detT), s1abs, f2(Q.detA)));
    thrCB = fma2(f2(Q.twoRcb), lat, f2(Q.circCB));
  }

  // ---- bore and pipes rt:1813-1872
  const f2 s2sum = fma2(sx, sx, sy * sy);
  const f2 p0x = fma2(-sx, f2(G.lengthB), ex), p0y = fma2(-sy, f2(G.lengthB), ey);
  const f2 mEnt = fma2(p0x, p0x, p0y * p0y) - f2(G.radiusCB2);
  const f2 pex = fma2(sx, f2(G.dzExitCB), ex), pey = fma2(sy, f2(G.dzExitCB), ey);
  const f2 mExit = fma2(pex, pex, pey * pey) - f2(G.radiusCB2);
  const bool hitEnt0 = mEnt.v.x < 0.0f, hitEnt1 = mEnt.v.y < 0.0f;
  const bool inExit0 = mExit.v.x < 0.0f, inExit1 = mExit.v.y < 0.0f;
  if (kMargins) {
    const f2 mA = abs2(mExit) - thrCB;
    const f2 mB = fma2(-f2(Q.entK), thrCB, abs2(mEnt));
    min_into(slack0, mA.v.x); min_into(slack1, mA.v.y);
    min_into(slack0, mB.v.x); min_into(slack1, mB.v.y);   // every ray: trace_f32.cuh, stage_a32
  }
  f2 path2 = f2(G.lengthB2) * (f2(1.0f) + s2sum);
  if (!(hitEnt0 && hitEnt1)) {   // a ray that enters through the bore wall rt:1820-1843
    const f2 hb = fma2(ex, sx, ey * sy), c = fma2(ex, ex, ey * ey) - f2(G.radiusCB2);
    if (kMargins) {
      const f2 mC = abs2(c) - thrCB;
      if (!hitEnt0) min_into(slack0, mC.v.x);
      if (!hitEnt1) min_into(slack1, mC.v.y);
    }
    const f2 disc = fma2(hb, hb, -(s2sum * c));
    const f2 rq = disc * rsqrt_nr2(disc);
    const f2 sq(disc.v.x > 1e-30f ? rq.v.x : 0.0f, disc.v.y > 1e-30f ? rq.v.y : 0.0f);
    const f2 ta = -(hb + sq) * rcp_nr2(s2sum), tb = c * rcp_nr2(sq - hb);
    const f2 t1(hb.v.x >= 0.0f ? ta.v.x : tb.v.x, hb.v.y >= 0.0f ? ta.v.y : tb.v.y);
    const f2 pw = t1 * t1 * (f2(1.0f) + s2sum);
    if (!hitEnt0) path2.v.x = pw.v.x;
    if (!hitEnt1) path2.v.y = pw.v.y;
  }
  f2 x0 = fma2(sx, f2(G.dzPipe2), ex), y0 = fma2(sy, f2(G.dzPipe2), ey);
  bool okP1_0 = true, okP1_1 = true, okP2_0 = true, okP2_1 = true;
  if (!P.pipesFree) {
    const f2 qx = fma2(sx, f2(G.dzPipe1), ex), qy = fma2(sy, f2(G.dzPipe1), ey);
    const f2 m1 = fma2(qx, qx, qy * qy) - f2(G.rPipe12);
    const f2 m2 = fma2(x0, x0, y0 * y0) - f2(G.rPipe12);   // quirk Q2: the radius of the first pipe
    okP1_0 = m1.v.x < 0.0f; okP1_1 = m1.v.y < 0.0f;
    okP2_0 = m2.v.x < 0.0f; okP2_1 = m2.v.y < 0.0f;
    if (kMargins) {
      const f2 thrPipe = fma2(f2(Q.twoRpipe), lat, f2(Q.circPipe));
      const f2 u1 = abs2(m1) - thrPipe, u2 = abs2(m2) - thrPipe;
      min_into(slack0, fminf(u1.v.x, u2.v.x)); min_into(slack1, fminf(u1.v.y, u2.v.y));
    }
  }
  // ---- telescope frame (not turned) rt:1888-1905
  x0 = x0 - f2(G.oeX); y0 = y0 - f2(G.oeY);
  const f2 tx = sx, ty = sy;
  const f2 rho0sq = fma2(x0, x0, y0 * y0);
  const f2 invRho0 = rsqrt_nr2(rho0sq);
  const f2 rho0 = rho0sq * invRho0;
  const f2 latRho = lat + f2(Q.rho);

  // ---- shell rt:1932-1957
  int code0 = -1, code1 = -1, hit0, hit1;
  {
    const f2 fb = (rho0 - f2(G.shellRhoMin)) * f2(G.shellInvStep);
    const int nG = P.nShellGuide - 1;
    const int b0 = max(0, min(int(fb.v.x), nG)), b1 = max(0, min(int(fb.v.y), nG));
    const uint2 ca = *reinterpret_cast<const uint2*>(S.shellTab + b0), cb = *reinterpret_cast<const uint2*>(S.shellTab + b1);
    const float B0 = __uint_as_float(ca.x), B1 = __uint_as_float(cb.x);
    const uint32_t pick0 = rho0.v.x < B0 ? 0x4440u : (rho0.v.x > B0 ? 0x4442u : 0x4441u);
    const uint32_t pick1 = rho0.v.y < B1 ? 0x4440u : (rho0.v.y > B1 ? 0x4442u : 0x4441u);
    hit0 = int(__byte_perm(ca.y, 0u, pick0)); hit1 = int(__byte_perm(cb.y, 0u, pick1));
    if (hit0 >= kShellCellFail) { code0 = hit0 - kShellCellFail; hit0 = 0; }
    if (hit1 >= kShellCellFail) { code1 = hit1 - kShellCellFail; hit1 = 0; }
    if (kMargins) {
      const f2 mS = abs2(rho0 - f2(B0, B1)) - latRho;
      min_into(slack0, mS.v.x); min_into(slack1, mS.v.y);
    }
  }
  if (!okP2_0) code0 = SART_EXIT_CLIP_PIPE_XRT;
  if (!okP2_1) code1 = SART_EXIT_CLIP_PIPE_XRT;
  if (!okP1_0) code0 = SART_EXIT_CLIP_PIPE_VT3;
  if (!okP1_1) code1 = SART_EXIT_CLIP_PIPE_VT3;
  if (!inExit0) code0 = hitEnt0 ? SART_EXIT_CLIP_EXIT_CB : SART_EXIT_MISSED_BORE;
  if (!inExit1) code1 = hitEnt1 ? SART_EXIT_CLIP_EXIT_CB : SART_EXIT_MISSED_BORE;
  bool unc0 = kMargins && slack0 <= 0.0f, unc1 = kMargins && slack1 <= 0.0f;   // as of the ray's exit
  bool clamped0 = false, clamped1 = false;

  RayResult out0, out1;
  if (code0 < 0 || code1 < 0) {
    // ---- energy index rt:464 (a passenger keeps index 0)
    int eIdx0 = 0, eIdx1 = 0;
    {
      bool slow;
      if (code0 < 0) {
        eIdx0 = cell_index(cell0, wa[5], slow);
        if (slow) eIdx0 = cell_index_slow(cell0, wa[5], thr_row(P, T, rI0), T.energyCDF + size_t(rI0) * P.nEnergies, P.nEnergies);
        if (eIdx0 > P.nEnergies - 1) { eIdx0 = P.nEnergies - 1; clamped0 = true; }
      }
      if (code1 < 0) {
        eIdx1 = cell_index(cell1, wb[5], slow);
        if (slow) eIdx1 = cell_index_slow(cell1, wb[5], thr_row(P, T, rI1), T.energyCDF + size_t(rI1) * P.nEnergies, P.nEnergies);
        if (eIdx1 > P.nEnergies - 1) { eIdx1 = P.nEnergies - 1; clamped1 = true; }
      }
    }
    const float4 elv0 = __ldg(reinterpret_cast<const float4*>(T.elut) + eIdx0);
    const float4 elv1 = __ldg(reinterpret_cast<const float4*>(T.elut) + eIdx1);
    const ShellF32& sh0 = S.shell[hit0];
    const ShellF32& sh1 = S.shell[hit1];
    const f2 below(hit0 > 0 ? S.shell[hit0 - 1].R1pT : 0.0f, hit1 > 0 ? S.shell[hit1 - 1].R1pT : 0.0f);
    const f2 lM(G.lMirror);
    const f2 t2sum = fma2(tx, tx, ty * ty);
    const f2 invLen = rsqrt_nr2(f2(1.0f) + t2sum);
    const f2 xt = fma2(x0, tx, y0 * ty);
    bool live0 = code0 < 0, live1 = code1 < 0;

    // ---- mirror 1 rt:1983-2020 (cone rho = r1 - tan(beta) z)
    const f2 tan1(sh0.tan1, sh1.tan1), R1(sh0.R1, sh1.R1);
    const f2 tolC1 = f2(sh0.twoR, sh1.twoR) * (lat + f2(Q.rho));
    f2 tolZ1;
    const f2 z1 = pick_root2<kUncMirror1, kMargins>(Q, fma2(-tan1, tan1, t2sum), fma2(tan1, R1, xt), (rho0 - R1) * (rho0 + R1),
                                                    f2(1.0f), f2(sh0.zmid1, sh1.zmid1), f2(sh0.zhalf1, sh1.zhalf1), tolC1,
                                                    f2(0.0f), live0, live1, slack0, slack1, tolZ1);
    if ((live0 && !(z1.v.x == z1.v.x)) || (live1 && !(z1.v.y == z1.v.y))) {   // kMiss
      if (live0 && !(z1.v.x == z1.v.x)) {
        code0 = miss1_code32<kMargins>(G, sh0, below.v.x, hit0, x0.v.x, y0.v.x, tx.v.x, ty.v.x, invLen.v.x, slack0);
        unc0 = kMargins && slack0 <= 0.0f; live0 = false;
      }
      if (live1 && !(z1.v.y == z1.v.y)) {
        code1 = miss1_code32<kMargins>(G, sh1, below.v.y, hit1, x0.v.y, y0.v.y, tx.v.y, ty.v.y, invLen.v.y, slack1);
        unc1 = kMargins && slack1 <= 0.0f; live1 = false;
      }
    }
    if (live0 || live1) {
      F3x2 pm = {fma2(tx, z1, x0), fma2(ty, z1, y0), z1};
      F3x2 v = {tx * invLen, ty * invLen, invLen};
      f2 sinA1, rhoM;
      {
        const f2 rr = fma2(pm.x, pm.x, pm.y * pm.y);
        const f2 ir = rsqrt_nr2(rr);
        rhoM = rr * ir;
        const f2 cosb(sh0.cosb, sh1.cosb);
        const F3x2 n = {pm.x * ir * cosb, pm.y * ir * cosb, f2(sh0.sinb, sh1.sinb)};
        sinA1 = reflect2(n, v);
      }
      // ---- mirror 2 rt:1994-2029 (cone rho = r4 - tan(3 beta) (z - distanceMirrors))
      const f2 tan2(sh0.tan2, sh1.tan2);
      const f2 mid2 = f2(sh0.zmid2, sh1.zmid2) - pm.z;
      const f2 pv = fma2(pm.x, v.x, pm.y * v.y), vv = fma2(v.x, v.x, v.y * v.y);
      const f2 tolC2 = (rhoM + rhoM) * fma2(f2(sh0.tan2p, sh1.tan2p), tolZ1, lat);
      const f2 rc = fma2(-tan2, pm.z - f2(sh0.dm, sh1.dm), f2(sh0.r4, sh1.r4));
      f2 tolZ2;
      const f2 t2 = pick_root2<kUncMirror2, kMargins>(Q, fma2(-(tan2 * tan2 * v.z), v.z, vv), fma2(tan2 * rc, v.z, pv),
                                                      (rhoM - rc) * (rhoM + rc), v.z, mid2, f2(sh0.zhalf2, sh1.zhalf2), tolC2,
                                                      tolZ1, live0, live1, slack0, slack1, tolZ2);
      // ---- nickel of the shell below rt:1706-1734, then the degenerate second hit
      {
        const f2 lhs = sinA1 * (lM - z1), rhs = R1 - below;
        const f2 m = fma2(lhs, lhs, -(rhs * rhs) * fma2(-sinA1, sinA1, f2(1.0f)));
        if (kMargins && ((SART_UNC_GROUPS >> kUncNickel) & 1u)) {
          const f2 mm = fma2(-(lhs + rhs), fma2(sinA1, tolZ1, f2(Q.nick)), abs2(m));
          if (live0 && hit0 > 0) min_into(slack0, mm.v.x);
          if (live1 && hit1 > 0) min_into(slack1, mm.v.y);
        }
        if (live0 && hit0 > 0 && m.v.x > 0.0f) { code0 = SART_EXIT_NICKEL; unc0 = kMargins && slack0 <= 0.0f; live0 = false; }
        if (live1 && hit1 > 0 && m.v.y > 0.0f) { code1 = SART_EXIT_NICKEL; unc1 = kMargins && slack1 <= 0.0f; live1 = false; }
      }
      if (live0 && !(t2.v.x == t2.v.x)) { code0 = SART_EXIT_NO_MIRROR_HIT; unc0 = kMargins && slack0 <= 0.0f; live0 = false; }
      if (live1 && !(t2.v.y == t2.v.y)) { code1 = SART_EXIT_NO_MIRROR_HIT; unc1 = kMargins && slack1 <= 0.0f; live1 = false; }
      if (live0 || live1) {
        pm.x = fma2(t2, v.x, pm.x); pm.y = fma2(t2, v.y, pm.y); pm.z = fma2(t2, v.z, pm.z);
        f2 sinA2;
        {
          const f2 rr = fma2(pm.x, pm.x, pm.y * pm.y);
          const f2 ir = rsqrt_nr2(rr);
          const f2 cos3b(sh0.cos3b, sh1.cos3b);
          const F3x2 n = {pm.x * ir * cos3b, pm.y * ir * cos3b, f2(sh0.sin3b, sh1.sin3b)};
          sinA2 = reflect2(n, v);
        }
        // ---- detector plane rt:797-814
        f2 xw, yw;
        {
          const f2 cP(G.cosPipe), sP(G.sinPipe);
          const f2 ax = fma2(pm.x, cP, pm.z * sP) - f2(G.dShift), az = fma2(pm.z, cP, -(pm.x * sP));
          const f2 wx = fma2(v.x, cP, v.z * sP), wz = fma2(v.z, cP, -(v.x * sP));
          const f2 iwz = rcp_nr2(wz);
          const f2 n = (f2(sh0.ddWin, sh1.ddWin) - az) * iwz;
          xw = fma2(n, wx, ax) - f2(G.lateralShift);
          yw = fma2(n, v.y, pm.y) - f2(G.transversalShift);
        }
        // ---- weights rt:2101-2128
        f2 wPreF;   // cos(yaw) x conversion probability, FP32
        {
          // yaw in degrees, fed to cos as radians (quirk Q3); atan_small's series for the slopes of solar rays
          f2 at;
          {
            const f2 x2 = ty * ty;
            at = ty * fma2(x2, fma2(x2, f2(0.2f), f2(-0.33333334f)), f2(1.0f));
            if (fabsf(ty.v.x) > 0.1f) at.v.x = atanf(ty.v.x);
            if (fabsf(ty.v.y) > 0.1f) at.v.y = atanf(ty.v.y);
          }
          const f2 ya = -at * f2(kDeg);
          out0.yaw = ya.v.x; out1.yaw = ya.v.y;
          const f2 pre(__cosf(ya.v.x), __cosf(ya.v.y));
          const f2 convVac = f2(P.convK) * path2;
          out0.pre = pre.v.x; out1.pre = pre.v.y;
          out0.convVac = convVac.v.x; out1.convVac = convVac.v.y;
          out0.path = sqrtf(path2.v.x); out1.path = sqrtf(path2.v.y);
          wPreF = pre;
        }
        auto asin2 = [](f2 x) { const f2 x2 = x * x; return x * fma2(x2, fma2(x2, f2(0.075f), f2(0.16666667f)), f2(1.0f)); };
        const f2 al1 = asin2(sinA1) * f2(kDeg), al2 = asin2(sinA2) * f2(kDeg);
        out0.a1 = al1.v.x; out1.a1 = al1.v.y; out0.a2 = al2.v.x; out1.a2 = al2.v.y;
        {
          const uint32_t nE1 = uint32_t(P.nEnergies + 1), nA = uint32_t(P.nAngles);
          const uint32_t rowOff0 = (uint32_t(sh0.coat & kCoatMask) * nE1 + uint32_t(eIdx0)) * nA;
          const uint32_t rowOff1 = (uint32_t(sh1.coat & kCoatMask) * nE1 + uint32_t(eIdx1)) * nA;
          clamped0 |= (sh0.coat & kCoatClamped) != 0; clamped1 |= (sh1.coat & kCoatClamped) != 0;
          clamped0 |= (__float_as_int(elv0.x) & (kLutClampRefl << 16)) != 0;
          clamped1 |= (__float_as_int(elv1.x) & (kLutClampRefl << 16)) != 0;
          if (kMargins && ((SART_UNC_GROUPS >> kUncAngle) & 1u)) {
            if (live0) min_into(slack0, Q.angLo - fmaxf(al1.v.x, al2.v.x));
            if (live1) min_into(slack1, Q.angLo - fmaxf(al1.v.y, al2.v.y));
          }
          const f2 r1 = refl_lookup2(P, T.reflE, al1, rowOff0, rowOff1, clamped0, clamped1);
          const f2 r2 = refl_lookup2(P, T.reflE, al2, rowOff0, rowOff1, clamped0, clamped1);
          // the product of two FP32 reflectivities can leave the FP32 range (1e-20 each at large angles)
          out0.refl = double(r1.v.x) * double(r2.v.x); out1.refl = double(r1.v.y) * double(r2.v.y);
          out0.wPre = out0.refl * double(wPreF.v.x) * double(out0.convVac);
          out1.wPre = out1.refl * double(wPreF.v.y) * double(out1.convVac);
        }
        out0.agas = elv0.w; out1.agas = elv1.w;
        out0.shell = hit0; out1.shell = hit1;
        out0.code = -1; out1.code = -1;
        out0.eIdx = eIdx0; out1.eIdx = eIdx1;
        out0.gasGamma = out0.gasE1 = out0.gasE2 = out0.gasInv2E = 0.f; out0.gasL = 0.0; out0.devDet = 0.f;
        out1.gasGamma = out1.gasE1 = out1.gasE2 = out1.gasInv2E = 0.f; out1.gasL = 0.0; out1.devDet = 0.f;
        // ---- window aperture rt:2139-2147
        const f2 rw2 = fma2(xw, xw, yw * yw);
        const f2 axw = abs2(xw), ayw = abs2(yw);
        if (kMargins && ((SART_UNC_GROUPS >> kUncWindow) & 1u)) {
          const f2 mW = abs2(rw2 - f2(G.radiusWindow2)) - fma2(f2(Q.twoRwin), det, f2(Q.circWin));
          if (live0) min_into(slack0, mW.v.x);
          if (live1) min_into(slack1, mW.v.y);
          if (Q.chipInside) {   // otherwise the window aperture lies inside the chip and decides alone
            const f2 ux = abs2(axw - f2(G.chipCX)), uy = abs2(ayw - f2(G.chipCY));
            if (live0) min_into(slack0, fminf(ux.v.x, uy.v.x) - det.v.x);
            if (live1) min_into(slack1, fminf(ux.v.y, uy.v.y) - det.v.y);
          }
        }
        const bool miss0 = rw2.v.x > G.radiusWindow2 || axw.v.x > G.chipCX || ayw.v.x > G.chipCY;
        const bool miss1 = rw2.v.y > G.radiusWindow2 || axw.v.y > G.chipCX || ayw.v.y > G.chipCY;
        // ---- strongback strips rt:2149-2185 (the margin of a ray that missed the aperture is not recorded)
        int sb0 = 2, sb1 = 2;
        if (P.nStripHalf > 0) {
          const f2 yt = abs2(fma2(yw, f2(G.cosTheta), -(xw * f2(G.sinTheta))));
          const f2 pitch(G.stripDist + G.stripWidth);
          const f2 u = yt - f2(0.5f * G.stripDist);
          const f2 us = u * f2(G.invStripPitch);
          const f2 fi(floorf(us.v.x), floorf(us.v.y));
          const f2 off = fma2(-fi, pitch, u);
          const float nS = float(P.nStripHalf);
          sb0 = (u.v.x > 0.0f && fi.v.x < nS && off.v.x > 0.0f && off.v.x < G.stripWidth) ? 1 : 0;
          sb1 = (u.v.y > 0.0f && fi.v.y < nS && off.v.y > 0.0f && off.v.y < G.stripWidth) ? 1 : 0;
          if (kMargins && ((SART_UNC_GROUPS >> kUncStrips) & 1u)) {
            const f2 e1 = abs2(off - f2(G.stripWidth)), e2 = pitch - off;
            if (live0 && !miss0) min_into(slack0, fminf(fminf(off.v.x, e1.v.x), e2.v.x) - det.v.x);
            if (live1 && !miss1) min_into(slack1, fminf(fminf(off.v.y, e1.v.y), e2.v.y) - det.v.y);
          }
        }
        const f2 xc = f2(G.chipCX) - xw, yc = yw + f2(G.chipCY);
        const f2 rr = rw2 * rsqrt_nr2(rw2);
        const f2 bx = xc * f2(G.invBinX), by = yc * f2(G.invBinY);
        auto tail = [&](bool miss, int sb, const float4& elv, float rw2h, float rrh, float xch, float ych, float bxh, float byh,
                        RayResult& out) {
          if (miss) {
            out.windowMiss = true; out.wPost = 0.0; out.x = out.y = out.r = 0.0; out.bin = -1;
            return;
          }
          out.windowMiss = false;
          const int sbe = __float_as_int(elv.x);
          const float tw = sb == 1 ? elv.z : (sb == 0 ? elv.y : 0.f);
          double post = double(tw);
          if (sbe != 0) {   // rare: an energy at which a Henke grid clamps, or soft X-rays on a strip (fast_params.h: EnergyLUT)
            const int cl = sbe >> 16;
            out.clamped |= (cl & (sb == 1 ? kLutClampStrongback : (sb == 0 ? kLutClampWindow : 0)) | (cl & kLutClampGas)) != 0;
            const int ex = (sbe << 16) >> 16;
            if (sb == 1 && ex != 0) post = __hiloint2double(__double2hiint(post) + ex * (1 << 20), __double2loint(post));
          }
          post *= double(elv.w);
          post *= double(P.exposure);
          out.wPost = post;
          out.r = double(rw2h > 1e-30f ? rrh : 0.0f);
          out.x = double(xch);
          out.y = double(ych);
          const int cx = int(floorf(bxh)), cy = int(floorf(byh));
          out.bin = (cx >= 0 && cx < SART_IMAGE_BINS && cy >= 0 && cy < SART_IMAGE_BINS) ? cy * SART_IMAGE_BINS + cx : -1;
        };
        out0.clamped = clamped0; out1.clamped = clamped1;
        tail(miss0, sb0, elv0, rw2.v.x, rr.v.x, xc.v.x, yc.v.x, bx.v.x, by.v.x, out0);
        tail(miss1, sb1, elv1, rw2.v.y, rr.v.y, xc.v.y, yc.v.y, bx.v.y, by.v.y, out1);
        if (live0) unc0 = kMargins && slack0 <= 0.0f;
        if (live1) unc1 = kMargins && slack1 <= 0.0f;
      }
    }
    // a ray that is still live here reached the weight stage: its outcome is out0 / out1
    if (live0) code0 = -1;
    if (live1) code1 = -1;
  }
  // ---- outcomes: one place per kind for both rays
  if (!(kMargins && unc0 && sink.defer(id0))) {
    if (code0 >= 0) sink.fail(code0); else sink.hit(out0);
  }
  if (has1 && !(kMargins && unc1 && sink.defer(id1))) {
    if (code1 >= 0) sink.fail(code1); else sink.hit(out1);
  }
}

}  // namespace fast
}  // namespace sart
