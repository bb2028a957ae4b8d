#!/usr/bin/env python
"""High-precision (mpmath, 60 digits) evaluation of traceAxion's GEOMETRY for single rays — development/validation aid.

The reference's formulas (src/raytracer.nim:481-534, 628-814, 1813-2083) evaluated without rounding noise: tells which
of two f64 results is closer to what the formulas mean. Only the happy path (entrance disc -> two mirrors -> window
plane) is followed; returns None where the ray leaves it.
"""
from __future__ import annotations

import mpmath as mp

mp.mp.dps = 60


def V(x, y, z):
    return mp.matrix([x, y, z])


def plane_point(p1, v, zc):
    lam = (zc - p1[2]) / v[2]
    return p1 + lam * v


def rot_in_x(v, ang, off):
    z = v[2] - off
    return V(v[0] * mp.cos(ang) + z * mp.sin(ang), v[1], z * mp.cos(ang) - v[0] * mp.sin(ang) + off)


def rot_in_y(v, ang, off):
    z = v[2] - off
    return V(v[0], v[1] * mp.cos(ang) - z * mp.sin(ang), z * mp.cos(ang) + v[1] * mp.sin(ang) + off)


def pick_root(p, d, a, hb, c, zmin, zmax):
    disc = hb * hb - a * c
    if disc < 0:
        return None
    sq = mp.sqrt(disc)
    for root in ((-hb - sq) / a, (-hb + sq) / a):
        z = p[2] + root * d[2]
        if zmin < z < zmax:
            return p + root * d
    return None


def reflect(n, frm, to):
    v = to - frm
    v = v / mp.norm(v)
    axis = V(n[1] * v[2] - n[2] * v[1], n[2] * v[0] - n[0] * v[2], n[0] * v[1] - n[1] * v[0])
    axis = axis / mp.norm(axis)
    alpha = mp.asin(abs((n.T * v)[0]) / mp.norm(n))
    vba = V(v[1] * axis[2] - v[2] * axis[1], v[2] * axis[0] - v[0] * axis[2], v[0] * axis[1] - v[1] * axis[0])
    return v * mp.cos(2 * alpha) - vba * mp.sin(2 * alpha), alpha


def trace(setup, O, E_xy):
    """setup: abi.Setup; O: (x, y, z) floats; E_xy: (x, y). Returns (x_chip, y_chip, shell, alpha1_deg, alpha2_deg)."""
    m, tel, pipes = setup.magnet, setup.telescope, setup.pipes
    mpf = mp.mpf
    O = V(*[mpf(float(c)) for c in O])
    E = V(mpf(float(E_xy[0])), mpf(float(E_xy[1])), mpf(m.lengthB))
    v = E - O
    # (rays that miss the entrance disc but enter through the bore wall follow the same line: rt:1820-1839 only
    # changes pathCB, which is a weight, not geometry)
    zE = mpf(m.lengthColdbore)
    zP1 = zE + mpf(pipes.cb2vt3_length)
    zP2 = zP1 + mpf(pipes.vt3xrt_length)
    pE = plane_point(O, v, zE)
    pP2 = plane_point(O, v, zP2)
    for pt, R in ((pE, m.radiusCB), (plane_point(O, v, zP1), pipes.cb2vt3_radius), (pP2, pipes.cb2vt3_radius)):
        if mp.sqrt(pt[0] ** 2 + pt[1] ** 2) >= R:
            return None
    rad = mp.pi / 180
    tX, tY = mpf(tel.telescope_turned_x) * rad, mpf(tel.telescope_turned_y) * rad
    b0 = mpf(tel.allAngles[0]) * rad
    half = ((mpf(tel.lMirror) + mpf(tel.allXsep[0]) / 2) * (mp.cos(b0) + mp.cos(3 * b0))) / 2
    oe = V(mpf(tel.optics_entrance[0]), mpf(tel.optics_entrance[1]), 0)
    pE = rot_in_y(rot_in_x(V(pE[0], pE[1], pE[2] - zP2), tX, half), tY, half) - oe
    pP2 = rot_in_y(rot_in_x(V(pP2[0], pP2[1], pP2[2] - zP2), tX, half), tY, half) - oe
    vX = pP2 - pE
    pEnt = pE + ((0 - pE[2]) / vX[2]) * vX
    rd = mp.sqrt(pEnt[0] ** 2 + pEnt[1] ** 2)
    hit = None
    for j in range(tel.nShells):
        if tel.allR1[j] < rd < tel.allR1[j] + tel.allThickness[j]:
            return None
        if hit is None and tel.allR1[j] > rd:
            hit = j
    if hit is None:
        return None
    r1 = mpf(tel.allR1[hit]); beta = mpf(tel.allAngles[hit]) * rad; xsep = mpf(tel.allXsep[hit]); l = mpf(tel.lMirror)
    beta3 = 3 * beta
    dm = mp.cos(beta) * (xsep + l)
    f = mpf(setup.detectorInstall.distanceDetectorXRT)
    wolter = tel.kind in (1, 3)
    p, d = pE, pEnt - pE
    if wolter:
        t = mp.tan(beta)
        r3 = -t * l + mp.sqrt(t * l * t * l + r1 * r1)
        e = 2 * r3 * t
        pm1 = pick_root(p, d, d[0] ** 2 + d[1] ** 2, p[0] * d[0] + p[1] * d[1] + e * d[2] / 2,
                        p[0] ** 2 + p[1] ** 2 - r3 ** 2 - e * l + e * p[2], 0, l * mp.cos(beta))
        if pm1 is None:
            return None
        mm = 1 / (r3 * t / mp.sqrt(r3 * r3 + r3 * 2 * t * (l - pm1[2])))
        n = V(pm1[0], pm1[1], mp.sqrt(pm1[0] ** 2 + pm1[1] ** 2) / mm)
    else:
        t = mp.tan(beta)
        k = t * t
        pm1 = pick_root(p, d, d[0] ** 2 + d[1] ** 2 - k * d[2] ** 2, p[0] * d[0] + p[1] * d[1] + r1 * t * d[2] - k * p[2] * d[2],
                        p[0] ** 2 + p[1] ** 2 - r1 ** 2 + 2 * r1 * t * p[2] - k * p[2] ** 2, 0, l * mp.cos(beta))
        if pm1 is None:
            return None
        n = V(pm1[0], pm1[1], t * mp.sqrt(pm1[0] ** 2 + pm1[1] ** 2))
    v1, a1 = reflect(n, pE, pEnt)
    p, d = pm1, 200 * v1
    if wolter:
        t3 = mp.tan(beta3 / 3)
        r3 = -t3 * l + mp.sqrt(t3 * l * t3 * l + r1 * r1)
        T = mp.tan(beta3)
        den = f + r3 / mp.tan(2 * beta3 / 3)
        e = 2 * r3 * T
        g = e / den
        pm2 = pick_root(p, d, d[0] ** 2 + d[1] ** 2 - g * d[2] ** 2,
                        p[0] * d[0] + p[1] * d[1] + g * d[2] * l - g * d[2] * p[2] + e * d[2] / 2,
                        p[0] ** 2 + p[1] ** 2 - r3 ** 2 - e * l + e * p[2] - g * l * l + 2 * g * p[2] * l - g * p[2] ** 2,
                        dm, dm + l * mp.cos(beta3))
        if pm2 is None:
            return None
        u = l - pm2[2]
        mm = 1 / (r3 * T * (1 + 2 * u / den) / mp.sqrt(r3 * r3 + r3 * 2 * T * u * (1 + u / den)))
        n = V(pm2[0], pm2[1], mp.sqrt(pm2[0] ** 2 + pm2[1] ** 2) / mm)
    else:
        r2 = r1 - l * mp.sin(beta)
        r3 = r2 - xsep / 2 * mp.tan(beta)
        r4 = r3 - xsep / 2 * mp.tan(beta3)
        t = mp.tan(beta3)
        k = t * t
        pz = p[2] - dm
        pm2 = pick_root(p, d, d[0] ** 2 + d[1] ** 2 - k * d[2] ** 2, p[0] * d[0] + p[1] * d[1] + r4 * t * d[2] - k * pz * d[2],
                        p[0] ** 2 + p[1] ** 2 - r4 ** 2 + 2 * r4 * t * pz - k * pz ** 2, dm, dm + l * mp.cos(beta3))
        if pm2 is None:
            return None
        n = V(pm2[0], pm2[1], t * mp.sqrt(pm2[0] ** 2 + pm2[1] ** 2))
    v2, a2 = reflect(n, pm1, pm1 + 200 * v1)
    distDet = dm - mpf(tel.allXsep[8]) / 2 * mp.cos(beta) + f - mpf(setup.detectorInstall.distanceWindowFocalPlane)
    pr = mpf(pipes.pipesTurned) * rad
    dsh = -mpf(tel.optics_entrance[0])
    a = rot_in_x(pm2, pr, 0) - V(dsh, 0, 0)
    b = rot_in_x(pm2 + 200 * v2, pr, 0) - V(dsh, 0, 0)
    w = b - a
    nn = (distDet / mp.cos(pr) - a[2]) / w[2]
    pdw = a + nn * w
    x = pdw[0] - mpf(setup.detectorInstall.lateralShift)
    y = pdw[1] - mpf(setup.detectorInstall.transversalShift)
    cx, cy = mpf(setup.consts.chipXMax) / 2, mpf(setup.consts.chipYMax) / 2
    return float(-x + cx), float(y + cy), hit, float(a1 / rad), float(a2 / rad)


# ---------------------------------------------------------------------------------------------------------------------
# Exit-code arbitration: the geometric decisions of traceAxion (rt:1813-2147) in 60-digit arithmetic, with the distance of
# the ray to the nearest decision boundary it met (in mm at the place of the decision). Away from every boundary — margin
# well above the 1e-4 .. 1e-3 mm rounding noise of the reference's f64 formulation — any correct evaluation of the
# reference's formulas must return this exit code, whatever its arithmetic.
EXIT = {"passed": 0, "missed_bore": 1, "clip_exit_cb": 2, "clip_pipe_vt3": 3, "clip_pipe_xrt": 4, "opaque": 5,
        "outside_shells": 6, "glass_front": 7, "nickel": 8, "no_mirror_hit": 9, "window_aperture": 10}


def classify(setup, O, E_xy):
    """Returns (exit code, margin [mm], ambiguous). `ambiguous`: mirror 1 was missed — the reference then carries on with
    pointMirror1 = pointExitCB and ends as nickel or no_mirror_hit depending on that garbage (rt:2040-2057); either is
    accepted. PASSED stands for "reached the detector window" (the weight decides passed / zero_weight)."""
    m, tel, pipes = setup.magnet, setup.telescope, setup.pipes
    mpf = mp.mpf
    O = V(*[mpf(float(c)) for c in O])
    E = V(mpf(float(E_xy[0])), mpf(float(E_xy[1])), mpf(m.lengthB))
    v = E - O
    margin = [mpf("1e9")]

    def near(d):
        margin[0] = min(margin[0], abs(d))

    def rho(pt):
        return mp.sqrt(pt[0] ** 2 + pt[1] ** 2)
    zE = mpf(m.lengthColdbore)
    zP1 = zE + mpf(pipes.cb2vt3_length)
    zP2 = zP1 + mpf(pipes.vt3xrt_length)
    R = mpf(m.radiusCB)
    pE = plane_point(O, v, zE)
    p0 = plane_point(O, v, 0)
    near(rho(pE) - R)
    if rho(pE) >= R:
        near(rho(p0) - R)
        return (EXIT["clip_exit_cb"] if rho(p0) < R else EXIT["missed_bore"]), float(margin[0]), False
    for z, code in ((zP1, "clip_pipe_vt3"), (zP2, "clip_pipe_xrt")):
        pt = plane_point(O, v, z)
        near(rho(pt) - mpf(pipes.cb2vt3_radius))          # quirk Q2: the first pipe's radius for both
        if rho(pt) >= mpf(pipes.cb2vt3_radius):
            return EXIT[code], float(margin[0]), False
    pP2 = plane_point(O, v, zP2)
    rad = mp.pi / 180
    tX, tY = mpf(tel.telescope_turned_x) * rad, mpf(tel.telescope_turned_y) * rad
    b0 = mpf(tel.allAngles[0]) * rad
    half = ((mpf(tel.lMirror) + mpf(tel.allXsep[0]) / 2) * (mp.cos(b0) + mp.cos(3 * b0))) / 2
    oe = V(mpf(tel.optics_entrance[0]), mpf(tel.optics_entrance[1]), 0)
    pE = rot_in_y(rot_in_x(V(pE[0], pE[1], pE[2] - zP2), tX, half), tY, half) - oe
    pP2 = rot_in_y(rot_in_x(V(pP2[0], pP2[1], pP2[2] - zP2), tX, half), tY, half) - oe
    vX = pP2 - pE
    pEnt = pE + ((0 - pE[2]) / vX[2]) * vX
    rd = rho(pEnt)
    if tel.kind in (1, 3):      # XMM / Abrixas: blocker, ring, spider (rt:1668-1701)
        xmm = tel.kind == 1
        pS = pE + (((-85 if xmm else -35) - pE[2]) / vX[2]) * vX
        arms, halfw, step = (17, mpf("1.145"), mpf("22.5")) if xmm else (7, mpf("3.75"), mpf(60))
        hit = False
        if xmm:
            for edge in (mpf("64.7"), mpf("151.6"), mpf("151.6") - mpf("20.9")):
                near(rd - edge)
            hit = rd <= mpf("64.7") or (mpf("151.6") - mpf("20.9") < rd < mpf("151.6"))
            if rd <= mpf("64.7") and tel.holeType != 0:
                # holes in the central blocker (rt:1674-1688; lineIntersectsObject rt:494-527 with centre z = 0: the point
                # of the line pointExitCB -> pointEntranceXRT in the plane z = 0, i.e. pointEntranceXRT itself)
                R = mpf(tel.holeInOptics)
                nH = tel.numberOfHoles
                half_n = nH - -(-nH // 2)
                s2 = mp.sqrt(2)
                for l in range(-half_n, half_n + 1):
                    cx = cy = mpf(0)
                    if l != 0:
                        if abs(l) % 2 == 0:
                            cy = 2 * mpf(l) * R
                        else:
                            cx = 2 * (mpf(l) + mpf(l) / abs(l)) * R
                    ix, iy = pEnt[0] - cx, pEnt[1] - cy
                    tx, ty = ix / s2 - iy / s2, ix / s2 + iy / s2
                    kind = tel.holeType      # 1 cross, 2 star, 3 circle, 4 square, 5 diamond (include/sart.h)
                    for a in (ix, iy):
                        if kind in (1, 2, 4): near(abs(a) - R)
                        if kind in (1, 2): near(abs(a) - 16 * R)
                    for a in (tx, ty):
                        if kind in (2, 5): near(abs(a) - R)
                        if kind == 2: near(abs(a) - 16 * R)
                    if kind == 3: near(mp.sqrt(ix * ix + iy * iy) - R)
                    cross = (abs(ix) < R and abs(iy) < 16 * R) or (abs(iy) < R and abs(ix) < 16 * R)
                    crossT = (abs(tx) < R and abs(ty) < 16 * R) or (abs(ty) < R and abs(tx) < 16 * R)
                    inside = {1: cross, 2: cross or crossT, 3: mp.sqrt(ix * ix + iy * iy) < R,
                              4: abs(ix) < R and abs(iy) < R, 5: abs(tx) < R and abs(ty) < R}[kind]
                    if inside:
                        hit = False
                        break
        else:
            near(rd - mpf("37.5"))
            hit = rd < mpf("37.5")
        central = rd <= mpf("64.7") if xmm else rd < mpf("37.5")   # the arm test is the branch for rays outside the centre (rt:1689-1701)
        if not hit and not central:
            for pt in (pEnt, pS):
                r = rho(pt)
                phi = mp.acos(pt[0] / r) / rad
                for i in range(arms):
                    for edge in (step * i - halfw, step * i + halfw):
                        near((phi - edge) * rad * r)
                    if step * i - halfw <= phi <= step * i + halfw:
                        hit = True
        if hit:
            return EXIT["opaque"], float(margin[0]), False
    n = tel.nShells
    near(rd - mpf(tel.allR1[n - 1]))
    if rd > mpf(tel.allR1[n - 1]):
        return EXIT["outside_shells"], float(margin[0]), False
    hit = None
    for j in range(n):
        r1j, edge = mpf(tel.allR1[j]), mpf(tel.allR1[j]) + mpf(tel.allThickness[j])
        near(rd - r1j); near(rd - edge)
        if r1j < rd < edge:
            return EXIT["glass_front"], float(margin[0]), False
        if hit is None and r1j > rd:
            hit = j
    if hit is None:
        return EXIT["no_mirror_hit"], float(margin[0]), False
    r1 = mpf(tel.allR1[hit]); beta = mpf(tel.allAngles[hit]) * rad; xsep = mpf(tel.allXsep[hit]); l = mpf(tel.lMirror)
    beta3 = 3 * beta
    dm = mp.cos(beta) * (xsep + l)
    f = mpf(setup.detectorInstall.distanceDetectorXRT)
    wolter = tel.kind in (1, 3)
    t = mp.tan(beta)

    def root(p, d, a, hb, c, zmin, zmax, slope):
        """pick_root with the margin of its interval test, converted to mm across the ray by the mirror's slope."""
        disc = hb * hb - a * c
        if disc < 0:
            near(mp.sqrt(-disc) / abs(a) * abs(d[2]) * slope)
            return None
        sq = mp.sqrt(disc)
        best = None
        for r in ((-hb - sq) / a, (-hb + sq) / a):
            z = p[2] + r * d[2]
            near(min(abs(z - zmin), abs(z - zmax)) * slope)
            if best is None and zmin < z < zmax:
                best = p + r * d
        return best
    p, d = pE, pEnt - pE
    if wolter:
        r3 = -t * l + mp.sqrt(t * l * t * l + r1 * r1)
        e = 2 * r3 * t
        pm1 = root(p, d, d[0] ** 2 + d[1] ** 2, p[0] * d[0] + p[1] * d[1] + e * d[2] / 2,
                   p[0] ** 2 + p[1] ** 2 - r3 ** 2 - e * l + e * p[2], 0, l * mp.cos(beta), t)
    else:
        k = t * t
        pm1 = root(p, d, d[0] ** 2 + d[1] ** 2 - k * d[2] ** 2, p[0] * d[0] + p[1] * d[1] + r1 * t * d[2] - k * p[2] * d[2],
                   p[0] ** 2 + p[1] ** 2 - r1 ** 2 + 2 * r1 * t * p[2] - k * p[2] ** 2, 0, l * mp.cos(beta), t)
    if pm1 is None:
        return EXIT["no_mirror_hit"], float(margin[0]), True
    if wolter:
        mm = 1 / (r3 * t / mp.sqrt(r3 * r3 + r3 * 2 * t * (l - pm1[2])))
        nrm = V(pm1[0], pm1[1], rho(pm1) / mm)
    else:
        nrm = V(pm1[0], pm1[1], t * rho(pm1))
    v1, a1 = reflect(nrm, pE, pEnt)
    p, d = pm1, 200 * v1
    if wolter:
        t3 = mp.tan(beta3 / 3)
        r3 = -t3 * l + mp.sqrt(t3 * l * t3 * l + r1 * r1)
        T = mp.tan(beta3)
        den = f + r3 / mp.tan(2 * beta3 / 3)
        e = 2 * r3 * T
        g = e / den
        pm2 = root(p, d, d[0] ** 2 + d[1] ** 2 - g * d[2] ** 2,
                   p[0] * d[0] + p[1] * d[1] + g * d[2] * l - g * d[2] * p[2] + e * d[2] / 2,
                   p[0] ** 2 + p[1] ** 2 - r3 ** 2 - e * l + e * p[2] - g * l * l + 2 * g * p[2] * l - g * p[2] ** 2,
                   dm, dm + l * mp.cos(beta3), T)
    else:
        r2 = r1 - l * mp.sin(beta)
        r3 = r2 - xsep / 2 * mp.tan(beta)
        r4 = r3 - xsep / 2 * mp.tan(beta3)
        T = mp.tan(beta3)
        k = T * T
        pz = p[2] - dm
        pm2 = root(p, d, d[0] ** 2 + d[1] ** 2 - k * d[2] ** 2, p[0] * d[0] + p[1] * d[1] + r4 * T * d[2] - k * pz * d[2],
                   p[0] ** 2 + p[1] ** 2 - r4 ** 2 + 2 * r4 * T * pz - k * pz ** 2, dm, dm + l * mp.cos(beta3), T)
    if hit > 0:     # lineHitsNickel rt:1706-1734, tested before the degenerate-hit test
        below = mpf(tel.allR1[hit - 1]) + mpf(tel.allThickness[hit - 1])
        lhs, rhs = mp.tan(a1) * (l - pm1[2]), r1 - below
        near(lhs - rhs)
        if lhs > rhs:
            return EXIT["nickel"], float(margin[0]), False
    if pm2 is None:
        return EXIT["no_mirror_hit"], float(margin[0]), False
    if wolter:
        u = l - pm2[2]
        mm = 1 / (r3 * T * (1 + 2 * u / den) / mp.sqrt(r3 * r3 + r3 * 2 * T * u * (1 + u / den)))
        nrm = V(pm2[0], pm2[1], rho(pm2) / mm)
    else:
        nrm = V(pm2[0], pm2[1], T * rho(pm2))
    v2, _ = reflect(nrm, pm1, pm1 + 200 * v1)
    distDet = dm - mpf(tel.allXsep[8]) / 2 * mp.cos(beta) + f - mpf(setup.detectorInstall.distanceWindowFocalPlane)
    pr = mpf(pipes.pipesTurned) * rad
    dsh = -mpf(tel.optics_entrance[0])
    a = rot_in_x(pm2, pr, 0) - V(dsh, 0, 0)
    b = rot_in_x(pm2 + 200 * v2, pr, 0) - V(dsh, 0, 0)
    w = b - a
    pdw = a + ((distDet / mp.cos(pr) - a[2]) / w[2]) * w
    x = pdw[0] - mpf(setup.detectorInstall.lateralShift)
    y = pdw[1] - mpf(setup.detectorInstall.transversalShift)
    cx, cy = mpf(setup.consts.chipXMax) / 2, mpf(setup.consts.chipYMax) / 2
    Rw = mpf(setup.detector.radiusWindow)
    near(mp.sqrt(x * x + y * y) - Rw); near(abs(x) - cx); near(abs(y) - cy)
    if mp.sqrt(x * x + y * y) > Rw or abs(x) > cx or abs(y) > cy:
        return EXIT["window_aperture"], float(margin[0]), False
    return EXIT["passed"], float(margin[0]), False
