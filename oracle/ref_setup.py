"""Independent Python transcription of the reference's setup constructors — TEST INFRASTRUCTURE ONLY.

Second reading of src/raytracer.nim:248-272, 1098-1157, 1251-1409, 1464-1496 (the first one is the product's C++
host_setup.cpp); tests compare the two field by field so that a typo in either shows up.
"""
from __future__ import annotations

import math

from solaraxionraytracing_b200 import abi

LLNL = dict(
    entrance=[-83.0, 0.0, 0.0], exit=[-83.0, 0.0, 454.0],
    thickness=[0.2] * 14,
    R1=[63.006, 65.606, 68.305, 71.105, 74.011, 77.027, 80.157, 83.405, 86.775, 90.272, 93.902, 97.668, 101.576,
        105.632],
    xsep=[4.171, 4.140, 4.221, 4.190, 4.228, 4.245, 4.288, 4.284, 4.306, 4.324, 4.373, 4.387, 4.403, 4.481],
    angles=[0.579, 0.603, 0.628, 0.654, 0.680, 0.708, 0.737, 0.767, 0.798, 0.830, 0.863, 0.898, 0.933, 0.970],
    lMirror=225.0, hole=0.0, nHoles=5, holeType=abi.HT_CROSS, refl=abi.RK_MULTI_COATING, layers=[2, 5, 9, 14])

XMM = dict(
    entrance=[0.0, -0.0, 0.0], exit=[0.0, -0.0, 600.0],
    thickness=[0.468, 0.475, 0.482, 0.490, 0.497, 0.504, 0.511, 0.519, 0.526, 0.534, 0.542, 0.549, 0.557, 0.566,
               0.574, 0.583, 0.591, 0.600, 0.609, 0.618, 0.627, 0.636, 0.646, 0.655, 0.665, 0.675, 0.684, 0.694,
               0.704, 0.714, 0.724, 0.735, 0.745, 0.756, 0.768, 0.779, 0.790, 0.802, 0.814, 0.826, 0.838, 0.850,
               0.862, 0.874, 0.887, 0.900, 0.913, 0.927, 0.941, 0.955, 0.968, 0.983, 0.997, 1.011, 1.026, 1.041,
               1.055, 1.070],
    R1=[153.118, 155.4105, 157.7235, 160.0565, 162.42, 164.803, 167.217, 169.651, 172.115, 174.5995, 177.1145,
        179.6495, 182.2145, 184.9615, 187.739, 190.5465, 193.3845, 196.253, 199.1515, 202.0805, 205.0395, 208.0795,
        211.1495, 214.25, 217.381, 220.542, 223.7435, 226.9755, 230.2375, 233.54, 236.873, 240.236, 243.6395,
        247.2855, 250.9715, 254.6985, 258.4655, 262.2625, 266.1005, 269.9785, 273.897, 277.856, 281.8555, 285.9055,
        289.9955, 294.178, 298.661, 303.0945, 307.5685, 312.093, 316.658, 321.2735, 325.939, 330.6555, 335.4225,
        340.23, 345.0875, 349.996],
    xsep=[0.0] * 58,
    angles=[0.29, 0.294, 0.298, 0.303, 0.307, 0.312, 0.316, 0.321, 0.325, 0.33, 0.335, 0.34, 0.345, 0.35, 0.355,
            0.36, 0.366, 0.371, 0.377, 0.382, 0.388, 0.393, 0.399, 0.405, 0.411, 0.417, 0.423, 0.429, 0.435, 0.441,
            0.448, 0.454, 0.461, 0.467, 0.474, 0.481, 0.489, 0.496, 0.503, 0.51, 0.518, 0.525, 0.533, 0.54, 0.548,
            0.556, 0.564, 0.573, 0.581, 0.59, 0.598, 0.607, 0.616, 0.625, 0.634, 0.643, 0.652, 0.661],
    lMirror=300.0, hole=0.2, nHoles=1, holeType=abi.HT_NONE, refl=abi.RK_SINGLE_COATING, layers=[])

ABRIXAS = dict(
    entrance=[0.0, -60.0, 0.0], exit=[0.0, -60.0, 600.0],
    thickness=[0.2] * 7 + [0.25] * 7 + [0.3] * 5 + [0.35] * 5 + [0.4] * 3,
    R1=[38.125, 39.353, 40.581, 41.809, 43.036, 44.292, 45.577, 46.894, 48.295, 49.731, 51.201, 52.707, 54.249,
        55.829, 57.447, 59.157, 60.909, 62.703, 64.540, 66.423, 68.403, 70.431, 72.509, 74.637, 76.817, 79.102,
        81.443],
    xsep=[0.0] * 27,
    angles=[0.3335, 0.3443, 0.3550, 0.3657, 0.3765, 0.3874, 0.3987, 0.4102, 0.4225, 0.4350, 0.4479, 0.4610, 0.4745,
            0.4883, 0.5024, 0.5174, 0.5327, 0.5484, 0.5644, 0.5809, 0.5982, 0.6159, 0.6340, 0.6526, 0.6716, 0.6916,
            0.7120],
    lMirror=150.0, hole=0.2, nHoles=1, holeType=abi.HT_NONE, refl=abi.RK_SINGLE_COATING, layers=[])

TELESCOPES = {abi.TK_LLNL: LLNL, abi.TK_XMM: XMM, abi.TK_ABRIXAS: ABRIXAS}
MAGNETS = {  # lengthColdbore, B, lengthB, radiusCB, pGasRoom, tGas
    abi.ES_CAST: (9756.0, 9.0, 9260.0, 21.5, 1.0, 1.7),
    abi.ES_BABYIAXO: (11300.0, 2.0, 11000.0, 500.0, 1.0, 100.0)}
PIPES = {  # cb->vt3 (len, rad), vt3->xrt (len, rad), axis dist, turned
    abi.TK_LLNL: (127.66, 39.89, 111.7, 23.935, 0.0, 2.75),
    abi.TK_ABRIXAS: (114.3, 66.65, 171.43, 47.62, 0.0, 0.0),
    abi.TK_XMM: (225.0, 370.0, 250.0, 370.0, 0.0, 0.0)}
DET_INSTALL = {abi.TK_LLNL: 1485.0, abi.TK_ABRIXAS: 1600.0, abi.TK_XMM: 7500.0}
TEST_SOURCE = {  # parallel, energy, distance, radius, up, left, activity, lengthCol
    abi.ES_CAST: (1, 1.0, 100.0, 10.0, 200.0, 0.0, 1.0, 50.0),
    abi.ES_BABYIAXO: (1, 0.021, 2000.0, 350.0, 0.0, 0.0, 0.125, 0.0)}


def calc_window_vals(radius: float, nStrips: int, openRatio: float):
    total = math.pi * radius * radius
    strips = total * (1.0 - openRatio)
    pitch = radius * 2.0 / (nStrips + 1.0)
    lall = 0.0
    for i in range(int(round(nStrips / 2))):
        off = i * pitch + 0.5 * pitch
        lall += math.sqrt(radius * radius - off * off) * 2.0
    lall *= 2.0
    width = strips / lall
    return width, pitch - width


def make_setup(experiment: int, detector: int, stage: int, telescope: int, flags: int = 0) -> abi.Setup:
    s = abi.Setup()
    s.abi_version = abi.ABI_VERSION
    s.flags, s.experiment, s.stage, s.detectorKind = flags, experiment, stage, detector
    m = MAGNETS[experiment]
    (s.magnet.lengthColdbore, s.magnet.B, s.magnet.lengthB, s.magnet.radiusCB, s.magnet.pGasRoom, s.magnet.tGas) = m
    p = PIPES[telescope]
    (s.pipes.cb2vt3_length, s.pipes.cb2vt3_radius, s.pipes.vt3xrt_length, s.pipes.vt3xrt_radius,
     s.pipes.distanceCBAxisXRTAxis, s.pipes.pipesTurned) = p
    t = TELESCOPES[telescope]
    tel = s.telescope
    tel.kind, tel.nShells = telescope, len(t["R1"])
    tel.numberOfHoles, tel.holeType, tel.reflKind = t["nHoles"], t["holeType"], t["refl"]
    tel.nCoatings = len(t["layers"]) if t["layers"] else 1
    for i, v in enumerate(t["layers"]):
        tel.layers[i] = v
    for i in range(3):
        tel.optics_entrance[i] = t["entrance"][i]
        tel.optics_exit[i] = t["exit"][i]
    tel.lMirror, tel.holeInOptics = t["lMirror"], t["hole"]
    for i in range(tel.nShells):
        tel.allThickness[i], tel.allR1[i], tel.allXsep[i], tel.allAngles[i] = (
            t["thickness"][i], t["R1"][i], t["xsep"][i], t["angles"][i])
    ts = TEST_SOURCE[experiment]
    src = s.testSource
    src.active = 1 if flags & abi.CF_XRAY_TEST else 0
    (src.parallel, src.energy, src.distance, src.radius, src.offAxisUp, src.offAxisLeft, src.activity,
     src.lengthCol) = ts
    s.detectorInstall.distanceDetectorXRT = DET_INSTALL[telescope]
    d = s.detector
    d.windowYear = {abi.DK_INGRID2017: abi.WY_2017, abi.DK_INGRID2018: abi.WY_2018, abi.DK_INGRIDIAXO: abi.WY_IAXO}[detector]
    d.numberOfStrips, d.radiusWindow, d.openApertureRatio = 4, 7.0, 0.838
    d.windowThickness, d.alThickness, d.depthDet, d.detectorWindowAperture = 0.3, 0.02, 30.0, 14.0
    d.stripWidthWindow, d.stripDistWindow = calc_window_vals(7.0, 4, 0.838)
    d.theta = math.radians(20.0 if d.windowYear == abi.WY_IAXO else 30.0)
    c = s.consts
    c.radiusSun, c.distanceSunEarth, c.roomTemp, c.mAxion, c.g_agamma = 6.9e11, 1.5e14, 293.15, 0.0853, 1e-12
    c.chipXMax = c.chipYMax = 14.0
    c.tesla_to_eV2, c.m_to_inv_eV = 195.35277121325237, 5067730.716548338
    if flags & abi.CF_XRAY_TEST:
        c.exposureFactor = 1.0
    elif experiment == abi.ES_CAST:
        c.exposureFactor = 3.585e3 * 3600.0 * 1.5 * 90.0
    else:
        c.exposureFactor = 9.5e6 * 3600.0 * 12.0 * 90.0
    return s
