#!/usr/bin/env python
"""Generate csrc/fk_generated.cuh: straight-line forward kinematics specialised per built-in robot.

The generic FK (armsim_device.cuh chain_fk) multiplies by every entry of each joint's fixed rotation and translation
(48 FP ops per joint).  For the URDFs the reference ships, those fixed transforms are signed axis permutations
(rpy are multiples of pi/2) and single-axis offsets, so R * Rf is a renaming of columns and p += R * t is <= 3 FMAs.
This script propagates the chain symbolically (entries are exact 0, +-1, a constant, or a register expression), folds
the constants and emits only the operations that remain -- "kernels built from the repo's URDFs".

Input : drl-on-robot-arm_b200/robots/*.json (written by tools/gen_robot_models.py).
Output: drl-on-robot-arm_b200/csrc/fk_generated.cuh  (committed; regenerate with `python tools/gen_fk_kernels.py`).
"""
import json
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "drl-on-robot-arm_b200")
EPS = 1e-9


def rpy_to_mat(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return [[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr]]


def snap(v):
    for t in (0.0, 1.0, -1.0):
        if abs(v - t) < EPS:
            return t
    return v


class Emitter:
    """values are ('c', float) constants or ('v', name, sign) registers"""

    def __init__(self):
        self.lines = []
        self.n = 0

    def tmp(self):
        self.n += 1
        return "t%d" % self.n

    @staticmethod
    def const(v):
        return ("c", snap(float(v)))

    @staticmethod
    def lit(v):
        return repr(float(v)) + "f"

    def ref(self, x):
        if x[0] == "c":
            return self.lit(x[1])
        return ("-" if x[2] < 0 else "") + x[1]

    def neg(self, x):
        if x[0] == "c":
            return ("c", -x[1])
        return ("v", x[1], -x[2])

    def mul(self, a, b):
        if a[0] == "c" and b[0] == "c":
            return ("c", snap(a[1] * b[1]))
        if a[0] == "c":
            a, b = b, a
        if b[0] == "c":
            if b[1] == 0.0:
                return ("c", 0.0)
            if b[1] == 1.0:
                return a
            if b[1] == -1.0:
                return self.neg(a)
            t = self.tmp()
            self.lines.append("const float %s = %s * %s;" % (t, a[1], self.lit(b[1] * a[2])))
            return ("v", t, 1)
        t = self.tmp()
        self.lines.append("const float %s = %s * %s;" % (t, a[1], b[1]))
        return ("v", t, a[2] * b[2])

    def add(self, a, b):
        if a[0] == "c" and b[0] == "c":
            return ("c", snap(a[1] + b[1]))
        if a[0] == "c":
            a, b = b, a
        if b[0] == "c" and b[1] == 0.0:
            return a
        t = self.tmp()
        self.lines.append("const float %s = %s + %s;" % (t, self.ref(a), self.ref(b)))
        return ("v", t, 1)

    def fma(self, a, b, c):
        """a * b + c with folding"""
        if (a[0] == "c" and a[1] == 0.0) or (b[0] == "c" and b[1] == 0.0):
            return c
        if a[0] == "c" and b[0] == "c":
            return self.add(("c", snap(a[1] * b[1])), c)
        if c[0] == "c" and c[1] == 0.0:
            return self.mul(a, b)
        if a[0] == "c":
            a, b = b, a
        if b[0] == "c" and abs(b[1]) == 1.0:
            return self.add(a if b[1] > 0 else self.neg(a), c)
        t = self.tmp()
        if b[0] == "c":
            self.lines.append("const float %s = fmaf(%s, %s, %s);" % (t, a[1], self.lit(b[1] * a[2]), self.ref(c)))
        else:
            sgn = a[2] * b[2]
            self.lines.append("const float %s = fmaf(%s%s, %s, %s);" % (t, "-" if sgn < 0 else "", a[1], b[1], self.ref(c)))
        return ("v", t, 1)

    def dot3(self, a, b, c0=None):
        acc = c0 if c0 is not None else ("c", 0.0)
        for x, y in zip(a, b):
            acc = self.fma(x, y, acc)
        return acc


def gen_robot(model):
    E = Emitter()
    L = E.lines
    Rb = rpy_to_mat(*model["base_rpy"])
    R = [[E.const(Rb[i][k]) for k in range(3)] for i in range(3)]
    p = [E.const(v) for v in model["base_xyz"]]
    out_P, out_Z = [], []
    for j, jt in enumerate(model["joints"]):
        L.append("// joint %d (%s)" % (j, jt["name"]))
        t = [E.const(v) for v in jt["xyz"]]
        p = [E.dot3(R[i], t, p[i]) for i in range(3)]
        Rf = rpy_to_mat(*jt["rpy"])
        Rfc = [[E.const(Rf[i][k]) for k in range(3)] for i in range(3)]
        M = [[E.dot3(R[i], [Rfc[0][k], Rfc[1][k], Rfc[2][k]]) for k in range(3)] for i in range(3)]
        out_P.append(list(p))
        out_Z.append([M[i][2] for i in range(3)])
        L.append("float s%d, c%d;" % (j, j))
        L.append("sincos_bounded(q[%d], s%d, c%d);" % (j, j, j))
        s, c = ("v", "s%d" % j, 1), ("v", "c%d" % j, 1)
        newR = []
        for i in range(3):
            r0 = E.fma(M[i][1], s, E.mul(M[i][0], c))
            r1 = E.fma(M[i][1], c, E.neg(E.mul(M[i][0], s)))
            newR.append([r0, r1, M[i][2]])
        R = newR
    body = ["  " + ln for ln in L]
    tail = []
    for i in range(3):
        tail.append("  p[%d] = %s;" % (i, E.ref(p[i])))
    for i in range(3):
        for k in range(3):
            tail.append("  R[%d] = %s;" % (3 * i + k, E.ref(R[i][k])))
    jac = []
    for j in range(7):
        for i in range(3):
            jac.append("    P[%d][%d] = %s; Z[%d][%d] = %s;" % (j, i, E.ref(out_P[j][i]), j, i, E.ref(out_Z[j][i])))
    nops = sum(1 for ln in L if ln.startswith("const float"))
    return body, tail, jac, nops


def main():
    robots = [("ARMSIM_ROBOT_KUKA_IIWA", "kuka_iiwa"), ("ARMSIM_ROBOT_DIANA_S1", "diana_s1")]
    out = ["// GENERATED by tools/gen_fk_kernels.py from drl-on-robot-arm_b200/robots/*.json -- do not edit.\n"
           "// Straight-line forward kinematics per built-in robot with the constant structure of its URDF folded in\n"
           "// (signed-permutation joint frames, single-axis offsets).  Same outputs as chain_fk<>: p, R = EE link frame;\n"
           "// P[j], Z[j] = origin and axis of joint j in the world.\n"
           "#pragma once\n\n"
           "template <int ROBOT>\nstruct RobotFK {\n  static constexpr bool kSpecialised = false;\n"
           "  template <bool WANT_JAC>\n  static __device__ __forceinline__ void run(const ChainParams& C, const float (&q)[NJ], float (&p)[3], float (&R)[9],\n"
           "                                             float (&P)[NJ][3], float (&Z)[NJ][3]) {\n"
           "    chain_fk<WANT_JAC>(C, q, p, R, P, Z);\n  }\n};\n\n"]
    for macro, name in robots:
        model = json.load(open(os.path.join(PKG, "robots", name + ".json")))
        body, tail, jac, nops = gen_robot(model)
        out.append("// %s: %d floating-point operations + 7 sincos (generic chain_fk: 336 + 7 sincos)\n" % (name, nops))
        out.append("template <>\nstruct RobotFK<%s> {\n  static constexpr bool kSpecialised = true;\n" % macro)
        out.append("  template <bool WANT_JAC>\n  static __device__ __forceinline__ void run(const ChainParams&, const float (&q)[NJ], float (&p)[3], float (&R)[9],\n"
                   "                                             float (&P)[NJ][3], float (&Z)[NJ][3]) {\n")
        out.append("\n".join("  " + ln for ln in body) + "\n")
        out.append("\n".join("  " + ln for ln in tail) + "\n")
        out.append("    if (WANT_JAC) {\n" + "\n".join("  " + ln for ln in jac) + "\n    }\n")
        out.append("  }\n};\n\n")
        print("%s: %d FP ops" % (name, nops))
    path = os.path.join(PKG, "csrc", "fk_generated.cuh")
    open(path, "w").write("".join(out))
    print("wrote", path)


if __name__ == "__main__":
    main()
