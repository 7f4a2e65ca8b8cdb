#!/usr/bin/env python
"""Generate the robot-model data the kernels, the oracle and the tests share.

Run in the BUILD container (needs /root/reference for the DianaS1 URDF and the
getJointInfo fixture); its outputs are committed so nothing on the GPU box reads
/root/reference:

  drl-on-robot-arm_b200/robots/kuka_iiwa.json   chain restated from the public pybullet_data
                                                kuka_iiwa/model.urdf (NOT in the reference tree)
  drl-on-robot-arm_b200/robots/diana_s1.json    chain parsed from models/diana/DianaS1_robot.urdf:29-216
  include/armsim_robot_models.h                 the same two chains as C tables (kernels + oracle)
  tests/golden/joint_info_fixture.json          envs/bmirobot_joints_info_pybullet.txt:1-16 parsed (p.getJointInfo dumps)
  tests/golden/ee_init_main_py.json             main.py:106 `initial_a` (EE position at init_joint_positions)

The Kuka table below is cross-checked here against the fixture: every
parentFramePos must equal (joint origin - parent inertial origin) and every
parentFrameOrn must be the inverse of the joint rpy.
"""
import ast
import json
import math
import os
import re
import sys
import xml.etree.ElementTree as ET

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("ARMSIM_REFERENCE", "/root/reference")
PKG = os.path.join(ROOT, "drl-on-robot-arm_b200")
PI = math.pi

# pybullet_data/kuka_iiwa/model.urdf restated: joint origin xyz / rpy in the
# parent link frame, limits, child-link inertial origin, mass, diagonal inertia.
KUKA = {
    "name": "kuka_iiwa",
    "base_xyz": [0.0, 0.0, 0.0],
    "base_rpy": [0.0, 0.0, 0.0],
    "base_inertial_xyz": [-0.1, 0.0, 0.07],
    "joints": [
        dict(name="lbr_iiwa_joint_1", xyz=[0, 0, 0.1575], rpy=[0, 0, 0], lower=-2.96705972839, upper=2.96705972839,
             com=[0, -0.03, 0.12], mass=4.0, inertia=[0.1, 0, 0, 0.09, 0, 0.02]),
        dict(name="lbr_iiwa_joint_2", xyz=[0, 0, 0.2025], rpy=[PI / 2, 0, PI], lower=-2.09439510239, upper=2.09439510239,
             com=[0.0003, 0.059, 0.042], mass=4.0, inertia=[0.05, 0, 0, 0.018, 0, 0.044]),
        dict(name="lbr_iiwa_joint_3", xyz=[0, 0.2045, 0], rpy=[PI / 2, 0, PI], lower=-2.96705972839, upper=2.96705972839,
             com=[0, 0.03, 0.13], mass=3.0, inertia=[0.08, 0, 0, 0.075, 0, 0.01]),
        dict(name="lbr_iiwa_joint_4", xyz=[0, 0, 0.2155], rpy=[PI / 2, 0, 0], lower=-2.09439510239, upper=2.09439510239,
             com=[0, 0.067, 0.034], mass=2.7, inertia=[0.03, 0, 0, 0.01, 0, 0.029]),
        dict(name="lbr_iiwa_joint_5", xyz=[0, 0.1845, 0], rpy=[-PI / 2, PI, 0], lower=-2.96705972839, upper=2.96705972839,
             com=[0.0001, 0.021, 0.076], mass=1.7, inertia=[0.02, 0, 0, 0.018, 0, 0.005]),
        dict(name="lbr_iiwa_joint_6", xyz=[0, 0, 0.2155], rpy=[PI / 2, 0, 0], lower=-2.09439510239, upper=2.09439510239,
             com=[0, 0.0006, 0.0004], mass=1.8, inertia=[0.005, 0, 0, 0.0036, 0, 0.0047]),
        dict(name="lbr_iiwa_joint_7", xyz=[0, 0.081, 0], rpy=[-PI / 2, PI, 0], lower=-3.05432619099, upper=3.05432619099,
             com=[0, 0, 0.02], mass=0.3, inertia=[0.001, 0, 0, 0.001, 0, 0.001]),
    ],
}
for j in KUKA["joints"]:
    j.update(effort=300.0, velocity=10.0, damping=0.5)


def rpy_to_mat(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    # R = Rz(yaw) * Ry(pitch) * Rx(roll)  (URDF convention)
    return [
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr],
    ]


def mat_to_quat_xyzw(m):
    tr = m[0][0] + m[1][1] + m[2][2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        return [(m[2][1] - m[1][2]) / s, (m[0][2] - m[2][0]) / s, (m[1][0] - m[0][1]) / s, 0.25 * s]
    i = max(range(3), key=lambda k: m[k][k])
    j, k = (i + 1) % 3, (i + 2) % 3
    s = math.sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0) * 2
    q = [0.0] * 4
    q[i] = 0.25 * s
    q[j] = (m[j][i] + m[i][j]) / s
    q[k] = (m[k][i] + m[i][k]) / s
    q[3] = (m[k][j] - m[j][k]) / s
    return q


def parse_fixture(path):
    """envs/bmirobot_joints_info_pybullet.txt -> list of two robots, each a list of getJointInfo tuples."""
    robots, cur = [], []
    for line in open(path, "r"):
        line = line.strip()
        if not line:
            if cur:
                robots.append(cur)
                cur = []
            continue
        line = re.sub(r"b'([^']*)'", r"'\1'", line)
        cur.append(list(ast.literal_eval(line)))
    if cur:
        robots.append(cur)
    return robots


def parse_urdf(path, name):
    """Minimal URDF reader: serial chain of revolute joints (fixed joints folded into the base)."""
    root = ET.parse(path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}

    def floats(s):
        return [float(x) for x in s.split()]

    joints = []
    for j in root.findall("joint"):
        if j.get("type") != "revolute":
            continue
        o = j.find("origin")
        lim = j.find("limit")
        dyn = j.find("dynamics")
        child = links[j.find("child").get("link")]
        inert = child.find("inertial")
        io = inert.find("origin")
        I = inert.find("inertia")
        axis = floats(j.find("axis").get("xyz"))
        assert axis == [0.0, 0.0, 1.0], "kernels assume joints revolve about local +z"
        joints.append(dict(
            name=j.get("name"), xyz=floats(o.get("xyz")), rpy=floats(o.get("rpy")),
            lower=float(lim.get("lower")), upper=float(lim.get("upper")),
            effort=float(lim.get("effort")), velocity=float(lim.get("velocity")),
            damping=float(dyn.get("damping")) if dyn is not None else 0.0,
            com=floats(io.get("xyz")), mass=float(inert.find("mass").get("value")),
            inertia=[float(I.get(k)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")],
        ))
    base = links["base_link"].find("inertial").find("origin")
    return {"name": name, "base_xyz": [0.0, 0.0, 0.0], "base_rpy": [0.0, 0.0, 0.0],
            "base_inertial_xyz": floats(base.get("xyz")), "joints": joints}


def check_against_fixture(robot, tuples, first_index):
    """parentFramePos == joint origin - parent inertial origin; parentFrameOrn == inverse(joint rpy)."""
    parent_com = robot["base_inertial_xyz"]
    worst = 0.0
    for k, j in enumerate(robot["joints"]):
        t = tuples[first_index + k]
        assert t[1] == j["name"], (t[1], j["name"])
        assert abs(t[8] - j["lower"]) < 1e-9 and abs(t[9] - j["upper"]) < 1e-9
        assert abs(t[10] - j["effort"]) < 1e-9 and abs(t[11] - j["velocity"]) < 1e-9
        assert abs(t[6] - j["damping"]) < 1e-12
        assert tuple(t[13]) == (0.0, 0.0, 1.0)
        exp_pos = [j["xyz"][i] - parent_com[i] for i in range(3)]
        for i in range(3):
            worst = max(worst, abs(exp_pos[i] - t[14][i]))
        R = rpy_to_mat(*j["rpy"])
        Rt = [[R[c][r] for c in range(3)] for r in range(3)]
        q = mat_to_quat_xyzw(Rt)
        fq = list(t[15])
        n = math.sqrt(sum(x * x for x in fq))  # Diana's first tuple is un-normalised (1,0,0,1e-13)
        fq = [x / n for x in fq]
        dot = abs(sum(a * b for a, b in zip(q, fq)))
        worst = max(worst, abs(1.0 - dot))
        parent_com = j["com"]
    return worst


def c_array(vals, fmt="%.17g"):
    return "{" + ", ".join(fmt % v for v in vals) + "}"


def emit_header(robots, path):
    out = []
    out.append("/* GENERATED by tools/gen_robot_models.py -- do not edit.\n"
               " * Built-in 7-DoF chains shared by the CUDA kernels (csrc/) and the CPU oracle (oracle/).\n"
               " * kuka_iiwa: pybullet_data/kuka_iiwa/model.urdf restated (SURVEY Appendix A), pinned by\n"
               " *            reference envs/bmirobot_joints_info_pybullet.txt:1-7 and main.py:106.\n"
               " * diana_s1 : reference models/diana/DianaS1_robot.urdf:29-216 (fixture lines 9-16).\n"
               " * Layout per joint: origin xyz, origin rpy (R = Rz(y)Ry(p)Rx(r)), then the joint revolves about local +z. */\n")
    out.append("#ifndef ARMSIM_ROBOT_MODELS_H\n#define ARMSIM_ROBOT_MODELS_H\n")
    out.append("#define ARMSIM_NJ 7\n")
    out.append("typedef struct ArmsimRobotModel {\n"
               "  const char* name;\n"
               "  double base_xyz[3], base_rpy[3];\n"
               "  double xyz[ARMSIM_NJ][3], rpy[ARMSIM_NJ][3];\n"
               "  double lower[ARMSIM_NJ], upper[ARMSIM_NJ], effort[ARMSIM_NJ], velocity[ARMSIM_NJ], damping[ARMSIM_NJ];\n"
               "  double mass[ARMSIM_NJ], com[ARMSIM_NJ][3];\n"
               "  double inertia[ARMSIM_NJ][6]; /* ixx ixy ixz iyy iyz izz about the COM, link axes */\n"
               "} ArmsimRobotModel;\n")
    for r in robots:
        J = r["joints"]
        assert len(J) == 7
        out.append("static const ArmsimRobotModel ARMSIM_MODEL_%s = {\n" % r["name"].upper())
        out.append('  "%s",\n' % r["name"])
        out.append("  %s, %s,\n" % (c_array(r["base_xyz"]), c_array(r["base_rpy"])))
        out.append("  {%s},\n" % ", ".join(c_array(j["xyz"]) for j in J))
        out.append("  {%s},\n" % ", ".join(c_array(j["rpy"]) for j in J))
        for key in ("lower", "upper", "effort", "velocity", "damping", "mass"):
            out.append("  %s,\n" % c_array([j[key] for j in J]))
        out.append("  {%s},\n" % ", ".join(c_array(j["com"]) for j in J))
        out.append("  {%s}\n" % ", ".join(c_array(j["inertia"]) for j in J))
        out.append("};\n")
    out.append("#endif\n")
    open(path, "w").write("".join(out))


def main():
    fixture = parse_fixture(os.path.join(REF, "envs", "bmirobot_joints_info_pybullet.txt"))
    assert len(fixture) == 2 and len(fixture[0]) == 7 and len(fixture[1]) == 8
    diana = parse_urdf(os.path.join(REF, "models", "diana", "DianaS1_robot.urdf"), "diana_s1")
    # envs/diana_cam_reach.py:201-204 loads the arm with base yaw pi
    diana["base_rpy"] = [0.0, 0.0, PI]
    w_k = check_against_fixture(KUKA, fixture[0], 0)
    w_d = check_against_fixture(diana, fixture[1], 1)
    print("fixture check: kuka worst |delta| = %.3g, diana worst |delta| = %.3g" % (w_k, w_d))
    assert w_k < 1e-9 and w_d < 1e-9

    os.makedirs(os.path.join(PKG, "robots"), exist_ok=True)
    for r in (KUKA, diana):
        with open(os.path.join(PKG, "robots", r["name"] + ".json"), "w") as f:
            json.dump(r, f, indent=1)
    emit_header([KUKA, diana], os.path.join(ROOT, "include", "armsim_robot_models.h"))

    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    with open(os.path.join(gold, "joint_info_fixture.json"), "w") as f:
        json.dump({"source": "reference envs/bmirobot_joints_info_pybullet.txt:1-16 (p.getJointInfo dumps)",
                   "fields": ["index", "name", "type", "qIndex", "uIndex", "flags", "damping", "friction", "lower", "upper",
                              "maxForce", "maxVelocity", "linkName", "axis", "parentFramePos", "parentFrameOrn", "parentIndex"],
                   "kuka_iiwa": fixture[0], "diana_s1": fixture[1]}, f, indent=1)
    # main.py:106
    src = open(os.path.join(REF, "main.py")).read()
    m = re.search(r"initial_a = \[([^\]]+)\]", src)
    ee = [float(x) for x in m.group(1).split(",")]
    with open(os.path.join(gold, "ee_init_main_py.json"), "w") as f:
        json.dump({"source": "reference main.py:106 initial_a (float32-rounded EE link-7 position at init_joint_positions, "
                             "envs/rl_reach_env.py:116-119)",
                   "init_joint_positions": [0.006418, 0.413184, -0.011401, -1.589317, 0.005379, 1.137684, -0.006539],
                   "ee": ee}, f, indent=1)
    print("wrote robots/*.json, include/armsim_robot_models.h, tests/golden/*.json")


if __name__ == "__main__":
    sys.exit(main())
