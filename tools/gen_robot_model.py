#!/usr/bin/env python3
"""URDF -> constant tables for the fixed-base chain robots (iiwa14).

Reproduces the model-building conventions that the reference relies on through
pinocchio::urdf::buildModel (reference: src/robot/robot.cpp:26):
  * no root joint is added: a fixed-base URDF gives a fixed-base model;
  * bodies behind `fixed` joints are merged into the parent joint's body;
  * joint placement = URDF <origin> of the joint in the parent *joint* frame;
  * rpy -> rotation through urdfdom's quaternion formula (Rotation::setFromRPY);
  * gravity (0,0,-9.81) in the world frame.

The output is DATA (numbers read from the URDF), written as a C header that both
the CUDA kernels (idocp_b200/csrc) and the CPU oracle (oracle/) include, and as
a JSON fixture for the tests.  Run from the repo root when /root/reference is
mounted:  python tools/gen_robot_model.py
"""
import json
import math
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IIWA_URDF = "/root/reference/examples/iiwa14/iiwa_description/urdf/iiwa14.urdf"


def rpy_to_matrix(r, p, y):
    # urdfdom: Rotation::setFromRPY (half-angle quaternion), then Eigen quaternion -> matrix
    phi, the, psi = r / 2.0, p / 2.0, y / 2.0
    x = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    yq = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    z = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    w = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    s = math.sqrt(x * x + yq * yq + z * z + w * w)
    x, yq, z, w = x / s, yq / s, z / s, w / s
    # Eigen::Quaternion::toRotationMatrix
    tx, ty, tz = 2 * x, 2 * yq, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * yq, tz * yq, tz * z
    return np.array([
        [1 - (tyy + tzz), txy - twz, txz + twy],
        [txy + twz, 1 - (txx + tzz), tyz - twx],
        [txz - twy, tyz + twx, 1 - (txx + tyy)],
    ])


def _vec(s, n=3):
    v = [float(x) for x in s.split()]
    assert len(v) == n
    return np.array(v)


def parse_origin(elem):
    o = elem.find("origin") if elem is not None else None
    if o is None:
        return np.eye(3), np.zeros(3)
    xyz = _vec(o.get("xyz", "0 0 0"))
    rpy = _vec(o.get("rpy", "0 0 0"))
    return rpy_to_matrix(*rpy), xyz


def skew(c):
    return np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])


class Inertia:
    """(m, com, I about com) in some frame; supports SE3 action and addition."""

    def __init__(self, m=0.0, c=None, I=None):
        self.m = m
        self.c = np.zeros(3) if c is None else c
        self.I = np.zeros((3, 3)) if I is None else I

    def transformed(self, R, p):
        return Inertia(self.m, R @ self.c + p, R @ self.I @ R.T)

    def __add__(self, o):
        m = self.m + o.m
        if m == 0.0:
            return Inertia()
        c = (self.m * self.c + o.m * o.c) / m
        I = np.zeros((3, 3))
        for b in (self, o):
            d = b.c - c
            I = I + b.I - b.m * skew(d) @ skew(d)
        return Inertia(m, c, I)


def load_chain(urdf_path):
    root = ET.parse(urdf_path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    children = {}
    child_links = set()
    for j in joints:
        children.setdefault(j.find("parent").get("link"), []).append(j)
        child_links.add(j.find("child").get("link"))
    roots = [n for n in links if n not in child_links]
    assert len(roots) == 1
    # urdfdom keeps child joints in a std::map keyed by name -> siblings sorted by joint name
    for k in children:
        children[k].sort(key=lambda j: j.get("name"))

    def link_inertia(name):
        ine = links[name].find("inertial")
        if ine is None:
            return Inertia()
        R, p = parse_origin(ine)
        m = float(ine.find("mass").get("value"))
        t = ine.find("inertia")
        ixx, ixy, ixz, iyy, iyz, izz = (float(t.get(k)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz"))
        I = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
        return Inertia(m, p, R @ I @ R.T)

    model = {"names": [], "parent": [], "R": [], "p": [], "axis": [], "inertia": [],
             "q_min": [], "q_max": [], "v_max": [], "effort": [], "frames": []}

    # frames are numbered as pinocchio does: universe, then per visited joint/body
    model["frames"].append(("universe", -1, np.eye(3), np.zeros(3)))
    # pinocchio 2.x UrdfVisitor::addRootJoint (fixed base): a FIXED_JOINT frame "root_joint"
    # precedes the root link's BODY frame, which is why iiwa_link_ee_kuka has id 22
    # (examples/iiwa14/task_space_ocp.cpp:67) and iiwa_link_3 has id 10.
    model["frames"].append(("root_joint", -1, np.eye(3), np.zeros(3)))

    def visit(link_name, parent_joint, R_acc, p_acc):
        """R_acc,p_acc: placement of `link_name` frame in the frame of movable joint `parent_joint` (-1 = universe)."""
        model["frames"].append((link_name, parent_joint, R_acc.copy(), p_acc.copy()))
        ine = link_inertia(link_name).transformed(R_acc, p_acc)
        if parent_joint >= 0:
            model["inertia"][parent_joint] = model["inertia"][parent_joint] + ine
        for j in children.get(link_name, []):
            Rj, pj = parse_origin(j)
            R_new, p_new = R_acc @ Rj, p_acc + R_acc @ pj
            typ = j.get("type")
            child = j.find("child").get("link")
            model["frames"].append((j.get("name"), parent_joint, R_new.copy(), p_new.copy()))
            if typ == "fixed":
                visit(child, parent_joint, R_new, p_new)
            elif typ == "floating":
                # pinocchio: URDF `floating` -> JointModelFreeFlyer (nq 7, nv 6); placement = joint origin
                assert parent_joint == -1 and "base" not in model
                model["base"] = {"name": j.get("name"), "R": R_new, "p": p_new, "inertia": Inertia()}
                model["base_frames_from"] = len(model["frames"])
                visit_base(child)
            elif typ in ("revolute", "continuous"):
                axis = _vec(j.find("axis").get("xyz"))
                lim = j.find("limit")
                idx = len(model["names"])
                model["names"].append(j.get("name"))
                model["parent"].append(parent_joint)
                model["R"].append(R_new)
                model["p"].append(p_new)
                model["axis"].append(axis)
                model["inertia"].append(Inertia())
                model["q_min"].append(float(lim.get("lower")))
                model["q_max"].append(float(lim.get("upper")))
                model["v_max"].append(float(lim.get("velocity")))
                model["effort"].append(float(lim.get("effort")))
                visit(child, idx, np.eye(3), np.zeros(3))
            else:
                raise NotImplementedError(typ)

    def visit_base(link_name):
        # bodies behind the free-flyer: the movable-joint index space restarts at -1 == "the base body"
        state = {"in_base": True}
        visit_in_base(link_name, np.eye(3), np.zeros(3))

    def visit_in_base(link_name, R_acc, p_acc):
        model["frames"].append((link_name, -1, R_acc.copy(), p_acc.copy()))
        model["base"]["inertia"] = model["base"]["inertia"] + link_inertia(link_name).transformed(R_acc, p_acc)
        for j in children.get(link_name, []):
            Rj, pj = parse_origin(j)
            R_new, p_new = R_acc @ Rj, p_acc + R_acc @ pj
            typ = j.get("type")
            child = j.find("child").get("link")
            model["frames"].append((j.get("name"), -1, R_new.copy(), p_new.copy()))
            if typ == "fixed":
                visit_in_base(child, R_new, p_new)
            elif typ in ("revolute", "continuous"):
                axis = _vec(j.find("axis").get("xyz"))
                lim = j.find("limit")
                idx = len(model["names"])
                model["names"].append(j.get("name"))
                model["parent"].append(-1)
                model["R"].append(R_new)
                model["p"].append(p_new)
                model["axis"].append(axis)
                model["inertia"].append(Inertia())
                model["q_min"].append(float(lim.get("lower")))
                model["q_max"].append(float(lim.get("upper")))
                model["v_max"].append(float(lim.get("velocity")))
                model["effort"].append(float(lim.get("effort")))
                visit(child, idx, np.eye(3), np.zeros(3))
            else:
                raise NotImplementedError(typ)

    visit(roots[0], -1, np.eye(3), np.zeros(3))
    return model


def c_array(name, arr, fmt="%.17g", qual="static const"):
    a = np.asarray(arr, dtype=float)
    dims = "".join("[%d]" % d for d in a.shape)
    def rec(x):
        if x.ndim == 1:
            return "{" + ", ".join(fmt % v for v in x) + "}"
        return "{\n  " + ",\n  ".join(rec(y) for y in x) + "}"
    return "%s double %s%s = %s;\n" % (qual, name, dims, rec(a))


def emit_iiwa14():
    m = load_chain(IIWA_URDF)
    n = len(m["names"])
    assert n == 7 and m["parent"] == [-1, 0, 1, 2, 3, 4, 5]
    for ax in m["axis"]:
        assert np.allclose(ax, [0, 0, 1])
    frame_names = [f[0] for f in m["frames"]]
    hdr = []
    hdr.append("/* GENERATED by tools/gen_robot_model.py from the iiwa14 URDF -- do not edit.\n"
               " * Fixed-base 7-joint chain, every joint revolute about its local Z axis.\n"
               " * Reference: src/robot/robot.cpp:26 (pinocchio::urdf::buildModel), robot.hxx:699-709 (limits).\n"
               " * PLACEMENT_R is row-major: placement of joint i in the frame of joint i-1 (world for i=0).\n"
               " * INERTIA is (xx,xy,xz,yy,yz,zz) about the body's centre of mass in the joint frame. */\n")
    hdr.append("#ifndef IDOCP_B200_MODEL_IIWA14_H_\n#define IDOCP_B200_MODEL_IIWA14_H_\n")
    hdr.append("#define IIWA14_NV 7\n#define IIWA14_GRAVITY 9.81\n")
    hdr.append(c_array("IIWA14_PLACEMENT_R", [R.reshape(9) for R in m["R"]]))
    hdr.append(c_array("IIWA14_PLACEMENT_P", m["p"]))
    hdr.append(c_array("IIWA14_MASS", [I.m for I in m["inertia"]]))
    hdr.append(c_array("IIWA14_COM", [I.c for I in m["inertia"]]))
    hdr.append(c_array("IIWA14_INERTIA", [[I.I[0, 0], I.I[0, 1], I.I[0, 2], I.I[1, 1], I.I[1, 2], I.I[2, 2]]
                                          for I in m["inertia"]]))
    hdr.append(c_array("IIWA14_Q_MIN", m["q_min"]))
    hdr.append(c_array("IIWA14_Q_MAX", m["q_max"]))
    hdr.append(c_array("IIWA14_V_MAX", m["v_max"]))
    hdr.append(c_array("IIWA14_EFFORT_MAX", m["effort"]))
    # end-effector frame used by examples/iiwa14/task_space_ocp.cpp:67 (frame id 22)
    ee = frame_names.index("iiwa_link_ee_kuka")
    hdr.append("/* frame %d = iiwa_link_ee_kuka (examples/iiwa14/task_space_ocp.cpp:67), attached to joint %d */\n"
               % (ee, m["frames"][ee][1]))
    hdr.append("#define IIWA14_EE_FRAME_ID %d\n#define IIWA14_EE_PARENT_JOINT %d\n" % (ee, m["frames"][ee][1]))
    hdr.append(c_array("IIWA14_EE_PLACEMENT_R", m["frames"][ee][2].reshape(9)))
    hdr.append(c_array("IIWA14_EE_PLACEMENT_P", m["frames"][ee][3]))
    hdr.append("#endif\n")
    text = "\n".join(hdr)
    for rel in ("idocp_b200/csrc/model_iiwa14.h", "oracle/model_iiwa14.h"):
        with open(os.path.join(REPO, rel), "w") as f:
            f.write(text)
    fixture = {
        "names": m["names"], "parent": m["parent"],
        "R": [R.tolist() for R in m["R"]], "p": [p.tolist() for p in m["p"]],
        "mass": [I.m for I in m["inertia"]], "com": [I.c.tolist() for I in m["inertia"]],
        "inertia": [I.I.tolist() for I in m["inertia"]],
        "q_min": m["q_min"], "q_max": m["q_max"], "v_max": m["v_max"], "effort": m["effort"],
        "frames": frame_names, "ee_frame": ee,
        "ee_R": m["frames"][ee][2].tolist(), "ee_p": m["frames"][ee][3].tolist(),
    }
    with open(os.path.join(REPO, "tests/golden/model_iiwa14.json"), "w") as f:
        json.dump(fixture, f, indent=1)
    print("iiwa14: %d joints, %d frames, ee frame id %d" % (n, len(frame_names), ee))
    for i, nm in enumerate(frame_names):
        print("  frame", i, nm)


ANYMAL_URDF = "/root/reference/examples/anymal/anymal_b_simple_description/urdf/anymal.urdf"
ANYMAL_CONTACT_FRAMES = [14, 24, 34, 44]  # LF, LH, RF, RH feet (examples/anymal/anymal_trotting.cpp:30)


def emit_anymal():
    m = load_chain(ANYMAL_URDF)
    n = len(m["names"])
    assert n == 12 and "base" in m
    assert m["names"] == ["LF_HAA", "LF_HFE", "LF_KFE", "LH_HAA", "LH_HFE", "LH_KFE",
                          "RF_HAA", "RF_HFE", "RF_KFE", "RH_HAA", "RH_HFE", "RH_KFE"]
    assert m["parent"] == [-1, 0, 1, -1, 3, 4, -1, 6, 7, -1, 9, 10]
    assert np.allclose(m["base"]["R"], np.eye(3)) and np.allclose(m["base"]["p"], 0)
    axis_id = []
    for ax in m["axis"]:
        k = int(np.argmax(np.abs(ax)))
        assert np.allclose(ax, np.eye(3)[k])
        axis_id.append(k)
    for R in m["R"]:
        assert np.allclose(R, np.eye(3))       # every ANYmal joint origin is a pure translation
    frame_names = [f[0] for f in m["frames"]]
    feet = ANYMAL_CONTACT_FRAMES
    assert [frame_names[f] for f in feet] == ["LF_FOOT", "LH_FOOT", "RF_FOOT", "RH_FOOT"]
    bodies = [m["base"]["inertia"]] + m["inertia"]
    total_mass = sum(b.m for b in bodies)
    hdr = []
    hdr.append("/* GENERATED by tools/gen_robot_model.py from the ANYmal-B URDF -- do not edit.\n"
               " * Free-flyer base (q: xyz + quaternion xyzw, v: body-frame linear+angular) + 4 legs x 3 revolute joints.\n"
               " * Reference: src/robot/robot.cpp:8-85 (pinocchio::urdf::buildModel + contact frames), robot.hxx:699-709.\n"
               " * Body 0 = base, body 1+j = child of joint j.  JOINT_PARENT: -1 = base.  JOINT_AXIS: 0/1/2 = X/Y/Z.\n"
               " * Every joint origin is a pure translation (JOINT_P, in the parent joint frame).\n"
               " * INERTIA is (xx,xy,xz,yy,yz,zz) about the body's centre of mass in the joint frame. */\n")
    hdr.append("#ifndef IDOCP_B200_MODEL_ANYMAL_H_\n#define IDOCP_B200_MODEL_ANYMAL_H_\n")
    hdr.append("#define ANYMAL_NJ 12\n#define ANYMAL_NV 18\n#define ANYMAL_NQ 19\n#define ANYMAL_NB 13\n"
               "#define ANYMAL_NCONTACT 4\n#define ANYMAL_GRAVITY 9.81\n#define ANYMAL_TOTAL_MASS %.17g\n" % total_mass)
    hdr.append("/* storage class of the tables: plain C by default, __device__ in the CUDA build (fb_math.cuh) */\n"
               "#ifndef ANYMAL_TABLE\n#define ANYMAL_TABLE static const\n#endif\n")
    hdr.append("ANYMAL_TABLE int ANYMAL_JOINT_PARENT[12] = {%s};\n" % ", ".join(str(p) for p in m["parent"]))
    hdr.append("ANYMAL_TABLE int ANYMAL_JOINT_AXIS[12] = {%s};\n" % ", ".join(str(a) for a in axis_id))
    hdr.append(c_array("ANYMAL_JOINT_P", m["p"], qual="ANYMAL_TABLE"))
    hdr.append(c_array("ANYMAL_MASS", [b.m for b in bodies], qual="ANYMAL_TABLE"))
    hdr.append(c_array("ANYMAL_COM", [b.c for b in bodies], qual="ANYMAL_TABLE"))
    hdr.append(c_array("ANYMAL_INERTIA", [[b.I[0, 0], b.I[0, 1], b.I[0, 2], b.I[1, 1], b.I[1, 2], b.I[2, 2]]
                                          for b in bodies], qual="ANYMAL_TABLE"))
    hdr.append(c_array("ANYMAL_Q_MIN", m["q_min"], qual="ANYMAL_TABLE"))
    hdr.append(c_array("ANYMAL_Q_MAX", m["q_max"], qual="ANYMAL_TABLE"))
    hdr.append(c_array("ANYMAL_V_MAX", m["v_max"], qual="ANYMAL_TABLE"))
    hdr.append(c_array("ANYMAL_EFFORT_MAX", m["effort"], qual="ANYMAL_TABLE"))
    hdr.append("/* contact frames %s (examples/anymal/anymal_trotting.cpp:30), parent joint = the leg's KFE */\n" % feet)
    hdr.append("ANYMAL_TABLE int ANYMAL_CONTACT_FRAME_ID[4] = {%s};\n" % ", ".join(str(f) for f in feet))
    hdr.append("ANYMAL_TABLE int ANYMAL_CONTACT_PARENT_JOINT[4] = {%s};\n"
               % ", ".join(str(m["frames"][f][1]) for f in feet))
    for f in feet:
        assert np.allclose(m["frames"][f][2], np.eye(3))
    hdr.append(c_array("ANYMAL_CONTACT_P", [m["frames"][f][3] for f in feet], qual="ANYMAL_TABLE"))
    hdr.append("#endif\n")
    text = "\n".join(hdr)
    for rel in ("idocp_b200/csrc/model_anymal.h", "oracle/model_anymal.h"):
        with open(os.path.join(REPO, rel), "w") as f:
            f.write(text)
    fixture = {
        "names": m["names"], "parent": m["parent"], "axis": axis_id,
        "p": [p.tolist() for p in m["p"]],
        "mass": [b.m for b in bodies], "com": [b.c.tolist() for b in bodies],
        "inertia": [b.I.tolist() for b in bodies], "total_mass": total_mass,
        "q_min": m["q_min"], "q_max": m["q_max"], "v_max": m["v_max"], "effort": m["effort"],
        "frames": frame_names, "contact_frames": feet,
        "contact_parent": [m["frames"][f][1] for f in feet],
        "contact_p": [m["frames"][f][3].tolist() for f in feet],
    }
    with open(os.path.join(REPO, "tests/golden/model_anymal.json"), "w") as f:
        json.dump(fixture, f, indent=1)
    print("anymal: %d joints, %d frames, total mass %.6f" % (n, len(frame_names), total_mass))


if __name__ == "__main__":
    if not os.path.exists(IIWA_URDF):
        sys.exit("reference URDF not found (run in the build container)")
    emit_iiwa14()
    emit_anymal()
