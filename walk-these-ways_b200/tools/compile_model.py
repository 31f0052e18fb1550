#!/usr/bin/env python
"""Robot-model compiler: Go1 URDF -> constant tables for the sim kernel and the oracle.

Reads the reference's DATA files (never code):
  resources/robots/go1/urdf/go1.urdf      kinematic tree, inertias, limits   
  resources/actuator_nets/unitree_go1.pt  actuator MLP 6->32->32->1 softsign (legged_robot.py:1238-1251)
and writes
  walk-these-ways_b200/resources/go1_model.json        human-readable model
  walk-these-ways_b200/resources/actuator_net_go1.bin  1313 float32 (W1[32x6] b1[32] W2[32x32] b2[32] W3[32] b3[1])
  walk-these-ways_b200/csrc/go1_model_generated.h      the same numbers as C initialisers

Collapsing follows Isaac Gym asset options used by the reference (legged_robot_config.py:227,
collapse_fixed_joints=True): imu_link merges into trunk, *_thigh_shoulder (massless) into hip; the
foot is kept as a separate *reported* body (dont_collapse, go1.urdf:188) but is rigidly attached, so
for dynamics its inertia is merged into the calf.  Leg order is Isaac Gym's DOF order FL, FR, RL, RR.

Run in the build container only (needs /root/reference); outputs are committed.
"""
import json
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = os.environ.get("GO1_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LEGS = ["FL", "FR", "RL", "RR"]


def vec(s):
    return np.array([float(x) for x in s.split()])


def inertial(link):
    i = link.find("inertial")
    if i is None:
        return 0.0, np.zeros(3), np.zeros((3, 3))
    m = float(i.find("mass").get("value"))
    o = i.find("origin")
    c = vec(o.get("xyz")) if o is not None else np.zeros(3)
    assert o is None or np.allclose(vec(o.get("rpy", "0 0 0")), 0)
    t = i.find("inertia")
    g = lambda k: float(t.get(k))
    I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
    return m, c, I


def merge(parts):
    """Combine rigidly attached (mass, com, I_com) parts, all expressed in one frame."""
    m = sum(p[0] for p in parts)
    c = sum(p[0] * p[1] for p in parts) / m
    I = np.zeros((3, 3))
    for pm, pc, pI in parts:
        d = pc - c
        I += pI + pm * (d @ d * np.eye(3) - np.outer(d, d))
    return m, c, I


def main():
    root = ET.parse(f"{REF}/resources/robots/go1/urdf/go1.urdf").getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = {j.get("name"): j for j in root.findall("joint")}
    jorigin = lambda n: vec(joints[n].find("origin").get("xyz"))

    model = {"legs": LEGS, "source": "go1.urdf (reference resources/robots/go1/urdf)"}
    # --- base: trunk + imu_link (fixed)
    tm, tc, tI = inertial(links["trunk"])
    im, ic, iI = inertial(links["imu_link"])
    bm, bc, bI = merge([(tm, tc, tI), (im, ic + jorigin("imu_joint"), iI)])
    model["base"] = {"mass": bm, "com_urdf": bc.tolist(), "inertia_com": bI.tolist(),
                     "box_half": (vec(links["trunk"].find("collision/geometry/box").get("size")) / 2).tolist()}
    model["hip"], model["thigh"], model["calf"] = [], [], []
    for L in LEGS:
        for part, jn in (("hip", f"{L}_hip_joint"), ("thigh", f"{L}_thigh_joint"), ("calf", f"{L}_calf_joint")):
            j = joints[jn]
            lim = j.find("limit")
            axis = vec(j.find("axis").get("xyz"))
            m, c, I = inertial(links[f"{L}_{part}"])
            if part == "calf":
                fm, fc, fI = inertial(links[f"{L}_foot"])
                m, c, I = merge([(m, c, I), (fm, fc + jorigin(f"{L}_foot_fixed"), fI)])
            model[part].append({
                "joint": jn, "origin": jorigin(jn).tolist(), "axis": int(np.argmax(np.abs(axis))),
                "lower": float(lim.get("lower")), "upper": float(lim.get("upper")),
                "velocity": float(lim.get("velocity")), "effort": float(lim.get("effort")),
                "mass": m, "com": c.tolist(), "inertia_com": I.tolist()})
        model.setdefault("foot_offset", []).append(jorigin(f"{L}_foot_fixed").tolist())
    model["foot_radius"] = float(links["FL_foot"].find("collision/geometry/sphere").get("radius"))
    model["hip_collision_offset"] = [vec(links[f"{L}_hip"].find("collision/origin").get("xyz")).tolist() for L in LEGS]
    model["hip_collision_radius"] = float(links["FL_hip"].find("collision/geometry/cylinder").get("radius"))
    model["thigh_box"] = vec(links["FL_thigh"].find("collision/geometry/box").get("size")).tolist()
    model["calf_box"] = vec(links["FL_calf"].find("collision/geometry/box").get("size")).tolist()
    model["total_mass"] = bm + sum(model[p][k]["mass"] for p in ("hip", "thigh", "calf") for k in range(4))

    os.makedirs(f"{PKG}/resources", exist_ok=True)
    with open(f"{PKG}/resources/go1_model.json", "w") as f:
        json.dump(model, f, indent=1)

    # --- actuator net
    import torch
    net = torch.jit.load(f"{REF}/resources/actuator_nets/unitree_go1.pt", map_location="cpu")
    sd = {k: v.detach().numpy().astype(np.float32) for k, v in net.state_dict().items()}
    flat = np.concatenate([sd["0.weight"].ravel(), sd["0.bias"], sd["2.weight"].ravel(), sd["2.bias"],
                           sd["4.weight"].ravel(), sd["4.bias"]]).astype(np.float32)
    assert flat.size == 1313
    flat.tofile(f"{PKG}/resources/actuator_net_go1.bin")

    # --- C header
    def arr(a):
        return "{" + ", ".join(repr(float(x)) for x in np.asarray(a, dtype=np.float64).ravel()) + "}"

    with open(f"{PKG}/csrc/go1_model_generated.h", "w") as f:
        f.write("// GENERATED by tools/compile_model.py from the reference's go1.urdf — do not edit.\n")
        f.write("// Leg order FL,FR,RL,RR; per-leg joint order hip(x-axis),thigh(y),calf(y). SI units.\n#pragma once\n")
        f.write(f"#define GO1_BASE_MASS {bm!r}\n")
        f.write(f"static const double GO1_BASE_COM_URDF[3] = {arr(bc)};\n")
        f.write(f"static const double GO1_BASE_INERTIA_COM[9] = {arr(bI)};\n")
        f.write(f"static const double GO1_BASE_BOX_HALF[3] = {arr(model['base']['box_half'])};\n")
        for part in ("hip", "thigh", "calf"):
            P = part.upper()
            f.write(f"static const double GO1_{P}_ORIGIN[4][3] = {arr([d['origin'] for d in model[part]])};\n")
            f.write(f"static const double GO1_{P}_MASS[4] = {arr([d['mass'] for d in model[part]])};\n")
            f.write(f"static const double GO1_{P}_COM[4][3] = {arr([d['com'] for d in model[part]])};\n")
            f.write(f"static const double GO1_{P}_INERTIA_COM[4][9] = {arr([d['inertia_com'] for d in model[part]])};\n")
            f.write(f"static const double GO1_{P}_LIMITS[4][2] = {arr([[d['lower'], d['upper']] for d in model[part]])};\n")
            f.write(f"#define GO1_{P}_AXIS {model[part][0]['axis']}\n")
            f.write(f"#define GO1_{P}_VEL_LIMIT {model[part][0]['velocity']!r}\n")
        f.write(f"static const double GO1_FOOT_OFFSET[4][3] = {arr(model['foot_offset'])};\n")
        f.write(f"#define GO1_FOOT_RADIUS {model['foot_radius']!r}\n")
        f.write(f"static const double GO1_HIP_COLL_OFFSET[4][3] = {arr(model['hip_collision_offset'])};\n")
        f.write(f"#define GO1_HIP_COLL_RADIUS {model['hip_collision_radius']!r}\n")
        f.write(f"#define GO1_EFFORT_LIMIT {model['hip'][0]['effort']!r}\n")
    print("total mass", model["total_mass"], "base", bm, bc)


if __name__ == "__main__":
    sys.exit(main())
