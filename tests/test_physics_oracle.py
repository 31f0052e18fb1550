"""Invariant tests that pin the fp64 physics oracle WITHOUT PhysX (parity unpinned).

The checks use an independent world-frame forward kinematics written here in numpy from
resources/go1_model.json (finite-difference velocities; no spatial algebra shared with the oracle).
"""
import json
import os

import numpy as np
import pytest

from oracle import physics as ph

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "walk-these-ways_b200")
MODEL = json.load(open(os.path.join(PKG, "resources", "go1_model.json")))


def quat_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def axis_R(axis, a):
    c, s = np.cos(a), np.sin(a)
    if axis == 0:
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def bodies_world(pos, quat, q, payload=0.0):
    """[(mass, com_world, R_world, I_com_body)] for the 13 dynamic bodies."""
    R0 = quat_R(quat)
    mb = MODEL["base"]["mass"] + payload
    out = [(mb, np.array(pos), R0, np.array(MODEL["base"]["inertia_com"]) * mb / MODEL["base"]["mass"])]  # com = com_disp = 0
    for L in range(4):
        R, p = R0, np.array(pos)
        for j, part in enumerate(("hip", "thigh", "calf")):
            d = MODEL[part][L]
            p = p + R @ np.array(d["origin"])
            R = R @ axis_R(d["axis"], q[3 * L + j])
            out.append((d["mass"], p + R @ np.array(d["com"]), R, np.array(d["inertia_com"])))
    return out


def advance_config(pos, quat, q, linvel, angvel, qd, eps):
    w = np.array(angvel) * eps
    dq = np.array([*(0.5 * w), 1.0])
    x1, y1, z1, w1 = dq
    x2, y2, z2, w2 = quat
    qn = np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                   w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
    return np.array(pos) + eps * np.array(linvel), qn / np.linalg.norm(qn), np.array(q) + eps * np.array(qd)


def momentum_energy(s, g=0.0):
    pos, quat, lv, av, q, qd = ph.state_arrays(s)
    eps = 1e-7
    b0 = bodies_world(pos, quat, q)
    b1 = bodies_world(*advance_config(pos, quat, q, lv, av, qd, eps))
    P = np.zeros(3); Lm = np.zeros(3); T = 0.0; V = 0.0
    for (m, c0, R0, Ic), (_, c1, R1, _) in zip(b0, b1):
        v = (c1 - c0) / eps
        Wx = (R1 - R0) @ R0.T / eps
        w = np.array([Wx[2, 1] - Wx[1, 2], Wx[0, 2] - Wx[2, 0], Wx[1, 0] - Wx[0, 1]]) / 2
        Iw = R0 @ Ic @ R0.T
        P += m * v
        Lm += Iw @ w + np.cross(c0, m * v)
        T += 0.5 * m * v @ v + 0.5 * w @ Iw @ w
        V += m * g * c0[2]
    return P, Lm, T, V


def free_params(dt=1e-4, g=0.0):
    p = ph.default_params()
    p.dt = dt
    p.gravity[2] = -g
    p.contact_margin = -1e9          # no foot contacts
    for k in range(4):
        p.pen_k[k] = 0; p.pen_c[k] = 0
    p.limit_k = 0; p.limit_c = 0
    return p


def random_state(rng, z=5.0):
    quat = rng.normal(size=4); quat /= np.linalg.norm(quat)
    q = ph.DEFAULT_DOF_POS + rng.uniform(-0.3, 0.3, 12)
    return ph.make_state([0.3, -0.2, z], quat, rng.uniform(-1, 1, 3), rng.uniform(-2, 2, 3), q, rng.uniform(-3, 3, 12))


def test_total_mass():
    assert abs(MODEL["total_mass"] - 11.309932) < 1e-6     # 11.31 kg


def test_free_fall_at_rest_has_no_joint_acceleration():
    p = free_params(dt=0.005, g=9.8)
    s = ph.make_state([0, 0, 3], [0, 0, 0, 1], [0, 0, 0], [0, 0, 0], ph.DEFAULT_DOF_POS, np.zeros(12))
    a0, qdd = ph.aba(p, ph.make_dr(), s, np.zeros(12))
    assert np.allclose(a0[:3], 0, atol=1e-9) and np.allclose(a0[3:], [0, 0, -9.8], atol=1e-9)
    assert np.allclose(qdd, 0, atol=1e-8)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_momentum_conserved_under_internal_torques(seed):
    rng = np.random.default_rng(seed)
    p = free_params()
    s = random_state(rng)
    P0, L0, _, _ = momentum_energy(s)
    tau = rng.uniform(-1.5, 1.5, 12)      # small: stay clear of the joint velocity clamps
    for _ in range(300):
        ph.substep(p, ph.make_dr(), s, tau)
    P1, L1, _, _ = momentum_energy(s)
    assert np.allclose(P0, P1, atol=1e-3), (P0, P1)      # O(dt) integrator error, see the convergence test
    assert np.allclose(L0, L1, atol=3e-3), (L0, L1)


def test_momentum_drift_is_first_order_in_dt():
    """Halving dt halves the drift: the residual is integrator error, not an inconsistent model."""
    errs = []
    for dt, n in ((2e-4, 150), (1e-4, 300), (5e-5, 600)):
        rng = np.random.default_rng(0)
        p = free_params(dt=dt)
        s = random_state(rng)
        P0, L0, _, _ = momentum_energy(s)
        tau = rng.uniform(-1.5, 1.5, 12)
        for _ in range(n):
            ph.substep(p, ph.make_dr(), s, tau)
        P1, L1, _, _ = momentum_energy(s)
        errs.append((np.abs(P1 - P0).max(), np.abs(L1 - L0).max()))
    for k in range(2):
        assert 1.7 < errs[0][k] / errs[1][k] < 2.3 and 1.7 < errs[1][k] / errs[2][k] < 2.3, errs


@pytest.mark.parametrize("seed", [3, 4])
def test_energy_conserved_in_free_flight(seed):
    rng = np.random.default_rng(seed)
    p = free_params(g=9.8)
    s = random_state(rng)
    _, _, T0, V0 = momentum_energy(s, 9.8)
    for _ in range(500):
        ph.substep(p, ph.make_dr(), s, np.zeros(12))
    _, _, T1, V1 = momentum_energy(s, 9.8)
    assert abs((T1 + V1) - (T0 + V0)) < 2e-3 * max(1.0, abs(T0)), (T0 + V0, T1 + V1)


def test_linear_momentum_rate_equals_weight():
    p = free_params(dt=1e-4, g=9.8)
    rng = np.random.default_rng(7)
    s = random_state(rng)
    P0, _, _, _ = momentum_energy(s)
    for _ in range(100):
        ph.substep(p, ph.make_dr(), s, rng.uniform(-5, 5, 12))
    P1, _, _, _ = momentum_energy(s)
    assert np.allclose((P1 - P0) / (100 * 1e-4), [0, 0, -9.8 * MODEL["total_mass"]], atol=5e-2)


def test_feet_kinematics_match_independent_fk():
    rng = np.random.default_rng(11)
    s = random_state(rng)
    pos, quat, lv, av, q, qd = ph.state_arrays(s)
    fp, fv = ph.feet(s)

    def feet_fk(pos, quat, q):
        R0 = quat_R(quat); out = []
        for L in range(4):
            R, p = R0, np.array(pos)
            for j, part in enumerate(("hip", "thigh", "calf")):
                d = MODEL[part][L]
                p = p + R @ np.array(d["origin"]); R = R @ axis_R(d["axis"], q[3 * L + j])
            out.append(p + R @ np.array(MODEL["foot_offset"][L]))
        return np.array(out)
    f0 = feet_fk(pos, quat, q)
    f1 = feet_fk(*advance_config(pos, quat, q, lv, av, qd, 1e-7))
    assert np.allclose(fp, f0, atol=1e-12)
    assert np.allclose(fv, (f1 - f0) / 1e-7, atol=1e-5)


def test_standing_equilibrium_supports_weight():
    """PD-held default pose on flat ground: settles, feet carry m*g, nothing else touches."""
    p = ph.default_params()
    s = ph.make_state([0, 0, 0.34], [0, 0, 0, 1], [0, 0, 0], [0, 0, 0], ph.DEFAULT_DOF_POS, np.zeros(12))
    for _ in range(600):
        tau = np.clip(20.0 * (ph.DEFAULT_DOF_POS - np.array(s.q)) - 0.5 * np.array(s.qd), -33.5, 33.5)
        cf = ph.substep(p, ph.make_dr(), s, tau)
    assert 0.2 < s.pos[2] < 0.34
    assert abs(cf[[4, 8, 12, 16], 2].sum() - 9.8 * MODEL["total_mass"]) < 6.0
    assert np.allclose(cf[[0, 1, 2, 3, 5, 6, 7]], 0)
    assert np.abs(np.array(s.angvel)).max() < 0.5


def test_drop_does_not_gain_energy_and_base_contact_reported():
    """Limp robot dropped from 0.5 m: ends on the ground, trunk/hip contacts report force."""
    p = ph.default_params()
    s = ph.make_state([0, 0, 0.5], [0, 0, 0, 1], [0, 0, 0], [0, 0, 0], ph.DEFAULT_DOF_POS, np.zeros(12))
    E0 = momentum_energy(s, 9.8)
    seen_base = False
    for _ in range(400):
        cf = ph.substep(p, ph.make_dr(), s, -0.5 * np.array(s.qd))
        seen_base |= np.linalg.norm(cf[0]) > 1.0
        assert np.all(np.isfinite(np.array(s.q)))
    E1 = momentum_energy(s, 9.8)
    assert E1[2] + E1[3] < E0[2] + E0[3]
    assert seen_base and s.pos[2] < 0.15
