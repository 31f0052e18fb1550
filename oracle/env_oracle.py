"""env_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT.

torch-CPU fp32 restatement of the non-physics part of LeggedRobot.step() for AoS tensors [N, ...]:
  compute_torques        legged_robot.py:907-946   (+ actuator net 1242-1251)
  step_contact_targets   legged_robot.py:826-905
  check_termination      legged_robot.py:138-148
  compute_reward         legged_robot.py:263-300 + go1_gym/envs/rewards/corl_rewards.py:15-201
  compute_observations   legged_robot.py:302-491
  reset_idx (device part)legged_robot.py:150-239, 645-665, 948-1001
It is pinned against vectors produced by the reference's own code (tests/golden/make_golden.py ->
tests/golden/env_*.npz; checked by tests/test_env_oracle.py) and is what the `-m gpu` parity tests compare
the CUDA kernel with.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.
"""
import math
import os

import numpy as np
import torch

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "walk-these-ways_b200")
HIP = [0, 3, 6, 9]
FEET = [4, 8, 12, 16]                 # Isaac Gym body indices (base, then hip/thigh/calf/foot per leg)
PENALISED = [2, 6, 10, 14, 3, 7, 11, 15]   # "thigh" names then "calf" names (legged_robot.py:1522-1524)
TERMINATION = [0]


def quat_rotate_inverse(q, v):
    q_w = q[:, -1:]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0)
    b = torch.cross(q_vec, v, dim=-1) * q_w * 2.0
    c = q_vec * (q_vec * v).sum(-1, keepdim=True) * 2.0
    return a - b + c


def quat_apply(q, v):
    xyz = q[:, :3]
    t = torch.cross(xyz, v, dim=-1) * 2
    return v + q[:, 3:] * t + torch.cross(xyz, t, dim=-1)


def quat_mul(a, b):
    x1, y1, z1, w1 = a.unbind(-1)
    x2, y2, z2, w2 = b.unbind(-1)
    return torch.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                        w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], -1)


def quat_from_angle_axis(angle, axis):
    th = (angle / 2).unsqueeze(-1)
    ax = axis / axis.norm()
    q = torch.cat([ax * th.sin(), th.cos()], -1)
    return q / q.norm(dim=-1, keepdim=True)


class ActuatorNet:
    def __init__(self):
        w = np.fromfile(os.path.join(_PKG, "resources", "actuator_net_go1.bin"), dtype=np.float32)
        t = torch.from_numpy(w.copy())
        self.W1, self.b1 = t[:192].view(32, 6), t[192:224]
        self.W2, self.b2 = t[224:1248].view(32, 32), t[1248:1280]
        self.W3, self.b3 = t[1280:1312].view(1, 32), t[1312:1313]

    def __call__(self, x):
        ss = torch.nn.functional.softsign
        return ss(ss(x @ self.W1.T + self.b1) @ self.W2.T + self.b2) @ self.W3.T + self.b3


def compute_torques(s, P, net=None):
    """One control substep. s: dict of tensors (mutated like the reference mutates self). Returns torques [N,12]."""
    a = s["actions"][:, :12] * P["action_scale"]
    a[:, HIP] *= P["hip_scale_reduction"]
    if P["use_lag"]:
        s["lag_buffer"] = s["lag_buffer"][1:] + [a.clone()]
        s["joint_pos_target"] = s["lag_buffer"][0] + P["default_dof_pos"]
    else:
        s["joint_pos_target"] = a + P["default_dof_pos"]
    if P["control_type"] == "actuator_net":
        err = s["dof_pos"] - s["joint_pos_target"] + s["motor_offsets"]
        vel = s["dof_vel"]
        xs = torch.stack((err, s["joint_pos_err_last"], s["joint_pos_err_last_last"], vel, s["joint_vel_last"],
                          s["joint_vel_last_last"]), -1)
        tq = (net or ActuatorNet())(xs.view(-1, 6)).view(-1, 12)
        s["joint_pos_err_last_last"] = s["joint_pos_err_last"].clone()
        s["joint_pos_err_last"] = err.clone()
        s["joint_vel_last_last"] = s["joint_vel_last"].clone()
        s["joint_vel_last"] = vel.clone()
    else:
        tq = P["kp"] * s["Kp_factors"] * (s["joint_pos_target"] - s["dof_pos"] + s["motor_offsets"]) \
             - P["kd"] * s["Kd_factors"] * s["dof_vel"]
    tq = tq * s["motor_strengths"]
    return torch.clip(tq, -P["torque_limit"], P["torque_limit"])


def step_contact_targets(s, P):
    cmd = s["commands"]
    f, ph, off, bnd, dur = cmd[:, 4], cmd[:, 5], cmd[:, 6], cmd[:, 7], cmd[:, 8]
    s["gait_indices"] = torch.remainder(s["gait_indices"] + P["dt"] * f, 1.0)
    g = s["gait_indices"]
    if P["pacing_offset"]:
        fi = [g + ph + off + bnd, g + bnd, g + off, g + ph]
    else:
        fi = [g + ph + off + bnd, g + off, g + bnd, g + ph]
    s["foot_indices"] = torch.remainder(torch.stack(fi, 1), 1.0)
    warped = []
    for idx in fi:
        r = torch.remainder(idx, 1)
        st, sw = r < dur, r > dur
        out = idx.clone()
        out[st] = r[st] * (0.5 / dur[st])
        out[sw] = 0.5 + (r[sw] - dur[sw]) * (0.5 / (1 - dur[sw]))
        warped.append(out)
    w = torch.stack(warped, 1)
    s["clock_inputs"] = torch.sin(2 * np.pi * w)
    s["doubletime_clock_inputs"] = torch.sin(4 * np.pi * w)
    s["halftime_clock_inputs"] = torch.sin(np.pi * w)
    kap = P["kappa_gait_probs"]
    cdf = torch.distributions.normal.Normal(0, kap).cdf
    r = torch.remainder(w, 1.0)
    s["desired_contact_states"] = cdf(r) * (1 - cdf(r - 0.5)) + cdf(r - 1) * (1 - cdf(r - 0.5 - 1))


def check_termination(s, P):
    reset = torch.any(torch.norm(s["contact_forces"][:, TERMINATION, :], dim=-1) > 1., dim=1)
    time_out = s["episode_length_buf"] > P["max_episode_length"]
    reset = reset | time_out
    if P["use_terminal_body_height"]:
        reset = reset | (s["root_states"][:, 2] < P["terminal_body_height"])
    return reset, time_out


def reward_terms(s, P):
    """All CoRLRewards terms (unscaled), keyed by name."""
    cf = s["contact_forces"]
    cmd, blv, bav, pg = s["commands"], s["base_lin_vel"], s["base_ang_vel"], s["projected_gravity"]
    q, qd = s["dof_pos"], s["dof_vel"]
    des = s["desired_contact_states"]
    fv, fp = s["foot_velocities"], s["foot_positions"]
    R = {}
    R["tracking_lin_vel"] = torch.exp(-torch.sum(torch.square(cmd[:, :2] - blv[:, :2]), 1) / P["tracking_sigma"])
    R["tracking_ang_vel"] = torch.exp(-torch.square(cmd[:, 2] - bav[:, 2]) / P["tracking_sigma_yaw"])
    R["lin_vel_z"] = torch.square(blv[:, 2])
    R["ang_vel_xy"] = torch.sum(torch.square(bav[:, :2]), 1)
    R["orientation"] = torch.sum(torch.square(pg[:, :2]), 1)
    R["torques"] = torch.sum(torch.square(s["torques"]), 1)
    R["dof_acc"] = torch.sum(torch.square((s["last_dof_vel"] - qd) / P["dt"]), 1)
    R["action_rate"] = torch.sum(torch.square(s["last_actions"] - s["actions"]), 1)
    R["collision"] = torch.sum(1. * (torch.norm(cf[:, PENALISED, :], dim=-1) > 0.1), 1)
    lim = -(q - P["dof_pos_limits"][:, 0]).clip(max=0.) + (q - P["dof_pos_limits"][:, 1]).clip(min=0.)
    R["dof_pos_limits"] = torch.sum(lim, 1)
    R["jump"] = -torch.square(s["root_states"][:, 2] - (cmd[:, 3] + P["base_height_target"]))
    ff = torch.norm(cf[:, FEET, :], dim=-1)
    R["tracking_contacts_shaped_force"] = sum(-(1 - des[:, i]) * (1 - torch.exp(-1 * ff[:, i] ** 2 / P["gait_force_sigma"])) for i in range(4)) / 4
    fvn = torch.norm(fv, dim=2)
    R["tracking_contacts_shaped_vel"] = sum(-(des[:, i] * (1 - torch.exp(-1 * fvn[:, i] ** 2 / P["gait_vel_sigma"]))) for i in range(4)) / 4
    R["dof_pos"] = torch.sum(torch.square(q - P["default_dof_pos"]), 1)
    R["dof_vel"] = torch.sum(torch.square(qd), 1)
    d1 = torch.square(s["joint_pos_target"] - s["last_joint_pos_target"]) * (s["last_actions"] != 0)
    R["action_smoothness_1"] = torch.sum(d1, 1)
    d2 = torch.square(s["joint_pos_target"] - 2 * s["last_joint_pos_target"] + s["last_last_joint_pos_target"])
    d2 = d2 * (s["last_actions"] != 0) * (s["last_last_actions"] != 0)
    R["action_smoothness_2"] = torch.sum(d2, 1)
    contact = cf[:, FEET, 2] > 1.
    filt = torch.logical_or(contact, s["last_contacts"])
    R["_new_last_contacts"] = contact
    R["feet_slip"] = torch.sum(filt * torch.square(torch.norm(fv[:, :, 0:2], dim=2)), 1)
    R["feet_contact_vel"] = torch.sum((fp[:, :, 2] < 0.03) * torch.square(torch.norm(fv, dim=2)), 1)
    R["feet_contact_forces"] = torch.sum((ff - P["max_contact_force"]).clip(min=0.), 1)
    phases = 1 - torch.abs(1.0 - torch.clip((s["foot_indices"] * 2.0) - 1.0, 0.0, 1.0) * 2.0)
    tgt = cmd[:, 9].unsqueeze(1) * phases + 0.02
    R["feet_clearance_cmd_linear"] = torch.sum(torch.square(tgt - fp[:, :, 2]) * (1 - des), 1)
    R["feet_impact_vel"] = torch.sum((ff > 1.0) * torch.square(torch.clip(s["prev_foot_velocities"][:, :, 2], -100, 0)), 1)
    # orientation_control
    dev = cmd.device
    qr = quat_from_angle_axis(-cmd[:, 11], torch.tensor([1., 0, 0], device=dev))
    qp = quat_from_angle_axis(-cmd[:, 10], torch.tensor([0., 1, 0], device=dev))
    dpg = quat_rotate_inverse(quat_mul(qr, qp), s["gravity_vec"])
    R["orientation_control"] = torch.sum(torch.square(pg[:, :2] - dpg[:, :2]), 1)
    # raibert heuristic
    rel = fp - s["root_states"][:, None, 0:3]
    qy = s["root_states"][:, 3:7].clone()
    qy[:, :3] *= -1                       # conjugate
    qy[:, :2] = 0.
    qy = qy / qy.norm(dim=-1, keepdim=True).clamp(min=1e-9)
    fb = torch.stack([quat_apply(qy, rel[:, i]) for i in range(4)], 1)
    wdt, ln = cmd[:, 12:13], cmd[:, 13:14]
    ys = torch.cat([wdt / 2, -wdt / 2, wdt / 2, -wdt / 2], 1)
    xs = torch.cat([ln / 2, ln / 2, -ln / 2, -ln / 2], 1)
    phs = torch.abs(1.0 - (s["foot_indices"] * 2.0)) * 1.0 - 0.5
    fr = cmd[:, 4]
    yv = cmd[:, 2:3] * ln / 2
    yo = phs * yv * (0.5 / fr.unsqueeze(1))
    yo[:, 2:4] *= -1
    xo = phs * cmd[:, 0:1] * (0.5 / fr.unsqueeze(1))
    err = torch.abs(torch.stack((xs + xo, ys + yo), 2) - fb[:, :, 0:2])
    R["raibert_heuristic"] = torch.sum(torch.square(err), dim=(1, 2))
    return R


def compute_reward(s, P):
    """Returns rew, rew_pos, rew_neg; updates episode_sums / command_sums / last_contacts in s."""
    R = reward_terms(s, P)
    n = s["commands"].shape[0]
    rew, pos, neg = torch.zeros(n), torch.zeros(n), torch.zeros(n)
    for name, scale in P["reward_scales"].items():
        if name == "termination":
            continue
        r = R[name] * scale
        rew += r
        if torch.sum(r) >= 0:
            pos += r
        elif torch.sum(r) <= 0:
            neg += r
        s["episode_sums"][name] = s["episode_sums"][name] + r
        if name in ("tracking_contacts_shaped_force", "tracking_contacts_shaped_vel"):
            s["command_sums"][name] = s["command_sums"][name] + scale + r
        else:
            s["command_sums"][name] = s["command_sums"][name] + r
        if name == "feet_slip":
            s["last_contacts"] = R["_new_last_contacts"]
    if P["only_positive_rewards"]:
        rew = torch.clip(rew, min=0.)
    elif P["only_positive_rewards_ji22_style"]:
        rew = pos * torch.exp(neg / P["sigma_rew_neg"])
    s["episode_sums"]["total"] = s["episode_sums"]["total"] + rew
    s["command_sums"]["lin_vel_raw"] = s["command_sums"]["lin_vel_raw"] + s["base_lin_vel"][:, 0]
    s["command_sums"]["ang_vel_raw"] = s["command_sums"]["ang_vel_raw"] + s["base_ang_vel"][:, 2]
    s["command_sums"]["lin_vel_residual"] = s["command_sums"]["lin_vel_residual"] + (s["base_lin_vel"][:, 0] - s["commands"][:, 0]) ** 2
    s["command_sums"]["ang_vel_residual"] = s["command_sums"]["ang_vel_residual"] + (s["base_ang_vel"][:, 2] - s["commands"][:, 2]) ** 2
    s["command_sums"]["ep_timesteps"] = s["command_sums"]["ep_timesteps"] + 1
    return rew, pos, neg


def compute_observations(s, P, noise_u=None):
    """obs [N,num_obs] (noise from injected uniforms `noise_u`), privileged obs; both clipped (legged_robot.py:84-87)."""
    obs = torch.cat((s["projected_gravity"], s["commands"] * P["commands_scale"],
                     (s["dof_pos"] - P["default_dof_pos"]) * P["obs_scale_dof_pos"], s["dof_vel"] * P["obs_scale_dof_vel"],
                     s["actions"]), -1) if P["observe_command"] else \
        torch.cat((s["projected_gravity"], (s["dof_pos"] - P["default_dof_pos"]) * P["obs_scale_dof_pos"],
                   s["dof_vel"] * P["obs_scale_dof_vel"], s["actions"]), -1)
    if P["observe_two_prev_actions"]:
        obs = torch.cat((obs, s["last_actions"]), -1)
    if P["observe_timing_parameter"]:
        obs = torch.cat((obs, s["gait_indices"].unsqueeze(1)), -1)
    if P["observe_clock_inputs"]:
        obs = torch.cat((obs, s["clock_inputs"]), -1)
    if P["observe_vel"]:
        obs = torch.cat((s["base_lin_vel"] * P["obs_scale_lin_vel"], s["base_ang_vel"] * P["obs_scale_ang_vel"], obs), -1)
    if P["observe_only_ang_vel"]:
        obs = torch.cat((s["base_ang_vel"] * P["obs_scale_ang_vel"], obs), -1)
    if P["observe_only_lin_vel"]:
        obs = torch.cat((s["base_lin_vel"] * P["obs_scale_lin_vel"], obs), -1)
    if P["observe_yaw"]:
        fwd = quat_apply(s["root_states"][:, 3:7], torch.tensor([1., 0, 0]).repeat(obs.shape[0], 1))
        obs = torch.cat((obs, torch.atan2(fwd[:, 1], fwd[:, 0]).unsqueeze(1)), -1)
    if P["observe_contact_states"]:
        obs = torch.cat((obs, (s["contact_forces"][:, FEET, 2] > 1.) * 1.0), 1)
    if P["add_noise"]:
        obs = obs + (2 * noise_u - 1) * P["noise_scale_vec"]
    priv = []
    col = lambda v: v.unsqueeze(1) if v.dim() == 1 else v
    for flag, ss, val in (("priv_friction", "friction_ss", lambda: col(s["friction_coeffs"])),
                          ("priv_restitution", "restitution_ss", lambda: col(s["restitutions"])),
                          ("priv_base_mass", "mass_ss", lambda: col(s["payloads"])),
                          ("priv_com_displacement", "com_ss", lambda: s["com_displacements"]),
                          ("priv_motor_strength", "motor_strength_ss", lambda: s["motor_strengths"]),
                          ("priv_motor_offset", "motor_offset_ss", lambda: s["motor_offsets"]),
                          ("priv_body_height", "body_height_ss", lambda: s["root_states"][:, 2:3]),
                          ("priv_body_velocity", "body_velocity_ss", lambda: s["base_lin_vel"])):
        if P.get(flag):
            sc, sh = P[ss]
            priv.append((val() - sh) * sc)
    if P.get("priv_gravity"):                  # legged_robot.py:466-472 divides by the scale
        sc, sh = P["gravity_ss"]
        priv.append((s["gravities"] - sh) / sc)
    if P.get("priv_clock_inputs"):
        priv.append(s["clock_inputs"])
    if P.get("priv_desired_contact_states"):
        priv.append(s["desired_contact_states"])
    priv = torch.cat(priv, 1) if priv else torch.zeros(obs.shape[0], 0)
    c = P["clip_obs"]
    return torch.clip(obs, -c, c), torch.clip(priv, -c, c)


def params_from_sim_config(c, active_scales, dt):
    """Oracle parameter dict from a Go1SimConfig (so kernel and oracle are driven by ONE resolved config)."""
    f = lambda a: torch.tensor(list(a), dtype=torch.float32)
    lim = torch.stack((f(c.soft_limit_lo), f(c.soft_limit_hi)), 1)
    return dict(
        action_scale=c.action_scale, hip_scale_reduction=c.hip_scale_reduction, use_lag=bool(c.use_lag),
        default_dof_pos=f(c.default_dof_pos).unsqueeze(0), control_type="actuator_net" if c.control_type == 0 else "P",
        kp=c.kp, kd=c.kd, torque_limit=c.torque_limit, dt=dt, pacing_offset=bool(c.pacing_offset),
        kappa_gait_probs=c.kappa_gait_probs, max_episode_length=c.max_episode_length,
        use_terminal_body_height=bool(c.use_terminal_body_height), terminal_body_height=c.terminal_body_height,
        tracking_sigma=c.tracking_sigma, tracking_sigma_yaw=c.tracking_sigma_yaw, dof_pos_limits=lim,
        base_height_target=c.base_height_target, gait_force_sigma=c.gait_force_sigma, gait_vel_sigma=c.gait_vel_sigma,
        max_contact_force=c.max_contact_force, reward_scales=dict(active_scales),
        only_positive_rewards=bool(c.only_positive_rewards), only_positive_rewards_ji22_style=bool(c.only_positive_rewards_ji22_style),
        sigma_rew_neg=c.sigma_rew_neg, observe_command=bool(c.observe_command), observe_two_prev_actions=bool(c.observe_two_prev_actions),
        observe_timing_parameter=bool(c.observe_timing_parameter), observe_clock_inputs=bool(c.observe_clock_inputs),
        observe_vel=bool(c.observe_vel), observe_only_ang_vel=bool(c.observe_only_ang_vel), observe_only_lin_vel=bool(c.observe_only_lin_vel),
        observe_yaw=bool(c.observe_yaw), observe_contact_states=bool(c.observe_contact_states), add_noise=bool(c.add_noise),
        commands_scale=f(c.commands_scale)[:c.num_commands], obs_scale_dof_pos=c.obs_scale_dof_pos, obs_scale_dof_vel=c.obs_scale_dof_vel,
        obs_scale_lin_vel=c.obs_scale_lin_vel, obs_scale_ang_vel=c.obs_scale_ang_vel, noise_scale_vec=f(c.noise_scale_vec)[:c.num_obs],
        clip_obs=c.clip_obs, clip_actions=c.clip_actions, num_train_envs=int(c.num_train_envs), rand_interval=int(c.rand_interval),
        custom_origins=bool(c.custom_origins), base_init_state=f(c.base_init_state),
        dr=[{k: (list(getattr(d, k)) if hasattr(getattr(d, k), "__len__") else getattr(d, k)) for k, _ in d._fields_} for d in c.dr],
        **{k: bool(getattr(c, k)) for k in ("priv_friction", "priv_restitution", "priv_base_mass", "priv_com_displacement", "priv_motor_strength",
                                            "priv_motor_offset", "priv_body_height", "priv_body_velocity", "priv_gravity", "priv_clock_inputs",
                                            "priv_desired_contact_states")},
        **{k: tuple(getattr(c, k)) for k in ("friction_ss", "restitution_ss", "mass_ss", "com_ss", "motor_strength_ss", "motor_offset_ss",
                                             "body_height_ss", "body_velocity_ss", "gravity_ss")})


# ---------------------------------------------------------------------------------------------------------------------
# domain randomisation, pushes, teleports and resets (legged_robot.py:150-239, 611-665, 948-1051).  Every torch.rand of the
# reference is replaced by the injected table U [N][48] (slot map: include/go1_b200.h, Go1SimBuffers.reset_rand); ranges are
# {low, float32(high - low)} as the resolved config stores them, evaluated as `u * span + low` in float32 like torch does.
# ---------------------------------------------------------------------------------------------------------------------
def _split(P, ids):
    """_call_train_eval (legged_robot.py:531-544): (ids, domain-rand table) for the train envs, then for the eval envs."""
    nt = P["num_train_envs"]
    return [(ids[ids < nt], P["dr"][0]), (ids[ids >= nt], P["dr"][1])]


def _draw(U, ids, slots, rng):
    return U[ids][:, slots] * rng[1] + rng[0]


def teleport_robots(s, P):
    n = s["root_states"].shape[0]
    for ids, D in _split(P, torch.arange(n)):
        if not D["teleport_robots"] or len(ids) == 0:
            continue
        rs = s["root_states"]
        lo = ids[rs[ids, 0] < D["teleport_x_lo"]]; rs[lo, 0] += D["teleport_dx"]
        hi = ids[rs[ids, 0] > D["teleport_x_hi"]]; rs[hi, 0] -= D["teleport_dx"]
        lo = ids[rs[ids, 1] < D["teleport_y_lo"]]; rs[lo, 1] += D["teleport_dy"]
        hi = ids[rs[ids, 1] > D["teleport_y_hi"]]; rs[hi, 1] -= D["teleport_dy"]


def push_robots(s, P, U):
    n = s["root_states"].shape[0]
    for ids, D in _split(P, torch.arange(n)):
        if not D["push_robots"] or len(ids) == 0:
            continue
        ids = ids[s["episode_length_buf"][ids] % int(D["push_interval"]) == 0]
        m = D["max_push_vel_xy"]
        s["root_states"][ids, 7:9] = (m - (-m)) * U[ids][:, [36, 37]] + (-m)


def randomize_dof_props(s, P, ids, U):
    for i, D in _split(P, ids):
        if len(i) == 0:
            continue
        if D["randomize_motor_strength"]:
            s["motor_strengths"][i, :] = _draw(U, i, [21], D["motor_strength_range"])
        if D["randomize_motor_offset"]:
            s["motor_offsets"][i, :] = _draw(U, i, list(range(24, 36)), D["motor_offset_range"])
        if D["randomize_Kp_factor"]:
            s["Kp_factors"][i, :] = _draw(U, i, [22], D["Kp_factor_range"])
        if D["randomize_Kd_factor"]:
            s["Kd_factors"][i, :] = _draw(U, i, [23], D["Kd_factor_range"])


def randomize_rigid_body_props(s, P, ids, U):
    for i, D in _split(P, ids):
        if len(i) == 0:
            continue
        if D["randomize_base_mass"]:
            s["payloads"][i] = _draw(U, i, [38], D["added_mass_range"])[:, 0]
        if D["randomize_com_displacement"]:
            s["com_displacements"][i, :] = _draw(U, i, [39, 40, 41], D["com_displacement_range"])
        if D["randomize_friction"]:
            s["friction_coeffs"][i] = _draw(U, i, [42], D["friction_range"])[:, 0]
        if D["randomize_restitution"]:
            s["restitutions"][i] = _draw(U, i, [43], D["restitution_range"])[:, 0]


def reset_idx(s, P, ids, new_commands, U):
    """legged_robot.py:150-239 for `ids` (tensor, ascending); the curriculum's part is `new_commands` [N][15].
    Returns (train/episode reward means dict or None, whether an eval env finished)."""
    if len(ids) == 0:
        return None, False
    s["commands"][ids] = new_commands[ids]                                 # _resample_commands: new commands, sums cleared
    for k in s["command_sums"]:
        s["command_sums"][k][ids] = 0.
    randomize_dof_props(s, P, ids, U)
    for i, D in _split(P, ids):
        if D["randomize_rigids_after_start"]:
            randomize_rigid_body_props(s, P, i, U)
    for i, D in _split(P, ids):                                            # _reset_dofs, _reset_root_states
        if len(i) == 0:
            continue
        s["dof_pos"][i] = P["default_dof_pos"] * ((1.5 - 0.5) * U[i][:, :12] + 0.5)
        s["dof_vel"][i] = 0.
    for i, D in _split(P, ids):
        if len(i) == 0:
            continue
        rs = s["root_states"]
        rs[i] = P["base_init_state"]
        rs[i, :3] += s["env_origins"][i]
        if P["custom_origins"]:
            rs[i, 0:1] += (D["x_init_range"] - (-D["x_init_range"])) * U[i][:, 12:13] + (-D["x_init_range"])
            rs[i, 1:2] += (D["y_init_range"] - (-D["y_init_range"])) * U[i][:, 13:14] + (-D["y_init_range"])
            rs[i, 0] += D["x_init_offset"]
            rs[i, 1] += D["y_init_offset"]
        yaw = (D["yaw_init_range"] - (-D["yaw_init_range"])) * U[i][:, 14:15] + (-D["yaw_init_range"])
        rs[i, 3:7] = quat_from_angle_axis(yaw[:, 0], torch.tensor([0., 0., 1.]))
        rs[i, 7:13] = (0.5 - (-0.5)) * U[i][:, 15:21] + (-0.5)
    for k in ("last_actions", "last_last_actions", "last_dof_vel"):
        s[k][ids] = 0.
    s["episode_length_buf"][ids] = 0
    nt = P["num_train_envs"]
    tr, ev = ids[ids < nt], ids[ids >= nt]
    means = None
    if len(tr) > 0:
        means = {}
        for k in s["episode_sums"]:
            means["rew_" + k] = torch.mean(s["episode_sums"][k][tr])
            s["episode_sums"][k][tr] = 0.
    if len(ev) > 0:
        for k in s["episode_sums"]:
            unset = ev[s["episode_sums_eval"][k][ev] == -1]
            s["episode_sums_eval"][k][unset] = s["episode_sums"][k][unset]
            s["episode_sums"][k][ev] = 0.
    s["gait_indices"][ids] = 0
    for b in s["lag_buffer"]:
        b[ids, :] = 0
    return means, len(ev) > 0


def post_physics_step(s, P, U_step, U_reset, noise_u, new_commands):
    """legged_robot.py:90-136 on physics outputs already in `s` (root_states, dof state, contact forces, foot kinematics):
    counters, base-frame quantities, callback (teleport, gait clock, push, periodic re-randomisation), termination, reward,
    reset_idx, observations, last_* rolls.  Returns dict(obs, priv, rew, reset, time_out, reset_ids, episode_means)."""
    s["episode_length_buf"] = s["episode_length_buf"] + 1
    quat = s["root_states"][:, 3:7]
    s["base_lin_vel"] = quat_rotate_inverse(quat, s["root_states"][:, 7:10])
    s["base_ang_vel"] = quat_rotate_inverse(quat, s["root_states"][:, 10:13])
    s["projected_gravity"] = quat_rotate_inverse(quat, s["gravity_vec"])
    teleport_robots(s, P)
    step_contact_targets(s, P)
    push_robots(s, P, U_step)
    if P["rand_interval"] > 0:
        due = torch.nonzero(s["episode_length_buf"] % P["rand_interval"] == 0).flatten()
        randomize_dof_props(s, P, due, U_step)
        for i, D in _split(P, due):
            if D["randomize_rigids_after_start"]:
                randomize_rigid_body_props(s, P, i, U_step)
    reset, time_out = check_termination(s, P)
    rew, pos, neg = compute_reward(s, P)
    ids = torch.nonzero(reset).flatten()
    means, _ = reset_idx(s, P, ids, new_commands, U_reset)
    obs, priv = compute_observations(s, P, noise_u)
    s["last_last_actions"] = s["last_actions"].clone(); s["last_actions"] = s["actions"].clone()
    s["last_last_joint_pos_target"] = s["last_joint_pos_target"].clone(); s["last_joint_pos_target"] = s["joint_pos_target"].clone()
    s["last_dof_vel"] = s["dof_vel"].clone()
    out_reset = reset.clone(); out_reset[ids] = True
    return dict(obs=obs, priv=priv, rew=rew, rew_pos=pos, rew_neg=neg, reset=out_reset, time_out=time_out, reset_ids=ids, episode_means=means)
