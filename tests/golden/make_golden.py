#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ by running the REFERENCE's own Python
(/root/reference, imported through tests/_stubs) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (committed):
  cfg_trees.json     Cfg defaults / after config_go1 / after scripts/train.py's overrides (key order kept)
  env_logic.npz      inputs + outputs of LeggedRobot._compute_torques (x2), _step_contact_targets,
                     check_termination, compute_reward, compute_observations on a mock env (N=64)
  kats.npz           known-answer tests of SURVEY.md §8c: actuator net, gait clock, curriculum, policy MLPs
  ppo.npz            one full ppo_cse act->process_env_step->compute_returns->update cycle (BASELINE config 1)
  terrain.npz        height fields + env origins built by the reference's Terrain class (curriculum / randomised / train.py
                     modes) on top of this repository's terrain_utils restatement (tests/_stubs/isaacgym/terrain_utils.py)
  resample.npz       14 rounds of LeggedRobot._resample_commands (curriculum update + sampling + gait remap) on a mock
                     env with scripts/train.py's config (N=64), incl. the torch.rand category draws of every round
"""
import io
import json
import os
import pickle
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tests", "_stubs"))
sys.path.insert(1, REF)
np.int = int            # the reference pins numpy 1.23 (legged_robot.py:1362)
import torch  # noqa: E402

torch.set_num_threads(1)


def clean(d):
    return {k: v for k, v in dict(d).items() if not k.startswith("_")}


def tree(C):
    out = {}
    for s, v in clean(vars(C)).items():
        if isinstance(v, type):
            out[s] = {k: (clean(vars(x)) if isinstance(x, type) else x) for k, x in clean(vars(v)).items()}
    return out


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    return o


def reference_train_cfg():
    """Cfg after config_go1 + the assignments of scripts/train.py (exec'd from the reference source)."""
    from go1_gym.envs.base.legged_robot_config import Cfg
    from go1_gym.envs.go1.go1_config import config_go1
    trees = {"defaults": jsonable(tree(Cfg))}
    config_go1(Cfg)
    trees["go1"] = jsonable(tree(Cfg))
    src = open(f"{REF}/scripts/train.py").read().splitlines()
    lines = [l.strip() for l in src if re.match(r"\s*Cfg\.\w+\.\w+\s*=", l)]
    exec("\n".join(lines), {"Cfg": Cfg})
    trees["train"] = jsonable(tree(Cfg))
    return Cfg, trees


def mock_env(Cfg, N, g):
    """A namespace carrying every attribute the reference's step-path methods read, filled with seeded state."""
    from go1_gym.envs.base.legged_robot import LeggedRobot
    R = lambda *s, lo=-1.0, hi=1.0: torch.rand(*s, generator=g) * (hi - lo) + lo
    dt = 4 * float(np.float32(0.005))
    env = types.SimpleNamespace()
    env.cfg = Cfg
    env.dt = dt
    env.num_envs = N
    env.num_dof = env.num_dofs = env.num_actions = env.num_actuated_dof = 12
    env.device = "cpu"
    env.num_bodies = 17
    env.obs_scales = Cfg.obs_scales
    env.feet_indices = torch.tensor([4, 8, 12, 16])
    env.penalised_contact_indices = torch.tensor([2, 6, 10, 14, 3, 7, 11, 15])
    env.termination_contact_indices = torch.tensor([0])
    names = [f"{l}_{p}_joint" for l in ("FL", "FR", "RL", "RR") for p in ("hip", "thigh", "calf")]
    env.default_dof_pos = torch.tensor([Cfg.init_state.default_joint_angles[n] for n in names]).unsqueeze(0)
    lim = torch.tensor([[-0.802851455917, 0.802851455917], [-1.0471975512, 4.18879020479], [-2.69653369433, -0.916297857297]] * 4)
    m = (lim[:, 0] + lim[:, 1]) / 2
    r = lim[:, 1] - lim[:, 0]
    env.dof_pos_limits = torch.stack((m - 0.5 * r * Cfg.rewards.soft_dof_pos_limit, m + 0.5 * r * Cfg.rewards.soft_dof_pos_limit), 1)
    env.torque_limits = torch.full((12,), 33.5)
    # ---- state
    quat = torch.randn(N, 4, generator=g) * torch.tensor([0.15, 0.15, 1.0, 1.0])
    quat = quat / quat.norm(dim=1, keepdim=True)
    env.root_states = torch.cat((R(N, 2, lo=-3, hi=3), R(N, 1, lo=0.02, hi=0.45), quat, R(N, 3, lo=-1.5, hi=1.5), R(N, 3, lo=-2, hi=2)), 1)
    env.base_pos = env.root_states[:, 0:3]
    env.base_quat = env.root_states[:, 3:7]
    env.dof_pos = env.default_dof_pos + R(N, 12, lo=-0.9, hi=0.9)
    env.dof_vel = R(N, 12, lo=-8, hi=8)
    env.actions = R(N, 12, lo=-3, hi=3)
    env.last_actions = R(N, 12, lo=-3, hi=3)
    env.last_actions[:5] = 0.0
    env.last_actions[5:10, ::2] = 0.0
    env.last_last_actions = R(N, 12, lo=-3, hi=3)
    env.last_last_actions[3:8] = 0.0
    env.last_dof_vel = R(N, 12, lo=-8, hi=8)
    env.last_joint_pos_target = env.default_dof_pos + R(N, 12, lo=-0.5, hi=0.5)
    env.last_last_joint_pos_target = env.default_dof_pos + R(N, 12, lo=-0.5, hi=0.5)
    env.lag_buffer = [R(N, 12, lo=-0.6, hi=0.6) for _ in range(7)]
    env.joint_pos_err_last = R(N, 12, lo=-0.5, hi=0.5)
    env.joint_pos_err_last_last = R(N, 12, lo=-0.5, hi=0.5)
    env.joint_vel_last = R(N, 12, lo=-8, hi=8)
    env.joint_vel_last_last = R(N, 12, lo=-8, hi=8)
    env.motor_offsets = R(N, 12, lo=-0.02, hi=0.02)
    env.motor_strengths = R(N, 1, lo=0.9, hi=1.1).repeat(1, 12)
    env.Kp_factors = torch.ones(N, 12)
    env.Kd_factors = torch.ones(N, 12)
    env.p_gains = torch.full((12,), 20.0)
    env.d_gains = torch.full((12,), 0.5)
    env.friction_coeffs = R(N, 1, lo=0.1, hi=3.0).repeat(1, 4)
    env.restitutions = R(N, 1, lo=0.0, hi=0.4).repeat(1, 4)
    env.payloads = R(N, lo=-1, hi=3)
    env.com_displacements = torch.zeros(N, 3)
    env.gravities = torch.zeros(N, 3)
    gv = torch.tensor([0.05, -0.03, -9.8])
    env.gravity_vec = (gv / gv.norm()).repeat(N, 1)
    cmd_lo = torch.tensor([-1, -0.6, -1, -0.25, 2.0, 0, 0, 0, 0.5, 0.03, -0.4, 0.0, 0.10, 0.35, 0.0])
    cmd_hi = torch.tensor([1, 0.6, 1, 0.15, 4.0, 1, 1, 1, 0.5, 0.35, 0.4, 0.0, 0.45, 0.45, 0.01])
    env.commands = torch.rand(N, 15, generator=g) * (cmd_hi - cmd_lo) + cmd_lo
    env.commands[:, 5:8] = torch.round(2 * env.commands[:, 5:8]) / 2.0 % 1
    env.commands[::7, 8] = 0.35                      # a few non-default stance durations
    env.gait_indices = torch.rand(N, generator=g)
    env.gait_indices[0] = 0.0
    env.clock_inputs = torch.zeros(N, 4)
    env.doubletime_clock_inputs = torch.zeros(N, 4)
    env.halftime_clock_inputs = torch.zeros(N, 4)
    env.desired_contact_states = torch.zeros(N, 4)
    env.contact_forces = torch.zeros(N, 17, 3)
    env.contact_forces[:, [4, 8, 12, 16]] = R(N, 4, 3, lo=-15, hi=15) * (torch.rand(N, 4, 1, generator=g) > 0.4)
    env.contact_forces[:, [4, 8, 12, 16], 2] = env.contact_forces[:, [4, 8, 12, 16], 2].abs() * 6
    env.contact_forces[:, [2, 6, 10, 14, 3, 7, 11, 15]] = R(N, 8, 3, lo=-2, hi=2) * (torch.rand(N, 8, 1, generator=g) > 0.8)
    env.contact_forces[:, 0] = R(N, 3, lo=-3, hi=3) * (torch.rand(N, 1, generator=g) > 0.8)
    env.foot_positions = torch.cat((R(N, 4, 2, lo=-3.5, hi=3.5), R(N, 4, 1, lo=0.0, hi=0.15)), 2)
    env.foot_positions[:, :, :2] = env.root_states[:, None, :2] + R(N, 4, 2, lo=-0.35, hi=0.35)
    env.foot_velocities = R(N, 4, 3, lo=-2, hi=2)
    env.prev_foot_velocities = R(N, 4, 3, lo=-2, hi=2)
    env.last_contacts = torch.rand(N, 4, generator=g) > 0.5
    env.episode_length_buf = torch.randint(0, 1010, (N,), generator=g)
    env.episode_length_buf[:3] = torch.tensor([1001, 1002, 1000])
    env.measured_heights = 0
    env.add_noise = True
    env.commands_scale = torch.tensor([2.0, 2.0, 0.25, 2.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.15, 0.3, 0.3, 1.0, 1.0, 1.0])
    env.noise_scale_vec = LeggedRobot._get_noise_scale_vec(env, Cfg)
    env.torques = torch.zeros(N, 12)
    env.joint_pos_target = torch.zeros(N, 12)
    env.rew_buf = torch.zeros(N); env.rew_buf_pos = torch.zeros(N); env.rew_buf_neg = torch.zeros(N)
    env.sim_params = types.SimpleNamespace(dt=float(np.float32(Cfg.sim.dt)))      # gymapi.SimParams.dt is a C float
    LeggedRobot._parse_cfg(env, Cfg)                                              # derives max_episode_length etc.
    assert abs(env.dt - dt) < 1e-12
    env.reward_scales = clean(vars(Cfg.reward_scales))
    env.curriculum_thresholds = clean(vars(Cfg.curriculum_thresholds))

    # actuator net evaluation closure as in _init_buffers (legged_robot.py:1238-1253)
    net = torch.jit.load(f"{REF}/resources/actuator_nets/unitree_go1.pt", map_location="cpu")

    def eval_actuator_network(jp, jpl, jpll, jv, jvl, jvll):
        xs = torch.cat((jp.unsqueeze(-1), jpl.unsqueeze(-1), jpll.unsqueeze(-1), jv.unsqueeze(-1), jvl.unsqueeze(-1), jvll.unsqueeze(-1)), dim=-1)
        return net(xs.view(N * 12, 6)).view(N, 12)
    env.actuator_network = eval_actuator_network
    return env


def make_env_logic(Cfg):
    from go1_gym.envs.base.legged_robot import LeggedRobot
    from isaacgym.torch_utils import quat_rotate_inverse

    N = 64
    g = torch.Generator().manual_seed(1234)
    env = mock_env(Cfg, N, g)
    out = {}

    def snap(prefix, names):
        for n in names:
            v = getattr(env, n)
            if isinstance(v, list):
                v = torch.stack(v)
            out[f"{prefix}/{n}"] = v.detach().clone().numpy()

    state_in = ["root_states", "dof_pos", "dof_vel", "actions", "last_actions", "last_last_actions", "last_dof_vel",
                "last_joint_pos_target", "last_last_joint_pos_target", "lag_buffer", "joint_pos_err_last", "joint_pos_err_last_last",
                "joint_vel_last", "joint_vel_last_last", "motor_offsets", "motor_strengths", "friction_coeffs", "restitutions",
                "payloads", "gravity_vec", "commands", "gait_indices", "contact_forces", "foot_positions", "foot_velocities",
                "prev_foot_velocities", "last_contacts", "episode_length_buf", "noise_scale_vec", "dof_pos_limits"]
    snap("in", state_in)
    with torch.no_grad():
        # ---- two control substeps (legged_robot.py:74-80 without the physics)
        for sub in range(2):
            tq = LeggedRobot._compute_torques(env, env.actions).view(N, 12)
            out[f"torques/sub{sub}"] = tq.numpy().copy()
            out[f"joint_pos_target/sub{sub}"] = env.joint_pos_target.numpy().copy()
        env.torques = tq
        snap("after_torques", ["lag_buffer", "joint_pos_err_last", "joint_pos_err_last_last", "joint_vel_last", "joint_vel_last_last"])
        # ---- post physics quantities (legged_robot.py:106-110)
        env.base_lin_vel = quat_rotate_inverse(env.base_quat, env.root_states[:, 7:10])
        env.base_ang_vel = quat_rotate_inverse(env.base_quat, env.root_states[:, 10:13])
        env.projected_gravity = quat_rotate_inverse(env.base_quat, env.gravity_vec)
        snap("post", ["base_lin_vel", "base_ang_vel", "projected_gravity"])
        LeggedRobot._step_contact_targets(env)
        snap("gait", ["gait_indices", "foot_indices", "clock_inputs", "doubletime_clock_inputs", "halftime_clock_inputs", "desired_contact_states"])
        LeggedRobot.check_termination(env)
        out["term/reset_buf"] = env.reset_buf.numpy().copy()
        out["term/time_out_buf"] = env.time_out_buf.numpy().copy()
        LeggedRobot._prepare_reward_function(env)
        scales = dict(env.reward_scales)
        out["reward/names"] = np.array(env.reward_names)
        out["reward/scales"] = np.array([scales[n] for n in env.reward_names])
        for n, fn in zip(env.reward_names, env.reward_functions):
            if n == "feet_slip":
                keep = env.last_contacts.clone()
            out[f"reward_raw/{n}"] = fn().numpy().copy()
            if n == "feet_slip":
                env.last_contacts = keep
        LeggedRobot.compute_reward(env)
        out["reward/rew_buf"] = env.rew_buf.numpy().copy()
        out["reward/rew_buf_pos"] = env.rew_buf_pos.numpy().copy()
        out["reward/rew_buf_neg"] = env.rew_buf_neg.numpy().copy()
        out["reward/last_contacts"] = env.last_contacts.numpy().copy()
        for k, v in env.episode_sums.items():
            out[f"episode_sums/{k}"] = v.numpy().copy()
        for k, v in env.command_sums.items():
            out[f"command_sums/{k}"] = v.numpy().copy()
        # ---- observations with injected noise: torch.rand_like is replaced for the call
        u = torch.rand(N, 70, generator=g)
        out["obs/noise_u"] = u.numpy().copy()
        orig = torch.rand_like
        torch.rand_like = lambda t, **k: u
        try:
            LeggedRobot.compute_observations(env)
        finally:
            torch.rand_like = orig
        out["obs/obs_buf"] = torch.clip(env.obs_buf, -100, 100).numpy().copy()
        out["obs/privileged_obs_buf"] = torch.clip(env.privileged_obs_buf, -100, 100).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "env_logic.npz"), **out)
    print("env_logic.npz:", len(out), "arrays")


def clone_cfg(C, name="CfgClone"):
    """A deep copy of a Cfg class tree (sections are plain classes under the params_proto stub)."""
    import copy
    sections = {}
    for k, v in clean(vars(C)).items():
        if isinstance(v, types.MappingProxyType):       # Cfg.command_ranges = vars(Cfg.commands), re-derived by _parse_cfg
            continue
        if isinstance(v, type):
            inner = {}
            for kk, vv in clean(vars(v)).items():
                inner[kk] = type(kk, (), {a: copy.deepcopy(b) for a, b in clean(vars(vv)).items()}) if isinstance(vv, type) else copy.deepcopy(vv)
            sections[k] = type(k, (), inner)
        else:
            sections[k] = copy.deepcopy(v)
    return type(name, (), sections)


# slot of every uniform draw in the [N][48] table the CUDA kernels read (include/go1_b200.h: Go1SimBuffers.reset_rand)
def _draw_slots(name, cfg, custom_origins):
    dr = cfg.domain_rand
    if name == "_push_robots":
        return [[36, 37]] if dr.push_robots else []
    if name == "_randomize_dof_props":
        return ([[21]] if dr.randomize_motor_strength else []) + ([list(range(24, 36))] if dr.randomize_motor_offset else []) + \
               ([[22]] if dr.randomize_Kp_factor else []) + ([[23]] if dr.randomize_Kd_factor else [])
    if name == "_randomize_rigid_body_props":
        return ([[38]] if dr.randomize_base_mass else []) + ([[39, 40, 41]] if dr.randomize_com_displacement else []) + \
               ([[42]] if dr.randomize_friction else []) + ([[43]] if dr.randomize_restitution else [])
    if name == "_reset_dofs":
        return [list(range(12))]
    if name == "_reset_root_states":
        return ([[12], [13]] if custom_origins else []) + [[14], list(range(15, 21))]
    return []


DR_OVERRIDES = dict(
    domain_rand=dict(randomize_rigids_after_start=True, randomize_com_displacement=True, com_displacement_range=[-0.15, 0.15],
                     randomize_restitution=True, restitution_range=[0.0, 0.4], randomize_Kp_factor=True, Kp_factor_range=[0.8, 1.3],
                     randomize_Kd_factor=True, Kd_factor_range=[0.5, 1.5], push_robots=True, push_interval_s=3, max_push_vel_xy=0.5),
    terrain=dict(teleport_robots=True, teleport_thresh=0.3, num_rows=4, num_cols=5, terrain_length=5.0, terrain_width=5.0,
                 x_init_range=0.2, y_init_range=0.2, yaw_init_range=3.14, x_init_offset=0.1, y_init_offset=-0.05),
    env=dict(priv_observe_friction=True, priv_observe_restitution=True, priv_observe_base_mass=True, priv_observe_com_displacement=True,
             priv_observe_motor_strength=True, priv_observe_motor_offset=True, priv_observe_body_height=True, priv_observe_body_velocity=True,
             priv_observe_gravity=True, priv_observe_clock_inputs=True, priv_observe_desired_contact_states=True, num_privileged_obs=45))
EVAL_OVERRIDES = dict(
    domain_rand=dict(motor_strength_range=[0.7, 0.8], motor_offset_range=[-0.05, 0.01], Kp_factor_range=[1.0, 1.1], Kd_factor_range=[0.9, 1.0],
                     added_mass_range=[2.0, 4.0], com_displacement_range=[0.0, 0.05], friction_range=[0.05, 0.2], restitution_range=[0.3, 0.5],
                     push_interval_s=2, max_push_vel_xy=1.5, randomize_Kd_factor=False),
    terrain=dict(teleport_thresh=0.5, num_rows=3, num_cols=4, x_init_range=0.5, y_init_range=0.1, yaw_init_range=1.0, x_init_offset=-0.3,
                 y_init_offset=0.2))


def apply_overrides(C, over):
    for sec, kv in over.items():
        for k, v in kv.items():
            setattr(getattr(C, sec), k, v)
    return C


def make_dr_step(Cfg0):
    """the reference's whole post_physics_step -- teleport, push, periodic re-randomisation, termination,
    rewards, reset_idx (_randomize_dof_props, _randomize_rigid_body_props, _reset_dofs, _reset_root_states, buffer clears,
    extras), compute_observations, last_* rolls -- on a mock env with full domain randomisation and a train/eval split
    (40 + 24 envs, different ranges), every torch.rand draw recorded in the slot the CUDA kernels read it from."""
    import math
    from go1_gym.envs.base.legged_robot import LeggedRobot
    import isaacgym.gymtorch as gymtorch
    gymtorch.unwrap_tensor = lambda t: t
    NT, NE = 40, 24
    N = NT + NE
    Cfg = apply_overrides(clone_cfg(Cfg0, "CfgDR"), DR_OVERRIDES)
    ECfg = apply_overrides(clone_cfg(Cfg, "CfgDREval"), EVAL_OVERRIDES)
    Cfg.terrain.x_offset, Cfg.terrain.rows_offset = 0, 0                     # set by Terrain.__init__ (go1_gym/utils/terrain.py:42-51)
    ECfg.terrain.x_offset, ECfg.terrain.rows_offset = 250, Cfg.terrain.num_rows
    g = torch.Generator().manual_seed(4321)
    env = mock_env(Cfg, N, g)

    class Env:                                    # unbound reference methods become bound methods of the mock
        def __getattr__(self, name):
            f = getattr(LeggedRobot, name, None)
            if callable(f):
                return types.MethodType(f, self)
            raise AttributeError(name)
    e = Env()
    e.__dict__.update(vars(env))
    env = e
    env.eval_cfg = ECfg
    env.num_train_envs, env.num_eval_envs = NT, NE
    LeggedRobot._parse_cfg(env, ECfg)
    LeggedRobot._parse_cfg(env, Cfg)
    env.reward_scales = clean(vars(Cfg.reward_scales))
    env.curriculum_thresholds = clean(vars(Cfg.curriculum_thresholds))
    R = lambda *s, lo=-1.0, hi=1.0: torch.rand(*s, generator=g) * (hi - lo) + lo

    class AnyCall:
        def __getattr__(self, name):
            return lambda *a, **k: None
    env.gym, env.sim, env.viewer, env.record_now, env.debug_viz, env.enable_viewer_sync = AnyCall(), None, None, False, False, False
    env._render_headless = lambda: None
    env._randomize_gravity = lambda *a, **k: None
    env.refresh_actor_rigid_shape_props = lambda ids, cfg: None      # pushes friction / restitution into PhysX: no arithmetic
    env.common_step_counter = 5
    env.custom_origins = True
    env.base_init_state = torch.tensor(list(Cfg.init_state.pos) + list(Cfg.init_state.rot) + list(Cfg.init_state.lin_vel) + list(Cfg.init_state.ang_vel))
    env.env_origins = torch.cat((R(N, 2, lo=0.0, hi=18.0), torch.zeros(N, 1)), 1)
    env.terrain_levels = torch.zeros(N, dtype=torch.long)
    # robots spread over the tile grids, some beyond the teleport thresholds (train grid 20 x 25 m; eval grid offset by x_offset * hscale)
    xo = torch.cat((torch.zeros(NT), torch.full((NE,), float(int(ECfg.terrain.x_offset * ECfg.terrain.horizontal_scale)))))
    env.root_states[:, 0] = R(N, lo=-0.5, hi=20.5) * torch.cat((torch.ones(NT), torch.full((NE,), 0.75))) + xo
    env.root_states[:, 1] = R(N, lo=-0.5, hi=25.5) * torch.cat((torch.ones(NT), torch.full((NE,), 0.8)))
    env.foot_positions[:, :, :2] = env.root_states[:, None, :2] + R(N, 4, 2, lo=-0.35, hi=0.35)
    env.com_displacements = R(N, 3, lo=-0.1, hi=0.1)
    env.Kp_factors = R(N, 1, lo=0.8, hi=1.3).repeat(1, 12)
    env.Kd_factors = R(N, 1, lo=0.5, hi=1.5).repeat(1, 12)
    env.gravities = torch.tensor([0.05, -0.03, 0.0]).repeat(N, 1)
    # episode lengths: pushes every 151 (train) / 101 (eval) steps, re-randomisation every 201, time-outs above 1001
    ep = torch.randint(1, 990, (N,), generator=g)
    ep[[1, 7, 13, 22]] = torch.tensor([150, 301, 452, 905])           # train pushes at 151, 302, 453, 906
    ep[[41, 44, 50]] = torch.tensor([100, 201, 504])                 # eval pushes at 101, 202, 505
    ep[[3, 9, 30, 45, 52]] = torch.tensor([200, 401, 803, 602, 200])  # re-randomisation at multiples of 201
    ep[[5, 47]] = torch.tensor([1001, 1003])                         # time-outs
    ep[11] = 603 - 1                                                  # 603 = 3*201: re-randomised, and (below) terminated in the same step
    env.episode_length_buf = ep
    env.contact_forces[11, 0] = torch.tensor([0.0, 3.0, 4.0])
    env.feet_air_time = R(N, 4, lo=0.0, hi=0.4)
    env.last_root_vel = torch.zeros(N, 6)
    env.base_lin_vel, env.base_ang_vel, env.projected_gravity = torch.zeros(N, 3), torch.zeros(N, 3), torch.zeros(N, 3)
    env.reset_buf = torch.ones(N, dtype=torch.long)
    env.time_out_buf = torch.zeros(N, dtype=torch.bool)
    env.extras = {}
    rbs = torch.zeros(N, 17, 13)
    rbs[:, [4, 8, 12, 16], 0:3] = env.foot_positions
    rbs[:, [4, 8, 12, 16], 7:10] = env.foot_velocities
    env.rigid_body_state = rbs.view(N * 17, 13)
    env.dof_state = torch.zeros(N * 12, 2)
    env.torques = R(N, 12, lo=-20, hi=20)
    env.joint_pos_target = env.default_dof_pos + R(N, 12, lo=-0.5, hi=0.5)
    env.noise_scale_vec = LeggedRobot._get_noise_scale_vec(env, Cfg)
    env.env_command_bins = np.arange(N)
    env.curricula, env.category_names = [], []
    # new commands the (stubbed) curriculum hands to resetting envs; command sums cleared like legged_robot.py:822-824
    cmd_lo = torch.tensor([-1, -0.6, -1, -0.25, 2.0, 0, 0, 0, 0.5, 0.03, -0.4, 0.0, 0.10, 0.35, 0.0])
    cmd_hi = torch.tensor([1, 0.6, 1, 0.15, 4.0, 1, 1, 1, 0.5, 0.35, 0.4, 0.0, 0.45, 0.45, 0.01])
    new_cmd = torch.rand(N, 15, generator=g) * (cmd_hi - cmd_lo) + cmd_lo
    resample_calls = []

    def resample(ids):
        resample_calls.append(ids.clone())
        if len(resample_calls) == 2:          # call 1 = the periodic resample of the callback (left alone), call 2 = reset_idx
            env.commands[ids] = new_cmd[ids]
            for k in env.command_sums:
                env.command_sums[k][ids] = 0.
    env._resample_commands = resample
    LeggedRobot._prepare_reward_function(env)
    for k in env.episode_sums:
        env.episode_sums[k] = R(N, lo=0.0, hi=2.0)
    for k in env.command_sums:
        env.command_sums[k] = R(N, lo=0.0, hi=2.0)
    for k in env.episode_sums_eval:
        env.episode_sums_eval[k][NT + 3] = 0.25                        # one eval env already holds a finished episode

    out = {}
    state_in = ["root_states", "dof_pos", "dof_vel", "actions", "last_actions", "last_last_actions", "last_dof_vel",
                "last_joint_pos_target", "last_last_joint_pos_target", "lag_buffer", "motor_offsets", "motor_strengths", "Kp_factors", "Kd_factors",
                "friction_coeffs", "restitutions", "payloads", "com_displacements", "gravities", "gravity_vec", "commands", "gait_indices",
                "contact_forces", "foot_positions", "foot_velocities", "prev_foot_velocities", "last_contacts", "episode_length_buf",
                "torques", "joint_pos_target", "env_origins"]

    def snap(prefix, names):
        for n in names:
            v = getattr(env, n)
            v = torch.stack(v) if isinstance(v, list) else v
            out[f"{prefix}/{n}"] = v.detach().clone().numpy()
    snap("in", state_in)
    for k, v in env.episode_sums.items():
        out[f"in/episode_sums/{k}"] = v.numpy().copy()
    for k, v in env.command_sums.items():
        out[f"in/command_sums/{k}"] = v.numpy().copy()
    out["in/new_commands"] = new_cmd.numpy().copy()

    # ---- recorder: every torch.rand of the randomisation / reset methods lands in its kernel slot
    U = {"step": torch.full((N, 48), 0.5), "reset": torch.full((N, 48), 0.5)}
    active = []
    orig_rand, orig_rand_like = torch.rand, torch.rand_like

    def rec_rand(*shape, **kw):
        kw.pop("device", None); kw.pop("requires_grad", None)
        v = orig_rand(*shape, generator=g, **kw)
        if active:
            plane, ids, slots = active[0]
            sl = slots.pop(0)
            assert v.shape[0] == len(ids) and v.numel() == len(ids) * len(sl), (v.shape, len(ids), sl)
            U[plane][ids.unsqueeze(1), torch.tensor(sl).unsqueeze(0)] = v.reshape(len(ids), len(sl))
        return v

    phase = ["step"]

    def hook(name):
        def f(ids, cfg):
            use = ids
            if name == "_push_robots" and cfg.domain_rand.push_robots:
                use = ids[env.episode_length_buf[ids] % int(cfg.domain_rand.push_interval) == 0]
            active.insert(0, (phase[0], use, _draw_slots(name, cfg, env.custom_origins)))
            try:
                return getattr(LeggedRobot, name)(env, ids, cfg)
            finally:
                assert not active[0][2] or len(use) == 0, (name, active[0][2])
                active.pop(0)
        return f
    for name in ("_push_robots", "_randomize_dof_props", "_randomize_rigid_body_props", "_reset_dofs", "_reset_root_states"):
        setattr(env, name, hook(name))
    reset_orig = LeggedRobot.reset_idx

    def reset_idx(ids):
        phase[0] = "reset"
        out["reset/ids"] = ids.numpy().copy()
        out["mid/episode_sums_total"] = env.episode_sums["total"].numpy().copy()
        reset_orig(env, ids)
        phase[0] = "step"
    env.reset_idx = reset_idx
    noise_u = orig_rand(N, 70, generator=g)
    out["obs/noise_u"] = noise_u.numpy().copy()
    torch.rand = rec_rand
    torch.rand_like = lambda t, **k: noise_u
    try:
        with torch.no_grad():
            LeggedRobot.post_physics_step(env)
    finally:
        torch.rand, torch.rand_like = orig_rand, orig_rand_like
    out["rand/step"] = U["step"].numpy().copy(); out["rand/reset"] = U["reset"].numpy().copy()
    out["meta/num_train_envs"] = np.int64(NT); out["meta/push_interval"] = np.array([Cfg.domain_rand.push_interval, ECfg.domain_rand.push_interval])
    out["meta/rand_interval"] = np.int64(Cfg.domain_rand.rand_interval)
    out["meta/interval_resample_ids"] = resample_calls[0].numpy().copy()
    snap("out", ["root_states", "dof_pos", "dof_vel", "last_actions", "last_last_actions", "last_dof_vel", "last_joint_pos_target",
                 "last_last_joint_pos_target", "lag_buffer", "motor_offsets", "motor_strengths", "Kp_factors", "Kd_factors", "friction_coeffs",
                 "restitutions", "payloads", "com_displacements", "commands", "gait_indices", "episode_length_buf", "base_lin_vel",
                 "base_ang_vel", "projected_gravity", "clock_inputs", "desired_contact_states", "foot_indices", "rew_buf", "rew_buf_pos",
                 "rew_buf_neg", "last_contacts", "last_root_vel", "feet_air_time"])
    out["out/reset_buf"] = env.reset_buf.numpy().astype(bool); out["out/time_out_buf"] = env.time_out_buf.numpy().copy()
    out["out/obs_buf"] = torch.clip(env.obs_buf, -100, 100).numpy().copy()
    out["out/privileged_obs_buf"] = torch.clip(env.privileged_obs_buf, -100, 100).numpy().copy()
    for k, v in env.episode_sums.items():
        out[f"out/episode_sums/{k}"] = v.numpy().copy()
        out[f"out/episode_sums_eval/{k}"] = env.episode_sums_eval[k].numpy().copy()
    for k, v in env.command_sums.items():
        out[f"out/command_sums/{k}"] = v.numpy().copy()
    for k, v in env.extras["train/episode"].items():
        out[f"out/extras_train_episode/{k}"] = np.asarray(v, dtype=np.float64)
    out["out/extras_time_outs"] = env.extras["time_outs"].numpy().copy()
    out["out/extras_has_eval_episode"] = np.bool_("eval/episode" in env.extras)
    # the resolved config trees of this case (the GPU test rebuilds the same Cfg pair from them)
    with open(os.path.join(HERE, "env_dr_cfg.json"), "w") as f:
        json.dump({"overrides": DR_OVERRIDES, "eval_overrides": EVAL_OVERRIDES,
                   "x_offset": [Cfg.terrain.x_offset, ECfg.terrain.x_offset]}, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "env_dr.npz"), **out)
    n_tel = int((out["out/root_states"][:, :2] != out["in/root_states"][:, :2]).any(1).sum())
    print("env_dr.npz:", len(out), "arrays; resets", out["reset/ids"].tolist(), "moved xy", n_tel,
          "pushed", int((out["rand/step"][:, 36] != 0.5).sum()), "re-randomised", int((out["rand/step"][:, 21] != 0.5).sum()))


def make_kats(Cfg):
    from go1_gym.envs.base.legged_robot import LeggedRobot
    from go1_gym.envs.base.curriculum import RewardThresholdCurriculum
    out = {}
    net = torch.jit.load(f"{REF}/resources/actuator_nets/unitree_go1.pt", map_location="cpu")
    torch.manual_seed(0)
    x = torch.randn(3, 6)
    out["actuator/x"] = x.numpy(); out["actuator/y"] = net(x).detach().numpy(); out["actuator/y0"] = net(torch.zeros(1, 6)).detach().numpy()
    # hardware log replay: MAE of the net vs measured tau_est
    class U(pickle.Unpickler):
        def find_class(self, mod, name):
            if mod == "torch.storage" and name == "_load_from_bytes":
                return lambda b: torch.load(io.BytesIO(b), map_location="cpu", weights_only=False)
            return super().find_class(mod, name)
    # gait clock KAT: f = 3 Hz, phase .5, 3 steps
    env = types.SimpleNamespace(cfg=Cfg, dt=0.02)
    env.commands = torch.zeros(1, 15); env.commands[0, 4] = 3.0; env.commands[0, 5] = 0.5; env.commands[0, 8] = 0.5
    env.gait_indices = torch.zeros(1)
    for n in ("clock_inputs", "doubletime_clock_inputs", "halftime_clock_inputs", "desired_contact_states"):
        setattr(env, n, torch.zeros(1, 4))
    for _ in range(3):
        LeggedRobot._step_contact_targets(env)
    out["gait/gait_indices"] = env.gait_indices.numpy().copy(); out["gait/clock_inputs"] = env.clock_inputs.numpy().copy()
    out["gait/desired_contact_states"] = env.desired_contact_states.numpy().copy()
    # curriculum KAT
    c = RewardThresholdCurriculum(100, x_vel=(-5, 5, 21), y_vel=(-.6, .6, 1), yaw_vel=(-5, 5, 21))
    c.set_to(np.array([-1, -.6, -1]), np.array([1, .6, 1]))
    cmds, bins = c.sample(5)
    out["curriculum/sample_cmds"] = cmds; out["curriculum/sample_bins"] = bins
    c.update(bins, [torch.tensor([1.0, 0.1, 1.0, 1.0, 0.2]), torch.tensor([1.0, 1.0, 1.0, 0.0, 1.0])], [0.5, 0.5],
             local_range=np.array([0.55, 0.55, 0.55]))
    out["curriculum/weights_after_update"] = c.weights.copy()
    cmds2, bins2 = c.sample(7)
    out["curriculum/sample2_cmds"] = cmds2; out["curriculum/sample2_bins"] = bins2
    # full 15-D train.py curriculum: a few samples/updates
    lim = Cfg.commands
    kw = dict(x_vel=(*lim.limit_vel_x, lim.num_bins_vel_x), y_vel=(*lim.limit_vel_y, lim.num_bins_vel_y), yaw_vel=(*lim.limit_vel_yaw, lim.num_bins_vel_yaw),
              body_height=(*lim.limit_body_height, lim.num_bins_body_height), gait_frequency=(*lim.limit_gait_frequency, lim.num_bins_gait_frequency),
              gait_phase=(*lim.limit_gait_phase, lim.num_bins_gait_phase), gait_offset=(*lim.limit_gait_offset, lim.num_bins_gait_offset),
              gait_bounds=(*lim.limit_gait_bound, lim.num_bins_gait_bound), gait_duration=(*lim.limit_gait_duration, lim.num_bins_gait_duration),
              footswing_height=(*lim.limit_footswing_height, lim.num_bins_footswing_height), body_pitch=(*lim.limit_body_pitch, lim.num_bins_body_pitch),
              body_roll=(*lim.limit_body_roll, lim.num_bins_body_roll), stance_width=(*lim.limit_stance_width, lim.num_bins_stance_width),
              stance_length=(*lim.limit_stance_length, lim.num_bins_stance_length), aux_reward_coef=(*lim.limit_aux_reward_coef, lim.num_bins_aux_reward_coef))
    c15 = RewardThresholdCurriculum(100, **kw)
    low = np.array([lim.lin_vel_x[0], lim.lin_vel_y[0], lim.ang_vel_yaw[0], lim.body_height_cmd[0], lim.gait_frequency_cmd_range[0], lim.gait_phase_cmd_range[0],
                    lim.gait_offset_cmd_range[0], lim.gait_bound_cmd_range[0], lim.gait_duration_cmd_range[0], lim.footswing_height_range[0],
                    lim.body_pitch_range[0], lim.body_roll_range[0], lim.stance_width_range[0], lim.stance_length_range[0], lim.aux_reward_coef_range[0]])
    high = np.array([lim.lin_vel_x[1], lim.lin_vel_y[1], lim.ang_vel_yaw[1], lim.body_height_cmd[1], lim.gait_frequency_cmd_range[1], lim.gait_phase_cmd_range[1],
                     lim.gait_offset_cmd_range[1], lim.gait_bound_cmd_range[1], lim.gait_duration_cmd_range[1], lim.footswing_height_range[1],
                     lim.body_pitch_range[1], lim.body_roll_range[1], lim.stance_width_range[1], lim.stance_length_range[1], lim.aux_reward_coef_range[1]])
    c15.set_to(low=low, high=high)
    out["curriculum15/weights0"] = c15.weights.copy()
    cm, bi = c15.sample(batch_size=6)
    out["curriculum15/cmds"] = cm; out["curriculum15/bins"] = bi
    lr = np.array([0.55, 0.55, 0.55, 0.55, 0.35, 0.25, 0.25, 0.25, 0.25, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0])
    c15.update(bi, [torch.tensor([1., 0, 1, 1, 0, 1]), torch.tensor([1., 1, 1, 1, 1, 0]), torch.tensor([1., 1, 1, 1, 1, 1]), torch.tensor([1., 1, 0, 1, 1, 1])],
               [0.5, 0.5, 0.5, 0.5], local_range=lr)
    out["curriculum15/weights1"] = c15.weights.copy()
    cm, bi = c15.sample(batch_size=9)
    out["curriculum15/cmds2"] = cm; out["curriculum15/bins2"] = bi
    # pretrained policy KAT
    run = f"{REF}/runs/gait-conditioned-agility/pretrain-v0/train/025417.456545/checkpoints"
    body = torch.jit.load(f"{run}/body_latest.jit", map_location="cpu")
    adapt = torch.jit.load(f"{run}/adaptation_module_latest.jit", map_location="cpu")
    h = torch.zeros(1, 2100)
    lat = adapt(h)
    out["policy/latent0"] = lat.detach().numpy(); out["policy/action0"] = body(torch.cat((h, lat), -1)).detach().numpy()
    np.savez_compressed(os.path.join(HERE, "kats.npz"), **out)
    print("kats.npz:", len(out), "arrays")


def make_ppo():
    """BASELINE.json config 1: ppo_cse GAE + ActorCritic + Adam update on a synthetic rollout (num_envs=4, T=24)."""
    from go1_gym_learn.ppo_cse.actor_critic import ActorCritic
    from go1_gym_learn.ppo_cse.ppo import PPO, PPO_Args
    N, T, NOBS, NH, NP, NA = 4, 24, 70, 2100, 2, 12
    torch.manual_seed(0)
    ac = ActorCritic(NOBS, NP, NH, NA)
    # deterministic, platform-independent initial weights (numpy PCG64), so the 12 MB need not be stored:
    # tests re-create them with the same helper (tests/ppo_golden_util.py)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ppo_golden_util import seeded_weights, sample_tensor
    with torch.no_grad():
        for k, v in seeded_weights({k: tuple(v.shape) for k, v in ac.state_dict().items()}).items():
            ac.state_dict()[k].copy_(torch.from_numpy(v))
    out = {}
    alg = PPO(ac, device="cpu")
    alg.init_storage(N, T, [NOBS], [NP], [NH], [NA])
    g = torch.Generator().manual_seed(7)
    obs = torch.randn(N, NOBS, generator=g); hist = torch.randn(N, NH, generator=g) * 0.3; priv = torch.randn(N, NP, generator=g)
    eps_all, rew_all, done_all = [], [], []
    hist_all, priv_all, obs_all = [], [], []
    import torch.distributions.normal as tn
    for t in range(T):
        hist_all.append(hist.numpy().copy()); priv_all.append(priv.numpy().copy()); obs_all.append(obs.numpy().copy())
        eps = torch.randn(N, NA, generator=g)
        eps_all.append(eps.numpy().copy())
        orig = tn.Normal.sample
        tn.Normal.sample = lambda self, sample_shape=torch.Size(): (self.mean + self.stddev * eps).detach()
        try:
            with torch.inference_mode():
                alg.act(obs, priv, hist)
        finally:
            tn.Normal.sample = orig
        rew = torch.randn(N, generator=g); done = (torch.rand(N, generator=g) < 0.1)
        rew_all.append(rew.numpy().copy()); done_all.append(done.numpy().copy())
        infos = {"env_bins": torch.zeros(N), "time_outs": torch.zeros(N, dtype=torch.bool)}
        with torch.inference_mode():
            alg.process_env_step(rew, done, infos)
        obs = torch.randn(N, NOBS, generator=g); hist = torch.randn(N, NH, generator=g) * 0.3; priv = torch.randn(N, NP, generator=g)
    out["last/hist"] = hist.numpy().copy(); out["last/priv"] = priv.numpy().copy()
    with torch.inference_mode():
        alg.compute_returns(hist, priv)
    st = alg.storage
    for n in ("actions", "values", "actions_log_prob", "mu", "sigma", "returns", "advantages", "rewards"):
        out[f"storage/{n}"] = getattr(st, n).detach().numpy().copy()
    out["storage/dones"] = st.dones.numpy().copy()
    out["in/eps"] = np.stack(eps_all); out["in/rew"] = np.stack(rew_all); out["in/done"] = np.stack(done_all)
    out["in/hist"] = np.stack(hist_all); out["in/priv"] = np.stack(priv_all); out["in/obs"] = np.stack(obs_all)
    # fix the minibatch permutation so it can be injected on the other side
    perm = torch.randperm(N * T, generator=g)
    out["in/perm"] = perm.numpy().copy()
    orig_rp = torch.randperm
    torch.randperm = lambda n, **k: perm
    try:
        losses = alg.update()
    finally:
        torch.randperm = orig_rp
    out["update/losses"] = np.array(losses, dtype=np.float64)
    out["update/learning_rate"] = np.array(alg.learning_rate)
    for k, v in ac.state_dict().items():
        out[f"final/{k}"] = sample_tensor(v.detach().numpy())     # strided samples + sum + sum of squares
    np.savez_compressed(os.path.join(HERE, "ppo.npz"), **out)
    print("ppo.npz:", len(out), "arrays; losses", losses[:3], "lr", alg.learning_rate)


def make_resample(Cfg):
    """the host side of reset_idx / the 500-step resample, through the reference's own code."""
    import math
    from go1_gym.envs.base.legged_robot import LeggedRobot
    N, dt = 64, 4 * float(np.float32(0.005))
    Cfg.env.max_episode_length = math.ceil(Cfg.env.episode_length_s / dt)
    keys = ["tracking_lin_vel", "tracking_ang_vel", "tracking_contacts_shaped_force", "tracking_contacts_shaped_vel"]
    env = types.SimpleNamespace(cfg=Cfg, dt=dt, device="cpu")
    env.reward_scales = {k: float(getattr(Cfg.reward_scales, k)) * dt for k in keys}        # legged_robot.py:1396-1400
    env.curriculum_thresholds = {k: v for k, v in clean(vars(Cfg.curriculum_thresholds)).items()}
    LeggedRobot._init_command_distribution(env, torch.arange(N))
    env.commands = torch.zeros(N, Cfg.commands.num_commands)
    env.command_sums = {k: torch.zeros(N) for k in keys + ["lin_vel_raw", "ang_vel_raw"]}
    ep_len = min(Cfg.env.max_episode_length, int(Cfg.commands.resampling_time / dt))
    out = {"meta/dt": np.float64(dt), "meta/ep_len": np.int64(ep_len), "meta/max_episode_length": np.int64(Cfg.env.max_episode_length)}
    rs = np.random.RandomState(7)
    rounds = 14
    for r in range(rounds):
        k = [N, 1, 5, 9, 2, 17, 6, 3, 12, 1, 8, 30, 4, 11][r]
        ids = torch.from_numpy(np.sort(rs.choice(N, k, replace=False)).astype(np.int64))
        sums = np.zeros((k, 4), dtype=np.float32)
        for j, key in enumerate(keys):      # around the success threshold so that both outcomes occur (more successes later)
            thr = env.curriculum_thresholds[key] * env.reward_scales[key] * ep_len
            sums[:, j] = (thr * (1.0 + rs.uniform(-0.25, 0.6 + 0.1 * r, size=k))).astype(np.float32)
            env.command_sums[key][ids] = torch.from_numpy(sums[:, j])
        torch.manual_seed(1000 + r)
        floats = torch.rand(k)
        torch.manual_seed(1000 + r)
        LeggedRobot._resample_commands(env, ids)
        out[f"r{r}/ids"] = ids.numpy(); out[f"r{r}/sums"] = sums; out[f"r{r}/rand"] = floats.numpy()
        out[f"r{r}/commands"] = env.commands[ids].numpy().copy()
        out[f"r{r}/bins"] = env.env_command_bins.copy(); out[f"r{r}/categories"] = env.env_command_categories.copy()
        for i, c in enumerate(env.curricula):
            out[f"r{r}/weights{i}"] = np.packbits(np.round(c.weights * 5).astype(np.uint8) > 0) if False else c.weights.astype(np.float32)
        assert all((env.command_sums[key][ids] == 0).all() for key in keys)
    out["meta/rounds"] = np.int64(rounds)
    np.savez_compressed(os.path.join(HERE, "resample.npz"), **out)
    print("resample.npz:", len(out), "arrays; successes per round:", [int((out[f"r{r}/weights0"] > 0).sum()) for r in range(rounds)])


TERRAIN_CASES = {
    # name: (overrides of Cfg.terrain, numpy seed)
    "curriculum": (dict(mesh_type="trimesh", curriculum=True, selected=False, num_rows=4, num_cols=6, border_size=5, terrain_length=8., terrain_width=8.,
                        terrain_proportions=[0.1, 0.1, 0.35, 0.25, 0.2], terrain_noise_magnitude=0.1, difficulty_scale=1.0), 3),
    "randomised": (dict(mesh_type="heightfield", curriculum=False, selected=False, num_rows=3, num_cols=5, border_size=4, terrain_length=8.,
                        terrain_width=8., terrain_proportions=[0.1, 0.15, 0.15, 0.15, 0.15, 0.1, 0.0, 0.0, 0.1, 0.1], terrain_noise_magnitude=0.08,
                        difficulty_scale=1.0), 11),
    "train_py": (dict(num_rows=3, num_cols=3, border_size=2), 0),       # scripts/train.py's own terrain settings otherwise (flat tiles)
}


EVAL_TERRAIN_OVERRIDES = dict(curriculum=False, selected=False, num_rows=2, num_cols=3, border_size=3, terrain_noise_magnitude=0.05,
                              terrain_proportions=[0.2, 0.2, 0.2, 0.2, 0.2])


def make_terrain(Cfg):
    """The reference's Terrain class (go1_gym/utils/terrain.py) driving OUR generators: pins tile layout, type/difficulty
    selection, numpy RNG consumption and env origins of walk-these-ways_b200/go1_gym/utils/terrain.py."""
    from go1_gym.utils.terrain import Terrain
    out = {}
    for name, (over, seed) in TERRAIN_CASES.items():
        saved = {k: getattr(Cfg.terrain, k) for k in over}
        for k, v in over.items():
            setattr(Cfg.terrain, k, v)
        np.random.seed(seed)
        t = Terrain(Cfg.terrain, 16)
        out[f"{name}/height_field_raw"] = t.height_field_raw.copy()
        out[f"{name}/env_origins"] = Cfg.terrain.env_origins.copy()
        if Cfg.terrain.mesh_type == "trimesh":
            out[f"{name}/vertices_sample"] = t.vertices[::997].copy(); out[f"{name}/triangles_sample"] = t.triangles[::997].copy()
        for k, v in saved.items():
            setattr(Cfg.terrain, k, v)
    # train + eval tile sets in one map (terrain.py:37-51: eval rows are appended below the train rows)
    import copy
    over, seed = TERRAIN_CASES["curriculum"]
    saved = {k: getattr(Cfg.terrain, k) for k in over}
    for k, v in over.items():
        setattr(Cfg.terrain, k, v)
    ev = type("eval_terrain", (), {k: copy.deepcopy(v) for k, v in vars(Cfg.terrain).items() if not k.startswith("__")})
    for k, v in EVAL_TERRAIN_OVERRIDES.items():
        setattr(ev, k, v)
    np.random.seed(seed)
    t = Terrain(Cfg.terrain, 16, ev, 8)
    out["train_eval/height_field_raw"] = t.height_field_raw.copy()
    out["train_eval/env_origins"] = Cfg.terrain.env_origins.copy()
    out["train_eval/eval_env_origins"] = ev.env_origins.copy()
    out["train_eval/eval_offsets"] = np.array([ev.x_offset, ev.rows_offset, t.tot_rows, t.tot_cols])
    for k, v in saved.items():
        setattr(Cfg.terrain, k, v)
    # measured heights (legged_robot.py:1756-1806) on the curriculum map: the reference's own _init_height_points / _get_heights
    from go1_gym.envs.base.legged_robot import LeggedRobot
    over, seed = TERRAIN_CASES["curriculum"]
    saved = {k: getattr(Cfg.terrain, k) for k in over}
    for k, v in over.items():
        setattr(Cfg.terrain, k, v)
    hf = out["curriculum/height_field_raw"]
    n = 48
    g = torch.Generator().manual_seed(9)
    env = types.SimpleNamespace(device="cpu", cfg=Cfg)
    env.terrain = types.SimpleNamespace(cfg=Cfg.terrain)
    env.height_samples = torch.tensor(hf)
    rp = torch.rand(n, 3, generator=g) * torch.tensor([30.0, 46.0, 0.6]) + torch.tensor([1.0, 1.0, 0.1])
    ang = (torch.rand(n, 3, generator=g) - 0.5) * torch.tensor([0.8, 0.8, 6.2])
    cr, sr, cp, sp, cyw, syw = [f(ang[:, i] / 2) for i in range(3) for f in (torch.cos, torch.sin)]
    q = torch.stack([sr * cp * cyw - cr * sp * syw, cr * sp * cyw + sr * cp * syw, cr * cp * syw - sr * sp * cyw, cr * cp * cyw + sr * sp * syw], 1)
    env.root_states = torch.cat([rp, q, torch.zeros(n, 6)], 1)
    env.base_quat = env.root_states[:, 3:7]
    ids = torch.arange(n)
    env.height_points = LeggedRobot._init_height_points(env, ids, Cfg)
    out["heights/base_pos"] = rp.numpy(); out["heights/base_quat"] = q.numpy()
    out["heights/measured"] = LeggedRobot._get_heights(env, ids, Cfg).numpy()
    for k, v in saved.items():
        setattr(Cfg.terrain, k, v)
    np.savez_compressed(os.path.join(HERE, "terrain.npz"), **out)
    print("terrain.npz:", {k: v.shape for k, v in out.items()}, "nonzero", {k: int(np.count_nonzero(v)) for k, v in out.items() if "height" in k})


def make_metrics_envelope():
    """the reward-term trajectories of the shipped training log (runs/.../metrics.pkl, one record per 10
    iterations, 4000 envs on Isaac Gym): the first 60 records of the terms the learning-level comparison looks at."""
    class U(pickle.Unpickler):
        def find_class(self, mod, name):
            if mod == "torch.storage" and name == "_load_from_bytes":
                return lambda b: torch.load(io.BytesIO(b), map_location="cpu", weights_only=False)
            return super().find_class(mod, name)
    rows = []
    with open(f"{REF}/runs/gait-conditioned-agility/pretrain-v0/train/025417.456545/metrics.pkl", "rb") as f:
        u = U(f)
        while len(rows) < 60:
            try:
                rows.append(u.load())
            except EOFError:
                break
    keys = ["train/episode/rew_total/mean", "train/episode/rew_tracking_lin_vel/mean", "train/episode/rew_tracking_ang_vel/mean",
            "train/episode/rew_tracking_contacts_shaped_force/mean", "train/episode/rew_tracking_contacts_shaped_vel/mean",
            "train/episode/rew_collision/mean", "train/episode/rew_action_rate/mean", "train/episode/rew_torques/mean",
            "train/episode/command_area_trot/mean", "adaptation_loss/mean", "mean_value_loss/mean", "mean_surrogate_loss/mean", "iterations", "timesteps"]
    out = {k: [float(r[k]) if k in r else None for r in rows] for k in keys}
    with open(os.path.join(HERE, "metrics_envelope.json"), "w") as f:
        json.dump(out, f)
    print("metrics_envelope.json:", len(rows), "records; rew_total", [round(x, 3) for x in out["train/episode/rew_total/mean"][:6]])


if __name__ == "__main__":
    Cfg, trees = reference_train_cfg()
    with open(os.path.join(HERE, "cfg_trees.json"), "w") as f:
        json.dump(trees, f, indent=0, sort_keys=False)
    make_env_logic(Cfg)
    make_dr_step(Cfg)
    make_kats(Cfg)
    make_ppo()
    make_resample(Cfg)
    make_terrain(Cfg)
    make_metrics_envelope()
