"""env_step_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT.

A complete CPU env step assembled from the pinned pieces: oracle/env_oracle.py (control, gait clock, termination,
rewards, observations: pinned to the reference's goldens) + oracle/go1_physics_oracle.c (fp64 rigid-body substep,
parity unpinned vs PhysX).  Used by bench.py (`cpu_baseline`, `--impl reference`) and by __graft_entry__.smoke().
Resets re-initialise terminated envs to the default pose with the last commands (the host curriculum is not part
of the timed CPU path)."""
import numpy as np
import torch

from . import env_oracle as eo
from . import physics as ph


class OracleEnv:
    def __init__(self, sim_cfg, active_scales, dt, n, seed=0):
        self.c, self.n = sim_cfg, n
        self.P = eo.params_from_sim_config(sim_cfg, active_scales, dt)
        self.net = eo.ActuatorNet()
        self.pp = ph.default_params()
        self.rng = np.random.default_rng(seed)
        z = lambda *s: torch.zeros(n, *s)
        self.s = dict(dof_pos=torch.tensor(ph.DEFAULT_DOF_POS, dtype=torch.float32).repeat(n, 1), dof_vel=z(12),
                      lag_buffer=[z(12) for _ in range(7)], motor_offsets=z(12), motor_strengths=torch.ones(n, 12),
                      Kp_factors=torch.ones(n, 12), Kd_factors=torch.ones(n, 12), joint_pos_err_last=z(12), joint_pos_err_last_last=z(12),
                      joint_vel_last=z(12), joint_vel_last_last=z(12), last_actions=z(12), last_last_actions=z(12), last_dof_vel=z(12),
                      last_joint_pos_target=z(12), last_last_joint_pos_target=z(12), gait_indices=z(), commands=z(15),
                      friction_coeffs=torch.ones(n), restitutions=z(), last_contacts=torch.zeros(n, 4, dtype=torch.bool),
                      episode_length_buf=torch.zeros(n, dtype=torch.long), prev_foot_velocities=z(4, 3),
                      gravity_vec=torch.tensor([0., 0., -1.]).repeat(n, 1))
        self.s["commands"][:, 4] = 3.0; self.s["commands"][:, 5] = 0.5; self.s["commands"][:, 8] = 0.5
        self.s["commands"][:, 9] = 0.08; self.s["commands"][:, 12] = 0.25; self.s["commands"][:, 13] = 0.4
        self.states = ph.state_array(n)
        self.drs = ph.dr_array(n)
        for i in range(n):
            self.states[i] = ph.make_state([0, 0, 0.34], [0, 0, 0, 1], [0, 0, 0], [0, 0, 0], ph.DEFAULT_DOF_POS, np.zeros(12))
            self.drs[i] = ph.make_dr(1.0, 0.0, 0.0)
        self._sview = np.frombuffer(self.states, dtype=np.float64).reshape(n, 37)      # pos3 quat4 lin3 ang3 q12 qd12
        self.s["episode_sums"], self.s["command_sums"] = None, None

    def _sync_from_physics(self):
        self.s["dof_pos"] = torch.tensor(self._sview[:, 13:25], dtype=torch.float32); self.s["dof_vel"] = torch.tensor(self._sview[:, 25:37], dtype=torch.float32)

    def step(self, actions):
        s, P, n = self.s, self.P, self.n
        s["actions"] = torch.clip(actions, -P["clip_actions"], P["clip_actions"])
        cf = None
        for _ in range(self.c.decimation):
            tau = eo.compute_torques(s, P, self.net)
            tq = tau.numpy().astype(np.float64)
            cf = ph.substep_batch(self.pp, self.drs, self.states, tq)
            self._sync_from_physics()
        s["torques"] = tau
        s["root_states"] = torch.tensor(self._sview[:, :13], dtype=torch.float32)
        fp, fv = ph.feet_batch(self.states)
        s["foot_positions"] = torch.tensor(fp, dtype=torch.float32)
        s["foot_velocities"] = torch.tensor(fv, dtype=torch.float32)
        s["contact_forces"] = torch.tensor(cf, dtype=torch.float32)
        s["episode_length_buf"] = s["episode_length_buf"] + 1
        quat = s["root_states"][:, 3:7]
        s["base_lin_vel"] = eo.quat_rotate_inverse(quat, s["root_states"][:, 7:10])
        s["base_ang_vel"] = eo.quat_rotate_inverse(quat, s["root_states"][:, 10:13])
        s["projected_gravity"] = eo.quat_rotate_inverse(quat, s["gravity_vec"])
        eo.step_contact_targets(s, P)
        reset, time_out = eo.check_termination(s, P)
        if s["episode_sums"] is None:
            keys = list(P["reward_scales"])
            s["episode_sums"] = {k: torch.zeros(n) for k in keys + ["total"]}
            s["command_sums"] = {k: torch.zeros(n) for k in keys + ["lin_vel_raw", "ang_vel_raw", "lin_vel_residual", "ang_vel_residual", "ep_timesteps"]}
        rew, _, _ = eo.compute_reward(s, P)
        for i in torch.nonzero(reset).flatten().tolist():
            self.states[i] = ph.make_state([0, 0, 0.34], [0, 0, 0, 1], [0, 0, 0], [0, 0, 0], ph.DEFAULT_DOF_POS * self.rng.uniform(0.5, 1.5, 12), np.zeros(12))
            s["episode_length_buf"][i] = 0; s["last_actions"][i] = 0; s["last_last_actions"][i] = 0; s["last_dof_vel"][i] = 0; s["gait_indices"][i] = 0
            for b in s["lag_buffer"]:
                b[i] = 0
        if reset.any():
            self._sync_from_physics()
        u = torch.from_numpy(self.rng.random((n, self.c.num_obs)).astype(np.float32))
        obs, priv = eo.compute_observations(s, P, u)
        s["last_last_actions"] = s["last_actions"].clone(); s["last_actions"] = s["actions"].clone()
        s["last_last_joint_pos_target"] = s["last_joint_pos_target"].clone(); s["last_joint_pos_target"] = s["joint_pos_target"].clone()
        s["last_dof_vel"] = s["dof_vel"].clone(); s["prev_foot_velocities"] = s["foot_velocities"].clone()
        return obs, priv, rew, reset
