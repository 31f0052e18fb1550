"""VelocityTrackingEasyEnv (reference go1_gym/envs/go1/velocity_tracking/__init__.py:10-49).  Same
constructor and 4-tuple step(); the 13 per-step device->host copies of the reference's `extras`
(joint_pos, body_linear_vel, ...) are produced lazily, only when a consumer reads them."""
import types

import torch

from go1_gym.envs.base.legged_robot import LeggedRobot
from go1_gym.envs.base.legged_robot_config import Cfg


class VelocityTrackingEasyEnv(LeggedRobot):
    def __init__(self, sim_device, headless, num_envs=None, prone=False, deploy=False, cfg: Cfg = None, eval_cfg: Cfg = None,
                 initial_dynamics_dict=None, physics_engine="SIM_PHYSX"):
        if num_envs is not None:
            cfg.env.num_envs = num_envs
        sim_params = types.SimpleNamespace(**{k: v for k, v in vars(cfg.sim).items() if not k.startswith("_")})
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless, eval_cfg, initial_dynamics_dict)

    def _register_lazy_extras(self):
        """The 12 per-step numpy extras of the reference (velocity_tracking/__init__.py:27-42): registered once; each
        one reads the live device state when (and only when) a consumer looks it up."""
        np_ = lambda t: t.detach().cpu().numpy()
        c, extras = self, self.extras
        extras.lazy("joint_pos", lambda: np_(c.dof_pos))
        extras.lazy("joint_vel", lambda: np_(c.dof_vel))
        extras.lazy("joint_pos_target", lambda: np_(c.joint_pos_target))
        extras.lazy("joint_vel_target", lambda: torch.zeros(12))
        extras.lazy("body_linear_vel", lambda: np_(c.base_lin_vel))
        extras.lazy("body_angular_vel", lambda: np_(c.base_ang_vel))
        extras.lazy("body_linear_vel_cmd", lambda: np_(c.commands)[:, 0:2])
        extras.lazy("body_angular_vel_cmd", lambda: np_(c.commands)[:, 2:])
        extras.lazy("contact_states", lambda: np_(c.contact_forces[:, c.feet_indices, 2] > 1.).copy())
        extras.lazy("foot_positions", lambda: np_(c.foot_positions).copy())
        extras.lazy("body_pos", lambda: np_(c.root_states[:, 0:3]))
        extras.lazy("torques", lambda: np_(c.torques))
        self._lazy_extras_of = id(extras)

    def step(self, actions):
        obs, priv, rew, reset, extras = super().step(actions)
        extras["privileged_obs"] = priv
        if self.__dict__.get("_lazy_extras_of") != id(extras):
            self._register_lazy_extras()
        return obs, rew, reset, extras

    def reset(self):
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        obs, _, _, _ = self.step(torch.zeros(self.num_envs, self.num_actions, device=self.device, requires_grad=False))
        return obs
