"""Loads tests/golden/env_logic.npz into the oracle's state dict / the kernel's buffers."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def load_gold():
    return np.load(os.path.join(HERE, "golden", "env_logic.npz"))


def train_sim_config(num_envs, **kw):
    import sys
    for m in [k for k in sys.modules if k.startswith("go1_gym.envs.base.legged_robot_config")]:
        del sys.modules[m]
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "walk-these-ways_b200", "compat"))
    from go1_gym.envs.base.legged_robot_config import Cfg
    from go1_b200.train_config import apply_train_config
    from go1_b200.config import build_sim_config
    apply_train_config(Cfg)
    for sec, attrs in kw.pop("cfg_overrides", {}).items():
        for k, v in attrs.items():
            setattr(getattr(Cfg, sec), k, v)
    c, info = build_sim_config(Cfg, num_envs=num_envs, **kw)
    return Cfg, c, info


def oracle_state(g):
    """State dict (torch CPU) for oracle/env_oracle.py from the golden inputs."""
    t = lambda k: torch.from_numpy(np.array(g[k]))
    s = {k[3:]: t(k) for k in g.files if k.startswith("in/")}
    s["lag_buffer"] = [x.clone() for x in s["lag_buffer"]]
    s["motor_strengths"] = s["motor_strengths"]
    s["friction_coeffs"] = s["friction_coeffs"][:, 0]
    s["restitutions"] = s["restitutions"][:, 0]
    s["Kp_factors"] = torch.ones_like(s["dof_pos"])
    s["Kd_factors"] = torch.ones_like(s["dof_pos"])
    return s


def zero_sums(P, n, extra=("lin_vel_raw", "ang_vel_raw", "lin_vel_residual", "ang_vel_residual", "ep_timesteps")):
    ep = {k: torch.zeros(n) for k in list(P["reward_scales"]) + ["total"]}
    cs = {k: torch.zeros(n) for k in list(P["reward_scales"]) + list(extra)}
    return ep, cs
