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


def clone_cfg(C, name="CfgClone"):
    """Deep copy of a Cfg class tree (used for the eval_cfg of the train/eval split)."""
    import copy
    import types
    sections = {}
    for k, v in vars(C).items():
        if k.startswith("_") or isinstance(v, types.MappingProxyType):
            continue
        if isinstance(v, type):
            inner = {}
            for kk, vv in vars(v).items():
                if kk.startswith("_"):
                    continue
                inner[kk] = type(kk, (), {a: copy.deepcopy(b) for a, b in vars(vv).items() if not a.startswith("_")}) if isinstance(vv, type) \
                    else copy.deepcopy(vv)
            sections[k] = type(k, (), inner)
        else:
            sections[k] = copy.deepcopy(v)
    return type(name, (), sections)


def dr_case():
    """tests/golden/env_dr.npz (the reference's post_physics_step with full domain randomisation and a 40 + 24 train/eval
    split) together with the Cfg pair and the resolved kernel config of that case."""
    import json
    g = np.load(os.path.join(HERE, "golden", "env_dr.npz"))
    with open(os.path.join(HERE, "golden", "env_dr_cfg.json")) as f:
        meta = json.load(f)
    import sys
    for m in [k for k in sys.modules if k.startswith("go1_gym.envs.base.legged_robot_config")]:
        del sys.modules[m]
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "walk-these-ways_b200", "compat"))
    from go1_gym.envs.base.legged_robot_config import Cfg
    from go1_b200.train_config import apply_train_config
    from go1_b200.config import build_sim_config
    apply_train_config(Cfg)

    def over(C, table):
        for sec, kv in table.items():
            for k, v in kv.items():
                setattr(getattr(C, sec), k, v)
    over(Cfg, meta["overrides"])
    ECfg = clone_cfg(Cfg, "EvalCfg")
    over(ECfg, meta["eval_overrides"])
    Cfg.terrain.x_offset, ECfg.terrain.x_offset = meta["x_offset"]
    nt = int(g["meta/num_train_envs"])
    n = g["in/root_states"].shape[0]
    Cfg.env.num_envs, ECfg.env.num_envs = nt, n - nt
    c, info = build_sim_config(Cfg, num_envs=n, num_train_envs=nt, eval_cfg=ECfg)
    return g, Cfg, ECfg, c, info


def dr_oracle_state(g, P):
    t = lambda k: torch.from_numpy(np.array(g[k]))
    names = [k[3:] for k in g.files if k.startswith("in/") and "/" not in k[3:]]
    s = {k: t("in/" + k) for k in names}
    s["lag_buffer"] = [x.clone() for x in s["lag_buffer"]]
    s["friction_coeffs"], s["restitutions"] = s["friction_coeffs"][:, 0].clone(), s["restitutions"][:, 0].clone()
    s["episode_sums"] = {k[len("in/episode_sums/"):]: t(k) for k in g.files if k.startswith("in/episode_sums/")}
    s["command_sums"] = {k[len("in/command_sums/"):]: t(k) for k in g.files if k.startswith("in/command_sums/")}
    n = s["root_states"].shape[0]
    s["episode_sums_eval"] = {k: -torch.ones(n) for k in s["episode_sums"]}
    s["episode_sums_eval"]["total"] = torch.zeros(n)          # legged_robot.py:1423: "total" starts at 0, so it is never recorded
    nt = P["num_train_envs"]
    for k in s["episode_sums_eval"]:
        s["episode_sums_eval"][k][nt + 3] = 0.25
    return s
