"""Cfg mirror == the reference's configuration tree (golden cfg_trees.json made by make_golden.py from the
reference's own legged_robot_config.py, go1_config.py and scripts/train.py), including key ORDER of
reward_scales (it fixes the reward summation order), and the derived intervals of the shipped run."""
import copy
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _fresh_cfg():
    for m in [k for k in sys.modules if k.startswith("go1_gym.envs.base.legged_robot_config")]:
        del sys.modules[m]
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "walk-these-ways_b200", "compat"))
    from go1_gym.envs.base.legged_robot_config import Cfg
    return Cfg


def _tree(C):
    clean = lambda d: {k: v for k, v in dict(d).items() if not k.startswith("_")}
    out = {}
    for s, v in clean(vars(C)).items():
        if isinstance(v, type):
            out[s] = {k: (clean(vars(x)) if isinstance(x, type) else x) for k, x in clean(vars(v)).items()}
    return json.loads(json.dumps(out))


def test_cfg_trees_match_reference():
    gold = json.load(open(os.path.join(HERE, "golden", "cfg_trees.json")))
    Cfg = _fresh_cfg()
    from go1_gym.envs.go1.go1_config import config_go1
    from go1_b200.train_config import apply_train_config
    assert _tree(Cfg) == gold["defaults"]
    c1 = copy.deepcopy(_tree(Cfg))
    config_go1(Cfg)
    assert _tree(Cfg) == gold["go1"] and c1 != gold["go1"]
    apply_train_config(Cfg)
    mine = _tree(Cfg)
    assert mine == gold["train"]
    assert list(mine["reward_scales"]) == list(gold["train"]["reward_scales"])      # order matters


def test_derived_intervals_match_shipped_parameters_pkl():
    """parameters.pkl of the shipped run: max_episode_length 1001, rand_interval 201, gravity 401/397."""
    Cfg = _fresh_cfg()
    from go1_b200.train_config import apply_train_config
    from go1_b200.config import build_sim_config
    apply_train_config(Cfg)
    c, info = build_sim_config(Cfg, num_envs=8)
    assert c.max_episode_length == 1001 and c.rand_interval == 201 and c.resampling_interval == 500
    assert Cfg.domain_rand.gravity_rand_interval == 401 and Cfg.domain_rand.gravity_rand_duration == 397
    assert c.num_obs == 70 and c.num_priv_obs == 2 and c.num_active_rewards == 19
    assert abs(info["dt"] - 4 * float(np.float32(0.005))) < 1e-15


def test_noise_scale_vec_matches_reference():
    gold = np.load(os.path.join(HERE, "golden", "env_logic.npz"))
    Cfg = _fresh_cfg()
    from go1_b200.train_config import apply_train_config
    from go1_b200.config import build_sim_config
    apply_train_config(Cfg)
    c, info = build_sim_config(Cfg, num_envs=8)
    assert np.allclose(np.array(c.noise_scale_vec)[:70], gold["in/noise_scale_vec"], atol=1e-8)
    lim = gold["in/dof_pos_limits"]
    assert np.allclose(np.array(c.soft_limit_lo), lim[:, 0], atol=1e-6) and np.allclose(np.array(c.soft_limit_hi), lim[:, 1], atol=1e-6)
    names = list(gold["reward/names"])
    from go1_b200 import capi
    assert [capi.REWARD_TERMS[i] for i in list(c.reward_order)[:c.num_active_rewards]] == names
    assert np.allclose([c.reward_scale[capi.REWARD_TERMS.index(n)] for n in names], gold["reward/scales"], rtol=1e-6)
