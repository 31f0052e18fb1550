"""world_size-2 gloo test (CPU) of the host half of the shared curriculum: env.reset() under the shared
curriculum is the collective LeggedRobot._resample_commands_host_all_ranks -- every rank replays the single-process call for
all world * N envs and keeps its slice -- so both ranks must end with the curricula, RandomState streams and category stream of
one process that owns all envs, and with that process's commands for their own envs."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _host_env(n):
    sys.path[:0] = [os.path.join(ROOT, "walk-these-ways_b200"), os.path.join(ROOT, "walk-these-ways_b200", "compat")]
    for m in [k for k in sys.modules if k.startswith("go1_gym.envs.base.legged_robot_config")]:
        del sys.modules[m]
    from go1_gym.envs.base.legged_robot_config import Cfg
    from go1_b200.train_config import apply_train_config
    from go1_b200.config import build_sim_config
    from go1_gym.envs.base import legged_robot as LR
    apply_train_config(Cfg)
    Cfg.env.num_envs = n
    _, info = build_sim_config(Cfg, num_envs=n, num_train_envs=n, seed=0)
    env = object.__new__(LR.LeggedRobot)
    env.cfg, env.dt, env.num_envs, env.device = Cfg, info["dt"], n, "cpu"
    env.reward_scales = dict(info["active_reward_scales"])
    env.curriculum_thresholds = LR.cfg_dict(Cfg.curriculum_thresholds)
    env.shared_curriculum = True
    env._init_command_distribution(np.arange(n))
    return env


def _state(env):
    return ([c.weights.copy() for c in env.curricula], [c.rng.get_state()[1].copy() for c in env.curricula], env._cat_rng.state,
            env.env_command_bins.copy(), env.env_command_categories.copy())


def _sums(n, seed):
    rs = np.random.RandomState(seed)
    return (rs.uniform(0.0, 12.0, size=(n, 4))).astype(np.float32)


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    env = _host_env(n)
    sums = _sums(world * n, 3)[rank * n:(rank + 1) * n]
    cmds = env._resample_commands_host_all_ranks(sums)
    cmds2 = env._resample_commands_host_all_ranks(sums * 0.5)          # a second collective reset continues the shared streams
    out[rank] = (cmds, cmds2, _state(env))
    dist.destroy_process_group()


def test_collective_host_reset_matches_single_process():
    world, n = 2, 48
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    one = _host_env(world * n)
    one.shared_curriculum = False
    sums = _sums(world * n, 3)
    want = one._resample_commands_host(np.arange(world * n), sums)
    want2 = one._resample_commands_host(np.arange(world * n), sums * 0.5)
    W, K, cat, bins, cats = _state(one)
    for r in range(world):
        cmds, cmds2, (w, k, c, b, ct) = out[r]
        sl = slice(r * n, (r + 1) * n)
        assert np.array_equal(cmds, want[sl]) and np.array_equal(cmds2, want2[sl])
        assert all(np.array_equal(a, bb) for a, bb in zip(w, W)) and all(np.array_equal(a, bb) for a, bb in zip(k, K)) and c == cat
        assert np.array_equal(b, bins[sl]) and np.array_equal(ct, cats[sl])
