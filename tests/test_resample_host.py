"""Host side of reset_idx / the 500-step command resample (SURVEY.md §8 a9 + a13) against the reference's own
LeggedRobot._resample_commands (legged_robot.py:710-824), replayed from tests/golden/resample.npz: curriculum weight
updates, bin/category bookkeeping, RandomState(100) sampling stream, gait-category remap of commands 5-7, the
small-command zeroing.  Bit-exact (float32 commands, int bins, float weights).  CPU only: no CUDA call is made."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "walk-these-ways_b200"), os.path.join(ROOT, "walk-these-ways_b200", "compat")]
G = np.load(os.path.join(ROOT, "tests", "golden", "resample.npz"))


class _Replay:
    """Stands in for the category RNG: returns the torch.rand draws the reference made in that round."""
    def __init__(self):
        self.queue = []

    def random(self, k):
        out = self.queue.pop(0)
        assert len(out) == k
        return out.astype(np.float64)


def _host_env():
    from go1_gym.envs.base.legged_robot_config import Cfg
    from go1_b200.train_config import apply_train_config
    from go1_b200.config import build_sim_config
    from go1_gym.envs.base import legged_robot as LR
    apply_train_config(Cfg)
    Cfg.env.num_envs = 64
    _, info = build_sim_config(Cfg, num_envs=64, num_train_envs=64, seed=0)
    env = object.__new__(LR.LeggedRobot)
    env.cfg, env.dt = Cfg, info["dt"]
    env.reward_scales = dict(info["active_reward_scales"])
    env.curriculum_thresholds = LR.cfg_dict(Cfg.curriculum_thresholds)
    env._init_command_distribution(np.arange(64))
    env._cat_rng = _Replay()
    return env, Cfg


def test_resample_commands_matches_reference_rounds():
    env, Cfg = _host_env()
    assert abs(env.dt - float(G["meta/dt"])) < 1e-15 and Cfg.env.max_episode_length == int(G["meta/max_episode_length"])
    commands = np.zeros((64, 15), dtype=np.float32)
    for r in range(int(G["meta/rounds"])):
        ids, sums = G[f"r{r}/ids"], G[f"r{r}/sums"]
        env._cat_rng.queue.append(G[f"r{r}/rand"])
        new = env._resample_commands_host(ids, sums)
        assert new.dtype == np.float32 and new.shape == (len(ids), 15)
        commands[ids] = new
        assert np.array_equal(new, G[f"r{r}/commands"]), (r, np.abs(new - G[f"r{r}/commands"]).max())
        assert np.array_equal(env.env_command_bins, G[f"r{r}/bins"]), r
        assert np.array_equal(env.env_command_categories, G[f"r{r}/categories"]), r
        for i, cur in enumerate(env.curricula):
            assert np.array_equal(cur.weights.astype(np.float32), G[f"r{r}/weights{i}"]), (r, i)
    # the rounds exercised both outcomes of the success test and all four gait categories
    assert (G[f"r{int(G['meta/rounds']) - 1}/weights0"] > 0).sum() > (G["r0/weights0"] > 0).sum()
    assert set(np.unique(env.env_command_categories)) == {0, 1, 2, 3}
