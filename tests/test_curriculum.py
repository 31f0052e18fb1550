"""Host curriculum mirror vs the reference's RewardThresholdCurriculum (tests/golden/kats.npz, produced by the reference's
own class with RandomState(100)): bit-identical sample streams and weight updates."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "walk-these-ways_b200", "compat"))


def test_three_dim_curriculum_known_answer():
    from go1_gym.envs.base.curriculum import RewardThresholdCurriculum
    k = np.load(os.path.join(HERE, "golden", "kats.npz"))
    c = RewardThresholdCurriculum(100, x_vel=(-5, 5, 21), y_vel=(-.6, .6, 1), yaw_vel=(-5, 5, 21))
    c.set_to(np.array([-1, -.6, -1]), np.array([1, .6, 1]))
    cmds, bins = c.sample(5)
    assert np.array_equal(bins, k["curriculum/sample_bins"]) and bins.tolist() == [221, 198, 218, 261, 176]
    assert np.array_equal(cmds, k["curriculum/sample_cmds"])
    c.update(bins, [np.array([1.0, 0.1, 1.0, 1.0, 0.2], dtype=np.float32), np.array([1.0, 1.0, 1.0, 0.0, 1.0], dtype=np.float32)], [0.5, 0.5],
             local_range=np.array([0.55, 0.55, 0.55]))
    assert np.array_equal(c.weights, k["curriculum/weights_after_update"])
    cmds2, bins2 = c.sample(7)
    assert np.array_equal(bins2, k["curriculum/sample2_bins"]) and np.array_equal(cmds2, k["curriculum/sample2_cmds"])


def test_train_py_curriculum_streams():
    """The 15-D, 441-bin curriculum of scripts/train.py built through LeggedRobot._init_command_distribution's tables."""
    from env_golden_util import train_sim_config
    from go1_gym.envs.base.curriculum import RewardThresholdCurriculum
    k = np.load(os.path.join(HERE, "golden", "kats.npz"))
    Cfg, c, info = train_sim_config(4)
    lim = Cfg.commands
    dims = [("x_vel", "vel_x"), ("y_vel", "vel_y"), ("yaw_vel", "vel_yaw"), ("body_height", "body_height"), ("gait_frequency", "gait_frequency"),
            ("gait_phase", "gait_phase"), ("gait_offset", "gait_offset"), ("gait_bounds", "gait_bound"), ("gait_duration", "gait_duration"),
            ("footswing_height", "footswing_height"), ("body_pitch", "body_pitch"), ("body_roll", "body_roll"), ("stance_width", "stance_width"),
            ("stance_length", "stance_length"), ("aux_reward_coef", "aux_reward_coef")]
    cur = RewardThresholdCurriculum(seed=lim.curriculum_seed, **{n: (*getattr(lim, f"limit_{key}"), getattr(lim, f"num_bins_{key}")) for n, key in dims})
    keys = ["lin_vel_x", "lin_vel_y", "ang_vel_yaw", "body_height_cmd", "gait_frequency_cmd_range", "gait_phase_cmd_range", "gait_offset_cmd_range",
            "gait_bound_cmd_range", "gait_duration_cmd_range", "footswing_height_range", "body_pitch_range", "body_roll_range", "stance_width_range",
            "stance_length_range", "aux_reward_coef_range"]
    cur.set_to(low=np.array([getattr(lim, q)[0] for q in keys]), high=np.array([getattr(lim, q)[1] for q in keys]))
    assert len(cur) == 441 and np.array_equal(cur.weights, k["curriculum15/weights0"])
    cm, bi = cur.sample(batch_size=6)
    assert np.array_equal(bi, k["curriculum15/bins"]) and np.array_equal(cm, k["curriculum15/cmds"])
    lr = np.array([0.55, 0.55, 0.55, 0.55, 0.35, 0.25, 0.25, 0.25, 0.25, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0])
    f = lambda *a: np.array(a, dtype=np.float32)
    cur.update(bi, [f(1., 0, 1, 1, 0, 1), f(1., 1, 1, 1, 1, 0), f(1., 1, 1, 1, 1, 1), f(1., 1, 0, 1, 1, 1)], [0.5, 0.5, 0.5, 0.5], local_range=lr)
    assert np.array_equal(cur.weights, k["curriculum15/weights1"])
    cm, bi = cur.sample(batch_size=9)
    assert np.array_equal(bi, k["curriculum15/bins2"]) and np.array_equal(cm, k["curriculum15/cmds2"])
