"""Pins oracle/env_oracle.py (the travelling restatement) to vectors produced by the reference's own
LeggedRobot / CoRLRewards code (tests/golden/env_logic.npz)."""
import numpy as np
import torch

from oracle import env_oracle as eo
from env_golden_util import load_gold, oracle_state, train_sim_config, zero_sums

TOL = dict(rtol=1e-5, atol=1e-6)


def _setup():
    g = load_gold()
    Cfg, c, info = train_sim_config(64)
    P = eo.params_from_sim_config(c, info["active_reward_scales"], info["dt"])
    return g, P, oracle_state(g)


def test_compute_torques_two_substeps():
    g, P, s = _setup()
    net = eo.ActuatorNet()
    for sub in range(2):
        tq = eo.compute_torques(s, P, net)
        assert np.allclose(tq.numpy(), g[f"torques/sub{sub}"], rtol=1e-5, atol=2e-5)
        assert np.allclose(s["joint_pos_target"].numpy(), g[f"joint_pos_target/sub{sub}"], **TOL)
    assert np.allclose(torch.stack(s["lag_buffer"]).numpy(), g["after_torques/lag_buffer"], **TOL)
    for k in ("joint_pos_err_last", "joint_pos_err_last_last", "joint_vel_last", "joint_vel_last_last"):
        assert np.allclose(s[k].numpy(), g[f"after_torques/{k}"], **TOL)


def _post(g, P, s):
    q = s["root_states"][:, 3:7]
    s["base_lin_vel"] = eo.quat_rotate_inverse(q, s["root_states"][:, 7:10])
    s["base_ang_vel"] = eo.quat_rotate_inverse(q, s["root_states"][:, 10:13])
    s["projected_gravity"] = eo.quat_rotate_inverse(q, s["gravity_vec"])
    eo.step_contact_targets(s, P)


def test_base_frame_and_gait_clock():
    g, P, s = _setup()
    _post(g, P, s)
    for k in ("base_lin_vel", "base_ang_vel", "projected_gravity"):
        assert np.allclose(s[k].numpy(), g[f"post/{k}"], **TOL)
    for k in ("gait_indices", "foot_indices", "clock_inputs", "doubletime_clock_inputs", "halftime_clock_inputs", "desired_contact_states"):
        assert np.allclose(s[k].numpy(), g[f"gait/{k}"], rtol=1e-5, atol=2e-6), k


def test_termination_masks_bit_exact():
    g, P, s = _setup()
    reset, time_out = eo.check_termination(s, P)
    assert np.array_equal(reset.numpy(), g["term/reset_buf"]) and np.array_equal(time_out.numpy(), g["term/time_out_buf"])
    assert g["term/time_out_buf"][:3].tolist() == [False, True, False]          # ep_len 1001, 1002, 1000 vs max 1001


def test_reward_terms_and_totals():
    g, P, s = _setup()
    net = eo.ActuatorNet()
    for _ in range(2):
        s["torques"] = eo.compute_torques(s, P, net)
    _post(g, P, s)
    R = eo.reward_terms(s, P)
    for name in g["reward/names"]:
        assert np.allclose(R[str(name)].numpy(), g[f"reward_raw/{name}"], rtol=2e-5, atol=1e-6), name
    s["episode_sums"], s["command_sums"] = zero_sums(P, 64)
    rew, pos, neg = eo.compute_reward(s, P)
    assert np.allclose(rew.numpy(), g["reward/rew_buf"], rtol=2e-5, atol=1e-7)
    assert np.allclose(pos.numpy(), g["reward/rew_buf_pos"], rtol=2e-5, atol=1e-7)
    assert np.allclose(neg.numpy(), g["reward/rew_buf_neg"], rtol=2e-5, atol=1e-7)
    assert np.array_equal(s["last_contacts"].numpy(), g["reward/last_contacts"])
    for k, v in s["episode_sums"].items():
        assert np.allclose(v.numpy(), g[f"episode_sums/{k}"], rtol=2e-5, atol=1e-7), k
    for k, v in s["command_sums"].items():
        assert np.allclose(v.numpy(), g[f"command_sums/{k}"], rtol=2e-5, atol=1e-7), k


def test_observations():
    g, P, s = _setup()
    net = eo.ActuatorNet()
    for _ in range(2):
        eo.compute_torques(s, P, net)
    _post(g, P, s)
    obs, priv = eo.compute_observations(s, P, torch.from_numpy(g["obs/noise_u"]))
    assert obs.shape == (64, 70) and priv.shape == (64, 2)
    assert np.allclose(obs.numpy(), g["obs/obs_buf"], rtol=1e-5, atol=2e-6)
    assert np.allclose(priv.numpy(), g["obs/privileged_obs_buf"], rtol=1e-5, atol=1e-6)


def test_actuator_net_known_answers():
    k = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "kats.npz"))
    net = eo.ActuatorNet()
    assert np.allclose(net(torch.from_numpy(k["actuator/x"])).numpy(), k["actuator/y"], rtol=1e-5, atol=1e-5)
    assert np.allclose(k["actuator/y"].ravel(), [18.1223, -21.0130, -20.5186], atol=1e-3)
    assert abs(float(net(torch.zeros(1, 6))) - (-0.0041)) < 1e-4


def test_gait_clock_known_answer():
    k = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "kats.npz"))
    P = dict(dt=0.02, pacing_offset=False, kappa_gait_probs=0.07)
    s = dict(commands=torch.zeros(1, 15), gait_indices=torch.zeros(1))
    s["commands"][0, 4] = 3.0; s["commands"][0, 5] = 0.5; s["commands"][0, 8] = 0.5
    for _ in range(3):
        eo.step_contact_targets(s, P)
    assert np.allclose(s["gait_indices"].numpy(), k["gait/gait_indices"], atol=1e-6) and abs(float(s["gait_indices"]) - 0.18) < 1e-6
    assert np.allclose(s["clock_inputs"].numpy(), k["gait/clock_inputs"], atol=1e-5)
    assert np.allclose(s["desired_contact_states"].numpy(), k["gait/desired_contact_states"], atol=1e-5)
    assert np.allclose(k["gait/clock_inputs"].ravel(), [-.9048, .9048, .9048, -.9048], atol=1e-4)


def test_full_dr_post_physics_step_matches_reference():
    """The reference's whole post_physics_step with teleports, pushes, periodic re-randomisation, reset_idx (dof props, rigid
    props, dof / root reset, buffer clears, episode means, eval sums) and a train/eval split: tests/golden/env_dr.npz.
    Injected-draw arithmetic is bit-exact; transcendental paths (gait clock, rewards, yaw quaternion) to 1e-6."""
    from env_golden_util import dr_case, dr_oracle_state
    g, Cfg, ECfg, c, info = dr_case()
    P = eo.params_from_sim_config(c, info["active_reward_scales"], info["dt"])
    s = dr_oracle_state(g, P)
    T = lambda k: torch.from_numpy(np.array(g[k]))
    r = eo.post_physics_step(s, P, T("rand/step"), T("rand/reset"), T("obs/noise_u"), T("in/new_commands"))
    assert np.array_equal(r["reset_ids"].numpy(), g["reset/ids"])
    assert np.array_equal(r["reset"].numpy(), g["out/reset_buf"]) and np.array_equal(r["time_out"].numpy(), g["out/time_out_buf"])
    assert np.array_equal(s["episode_length_buf"].numpy(), g["out/episode_length_buf"])
    exact = ["dof_pos", "dof_vel", "last_actions", "last_last_actions", "last_dof_vel", "last_joint_pos_target", "last_last_joint_pos_target",
             "motor_offsets", "motor_strengths", "Kp_factors", "Kd_factors", "payloads", "com_displacements", "commands"]
    for k in exact:
        assert np.array_equal(s[k].numpy(), g[f"out/{k}"]), k
    assert np.array_equal(s["friction_coeffs"].numpy(), g["out/friction_coeffs"][:, 0]) and np.array_equal(s["restitutions"].numpy(), g["out/restitutions"][:, 0])
    assert np.array_equal(torch.stack(s["lag_buffer"]).numpy(), g["out/lag_buffer"])
    rs, want = s["root_states"].numpy(), g["out/root_states"]
    assert np.array_equal(rs[:, [0, 1, 2, 7, 8, 9, 10, 11, 12]], want[:, [0, 1, 2, 7, 8, 9, 10, 11, 12]])        # positions / twists: exact
    assert np.allclose(rs[:, 3:7], want[:, 3:7], rtol=0, atol=1e-6)                                              # yaw quaternion: sin / cos
    for k in ("base_lin_vel", "base_ang_vel", "projected_gravity", "gait_indices", "clock_inputs", "desired_contact_states", "foot_indices"):
        assert np.allclose(s[k].numpy(), g[f"out/{k}"], rtol=1e-5, atol=2e-6), k
    assert np.allclose(r["rew"].numpy(), g["out/rew_buf"], rtol=1e-5, atol=1e-6)
    assert np.allclose(r["rew_pos"].numpy(), g["out/rew_buf_pos"], rtol=1e-5, atol=1e-6) and np.allclose(r["rew_neg"].numpy(), g["out/rew_buf_neg"], rtol=1e-5, atol=1e-6)
    assert np.allclose(r["obs"].numpy(), g["out/obs_buf"], rtol=1e-5, atol=2e-6)
    assert np.allclose(r["priv"].numpy(), g["out/privileged_obs_buf"], rtol=1e-5, atol=2e-6)
    for k in s["episode_sums"]:
        assert np.allclose(s["episode_sums"][k].numpy(), g[f"out/episode_sums/{k}"], rtol=1e-5, atol=1e-6), k
        assert np.allclose(s["episode_sums_eval"][k].numpy(), g[f"out/episode_sums_eval/{k}"], rtol=1e-5, atol=1e-6), k
    for k in s["command_sums"]:
        assert np.allclose(s["command_sums"][k].numpy(), g[f"out/command_sums/{k}"], rtol=1e-5, atol=1e-6), k
    for k, v in r["episode_means"].items():
        assert np.allclose(float(v), g[f"out/extras_train_episode/{k}"], rtol=1e-5, atol=1e-6), k
    # the case exercises every branch
    moved = (g["out/root_states"][:, :2] != g["in/root_states"][:, :2]).any(1)
    moved[g["reset/ids"]] = False
    assert moved.sum() >= 4 and (g["rand/step"][:, 36] != 0.5).sum() >= 5 and (g["rand/step"][:, 21] != 0.5).sum() >= 5
    assert (g["reset/ids"] >= int(g["meta/num_train_envs"])).sum() >= 3
