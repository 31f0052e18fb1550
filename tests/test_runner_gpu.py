"""The scripts/train.py flow end to end through the drop-in packages: Cfg + config_go1-style overrides ->
VelocityTrackingEasyEnv -> HistoryWrapper -> Runner.learn() with logging, curriculum dumps and checkpoints, then the
artefacts scripts/play.py consumes (ac_weights_last.pt, adaptation_module_latest.jit, body_latest.jit) reproduce the
policy (play.py:24-45) and a resumed Runner starts from the saved weights and curriculum."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "walk-these-ways_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "walk-these-ways_b200", "compat"))


def _make(tmp_path, n=64):
    for m in [k for k in sys.modules if k.startswith("go1_gym.envs.base.legged_robot_config")]:
        del sys.modules[m]
    from go1_gym.envs.base.legged_robot_config import Cfg
    from go1_b200.train_config import apply_train_config
    from go1_gym.envs.go1.velocity_tracking import VelocityTrackingEasyEnv
    from go1_gym.envs.wrappers.history_wrapper import HistoryWrapper
    from go1_gym_learn.ppo_cse import Runner, RunnerArgs
    from ml_logger import logger
    apply_train_config(Cfg)
    Cfg.env.num_envs = n
    logger.configure(prefix="run", root=str(tmp_path))
    env = HistoryWrapper(VelocityTrackingEasyEnv(sim_device="cuda:0", headless=True, cfg=Cfg))
    return env, Runner, RunnerArgs, logger


def test_runner_learn_checkpoints_and_play_artifacts(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                      # Runner.save writes ./tmp/legged_data like the reference
    env, Runner, RunnerArgs, logger = _make(tmp_path)
    RunnerArgs.num_steps_per_env, RunnerArgs.save_interval, RunnerArgs.log_freq, RunnerArgs.save_video_interval = 8, 2, 1, 100
    RunnerArgs.resume = False
    runner = Runner(env, device="cuda:0")
    w0 = runner.alg.actor_critic.flat_params.clone()
    runner.learn(num_learning_iterations=3, init_at_random_ep_len=True, eval_freq=100)
    ac = runner.alg.actor_critic
    assert torch.isfinite(ac.flat_params).all() and not torch.equal(ac.flat_params, w0)
    run = os.path.join(str(tmp_path), "run")
    ck = os.path.join(run, "checkpoints")
    for f in ("ac_weights_last.pt", "ac_weights_000000.pt", "ac_weights_000002.pt", "adaptation_module_latest.jit", "body_latest.jit"):
        assert os.path.exists(os.path.join(ck, f)), (f, os.listdir(ck))
    assert os.path.exists(os.path.join(run, "curriculum", "distribution.pkl"))
    # play.py:24-45: policy = body(cat(obs_history, adaptation_module(obs_history)))
    sd = torch.load(os.path.join(ck, "ac_weights_last.pt"), map_location="cpu")
    assert set(sd) == set(ac.state_dict()) and all(torch.equal(sd[k].cpu(), v.cpu()) for k, v in ac.state_dict().items())
    body = torch.jit.load(os.path.join(ck, "body_latest.jit"))
    adapt = torch.jit.load(os.path.join(ck, "adaptation_module_latest.jit"))
    h = torch.randn(5, env.num_obs_history)
    want = ac.act_student(h.cuda()).cpu()
    lat = adapt(h)
    got = body(torch.cat((h, lat), dim=-1))
    assert torch.allclose(got, want, rtol=2e-2, atol=2e-2)          # TorchScript runs fp32 on the CPU, the kernels TF32
    # resume (ppo_cse/__init__.py:76-91): weights and curriculum distribution come back
    weights_before = [c.weights.copy() for c in env.curricula]
    env.env._curriculum_to_host()
    weights_before = [c.weights.copy() for c in env.curricula]
    RunnerArgs.resume, RunnerArgs.resume_path = True, run
    try:
        env2, Runner2, _, _ = _make(tmp_path)
        r2 = Runner2(env2, device="cuda:0")
        assert torch.equal(r2.alg.actor_critic.flat_params, ac.flat_params)
        for c, w in zip(env2.curricula, weights_before):
            assert c.weights.shape == w.shape
    finally:
        RunnerArgs.resume = False


def test_train_eval_split_rollout_and_update(tmp_path, monkeypatch):
    """SURVEY.md §8f row 3 (legged_robot.py:531-544, ppo_cse/__init__.py:142-146): eval envs are appended after the train envs,
    take their randomisation / reset ranges from eval_cfg, are driven by the student policy, never enter the rollout storage,
    and their first finished episode lands in episode_sums_eval."""
    monkeypatch.chdir(tmp_path)
    for m in [k for k in sys.modules if k.startswith("go1_gym.envs.base.legged_robot_config")]:
        del sys.modules[m]
    sys.path.insert(0, HERE)
    from env_golden_util import clone_cfg
    from go1_gym.envs.base.legged_robot_config import Cfg
    from go1_b200.train_config import apply_train_config
    from go1_gym.envs.go1.velocity_tracking import VelocityTrackingEasyEnv
    from go1_gym.envs.wrappers.history_wrapper import HistoryWrapper
    from go1_gym_learn.ppo_cse import Runner, RunnerArgs
    from ml_logger import logger
    apply_train_config(Cfg)
    Cfg.env.num_envs = 96
    ECfg = clone_cfg(Cfg, "EvalCfg")
    ECfg.env.num_envs = 32
    ECfg.domain_rand.friction_range = [2.9, 3.0]
    ECfg.domain_rand.added_mass_range = [2.5, 3.0]
    ECfg.domain_rand.motor_strength_range = [0.5, 0.6]
    ECfg.terrain.yaw_init_range = 0.0
    logger.configure(prefix="run_eval", root=str(tmp_path))
    env = HistoryWrapper(VelocityTrackingEasyEnv(sim_device="cuda:0", headless=True, cfg=Cfg, eval_cfg=ECfg))
    assert (env.num_envs, env.num_train_envs, env.num_eval_envs) == (128, 96, 32)
    RunnerArgs.num_steps_per_env, RunnerArgs.resume = 12, False
    runner = Runner(env, device="cuda:0")
    st = runner.alg.storage
    assert st.observations.shape[:2] == (12, 96) and st.observation_histories.shape == (12, 96, env.num_obs_history)
    core = env.env.core
    fr, pay = core.env("friction_coeffs")[0], core.env("payloads")[0]
    assert (fr[96:] >= 2.9).all() and (fr[:96] < 3.0001).all() and (fr[:96] < 2.9).any()          # creation-time draws per cfg
    assert (pay[96:] >= 2.5).all() and (pay[:96].min() < 2.0)
    assert torch.equal(core.env("rigid_payload")[0], pay)
    ms = core.env("motor_strengths")[0]
    assert ((ms[96:] >= 0.5) & (ms[96:] <= 0.6)).all() and (ms[:96] >= 0.9).all()                # env.reset() re-drew the dof props per cfg
    q = core.env("root_quat")
    assert (q[2, 96:].abs() < 0.05).all() and (q[2, :96].abs() > 0.3).any()       # eval yaw_init_range = 0 (then one settling step), train +-3.14
    od = env.get_observations()
    obs, priv, hist = od["obs"], od["privileged_obs"], od["obs_history"]
    # make the eval robots fall so that eval episodes finish inside the rollout
    core.env("root_quat")[0, 96:112] = 1.0; core.env("root_quat")[3, 96:112] = 0.0; core.env("root_quat")[2, 96:112] = 0.0
    obs, priv, hist, infos = runner.rollout(obs, priv, hist)
    assert "eval/episode" in infos and infos["env_bins"].shape[0] == 96 and infos["time_outs"].shape[0] == 96
    with torch.inference_mode():
        runner.alg.compute_returns(hist[:96], priv[:96])
    losses = runner.alg.update()
    assert all(np.isfinite(losses))
    ev = env.env.episode_sums_eval
    assert set(ev) >= {"tracking_lin_vel", "total"}
    rec = ev["tracking_lin_vel"]
    assert (rec[:96] == -1).all() and (rec[96:112] != -1).all()                                   # train envs never write it
    assert (ev["total"] == 0).all()                                                               # "total" starts at 0: never recorded (:1423)


def test_graph_replayed_rollout_equals_eager_rollout(tmp_path, monkeypatch):
    """Runner.rollout as CUDA-graph replays (one captured env step per history-buffer parity, slot / step counter / gravity read
    from device memory) must fill the rollout storage with exactly what the launch-by-launch path stores: same seeds, same
    Philox streams -> bit-identical observations, actions, rewards, dones, values, histories, env state and curriculum."""
    monkeypatch.chdir(tmp_path)
    from go1_gym_learn.ppo_cse import Runner as _R
    out = []
    for graphed in (False, True):
        torch.manual_seed(0); np.random.seed(0)
        env, Runner, RunnerArgs, logger = _make(tmp_path, n=256)
        RunnerArgs.num_steps_per_env, RunnerArgs.resume = 24, False
        runner = Runner(env, device="cuda:0")
        runner.step_graph = graphed
        if not graphed:       # launch by launch, and without the policy-only graph (its capture warm-up draws from the action-noise stream)
            runner.alg.use_cuda_graph = False
        g = torch.Generator().manual_seed(1)
        env.episode_length_buf = torch.randint(0, 1001, (256,), generator=g)
        od = env.get_observations()
        state = (od["obs"], od["privileged_obs"], od["obs_history"])
        snaps = []
        for it in range(2):                  # the second rollout replays graphs captured during the first
            obs, priv, hist, infos = runner.rollout(*state)
            state = (obs, priv, hist)
            torch.cuda.synchronize()
            st = runner.alg.storage
            snaps.append({k: getattr(st, k).clone() for k in ("observations", "privileged_observations", "observation_histories", "actions",
                                                              "rewards", "dones", "values", "actions_log_prob", "mu", "sigma", "env_bins")})
            snaps[-1]["hist"] = hist.clone(); snaps[-1]["env_f32"] = env.env.core.env_f32.clone(); snaps[-1]["leg_f32"] = env.env.core.leg_f32.clone()
            snaps[-1]["ep"] = env.env.core.episode_length_buf.clone()
            env.env._curriculum_to_host(keep_device=True)
            snaps[-1]["weights"] = torch.tensor(np.stack([c.weights for c in env.curricula]))
            snaps[-1]["rew_total"] = float(infos["train/episode"].get("rew_total", torch.tensor(0.0)))
            runner.alg.storage.clear()
        sg = runner.__dict__.get("_sg")
        assert (sg is not None and len(sg["graphs"]) == 2) if graphed else (not sg or not sg["graphs"])
        assert env.env.common_step_counter >= 48
        out.append(snaps)
    for it in range(2):
        a, b = out[0][it], out[1][it]
        assert int(a["dones"].sum()) > 0                                   # resets happened inside the rollout
        for k in a:
            if torch.is_tensor(a[k]):
                assert torch.equal(a[k], b[k]), (it, k, float((a[k].float() - b[k].float()).abs().max()))
            else:
                assert a[k] == b[k], (it, k, a[k], b[k])
