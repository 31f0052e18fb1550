#!/usr/bin/env python
"""Steps a bench configuration with random-policy actions and reports the first non-finite value in the sim state / outputs
(step, env, row names), optionally with features of the configuration switched off one at a time (--off a,b,c).
    python walk-these-ways_b200/tools/nan_hunt.py --config rough_dr --envs 4096 --steps 300"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (sets sys.path for the package)
import torch  # noqa: E402


def build(config, envs, off):
    import numpy as np
    for m in [k for k in sys.modules if k.startswith("go1_gym.envs.base.legged_robot_config")]:
        del sys.modules[m]
    from go1_gym.envs.base.legged_robot_config import Cfg
    from go1_b200.train_config import apply_train_config
    from go1_gym.envs.go1.velocity_tracking import VelocityTrackingEasyEnv
    apply_train_config(Cfg)
    if config == "rough_dr":
        for sec, kv in bench.ROUGH_DR_OVERRIDES.items():
            for k, v in kv.items():
                setattr(getattr(Cfg, sec), k, v)
    for name in off:
        if name == "terrain":
            Cfg.terrain.mesh_type = "trimesh"; Cfg.terrain.terrain_proportions = [0, 0, 0, 0, 0, 0, 0, 0, 1.0]; Cfg.terrain.terrain_noise_magnitude = 0.0
        elif name == "teleport":
            Cfg.terrain.teleport_robots = False
        elif name == "push":
            Cfg.domain_rand.push_robots = False
        elif name == "rigids":
            Cfg.domain_rand.randomize_rigids_after_start = False
        elif name == "com":
            Cfg.domain_rand.randomize_com_displacement = False
        elif name == "init_range":
            Cfg.terrain.x_init_range = Cfg.terrain.y_init_range = 0.2
        elif name == "gravity":
            Cfg.domain_rand.randomize_gravity = False
    np.random.seed(0); torch.manual_seed(0)
    Cfg.env.num_envs = envs
    return VelocityTrackingEasyEnv(sim_device="cuda:0", headless=True, cfg=Cfg)


def hunt(config, envs, steps, off, scale):
    env = build(config, envs, off)
    core = env.core
    from go1_b200.sim import _FIELD_NAMES
    from go1_b200 import capi
    env.reset()
    env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length))
    g = torch.Generator(device="cuda").manual_seed(1)
    worst = dict(z=0.0, v=0.0)
    for t in range(steps):
        a = torch.randn(envs, 12, device="cuda", generator=g) * scale
        obs, rew, done, ex = env.step(a)
        bad = None
        for kind, buf in ((0, core.env_f32), (1, core.leg_f32)):
            fin = torch.isfinite(buf)
            if not fin.all():
                r, c = torch.nonzero(~fin)[0].tolist()
                names = sorted((capi.row(kind, n), n) for n in _FIELD_NAMES[kind])
                field = [n for rr, n in names if rr <= r][-1]
                bad = (kind, r, field, c if kind == 0 else c // 4)
                break
        if bad is None and not (torch.isfinite(obs).all() and torch.isfinite(rew).all()):
            bad = ("obs/rew", int(torch.nonzero(~torch.isfinite(obs).all(1) | ~torch.isfinite(rew))[0]))
        worst["z"] = max(worst["z"], float(core.env("root_pos")[2].abs().max())); worst["v"] = max(worst["v"], float(core.env("root_lin_vel").abs().max()))
        if bad is not None:
            e = bad[-1]
            print(f"[{config} off={off}] first non-finite at step {t}: {bad}; env {e}: pos {core.env('root_pos')[:, e].tolist()} quat {core.env('root_quat')[:, e].tolist()} "
                  f"vel {core.env('root_lin_vel')[:, e].tolist()} origin {core.env('env_origins')[:, e].tolist()} com {core.env('rigid_com')[:, e].tolist()} "
                  f"payload {float(core.env('rigid_payload')[0, e])} ep_len {int(core.episode_length_buf[e])}")
            return False
    print(f"[{config} off={off}] {steps} steps finite; max |z| {worst['z']:.2f}, max |v| {worst['v']:.2f}, resets/step {float(done.float().mean()) * envs:.1f}")
    return True


def hunt_train(config, envs, iterations):
    """Whole training iterations of a bench configuration with a finiteness check after every phase."""
    env, runner = bench.build_training(envs, "cuda:0", 1, config)
    od = env.get_observations()
    state = [od["obs"], od["privileged_obs"], od["obs_history"]]
    st = runner.alg.storage
    ac = runner.alg.actor_critic

    def check(tag, tensors):
        for name, t in tensors.items():
            if not torch.isfinite(t.float()).all():
                bad = torch.nonzero(~torch.isfinite(t.float()))[0].tolist()
                print(f"[{config}] iteration {it}: non-finite {name} after {tag} at {bad}: {t[tuple(bad)].item() if len(bad) == t.dim() else '?'}")
                return False
        return True
    for it in range(iterations):
        obs, priv, hist, infos = runner.rollout(*state)
        state[:] = [obs, priv, hist]
        core = env.env.core
        if not check("rollout", dict(obs=st.observations, priv=st.privileged_observations, hist=st.observation_histories, actions=st.actions,
                                     rewards=st.rewards, values=st.values, logp=st.actions_log_prob, mu=st.mu, env_bins=st.env_bins,
                                     env_f32=core.env_f32, leg_f32=core.leg_f32)):
            return False
        with torch.inference_mode():
            runner.alg.compute_returns(hist[:env.num_train_envs], priv[:env.num_train_envs])
        if not check("compute_returns", dict(returns=st.returns, advantages=st.advantages)):
            print("   rewards min/max", float(st.rewards.min()), float(st.rewards.max()), "values min/max", float(st.values.min()), float(st.values.max()))
            return False
        losses = runner.alg.update()
        if not check("update", dict(params=ac.flat_params, losses=torch.tensor(losses[:6]))):
            print("   losses", losses, "grad head", ac.flat_grads[:10].tolist(), "lr", runner.alg.learning_rate)
            print("   rewards min/max", float(st.rewards.min()), float(st.rewards.max()), "adv min/max", float(st.advantages.min()), float(st.advantages.max()),
                  "returns min/max", float(st.returns.min()), float(st.returns.max()), "logp min", float(st.actions_log_prob.min()))
            return False
    print(f"[{config}] {iterations} training iterations finite; last losses {[round(x, 4) for x in losses[:3]]}")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="rough_dr")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--bisect", action="store_true", help="if the full configuration fails, retry with each feature switched off")
    ap.add_argument("--train", type=int, default=0, help="instead: run this many whole training iterations with finiteness checks")
    a = ap.parse_args()
    if a.train:
        raise SystemExit(0 if hunt_train(a.config, a.envs, a.train) else 1)
    ok = hunt(a.config, a.envs, a.steps, [], a.scale)
    if not ok and a.bisect:
        for f in ("terrain", "com", "rigids", "push", "teleport", "init_range", "gravity"):
            hunt(a.config, a.envs, a.steps, [f], a.scale)
