"""GPU parity tests of the fused step kernel (through the C-ABI) against
  * vectors produced by the reference's own code (tests/golden/env_logic.npz) for the control / post-physics parts,
  * the fp64 physics oracle (oracle/go1_physics_oracle.c) for the rigid-body substeps (parity unpinned vs PhysX).
Tolerances: fp32, 1e-5 relative on env logic (stated in-line); masks bit-exact."""
import numpy as np
import pytest
import torch

from env_golden_util import load_gold, oracle_state, train_sim_config, zero_sums

pytestmark = pytest.mark.gpu
FEET, THIGH, CALF = [4, 8, 12, 16], [2, 6, 10, 14], [3, 7, 11, 15]


def _sim(n, inject=True, **kw):
    from go1_b200.sim import SimCore
    Cfg, c, info = train_sim_config(n, **kw)
    c.rand_interval = 0                      # periodic motor DR is tested separately
    sim = SimCore(c, inject_noise=inject, inject_reset_rand=inject)
    return Cfg, c, info, sim


def _load_golden_state(sim, g, substeps_done=0):
    dev = sim.device
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    rs = T(g["in/root_states"])
    sim.env("root_pos").copy_(rs[:, 0:3].t()); sim.env("root_quat").copy_(rs[:, 3:7].t())
    sim.env("root_lin_vel").copy_(rs[:, 7:10].t()); sim.env("root_ang_vel").copy_(rs[:, 10:13].t())
    for name in ("dof_pos", "dof_vel", "last_actions", "last_last_actions", "last_dof_vel", "last_joint_pos_target",
                 "last_last_joint_pos_target", "joint_pos_err_last", "joint_pos_err_last_last", "joint_vel_last",
                 "joint_vel_last_last", "motor_offsets"):
        sim.set_joint_aos(name, T(g[f"in/{name}"]))
    lag = T(g["in/lag_buffer"])                       # [7][N][12]; entry 0 is never read (legged_robot.py:923-924)
    for i in range(6):
        sim.leg("lag_buffer")[3 * i:3 * i + 3].copy_(lag[i + 1].reshape(-1, 4, 3).permute(2, 0, 1))
    sim.env("motor_strengths")[0].copy_(T(g["in/motor_strengths"])[:, 0])
    sim.env("friction_coeffs")[0].copy_(T(g["in/friction_coeffs"])[:, 0])
    sim.env("restitutions")[0].copy_(T(g["in/restitutions"])[:, 0])
    sim.env("payloads")[0].copy_(T(g["in/payloads"]))
    sim.sync_rigid_props()
    sim.env("commands").copy_(T(g["in/commands"]).t())
    sim.env("gait_indices")[0].copy_(T(g["in/gait_indices"]))
    cf = T(g["in/contact_forces"])
    sim.set_foot_aos("foot_contact_forces", cf[:, FEET]); sim.set_foot_aos("thigh_contact_forces", cf[:, THIGH])
    sim.set_foot_aos("calf_contact_forces", cf[:, CALF])
    part = torch.zeros(sim.N, 4, 3, device=dev); part[:, 0] = cf[:, 0]
    sim.set_foot_aos("base_contact_forces_part", part)
    sim.set_foot_aos("foot_positions", T(g["in/foot_positions"])); sim.set_foot_aos("foot_velocities", T(g["in/foot_velocities"]))
    sim.set_foot_aos("prev_foot_velocities", T(g["in/prev_foot_velocities"]))
    sim.leg("last_contacts")[0].copy_(T(g["in/last_contacts"].astype(np.float32)))
    sim.episode_length_buf.copy_(T(g["in/episode_length_buf"]).to(torch.int32) - 1)   # the kernel does the += 1 itself
    gv = g["in/gravity_vec"][0]
    sim.set_gravity([0.05, -0.03, -9.8], gv)
    return T(np.clip(g["in/actions"], -10, 10))


def test_compute_torques_matches_reference():
    g = load_gold()
    Cfg, c, info, sim = _sim(64)
    actions = _load_golden_state(sim, g)
    for sub in range(2):
        sim.step(actions, mode=1)
        torch.cuda.synchronize()
        assert np.allclose(sim.joint_aos("torques").cpu().numpy(), g[f"torques/sub{sub}"], rtol=1e-5, atol=5e-5)
        assert np.allclose(sim.joint_aos("joint_pos_target").cpu().numpy(), g[f"joint_pos_target/sub{sub}"], rtol=1e-6, atol=1e-6)
    lag = g["after_torques/lag_buffer"]
    for i in range(6):
        got = sim.leg("lag_buffer")[3 * i:3 * i + 3].permute(1, 2, 0).reshape(64, 12).cpu().numpy()
        assert np.allclose(got, lag[i + 1], atol=1e-6)
    for k in ("joint_pos_err_last", "joint_pos_err_last_last", "joint_vel_last", "joint_vel_last_last"):
        assert np.allclose(sim.joint_aos(k).cpu().numpy(), g[f"after_torques/{k}"], rtol=1e-6, atol=1e-6), k


def test_post_physics_matches_reference():
    """check_termination / gait clock / 19 reward terms / sums / observations on injected physics outputs."""
    g = load_gold()
    Cfg, c, info, sim = _sim(64)
    actions = _load_golden_state(sim, g)
    sim.set_joint_aos("torques", torch.from_numpy(g["torques/sub1"]).cuda())
    sim.set_joint_aos("joint_pos_target", torch.from_numpy(g["joint_pos_target/sub1"]).cuda())
    sim.noise.copy_(torch.from_numpy(g["obs/noise_u"]))
    sim.step(actions, common_step=5, mode=2)
    torch.cuda.synchronize()
    N = 64
    cpu = lambda t: t.detach().cpu().numpy()
    for k, row in (("base_lin_vel", "base_lin_vel"), ("base_ang_vel", "base_ang_vel"), ("projected_gravity", "projected_gravity")):
        assert np.allclose(cpu(sim.env(row).t()), g[f"post/{k}"], rtol=1e-5, atol=2e-6), k
    assert np.allclose(cpu(sim.env("gait_indices")[0]), g["gait/gait_indices"], atol=2e-6)
    for k in ("foot_indices", "clock_inputs", "doubletime_clock_inputs", "halftime_clock_inputs", "desired_contact_states"):
        assert np.allclose(cpu(sim.leg(k)[0]), g[f"gait/{k}"], rtol=1e-5, atol=1e-5), k
    # masks: bit-exact
    assert np.array_equal(cpu(sim.reset_u8).astype(bool), g["term/reset_buf"])
    assert np.array_equal(cpu(sim.timeout_u8).astype(bool), g["term/time_out_buf"])
    assert np.array_equal(cpu(sim.episode_length_buf), g["in/episode_length_buf"])
    # rewards
    assert np.allclose(cpu(sim.rew), g["reward/rew_buf"], rtol=3e-5, atol=1e-7)
    assert np.allclose(cpu(sim.env("rew_buf_pos")[0]), g["reward/rew_buf_pos"], rtol=3e-5, atol=1e-7)
    assert np.allclose(cpu(sim.env("rew_buf_neg")[0]), g["reward/rew_buf_neg"], rtol=3e-5, atol=1e-7)
    assert np.array_equal(cpu(sim.leg("last_contacts")[0]) != 0, g["reward/last_contacts"])
    from go1_b200 import capi
    es, cs = cpu(sim.env("episode_sums")), cpu(sim.env("command_sums"))
    for name in g["reward/names"]:
        tid = capi.REWARD_TERMS.index(str(name))
        assert np.allclose(es[tid], g[f"episode_sums/{name}"], rtol=3e-5, atol=1e-7), name
        assert np.allclose(cs[tid], g[f"command_sums/{name}"], rtol=3e-5, atol=1e-7), name
    assert np.allclose(es[capi.NUM_REWARD_TERMS], g["episode_sums/total"], rtol=3e-5, atol=1e-7)
    for i, k in enumerate(capi.COMMAND_SUM_EXTRAS):
        assert np.allclose(cs[capi.NUM_REWARD_TERMS + i], g[f"command_sums/{k}"], rtol=1e-5, atol=1e-6), k
    # observations: produced in-kernel for envs that do not reset this step
    keep = ~g["term/reset_buf"]
    assert keep.sum() > 20
    assert np.allclose(cpu(sim.obs)[keep], g["obs/obs_buf"][keep], rtol=1e-5, atol=5e-6)
    assert np.allclose(cpu(sim.priv_obs)[keep], g["obs/privileged_obs_buf"][keep], rtol=1e-5, atol=1e-6)
    # last_* rolls (legged_robot.py:126-131) for continuing envs, untouched for resetting ones
    la = cpu(sim.joint_aos("last_actions"))
    assert np.allclose(la[keep], np.clip(g["in/actions"], -10, 10)[keep]) and np.allclose(la[~keep], g["in/last_actions"][~keep])
    assert np.allclose(cpu(sim.joint_aos("last_last_actions"))[keep], g["in/last_actions"][keep])
    assert np.allclose(cpu(sim.joint_aos("last_joint_pos_target"))[keep], g["joint_pos_target/sub1"][keep])
    assert np.allclose(cpu(sim.joint_aos("last_dof_vel"))[keep], g["in/dof_vel"][keep])
    # event list = reset env ids with their 4 curriculum sums
    rid, rsum, iid, isum = sim.fetch_events()
    assert np.array_equal(rid, np.nonzero(g["term/reset_buf"])[0])
    keys = ["tracking_lin_vel", "tracking_ang_vel", "tracking_contacts_shaped_force", "tracking_contacts_shaped_vel"]
    want = np.stack([g[f"command_sums/{k}"][rid] for k in keys], 1)
    assert np.allclose(rsum, want, rtol=3e-5, atol=1e-7)
    exp_interval = np.nonzero(~g["term/reset_buf"] & ((g["in/episode_length_buf"] + 1) % 500 == 0))[0]
    assert np.array_equal(iid, exp_interval)


def _random_phys_state(n, seed):
    rng = np.random.default_rng(seed)
    from oracle import physics as ph
    yaw = rng.uniform(-3.14, 3.14, n); roll = rng.uniform(-0.2, 0.2, n); pitch = rng.uniform(-0.2, 0.2, n)
    cy, sy, cp, sp, cr, sr = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
    quat = np.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy], 1)
    st = dict(pos=np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(0.22, 0.40, n)], 1), quat=quat,
              linvel=rng.uniform(-0.8, 0.8, (n, 3)), angvel=rng.uniform(-1.0, 1.0, (n, 3)),
              q=ph.DEFAULT_DOF_POS + rng.uniform(-0.25, 0.25, (n, 12)), qd=rng.uniform(-2, 2, (n, 12)),
              friction=rng.uniform(0.1, 3.0, n), restitution=rng.uniform(0, 0.4, n), payload=rng.uniform(-1, 3, n))
    return st, rng


def _load_phys_state(sim, st):
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    sim.env("root_pos").copy_(T(st["pos"]).t()); sim.env("root_quat").copy_(T(st["quat"]).t())
    sim.env("root_lin_vel").copy_(T(st["linvel"]).t()); sim.env("root_ang_vel").copy_(T(st["angvel"]).t())
    sim.set_joint_aos("dof_pos", T(st["q"])); sim.set_joint_aos("dof_vel", T(st["qd"]))
    sim.env("friction_coeffs")[0].copy_(T(st["friction"])); sim.env("restitutions")[0].copy_(T(st["restitution"]))
    sim.env("payloads")[0].copy_(T(st["payload"]))
    sim.sync_rigid_props()


@pytest.mark.parametrize("control", ["P", "actuator_net"])
def test_physics_step_matches_fp64_oracle(control):
    """One policy step (4 substeps) from random near-ground states vs the fp64 oracle driven by the same torques."""
    physics_vs_oracle(control, None)


def physics_vs_oracle(control, hf):
    """hf: None (flat) or (int16 samples [rows][cols], horizontal_scale, vertical_scale, border) for a height-field terrain."""
    import ctypes
    from oracle import physics as ph
    from oracle import env_oracle as eo
    from go1_b200.sim import SimCore
    n = 96
    Cfg, c, info = train_sim_config(n, cfg_overrides={"control": {"control_type": control}})
    c.rand_interval = 0
    st, rng = _random_phys_state(n, 3)
    hf_dev = None
    if hf is not None:
        samples, hs, vs, border = hf
        hf_dev = torch.from_numpy(np.ascontiguousarray(samples)).cuda()
        c.hf, c.hf_rows, c.hf_cols, c.hf_hscale, c.hf_vscale, c.hf_border = hf_dev.data_ptr(), samples.shape[0], samples.shape[1], hs, vs, border
        # start every robot the same distance above ITS ground point (bilinear height, as both simulators sample it)
        fx, fy = (st["pos"][:, 0] + border) / hs, (st["pos"][:, 1] + border) / hs
        ix, iy = fx.astype(int), fy.astype(int)
        ax, ay = fx - ix, fy - iy
        h = (samples[ix, iy] * (1 - ax) * (1 - ay) + samples[ix + 1, iy] * ax * (1 - ay) + samples[ix, iy + 1] * (1 - ax) * ay
             + samples[ix + 1, iy + 1] * ax * ay) * vs
        st["pos"][:, 2] += h
    sim = SimCore(c, inject_noise=True, inject_reset_rand=True)
    _load_phys_state(sim, st)
    actions = torch.from_numpy(rng.uniform(-1.5, 1.5, (n, 12)).astype(np.float32)).cuda()
    g = [0.0, 0.0, -9.8]
    sim.set_gravity(g, [0, 0, -1])
    sim.step(actions, mode=0)
    torch.cuda.synchronize()
    # oracle
    P = eo.params_from_sim_config(c, info["active_reward_scales"], info["dt"])
    s = dict(actions=actions.cpu().clone(), dof_pos=torch.tensor(st["q"], dtype=torch.float32), dof_vel=torch.tensor(st["qd"], dtype=torch.float32),
             lag_buffer=[torch.zeros(n, 12) for _ in range(7)], motor_offsets=torch.zeros(n, 12), motor_strengths=torch.ones(n, 12),
             Kp_factors=torch.ones(n, 12), Kd_factors=torch.ones(n, 12))
    for k in ("joint_pos_err_last", "joint_pos_err_last_last", "joint_vel_last", "joint_vel_last_last"):
        s[k] = torch.zeros(n, 12)
    net = eo.ActuatorNet()
    pp = ph.default_params()
    if hf is not None:
        hf_host = np.ascontiguousarray(samples, dtype=np.int16)
        pp.hf = hf_host.ctypes.data_as(ctypes.POINTER(ctypes.c_short))
        pp.hf_rows, pp.hf_cols, pp.hf_hscale, pp.hf_vscale, pp.hf_border = samples.shape[0], samples.shape[1], hs, vs, border
    states = [ph.make_state(st["pos"][i], st["quat"][i], st["linvel"][i], st["angvel"][i], st["q"][i], st["qd"][i]) for i in range(n)]
    drs = [ph.make_dr(st["friction"][i], st["restitution"][i], st["payload"][i]) for i in range(n)]
    cf = None
    for sub in range(4):
        tau = eo.compute_torques(s, P, net).numpy().astype(np.float64)
        cf = np.stack([ph.substep(pp, drs[i], states[i], tau[i]) for i in range(n)])
        s["dof_pos"] = torch.tensor(np.array([np.array(x.q) for x in states]), dtype=torch.float32)
        s["dof_vel"] = torch.tensor(np.array([np.array(x.qd) for x in states]), dtype=torch.float32)
    want = {k: np.array([np.array(getattr(x, k)) for x in states]) for k in ("pos", "quat", "linvel", "angvel", "q", "qd")}
    got = dict(pos=sim.env("root_pos").t().cpu().numpy(), quat=sim.env("root_quat").t().cpu().numpy(),
               linvel=sim.env("root_lin_vel").t().cpu().numpy(), angvel=sim.env("root_ang_vel").t().cpu().numpy(),
               q=sim.joint_aos("dof_pos").cpu().numpy(), qd=sim.joint_aos("dof_vel").cpu().numpy())
    tol = dict(pos=2e-4, quat=3e-4, linvel=5e-3, angvel=2e-2, q=1e-3, qd=5e-2)
    for k in want:
        err = np.abs(got[k] - want[k]).max(axis=1)
        frac = (err < tol[k]).mean()
        assert frac >= 0.95, (k, frac, np.sort(err)[-5:])
        assert np.median(err) < 0.1 * tol[k], (k, np.median(err))
    ff = sim.foot_aos("foot_contact_forces").cpu().numpy()
    errf = np.abs(ff - cf[:, FEET]).reshape(n, -1).max(axis=1)
    assert (errf < 0.5).mean() >= 0.9, np.sort(errf)[-8:]
    assert np.abs(sim.joint_aos("torques").cpu().numpy() - tau).max() < 0.05


def test_standing_robots_stay_up_and_fallen_robots_terminate():
    n = 256
    Cfg, c, info, sim = _sim(n, inject=False, cfg_overrides={"control": {"control_type": "P"}})
    from oracle import physics as ph
    sim.env("root_pos")[2].fill_(0.32)
    sim.set_joint_aos("dof_pos", torch.tensor(ph.DEFAULT_DOF_POS, dtype=torch.float32).repeat(n, 1).cuda())
    # second half starts upside down 0.3 m above ground -> trunk contact -> termination
    sim.env("root_quat")[0, n // 2:] = 1.0; sim.env("root_quat")[3, n // 2:] = 0.0
    sim.env("commands")[4].fill_(3.0); sim.env("commands")[8].fill_(0.5)
    actions = torch.zeros(n, 12, device="cuda")
    sim.set_gravity([0, 0, -9.8], [0, 0, -1])
    resets = torch.zeros(n, dtype=torch.bool, device="cuda")
    for t in range(60):
        sim.step(actions, common_step=t, mode=0)
        resets |= sim.reset_u8.bool()
    torch.cuda.synchronize()
    z = sim.env("root_pos")[2]
    assert torch.isfinite(sim.env_f32).all() and torch.isfinite(sim.leg_f32).all()
    assert not resets[:n // 2].any() and (z[:n // 2] > 0.2).all() and (z[:n // 2] < 0.34).all()
    assert resets[n // 2:].all()
    fz = sim.foot_aos("foot_contact_forces")[:n // 2, :, 2].sum(1)
    assert ((fz - 9.8 * 11.31).abs() < 15).all()
