"""GPU parity of the fused step kernel + the reset kernel against the reference's own post_physics_step run with full domain
randomisation and a train/eval split (tests/golden/env_dr.npz, produced by tests/golden/make_golden.py::make_dr_step from
/root/reference): _teleport_robots, _push_robots, the periodic _randomize_dof_props / _randomize_rigid_body_props,
check_termination, compute_reward, reset_idx (_randomize_dof_props, _randomize_rigid_body_props, _reset_dofs,
_reset_root_states, buffer clears, extras means, episode_sums_eval), compute_observations with all eleven privileged groups,
and the last_* rolls.  Every torch.rand of the reference is injected through Go1SimBuffers.reset_rand, so: bit-exact for
everything that is draws, copies, masks and counters; 1e-6 absolute for poses (yaw quaternion: sin / cos / normalise);
1e-5 relative for the reward / observation arithmetic (same tolerances as tests/test_sim_gpu.py)."""
import numpy as np
import pytest
import torch

from env_golden_util import dr_case

pytestmark = pytest.mark.gpu
FEET, THIGH, CALF = [4, 8, 12, 16], [2, 6, 10, 14], [3, 7, 11, 15]


def _load(sim, g, nt):
    dev = sim.device
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    rs = T(g["in/root_states"])
    sim.env("root_pos").copy_(rs[:, 0:3].t()); sim.env("root_quat").copy_(rs[:, 3:7].t())
    sim.env("root_lin_vel").copy_(rs[:, 7:10].t()); sim.env("root_ang_vel").copy_(rs[:, 10:13].t())
    for name in ("dof_pos", "dof_vel", "last_actions", "last_last_actions", "last_dof_vel", "last_joint_pos_target",
                 "last_last_joint_pos_target", "motor_offsets", "torques", "joint_pos_target"):
        sim.set_joint_aos(name, T(g[f"in/{name}"]))
    lag = T(g["in/lag_buffer"])
    for i in range(6):
        sim.leg("lag_buffer")[3 * i:3 * i + 3].copy_(lag[i + 1].reshape(-1, 4, 3).permute(2, 0, 1))
    for row, key in (("motor_strengths", "motor_strengths"), ("Kp_factors", "Kp_factors"), ("Kd_factors", "Kd_factors"),
                     ("friction_coeffs", "friction_coeffs"), ("restitutions", "restitutions")):
        sim.env(row)[0].copy_(T(g[f"in/{key}"])[:, 0])
    sim.env("payloads")[0].copy_(T(g["in/payloads"]))
    sim.env("com_displacements").copy_(T(g["in/com_displacements"]).t())
    sim.sync_rigid_props()
    sim.env("env_origins").copy_(T(g["in/env_origins"]).t())
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
    sim.episode_length_buf.copy_(T(g["in/episode_length_buf"]).to(torch.int32))          # pre-increment, like the reference's buffer
    from go1_b200 import capi
    es, cs = sim.env("episode_sums"), sim.env("command_sums")
    for k in g.files:
        if k.startswith("in/episode_sums/"):
            name = k[len("in/episode_sums/"):]
            es[capi.NUM_REWARD_TERMS if name == "total" else capi.REWARD_TERMS.index(name)].copy_(T(g[k]))
        if k.startswith("in/command_sums/"):
            name = k[len("in/command_sums/"):]
            row = capi.NUM_REWARD_TERMS + capi.COMMAND_SUM_EXTRAS.index(name) if name in capi.COMMAND_SUM_EXTRAS else capi.REWARD_TERMS.index(name)
            cs[row].copy_(T(g[k]))
    sim.episode_sums_eval[:, nt + 3] = 0.25
    grav = g["in/gravities"][0] + np.array([0.0, 0.0, -9.8], dtype=np.float32)
    sim.set_gravity(grav.tolist(), g["in/gravity_vec"][0])
    return T(np.clip(g["in/actions"], -10, 10))


def test_full_dr_step_and_reset_match_reference():
    from go1_b200 import capi
    from go1_b200.sim import SimCore
    g, Cfg, ECfg, c, info = dr_case()
    nt, N = int(g["meta/num_train_envs"]), g["in/root_states"].shape[0]
    assert c.dr[0].push_interval == int(g["meta/push_interval"][0]) and c.dr[1].push_interval == int(g["meta/push_interval"][1])
    assert c.rand_interval == int(g["meta/rand_interval"])
    sim = SimCore(c, inject_noise=True, inject_reset_rand=True)
    sim.enable_eval_sums()
    actions = _load(sim, g, nt)
    sim.noise.copy_(torch.from_numpy(g["obs/noise_u"]))
    sim.reset_rand.copy_(torch.from_numpy(g["rand/step"]))
    sim.step(actions, common_step=5, mode=2)
    rid, rsum, iid, isum = sim.fetch_events()
    assert np.array_equal(rid, g["reset/ids"])
    assert np.allclose(sim.env("episode_sums")[capi.NUM_REWARD_TERMS].cpu().numpy(), g["mid/episode_sums_total"], rtol=1e-5, atol=1e-6)
    sim.reset_rand.copy_(torch.from_numpy(g["rand/reset"]))
    sim.episode_acc.zero_()
    sim.reset_idx(rid, g["in/new_commands"][rid], actions=actions, post_step=True, common_step=5)
    torch.cuda.synchronize()
    cpu = lambda t: t.detach().cpu().numpy()
    # ---- masks / counters: bit-exact
    assert np.array_equal(cpu(sim.reset_u8).astype(bool), g["out/reset_buf"])
    assert np.array_equal(cpu(sim.timeout_u8).astype(bool), g["out/time_out_buf"])
    assert np.array_equal(cpu(sim.episode_length_buf), g["out/episode_length_buf"])
    # ---- draws and copies: bit-exact
    for k in ("dof_pos", "dof_vel", "last_actions", "last_last_actions", "last_dof_vel", "last_joint_pos_target",
              "last_last_joint_pos_target", "motor_offsets"):
        assert np.array_equal(cpu(sim.joint_aos(k)), g[f"out/{k}"]), k
    for row in ("motor_strengths", "Kp_factors", "Kd_factors", "friction_coeffs", "restitutions"):
        assert np.array_equal(cpu(sim.env(row)[0]), g[f"out/{row}"][:, 0]), row
    assert np.array_equal(cpu(sim.env("payloads")[0]), g["out/payloads"])
    assert np.array_equal(cpu(sim.env("com_displacements").t()), g["out/com_displacements"])
    assert np.array_equal(cpu(sim.env("commands").t()), g["out/commands"])
    lag = g["out/lag_buffer"]
    for i in range(6):
        assert np.array_equal(cpu(sim.leg("lag_buffer")[3 * i:3 * i + 3].permute(1, 2, 0).reshape(N, 12)), lag[i + 1])
    want = g["out/root_states"]
    assert np.array_equal(cpu(sim.env("root_pos").t()), want[:, 0:3])
    assert np.array_equal(cpu(sim.env("root_lin_vel").t()), want[:, 7:10]) and np.array_equal(cpu(sim.env("root_ang_vel").t()), want[:, 10:13])
    assert np.allclose(cpu(sim.env("root_quat").t()), want[:, 3:7], rtol=0, atol=1e-6)
    # the physics keeps the creation-time mass / centre of mass (re-draws only change the observed buffers)
    assert np.array_equal(cpu(sim.env("rigid_payload")[0]), g["in/payloads"]) and np.array_equal(cpu(sim.env("rigid_com").t()), g["in/com_displacements"])
    # ---- arithmetic: fp32 tolerances
    for k in ("base_lin_vel", "base_ang_vel", "projected_gravity"):
        assert np.allclose(cpu(sim.env(k).t()), g[f"out/{k}"], rtol=1e-5, atol=2e-6), k
    assert np.allclose(cpu(sim.env("gait_indices")[0]), g["out/gait_indices"], atol=2e-6)
    for k in ("foot_indices", "clock_inputs", "desired_contact_states"):
        assert np.allclose(cpu(sim.leg(k)[0]), g[f"out/{k}"], rtol=1e-5, atol=1e-5), k
    assert np.allclose(cpu(sim.rew), g["out/rew_buf"], rtol=3e-5, atol=1e-6)
    assert np.allclose(cpu(sim.env("rew_buf_pos")[0]), g["out/rew_buf_pos"], rtol=3e-5, atol=1e-6)
    assert np.allclose(cpu(sim.env("rew_buf_neg")[0]), g["out/rew_buf_neg"], rtol=3e-5, atol=1e-6)
    assert np.array_equal(cpu(sim.leg("last_contacts")[0]) != 0, g["out/last_contacts"])
    es, cs, ev = cpu(sim.env("episode_sums")), cpu(sim.env("command_sums")), cpu(sim.episode_sums_eval)
    for k in g.files:
        if k.startswith("out/episode_sums/"):
            name = k[len("out/episode_sums/"):]
            row = capi.NUM_REWARD_TERMS if name == "total" else capi.REWARD_TERMS.index(name)
            assert np.allclose(es[row], g[k], rtol=3e-5, atol=1e-6), name
            assert np.allclose(ev[row], g[f"out/episode_sums_eval/{name}"], rtol=3e-5, atol=1e-6), name
        if k.startswith("out/command_sums/"):
            name = k[len("out/command_sums/"):]
            row = capi.NUM_REWARD_TERMS + capi.COMMAND_SUM_EXTRAS.index(name) if name in capi.COMMAND_SUM_EXTRAS else capi.REWARD_TERMS.index(name)
            assert np.allclose(cs[row], g[k], rtol=3e-5, atol=1e-6), name
    # observations of ALL envs: continuing ones from the step kernel, reset ones from the reset kernel
    assert np.allclose(cpu(sim.obs), g["out/obs_buf"], rtol=1e-5, atol=5e-6)
    assert np.allclose(cpu(sim.priv_obs), g["out/privileged_obs_buf"], rtol=1e-5, atol=5e-6)
    # extras["train/episode"] reward means from the device accumulator
    acc = cpu(sim.episode_acc)
    n_train_reset = int((g["reset/ids"] < nt).sum())
    assert acc[capi.NUM_EPISODE_SUMS] == n_train_reset
    for k in g.files:
        if k.startswith("out/extras_train_episode/rew_"):
            name = k[len("out/extras_train_episode/rew_"):]
            row = capi.NUM_REWARD_TERMS if name == "total" else capi.REWARD_TERMS.index(name)
            assert np.allclose(acc[row] / n_train_reset, g[k], rtol=3e-5, atol=1e-6), name
