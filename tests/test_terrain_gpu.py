"""Rough terrain on the GPU: the step kernel on a height field vs the fp64 physics oracle, the
measured-heights body-height termination vs the torch restatement of _get_heights (legged_robot.py:1772-1806), and a whole
env on the reference's curriculum tile set."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "walk-these-ways_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "walk-these-ways_b200", "compat"))


def _smooth_field():
    hs, vs, border = 0.1, 0.005, 3.2
    x = np.arange(64) * hs - border
    xx, yy = np.meshgrid(x, x, indexing="ij")
    h = 0.06 * np.sin(1.3 * xx) * np.cos(0.9 * yy) + 0.12 * xx + 0.05 * yy
    return np.rint(h / vs).astype(np.int16), hs, vs, border


def test_physics_on_height_field_matches_fp64_oracle():
    from test_sim_gpu import physics_vs_oracle
    physics_vs_oracle("actuator_net", _smooth_field())


def _rough_env(n, measure_heights):
    for m in [k for k in sys.modules if k.startswith("go1_gym.envs.base.legged_robot_config")]:
        del sys.modules[m]
    from go1_gym.envs.base.legged_robot_config import Cfg
    from go1_b200.train_config import apply_train_config
    from go1_gym.envs.go1.velocity_tracking import VelocityTrackingEasyEnv
    apply_train_config(Cfg)
    Cfg.env.num_envs = n
    t = Cfg.terrain
    t.mesh_type, t.curriculum, t.selected = "trimesh", True, False
    t.num_rows, t.num_cols, t.border_size, t.terrain_length, t.terrain_width = 4, 4, 5, 8., 8.
    t.terrain_proportions, t.terrain_noise_magnitude = [0.1, 0.1, 0.35, 0.25, 0.2], 0.1
    t.center_robots, t.measure_heights = False, measure_heights
    t.max_init_terrain_level = 3
    np.random.seed(4); torch.manual_seed(4); torch.cuda.manual_seed_all(4)
    return VelocityTrackingEasyEnv(sim_device="cuda:0", headless=True, cfg=Cfg), Cfg


def test_env_on_curriculum_tiles_stands_and_uses_tile_origins():
    env, Cfg = _rough_env(64, False)
    assert env.sim_cfg.hf and not env.terrain.is_flat and env.height_samples.shape == (env.terrain.tot_rows, env.terrain.tot_cols)
    env.reset()
    origins = env.env_origins.clone()
    tile_z = torch.from_numpy(Cfg.terrain.env_origins).float().cuda()[env.terrain_levels, env.terrain_types][:, 2]
    assert torch.allclose(origins[:, 2], tile_z) and float(origins[:, 2].max()) > 0.05          # spawn above the tile's highest point
    resets = torch.zeros(64, device="cuda")
    for i in range(60):
        obs, rew, done, info = env.step(torch.zeros(64, 12, device="cuda"))
        resets += done.float()
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    # standing robots rest on THEIR terrain: base height above the ground sample right under the base stays near the
    # nominal 0.3 m (whatever the tile: slope, stairs, obstacles, at any curriculum level)
    t = Cfg.terrain
    ix = ((env.base_pos[:, 0] + t.border_size) / t.horizontal_scale).round().long().clamp(0, env.height_samples.shape[0] - 1)
    iy = ((env.base_pos[:, 1] + t.border_size) / t.horizontal_scale).round().long().clamp(0, env.height_samples.shape[1] - 1)
    ground = env.height_samples[ix, iy].float() * t.vertical_scale
    clearance = env.base_pos[:, 2] - ground
    # the reference spawns every robot at the HIGHEST point of its tile (terrain.py:176-178), i.e. in mid-air above pit-type tiles
    # (down-slope pyramids, descending stairs); judge standing on the tiles whose top is the spawn platform
    on_top = (ground - origins[:, 2]).abs() < 0.03
    assert int(on_top.sum()) >= 8 and int((~on_top).sum()) >= 8, (int(on_top.sum()), ground, origins[:, 2])
    cl = clearance[on_top]
    assert float(cl.min()) > 0.15 and float(cl.max()) < 0.5, cl
    assert float(origins[:, 2].max() - origins[:, 2].min()) > 0.1            # the robots really are at different terrain heights
    # robots dropped into the pits may crash and respawn; the ones standing on their platform never terminate
    assert float(resets[on_top].sum()) <= 2, resets


def test_measured_heights_termination_matches_torch_restatement():
    env, Cfg = _rough_env(128, True)
    assert env.sim_cfg.measure_heights == 1 and env.sim_cfg.num_height_points_x == 17 and env.sim_cfg.num_height_points_y == 11
    env.reset()
    core = env.core
    g = torch.Generator().manual_seed(0)
    # hover the robots 0.0 .. 0.5 m above their tile origin height, yawed, no contact: only the body-height test can fire
    z = env.env_origins[:, 2] + 0.25 + 0.5 * torch.rand(128, generator=g).cuda()
    core.env("root_pos")[2].copy_(z)
    yaw = (6.28 * torch.rand(128, generator=g) - 3.14).cuda()
    core.env("root_quat")[0].zero_(); core.env("root_quat")[1].zero_()
    core.env("root_quat")[2].copy_(torch.sin(yaw / 2)); core.env("root_quat")[3].copy_(torch.cos(yaw / 2))
    core.env("root_lin_vel").zero_(); core.env("root_ang_vel").zero_()
    env.sim_cfg.terminal_body_height = 0.45
    core.update_config()
    core.episode_length_buf.fill_(10)
    core.step(torch.zeros(128, 12, device="cuda"), common_step=7, mode=0)
    torch.cuda.synchronize()
    heights = env._get_heights(torch.arange(128, device="cuda"))
    body_height = (env.base_pos[:, 2:3] - heights).mean(1)
    want = body_height < 0.45
    margin = (body_height - 0.45).abs() > 2e-3           # the kernel evaluates the test before this step's (tiny) free fall is undone
    base_contact = env.contact_forces[:, 0].norm(dim=1) > 1.0
    assert not base_contact.any()
    got = core.reset_u8.bool()
    assert torch.equal(got[margin], want[margin]) and 10 < int(want.sum()) < 118
