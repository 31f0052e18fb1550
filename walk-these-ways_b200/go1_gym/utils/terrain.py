"""Terrain: the grid of sub-terrain tiles (rows = difficulty levels, columns = terrain types), the int16 height field the
simulator samples, and the spawn origin of every tile.  Same observable behaviour as the reference class
(go1_gym/utils/terrain.py:12-180) -- attribute names, config side effects, tile order, numpy RNG consumption -- organised
around a table of terrain-type bands instead of an if-chain.  tests/test_terrain.py runs the reference's own class on the
same generators and seeds and compares height fields and origins bit for bit (tests/golden/terrain.npz)."""
from types import SimpleNamespace

import numpy as np

from . import terrain_utils as tu


def _tile_parameters(cfg, difficulty):
    """Difficulty -> generator parameters (terrain.py:110-114)."""
    return SimpleNamespace(slope=0.4 * difficulty, step=0.05 + 0.18 * difficulty,
                           obstacle=0.05 + difficulty * (cfg.max_platform_height - 0.05),
                           stone=1.5 * (1.05 - difficulty), gap=0.05 if difficulty == 0 else 0.1)


def _smooth_slope(T, tile, p, cfg, choice, bands):
    tu.pyramid_sloped_terrain(tile, slope=-p.slope if choice < bands[0] / 2 else p.slope, platform_size=3.)


def _rough_slope(T, tile, p, cfg, choice, bands):
    tu.pyramid_sloped_terrain(tile, slope=p.slope, platform_size=3.)
    tu.random_uniform_terrain(tile, min_height=-0.05, max_height=0.05, step=T.cfg.terrain_smoothness, downsampled_scale=0.2)


def _stairs(T, tile, p, cfg, choice, bands):
    tu.pyramid_stairs_terrain(tile, step_width=0.31, step_height=-p.step if choice < bands[2] else p.step, platform_size=3.)


def _obstacles(T, tile, p, cfg, choice, bands):
    tu.discrete_obstacles_terrain(tile, p.obstacle, 1., 2., 20, platform_size=3.)


def _stones(T, tile, p, cfg, choice, bands):
    tu.stepping_stones_terrain(tile, stone_size=p.stone, stone_distance=p.gap, max_height=0., platform_size=4.)


def _flat(T, tile, p, cfg, choice, bands):
    pass


def _noise(T, tile, p, cfg, choice, bands):
    m = cfg.terrain_noise_magnitude
    tu.random_uniform_terrain(tile, min_height=-m, max_height=m, step=0.005, downsampled_scale=0.2)


def _half_rough(T, tile, p, cfg, choice, bands):
    tu.random_uniform_terrain(tile, min_height=-0.05, max_height=0.05, step=T.cfg.terrain_smoothness, downsampled_scale=0.2)
    tile.height_field_raw[:tile.length // 2, :] = 0


# upper band index (into the cumulative terrain_proportions) -> generator.  Bands 2 and 3 are both stairs (descending below
# band 2's bound, ascending up to band 3's), which is why index 2 is absent: terrain.py:126-129 tests proportions[3] first.
_BANDS = ((0, _smooth_slope), (1, _rough_slope), (3, _stairs), (4, _obstacles), (5, _stones), (6, _flat), (7, _flat),
          (8, _noise), (9, _half_rough))


class Terrain:
    def __init__(self, cfg, num_robots, eval_cfg=None, num_eval_robots=0):
        self.cfg, self.eval_cfg, self.num_robots, self.type = cfg, eval_cfg, num_robots, cfg.mesh_type
        if self.type in ("none", "plane"):
            return
        self.train_rows, self.train_cols, self.eval_rows, self.eval_cols = self.load_cfgs()
        self.tot_rows = len(self.train_rows) + len(self.eval_rows)
        self.tot_cols = max(len(self.train_cols), len(self.eval_cols))
        cfg.env_length, cfg.env_width = cfg.terrain_length, cfg.terrain_width
        self.height_field_raw = np.zeros((self.tot_rows, self.tot_cols), dtype=np.int16)
        self.initialize_terrains()
        self.heightsamples = self.height_field_raw
        self.is_flat = not self.height_field_raw.any()
        if self.type == "trimesh":
            # API parity only: this simulator samples the height field itself (bilinear), so a stair riser is a one-cell ramp
            # rather than the vertical face slope_treshold cuts into the PhysX mesh
            self.vertices, self.triangles = tu.convert_heightfield_to_trimesh(self.height_field_raw, cfg.horizontal_scale,
                                                                              cfg.vertical_scale, cfg.slope_treshold)

    # ------------------------------------------------------------------ pixel geometry of one config's block of tiles
    @staticmethod
    def _load_cfg(c):
        px = lambda metres: int(metres / c.horizontal_scale)
        c.proportions = [np.sum(c.terrain_proportions[:k + 1]) for k in range(len(c.terrain_proportions))]
        c.num_sub_terrains = c.num_rows * c.num_cols
        c.env_origins = np.zeros((c.num_rows, c.num_cols, 3))
        c.width_per_env_pixels, c.length_per_env_pixels, c.border = px(c.terrain_length), px(c.terrain_width), px(c.border_size)
        c.tot_cols = int(c.num_cols * c.width_per_env_pixels) + 2 * c.border
        c.tot_rows = int(c.num_rows * c.length_per_env_pixels) + 2 * c.border

    def load_cfgs(self):
        """Train tiles first; eval tiles (if any) are appended below them along the row axis."""
        blocks, first_row, first_level = [], 0, 0
        for c in (self.cfg, self.eval_cfg):
            if c is None:
                blocks += [[], []]
                continue
            self._load_cfg(c)
            c.row_indices, c.col_indices = np.arange(first_row, first_row + c.tot_rows), np.arange(0, c.tot_cols)
            c.x_offset, c.rows_offset = first_row, first_level
            blocks += [c.row_indices, c.col_indices]
            first_row, first_level = first_row + c.tot_rows, first_level + c.num_rows
        return tuple(blocks)

    # ------------------------------------------------------------------ which tile goes where
    def initialize_terrains(self):
        for c in (self.cfg, self.eval_cfg):
            if c is not None:
                self._initialize_terrain(c)

    def _initialize_terrain(self, c):
        (self.curriculum if c.curriculum else self.selected_terrain if c.selected else self.randomized_terrain)(c)

    def curriculum(self, c):
        """Column = terrain type (choice sweeps the proportion bands), row = difficulty level."""
        for col in range(c.num_cols):
            for row in range(c.num_rows):
                tile = self.make_terrain(c, col / c.num_cols + 0.001, row / c.num_rows * c.difficulty_scale, c.proportions)
                self.add_terrain_to_map(c, tile, row, col)

    def randomized_terrain(self, c):
        for k in range(c.num_sub_terrains):
            row, col = np.unravel_index(k, (c.num_rows, c.num_cols))
            choice = np.random.uniform(0, 1)
            difficulty = np.random.choice([0.5, 0.75, 0.9])
            self.add_terrain_to_map(c, self.make_terrain(c, choice, difficulty, c.proportions), row, col)

    def selected_terrain(self, c):
        spec = c.terrain_kwargs
        generator = getattr(tu, spec.pop('type').split(".")[-1])           # the reference eval()s "terrain_utils.<name>"
        kwargs = spec.get("terrain_kwargs", spec) if isinstance(spec, dict) else spec.terrain_kwargs
        for k in range(c.num_sub_terrains):
            row, col = np.unravel_index(k, (c.num_rows, c.num_cols))
            tile = self._blank_tile(c)
            generator(tile, **kwargs)
            self.add_terrain_to_map(c, tile, row, col)

    @staticmethod
    def _blank_tile(c):
        return tu.SubTerrain("terrain", width=c.width_per_env_pixels, length=c.width_per_env_pixels,
                             vertical_scale=c.vertical_scale, horizontal_scale=c.horizontal_scale)

    def make_terrain(self, c, choice, difficulty, proportions):
        """The tile type is the first band whose cumulative proportion exceeds `choice` (none: a flat tile)."""
        tile = self._blank_tile(c)
        params = _tile_parameters(c, difficulty)
        for upper, generate in _BANDS:
            if upper < len(proportions) and choice < proportions[upper]:
                generate(self, tile, params, c, choice, proportions)
                break
        return tile

    def add_terrain_to_map(self, c, tile, row, col):
        """Paste the tile into the map; the env origin is the tile centre at the height of the tile's highest sample."""
        r0 = c.border + c.x_offset + row * c.length_per_env_pixels
        c0 = c.border + col * c.width_per_env_pixels
        window = self.height_field_raw[r0:r0 + c.length_per_env_pixels, c0:c0 + c.width_per_env_pixels]
        window[...] = tile.height_field_raw
        c.env_origins[row, col] = ((row + 0.5) * c.terrain_length + c.x_offset * tile.horizontal_scale,
                                   (col + 0.5) * c.terrain_width, np.max(window) * tile.vertical_scale)
