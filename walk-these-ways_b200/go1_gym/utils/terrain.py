"""Terrain (reference go1_gym/utils/terrain.py:12-180): the grid of num_rows x num_cols sub-terrain tiles (rows = difficulty
levels, cols = terrain types), the int16 height field the simulator samples, and the env origin of every tile.

Tile selection follows the reference exactly (curriculum / selected / randomised modes, the cumulative
`terrain_proportions` thresholds and the per-tile parameters of make_terrain, terrain.py:105-156); the generators live in
terrain_utils.py (a restatement of isaacgym.terrain_utils).  tests/test_terrain.py runs the reference's own Terrain class
on the same generators and seeds and compares height fields and origins bit for bit."""
import numpy as np

from . import terrain_utils


class Terrain:
    def __init__(self, cfg, num_robots, eval_cfg=None, num_eval_robots=0):
        self.cfg, self.eval_cfg, self.num_robots, self.type = cfg, eval_cfg, num_robots, cfg.mesh_type
        if self.type in ["none", "plane"]:
            return
        self.train_rows, self.train_cols, self.eval_rows, self.eval_cols = self.load_cfgs()
        self.tot_rows = len(self.train_rows) + len(self.eval_rows)
        self.tot_cols = max(len(self.train_cols), len(self.eval_cols))
        self.cfg.env_length, self.cfg.env_width = cfg.terrain_length, cfg.terrain_width
        self.height_field_raw = np.zeros((self.tot_rows, self.tot_cols), dtype=np.int16)
        self.initialize_terrains()
        self.heightsamples = self.height_field_raw
        self.is_flat = not self.height_field_raw.any()
        if self.type == "trimesh":
            # kept for API parity (terrain.py:31-35); this simulator samples the height field itself (bilinear), so a step
            # is a one-cell ramp rather than the vertical face slope_treshold would cut into the mesh
            self.vertices, self.triangles = terrain_utils.convert_heightfield_to_trimesh(
                self.height_field_raw, self.cfg.horizontal_scale, self.cfg.vertical_scale, self.cfg.slope_treshold)

    # ------------------------------------------------------------------ configuration (terrain.py:37-66)
    def load_cfgs(self):
        self._load_cfg(self.cfg)
        self.cfg.row_indices = np.arange(0, self.cfg.tot_rows)
        self.cfg.col_indices = np.arange(0, self.cfg.tot_cols)
        self.cfg.x_offset = 0
        self.cfg.rows_offset = 0
        if self.eval_cfg is None:
            return self.cfg.row_indices, self.cfg.col_indices, [], []
        self._load_cfg(self.eval_cfg)
        self.eval_cfg.row_indices = np.arange(self.cfg.tot_rows, self.cfg.tot_rows + self.eval_cfg.tot_rows)
        self.eval_cfg.col_indices = np.arange(0, self.eval_cfg.tot_cols)
        self.eval_cfg.x_offset = self.cfg.tot_rows
        self.eval_cfg.rows_offset = self.cfg.num_rows
        return self.cfg.row_indices, self.cfg.col_indices, self.eval_cfg.row_indices, self.eval_cfg.col_indices

    @staticmethod
    def _load_cfg(cfg):
        cfg.proportions = [np.sum(cfg.terrain_proportions[:i + 1]) for i in range(len(cfg.terrain_proportions))]
        cfg.num_sub_terrains = cfg.num_rows * cfg.num_cols
        cfg.env_origins = np.zeros((cfg.num_rows, cfg.num_cols, 3))
        cfg.width_per_env_pixels = int(cfg.terrain_length / cfg.horizontal_scale)
        cfg.length_per_env_pixels = int(cfg.terrain_width / cfg.horizontal_scale)
        cfg.border = int(cfg.border_size / cfg.horizontal_scale)
        cfg.tot_cols = int(cfg.num_cols * cfg.width_per_env_pixels) + 2 * cfg.border
        cfg.tot_rows = int(cfg.num_rows * cfg.length_per_env_pixels) + 2 * cfg.border

    # ------------------------------------------------------------------ tile selection (terrain.py:68-103)
    def initialize_terrains(self):
        self._initialize_terrain(self.cfg)
        if self.eval_cfg is not None:
            self._initialize_terrain(self.eval_cfg)

    def _initialize_terrain(self, cfg):
        if cfg.curriculum:
            self.curriculum(cfg)
        elif cfg.selected:
            self.selected_terrain(cfg)
        else:
            self.randomized_terrain(cfg)

    def randomized_terrain(self, cfg):
        for k in range(cfg.num_sub_terrains):
            i, j = np.unravel_index(k, (cfg.num_rows, cfg.num_cols))
            choice = np.random.uniform(0, 1)
            difficulty = np.random.choice([0.5, 0.75, 0.9])
            self.add_terrain_to_map(cfg, self.make_terrain(cfg, choice, difficulty, cfg.proportions), i, j)

    def curriculum(self, cfg):
        for j in range(cfg.num_cols):
            for i in range(cfg.num_rows):
                difficulty = i / cfg.num_rows * cfg.difficulty_scale
                choice = j / cfg.num_cols + 0.001
                self.add_terrain_to_map(cfg, self.make_terrain(cfg, choice, difficulty, cfg.proportions), i, j)

    def selected_terrain(self, cfg):
        terrain_type = cfg.terrain_kwargs.pop('type')
        generator = getattr(terrain_utils, terrain_type.split(".")[-1])          # the reference eval()s "terrain_utils.<name>"
        kwargs = cfg.terrain_kwargs.get("terrain_kwargs", cfg.terrain_kwargs) if isinstance(cfg.terrain_kwargs, dict) \
            else cfg.terrain_kwargs.terrain_kwargs
        for k in range(cfg.num_sub_terrains):
            i, j = np.unravel_index(k, (cfg.num_rows, cfg.num_cols))
            tile = self._new_tile(cfg)
            generator(tile, **kwargs)
            self.add_terrain_to_map(cfg, tile, i, j)

    @staticmethod
    def _new_tile(cfg):
        return terrain_utils.SubTerrain("terrain", width=cfg.width_per_env_pixels, length=cfg.width_per_env_pixels,
                                        vertical_scale=cfg.vertical_scale, horizontal_scale=cfg.horizontal_scale)

    def make_terrain(self, cfg, choice, difficulty, proportions):
        """terrain.py:105-156: the tile type is the first cumulative proportion above `choice`; difficulty scales slope,
        step height, obstacle height and stepping-stone size."""
        tile = self._new_tile(cfg)
        slope = difficulty * 0.4
        step_height = 0.05 + 0.18 * difficulty
        discrete_obstacles_height = 0.05 + difficulty * (cfg.max_platform_height - 0.05)
        stepping_stones_size = 1.5 * (1.05 - difficulty)
        stone_distance = 0.05 if difficulty == 0 else 0.1
        if choice < proportions[0]:                       # smooth pyramid slope (downhill for the first half of the band)
            if choice < proportions[0] / 2:
                slope *= -1
            terrain_utils.pyramid_sloped_terrain(tile, slope=slope, platform_size=3.)
        elif choice < proportions[1]:                     # rough pyramid slope
            terrain_utils.pyramid_sloped_terrain(tile, slope=slope, platform_size=3.)
            terrain_utils.random_uniform_terrain(tile, min_height=-0.05, max_height=0.05, step=self.cfg.terrain_smoothness,
                                                 downsampled_scale=0.2)
        elif choice < proportions[3]:                     # stairs: down for band 2, up for band 3
            if choice < proportions[2]:
                step_height *= -1
            terrain_utils.pyramid_stairs_terrain(tile, step_width=0.31, step_height=step_height, platform_size=3.)
        elif choice < proportions[4]:                     # discrete obstacles
            terrain_utils.discrete_obstacles_terrain(tile, discrete_obstacles_height, 1., 2., 20, platform_size=3.)
        elif choice < proportions[5]:
            terrain_utils.stepping_stones_terrain(tile, stone_size=stepping_stones_size, stone_distance=stone_distance,
                                                  max_height=0., platform_size=4.)
        elif choice < proportions[6]:
            pass
        elif choice < proportions[7]:
            pass
        elif choice < proportions[8]:                     # uniform noise of Cfg.terrain.terrain_noise_magnitude (train.py: 0 -> flat)
            terrain_utils.random_uniform_terrain(tile, min_height=-cfg.terrain_noise_magnitude, max_height=cfg.terrain_noise_magnitude,
                                                 step=0.005, downsampled_scale=0.2)
        elif choice < proportions[9]:                     # half flat, half rough
            terrain_utils.random_uniform_terrain(tile, min_height=-0.05, max_height=0.05, step=self.cfg.terrain_smoothness,
                                                 downsampled_scale=0.2)
            tile.height_field_raw[0:tile.length // 2, :] = 0
        return tile

    def add_terrain_to_map(self, cfg, tile, row, col):
        """terrain.py:158-180: paste the tile, record the env origin (tile centre, z = highest point of the tile)."""
        start_x = cfg.border + row * cfg.length_per_env_pixels + cfg.x_offset
        end_x = cfg.border + (row + 1) * cfg.length_per_env_pixels + cfg.x_offset
        start_y = cfg.border + col * cfg.width_per_env_pixels
        end_y = cfg.border + (col + 1) * cfg.width_per_env_pixels
        self.height_field_raw[start_x:end_x, start_y:end_y] = tile.height_field_raw
        env_origin_x = (row + 0.5) * cfg.terrain_length + cfg.x_offset * tile.horizontal_scale
        env_origin_y = (col + 0.5) * cfg.terrain_width
        env_origin_z = np.max(self.height_field_raw[start_x:end_x, start_y:end_y]) * tile.vertical_scale
        cfg.env_origins[row, col] = [env_origin_x, env_origin_y, env_origin_z]
