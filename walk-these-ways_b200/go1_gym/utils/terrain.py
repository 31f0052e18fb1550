"""Terrain (reference go1_gym/utils/terrain.py:12-180): tile grid, env origins and the int16 height field.
Round-1 scope: the flat tiles of scripts/train.py (terrain_proportions [0,...,1.0] with
terrain_noise_magnitude 0 -> random_uniform_terrain(-0, 0) == flat).  The rough generators
(isaacgym.terrain_utils: slopes, stairs, obstacles, stepping stones) are the next row of SURVEY.md §8f."""
import numpy as np


class Terrain:
    def __init__(self, cfg, num_robots, eval_cfg=None, num_eval_robots=0):
        if eval_cfg is not None:
            raise NotImplementedError("eval_cfg terrains (SURVEY.md §8f row 3)")
        self.cfg, self.num_robots, self.type = cfg, num_robots, cfg.mesh_type
        if self.type in ["none", "plane"]:
            return
        cfg.proportions = [np.sum(cfg.terrain_proportions[:i + 1]) for i in range(len(cfg.terrain_proportions))]
        cfg.num_sub_terrains = cfg.num_rows * cfg.num_cols
        cfg.env_origins = np.zeros((cfg.num_rows, cfg.num_cols, 3))
        cfg.width_per_env_pixels = int(cfg.terrain_length / cfg.horizontal_scale)
        cfg.length_per_env_pixels = int(cfg.terrain_width / cfg.horizontal_scale)
        cfg.border = int(cfg.border_size / cfg.horizontal_scale)
        cfg.tot_cols = int(cfg.num_cols * cfg.width_per_env_pixels) + 2 * cfg.border
        cfg.tot_rows = int(cfg.num_rows * cfg.length_per_env_pixels) + 2 * cfg.border
        cfg.x_offset = 0
        cfg.rows_offset = 0
        self.tot_rows, self.tot_cols = cfg.tot_rows, cfg.tot_cols
        cfg.env_length, cfg.env_width = cfg.terrain_length, cfg.terrain_width
        self.height_field_raw = np.zeros((self.tot_rows, self.tot_cols), dtype=np.int16)
        flat = (not cfg.curriculum and not cfg.selected and len(cfg.proportions) >= 9 and cfg.proportions[7] == 0
                and cfg.terrain_noise_magnitude == 0.0)
        if not flat:
            raise NotImplementedError("only the flat tile set of scripts/train.py is generated in this round "
                                      "(rough-terrain generators: SURVEY.md §8f row 2)")
        for i in range(cfg.num_rows):
            for j in range(cfg.num_cols):
                cfg.env_origins[i, j] = [(i + 0.5) * cfg.terrain_length, (j + 0.5) * cfg.terrain_width, 0.0]
        self.heightsamples = self.height_field_raw
        self.is_flat = True
