"""Height-field generators used by go1_gym/utils/terrain.py.

The reference imports these from `isaacgym.terrain_utils` (Isaac Gym Preview 4, proprietary, not in the reference tree and
not installable here), so this is a restatement of that module's published algorithms -- same function names, arguments
and discrete-unit arithmetic on an int16 `SubTerrain.height_field_raw` of shape [width][length] -- anchored on the
reference's call sites (go1_gym/utils/terrain.py:105-156).  Parity with the original module is UNPINNED (no golden
vectors exist for it); what IS pinned is the reference's Terrain class driving these functions
(tests/test_terrain.py runs the reference class on top of this module and compares with ours)."""
import numpy as np


class SubTerrain:
    def __init__(self, terrain_name="terrain", width=256, length=256, vertical_scale=1.0, horizontal_scale=1.0):
        self.terrain_name = terrain_name
        self.vertical_scale = vertical_scale
        self.horizontal_scale = horizontal_scale
        self.width = width
        self.length = length
        self.height_field_raw = np.zeros((self.width, self.length), dtype=np.int16)


def _bilinear_resample(coarse, width, length):
    """Values of the piecewise-bilinear surface through `coarse` (samples equally spaced over the tile, end points
    included) at `width` x `length` equally spaced points -- what scipy's linear interp2d did in the original."""
    n0, n1 = coarse.shape
    u = np.linspace(0.0, n0 - 1.0, width) if n0 > 1 else np.zeros(width)
    v = np.linspace(0.0, n1 - 1.0, length) if n1 > 1 else np.zeros(length)
    i0 = np.minimum(np.floor(u).astype(int), max(n0 - 2, 0)); i1 = np.minimum(i0 + 1, n0 - 1)
    j0 = np.minimum(np.floor(v).astype(int), max(n1 - 2, 0)); j1 = np.minimum(j0 + 1, n1 - 1)
    a = (u - i0)[:, None]; b = (v - j0)[None, :]
    c = coarse.astype(np.float64)
    return (c[np.ix_(i0, j0)] * (1 - a) * (1 - b) + c[np.ix_(i1, j0)] * a * (1 - b)
            + c[np.ix_(i0, j1)] * (1 - a) * b + c[np.ix_(i1, j1)] * a * b)


def random_uniform_terrain(terrain, min_height, max_height, step=1, downsampled_scale=None):
    """Uniform noise sampled on a coarse grid (one sample per `downsampled_scale` metres, heights quantised to `step`)
    and bilinearly upsampled to the tile."""
    if downsampled_scale is None:
        downsampled_scale = terrain.horizontal_scale
    min_height = int(min_height / terrain.vertical_scale)
    max_height = int(max_height / terrain.vertical_scale)
    step = max(int(step / terrain.vertical_scale), 1)
    heights_range = np.arange(min_height, max_height + step, step)
    coarse = np.random.choice(heights_range, (int(terrain.width * terrain.horizontal_scale / downsampled_scale),
                                              int(terrain.length * terrain.horizontal_scale / downsampled_scale)))
    terrain.height_field_raw += np.rint(_bilinear_resample(coarse, terrain.width, terrain.length)).astype(np.int16)
    return terrain


def sloped_terrain(terrain, slope=1):
    xx = np.arange(0, terrain.width).reshape(terrain.width, 1)
    max_height = int(slope * (terrain.horizontal_scale / terrain.vertical_scale) * terrain.width)
    terrain.height_field_raw[:, np.arange(terrain.length)] += (max_height * xx / terrain.width).astype(terrain.height_field_raw.dtype)
    return terrain


def pyramid_sloped_terrain(terrain, slope=1, platform_size=1.):
    x, y = np.arange(0, terrain.width), np.arange(0, terrain.length)
    center_x, center_y = int(terrain.width / 2), int(terrain.length / 2)
    xx = ((center_x - np.abs(center_x - x)) / center_x).reshape(terrain.width, 1)
    yy = ((center_y - np.abs(center_y - y)) / center_y).reshape(1, terrain.length)
    max_height = int(slope * (terrain.horizontal_scale / terrain.vertical_scale) * (terrain.width / 2))
    terrain.height_field_raw += (max_height * xx * yy).astype(terrain.height_field_raw.dtype)
    platform_size = int(platform_size / terrain.horizontal_scale / 2)
    x1, y1 = terrain.width // 2 - platform_size, terrain.length // 2 - platform_size
    min_h = min(terrain.height_field_raw[x1, y1], 0)
    max_h = max(terrain.height_field_raw[x1, y1], 0)
    terrain.height_field_raw = np.clip(terrain.height_field_raw, min_h, max_h)
    return terrain


def discrete_obstacles_terrain(terrain, max_height, min_size, max_size, num_rects, platform_size=1.):
    max_height = int(max_height / terrain.vertical_scale)
    min_size = int(min_size / terrain.horizontal_scale)
    max_size = int(max_size / terrain.horizontal_scale)
    platform_size = int(platform_size / terrain.horizontal_scale)
    (i, j) = terrain.height_field_raw.shape
    height_range = [-max_height, -max_height // 2, max_height // 2, max_height]
    width_range = range(min_size, max_size, 4)
    length_range = range(min_size, max_size, 4)
    for _ in range(num_rects):
        width = np.random.choice(width_range)
        length = np.random.choice(length_range)
        start_i = np.random.choice(range(0, i - width, 4))
        start_j = np.random.choice(range(0, j - length, 4))
        terrain.height_field_raw[start_i:start_i + width, start_j:start_j + length] = np.random.choice(height_range)
    x1, x2 = (terrain.width - platform_size) // 2, (terrain.width + platform_size) // 2
    y1, y2 = (terrain.length - platform_size) // 2, (terrain.length + platform_size) // 2
    terrain.height_field_raw[x1:x2, y1:y2] = 0
    return terrain


def wave_terrain(terrain, num_waves=1, amplitude=1.):
    amplitude = int(0.5 * amplitude / terrain.vertical_scale)
    if num_waves > 0:
        div = terrain.length / (num_waves * np.pi * 2)
        xx = np.arange(0, terrain.width).reshape(terrain.width, 1)
        yy = np.arange(0, terrain.length).reshape(1, terrain.length)
        terrain.height_field_raw += (amplitude * np.cos(yy / div) + amplitude * np.sin(xx / div)).astype(terrain.height_field_raw.dtype)
    return terrain


def stairs_terrain(terrain, step_width, step_height):
    step_width = int(step_width / terrain.horizontal_scale)
    step_height = int(step_height / terrain.vertical_scale)
    num_steps = terrain.width // step_width
    height = step_height
    for i in range(num_steps):
        terrain.height_field_raw[i * step_width: (i + 1) * step_width, :] += height
        height += step_height
    return terrain


def pyramid_stairs_terrain(terrain, step_width, step_height, platform_size=1.):
    step_width = int(step_width / terrain.horizontal_scale)
    step_height = int(step_height / terrain.vertical_scale)
    platform_size = int(platform_size / terrain.horizontal_scale)
    height = 0
    start_x, stop_x, start_y, stop_y = 0, terrain.width, 0, terrain.length
    while (stop_x - start_x) > platform_size and (stop_y - start_y) > platform_size:
        start_x += step_width; stop_x -= step_width
        start_y += step_width; stop_y -= step_width
        height += step_height
        terrain.height_field_raw[start_x: stop_x, start_y: stop_y] = height
    return terrain


def stepping_stones_terrain(terrain, stone_size, stone_distance, max_height, platform_size=1., depth=-10):
    stone_size = int(stone_size / terrain.horizontal_scale)
    stone_distance = int(stone_distance / terrain.horizontal_scale)
    max_height = int(max_height / terrain.vertical_scale)
    platform_size = int(platform_size / terrain.horizontal_scale)
    height_range = np.arange(-max_height - 1, max_height, step=1)
    start_x = start_y = 0
    terrain.height_field_raw[:, :] = int(depth / terrain.vertical_scale)
    if terrain.length >= terrain.width:
        while start_y < terrain.length:
            stop_y = min(terrain.length, start_y + stone_size)
            start_x = np.random.randint(0, stone_size)
            stop_x = max(0, start_x - stone_distance)                       # the first (partial) stone of the row
            terrain.height_field_raw[0: stop_x, start_y: stop_y] = np.random.choice(height_range)
            while start_x < terrain.width:
                stop_x = min(terrain.width, start_x + stone_size)
                terrain.height_field_raw[start_x: stop_x, start_y: stop_y] = np.random.choice(height_range)
                start_x += stone_size + stone_distance
            start_y += stone_size + stone_distance
    else:
        while start_x < terrain.width:
            stop_x = min(terrain.width, start_x + stone_size)
            start_y = np.random.randint(0, stone_size)
            stop_y = max(0, start_y - stone_distance)
            terrain.height_field_raw[start_x: stop_x, 0: stop_y] = np.random.choice(height_range)
            while start_y < terrain.length:
                stop_y = min(terrain.length, start_y + stone_size)
                terrain.height_field_raw[start_x: stop_x, start_y: stop_y] = np.random.choice(height_range)
                start_y += stone_size + stone_distance
            start_x += stone_size + stone_distance
    x1, x2 = (terrain.width - platform_size) // 2, (terrain.width + platform_size) // 2
    y1, y2 = (terrain.length - platform_size) // 2, (terrain.length + platform_size) // 2
    terrain.height_field_raw[x1:x2, y1:y2] = 0
    return terrain


def convert_heightfield_to_trimesh(height_field_raw, horizontal_scale, vertical_scale, slope_threshold=None):
    """Vertices [rows*cols, 3] float32 and triangles [2*(rows-1)*(cols-1), 3] uint32 of the height field; where the slope
    between neighbours exceeds `slope_threshold` the lower vertex is moved under the upper one (vertical faces)."""
    hf = height_field_raw
    num_rows, num_cols = hf.shape
    y = np.linspace(0, (num_cols - 1) * horizontal_scale, num_cols)
    x = np.linspace(0, (num_rows - 1) * horizontal_scale, num_rows)
    yy, xx = np.meshgrid(y, x)
    if slope_threshold is not None:
        slope_threshold *= horizontal_scale / vertical_scale
        move_x = np.zeros((num_rows, num_cols)); move_y = np.zeros((num_rows, num_cols)); move_corners = np.zeros((num_rows, num_cols))
        move_x[:num_rows - 1, :] += (hf[1:num_rows, :] - hf[:num_rows - 1, :] > slope_threshold)
        move_x[1:num_rows, :] -= (hf[:num_rows - 1, :] - hf[1:num_rows, :] > slope_threshold)
        move_y[:, :num_cols - 1] += (hf[:, 1:num_cols] - hf[:, :num_cols - 1] > slope_threshold)
        move_y[:, 1:num_cols] -= (hf[:, :num_cols - 1] - hf[:, 1:num_cols] > slope_threshold)
        move_corners[:num_rows - 1, :num_cols - 1] += (hf[1:num_rows, 1:num_cols] - hf[:num_rows - 1, :num_cols - 1] > slope_threshold)
        move_corners[1:num_rows, 1:num_cols] -= (hf[:num_rows - 1, :num_cols - 1] - hf[1:num_rows, 1:num_cols] > slope_threshold)
        xx += (move_x + move_corners * (move_x == 0)) * horizontal_scale
        yy += (move_y + move_corners * (move_y == 0)) * horizontal_scale
    vertices = np.zeros((num_rows * num_cols, 3), dtype=np.float32)
    vertices[:, 0] = xx.flatten()
    vertices[:, 1] = yy.flatten()
    vertices[:, 2] = hf.flatten() * vertical_scale
    triangles = -np.ones((2 * (num_rows - 1) * (num_cols - 1), 3), dtype=np.uint32)
    for i in range(num_rows - 1):
        ind0 = np.arange(0, num_cols - 1) + i * num_cols
        ind1 = ind0 + 1
        ind2 = ind0 + num_cols
        ind3 = ind2 + 1
        start = 2 * i * (num_cols - 1)
        stop = start + 2 * (num_cols - 1)
        triangles[start:stop:2, 0] = ind0; triangles[start:stop:2, 1] = ind3; triangles[start:stop:2, 2] = ind1
        triangles[start + 1:stop:2, 0] = ind0; triangles[start + 1:stop:2, 1] = ind2; triangles[start + 1:stop:2, 2] = ind3
    return vertices, triangles
