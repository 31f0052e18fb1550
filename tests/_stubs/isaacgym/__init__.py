"""Test stub for Isaac Gym Preview 4 (proprietary, absent). Only torch_utils carries arithmetic:
public xyzw-quaternion formulas restated from their definitions."""
from . import gymapi, gymtorch, gymutil, terrain_utils, torch_utils  # noqa
