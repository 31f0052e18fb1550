"""Small torch helpers of the reference (go1_gym/utils/math_utils.py:12-38)."""
import numpy as np
import torch


def quat_apply(q, v):
    xyz = q[:, :3]
    t = torch.cross(xyz, v, dim=-1) * 2
    return v + q[:, 3:] * t + torch.cross(xyz, t, dim=-1)


def quat_apply_yaw(quat, vec):
    q = quat.clone().view(-1, 4)
    q[:, :2] = 0.
    q = q / q.norm(dim=-1, keepdim=True).clamp(min=1e-9)
    return quat_apply(q, vec)


def wrap_to_pi(angles):
    angles %= 2 * np.pi
    angles -= 2 * np.pi * (angles > np.pi)
    return angles


def get_scale_shift(range):
    return 2. / (range[1] - range[0]), (range[1] + range[0]) / 2.
