"""Device-resident command curriculum: host management of the state that go1_curriculum_resample (csrc/curriculum.cu)
works on.

The curriculum objects of the reference API (`env.curricula[i]` with `.weights`, `.rng`, `env.env_command_bins`, ...) stay
ordinary host objects; while the rollout runs they are mirrored on the device and only the device copy advances.  `to_host()`
brings the host objects up to date (before logging, checkpoints, or any host-side resample); `to_device()` pushes them back."""
import ctypes as C

import numpy as np
import torch

from . import capi

_KIND = {"nominal": 0, "pronk": 1, "trot": 2, "pace": 3, "bound": 4}
_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_MASK = (1 << 64) - 1


class SplitMix64:
    """The category-draw generator shared by the host path and the device kernel (the reference uses torch.rand on the GPU
    here, so only the distribution is specified).  `random(k)` = k doubles in [0, 1) with 53 random bits."""

    def __init__(self, seed):
        self.state = int(seed) & _MASK

    def random(self, k):
        k = int(k)
        with np.errstate(over="ignore"):
            z = np.uint64(self.state) + _GAMMA * np.arange(1, k + 1, dtype=np.uint64)
            self.state = (self.state + 0x9E3779B97F4A7C15 * k) & _MASK
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


class DeviceCurriculum:
    def __init__(self, env, local_range, task_keys, process_group=None, emulate_world_rank=None):
        core, cfg = env.core, env.cfg
        # cross-rank replay: one curriculum shared by the envs of all ranks, kept identical on every rank by
        # running the (deterministic) kernel on the all-gathered event records
        self.group = process_group
        self.world = self.rank = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        if emulate_world_rank is not None:           # tests: the kernel side of the cross-rank replay without a process group
            self.world, self.rank = emulate_world_rank
        self.shared = self.world > 1
        self.env, self.core = env, core
        dev, N = core.device, core.N
        curs = env.curricula
        ncat, L, D = len(curs), len(curs[0]), curs[0].grid.shape[0]
        if ncat > capi.CUR_MAX_CATEGORIES:
            raise capi.Go1Error("too many curriculum categories for the device curriculum")
        for cur in curs:
            if len(cur) != L or not np.array_equal(cur.grid, curs[0].grid):
                raise capi.Go1Error("device curriculum: all categories must share one grid")
        self.ncat, self.L, self.D, self.N = ncat, L, D, N
        c = self.cfg = capi.Go1CurriculumConfig()
        c.num_categories, c.num_bins, c.num_dims, c.num_commands = ncat, L, D, int(cfg.commands.num_commands)
        for i, name in enumerate(env.category_names):
            c.category_kind[i] = _KIND[name]
        ep_len, cols, thr = env._resample_constants()
        c.num_task_keys = len(cols)
        for q, (col, th) in enumerate(zip(cols, thr)):
            c.task_col[q], c.threshold[q] = int(col), float(th)
        c.ep_len = float(ep_len)
        cc = cfg.commands
        c.gaitwise_curricula, c.exclusive_phase_offset = int(bool(cc.gaitwise_curricula)), int(bool(cc.exclusive_phase_offset))
        c.balance_gait_distribution, c.binary_phases = int(bool(cc.balance_gait_distribution)), int(bool(cc.binary_phases))
        c.num_train_envs, c.snapshot_time_outs = int(env.num_train_envs), int(bool(cfg.env.send_timeouts))
        NT = N * self.world if self.shared else N
        if self.shared:
            if self.world > capi.CUR_MAX_CATEGORIES:
                raise capi.Go1Error("cross-rank curriculum replay supports up to 8 ranks")
            c.xr_world, c.xr_rank, c.xr_cap = self.world, self.rank, N

        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.weights = torch.zeros(ncat, L, **f64)
        self.grid = torch.as_tensor(np.ascontiguousarray(curs[0].grid.T)).to(dev)            # [L][D]
        self.half_bins = torch.as_tensor(np.array([*curs[0].bin_sizes.values()]) / 2).to(dev)
        self.local_range = torch.as_tensor(np.asarray(local_range, dtype=np.float64)).to(dev)
        self.mt = torch.zeros(ncat, 625, **i32)                                              # uint32 words in int32 storage
        self.cat_rng = torch.zeros(1, dtype=torch.int64, device=dev)
        self.env_bins = torch.zeros(N, **i32)
        self.env_categories = torch.zeros(N, **i32)
        self.env_bins_f32 = torch.zeros(env.num_train_envs, device=dev)
        self.time_outs_u8 = torch.zeros(env.num_train_envs, dtype=torch.uint8, device=dev)
        self.time_outs = self.time_outs_u8.view(torch.bool)
        self.cdf = torch.zeros(ncat, L, **f64)
        self.cdf_valid = torch.zeros(ncat, **i32)
        self.scratch_i32 = torch.zeros(8 * NT + 64, **i32)
        self.scratch_u32 = torch.zeros(2 * (D + 1) * max(NT, 1024), **i32)
        self.scratch_f64 = torch.zeros((D + 2) * max(NT, 1024), **f64)
        self.out_count = torch.zeros(1, **i32)
        self.out_ids = torch.zeros(N, **i32)
        self.out_commands = torch.zeros(N, capi.NUM_COMMANDS, device=dev)
        self.xr_send = self.xr_recv = self.xr_ids = self.xr_commands = None
        if self.shared:
            blk = 1 + N * capi.XR_STRIDE
            self.xr_send = torch.zeros(2, blk, device=dev)
            self.xr_recv = torch.zeros(self.world, 2, blk, device=dev)
            self.xr_ids = torch.zeros(NT, **i32)
            self.xr_commands = torch.zeros(NT, capi.NUM_COMMANDS, device=dev)
        b = self.bufs = capi.Go1CurriculumBuffers()
        if self.shared:
            b.xr_send, b.xr_events = self.xr_send.data_ptr(), self.xr_recv.data_ptr()
            b.xr_ids, b.xr_commands = self.xr_ids.data_ptr(), self.xr_commands.data_ptr()
        for name, t in (("weights", self.weights), ("grid", self.grid), ("half_bins", self.half_bins), ("local_range", self.local_range),
                        ("mt", self.mt), ("cat_rng", self.cat_rng), ("env_bins", self.env_bins), ("env_categories", self.env_categories),
                        ("env_bins_f32", self.env_bins_f32), ("time_outs_snapshot", self.time_outs_u8), ("cdf", self.cdf),
                        ("cdf_valid", self.cdf_valid), ("scratch_i32", self.scratch_i32), ("scratch_u32", self.scratch_u32),
                        ("scratch_f64", self.scratch_f64), ("out_count", self.out_count), ("out_ids", self.out_ids),
                        ("out_commands", self.out_commands)):
            setattr(b, name, t.data_ptr())
        self._cfg_ref, self._bufs_ref = C.byref(self.cfg), C.byref(self.bufs)
        self.on_device = False

    # ------------------------------------------------------------------ host <-> device
    def to_device(self):
        """Push the host curriculum objects (weights, RandomState streams, per-env bins/categories, category rng)."""
        if self.on_device:
            return
        env, dev = self.env, self.core.device
        self.weights.copy_(torch.as_tensor(np.stack([cur.weights for cur in env.curricula])))
        mt = np.zeros((self.ncat, 625), dtype=np.uint32)
        for i, cur in enumerate(env.curricula):
            kind, key, pos, has_gauss, _ = cur.rng.get_state()
            assert kind == "MT19937" and has_gauss == 0
            mt[i, :624], mt[i, 624] = key, pos
        self.mt.copy_(torch.as_tensor(mt.view(np.int32)))
        self.cat_rng.copy_(torch.as_tensor(np.array([env._cat_rng.state], dtype=np.uint64).view(np.int64)))
        self.env_bins.copy_(torch.as_tensor(np.asarray(env.env_command_bins, dtype=np.int32)))
        self.env_categories.copy_(torch.as_tensor(np.asarray(env.env_command_categories, dtype=np.int32)))
        self.env_bins_f32.copy_(torch.as_tensor(np.asarray(env.env_command_bins[:env.num_train_envs], dtype=np.float32)))
        self.cdf_valid.zero_()
        self.on_device = True

    def to_host(self, keep_device=False):
        """Bring env.curricula / env_command_bins / env_command_categories / the category rng up to date (one sync)."""
        if not self.on_device:
            return
        env = self.env
        w = self.weights.cpu().numpy()
        mt = self.mt.cpu().numpy().view(np.uint32)
        for i, cur in enumerate(env.curricula):
            cur.weights[:] = w[i]
            cur.rng.set_state(("MT19937", mt[i, :624].copy(), int(mt[i, 624]), 0, 0.0))
        env._cat_rng.state = int(self.cat_rng.cpu().numpy().view(np.uint64)[0])
        env.env_command_bins[:] = self.env_bins.cpu().numpy()
        env.env_command_categories[:] = self.env_categories.cpu().numpy()
        if not keep_device:
            self.on_device = False

    # ------------------------------------------------------------------ calls
    def resample(self, which):
        """which = 0: terminated envs -> out_ids/out_commands/out_count; 1: periodic resample, applied in place."""
        capi.check(self.core.L.go1_curriculum_resample(self.core._handle, self._cfg_ref, self._bufs_ref, int(which), capi.stream_ptr()),
                   "go1_curriculum_resample")

    def gather(self):
        """Cross-rank replay: pack this rank's two event lists (global ids, current bins / categories) and all-gather them.  Called
        once per env step right after go1_sim_step; the step's resample(0) and the next step's resample(1) consume the result."""
        if not self.shared:
            return
        import torch.distributed as dist
        capi.check(self.core.L.go1_curriculum_pack(self.core._handle, self._cfg_ref, self._bufs_ref, capi.stream_ptr()), "go1_curriculum_pack")
        dist.all_gather_into_tensor(self.xr_recv.view(-1, self.xr_recv.shape[-1]), self.xr_send, group=self.group)     # [world * 2][block]

    def reset_envs(self, actions, post_step, common_step, episode_acc):
        capi.check(self.core.L.go1_sim_reset_idx_dev(self.core._handle, self.out_ids.data_ptr(), self.out_count.data_ptr(),
                                                     self.out_commands.data_ptr(), capi.ptr(actions) if actions is not None else None,
                                                     int(bool(post_step)), int(common_step), capi.ptr(episode_acc), capi.stream_ptr()),
                   "go1_sim_reset_idx_dev")
