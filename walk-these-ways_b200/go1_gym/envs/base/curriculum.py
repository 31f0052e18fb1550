"""Command curricula (reference go1_gym/envs/base/curriculum.py:17-159), host-side numpy, same RandomState
call sequence so that sampling is bit-identical for a given seed (pinned by tests/golden/kats.npz)."""
import numpy as np


class Curriculum:
    def __init__(self, seed, **key_ranges):
        self.rng = np.random.RandomState(seed)
        self.cfg, idx = {}, {}
        for key, (lo, hi, n) in key_ranges.items():
            half = (hi - lo) / n / 2
            self.cfg[key] = np.linspace(lo + half, hi - half, n)
            idx[key] = np.linspace(0, n - 1, n)
        self.lows = np.array([r[0] for r in key_ranges.values()])
        self.highs = np.array([r[1] for r in key_ranges.values()])
        self.bin_sizes = {key: (hi - lo) / n for key, (lo, hi, n) in key_ranges.items()}
        self._raw_grid = np.stack(np.meshgrid(*self.cfg.values(), indexing='ij'))
        self._idx_grid = np.stack(np.meshgrid(*idx.values(), indexing='ij'))
        self.keys = [*key_ranges.keys()]
        self.grid = self._raw_grid.reshape([len(self.keys), -1])
        self.idx_grid = self._idx_grid.reshape([len(self.keys), -1])
        self._l = len(self.grid[0])
        self.ls = {key: len(self.cfg[key]) for key in self.cfg}
        self.weights = np.zeros(self._l)
        self.indices = np.arange(self._l)
        self._grid_rows = np.ascontiguousarray(self.grid.T)               # [L][D] bin centroids
        self._half_bins = np.array([*self.bin_sizes.values()]) / 2

    def __len__(self):
        return self._l

    def set_to(self, low, high, value=1.0):
        inds = np.logical_and(self.grid >= low[:, None], self.grid <= high[:, None]).all(axis=0)
        assert len(inds) != 0, "You are intializing your distribution with an empty domain!"
        self.weights[inds] = value

    def update(self, **kwargs):
        pass

    def _cdf(self):
        """cdf of the bin weights exactly as RandomState.choice(p=w/sum) builds it, cached until the weights change."""
        key = self.weights.tobytes()
        if getattr(self, "_cdf_key", None) != key:
            p = self.weights / self.weights.sum()
            cdf = p.cumsum()
            cdf /= cdf[-1]
            self._cdf_key, self._cdf_val = key, cdf
        return self._cdf_val

    def sample_bins(self, batch_size, low=None, high=None):
        if low is not None and high is not None:
            valid = np.logical_and(self.grid >= low[:, None], self.grid <= high[:, None]).all(axis=0)
            w = np.zeros_like(self.weights)
            w[valid] = self.weights[valid]
            inds = self.rng.choice(self.indices, batch_size, p=w / w.sum())
        else:
            # RandomState.choice(a, n, p=p) == cdf.searchsorted(random_sample(n), side='right') (numpy legacy generator):
            # same draws from the same stream, without rebuilding the cdf on every call
            inds = self._cdf().searchsorted(self.rng.random_sample(batch_size), side='right')
        return self._grid_rows[inds], inds

    def sample_uniform_from_cell(self, centroids):
        bin_sizes = np.array([*self.bin_sizes.values()])
        return self.rng.uniform(centroids + bin_sizes / 2, centroids - bin_sizes / 2)

    def sample(self, batch_size, low=None, high=None):
        centroids, inds = self.sample_bins(batch_size, low=low, high=high)
        # one vectorised draw: RandomState.uniform fills its output in C order, so a [batch, D] call consumes the stream
        # exactly like the reference's `batch` sequential D-wide calls (curriculum.py:87-89) — bit-identical samples
        return self.rng.uniform(centroids + self._half_bins, centroids - self._half_bins), inds


class SumCurriculum(Curriculum):
    def __init__(self, seed, **kwargs):
        super().__init__(seed, **kwargs)
        self.success = np.zeros(len(self))
        self.trials = np.zeros(len(self))

    def update(self, bin_inds, l1_error, threshold):
        ok = l1_error < threshold
        self.success[bin_inds[ok]] += 1
        self.trials[bin_inds] += 1

    def success_rates(self, *keys):
        s_rate = (self.success / (self.trials + 1e-6)).reshape(list(self.ls.values()))
        marginals = tuple(i for i, key in enumerate(self.keys) if key not in keys)
        return s_rate.mean(axis=marginals) if marginals else s_rate


class RewardThresholdCurriculum(Curriculum):
    def __init__(self, seed, **kwargs):
        super().__init__(seed, **kwargs)
        n = len(self)
        self.episode_reward_lin, self.episode_reward_ang = np.zeros(n), np.zeros(n)
        self.episode_lin_vel_raw, self.episode_ang_vel_raw, self.episode_duration = np.zeros(n), np.zeros(n), np.zeros(n)

    def get_local_bins(self, bin_inds, ranges=0.1):
        if isinstance(ranges, float):
            ranges = np.ones(self.grid.shape[0]) * ranges
        bin_inds = bin_inds.reshape(-1)
        centre = self.grid[:, bin_inds, None]                    # [D, k, 1]
        g = self.grid[:, None, :]                                # [D, 1, L]
        r = ranges.reshape(-1, 1, 1)
        return np.logical_and(g >= centre - r, g <= centre + r).all(axis=0)

    def update(self, bin_inds, task_rewards, success_thresholds, local_range=0.5):
        """task_rewards: list of float32 arrays (or torch tensors); success = every reward above its threshold."""
        if len(success_thresholds) == 0:
            ok = np.array([False] * len(bin_inds))
        else:
            ok = np.ones(len(bin_inds), dtype=bool)
            for rew, thr in zip(task_rewards, success_thresholds):
                rew = rew.cpu().numpy() if hasattr(rew, "cpu") else np.asarray(rew)
                ok &= rew.astype(np.float32) > np.float32(thr)
        self.apply_successes(bin_inds[ok], local_range)

    def apply_successes(self, ok_bins, local_range):
        """Weight update for the bins of successful envs (curriculum.py:141-154): +0.2 on the bins themselves (once per
        distinct bin), then +0.2 on every bin within local_range of each successful env's bin, one env after another.
        The neighbour lists depend only on the constant grid, so they are built once per (bin, range) and cached."""
        if len(ok_bins) == 0:
            return
        w = self.weights
        w[ok_bins] = np.clip(w[ok_bins] + 0.2, 0, 1)
        cache = self.__dict__.setdefault("_adjacent", {})
        rkey = local_range if isinstance(local_range, float) else np.asarray(local_range).tobytes()
        for b in ok_bins.reshape(-1).tolist():
            adj = cache.get((b, rkey))
            if adj is None:
                adj = cache[(b, rkey)] = np.flatnonzero(self.get_local_bins(np.array([b]), ranges=local_range)[0])
            w[adj] = np.clip(w[adj] + 0.2, 0, 1)

    def log(self, bin_inds, lin_vel_raw=None, ang_vel_raw=None, episode_duration=None):
        self.episode_lin_vel_raw[bin_inds] = lin_vel_raw.cpu().numpy()
        self.episode_ang_vel_raw[bin_inds] = ang_vel_raw.cpu().numpy()
        self.episode_duration[bin_inds] = episode_duration.cpu().numpy()
