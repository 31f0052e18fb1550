"""Running-mean caches the Runner dumps with the curriculum (same API as the reference's go1_gym_learn/ppo/metrics_caches.py:
`log(**values)` / `log(slots, **values)` fold new samples into per-key running means, `get_summary()` returns the means and
resets).  One implementation serves both: DistCache averages whole arrays, SlotCache averages per slot of an [n] vector."""
import numpy as np


class _RunningMeans:
    """mean_k <- mean_k + (x - mean_k) / count_k, kept per key (and per slot when `index` selects a subset)."""

    def __init__(self, zero):
        self._zero = zero
        self.cache = {}            # key -> running mean           (public, like the reference's attribute)
        self._seen = {}            # key -> number of samples folded in

    def _fold(self, values, index):
        for key, x in values.items():
            if key not in self.cache:
                self.cache[key], self._seen[key] = self._zero(), self._zero()
            if index is None:
                self._seen[key] = self._seen[key] + 1
                self.cache[key] = (x + (self._seen[key] - 1) * self.cache[key]) / self._seen[key]
            else:
                k = self._seen[key][index] + 1
                self._seen[key][index] = k
                self.cache[key][index] = (x + (k - 1) * self.cache[key][index]) / k

    def get_summary(self):
        out, self.cache, self._seen = self.cache, {}, {}
        return out


class DistCache(_RunningMeans):
    def __init__(self):
        super().__init__(lambda: 0)

    def log(self, **key_vals):
        self._fold(key_vals, None)


class SlotCache(_RunningMeans):
    def __init__(self, n):
        self.n = n
        super().__init__(lambda: np.zeros([n]))

    def log(self, slots=None, **key_vals):
        self._fold(key_vals, range(self.n) if slots is None else slots)
