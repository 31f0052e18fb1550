"""Stand-in for ml_logger (setup.py:12 of the reference): stores metrics in memory, writes files under
`root/prefix`, and never forces a device->host sync per step (tensors are kept and reduced lazily in
log_metrics_summary)."""
import os
import pickle
import time
from collections import defaultdict
from contextlib import contextmanager
from datetime import datetime


class ML_Logger:
    def __init__(self, root=None, prefix=None, **kw):
        self.root = str(root) if root is not None else os.path.abspath("./runs")
        self.prefix = prefix or "default"
        self._metrics = defaultdict(list)
        self._lazy = []
        self._metric_prefix = ""
        self._timers = {}
        self._every = defaultdict(int)
        self.summaries = []
        self.print_summary = False

    # ---- configuration
    def configure(self, prefix=None, root=None, **kw):
        if prefix is not None:
            self.prefix = str(prefix)
        if root is not None:
            self.root = str(root)
        return self

    @staticmethod
    def utcnow(fmt="%Y-%m-%d/%H%M%S.%f"):
        return datetime.utcnow().strftime(fmt)

    def _path(self, rel):
        p = os.path.join(self.root, self.prefix, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        return p

    # ---- timers
    def start(self, *keys):
        now = time.perf_counter()
        for k in keys:
            self._timers[k] = now

    def since(self, key):
        return time.perf_counter() - self._timers.get(key, time.perf_counter())

    def split(self, key):
        now = time.perf_counter()
        dt = now - self._timers.get(key, now)
        self._timers[key] = now
        return dt

    # ---- metrics
    @contextmanager
    def Prefix(self, metrics=None, **kw):
        old = self._metric_prefix
        self._metric_prefix = (metrics.rstrip("/") + "/") if metrics else old
        try:
            yield
        finally:
            self._metric_prefix = old

    @contextmanager
    def Sync(self, *a, **kw):
        yield

    def store_metrics(self, metrics=None, **kw):
        d = dict(metrics or {})
        d.update(kw)
        for k, v in d.items():
            self._metrics[self._metric_prefix + k].append(v)

    def store_metrics_lazy(self, prefix, mapping):
        """Keep a reference to a (possibly lazily built) dict; it is expanded at summary time."""
        self._lazy.append((prefix.rstrip("/") + "/", mapping))

    def every(self, n, key="default", start_on=0):
        self._every[key] += 1
        c = self._every[key]
        return c >= start_on and (c - start_on) % n == 0

    @staticmethod
    def _to_float(v):
        try:
            import torch
            if isinstance(v, torch.Tensor):
                return float(v.detach().float().mean().item())
        except ImportError:
            pass
        try:
            return float(v)
        except (TypeError, ValueError):
            return None

    def log_metrics_summary(self, key_values=None, **kw):
        row = dict(key_values or {})
        seen = set()
        for prefix, mapping in self._lazy:
            if id(mapping) in seen:
                continue
            seen.add(id(mapping))
            for k, v in mapping.items():
                self._metrics[prefix + k].append(v)
        self._lazy = []
        for k, vals in self._metrics.items():
            fs = [f for f in (self._to_float(v) for v in vals) if f is not None]
            if fs:
                row[k + "/mean"] = sum(fs) / len(fs)
        self._metrics.clear()
        self.summaries.append(row)
        with open(self._path("metrics.pkl"), "ab") as f:
            pickle.dump(row, f)
        if self.print_summary:
            print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()})
        return row

    # ---- files
    def log_params(self, **kw):
        def clean(o):
            if isinstance(o, dict):
                return {k: clean(v) for k, v in o.items() if not str(k).startswith("_")}
            if isinstance(o, type):
                return clean(vars(o))
            return o
        with open(self._path("parameters.pkl"), "ab") as f:
            pickle.dump(clean(kw), f)

    def log_text(self, text, filename="log.txt", dedent=False, **kw):
        import textwrap
        with open(self._path(filename), "a") as f:
            f.write(textwrap.dedent(text) if dedent else text)

    def save_pkl(self, data, path, append=False):
        with open(self._path(path), "ab" if append else "wb") as f:
            pickle.dump(data, f)

    def load_pkl(self, path):
        out = []
        with open(self._path(path), "rb") as f:
            while True:
                try:
                    out.append(pickle.load(f))
                except EOFError:
                    return out

    def torch_save(self, obj, path):
        import torch
        torch.save(obj, self._path(path))

    def load_torch(self, path, **kw):
        import torch
        return torch.load(self._path(path), **kw)

    def duplicate(self, src, dst):
        import shutil
        shutil.copyfile(self._path(src), self._path(dst))

    def upload_file(self, file_path, target_path="", once=True):
        import shutil
        dst = self._path(os.path.join(target_path, os.path.basename(file_path)))
        shutil.copyfile(file_path, dst)

    def save_video(self, frames, path, fps=30, **kw):
        pass      # rendering is out of scope

    def job_running(self):
        pass

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)

        def _noop(*a, **k):
            return None
        return _noop


logger = ML_Logger()
