"""Test stub for ml_logger: a no-op logger with a truthy prefix."""


class _Ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class ML_Logger:
    prefix = "stub"

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        def _noop(*a, **k):
            return _Ctx()
        return _noop

    def every(self, *a, **k):
        return False

    def since(self, *a, **k):
        return 0.0

    def split(self, *a, **k):
        return 0.0


logger = ML_Logger()
