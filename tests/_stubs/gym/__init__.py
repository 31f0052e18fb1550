"""Test stub for gym: Env / Wrapper with attribute forwarding."""


class Env:
    pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)


class spaces:
    pass
