"""Stand-in for params_proto 2.10 (setup.py:15 of the reference): config classes whose attributes are
read and mutated as class attributes (`Cfg.env.num_envs = 4000`) and whose `vars(cls)` is a plain dict."""


_RAW_DICT = type.__dict__["__dict__"]      # the real class-dict descriptor (the property below shadows it)


class Meta(type):
    def __new__(mcls, name, bases, ns, cli=False, **kw):
        return super().__new__(mcls, name, bases, ns)

    def __init__(cls, name, bases, ns, cli=False, **kw):
        super().__init__(name, bases, ns)

    @property
    def __dict__(cls):
        out = {}
        for klass in reversed(cls.__mro__):
            if klass is object:
                continue
            for k, v in _RAW_DICT.__get__(klass).items():
                if not k.startswith("_"):
                    out[k] = v
        return out


class PrefixProto(metaclass=Meta):
    pass


class ParamsProto(metaclass=Meta):
    pass


class Proto:
    def __init__(self, default=None, **kw):
        self.default = default
