"""Test stub for params_proto (absent in this image). Config classes become plain classes."""


class _Proto:
    def __init_subclass__(cls, cli=False, **kw):
        super().__init_subclass__()


class PrefixProto(_Proto):
    pass


class ParamsProto(_Proto):
    pass


class Meta(type):
    pass
