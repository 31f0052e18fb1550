"""Marker module: the reference's scripts start with `import isaacgym; assert isaacgym`.  The PhysX-based
simulator is not used — go1_gym in this repo drives libgo1b200.so instead."""
__all__ = []
B200_NATIVE = True
