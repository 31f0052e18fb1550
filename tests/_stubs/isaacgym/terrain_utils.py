"""Stands in for isaacgym.terrain_utils when the REFERENCE's go1_gym/utils/terrain.py is imported by the golden generator:
re-exports this repository's restatement of the published generators (walk-these-ways_b200/go1_gym/utils/terrain_utils.py),
loaded by path because `go1_gym` resolves to the reference package in that process."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))),
                  "walk-these-ways_b200", "go1_gym", "utils", "terrain_utils.py")
_spec = importlib.util.spec_from_file_location("_b200_terrain_utils", _p)
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
for _k in dir(_m):
    if not _k.startswith("_"):
        globals()[_k] = getattr(_m, _k)
