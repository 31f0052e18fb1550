"""go1_gym — drop-in namespace of the reference's env package, driving libgo1b200.so instead of Isaac Gym."""
import os

MINI_GYM_ROOT_DIR = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
MINI_GYM_ENVS_DIR = os.path.join(MINI_GYM_ROOT_DIR, "go1_gym", "envs")
