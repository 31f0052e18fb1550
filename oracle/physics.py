"""ctypes front-end of the fp64 physics oracle (TEST INFRASTRUCTURE — see go1_physics_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libgo1oracle.so")


class PhysParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("gravity", C.c_double * 3),
                ("erp", C.c_double), ("cfm", C.c_double), ("max_depen_vel", C.c_double),
                ("contact_margin", C.c_double), ("bounce_threshold", C.c_double),
                ("pgs_iters", C.c_int), ("_pad", C.c_int),
                ("terrain_friction", C.c_double), ("terrain_restitution", C.c_double),
                ("pen_k", C.c_double * 4), ("pen_c", C.c_double * 4), ("pen_mt", C.c_double),
                ("limit_k", C.c_double), ("limit_c", C.c_double),
                ("hf", C.POINTER(C.c_short)), ("hf_rows", C.c_int), ("hf_cols", C.c_int),
                ("hf_hscale", C.c_double), ("hf_vscale", C.c_double), ("hf_border", C.c_double)]


class PhysState(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("quat", C.c_double * 4), ("linvel", C.c_double * 3),
                ("angvel", C.c_double * 3), ("q", C.c_double * 12), ("qd", C.c_double * 12)]


class PhysDR(C.Structure):
    _fields_ = [("friction", C.c_double), ("restitution", C.c_double), ("payload", C.c_double),
                ("com_disp", C.c_double * 3)]


class PhysOut(C.Structure):
    _fields_ = [("contact_force", (C.c_double * 3) * 17)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.go1_oracle_default_params.argtypes = [C.POINTER(PhysParams)]
        _lib.go1_oracle_substep.argtypes = [C.POINTER(PhysParams), C.POINTER(PhysDR), C.POINTER(PhysState),
                                            C.POINTER(C.c_double), C.POINTER(PhysOut)]
        _lib.go1_oracle_feet.argtypes = [C.POINTER(PhysState), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib.go1_oracle_substep_batch.argtypes = [C.POINTER(PhysParams), C.c_int, C.POINTER(PhysDR), C.POINTER(PhysState), C.POINTER(C.c_double), C.POINTER(PhysOut)]
        _lib.go1_oracle_feet_batch.argtypes = [C.c_int, C.POINTER(PhysState), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib.go1_oracle_aba.argtypes = [C.POINTER(PhysParams), C.POINTER(PhysDR), C.POINTER(PhysState),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return _lib


def set_threads(t):
    """Worker threads of the batched calls (bench.py pins them to the torch thread count: no oversubscription)."""
    lib().go1_oracle_set_threads(int(t))


def substep_batch(p, drs, states, tau):
    """All envs in one C call (OpenMP over envs). drs/states: ctypes arrays; tau: float64 [n,12]. Returns contact forces [n,17,3]."""
    n = len(states)
    L = lib()
    out = (PhysOut * n)()
    t = np.ascontiguousarray(tau, dtype=np.float64)
    L.go1_oracle_substep_batch(C.byref(p), n, drs, states, t.ctypes.data_as(C.POINTER(C.c_double)), out)
    return np.ctypeslib.as_array(out).view(np.float64).reshape(n, 17, 3).copy() if False else np.frombuffer(out, dtype=np.float64).reshape(n, 17, 3).copy()


def feet_batch(states):
    n = len(states)
    fp = np.zeros((n, 4, 3)); fv = np.zeros((n, 4, 3))
    lib().go1_oracle_feet_batch(n, states, fp.ctypes.data_as(C.POINTER(C.c_double)), fv.ctypes.data_as(C.POINTER(C.c_double)))
    return fp, fv


def state_array(n):
    return (PhysState * n)()


def dr_array(n):
    return (PhysDR * n)()


def default_params():
    p = PhysParams()
    lib().go1_oracle_default_params(C.byref(p))
    return p


def make_state(pos, quat, linvel, angvel, q, qd):
    s = PhysState()
    s.pos[:] = list(pos); s.quat[:] = list(quat); s.linvel[:] = list(linvel)
    s.angvel[:] = list(angvel); s.q[:] = list(q); s.qd[:] = list(qd)
    return s


def state_arrays(s):
    return (np.array(s.pos), np.array(s.quat), np.array(s.linvel), np.array(s.angvel), np.array(s.q), np.array(s.qd))


def make_dr(friction=1.0, restitution=0.0, payload=0.0, com_disp=(0, 0, 0)):
    d = PhysDR()
    d.friction, d.restitution, d.payload = friction, restitution, payload
    d.com_disp[:] = list(com_disp)
    return d


def substep(p, dr, s, tau):
    out = PhysOut()
    t = (C.c_double * 12)(*[float(x) for x in tau])
    lib().go1_oracle_substep(C.byref(p), C.byref(dr), C.byref(s), t, C.byref(out))
    return np.array(out.contact_force)


def feet(s):
    fp = (C.c_double * 12)(); fv = (C.c_double * 12)()
    lib().go1_oracle_feet(C.byref(s), fp, fv)
    return np.array(fp).reshape(4, 3), np.array(fv).reshape(4, 3)


def aba(p, dr, s, tau):
    a0 = (C.c_double * 6)(); qdd = (C.c_double * 12)()
    t = (C.c_double * 12)(*[float(x) for x in tau])
    lib().go1_oracle_aba(C.byref(p), C.byref(dr), C.byref(s), t, a0, qdd)
    return np.array(a0), np.array(qdd)


DEFAULT_DOF_POS = np.array([0.1, 0.8, -1.5, -0.1, 0.8, -1.5, 0.1, 1.0, -1.5, -0.1, 1.0, -1.5])  # go1_config.py:12-27
