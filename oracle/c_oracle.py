"""ctypes wrapper of oracle/nmf_oracle.c (TEST INFRASTRUCTURE / CPU BASELINE ONLY)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libfzoracle.so")


class Geom(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int), ("C", ctypes.c_int), ("n", ctypes.c_int * 3), ("p", ctypes.c_int * 3),
                ("d", ctypes.c_int), ("S", ctypes.c_int), ("sh", (ctypes.c_int * 3) * 8)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = ctypes.CDLL(LIB)
        _lib.fzo_num_threads.restype = ctypes.c_int
    return _lib


def make_geom(x_shape, head_dim, patch, shifts) -> Geom:
    g = Geom()
    g.B, g.C, g.d, g.S = x_shape[0], x_shape[1], head_dim, len(shifts)
    for k in range(3):
        g.n[k], g.p[k] = x_shape[2 + k], patch[k]
    for s, sh in enumerate(shifts):
        for k in range(3):
            g.sh[s][k] = sh[k]
    return g


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def swnmf_forward(x, v0, head_dim, patch, shifts, relu=True, num_iters=5, eps=1e-16):
    x = np.ascontiguousarray(x, dtype=np.float32)
    v0 = np.ascontiguousarray(v0.reshape(-1), dtype=np.float32)
    y = np.empty_like(x)
    g = make_geom(x.shape, head_dim, patch, shifts)
    lib().fzo_swnmf_forward(_p(x), _p(v0), _p(y), ctypes.byref(g), int(relu), int(num_iters), ctypes.c_float(eps))
    return y


def swnmf_backward(x, gy, v0, head_dim, patch, shifts, relu=True, num_iters=5, num_grad_steps=None, eps=1e-16):
    x = np.ascontiguousarray(x, dtype=np.float32)
    gy = np.ascontiguousarray(gy, dtype=np.float32)
    v0 = np.ascontiguousarray(v0.reshape(-1), dtype=np.float32)
    gx = np.empty_like(x)
    g = make_geom(x.shape, head_dim, patch, shifts)
    k = -1 if num_grad_steps is None else int(num_grad_steps)
    lib().fzo_swnmf_backward(_p(x), _p(gy), _p(v0), _p(gx), ctypes.byref(g), int(relu), int(num_iters), k,
                             ctypes.c_float(eps))
    return gx


def num_threads() -> int:
    return int(lib().fzo_num_threads())


def use_all_cores() -> int:
    """Run the OpenMP loops on every core this process may use (launchers such as torchrun set OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().fzo_set_num_threads(int(n))
    return num_threads()
