"""ctypes prototypes for the graph / table / Gibbs entry points of include/btgpu.h."""
from __future__ import annotations

import ctypes as C

vp = C.c_void_p


def bind(L):
    pass
