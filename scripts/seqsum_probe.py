#!/usr/bin/env python
"""Timing probe of patch_seqsum_kernel on synthetic sequences (run under ncu --metrics gpu__time_duration.sum):
launch order = the cases below."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import xreg_b200
from xreg_b200 import _lib

lib = _lib.load()
ctx = xreg_b200.Context(0)
FP = C.POINTER(C.c_float)


def run(v, serial=0):
    v = np.ascontiguousarray(v, dtype=np.float32)
    out = np.zeros(v.shape[0], dtype=np.float32)
    _lib.check(lib.xrc_seqsum_f32(ctx.handle, v.ctypes.data_as(FP), v.shape[0], v.shape[1], serial, out.ctypes.data_as(FP)))
    return out


rng = np.random.default_rng(0)
n = 206116
s = np.where(rng.random((200, n)) < 0.3, 1.0, 1.0 - rng.uniform(-0.2, 0.9, (200, n))).astype(np.float32)
cases = [("200 x C2-like", s), ("1 x C2-like", s[:1]), ("148 x C2-like", s[:148]), ("200 x ones", np.ones((200, n), np.float32)),
         ("200 x C2-like, n=51200", s[:, :51200]), ("296 x C2-like", np.concatenate([s, s[:96]]))]
for name, v in cases:
    o = run(v)
    print(name, o[:2])
