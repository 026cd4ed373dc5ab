#!/usr/bin/env python
"""Timing probe of the sequential-sum emulations on C2-like sequences (run under ncu --metrics gpu__time_duration.sum).
Launch order = for each sequence count in COUNTS: chained (serial=2), two-phase with 1 / 2 / 4 / 8 CTAs per sequence
(serial = 11 / 12 / 14 / 18), two-phase automatic (serial=0)."""
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
COUNTS = [2, 26, 50, 100, 200]
for cnt in COUNTS:
    want = run(s[:cnt], 2)
    for mode in (11, 12, 14, 18, 0):
        got = run(s[:cnt], mode)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (cnt, mode)
    print(cnt, "ok", want[:2])
