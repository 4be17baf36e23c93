"""TEST INFRASTRUCTURE ONLY — ctypes front-end of oracle/knn_oracle.cpp (brute-force distCUDA2)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _lib():
    path = os.path.join(_HERE, "libknn_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", _HERE, "libknn_oracle.so"], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(path)
    lib.knn_oracle_dist2.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    lib.knn_oracle_dist2.restype = None
    return lib


def dist2(points) -> np.ndarray:
    pts = np.ascontiguousarray(points, np.float32)
    out = np.empty(pts.shape[0], np.float32)
    _lib().knn_oracle_dist2(pts.ctypes.data, pts.shape[0], out.ctypes.data)
    return out
