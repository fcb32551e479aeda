"""TEST INFRASTRUCTURE ONLY -- builds (gcc) and loads oracle/cc_oracle.c, the plain-C restatement
of sam2/csrc/connected_components.cu, and exposes it with the reference's call shape
(sam2/utils/misc.py:47-63): uint8/bool [N,1,H,W] -> (labels i32, counts i32)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.cc_oracle_label.restype = ctypes.c_int
        _lib.cc_oracle_label.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def cc_label(mask: torch.Tensor):
    m = np.ascontiguousarray(mask.detach().cpu().numpy().astype(np.uint8))
    n, c, h, w = m.shape
    assert c == 1
    labels = np.zeros((n, 1, h, w), np.int32)
    counts = np.zeros((n, 1, h, w), np.int32)
    rc = _load().cc_oracle_label(m.ctypes.data, n, h, w, labels.ctypes.data, counts.ctypes.data)
    if rc != 0:
        raise RuntimeError("height and width must be even numbers" if rc == 1 else f"cc_oracle rc={rc}")
    return torch.from_numpy(labels), torch.from_numpy(counts)
