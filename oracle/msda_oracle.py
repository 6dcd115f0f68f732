"""numpy/ctypes front end of the plain-C oracle (oracle/msda_oracle.c).

TEST INFRASTRUCTURE ONLY.  Nothing under mdqe_cvpr2023_b200/ may import this module; it is the
checker used by tests/, by ``__graft_entry__.smoke()`` and by bench.py's ``cpu_baseline`` /
``--impl reference`` legs.  It restates ``ms_deform_attn_core_pytorch``
(/root/reference/mdqe/models/ops/functions/ms_deform_attn_func.py:45-65) plus its autograd
gradients, and the mask einsum 'bqm,bmthw->bqthw' (mdqe/models/matcher.py:182).

Pinned against the reference's own Python by tests/test_oracle_golden.py (fixtures from
tests/golden/make_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "msda_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_OUT_DIR, "libmsda_oracle.so")
_lib = None


def build(force=False):
    """Compile the C restatement with gcc (OpenMP).  Returns the path of the shared object."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(_SRC)):
        return _LIB_PATH
    cmd = ["gcc", "-O2", "-fopenmp", "-fno-fast-math", "-shared", "-fPIC", "-o", _LIB_PATH, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, loc, aw, level_start=None):
    dt = value.dtype
    assert dt in (np.float32, np.float64), dt
    value = np.ascontiguousarray(value, dtype=dt)
    loc = np.ascontiguousarray(loc, dtype=dt)
    aw = np.ascontiguousarray(aw, dtype=dt)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    assert shapes.shape == (L, 2)
    if level_start is None:
        assert int((shapes[:, 0] * shapes[:, 1]).sum()) == S
        lsi = np.concatenate([[0], np.cumsum(shapes[:, 0] * shapes[:, 1])[:-1]]).astype(np.int64)
    else:
        # explicit starts: levels may be any sub-windows of the S value rows (temporal mode)
        lsi = np.ascontiguousarray(level_start, dtype=np.int64).reshape(L)
        assert L == 0 or int((lsi + shapes[:, 0] * shapes[:, 1]).max()) <= S
    return value, shapes, lsi, loc, aw, (N, S, M, D, L, Lq, P)


def msda_forward(value, shapes, loc, aw, level_start=None):
    """out[N,Lq,M*D] for value[N,S,M,D], shapes[L,2], loc[N,Lq,M,L,P,2], aw[N,Lq,M,L,P]."""
    value, shapes, lsi, loc, aw, dims = _prep(value, shapes, loc, aw, level_start)
    N, S, M, D, L, Lq, P = dims
    out = np.empty((N, Lq, M * D), dtype=value.dtype)
    fn = getattr(_load(), "msda_oracle_forward_f32" if value.dtype == np.float32 else "msda_oracle_forward_f64")
    fn(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(aw),
       *[ctypes.c_int(v) for v in (N, S, M, D, L, Lq, P)], _ptr(out))
    return out


def msda_backward(value, shapes, loc, aw, grad_out, level_start=None):
    """(grad_value, grad_loc, grad_aw) of sum(out * grad_out)."""
    value, shapes, lsi, loc, aw, dims = _prep(value, shapes, loc, aw, level_start)
    N, S, M, D, L, Lq, P = dims
    go = np.ascontiguousarray(grad_out, dtype=value.dtype).reshape(N, Lq, M * D)
    gv = np.zeros_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(aw)
    fn = getattr(_load(), "msda_oracle_backward_f32" if value.dtype == np.float32 else "msda_oracle_backward_f64")
    fn(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(aw), _ptr(go),
       *[ctypes.c_int(v) for v in (N, S, M, D, L, Lq, P)], _ptr(gv), _ptr(gl), _ptr(ga))
    return gv, gl, ga


def mask_forward(coeff, proto):
    """einsum('bqm,bmthw->bqthw') with double accumulation."""
    dt = coeff.dtype
    assert dt in (np.float32, np.float64)
    coeff = np.ascontiguousarray(coeff, dtype=dt)
    proto = np.ascontiguousarray(proto, dtype=dt)
    B, Q, K = coeff.shape
    assert proto.shape[:2] == (B, K)
    ncols = int(np.prod(proto.shape[2:]))
    out = np.empty((B, Q) + proto.shape[2:], dtype=dt)
    fn = getattr(_load(), "mask_oracle_forward_f32" if dt == np.float32 else "mask_oracle_forward_f64")
    fn(_ptr(coeff), _ptr(proto), ctypes.c_int(B), ctypes.c_int(Q), ctypes.c_int(K),
       ctypes.c_int64(ncols), _ptr(out))
    return out


def mask_backward(coeff, proto, grad_out):
    dt = coeff.dtype
    coeff = np.ascontiguousarray(coeff, dtype=dt)
    proto = np.ascontiguousarray(proto, dtype=dt)
    go = np.ascontiguousarray(grad_out, dtype=dt)
    B, Q, K = coeff.shape
    ncols = int(np.prod(proto.shape[2:]))
    gc = np.empty_like(coeff)
    gp = np.empty_like(proto)
    fn = getattr(_load(), "mask_oracle_backward_f32" if dt == np.float32 else "mask_oracle_backward_f64")
    fn(_ptr(coeff), _ptr(proto), _ptr(go), ctypes.c_int(B), ctypes.c_int(Q), ctypes.c_int(K),
       ctypes.c_int64(ncols), _ptr(gc), _ptr(gp))
    return gc, gp
