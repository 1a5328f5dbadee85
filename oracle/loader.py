"""ctypes loader of the CPU oracle (oracle/_build/libart_oracle.so).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never from adiabatic_raytracer_b200/.
Parity status of the oracle: UNPINNED against the real Julia reference (see oracle/README.md).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from adiabatic_raytracer_b200.params import ArtParams, ArtStats, CROSS_REC

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libart_oracle.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("art_oracle.cpp", "physics.hpp", "tracer.hpp", "dual.hpp", "philox.hpp", "vern6.hpp")]
    srcs.append(os.path.join(_HERE, "..", "include", "art.h"))
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        PP = C.POINTER(ArtParams)
        L.orc_max_radius.argtypes = [PP, C.POINTER(C.c_double)]
        L.orc_pointwise.argtypes = [PP, C.c_int, C.c_int64, _dp, _dp]
        L.orc_rhs.argtypes = [PP, C.c_int64, _ip, _dp, _dp, _dp, C.c_int32, _dp]
        L.orc_sample_from_uniforms.argtypes = [PP, C.c_int64, _dp, _dp, _dp, _dp, _ip, _ip]
        L.orc_sample.argtypes = [PP, C.c_uint64, C.c_int64, C.c_int64, _dp, _dp, _dp, _ip, _ip]
        L.orc_weights.argtypes = [PP, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_conv_prob.argtypes = [PP, C.c_int64, _dp, _dp, _dp, C.c_int32, _dp]
        L.orc_propagate.argtypes = [PP, C.c_int64, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int32, C.c_int32,
                                    _dp, _dp, _dp, _ip, _ip, _dp, C.c_int32, C.c_void_p]
        L.orc_run.argtypes = [PP, C.c_uint64, C.c_int64, C.c_int64, _dp, C.c_int64, C.POINTER(C.c_int64),
                              C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(ArtStats)]
        L.orc_run_events.argtypes = L.orc_run.argtypes + [_dp]
        L.orc_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sampler_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, _dp]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise RuntimeError(f"oracle returned status {rc}")


def max_radius(p):
    r = C.c_double()
    rc = lib().orc_max_radius(C.byref(p), C.byref(r))
    return rc, r.value


def pointwise(p, which, inp, n_out=1):
    inp = np.ascontiguousarray(inp, dtype=np.float64)
    n = inp.shape[0]
    out = np.zeros((n, n_out) if n_out > 1 else n)
    _chk(lib().orc_pointwise(C.byref(p), which, n, inp, out))
    return out


def rhs(p, species, u, erg, lnt, b_sign=1):
    u = np.ascontiguousarray(u, dtype=np.float64)
    n = u.shape[0]
    du = np.zeros_like(u)
    _chk(lib().orc_rhs(C.byref(p), n, np.ascontiguousarray(species, np.int32), u,
                       np.ascontiguousarray(erg, np.float64), np.ascontiguousarray(lnt, np.float64), b_sign, du))
    return du


def sample_from_uniforms(p, u):
    u = np.ascontiguousarray(u, dtype=np.float64)
    n = u.shape[0]
    x, vl, vi = np.zeros((3, n)), np.zeros((3, n)), np.zeros((3, n))
    acc, nc = np.zeros(n, np.int32), np.zeros(n, np.int32)
    _chk(lib().orc_sample_from_uniforms(C.byref(p), n, u, x, vl, vi, acc, nc))
    return dict(x=x, v_loc=vl, v_inf=vi, accepted=acc, n_cross=nc)


def sample(p, seed, first_event, n):
    x, vl, vi = np.zeros((3, n)), np.zeros((3, n)), np.zeros((3, n))
    nc, nf = np.zeros(n, np.int32), np.zeros(n, np.int32)
    _chk(lib().orc_sample(C.byref(p), seed, first_event, n, x, vl, vi, nc, nf))
    return dict(x=x, v_loc=vl, v_inf=vi, n_cross=nc, n_failed=nf)


def weights(p, x, v_loc, v_inf):
    x = np.ascontiguousarray(x, np.float64)
    n = x.shape[1]
    k = np.zeros((3, n))
    cw, jg, sp, eg, ve = (np.zeros(n) for _ in range(5))
    _chk(lib().orc_weights(C.byref(p), n, x, np.ascontiguousarray(v_loc), np.ascontiguousarray(v_inf), k, cw, jg, sp, eg, ve))
    return dict(k_init=k, cos_w=cw, jac_gr=jg, sln_prob=sp, erg_inf=eg, vel_eng=ve)


def conv_prob(p, x, k, erg_scaled, b_sign=1):
    x = np.ascontiguousarray(x, np.float64)
    n = x.shape[1]
    out = np.zeros(n)
    _chk(lib().orc_conv_prob(C.byref(p), n, x, np.ascontiguousarray(k, np.float64),
                             np.ascontiguousarray(erg_scaled, np.float64), b_sign, out))
    return out


def propagate(p, species, x0, k0, erg, t_start, dOmega, b_sign=1, max_crossings=-1, cross_cap=16):
    x0 = np.ascontiguousarray(x0, np.float64)
    n = x0.shape[1]
    xf, kf = np.zeros((3, n)), np.zeros((3, n))
    u7, st, nc = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
    rec = np.zeros((n, cross_cap, CROSS_REC))
    cnt = np.zeros((n, 4), np.int32)
    _chk(lib().orc_propagate(C.byref(p), n, np.ascontiguousarray(species, np.int32), x0,
                             np.ascontiguousarray(k0, np.float64), np.ascontiguousarray(erg, np.float64),
                             np.ascontiguousarray(t_start, np.float64), np.ascontiguousarray(dOmega, np.float64),
                             b_sign, max_crossings, xf, kf, u7, st, nc, rec, cross_cap, cnt.ctypes.data))
    return dict(x_f=xf, k_f=kf, u7_f=u7, status=st, n_cross=nc, cross_rec=rec, counters=cnt)


def run(p, seed, first_event, n_events, rows_cap=None):
    ncols = 31 if p.save_mode > 1 else (29 if p.save_mode > 0 else 13)
    cap = rows_cap or (n_events * 8 + 64)
    rows = np.zeros((cap, ncols))
    events = np.zeros((max(n_events, 1), 12))  # ART_EVENT_COLS
    nr, fi, tc = C.c_int64(), C.c_int64(), C.c_int64()
    st = ArtStats()
    _chk(lib().orc_run_events(C.byref(p), seed, first_event, n_events, rows, cap, C.byref(nr), C.byref(fi), C.byref(tc),
                              C.byref(st), events))
    return dict(rows=rows[: nr.value].copy(), f_inx=fi.value, tot_count=tc.value, stats=st.as_dict(), events=events[:n_events])


def philox(ctr, key):
    ctr = np.ascontiguousarray(ctr, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().orc_philox(ctr.ctypes.data, key.ctypes.data, out.ctypes.data)
    return out


def sampler_uniforms(seed, event, attempt):
    U = np.zeros(10)
    lib().orc_sampler_uniforms(seed, event, attempt, U)
    return U


def set_literal_start_events(on):
    """True: locate-then-discard start-point events like the reference; False (default): start-box rule."""
    lib().orc_set_literal_start_events(int(bool(on)))


def num_threads():
    return lib().orc_num_threads()
