"""ctypes loader of tests/hostcheck/_build/libart_hostcheck.so — the product's device headers
compiled for the CPU (test infrastructure; never imported by the package)."""
import ctypes as C
import os
import subprocess

import numpy as np

from adiabatic_raytracer_b200.params import ArtParams, CROSS_REC

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libart_hostcheck.so")
_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        L = C.CDLL(_SO)
        PP = C.POINTER(ArtParams)
        L.hc_max_radius.argtypes = [PP, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.hc_rhs.argtypes = [PP, C.c_int64, _ip, _dp, _dp, _dp, _dp]
        L.hc_condition.argtypes = [PP, C.c_int64, _dp, _dp]
        L.hc_condition_F.argtypes = [PP, C.c_int64, _dp, _dp]
        L.hc_fastmath.argtypes = [C.c_int64, _dp, _dp, _dp, _dp]
        L.hc_flog.argtypes = [C.c_int64, _dp, _dp]
        L.hc_sample_from_uniforms.argtypes = [PP, C.c_int64, _dp, _dp, _dp, _dp, _ip, _ip, _ip]
        L.hc_weights.argtypes = [PP, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.hc_conv_prob.argtypes = [PP, C.c_int64, _dp, _dp, _dp, C.c_int32, _dp]
        L.hc_propagate.argtypes = [PP, C.c_int64, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int32, C.c_int32,
                                   _dp, _dp, _dp, _ip, _ip, _dp, C.c_int32, C.c_void_p]
        L.hc_run.argtypes = [PP, C.c_uint64, C.c_int64, C.c_int64, _dp, C.c_int64, C.POINTER(C.c_int64),
                             C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise RuntimeError(f"hostcheck returned status {rc}")


def max_radius(p):
    a, b = C.c_double(), C.c_double()
    rc = lib().hc_max_radius(C.byref(p), C.byref(a), C.byref(b))
    return rc, a.value, b.value


def rhs(p, species, u, erg, lnt):
    u = np.ascontiguousarray(u, np.float64)
    du = np.zeros_like(u)
    _chk(lib().hc_rhs(C.byref(p), u.shape[0], np.ascontiguousarray(species, np.int32), u,
                      np.ascontiguousarray(erg, np.float64), np.ascontiguousarray(lnt, np.float64), du))
    return du


def fastmath(x):
    x = np.ascontiguousarray(x, np.float64)
    sn, cs, ex = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
    _chk(lib().hc_fastmath(x.size, x, sn, cs, ex))
    return sn, cs, ex


def flog(x):
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros_like(x)
    _chk(lib().hc_flog(x.size, x, out))
    return out


def condition_F(p, inp):
    inp = np.ascontiguousarray(inp, np.float64)
    out = np.zeros(inp.shape[0])
    _chk(lib().hc_condition_F(C.byref(p), inp.shape[0], inp, out))
    return out


def condition(p, inp):
    inp = np.ascontiguousarray(inp, np.float64)
    out = np.zeros(inp.shape[0])
    _chk(lib().hc_condition(C.byref(p), inp.shape[0], inp, out))
    return out


def sample_from_uniforms(p, u):
    u = np.ascontiguousarray(u, np.float64)
    n = u.shape[0]
    x, vl, vi = np.zeros((3, n)), np.zeros((3, n)), np.zeros((3, n))
    acc, nc, ncond = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    _chk(lib().hc_sample_from_uniforms(C.byref(p), n, u, x, vl, vi, acc, nc, ncond))
    return dict(x=x, v_loc=vl, v_inf=vi, accepted=acc, n_cross=nc, n_cond=ncond)


def weights(p, x, v_loc, v_inf):
    x = np.ascontiguousarray(x, np.float64)
    n = x.shape[1]
    k = np.zeros((3, n))
    cw, jg, sp, eg, ve = (np.zeros(n) for _ in range(5))
    _chk(lib().hc_weights(C.byref(p), n, x, np.ascontiguousarray(v_loc), np.ascontiguousarray(v_inf), k, cw, jg, sp, eg, ve))
    return dict(k_init=k, cos_w=cw, jac_gr=jg, sln_prob=sp, erg_inf=eg, vel_eng=ve)


def conv_prob(p, x, k, erg_scaled, b_sign=1):
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros(x.shape[1])
    _chk(lib().hc_conv_prob(C.byref(p), x.shape[1], x, np.ascontiguousarray(k, np.float64),
                            np.ascontiguousarray(erg_scaled, np.float64), b_sign, out))
    return out


def propagate(p, species, x0, k0, erg, t_start, dOmega, b_sign=1, max_crossings=-1, cross_cap=16):
    x0 = np.ascontiguousarray(x0, np.float64)
    n = x0.shape[1]
    xf, kf = np.zeros((3, n)), np.zeros((3, n))
    u7, st, nc = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
    rec = np.zeros((n, cross_cap, CROSS_REC))
    cnt = np.zeros((n, 4), np.int32)
    _chk(lib().hc_propagate(C.byref(p), n, np.ascontiguousarray(species, np.int32), x0,
                            np.ascontiguousarray(k0, np.float64), np.ascontiguousarray(erg, np.float64),
                            np.ascontiguousarray(t_start, np.float64), np.ascontiguousarray(dOmega, np.float64),
                            b_sign, max_crossings, xf, kf, u7, st, nc, rec, cross_cap, cnt.ctypes.data))
    return dict(x_f=xf, k_f=kf, u7_f=u7, status=st, n_cross=nc, cross_rec=rec, counters=cnt)


def run(p, seed, first_event, n_events):
    ncols = 31 if p.save_mode > 1 else (29 if p.save_mode > 0 else 13)
    cap = n_events * 8 + 64
    rows = np.zeros((cap, ncols))
    nr, fi, tc = C.c_int64(), C.c_int64(), C.c_int64()
    _chk(lib().hc_run(C.byref(p), seed, first_event, n_events, rows, cap, C.byref(nr), C.byref(fi), C.byref(tc)))
    return dict(rows=rows[: nr.value].copy(), f_inx=fi.value, tot_count=tc.value)


def certificates(p, species, x0, k0, erg, max_crossings=-1):
    """Certified-sign shortcuts vs the literal scan: [accepted steps, certified, contradicted, in-box, in-box contradicted]."""
    x0 = np.ascontiguousarray(x0, np.float64)
    n = x0.shape[1]
    counts = np.zeros(12, np.int64)
    _chk(lib().hc_certificates(C.byref(p), n, np.ascontiguousarray(species, np.int32).ctypes.data_as(C.c_void_p),
                               x0.ctypes.data_as(C.c_void_p), np.ascontiguousarray(k0, np.float64).ctypes.data_as(C.c_void_p),
                               np.ascontiguousarray(erg, np.float64).ctypes.data_as(C.c_void_p), max_crossings,
                               counts.ctypes.data_as(C.c_void_p)))
    return counts


def node_traj(p, species, x0, k0, erg, t_start=0.0, dOmega=-1.0, max_crossings=-1, ntimes=100):
    """Trajectory (x, y, z, ln t) of one node at the saveat points of saveMode 3."""
    out = np.zeros((ntimes + 70, 4))
    L = lib()
    L.hc_node_traj.restype = C.c_int
    n = L.hc_node_traj(C.byref(p), int(species), np.ascontiguousarray(x0, np.float64).ctypes.data_as(C.c_void_p),
                       np.ascontiguousarray(k0, np.float64).ctypes.data_as(C.c_void_p), C.c_double(erg), C.c_double(t_start),
                       C.c_double(dOmega), int(max_crossings), int(ntimes), out.ctypes.data_as(C.c_void_p), out.shape[0])
    assert n > 0
    return out[:n]


def sampler_trips(p, u):
    """How the sampler's attempts are spent: dict of pieces of work per kind, condition evaluations, and how many scanned
    steps a sharper momentum-free bound could still pass over (hc_sampler_trips)."""
    u = np.ascontiguousarray(u, np.float64)
    counts = np.zeros(12, np.int64)
    L = lib()
    L.hc_sampler_trips.restype = C.c_int
    _chk(L.hc_sampler_trips(C.byref(p), C.c_int64(u.shape[0]), u.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p)))
    return dict(scan=int(counts[0]), root=int(counts[1]), nudge=int(counts[2]), init=int(counts[3]), skip=int(counts[4]),
                n_cond=int(counts[5]), missed=int(counts[6]), scanned_dark=int(counts[7]), scanned_lit=int(counts[8]))
