"""ctypes binding of the C ABI (include/art.h) — the same calls a Julia host makes via ccall.

There is no CPU fallback: if libart_b200.so is missing the import of this module raises, and
every compute entry point returns ART_ERR_CUDA without a device.
"""
import ctypes as C
import os

import numpy as np

from .params import ART_ERR_CAPACITY, ArtParams, ArtStats, CROSS_REC

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ART_SO", os.path.join(_HERE, "libart_b200.so"))  # ART_SO: tuning variants only

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_PP = C.POINTER(ArtParams)
_H = C.c_void_p
_i64p = C.POINTER(C.c_int64)

# every symbol include/art.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "art_version": (C.c_char_p, []),
    "art_params_default": (C.c_int, [_PP]),
    "art_max_radius": (C.c_int, [_PP, C.POINTER(C.c_double)]),
    "art_create": (C.c_int, [_PP, C.c_int, C.POINTER(_H)]),
    "art_destroy": (None, [_H]),
    "art_last_error": (C.c_char_p, [_H]),
    "art_sample_from_uniforms": (C.c_int, [_H, C.c_int64, _dp, _dp, _dp, _dp, _ip, _ip]),
    "art_sample": (C.c_int, [_H, C.c_uint64, C.c_int64, C.c_int64, _dp, _dp, _dp, _ip, _ip]),
    "art_weights": (C.c_int, [_H, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "art_conv_prob": (C.c_int, [_H, C.c_int64, _dp, _dp, _dp, C.c_int32, _dp]),
    "art_propagate": (C.c_int, [_H, C.c_int64, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int32, C.c_int32, _dp, _dp, _dp,
                                _ip, _ip, _dp, C.c_int32, C.c_void_p]),
    "art_rhs": (C.c_int, [_H, C.c_int64, _ip, _dp, _dp, _dp, _dp]),
    "art_run": (C.c_int, [_H, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, _i64p, _i64p, _i64p,
                          C.POINTER(ArtStats)]),
    "art_run_device": (C.c_int, [_H, C.c_uint64, C.c_int64, C.c_int64, _i64p, _i64p, _i64p, C.POINTER(ArtStats)]),
    "art_run_device_batches": (C.c_int, [_H, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _i64p, _i64p, _i64p,
                                         C.POINTER(ArtStats)]),
    "art_fetch_rows": (C.c_int, [_H, C.c_void_p, C.c_int64, _i64p]),
    "art_fetch_events": (C.c_int, [_H, _dp, C.c_int64, _i64p]),
    "art_write_event_files": (C.c_int, [C.c_char_p, C.c_char_p, _dp, C.c_int64, _dp, C.c_int64]),
    "art_write_tree_files": (C.c_int, [_H, C.c_char_p, C.c_char_p, _i64p]),
    "art_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "art_host_free": (C.c_int, [C.c_void_p]),
    "art_normalize_rows": (C.c_int, [_dp, C.c_int64, C.c_int32, C.c_int64]),
    "art_write_npy": (C.c_int, [C.c_char_p, _dp, C.c_int64, C.c_int32]),
    "art_npy_name": (C.c_int, [_PP, C.c_int64, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64]),
    "art_hist_config": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "art_hist_config_spectrum": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double]),
    "art_hist_len": (C.c_int64, [_H]),
    "art_hist_reset": (C.c_int, [_H]),
    "art_hist_get": (C.c_int, [_H, _dp, C.c_int64]),
    "art_hist_get_global": (C.c_int, [_H, _dp, C.c_int64]),
    "art_comm_unique_id": (C.c_int, [C.c_char_p]),
    "art_comm_init": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_char_p]),
    "art_comm_allreduce_hist": (C.c_int, [_H]),
    "art_comm_destroy": (C.c_int, [_H]),
    "art_fp64_peak": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load libart_b200.so (built by __graft_entry__.build() / adiabatic_raytracer_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -m adiabatic_raytracer_b200.build` "
                "(nvcc, sm_100a).  This package has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class ArtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"art status {code}: {msg}")
        self.code = code


class Handle:
    """Opaque per-GPU handle (art_create / art_destroy)."""

    def __init__(self, params: ArtParams, device: int = 0):
        L = load()
        self._L = L
        self.params = params
        self._h = _H()
        rc = L.art_create(C.byref(params), device, C.byref(self._h))
        if rc != 0:
            raise ArtError(rc, L.art_last_error(None).decode())
        self.ncols = 31 if params.save_mode > 1 else (29 if params.save_mode > 0 else 13)

    def close(self):
        if self._h:
            self._L.art_destroy(self._h)
            self._h = _H()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise ArtError(rc, self._L.art_last_error(self._h).decode())

    # ---- fine-grained ---------------------------------------------------------------------
    def sample_from_uniforms(self, u):
        u = np.ascontiguousarray(u, np.float64)
        n = u.shape[0]
        x, vl, vi = np.zeros((3, n)), np.zeros((3, n)), np.zeros((3, n))
        acc, nc = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._chk(self._L.art_sample_from_uniforms(self._h, n, u, x, vl, vi, acc, nc))
        return dict(x=x, v_loc=vl, v_inf=vi, accepted=acc, n_cross=nc)

    def sample(self, seed, first_event, n):
        x, vl, vi = np.zeros((3, n)), np.zeros((3, n)), np.zeros((3, n))
        nc, nf = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._chk(self._L.art_sample(self._h, seed, first_event, n, x, vl, vi, nc, nf))
        return dict(x=x, v_loc=vl, v_inf=vi, n_cross=nc, n_failed=nf)

    def weights(self, x, v_loc, v_inf):
        x = np.ascontiguousarray(x, np.float64)
        n = x.shape[1]
        k = np.zeros((3, n))
        cw, jg, sp, eg, ve = (np.zeros(n) for _ in range(5))
        self._chk(self._L.art_weights(self._h, n, x, np.ascontiguousarray(v_loc, np.float64),
                                      np.ascontiguousarray(v_inf, np.float64), k, cw, jg, sp, eg, ve))
        return dict(k_init=k, cos_w=cw, jac_gr=jg, sln_prob=sp, erg_inf=eg, vel_eng=ve)

    def conv_prob(self, x, k, erg_scaled, b_sign=1):
        x = np.ascontiguousarray(x, np.float64)
        out = np.zeros(x.shape[1])
        self._chk(self._L.art_conv_prob(self._h, x.shape[1], x, np.ascontiguousarray(k, np.float64),
                                        np.ascontiguousarray(erg_scaled, np.float64), b_sign, out))
        return out

    def rhs(self, species, u, erg, lnt):
        u = np.ascontiguousarray(u, np.float64)
        du = np.zeros_like(u)
        self._chk(self._L.art_rhs(self._h, u.shape[0], np.ascontiguousarray(species, np.int32), u,
                                  np.ascontiguousarray(erg, np.float64), np.ascontiguousarray(lnt, np.float64), du))
        return du

    def propagate(self, species, x0, k0, erg, t_start, dOmega, b_sign=1, max_crossings=-1, cross_cap=16):
        x0 = np.ascontiguousarray(x0, np.float64)
        n = x0.shape[1]
        xf, kf = np.zeros((3, n)), np.zeros((3, n))
        u7, st, nc = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
        rec = np.zeros((n, cross_cap, CROSS_REC))
        cnt = np.zeros((n, 4), np.int32)
        self._chk(self._L.art_propagate(
            self._h, n, np.ascontiguousarray(species, np.int32), x0, np.ascontiguousarray(k0, np.float64),
            np.ascontiguousarray(erg, np.float64), np.ascontiguousarray(t_start, np.float64),
            np.ascontiguousarray(dOmega, np.float64), b_sign, max_crossings, xf, kf, u7, st, nc, rec, cross_cap,
            cnt.ctypes.data))
        return dict(x_f=xf, k_f=kf, u7_f=u7, status=st, n_cross=nc, cross_rec=rec, counters=cnt)

    # ---- fused pipeline -------------------------------------------------------------------
    def run(self, seed, first_event, n_events, rows_out=None):
        """art_run: host rows out (row-major n_rows × ncols, column 7 not yet divided by f_inx).  The reference appends rows
        without bound; a buffer that turns out too small is re-allocated and the rows fetched again (no recomputation)."""
        cap = n_events * 3 + 1024
        if rows_out is None:
            rows_out = np.empty((cap, self.ncols))
        nr, fi, tc = C.c_int64(), C.c_int64(), C.c_int64()
        st = ArtStats()
        rc = self._L.art_run(self._h, seed, first_event, n_events, rows_out.ctypes.data, rows_out.shape[0],
                             C.byref(nr), C.byref(fi), C.byref(tc), C.byref(st))
        if rc == ART_ERR_CAPACITY and nr.value > rows_out.shape[0]:
            rows_out = np.empty((nr.value, self.ncols))
            rc = self._L.art_fetch_rows(self._h, rows_out.ctypes.data, rows_out.shape[0], C.byref(nr))
        self._chk(rc)
        return dict(rows=rows_out[: nr.value], f_inx=fi.value, tot_count=tc.value, stats=st.as_dict())

    def run_device_batches(self, seed, first_event, batch_events, n_batches, batch_stride=None):
        """art_run_device_batches: n_batches batches of batch_events events (ids first_event + b·batch_stride + [0, batch_events))
        as one pipelined job; rows stay on the device (fetch_rows)."""
        nr, fi, tc = C.c_int64(), C.c_int64(), C.c_int64()
        st = ArtStats()
        self._chk(self._L.art_run_device_batches(self._h, seed, first_event, batch_events, n_batches,
                                                 batch_events if batch_stride is None else batch_stride,
                                                 C.byref(nr), C.byref(fi), C.byref(tc), C.byref(st)))
        return dict(n_rows=nr.value, f_inx=fi.value, tot_count=tc.value, stats=st.as_dict())

    def run_device(self, seed, first_event, n_events):
        nr, fi, tc = C.c_int64(), C.c_int64(), C.c_int64()
        st = ArtStats()
        self._chk(self._L.art_run_device(self._h, seed, first_event, n_events, C.byref(nr), C.byref(fi), C.byref(tc),
                                         C.byref(st)))
        return dict(n_rows=nr.value, f_inx=fi.value, tot_count=tc.value, stats=st.as_dict())

    def fetch_rows(self, n_rows, rows_out=None):
        rows = np.empty((max(n_rows, 1), self.ncols)) if rows_out is None else rows_out
        nr = C.c_int64()
        self._chk(self._L.art_fetch_rows(self._h, rows.ctypes.data, rows.shape[0], C.byref(nr)))
        return rows[: nr.value]

    def fetch_events(self, n_events):
        """Event table of the last run (n_events × 12: id, v_inf[3], sln_prob, x[3], k_init[3], count)."""
        ev = np.zeros((max(n_events, 1), 12))
        n = C.c_int64()
        self._chk(self._L.art_fetch_events(self._h, ev, ev.shape[0], C.byref(n)))
        return ev[: n.value]

    def write_tree_files(self, dir_tag: str, file_tag: str) -> int:
        """saveMode 3: <dir_tag>/tree/tree_<file_tag><event> for every event of the last run; returns the number of files."""
        n = C.c_int64()
        self._chk(self._L.art_write_tree_files(self._h, dir_tag.encode(), file_tag.encode(), C.byref(n)))
        return n.value

    # ---- histogram / multi-GPU ------------------------------------------------------------
    def hist_config(self, n_costheta, n_phi, n_domega=1, domega_lo=0.0, domega_hi=0.0):
        """Bins [species][cosθf][ϕf][Δω][4]; n_domega > 1 adds the energy axis of the binned spectra."""
        self._chk(self._L.art_hist_config_spectrum(self._h, n_costheta, n_phi, n_domega, domega_lo, domega_hi))

    def hist_reset(self):
        self._chk(self._L.art_hist_reset(self._h))

    def hist_get(self):
        n = self._L.art_hist_len(self._h)
        out = np.zeros(n)
        self._chk(self._L.art_hist_get(self._h, out, n))
        return out

    def hist_get_global(self):
        """The all-reduced block of the last comm_allreduce_hist()."""
        n = self._L.art_hist_len(self._h)
        out = np.zeros(n)
        self._chk(self._L.art_hist_get_global(self._h, out, n))
        return out

    def comm_init(self, nranks, rank, uid: bytes):
        self._chk(self._L.art_comm_init(self._h, nranks, rank, uid))

    def comm_allreduce_hist(self):
        self._chk(self._L.art_comm_allreduce_hist(self._h))

    def fp64_peak(self):
        t, ms = C.c_double(), C.c_double()
        self._chk(self._L.art_fp64_peak(self._h, C.byref(t), C.byref(ms)))
        return t.value, ms.value


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = load().art_comm_unique_id(buf)
    if rc != 0:
        raise ArtError(rc, load().art_last_error(None).decode())
    return buf.raw


def max_radius(p: ArtParams):
    r = C.c_double()
    rc = load().art_max_radius(C.byref(p), C.byref(r))
    return rc, r.value


def npy_name(p: ArtParams, Ntajs: int, dir_tag: str = "results", file_tag: str = "") -> str:
    buf = C.create_string_buffer(1024)
    rc = load().art_npy_name(C.byref(p), Ntajs, dir_tag.encode(), file_tag.encode(), buf, 1024)
    if rc != 0:
        raise ArtError(rc, "art_npy_name")
    return buf.value.decode()


def write_npy(path: str, rows: np.ndarray):
    rows = np.ascontiguousarray(rows, np.float64)
    rc = load().art_write_npy(path.encode(), rows, rows.shape[0], rows.shape[1])
    if rc != 0:
        raise ArtError(rc, f"art_write_npy({path})")


def write_event_files(dir_tag: str, file_tag: str, rows: np.ndarray, events: np.ndarray):
    """<dir_tag>/event/final_<file_tag> and event_<file_tag> (saveMode > 1, MainRunner.jl:435-444)."""
    rows = np.ascontiguousarray(rows, np.float64)
    events = np.ascontiguousarray(events, np.float64)
    assert rows.shape[1] == 31 and events.shape[1] == 12
    rc = load().art_write_event_files(dir_tag.encode(), file_tag.encode(), rows, rows.shape[0], events, events.shape[0])
    if rc != 0:
        raise ArtError(rc, f"art_write_event_files({dir_tag}/event/*_{file_tag})")


def normalize_rows(rows: np.ndarray, f_inx_total: int):
    assert rows.flags["C_CONTIGUOUS"]
    rc = load().art_normalize_rows(rows, rows.shape[0], rows.shape[1], f_inx_total)
    if rc != 0:
        raise ArtError(rc, "art_normalize_rows")


class PinnedRows:
    """A (cap × ncols) float64 row buffer in page-locked host memory (art_host_alloc): `.array` is a numpy
    view to pass as `rows_out`; free with `.close()` (or use as a context manager)."""

    def __init__(self, cap: int, ncols: int):
        L = load()
        self._p = C.c_void_p()
        rc = L.art_host_alloc(C.byref(self._p), int(cap) * int(ncols) * 8)
        if rc != 0:
            raise ArtError(rc, L.art_last_error(None).decode())
        buf = (C.c_double * (int(cap) * int(ncols))).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.float64).reshape(int(cap), int(ncols))

    def close(self):
        if self._p:
            self.array = None
            load().art_host_free(self._p)
            self._p = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
