"""ctypes binding of ``oracle/c/auvpf_oracle.c`` (TEST INFRASTRUCTURE, see
``oracle/__init__.py``).  The shared object is built by ``make -C oracle``
(``__graft_entry__.build()`` does it) into ``oracle/_build/``."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libauvpf_oracle.so")
_lib = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_quantise_logw.restype = C.c_int
        _lib.orc_systematic.restype = C.c_int
    return _lib


def _dem_args(dem):
    z = np.ascontiguousarray(dem["z"], np.float32)
    return z, (z.ctypes.data_as(C.c_void_p), C.c_int(z.shape[0]), C.c_int(z.shape[1]),
               C.c_double(dem["x0"]), C.c_double(dem["y0"]), C.c_double(dem["res"]))


def raycast(dem, origins, dirs, range_max, want_diag=False):
    """Rays (n,3)+(n,3) float64 -> ranges (n,) float64 [, diag (n,3)]."""
    L = lib()
    z, da = _dem_args(dem)
    o = np.ascontiguousarray(origins, np.float64)
    d = np.ascontiguousarray(dirs, np.float64)
    n = o.shape[0]
    out = np.empty(n, np.float64)
    diag = np.empty((n, 3), np.float64) if want_diag else None
    L.orc_raycast(*da, o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.c_long(n),
                  C.c_double(range_max), out.ctypes.data_as(C.c_void_p),
                  diag.ctypes.data_as(C.c_void_p) if want_diag else None)
    return (out, diag) if want_diag else out


def update_mbes(dem, poses, m2o, beam_dirs, r_meas, meas_std, range_min, range_max,
                want_ranges=False, want_diag=False):
    """Fused raycast + beam likelihood.  Returns dict(logw, ranges?, diag?)."""
    L = lib()
    z, da = _dem_args(dem)
    P = np.ascontiguousarray(poses, np.float64)
    N = P.shape[0]
    bd = np.ascontiguousarray(beam_dirs, np.float64)
    B = bd.shape[0]
    rm = None if r_meas is None else np.ascontiguousarray(r_meas, np.float64)
    m = np.ascontiguousarray(m2o, np.float64).reshape(16)
    logw = np.empty(N, np.float64)
    ranges = np.empty((N, B), np.float64) if want_ranges else None
    diag = np.empty((N, B, 3), np.float64) if want_diag else None
    L.orc_update_mbes(*da, P.ctypes.data_as(C.c_void_p), C.c_long(N), m.ctypes.data_as(C.c_void_p),
                      bd.ctypes.data_as(C.c_void_p),
                      rm.ctypes.data_as(C.c_void_p) if rm is not None else None, C.c_int(B),
                      C.c_double(meas_std), C.c_double(range_min), C.c_double(range_max),
                      ranges.ctypes.data_as(C.c_void_p) if want_ranges else None,
                      diag.ctypes.data_as(C.c_void_p) if want_diag else None,
                      logw.ctypes.data_as(C.c_void_p))
    out = {"logw": logw}
    if want_ranges:
        out["ranges"] = ranges
    if want_diag:
        out["diag"] = diag
    return out


def motion_pred(poses, v, wz, roll, pitch, z, dt, cov, noise):
    L = lib()
    P = np.array(poses, np.float64, copy=True, order="C")
    nz = np.ascontiguousarray(noise, np.float64)
    v3 = (C.c_double * 3)(*[float(x) for x in v])
    c6 = (C.c_double * 6)(*[float(x) for x in cov])
    L.orc_motion_pred(P.ctypes.data_as(C.c_void_p), C.c_long(P.shape[0]), v3, C.c_double(wz),
                      C.c_double(roll), C.c_double(pitch), C.c_double(z), C.c_double(dt), c6,
                      nz.ctypes.data_as(C.c_void_p))
    return P


def quantise_logw(logw32):
    """-> (qq uint64[N] summing to 2^52, m float32, A int)."""
    L = lib()
    lw = np.ascontiguousarray(logw32, np.float32)
    qq = np.empty(lw.shape[0], np.uint64)
    m = C.c_float()
    A = (C.c_uint64 * 2)()
    rc = L.orc_quantise_logw(lw.ctypes.data_as(C.c_void_p), C.c_long(lw.shape[0]),
                             qq.ctypes.data_as(C.c_void_p), C.byref(m), A)
    if rc != 0:
        raise FloatingPointError("degenerate weights")
    return qq, np.float32(m.value), int(A[0]) | (int(A[1]) << 64)


def integer_weights(logw32, m):
    """a_i = floor(exp(logw_i - m) 2^60) for a GIVEN maximum m (a shard of a sharded filter)."""
    L = lib()
    lw = np.ascontiguousarray(logw32, np.float32)
    a = np.empty(lw.shape[0], np.uint64)
    L.orc_integer_weights(lw.ctypes.data_as(C.c_void_p), C.c_long(lw.shape[0]), C.c_float(float(m)),
                          a.ctypes.data_as(C.c_void_p))
    return a


def systematic(w, u):
    L = lib()
    w = np.ascontiguousarray(w, np.float64)
    idx = np.empty(w.shape[0], np.int32)
    rc = L.orc_systematic(w.ctypes.data_as(C.c_void_p), C.c_long(w.shape[0]), C.c_double(u),
                          idx.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise IndexError("reference would raise IndexError (cumsum[-1] <= positions[-1])")
    return idx


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(int(n)))
