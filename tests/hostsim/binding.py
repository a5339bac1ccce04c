"""ctypes binding of the TEST-ONLY host simulator (tests/hostsim/hostsim.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libauvpf_hostsim.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "hostsim.cpp")
        hdr = os.path.join(_HERE, "..", "..", "smarc_navigation_b200", "csrc", "auvpf_math.cuh")
        if (not os.path.isfile(_SO)) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_SO)
        _lib.hs_resample.restype = C.c_int
        _lib.hs_count_below.restype = C.c_long
        _lib.hs_wrap_pi.restype = C.c_float
        _lib.hs_cdf_exact.restype = C.c_uint64
        _lib.hs_cdf_estimate.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def raycast_poses(dem, shift, range_max, m2o, poses, beams, want_mip=False, shift_hi=0, fan=True, want_swept=False):
    L = lib()
    z = np.ascontiguousarray(dem["z"], np.float32)
    H, W = z.shape
    poses = np.ascontiguousarray(poses, np.float64)
    beams = np.ascontiguousarray(beams, np.float32)
    m2o = np.ascontiguousarray(m2o, np.float64).reshape(16)
    n, B = poses.shape[0], beams.shape[0]
    out = np.empty((n, B), np.float32)
    oH, oW = C.c_int(), C.c_int()
    L.hs_mip_dims(C.c_int(H), C.c_int(W), C.c_int(shift), C.byref(oH), C.byref(oW))
    mip = np.empty((oH.value, oW.value), np.float32)
    swept = C.c_long()
    L.hs_raycast_poses(_p(z), C.c_int(H), C.c_int(W), C.c_double(dem["x0"]), C.c_double(dem["y0"]),
                       C.c_double(dem["res"]), C.c_int(shift), C.c_int(shift_hi), C.c_double(range_max), _p(m2o), _p(poses),
                       C.c_long(n), _p(beams), C.c_int(B), _p(out), _p(mip), C.c_int(int(fan)), C.byref(swept))
    if want_swept:
        return out, int(swept.value)
    return (out, mip) if want_mip else out


def propagate(poses, params, noise):
    """poses (n,6) f64 -> (n,6) f64 through the fp32/fp64 SoA device math."""
    L = lib()
    P = np.asarray(poses, np.float64)
    x, y = P[:, 0].copy(), P[:, 1].copy()
    z, r, p, w = (P[:, k].astype(np.float32) for k in (2, 3, 4, 5))
    nz = np.ascontiguousarray(noise, np.float32)
    f = [C.c_float(float(v)) for v in params]
    L.hs_propagate(_p(x), _p(y), _p(z), _p(r), _p(p), _p(w), C.c_long(len(x)), *f, _p(nz))
    return np.stack([x, y, z.astype(np.float64), r.astype(np.float64), p.astype(np.float64), w.astype(np.float64)], -1)


def philox(ctr, key):
    L = lib()
    ctr = np.ascontiguousarray(ctr, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    out = np.empty_like(ctr)
    L.hs_philox(_p(ctr), _p(key), C.c_long(ctr.shape[0]), _p(out))
    return out


def normals6(seed, ids, step, stream):
    L = lib()
    ids = np.ascontiguousarray(ids, np.int64)
    out = np.empty((len(ids), 6), np.float32)
    L.hs_normals6(C.c_uint64(seed), _p(ids), C.c_long(len(ids)), C.c_uint32(step), C.c_uint32(stream), _p(out))
    return out


def resample(logw, u):
    L = lib()
    lw = np.ascontiguousarray(logw, np.float32)
    qq = np.empty(len(lw), np.uint64)
    anc = np.empty(len(lw), np.int32)
    rc = L.hs_resample(_p(lw), C.c_long(len(lw)), C.c_double(u), _p(qq), _p(anc))
    if rc:
        raise FloatingPointError("degenerate")
    return qq, anc


def prof_reasons(reset=True):
    """Why profile marches gave up since the last reset: [steep triangle, turn-back, left the map,
    out of steps, steep anchor, anchor mismatch, centre beam missed, -]."""
    out = (C.c_long * 8)()
    lib().hs_prof_reasons(out, C.c_int(1 if reset else 0))
    return list(out)


def residual(qq, U, strict=False):
    qq = np.ascontiguousarray(qq, np.uint64)
    U = np.ascontiguousarray(U, np.float64)
    idx = np.zeros(len(qq), np.int32)
    lib().hs_residual(_p(qq), C.c_long(len(qq)), _p(U), C.c_int(1 if strict else 0), _p(idx))
    return idx


def searchsorted(arr, keys):
    arr = np.ascontiguousarray(arr, np.float64)
    keys = np.ascontiguousarray(keys, np.float64)
    out = np.zeros(len(keys), np.int64)
    lib().hs_searchsorted(_p(arr), C.c_long(len(arr)), _p(keys), C.c_long(len(keys)), _p(out))
    return out


def cdf_exact(Cv, A):
    """floor(C 2^52 / A) by the device function (float64 estimate + wrapping 128-bit multiply-back)."""
    m = 0xFFFFFFFFFFFFFFFF
    return int(lib().hs_cdf_exact(C.c_uint64(Cv & m), C.c_uint64(Cv >> 64), C.c_uint64(A & m), C.c_uint64(A >> 64)))


def cdf_estimate(Cv, A):
    m = 0xFFFFFFFFFFFFFFFF
    return float(lib().hs_cdf_estimate(C.c_uint64(Cv & m), C.c_uint64(Cv >> 64), C.c_uint64(A & m), C.c_uint64(A >> 64)))


def count_below(u, N, c):
    return int(lib().hs_count_below(C.c_double(u), C.c_long(N), C.c_uint64(c)))


def wrap_pi(a):
    return float(lib().hs_wrap_pi(C.c_float(a)))


def neff_fixed(a):
    q = np.ascontiguousarray(a, np.uint64)
    fn = lib().hs_neff_fixed
    fn.restype = C.c_double
    return float(fn(_p(q), C.c_long(len(q))))


def select_source_side(a, u, G=1, want_stats=False, n_out=None):
    """Ancestors by the host twin of k_select (chunk tables, source-side run ends from the cdf estimate,
    exact fix-up, run heads + max-scan) for ``G`` shards; ``a``: the integer weights of all G*M sources.
    ``n_out``: number of positions (default: as many as sources)."""
    a = np.ascontiguousarray(a, np.uint64)
    assert len(a) % G == 0
    n_out = len(a) if n_out is None else int(n_out)
    anc = np.zeros(n_out, np.int32)
    stats = (C.c_long * 2)()
    lib().hs_select_source_side_n(_p(a), C.c_long(len(a) // G), C.c_int(G), C.c_double(u), C.c_long(n_out), _p(anc), stats)
    return (anc, (int(stats[0]), int(stats[1]))) if want_stats else anc
