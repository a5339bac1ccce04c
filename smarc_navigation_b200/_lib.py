"""ctypes loader of ``libauvpf.so`` (C ABI in ``include/auvpf.h``).

There is no fallback: if the CUDA library is missing or fails to load, importing
anything that needs it raises.  ``__graft_entry__.build()`` (or
``make -C smarc_navigation_b200/csrc``) builds it in-tree for sm_100a.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AUVPF_LIB") or os.path.join(_HERE, "libauvpf.so")   # AUVPF_LIB: kernel-tuning builds

OK, ERR_BADARG, ERR_CUDA, ERR_STATE, ERR_DEGENERATE, ERR_NOMEM = 0, -1, -2, -3, -4, -5
SYSTEMATIC, RESIDUAL = 0, 1
RS_SORTED, RS_INPLACE_REF, RS_STRICT_REF, RS_ASYNC = 0, 1, 2, 4
IPC_BLOB_BYTES = 256
EST_REFERENCE, EST_WEIGHTED, EST_CIRCULAR_YAW, EST_FULL_COV = 0, 1, 2, 4
K_NAMES = ("propagate", "update_mbes", "update_gps", "normalise", "scan", "select_gather", "roughen",
           "estimate", "other")


class Config(C.Structure):
    """``auvpf_config`` (include/auvpf.h)."""
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("particle_count", C.c_int64),
                ("n_filters", C.c_int32), ("coarse_level", C.c_int32), ("global_offset", C.c_int64),
                ("global_count", C.c_int64), ("motion_cov", C.c_double * 6), ("meas_std", C.c_double),
                ("m2o", C.c_double * 16), ("seed", C.c_uint64), ("range_min", C.c_double),
                ("range_max", C.c_double)]


class AuvpfError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "auvpf error %d: %s" % (code, msg))
        self.code = code


class DegenerateWeights(AuvpfError):
    pass


# every symbol include/auvpf.h declares: (name, restype, argtypes)
_vp, _i32, _i64, _u64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
SYMBOLS = [
    ("auvpf_version", C.c_int, []),
    ("auvpf_last_error", C.c_char_p, []),
    ("auvpf_create", C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    ("auvpf_destroy", C.c_int, [_vp]),
    ("auvpf_sync", C.c_int, [_vp]),
    ("auvpf_stream", _vp, [_vp]),
    ("auvpf_particle_count", _i64, [_vp]),
    ("auvpf_particle_capacity", _i64, [_vp]),
    ("auvpf_kld_bins", C.c_int, [_vp, C.c_double, C.c_double, _vp]),
    ("auvpf_kld_target", _i64, [_i64, C.c_double, C.c_double, _i64, _i64]),
    ("auvpf_set_next_particle_count", C.c_int, [_vp, _i64]),
    ("auvpf_set_dem", C.c_int, [_vp, _vp, _i32, _i32, _dbl, _dbl, _dbl]),
    ("auvpf_rasterise_mesh", C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _i32, _dbl, _dbl, _dbl, C.c_float, _vp]),
    ("auvpf_set_beams", C.c_int, [_vp, _vp, _i32]),
    ("auvpf_init_particles", C.c_int, [_vp, _vp, _vp, _vp]),
    ("auvpf_set_poses", C.c_int, [_vp, _vp]),
    ("auvpf_get_poses", C.c_int, [_vp, _vp]),
    ("auvpf_get_logw", C.c_int, [_vp, _vp]),
    ("auvpf_set_logw", C.c_int, [_vp, _vp]),
    ("auvpf_predict", C.c_int, [_vp, _vp, _i32, _vp]),
    ("auvpf_predict_many", C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    ("auvpf_update_gps", C.c_int, [_vp, _vp, _vp]),
    ("auvpf_update_mbes", C.c_int, [_vp, _vp, _vp]),
    ("auvpf_mbes_flagged", C.c_int, [_vp, _vp]),
    ("auvpf_raycast", C.c_int, [_vp, _vp, _i64, _vp]),
    ("auvpf_resample", C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    ("auvpf_resample_weights", C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    ("auvpf_get_weights_q", C.c_int, [_vp, _vp]),
    ("auvpf_apply_indices", C.c_int, [_vp, _vp, _i32]),
    ("auvpf_roughen", C.c_int, [_vp, _vp, _vp]),
    ("auvpf_mean_cov", C.c_int, [_vp, _vp, _vp]),
    ("auvpf_estimate", C.c_int, [_vp, _i32, _vp, _vp]),
    ("auvpf_pose_array", C.c_int, [_vp, _i64, _i64, _vp, C.POINTER(_i64)]),
    ("auvpf_pose_array_topk", C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    ("auvpf_pose_sums_buffer", C.c_int, [_vp, C.POINTER(_vp)]),
    ("auvpf_pose_sums_dev", C.c_int, [_vp]),
    ("auvpf_resample_status", C.c_int, [_vp, _vp]),
    ("auvpf_neff", C.c_int, [_vp, _vp]),
    ("auvpf_set_neff_gate", C.c_int, [_vp, _dbl]),
    ("auvpf_neff_fixed", C.c_int, [_vp, _vp]),
    ("auvpf_resample_skipped", C.c_int, [_vp, _vp]),
    ("auvpf_replay", C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    ("auvpf_ipc_export", C.c_int, [_vp, _vp]),
    ("auvpf_ipc_import", C.c_int, [_vp, _vp, _i32, _i32]),
    ("auvpf_peer_attach", C.c_int, [_vp, C.POINTER(_vp), _i32, _i32]),
    ("auvpf_shard_resample", C.c_int, [_vp, _dbl, _vp, _vp, _vp]),
    ("auvpf_shard_degenerate", C.c_int, [_vp, C.POINTER(_i32)]),
    ("auvpf_group_resample", C.c_int, [C.POINTER(_vp), _i32, _dbl, _vp, C.POINTER(_vp)]),
    ("auvpf_shard_pose_sums", C.c_int, [_vp, _vp]),
    ("auvpf_group_pose_sums", C.c_int, [C.POINTER(_vp), _i32, _vp]),
    ("auvpf_shard_buffers", C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    ("auvpf_shard_max_dev", C.c_int, [_vp]),
    ("auvpf_shard_weights_dev", C.c_int, [_vp]),
    ("auvpf_shard_max", C.c_int, [_vp, _vp]),
    ("auvpf_shard_sum", C.c_int, [_vp, _vp, _vp]),
    ("auvpf_shard_emit", C.c_int, [_vp, _dbl, _vp, _i32, _i32, _i64, _i64, _vp, _vp]),
    ("auvpf_shard_load", C.c_int, [_vp, _vp, _i64]),
    ("auvpf_shard_load_roughen", C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    ("auvpf_profile_enable", C.c_int, [_vp, _i32]),
    ("auvpf_profile_read", C.c_int, [_vp, _vp, _vp, _i32]),
]

_lib = None


def load():
    """Load the shared library and bind every declared symbol.  Needs only the
    file and libcudart's static copy inside it -- no GPU is touched."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` or "
                          "`make -C smarc_navigation_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the ABI drifted
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc == OK:
        return
    msg = load().auvpf_last_error().decode("utf-8", "replace")
    if rc == ERR_DEGENERATE:
        raise DegenerateWeights(rc, msg)
    raise AuvpfError(rc, msg)
