"""Python facade: the reference filter API over the CUDA library.

Mirrors ``auv_particle_filter/scripts/auv_pf.py`` (class ``auv_pf``): the same
constructor parameters (the ROS params of ``auv_pf.py:27-56`` /
``auv_pf.launch:15-20``, same names, same "list of VARIANCES in
[x, y, z, roll, pitch, yaw] order" meaning, covariance strings accepted) and the
same three methods ``predict(odom_t)``, ``update(meas) -> weights``,
``resample(weights)``, plus ``update_loc_pose``'s numbers (``mean_cov``).  A ROS
node keeps its callbacks and swaps the object.

All arithmetic happens in ``libauvpf.so`` on the GPU; this file only marshals
arguments.  There is no CPU path.
"""
import ctypes as C
import math
import threading

import numpy as np

from . import _lib

_TWO_PI = 2.0 * math.pi


def parse_covariance(value):
    """The reference's covariance parameters arrive as strings like
    ``"[0.1, 0.1, 0.0, 0.0, 0.0, 0.0]"`` (auv_pf.py:40-56); lists pass through."""
    if isinstance(value, str):
        value = value.replace('[', '').replace(']', '')
        value = list(map(float, value.split(",")))
    v = [float(x) for x in value]
    if len(v) != 6:
        raise ValueError("covariance must have 6 entries [x, y, z, roll, pitch, yaw]")
    return v


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _odom_row(odom_t, dt):
    """Accept a ROS ``nav_msgs/Odometry``-shaped object (the attributes
    ``Particle.motion_pred`` reads, auv_particle.py:43-70) or a plain object /
    tuple ``(v[3], wz, quat_xyzw[4], z)``."""
    if hasattr(odom_t, "twist"):
        tw, po = odom_t.twist.twist, odom_t.pose.pose
        return [tw.linear.x, tw.linear.y, tw.linear.z, tw.angular.z, po.orientation.x, po.orientation.y,
                po.orientation.z, po.orientation.w, po.position.z, dt, 0.0]
    if hasattr(odom_t, "v"):
        v, wz, q, z = odom_t.v, odom_t.wz, odom_t.quat, odom_t.z
        if dt is None:
            dt = odom_t.dt
    else:
        v, wz, q, z = odom_t
    return [v[0], v[1], v[2], wz, q[0], q[1], q[2], q[3], z, dt, 0.0]


class ParticleFilter(object):
    """Drop-in for the filter half of ``auv_pf`` (auv_pf.py:23-260).

    Parameters (reference names, auv_pf.py:27-56): ``particle_count``,
    ``init_covariance``, ``motion_covariance``, ``resampling_noise_covariance``,
    ``measurement_std``, ``m2o_matrix`` (4x4 map<-odom, auv_pf.py:79-81).
    GPU-side additions: ``n_filters``, ``seed``, ``device``, ``range_min``,
    ``range_max``, ``resampler`` ("systematic", the north-star default, or
    "residual", the one the reference calls, auv_pf.py:182), ``inplace_reference``
    (keep the reference's survivors-stay-in-slot map, auv_pf.py:183-198),
    ``strict_reference`` (residual only: reproduce resampling.py:70-74 literally,
    particle_count <= 65536).  ``resampler="residual", inplace_reference=True,
    strict_reference=True`` is the reference's ``resample()`` bit for bit.
    ``resample_neff_ratio``: the gate the reference carries commented out
    (auv_pf.py:174-181, ``if N_eff < self.pc/2.``): 0 (default) resamples on every call like
    the live reference code, 0.5 is the commented value; see ``set_neff_gate``.
    """

    def __init__(self, particle_count=10, init_covariance=(0.,) * 6, motion_covariance=(0.,) * 6,
                 resampling_noise_covariance=(0.,) * 6, measurement_std=0.01, m2o_matrix=None,
                 n_filters=1, seed=0, device=0, range_min=0.0, range_max=120.0, coarse_level=-1,
                 resampler="systematic", inplace_reference=False, strict_reference=False, init_mean=None,
                 init_noise=None,
                 global_offset=0, global_count=0, resample_neff_ratio=0.0):
        self._lib = _lib.load()
        self.pc = int(particle_count)
        self.n_filters = int(n_filters)
        self.init_cov = parse_covariance(init_covariance)
        self.motion_cov = parse_covariance(motion_covariance)
        self.res_noise_cov = parse_covariance(resampling_noise_covariance)
        self.meas_std = float(measurement_std)
        self.m2o_mat = np.identity(4) if m2o_matrix is None else np.asarray(m2o_matrix, np.float64).reshape(4, 4)
        self.resampler = {"systematic": _lib.SYSTEMATIC, "residual": _lib.RESIDUAL}[resampler]
        self.inplace_reference = bool(inplace_reference)
        self.strict_reference = bool(strict_reference)
        self._lock = threading.Lock()      # the reference's callbacks race on self.particles
        self.time = None
        self.old_time = None
        cfg = _lib.Config()
        cfg.struct_size = C.sizeof(_lib.Config)
        cfg.device = int(device)
        cfg.particle_count = self.pc
        cfg.n_filters = self.n_filters
        if isinstance(coarse_level, (tuple, list)):          # (finer, coarser) max-mip levels
            coarse_level = int(coarse_level[0]) | (int(coarse_level[1]) << 8)
        cfg.coarse_level = int(coarse_level)
        cfg.global_offset = int(global_offset)
        cfg.global_count = int(global_count)
        cfg.motion_cov[:] = self.motion_cov
        cfg.meas_std = self.meas_std
        cfg.m2o[:] = self.m2o_mat.reshape(16).tolist()
        cfg.seed = int(seed)
        cfg.range_min = float(range_min)
        cfg.range_max = float(range_max)
        self._h = C.c_void_p()
        _lib.check(self._lib.auvpf_create(C.byref(cfg), C.byref(self._h)))
        self.n_rows = self.pc * self.n_filters
        self.n_beams = 0
        self.resample_neff_ratio = 0.0
        if resample_neff_ratio:
            self.set_neff_gate(resample_neff_ratio)
        # Particle.__init__: zero pose + add_noise(init_cov) (auv_particle.py:24,30)
        mean = np.zeros(6) if init_mean is None else np.asarray(init_mean, np.float64)
        self.init_particles(mean, self.init_cov, init_noise)

    # -- lifecycle -----------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.auvpf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _lib.check(self._lib.auvpf_sync(self._h))

    @property
    def stream_ptr(self):
        return self._lib.auvpf_stream(self._h)

    # -- map / sensor --------------------------------------------------------------------
    def set_dem(self, z, x0=0.0, y0=0.0, res=1.0):
        if isinstance(z, dict):
            z, x0, y0, res = z["z"], z["x0"], z["y0"], z["res"]
        z = np.ascontiguousarray(z, np.float32)
        with self._lock:
            _lib.check(self._lib.auvpf_set_dem(self._h, _ptr(z), z.shape[0], z.shape[1], x0, y0, res))

    def load_dem(self, path):
        """Read a DEM file (``dem_io``: float32 or uint16-quantised heights + georeference) and make it the map.
        Returns the dict that was loaded -- the float32 heights in it are exactly what the kernels see."""
        from . import dem_io
        dem = dem_io.load_dem(path)
        self.set_dem(dem)
        return dem

    def rasterise_mesh(self, vertices, faces, shape, x0=0.0, y0=0.0, res=1.0, fill=np.nan):
        """Triangle mesh -> (H, W) float32 DEM on the GPU (highest surface wins; uncovered vertices
        get ``fill``); feed the result to ``set_dem``."""
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        f = np.ascontiguousarray(faces, np.int32).reshape(-1, 3)
        if f.size and (f.min() < 0 or f.max() >= v.shape[0]):
            raise ValueError("face index out of range")
        z = np.empty((int(shape[0]), int(shape[1])), np.float32)
        with self._lock:
            _lib.check(self._lib.auvpf_rasterise_mesh(self._h, _ptr(v), v.shape[0], _ptr(f), f.shape[0], z.shape[0],
                                                      z.shape[1], x0, y0, res, fill, _ptr(z)))
        return z

    def set_beams(self, dirs_xyz):
        d = np.ascontiguousarray(dirs_xyz, np.float32).reshape(-1, 3)
        with self._lock:
            _lib.check(self._lib.auvpf_set_beams(self._h, _ptr(d), d.shape[0]))
        self.n_beams = d.shape[0]

    def set_scan_geometry(self, angle_min, angle_increment, n):
        """Beam table of a ``sensor_msgs/LaserScan``-style fan in the sensor y-z plane
        (the only MBES wire format in the reference,
        mbes_processors/mbes_mapper/include/mbes_mapper/mbes_receptor.hpp:26,82)."""
        th = angle_min + angle_increment * np.arange(n)
        self.set_beams(np.stack([np.zeros(n), np.sin(th), -np.cos(th)], -1))

    # -- state ---------------------------------------------------------------------------
    def init_particles(self, mean, var, noise=None):
        m = (C.c_double * 6)(*[float(x) for x in mean])
        v = (C.c_double * 6)(*parse_covariance(var))
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32).reshape(self.n_rows, 6)
        with self._lock:
            _lib.check(self._lib.auvpf_init_particles(self._h, m, v, _ptr(nz)))

    def poses(self):
        """(F*M, 6) float64 ``[x, y, z, roll, pitch, yaw]`` = every ``particles[i].p_pose``."""
        out = np.empty((self.n_rows, 6), np.float64)
        with self._lock:
            _lib.check(self._lib.auvpf_get_poses(self._h, _ptr(out)))
        return out

    def set_poses(self, poses):
        p = np.ascontiguousarray(poses, np.float64).reshape(self.n_rows, 6)
        with self._lock:
            _lib.check(self._lib.auvpf_set_poses(self._h, _ptr(p)))

    def log_weights(self):
        out = np.empty(self.n_rows, np.float32)
        with self._lock:
            _lib.check(self._lib.auvpf_get_logw(self._h, _ptr(out)))
        return out

    def set_log_weights(self, logw):
        lw = np.ascontiguousarray(logw, np.float32).reshape(self.n_rows)
        with self._lock:
            _lib.check(self._lib.auvpf_set_logw(self._h, _ptr(lw)))
            _lib.check(self._lib.auvpf_sync(self._h))

    # -- predict: auv_pf.predict (auv_pf.py:213-216) -----------------------------------------
    def predict(self, odom_t, dt=None, noise=None):
        """``dt`` defaults to ``self.time - self.old_time`` like the reference
        (:214) when the caller maintains those, else must be given (or carried by
        the odom object).  ``odom_t`` may be a list of F odoms (one per filter), or an ``(F, 11)``
        float64 array of ready-made ABI rows ``[vx vy vz wz qx qy qz qw z dt 0]``."""
        if dt is None and self.time is not None and self.old_time is not None:
            dt = self.time - self.old_time
        if isinstance(odom_t, np.ndarray) and odom_t.ndim == 2 and odom_t.shape[1] == 11:
            od = np.ascontiguousarray(odom_t, np.float64)         # ready-made (F, 11) / (1, 11) rows of the C ABI
            if od.shape[0] not in (1, self.n_filters):
                raise ValueError("odometry rows: 1 or n_filters")
        else:
            many = isinstance(odom_t, (list, tuple)) and len(odom_t) == self.n_filters and self.n_filters > 1 \
                and not np.isscalar(odom_t[1])
            rows = [_odom_row(o, dt) for o in odom_t] if many else [_odom_row(odom_t, dt)]
            if any(r[9] is None for r in rows):
                raise ValueError("predict needs dt")
            od = np.ascontiguousarray(rows, np.float64)
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32).reshape(self.n_rows, 6)
        with self._lock:
            _lib.check(self._lib.auvpf_predict(self._h, _ptr(od), od.shape[0], _ptr(nz)))

    def predict_many(self, odoms, dt=None, noise=None):
        """Several consecutive ``predict`` calls (the odometry messages queued between two pings) in
        ONE pass over the particle state; bit-identical to calling ``predict`` for each.  ``odoms``:
        a list of what ``predict`` accepts, or an ``(n_steps, F or 1, 11)`` array of ABI rows."""
        if isinstance(odoms, np.ndarray) and odoms.ndim == 3 and odoms.shape[2] == 11:
            od = np.ascontiguousarray(odoms, np.float64)
        else:
            steps = []
            for o in odoms:
                if isinstance(o, np.ndarray) and o.ndim == 2 and o.shape[1] == 11:
                    steps.append(np.asarray(o, np.float64))
                else:
                    many = isinstance(o, (list, tuple)) and len(o) == self.n_filters and self.n_filters > 1 \
                        and not np.isscalar(o[1])
                    rows = [_odom_row(x, dt) for x in o] if many else [_odom_row(o, dt)]
                    if any(r[9] is None for r in rows):
                        raise ValueError("predict needs dt")
                    steps.append(np.asarray(rows, np.float64))
            od = np.ascontiguousarray(np.stack(steps), np.float64)
        if od.shape[1] not in (1, self.n_filters):
            raise ValueError("odometry rows per step: 1 or n_filters")
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32).reshape(od.shape[0], self.n_rows, 6)
        with self._lock:
            _lib.check(self._lib.auvpf_predict_many(self._h, _ptr(od), od.shape[0], od.shape[1], _ptr(nz)))

    # -- update: auv_pf.update (auv_pf.py:135-167) -------------------------------------------
    def update(self, meas, return_weights=True):
        """GPS form (what the reference does): ``meas`` is a ROS Odometry-shaped GPS
        fix (``pose.pose.position.{x,y}``, already in the map frame -- the
        utm->map ``transformPoint`` of :151 stays in the node) or an ``(x, y)``
        pair.  MBES form: ``meas`` is an array of B ranges (or (F, B)), or a
        LaserScan-shaped object with ``.ranges``.  Returns the reference's
        ``weights`` array (linear float64, ``+ 1e-200``, :165) unless
        ``return_weights=False`` (weights stay on the GPU for ``resample()``)."""
        if hasattr(meas, "ranges"):
            return self.update_mbes(np.asarray(meas.ranges, np.float32), return_weights)
        if hasattr(meas, "pose"):
            p = meas.pose.pose.position
            return self.update_gps((p.x, p.y), return_weights)
        a = np.asarray(meas)
        if a.size == 2 * self.n_filters and self.n_beams != 2 * self.n_filters:
            return self.update_gps(a, return_weights)
        return self.update_mbes(a, return_weights)

    def _weights_out(self, logw):
        return np.exp(logw.astype(np.float64)) + 1.0e-200          # auv_pf.py:165

    def update_gps(self, gps_xy, return_weights=True):
        g = np.ascontiguousarray(np.broadcast_to(np.asarray(gps_xy, np.float64).reshape(-1, 2),
                                                 (self.n_filters, 2)))
        out = np.empty(self.n_rows, np.float32) if return_weights else None
        with self._lock:
            _lib.check(self._lib.auvpf_update_gps(self._h, _ptr(g), _ptr(out)))
        return self._weights_out(out) if return_weights else None

    def update_mbes(self, ranges, return_weights=True):
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(ranges, np.float32).reshape(-1, self.n_beams),
                                                 (self.n_filters, self.n_beams)))
        out = np.empty(self.n_rows, np.float32) if return_weights else None
        with self._lock:
            _lib.check(self._lib.auvpf_update_mbes(self._h, _ptr(r), _ptr(out)))
        return self._weights_out(out) if return_weights else None

    def mbes_flagged(self):
        """Per filter ``[second-pass (map-border) particles, particles recast ray by ray]`` of the last
        ``update_mbes`` (synchronises)."""
        out = np.zeros((self.n_filters, 2), np.int64)
        with self._lock:
            _lib.check(self._lib.auvpf_mbes_flagged(self._h, _ptr(out)))
        return out

    def raycast(self, poses):
        """Expected ranges (n, B) float32 for explicit poses (parity probe)."""
        p = np.ascontiguousarray(poses, np.float64).reshape(-1, 6)
        out = np.empty((p.shape[0], self.n_beams), np.float32)
        with self._lock:
            _lib.check(self._lib.auvpf_raycast(self._h, _ptr(p), p.shape[0], _ptr(out)))
        return out

    # -- resample: auv_pf.resample (auv_pf.py:169-192) ---------------------------------------
    def resample(self, weights=None, u=None, noise=None, return_ancestors=False, roughen=True, wait=True, new_count=None):
        """Normalise (:172), draw ancestors, reassign poses (:183-189,195-198) and add
        ``resampling_noise_covariance`` noise to every particle (:191-192).
        ``weights=None`` uses the device-resident log-weights of the last update;
        a float64 array reproduces the reference call ``resample(weights)``.
        ``u``: the ``random()`` of resampling.py:157 (one per filter) for the systematic
        resampler, the ``random(N - k)`` array of :74 (``(F, M)``; the first ``M - k`` of a row
        are consumed) for the residual one; default Philox.
        ``new_count`` (one filter, systematic, not in place): draw that many particles instead of ``particle_count``
        -- positions ``(u + i) / new_count`` over the CDF of the current ones -- and continue with them (KLD
        adaptation, ``kld_particle_count``); between 1 and the count the filter was created with.  A resample held
        back by the N_eff gate leaves the count alone."""
        if new_count is not None:
            with self._lock:
                _lib.check(self._lib.auvpf_set_next_particle_count(self._h, int(new_count)))
        anc = np.empty(max(self.n_rows, int(new_count or 0)), np.int32) if return_ancestors else None
        if u is None:
            uu = None
        elif self.resampler == _lib.RESIDUAL:
            uu = np.zeros((self.n_filters, self.pc), np.float64)
            ua = np.asarray(u, np.float64).reshape(self.n_filters, -1)
            uu[:, :min(ua.shape[1], self.pc)] = ua[:, :self.pc]
        else:
            uu = np.ascontiguousarray(np.broadcast_to(np.asarray(u, np.float64).reshape(-1), (self.n_filters,)))
        rv = (C.c_double * 6)(*self.res_noise_cov) if roughen else None
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32).reshape(self.n_rows, 6)
        flags = _lib.RS_INPLACE_REF if self.inplace_reference else _lib.RS_SORTED
        if self.strict_reference and self.resampler == _lib.RESIDUAL:
            flags |= _lib.RS_STRICT_REF
        if not wait:          # enqueue only; degenerate filters keep their state, see resample_status()
            flags |= _lib.RS_ASYNC
        with self._lock:
            if weights is None:
                rc = self._lib.auvpf_resample(self._h, self.resampler, flags, _ptr(uu), _ptr(anc), rv, _ptr(nz))
            else:
                w = np.ascontiguousarray(weights, np.float64).reshape(self.n_rows)
                rc = self._lib.auvpf_resample_weights(self._h, _ptr(w), self.resampler, flags, _ptr(uu),
                                                      _ptr(anc), rv, _ptr(nz))
            _lib.check(rc)
            if new_count is not None:
                self.n_rows = int(self._lib.auvpf_particle_count(self._h))
                self.pc = self.n_rows // self.n_filters
                self._lib.auvpf_set_next_particle_count(self._h, 0)        # (a held-back resample: no stale request left behind)
        return anc if anc is None else anc[:self.n_rows]

    # -- KLD-style adaptation of the particle count (not in the reference) ------------------------
    @property
    def capacity(self):
        """The particle count the filter was created with: the upper bound of ``resample(new_count=...)``."""
        return int(self._lib.auvpf_particle_capacity(self._h)) // self.n_filters

    def kld_bins(self, bin_xy=0.5, bin_yaw=np.radians(5.0)):
        """Occupied bins of size ``bin_xy`` x ``bin_xy`` [m] x ``bin_yaw`` [rad] per filter (counted on the device)."""
        out = np.zeros(self.n_filters, np.int64)
        with self._lock:
            _lib.check(self._lib.auvpf_kld_bins(self._h, float(bin_xy), float(bin_yaw), _ptr(out)))
        return out

    def kld_particle_count(self, bin_xy=0.5, bin_yaw=np.radians(5.0), eps=0.05, z_quantile=2.3263478740408408,
                           n_min=1000, n_max=None):
        """Fox's KLD bound for the current cloud: the particle count that keeps the KL divergence of the sample-based
        posterior below ``eps`` with the confidence of ``z_quantile`` (default 99 %), clamped to ``[n_min, n_max]``
        (default: the capacity).  Feed it to ``resample(new_count=...)``."""
        n_max = self.capacity if n_max is None else int(n_max)
        k = self.kld_bins(bin_xy, bin_yaw)
        return np.array([self._lib.auvpf_kld_target(int(b), float(eps), float(z_quantile), int(n_min), n_max) for b in k], np.int64)

    def resample_status(self):
        """Per-filter degenerate-weights flag of the last resample (synchronises)."""
        out = np.zeros(self.n_filters, np.int32)
        with self._lock:
            _lib.check(self._lib.auvpf_resample_status(self._h, _ptr(out)))
        return out.astype(bool)

    def weights_q(self):
        """Fixed-point normalised weights of the last resample: uint64 q, w = q * 2^-52."""
        out = np.empty(self.n_rows, np.uint64)
        with self._lock:
            _lib.check(self._lib.auvpf_get_weights_q(self._h, _ptr(out)))
        return out

    def apply_indices(self, indices):
        """The reference's keep/lost/dupes reassignment for given indices (auv_pf.py:183-198)."""
        idx = np.ascontiguousarray(indices, np.int32).reshape(self.n_rows)
        with self._lock:
            _lib.check(self._lib.auvpf_apply_indices(self._h, _ptr(idx), 0))

    def reassign_poses(self, lost, dupes):
        """auv_pf.reassign_poses (auv_pf.py:195-198) for explicit lists."""
        idx = np.arange(self.n_rows, dtype=np.int32)
        n = min(len(lost), len(dupes))
        idx[np.asarray(lost[:n], np.int64)] = np.asarray(dupes[:n], np.int32)
        # an index array in which every slot names its source is its own ancestor map
        p = self.poses()
        self.set_poses(p[idx])

    def roughen(self, var=None, noise=None):
        v = (C.c_double * 6)(*parse_covariance(self.res_noise_cov if var is None else var))
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32).reshape(self.n_rows, 6)
        with self._lock:
            _lib.check(self._lib.auvpf_roughen(self._h, v, _ptr(nz)))

    # -- estimate: auv_pf.update_loc_pose (auv_pf.py:218-252) --------------------------------
    def mean_cov(self):
        """``(mean (F,6), cov (F,3,3))``: unweighted mean pose with the arithmetic mean of
        the wrapped yaws, and the reference's partially filled 3x3 covariance."""
        mean = np.empty((self.n_filters, 6), np.float64)
        cov = np.empty((self.n_filters, 9), np.float64)
        with self._lock:
            _lib.check(self._lib.auvpf_mean_cov(self._h, _ptr(mean), _ptr(cov)))
        return mean, cov.reshape(self.n_filters, 3, 3)

    def estimate(self, weighted=False, circular_yaw=False, full_cov=False):
        """``mean_cov`` with options: weighted by the log-weights of the last update,
        circular yaw mean (the reference's arithmetic mean of wrapped yaws breaks at +-pi), full symmetric
        covariance.  All False = the reference's numbers."""
        flags = (_lib.EST_WEIGHTED if weighted else 0) | (_lib.EST_CIRCULAR_YAW if circular_yaw else 0) | \
            (_lib.EST_FULL_COV if full_cov else 0)
        mean = np.empty((self.n_filters, 6), np.float64)
        cov = np.empty((self.n_filters, 9), np.float64)
        with self._lock:
            _lib.check(self._lib.auvpf_estimate(self._h, flags, _ptr(mean), _ptr(cov)))
        return mean, cov.reshape(self.n_filters, 3, 3)

    def pose_array(self, max_poses=2000, stride=0):
        """The ``PoseArray`` of ``loc_loop`` (auv_pf.py:264-285) for a strided subsample, quaternions computed on
        the device: ``(F, n, 7)`` rows ``[x y z qx qy qz qw]``, ``n <= max_poses``."""
        n_max = min(int(max_poses), self.pc)
        out = np.empty((self.n_filters, n_max, 7), np.float64)
        n = C.c_int64(0)
        with self._lock:
            _lib.check(self._lib.auvpf_pose_array(self._h, int(stride), n_max, _ptr(out), C.byref(n)))
        return out.reshape(-1)[: self.n_filters * n.value * 7].reshape(self.n_filters, n.value, 7)

    def pose_array_topk(self, k=2000):
        """The ``k`` particles of largest weight per filter (between an update and its resample), selected on the
        device: ``(poses (F, k, 7) [x y z qx qy qz qw], log_weights (F, k), indices (F, k))``, rows ordered by
        weight descending, index ascending among equal weights."""
        k = min(int(k), self.pc)
        poses = np.empty((self.n_filters, k, 7), np.float64)
        lw = np.empty((self.n_filters, k), np.float32)
        idx = np.empty((self.n_filters, k), np.int32)
        with self._lock:
            _lib.check(self._lib.auvpf_pose_array_topk(self._h, k, _ptr(poses), _ptr(lw), _ptr(idx)))
        for f in range(self.n_filters):
            key = np.where(np.isnan(lw[f]), -np.inf, lw[f])
            order = np.lexsort((idx[f], -key))
            poses[f], lw[f], idx[f] = poses[f][order], lw[f][order], idx[f][order]
        return poses, lw, idx

    def neff(self):
        out = np.empty(self.n_filters, np.float64)
        with self._lock:
            _lib.check(self._lib.auvpf_neff(self._h, _ptr(out)))
        return out

    # -- N_eff-gated resampling: the block commented out at auv_pf.py:174-181 -------------------
    def set_neff_gate(self, ratio):
        """``resample()`` holds a filter's resample back (poses, weights and roughening untouched)
        whenever its ``N_eff = 1 / sum w^2`` is NOT below ``ratio * particle_count``; 0 switches the
        gate off.  The decision is taken on the device (``neff_fixed``: exact integer arithmetic, so
        it is reproducible and the CPU oracle takes the same one)."""
        with self._lock:
            _lib.check(self._lib.auvpf_set_neff_gate(self._h, float(ratio)))
        self.resample_neff_ratio = float(ratio)

    def neff_fixed(self):
        """The integer-exact N_eff the gate compares (per filter) of the last normalisation."""
        out = np.empty(self.n_filters, np.float64)
        with self._lock:
            _lib.check(self._lib.auvpf_neff_fixed(self._h, _ptr(out)))
        return out

    def resample_skipped(self):
        """Per filter: did the N_eff gate hold the last ``resample()`` back?  (synchronises)"""
        out = np.zeros(self.n_filters, np.int32)
        with self._lock:
            _lib.check(self._lib.auvpf_resample_skipped(self._h, _ptr(out)))
        return out.astype(bool)

    # -- replay of a recorded sequence -------------------------------------------------------
    def replay(self, odoms, ranges, u=None, roughen=True, estimates=True):
        """Run ``n_cycles`` whole filter cycles -- per cycle: the odometry messages between two
        pings (``predict_many``), the MBES ``update``, ``resample`` and the ``mean_cov`` estimate --
        without the host in the loop: the sequence is uploaded once, the estimates are read back once.
        ``odoms``: ``(n_cycles, n_steps, F or 1, 11)`` ABI rows (see ``predict``); ``ranges``:
        ``(n_cycles, F or 1, B)``; ``u``: ``(n_cycles, F)`` uniforms or None (Philox).  Returns
        ``(means (n_cycles, F, 6), covs (n_cycles, F, 3, 3))`` or None with ``estimates=False``
        (enqueue only).  Bit-identical to making the calls one by one."""
        if self.resampler != _lib.SYSTEMATIC or self.inplace_reference:
            raise ValueError("replay runs the enqueue-only cycle: systematic resampler, sorted gather")
        od = np.ascontiguousarray(odoms, np.float64)
        if od.ndim != 4 or od.shape[3] != 11 or od.shape[2] not in (1, self.n_filters):
            raise ValueError("odoms: (n_cycles, n_steps, F or 1, 11)")
        n_cycles = od.shape[0]
        r = np.asarray(ranges, np.float32).reshape(n_cycles, -1, self.n_beams)
        r = np.ascontiguousarray(np.broadcast_to(r, (n_cycles, self.n_filters, self.n_beams)))
        uu = None if u is None else np.ascontiguousarray(
            np.broadcast_to(np.asarray(u, np.float64).reshape(n_cycles, -1), (n_cycles, self.n_filters)))
        rv = (C.c_double * 6)(*self.res_noise_cov) if roughen else None
        mean = np.empty((n_cycles, self.n_filters, 6), np.float64) if estimates else None
        cov = np.empty((n_cycles, self.n_filters, 9), np.float64) if estimates else None
        with self._lock:
            _lib.check(self._lib.auvpf_replay(self._h, n_cycles, od.shape[1], od.shape[2], _ptr(od), _ptr(r),
                                              self.resampler, _lib.RS_SORTED, _ptr(uu), rv, _ptr(mean), _ptr(cov)))
        if not estimates:
            return None
        return mean, cov.reshape(n_cycles, self.n_filters, 3, 3)

    # -- measurement support -----------------------------------------------------------------
    def profile(self, on=True):
        _lib.check(self._lib.auvpf_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset=True):
        n = len(_lib.K_NAMES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        with self._lock:
            _lib.check(self._lib.auvpf_profile_read(self._h, ms, cnt, 1 if reset else 0))
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(_lib.K_NAMES)}
